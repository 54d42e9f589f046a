"""Drop-in ``qarray`` import surface for the names the reference uses
(src/qadapt/environment/qarray_base_class.py:12): ``ChargeSensedDotArray``, ``LatchingModel``, ``TelegraphNoise``,
``WhiteNoise`` -- backed by the B200-native libqdsim.so."""
from qdsim.composer import GateVoltageComposer  # noqa: F401

from .charge_sensed import ChargeSensedDotArray  # noqa: F401
from .latching_models import LatchingBaseModel, LatchingModel  # noqa: F401
from .noise_models import BaseNoiseModel, NoiseModelSum, TelegraphNoise, WhiteNoise  # noqa: F401

__version__ = "1.6.0+qdsim"
