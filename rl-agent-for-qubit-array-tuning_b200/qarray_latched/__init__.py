"""Drop-in ``qarray_latched`` import surface (the reference's vendored fork, src/qarray_latched/) for the names
src/qadapt/environment/qarray_base_class.py:13-17 imports -- backed by libqdsim.so."""
from .DotArrays import BarrierVoltageModel, GateVoltageComposer, TunnelCoupledChargeSensed  # noqa: F401
