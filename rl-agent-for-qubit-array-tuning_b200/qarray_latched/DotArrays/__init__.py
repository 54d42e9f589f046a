from .barrier_voltage_model import BarrierVoltageModel  # noqa: F401
from .GateVoltageComposer import GateVoltageComposer  # noqa: F401
from .TunnelCoupledChargeSensed import TunnelCoupledChargeSensed  # noqa: F401
from .voltage_dependent_capacitance import (VoltageDependendentCapacitanceModel,  # noqa: F401
                                            create_linear_capacitance_model, create_quadratic_capacitance_model,
                                            create_sigmoid_capacitance_model)
