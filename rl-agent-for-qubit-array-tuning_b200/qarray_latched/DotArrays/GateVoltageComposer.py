from qdsim.composer import GateVoltageComposer  # noqa: F401
