"""Process-wide engines for the single-device drop-in classes: one ``Engine`` (qd_ctx) per CUDA device, shared by every
``ChargeSensedDotArray`` / ``TunnelCoupledChargeSensed`` of the process.  The model object that used an engine last is
its "owner"; another object re-uploads its own constants first (a few KB)."""
from __future__ import annotations

import itertools
import os

import numpy as np

from .engine import Engine

_engines: dict[int, Engine] = {}
_owner: dict[int, tuple] = {}
_tokens = itertools.count(1)


def default_device() -> int:
    return int(os.environ.get("QDSIM_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def engine_for(model, device: int | None = None) -> Engine:
    """Engine of ``device`` with ``model``'s constants resident (``model._model_batch()`` builds them)."""
    dev = default_device() if device is None else device
    eng = _engines.get(dev)
    if eng is None:
        eng = _engines[dev] = Engine(dev)
    token = model.__dict__.get("_qd_token")
    if token is None:                    # never id(model): the address of a collected model is reused by the next one
        token = next(_tokens)
        object.__setattr__(model, "_qd_token", token)
    key = (token, model._version)
    if _owner.get(dev) != key:
        eng.set_models(model._model_batch())
        _owner[dev] = key
    return eng


def fresh_seed() -> int:
    """A 64-bit scan seed drawn from the process-global ``np.random`` state -- the stream the reference's noise and
    latching classes consume -- so ``np.random.seed(...)`` makes a run reproducible."""
    return int.from_bytes(np.random.bytes(8), "little")


def shutdown():
    for eng in _engines.values():
        eng.close()
    _engines.clear()
    _owner.clear()
