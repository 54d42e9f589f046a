"""CPU oracle: a plain NumPy fp64 restatement of the reference's charge-stability hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``rl-agent-for-qubit-array-tuning_b200/``) never imports ``oracle`` and fails loudly when the CUDA
library is missing.

PARITY STATUS, by stage (details: DESIGN.md section 2).

* PINNED against the reference itself, run in this container (``tests/golden/ref_*.npz`` made by
  ``tests/golden/make_reference_golden.py``; checked by ``tests/test_reference_golden.py``): Maxwell conversion, scan
  grids, optimal gate voltages / virtual gate matrix, the whole tunnel-coupled "Path B" ground state
  (``src/qarray_latched/DotArrays/*.py`` executed unmodified) and the sensor stack; for "Path A" the free energy and
  floor/ceil candidate enumeration through the reference's in-tree mirrors (``src/qarray_latched/functions.py:30-47``)
  with the relaxation QP of ``functions.py:66-81`` solved exactly.
* PARITY UNPINNED for what only exists in the absent wheel: the ADMM/OSQP solver tolerance of the relaxation, the
  thresholded / brute-force variants, ``LatchingModel``, ``WhiteNoise``, ``TelegraphNoise``.

The arithmetic of "Path A" lives in the third-party package ``qarray==1.6.0`` (+
``qarray-rust-core==1.3.1``), pinned by the reference at ``pyproject.toml:8`` / ``uv.lock:2015-2053`` but absent
from ``/root/reference`` and from this image; the reference ships no test that pins a number on this path
(SURVEY.md section 4).  The oracle therefore restates

* literally, from in-tree source, everything that *is* in ``/root/reference`` (Maxwell conversion, scan grids,
  the sensor stack, the whole tunnel-coupled "Path B": ``src/qarray_latched/DotArrays/*.py``), and
* from the published algorithm of qarray 1.6.0 the pieces that are not (continuous relaxation + floor/ceil /
  thresholded / brute-force search, LatchingModel, WhiteNoise, TelegraphNoise), anchored on the reference's call
  sites (``src/qadapt/environment/qarray_base_class.py:726-756``) and on the in-tree mirrors of the upstream code
  (``src/qarray_latched/functions.py:30-81``, ``src/qarray_latched/latched.py:65-170``).

Every recalled (not read) semantic is an explicit keyword switch, listed in DESIGN.md "Oracle switches".

Randomness: every draw comes from a counter-based Philox4x32-10 stream keyed by a per-scan 64-bit seed with
counter = flattened pixel index (``oracle.philox``), so the CUDA kernels can be compared draw-for-draw.
"""

from . import philox, capacitance, composer, path_a, latching, noise, sensor, scan  # noqa: F401
