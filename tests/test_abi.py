"""The C-ABI shared library: it is built, loads, and exports every symbol include/qdsim.h declares; the ctypes / numpy
mirrors of its structs have the header's sizes.  No compute calls (CPU only)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qdsim.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("qd_create", "qd_destroy", "qd_last_error", "qd_set_models", "qd_scan_open", "qd_scan_open_host",
                 "qd_points_open_host", "qd_scan_upload", "qd_scan_launch"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from qdsim import _lib
    lib = _lib.load()
    assert lib.qd_abi_version() == 3
    for name in _declared_functions():
        assert hasattr(lib, name), f"libqdsim.so does not export {name}"
    assert set(_lib.EXPORTS) == set(_declared_functions())


def test_struct_mirrors_match_the_header(tmp_path):
    from qdsim import PARAMS_DTYPE, SCAN_DTYPE
    from qdsim._lib import ModelDesc
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qdsim.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(qd_scan),sizeof(qd_env_params),sizeof(qd_model_desc),offsetof(qd_scan,seed),'
                    'offsetof(qd_scan,env_id),offsetof(qd_env_params,max_charge_carriers));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    s_scan, s_par, s_desc, o_seed, o_env, o_maxc = map(int, subprocess.check_output([str(exe)]).split())
    assert SCAN_DTYPE.itemsize == s_scan == 480
    assert PARAMS_DTYPE.itemsize == s_par
    assert ctypes.sizeof(ModelDesc) == s_desc
    assert SCAN_DTYPE.fields["seed"][1] == o_seed and SCAN_DTYPE.fields["env_id"][1] == o_env
    assert PARAMS_DTYPE.fields["max_charge_carriers"][1] == o_maxc


def test_no_cpu_fallback_without_the_library(monkeypatch, tmp_path):
    """The product path must fail loudly when the CUDA library is missing."""
    from qdsim import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("QDSIM_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(OSError):
        _lib.load()


def test_engine_raises_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qdsim import Engine, QdError
    with pytest.raises(QdError):
        Engine(0)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rl-agent-for-qubit-array-tuning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
