"""The C-ABI boundary: libjalla_b200.so loads, exports every symbol include/jalla_b200.h declares
(including the unprefixed cblas_* / LAPACKE_* names Jalla's `foreign import ccall` resolves), the product
never links the oracle, and without a GPU every entry point fails loudly instead of computing on the CPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jalla_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(?:int|void|void \*|const char \*|uint64_t)\s*\*?\s*((?:jb_|cblas_|LAPACKE_)\w+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_jalla_boundary():
    names = declared_symbols()
    for t in "sdcz":
        assert f"cblas_{t}gemm" in names and f"jb_cblas_{t}gemm" in names
        for r in ("gesv", "getrf", "getrs", "getri", "potrf", "potrs", "posv"):
            assert f"LAPACKE_{t}{r}" in names and f"jb_LAPACKE_{t}{r}" in names
        for r in ("getrf_batched", "getrs_batched", "gesv_batched", "potrf_batched", "potrs_batched"):
            assert f"jb_{t}{r}" in names
    assert len(names) >= 105


def test_library_exports_every_declared_symbol():
    from jalla_b200 import _lib
    missing = [n for n in declared_symbols() if not hasattr(_lib.lib, n)]
    assert not missing, missing
    assert b"sm_100a" in _lib.lib.jb_version()


def test_header_compiles_as_c_and_cxx(tmp_path):
    for comp, lang in (("gcc", "c"), ("g++", "c++")):
        r = subprocess.run([comp, "-fsyntax-only", "-Wall", "-x", lang, HEADER], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_product_does_not_link_the_oracle():
    from jalla_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "openblas" not in out.lower() and "cublas" not in out.lower() and "cusolver" not in out.lower()
    syms = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms and "cublas" not in syms.lower() and "cusolver" not in syms.lower()
    for root, _, files in os.walk(os.path.join(ROOT, "jalla_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(root, f)).read()
                assert "from oracle" not in text and "import oracle" not in text, f


def test_no_gpu_fails_loudly():
    from jalla_b200 import _lib
    lib = _lib.lib
    if lib.jb_device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU-only container")
    assert lib.jb_init(-1) == _lib.JB_ERR_CUDA
    a = np.ones((4, 4), order="F"); c = np.full((4, 4), 7.0, order="F")
    lib.jb_last_error()
    lib.cblas_dgemm(102, 111, 111, 4, 4, 4, 1.0, a.ctypes.data, 4, a.ctypes.data, 4, 0.0, c.ctypes.data, 4)
    assert lib.jb_last_error() == _lib.JB_ERR_CUDA          # recorded, and C untouched: no CPU arithmetic path
    assert (c == 7.0).all()
    piv = np.zeros(4, np.int32)
    assert lib.LAPACKE_dgesv(102, 4, 1, a.ctypes.data, 4, piv.ctypes.data, c.ctypes.data, 4) == _lib.JB_ERR_CUDA
    assert lib.jb_dgetrf_batched(4, a.ctypes.data, 4, 16, piv.ctypes.data, piv.ctypes.data, 1) == _lib.JB_ERR_CUDA
    with pytest.raises(_lib.JallaB200Error):
        from jalla_b200 import matrix as J
        J.Matrix((4, 4), "d")


def test_argument_errors_need_no_gpu():
    """LAPACKE-style argument screening happens before any device work."""
    from jalla_b200 import _lib
    lib = _lib.lib
    a = np.ones((4, 4), order="F"); piv = np.zeros(4, np.int32)
    assert lib.LAPACKE_dgetrf(100, 4, 4, a.ctypes.data, 4, piv.ctypes.data) == -1
    assert lib.LAPACKE_dgetrf(102, -1, 4, a.ctypes.data, 4, piv.ctypes.data) == -2
    assert lib.LAPACKE_dgetrf(102, 4, 4, a.ctypes.data, 3, piv.ctypes.data) == -5
    assert lib.LAPACKE_dpotrf(102, b"X", 4, a.ctypes.data, 4) == -2
    assert lib.LAPACKE_dgetrs(102, b"Q", 4, 1, a.ctypes.data, 4, piv.ctypes.data, a.ctypes.data, 4) == -2
    assert lib.LAPACKE_dgesv(102, 4, 1, a.ctypes.data, 4, piv.ctypes.data, a.ctypes.data, 3) == -8
    lib.jb_last_error()
    lib.cblas_dgemm(102, 111, 111, 4, 4, 4, 1.0, a.ctypes.data, 3, a.ctypes.data, 4, 0.0, a.ctypes.data, 4)
    assert lib.jb_last_error() == _lib.JB_ERR_ARG


def test_synth_twin_is_deterministic_and_shardable():
    from jalla_b200 import synth
    full = synth.uniform(40, 30, 99, "d")
    part = synth.uniform(10, 7, 99, "d", row0=13, col0=5)
    np.testing.assert_array_equal(part, full[13:23, 5:12])
    assert -1.0 <= full.min() and full.max() < 1.0 and abs(full.mean()) < 0.1
    s = synth.uniform(20, 20, 5, "d", sym=True, diag=20.0)
    np.testing.assert_array_equal(s, s.T)
    assert np.all(np.linalg.eigvalsh(s) > 0)
    z = synth.uniform(8, 8, 3, "z")
    assert z.dtype == np.complex128 and np.abs(z.imag).max() > 0
