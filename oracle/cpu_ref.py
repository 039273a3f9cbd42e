"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Uniform ctypes front end over the CPU implementations of the C boundary Jalla calls
(cblas_?gemm, LAPACKE_?{gesv,getrf,getrs,getri,potrf,potrs,posv}):

  * `oracle()`   — this repo's plain-C restatement, oracle/libjalla_oracle.so
                   (oracle/jalla_oracle.c, built by `make -C oracle`);
  * `openblas()` — the third-party library Jalla itself would link
                   (/root/reference/jalla.cabal:75 `extra-libraries: lapacke, cblas, blas`):
                   OpenBLAS 0.3.31.dev bundled with scipy (LP64, `scipy_` symbol prefix).
                   It pins the restatement (tests/test_oracle_pin.py), generated the golden
                   vectors (tests/golden/make_golden.py) and is the timed CPU arm of bench.py;
  * `openblas_second()` — OpenBLAS 0.3.15 bundled with opencv (unprefixed symbols): a second
                   opinion only.

All matrices are numpy arrays; pointers and leading dimensions are passed exactly as
Jalla's FFI does (order 101/102, trans 111/112/113, ipiv int32 1-based, info returned).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
_i, _p, _c = C.c_int, C.c_void_p, C.c_char


def _ptr(a):
    return a.ctypes.data_as(_p) if isinstance(a, np.ndarray) else a


class Backend:
    def __init__(self, lib, gemm_name, lapacke_name, label, info=None):
        self.lib, self.label, self.info = lib, label, info or {}
        self._gemm_name, self._lapacke_name = gemm_name, lapacke_name
        self._cache = {}

    def _fn(self, key, name, res, args):
        f = self._cache.get(key)
        if f is None:
            f = getattr(self.lib, name)
            f.restype, f.argtypes = res, args
            self._cache[key] = f
        return f

    # cblas_?gemm(order, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
    def gemm(self, t, order, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        if t in "sd":
            sc = C.c_float if t == "s" else C.c_double
            f = self._fn(("gemm", t), self._gemm_name(t), None, [_i] * 6 + [sc, _p, _i, _p, _i, sc, _p, _i])
            f(order, ta, tb, m, n, k, alpha, _ptr(A), lda, _ptr(B), ldb, beta, _ptr(Cm), ldc)
        else:
            f = self._fn(("gemm", t), self._gemm_name(t), None, [_i] * 6 + [_p, _p, _i, _p, _i, _p, _p, _i])
            sc = np.array([alpha, beta], dtype=NP[t])
            f(order, ta, tb, m, n, k, sc.ctypes.data, _ptr(A), lda, _ptr(B), ldb, sc.ctypes.data + sc.itemsize, _ptr(Cm), ldc)

    # cblas_?trsm (OpenBLAS backends only; the C restatement has no trsm)
    def trsm(self, t, order, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb):
        name = self._gemm_name(t).replace("gemm", "trsm")
        if t in "sd":
            sc = C.c_float if t == "s" else C.c_double
            f = self._fn(("trsm", t), name, None, [_i] * 7 + [sc, _p, _i, _p, _i])
            f(order, side, uplo, trans, diag, m, n, alpha, _ptr(A), lda, _ptr(B), ldb)
        else:
            f = self._fn(("trsm", t), name, None, [_i] * 7 + [_p, _p, _i, _p, _i])
            sc = np.array([alpha], dtype=NP[t])
            f(order, side, uplo, trans, diag, m, n, sc.ctypes.data, _ptr(A), lda, _ptr(B), ldb)

    # cblas_?syrk / cblas_?herk (OpenBLAS backends only)
    def syrk(self, t, order, uplo, trans, n, k, alpha, A, lda, beta, Cm, ldc, herm=False):
        name = self._gemm_name(t).replace("gemm", "herk" if herm else "syrk")
        if herm or t in "sd":
            sc = C.c_float if t in "sc" else C.c_double
            f = self._fn(("syrk", t, herm), name, None, [_i] * 5 + [sc, _p, _i, sc, _p, _i])
            f(order, uplo, trans, n, k, alpha, _ptr(A), lda, beta, _ptr(Cm), ldc)
        else:
            f = self._fn(("syrk", t, herm), name, None, [_i] * 5 + [_p, _p, _i, _p, _p, _i])
            sc = np.array([alpha, beta], dtype=NP[t])
            f(order, uplo, trans, n, k, sc.ctypes.data, _ptr(A), lda, sc.ctypes.data + sc.itemsize, _ptr(Cm), ldc)

    def gesv(self, t, order, n, nrhs, a, lda, ipiv, b, ldb):
        f = self._fn(("gesv", t), self._lapacke_name(t, "gesv"), _i, [_i, _i, _i, _p, _i, _p, _p, _i])
        return f(order, n, nrhs, _ptr(a), lda, _ptr(ipiv), _ptr(b), ldb)

    def getrf(self, t, order, m, n, a, lda, ipiv):
        f = self._fn(("getrf", t), self._lapacke_name(t, "getrf"), _i, [_i, _i, _i, _p, _i, _p])
        return f(order, m, n, _ptr(a), lda, _ptr(ipiv))

    def getrs(self, t, order, trans, n, nrhs, a, lda, ipiv, b, ldb):
        f = self._fn(("getrs", t), self._lapacke_name(t, "getrs"), _i, [_i, _c, _i, _i, _p, _i, _p, _p, _i])
        return f(order, trans.encode(), n, nrhs, _ptr(a), lda, _ptr(ipiv), _ptr(b), ldb)

    def getri(self, t, order, n, a, lda, ipiv):
        f = self._fn(("getri", t), self._lapacke_name(t, "getri"), _i, [_i, _i, _p, _i, _p])
        return f(order, n, _ptr(a), lda, _ptr(ipiv))

    def potrf(self, t, order, uplo, n, a, lda):
        f = self._fn(("potrf", t), self._lapacke_name(t, "potrf"), _i, [_i, _c, _i, _p, _i])
        return f(order, uplo.encode(), n, _ptr(a), lda)

    def potrs(self, t, order, uplo, n, nrhs, a, lda, b, ldb):
        f = self._fn(("potrs", t), self._lapacke_name(t, "potrs"), _i, [_i, _c, _i, _i, _p, _i, _p, _i])
        return f(order, uplo.encode(), n, nrhs, _ptr(a), lda, _ptr(b), ldb)

    def posv(self, t, order, uplo, n, nrhs, a, lda, b, ldb):
        f = self._fn(("posv", t), self._lapacke_name(t, "posv"), _i, [_i, _c, _i, _i, _p, _i, _p, _i])
        return f(order, uplo.encode(), n, nrhs, _ptr(a), lda, _ptr(b), ldb)


_oracle = None
_openblas = None
_openblas2 = None


def oracle() -> Backend:
    global _oracle
    if _oracle is None:
        path = os.path.join(_HERE, "libjalla_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        _oracle = Backend(lib, lambda t: f"oracle_{t}gemm", lambda t, r: f"oracle_{t}LAPACKE_{r}", "oracle (C restatement)")
    return _oracle


def _site_packages():
    return [p for p in sys.path if p.endswith("site-packages")]


def openblas() -> Backend:
    """scipy-bundled OpenBLAS (the library Jalla's jalla.cabal:75 would resolve to here)."""
    global _openblas
    if _openblas is None:
        cands = []
        for sp in _site_packages():
            cands += sorted(glob.glob(os.path.join(sp, "scipy.libs", "libscipy_openblas-*.so")))
        if not cands:
            raise RuntimeError("scipy-bundled OpenBLAS not found")
        lib = C.CDLL(cands[0], mode=C.RTLD_LOCAL)
        info = {"path": cands[0]}
        try:
            lib.scipy_openblas_get_config.restype = C.c_char_p
            lib.scipy_openblas_get_corename.restype = C.c_char_p
            lib.scipy_openblas_get_num_threads.restype = C.c_int
            info.update(config=lib.scipy_openblas_get_config().decode(), core=lib.scipy_openblas_get_corename().decode(),
                        threads=lib.scipy_openblas_get_num_threads())
        except AttributeError:
            pass
        _openblas = Backend(lib, lambda t: f"scipy_cblas_{t}gemm", lambda t, r: f"scipy_LAPACKE_{t}{r}",
                            "OpenBLAS (scipy-bundled)", info)
    return _openblas


def openblas_set_threads(n: int) -> None:
    lib = openblas().lib
    lib.scipy_openblas_set_num_threads.argtypes = [C.c_int]
    lib.scipy_openblas_set_num_threads(int(n))
    openblas().info["threads"] = int(n)


def openblas_second() -> Backend:
    global _openblas2
    if _openblas2 is None:
        for sp in _site_packages():
            d = os.path.join(sp, "opencv_python_headless.libs")
            libs = sorted(glob.glob(os.path.join(d, "libopenblasp-*.so")))
            if libs:
                for dep in sorted(glob.glob(os.path.join(d, "libquadmath-*.so*"))) + sorted(glob.glob(os.path.join(d, "libgfortran-*.so*"))):
                    C.CDLL(dep, mode=C.RTLD_GLOBAL)
                lib = C.CDLL(libs[0], mode=C.RTLD_LOCAL)
                _openblas2 = Backend(lib, lambda t: f"cblas_{t}gemm", lambda t, r: f"LAPACKE_{t}{r}",
                                     "OpenBLAS 0.3.15 (opencv-bundled)", {"path": libs[0]})
                break
        if _openblas2 is None:
            raise RuntimeError("opencv-bundled OpenBLAS not found")
    return _openblas2
