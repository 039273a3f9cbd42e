"""Adapter giving libjalla_b200.so the same numpy-facing call surface as
oracle.cpu_ref.Backend, so a parity test is `gpu.f(args)` vs `oracle.f(args)` on identical
arguments.  mode="device": operands are copied into jb_malloc'd device buffers and device
pointers cross the C ABI; mode="host": numpy pointers cross the ABI and the library stages."""
from __future__ import annotations

import ctypes as C

import numpy as np

from jalla_b200 import _lib
from jalla_b200._lib import lib, check, check_last

NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


class DevBuf:
    def __init__(self, arr: np.ndarray):
        self.arr = arr
        self.nbytes = max(arr.nbytes, 16)
        self.ptr = lib.jb_malloc(self.nbytes, _lib.JB_MEM_DEVICE)
        assert self.ptr, lib.jb_last_error_string()
        if arr.nbytes:
            check(lib.jb_memcpy_h2d(self.ptr, arr.ctypes.data, arr.nbytes), "h2d")

    def back(self):
        if self.arr.nbytes:
            check(lib.jb_memcpy_d2h(self.arr.ctypes.data, self.ptr, self.arr.nbytes), "d2h")

    def free(self):
        lib.jb_free(self.ptr)


class GpuBackend:
    label = "jalla_b200 (B200)"

    def __init__(self, mode="device", ipiv_on_device=False):
        self.mode, self.ipiv_dev = mode, ipiv_on_device

    def _run(self, fn, args, ins, outs, what, returns_info=True):
        """args: list with numpy arrays in operand positions.  ins/outs: indices of arrays."""
        bufs = {}
        call = list(args)
        for i, a in enumerate(args):
            if isinstance(a, np.ndarray):
                is_piv = a.dtype == np.int32
                if self.mode == "device" and (not is_piv or self.ipiv_dev):
                    bufs[i] = DevBuf(a)
                    call[i] = bufs[i].ptr
                else:
                    call[i] = a.ctypes.data
        r = fn(*call)
        if returns_info:
            check(r, what)
        else:
            check_last(what)
        check(lib.jb_sync(), "sync")
        for i, b in bufs.items():
            if i in outs:
                b.back()
            b.free()
        return r

    def gemm(self, t, order, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        fn = getattr(lib, f"cblas_{t}gemm")
        if t in "sd":
            al, be, keep = alpha, beta, None
        else:
            keep = np.array([alpha, beta], dtype=NP[t])
            al, be = keep.ctypes.data, keep.ctypes.data + keep.itemsize
        self._run(fn, [order, ta, tb, m, n, k, al, A, lda, B, ldb, be, Cm, ldc], {7, 9, 12}, {12}, "gemm", returns_info=False)

    def trsm(self, t, order, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb):
        fn = getattr(lib, f"cblas_{t}trsm")
        if t in "sd":
            al, keep = alpha, None
        else:
            keep = np.array([alpha], dtype=NP[t]); al = keep.ctypes.data
        self._run(fn, [order, side, uplo, trans, diag, m, n, al, A, lda, B, ldb], {8, 10}, {10}, "trsm", returns_info=False)

    def syrk(self, t, order, uplo, trans, n, k, alpha, A, lda, beta, Cm, ldc, herm=False):
        fn = getattr(lib, f"cblas_{t}{'herk' if herm else 'syrk'}")
        if herm or t in "sd":
            al, be, keep = float(alpha), float(beta), None
        else:
            keep = np.array([alpha, beta], dtype=NP[t]); al, be = keep.ctypes.data, keep.ctypes.data + keep.itemsize
        self._run(fn, [order, uplo, trans, n, k, al, A, lda, be, Cm, ldc], {6, 9}, {9}, "syrk", returns_info=False)

    def gesv(self, t, order, n, nrhs, a, lda, ipiv, b, ldb):
        return self._run(getattr(lib, f"LAPACKE_{t}gesv"), [order, n, nrhs, a, lda, ipiv, b, ldb], {3, 6}, {3, 5, 6}, "gesv")

    def getrf(self, t, order, m, n, a, lda, ipiv):
        return self._run(getattr(lib, f"LAPACKE_{t}getrf"), [order, m, n, a, lda, ipiv], {3}, {3, 5}, "getrf")

    def getrs(self, t, order, trans, n, nrhs, a, lda, ipiv, b, ldb):
        return self._run(getattr(lib, f"LAPACKE_{t}getrs"), [order, trans.encode(), n, nrhs, a, lda, ipiv, b, ldb], {4, 6, 7}, {7}, "getrs")

    def getri(self, t, order, n, a, lda, ipiv):
        return self._run(getattr(lib, f"LAPACKE_{t}getri"), [order, n, a, lda, ipiv], {2, 4}, {2}, "getri")

    def potrf(self, t, order, uplo, n, a, lda):
        return self._run(getattr(lib, f"LAPACKE_{t}potrf"), [order, uplo.encode(), n, a, lda], {3}, {3}, "potrf")

    def potrs(self, t, order, uplo, n, nrhs, a, lda, b, ldb):
        return self._run(getattr(lib, f"LAPACKE_{t}potrs"), [order, uplo.encode(), n, nrhs, a, lda, b, ldb], {4, 6}, {6}, "potrs")

    def posv(self, t, order, uplo, n, nrhs, a, lda, b, ldb):
        return self._run(getattr(lib, f"LAPACKE_{t}posv"), [order, uplo.encode(), n, nrhs, a, lda, b, ldb], {4, 6}, {4, 6}, "posv")
