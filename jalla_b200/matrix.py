"""Host-side mirror of Jalla's matrix interface for a GPU-resident matrix type.

Names, argument meaning and error behaviour follow /root/reference/Numeric/Jalla/Matrix.hs
so that the parity tests read like the reference's own (tests/Test.hs).  The reference is
Haskell and no GHC exists in this image, so this mirror is Python above the C ABI
(include/jalla_b200.h); haskell/Numeric/Jalla/B200.hs holds the uncompiled Haskell
instance a maintainer would add.

Every arithmetic call goes through the same C symbols Jalla binds (cblas_?gemm,
LAPACKE_?gesv, LAPACKE_?getrf, LAPACKE_?getri) — here resolved in libjalla_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import lib, check, check_last

# Numeric/Jalla/Types.hs:78-89 (Transpose, Order); BLAS enum values Matrix.hs:154-175
NoTrans, Trans = _lib.NO_TRANS, _lib.TRANS
RowMajor, ColumnMajor = _lib.ROW_MAJOR, _lib.COL_MAJOR

_NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
_LETTER = {np.dtype(v): k for k, v in _NP.items()}


def _letter(dtype) -> str:
    if isinstance(dtype, str) and dtype in _NP:
        return dtype
    return _LETTER[np.dtype(dtype)]


class Matrix:
    """GPU-resident counterpart of `data Matrix e` (Matrix.hs:329-332): device buffer,
    shape, leading dimension, order.  Always ColumnMajor with lda = rows, like
    matrixAlloc' (Matrix.hs:785-787); the buffer is uninitialised on allocation."""

    __slots__ = ("ptr", "shape", "lda", "order", "t", "_own")

    def __init__(self, shape: Tuple[int, int], dtype="d", ptr: Optional[int] = None, lda: Optional[int] = None):
        r, c = int(shape[0]), int(shape[1])
        self.shape = (r, c)
        self.t = _letter(dtype)
        self.lda = int(lda) if lda is not None else max(r, 1)
        self.order = ColumnMajor
        if ptr is None:
            nbytes = self.lda * max(c, 1) * self.itemsize
            p = lib.jb_malloc(nbytes, _lib.JB_MEM_DEVICE)
            if not p:
                check(lib.jb_last_error(), "matrixAlloc")
                raise MemoryError("matrixAlloc: device allocation failed")
            self.ptr, self._own = p, True
        else:
            self.ptr, self._own = int(ptr), False

    def __del__(self):  # the ForeignPtr finalizer of the Haskell instance: jb_free
        if getattr(self, "_own", False) and self.ptr:
            lib.jb_free(self.ptr)
            self.ptr = 0

    @property
    def dtype(self):
        return np.dtype(_NP[self.t])

    @property
    def itemsize(self) -> int:
        return self.dtype.itemsize

    # --- listMatrix / matrixList analogues: bulk host<->device traffic
    @staticmethod
    def from_numpy(a: np.ndarray) -> "Matrix":
        a = np.asfortranarray(a)
        if a.ndim != 2:
            raise ValueError("Matrix.from_numpy: need a 2-D array")
        m = Matrix(a.shape, a.dtype)
        if a.size:
            check(lib.jb_memcpy_h2d(m.ptr, a.ctypes.data, a.nbytes), "listMatrix")
        return m

    def to_numpy(self) -> np.ndarray:
        r, c = self.shape
        buf = np.empty((self.lda, c), dtype=self.dtype, order="F")
        if buf.size:
            check(lib.jb_memcpy_d2h(buf.ctypes.data, self.ptr, buf.nbytes), "matrixList")
        return np.asfortranarray(buf[:r, :])

    def __matmul__(self, other: "Matrix") -> "Matrix":  # (##), Matrix.hs:198-199
        return mm(self, other)


def rowCount(m: Matrix) -> int:
    return m.shape[0]


def colCount(m: Matrix) -> int:
    return m.shape[1]


def matrixAlloc(shape, dtype="d") -> Matrix:
    return Matrix(shape, dtype)


def _shape_trans(t: int, shape):  # Types.hs:50-62 shapeTrans
    return shape if t == NoTrans else (shape[1], shape[0])


def _scalar_args(t: str, alpha, beta):
    if t == "s":
        return C.c_float(alpha), C.c_float(beta), None
    if t == "d":
        return C.c_double(alpha), C.c_double(beta), None
    # complex scalars travel by pointer to call-scoped temporaries (BlasOps.hs:156,180)
    arr = np.array([alpha, beta], dtype=_NP[t])
    return C.c_void_p(arr.ctypes.data), C.c_void_p(arr.ctypes.data + arr.itemsize), arr


def unsafeMatrixMult(alpha, transA: int, a: Matrix, transB: int, b: Matrix, beta, c: Matrix) -> None:
    """C <- alpha*op(A)*op(B) + beta*C in place (Matrix.hs:467-490).  Like the reference the
    size of C is not checked."""
    if a.order != b.order:
        raise ValueError("unsafeMatrixMult: order of matrices must be equal.")
    if not (a.t == b.t == c.t):
        raise TypeError("unsafeMatrixMult: element types differ")
    m, k = _shape_trans(transA, a.shape)
    n = _shape_trans(transB, b.shape)[1]
    al, be, keep = _scalar_args(a.t, alpha, beta)
    fn = getattr(lib, f"cblas_{a.t}gemm")
    fn(a.order, transA, transB, m, n, k, al, a.ptr, a.lda, b.ptr, b.lda, be, c.ptr, c.lda)
    del keep
    check_last("gemm")


def matrixMult(alpha, transA: int, a: Matrix, transB: int, b: Matrix) -> Matrix:
    """alpha * op(A) * op(B) into a fresh, uninitialised matrix with beta = 0 (Matrix.hs:446-461)."""
    shape = (_shape_trans(transA, a.shape)[0], _shape_trans(transB, b.shape)[1])
    ret = matrixAlloc(shape, a.t)
    unsafeMatrixMult(alpha, transA, a, transB, b, 0, ret)
    return ret


def mm(a: Matrix, b: Matrix) -> Matrix:
    """`a ## b` (Matrix.hs:198-199)."""
    if colCount(a) != rowCount(b):
        raise ValueError("(##): shapes do not match")
    return matrixMult(1, NoTrans, a, NoTrans, b)


def mmT(at: Tuple[Matrix, int], bt: Tuple[Matrix, int]) -> Matrix:
    """`(a,ta) ##! (b,tb)` (Matrix.hs:200-203)."""
    (a, ta), (b, tb) = at, bt
    if _shape_trans(ta, a.shape)[1] != _shape_trans(tb, b.shape)[0]:
        raise ValueError("(##!): shapes do not match")
    return matrixMult(1, ta, a, tb, b)


def matrixCopy(a: Matrix, trans: int = NoTrans) -> Matrix:
    """matrixCopy (Matrix.hs:877-878): one device lacpy / transpose-copy instead of the
    reference's per-row cblas_?copy loop (Matrix.hs:855-871)."""
    ret = matrixAlloc(_shape_trans(trans, a.shape), a.t)
    check(getattr(lib, f"jb_{a.t}lacpy")(trans, a.shape[0], a.shape[1], a.ptr, a.lda, ret.ptr, ret.lda), "matrixCopy")
    return ret


def referenceMatrixCopy(a: Matrix, trans: int = NoTrans) -> Matrix:
    """matrixCopy exactly as the unmodified reference performs it on a CMatrix (Matrix.hs:855-878): one strided
    cblas_?copy per row of the destination (rowsRef / columnsRef vectors, CVector.hs:284-288) — here on device pointers,
    which libjalla_b200's level-1 entry points accept.  r kernel launches; matrixCopy above is the one-launch form."""
    r, c = a.shape
    ret = matrixAlloc(_shape_trans(trans, a.shape), a.t)
    copy = getattr(lib, f"cblas_{a.t}copy")
    es = a.itemsize
    if trans == NoTrans:
        for i in range(r):      # row i of A (stride lda) -> row i of the copy
            copy(c, a.ptr + es * i, a.lda, ret.ptr + es * i, ret.lda)
    else:
        for j in range(c):      # column j of A (stride 1) -> row j of the copy
            copy(r, a.ptr + es * a.lda * j, 1, ret.ptr + es * j, ret.lda)
    check_last("matrixCopy")
    return ret


def _scalar_ptrs(t: str, *vals):
    arr = np.array(vals, dtype=_NP[t])
    return [C.c_void_p(arr.ctypes.data + i * arr.itemsize) for i in range(len(vals))], arr


def fillMatrix(shape, value, dtype="d") -> Matrix:
    """createMatrix shape (fill value) on the device (Matrix.hs:798-819,1050-1060 do it with per-element pokes)."""
    t = _letter(dtype)
    m = matrixAlloc(shape, t)
    if t in "sd":
        check(getattr(lib, f"jb_{t}laset")(shape[0], shape[1], value, value, m.ptr, m.lda), "fill")
    else:
        (p,), keep = _scalar_ptrs(t, value)
        check(getattr(lib, f"jb_{t}laset")(shape[0], shape[1], p, p, m.ptr, m.lda), "fill")
    return m


def idMatrix(n: int, dtype="d") -> Matrix:
    """idMatrix n = createMatrix (n,n) $ fill 0 >> setDiag 0 (repeat 1)  (Matrix.hs:525-526), one device kernel."""
    t = _letter(dtype)
    m = matrixAlloc((n, n), t)
    if t in "sd":
        check(getattr(lib, f"jb_{t}laset")(n, n, 0.0, 1.0, m.ptr, m.lda), "idMatrix")
    else:
        (p0, p1), keep = _scalar_ptrs(t, 0, 1)
        check(getattr(lib, f"jb_{t}laset")(n, n, p0, p1, m.ptr, m.lda), "idMatrix")
    return m


def setDiag(m: Matrix, d: int, values) -> None:
    """setDiag d values (Matrix.hs:1023-1036): only as many values as fit in the diagonal are used."""
    v = np.ascontiguousarray(values, dtype=m.dtype)
    check(getattr(lib, f"jb_{m.t}set_diag")(m.shape[0], m.shape[1], d, v.ctypes.data, v.size, m.ptr, m.lda), "setDiag")


def matrixMultDiag(at: Tuple[Matrix, int], d) -> Matrix:
    """matrixMultDiag (a,t) d = op(A) * diag(d)  (Matrix.hs:660-665: modifyMatrix a t $ zipWithM_ scaleColumn d [0..c-1]),
    one fused transpose-copy-and-scale kernel instead of a copy plus c cblas_?scal calls."""
    a, t = at
    r, c = _shape_trans(t, a.shape)
    v = np.ascontiguousarray(list(d)[:c] if not isinstance(d, np.ndarray) else d[:c], dtype=a.dtype)
    if v.size < c:      # zipWithM_ stops at the shorter list: the remaining columns are copied unscaled
        v = np.concatenate([v, np.ones(c - v.size, dtype=a.dtype)])
    out = matrixAlloc((r, c), a.t)
    check(getattr(lib, f"jb_{a.t}mult_diag")(t, r, c, a.ptr, a.lda, v.ctypes.data, out.ptr, out.lda), "matrixMultDiag")
    return out


def pseudoInverseFromSVD(u: Matrix, s, vt: Matrix) -> Matrix:
    """The GEMM tail of pseudoInverse (Matrix.hs:544-551): (matrixMultDiag (vt,Trans) s', NoTrans) ##! (u,Trans) with
    s' = 1/s where s /= 0.  The SVD itself (gesvd) is outside the hot path; u, s, vt come from the caller."""
    s = np.asarray(s, dtype=np.float64)
    sinv = np.where(s != 0, 1.0 / np.where(s != 0, s, 1.0), 0.0)
    return mmT((matrixMultDiag((vt, Trans), sinv.astype(vt.dtype)), NoTrans), (u, Trans))


def unsafeSolveLinearSystem(a: Matrix, b: Matrix) -> None:
    """Solve A X = B with ?gesv, overwriting A with its LU factors and B with X
    (Matrix.hs:495-508).  ipiv is a call-scoped host array like the reference's allocaArray."""
    if not (rowCount(a) == colCount(a) and rowCount(a) == rowCount(b)):
        raise ValueError("unsafeSolveLinearSystem: The shapes of the arguments do not match.")
    n, nrhs = colCount(a), colCount(b)
    ipiv = (C.c_int * max(n, 1))()
    ret = getattr(lib, f"LAPACKE_{a.t}gesv")(a.order, n, nrhs, a.ptr, a.lda, ipiv, b.ptr, b.lda)
    check(ret, "gesv")
    if ret != 0:
        raise ArithmeticError("unsafeSolveLinearSystem: ret /= 0")


def solveLinearSystem(a: Matrix, b: Matrix) -> Matrix:
    """solveLinearSystem (Matrix.hs:513-521): copies B and A, solves in place, returns X only."""
    x = matrixCopy(b, NoTrans)
    a2 = matrixCopy(a, NoTrans)
    unsafeSolveLinearSystem(a2, x)
    return x


def unsafeInvert(a: Matrix) -> Optional[Matrix]:
    """getrf then getri in place; None when either reports info /= 0 (Matrix.hs:763-776)."""
    m, n = a.shape
    ipiv = (C.c_int * max(min(m, n), 1))()
    ret = getattr(lib, f"LAPACKE_{a.t}getrf")(a.order, m, n, a.ptr, a.lda, ipiv)
    check(ret, "getrf")
    if ret != 0:
        return None
    ret = getattr(lib, f"LAPACKE_{a.t}getri")(a.order, n, a.ptr, a.lda, ipiv)
    check(ret, "getri")
    return a if ret == 0 else None


def invert(a: Matrix) -> Optional[Matrix]:
    """invert (Matrix.hs:537-540): None for non-square or singular input."""
    if colCount(a) != rowCount(a):
        return None
    return unsafeInvert(matrixCopy(a, NoTrans))


def invert_(a: Matrix) -> Matrix:
    """invert' (Matrix.hs:530-532): inverse by solving A X = I."""
    if colCount(a) != rowCount(a):
        raise ValueError("Cannot invert non-square matrix.")
    return solveLinearSystem(a, idMatrix(colCount(a), a.t))


def frobNorm(m: Matrix) -> float:
    """Host-side check helper (the reference computes it with BLAS dot, off the hot path)."""
    return float(np.linalg.norm(m.to_numpy()))
