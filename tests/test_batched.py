"""Batched small-system parity (BASELINE.json configs[2]: independent 32x32 Double systems,
getrf + getrs).  The device kernels follow the oracle's arithmetic contract exactly
(oracle/oracle_impl.inc: r = 1/pivot, l = a*r, a_ij = fma(-l,u,a_ij)), so LU factors,
pivots and solutions are compared BIT-EXACTLY for real types — on every system, including
exact ties — and the full-size run is checked through size-independent properties."""
import ctypes as C

import numpy as np
import pytest

from jalla_b200 import _lib, synth
from jalla_b200._lib import lib, check
from tests.util import NP, EPS, fro

pytestmark = pytest.mark.gpu


class Dev:
    def __init__(self, arr):
        self.arr = np.ascontiguousarray(arr) if arr.ndim == 1 else arr
        self.ptr = lib.jb_malloc(max(self.arr.nbytes, 16), _lib.JB_MEM_DEVICE)
        assert self.ptr
        if self.arr.nbytes:
            check(lib.jb_memcpy_h2d(self.ptr, self.arr.ctypes.data, self.arr.nbytes), "h2d")

    def get(self):
        check(lib.jb_memcpy_d2h(self.arr.ctypes.data, self.ptr, self.arr.nbytes), "d2h")
        return self.arr

    def __del__(self):
        lib.jb_free(self.ptr)


def assert_same_bits(x, y):
    """Bit-identical except that any NaN matches any NaN (x86 and the GPU propagate different NaN payloads)."""
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    nx, ny = np.isnan(x), np.isnan(y)
    np.testing.assert_array_equal(nx, ny)
    np.testing.assert_array_equal(np.where(nx, 0, x.view(np.int64)), np.where(ny, 0, y.view(np.int64)))


def batch_rand(rng, t, batch, n, cols=None, lo=-1.0, hi=1.0):
    cols = n if cols is None else cols
    a = rng.uniform(lo, hi, (batch, cols, n))
    if t in "cz":
        a = a + 1j * rng.uniform(lo, hi, (batch, cols, n))
    return a.astype(NP[t])   # [batch][col][row]: column-major matrices, contiguous stride n*cols


def oracle_gesv_batch(orc, t, A, b):
    batch, n = A.shape[0], A.shape[2]
    LU, X = A.copy(), b.copy()
    piv = np.zeros((batch, n), np.int32); info = np.zeros(batch, np.int32)
    for i in range(batch):
        info[i] = orc.gesv(t, 102, n, b.shape[1], LU[i], n, piv[i], X[i], n)
    return LU, piv, X, info


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("n", [32, 1, 5, 17, 31])
def test_gesv_batched_bit_exact(orc, rng, t, n):
    lib.jb_init(-1)
    batch = 301
    A = batch_rand(rng, t, batch, n); b = batch_rand(rng, t, batch, n, 1)
    if t in "sd":   # exact ties: integer matrices exercise the lowest-index tie-break
        A[:40] = rng.integers(-2, 3, A[:40].shape).astype(NP[t])
    LUo, pivo, Xo, infoo = oracle_gesv_batch(orc, t, A, b)
    dA, dB = Dev(A.copy()), Dev(b.copy())
    dP, dI = Dev(np.zeros((batch, n), np.int32)), Dev(np.full(batch, -7, np.int32))
    rc = getattr(lib, f"jb_{t}gesv_batched")(n, 1, dA.ptr, n, n * n, dP.ptr, dB.ptr, n, n, dI.ptr, batch)
    check(rc, "gesv_batched"); check(lib.jb_sync(), "sync")
    LU, piv, X, info = dA.get(), dP.get(), dB.get(), dI.get()
    np.testing.assert_array_equal(info, infoo)
    np.testing.assert_array_equal(piv, pivo)
    ok = infoo == 0
    if t in "sd":
        np.testing.assert_array_equal(LU, LUo)
        np.testing.assert_array_equal(X[ok], Xo[ok])
    else:
        np.testing.assert_array_equal(LU.view(NP["s" if t == "c" else "d"]), LUo.view(NP["s" if t == "c" else "d"]))
        np.testing.assert_array_equal(X[ok], Xo[ok])
    np.testing.assert_array_equal(X[~ok], b[~ok])      # singular systems: B untouched, like ?gesv


@pytest.mark.parametrize("t", "sd")
def test_getrf_getrs_batched_multi_rhs(orc, rng, t):
    lib.jb_init(-1)
    batch, n, nrhs = 77, 32, 3
    A = batch_rand(rng, t, batch, n); B = batch_rand(rng, t, batch, n, nrhs)
    dA, dB = Dev(A.copy()), Dev(B.copy())
    dP, dI = Dev(np.zeros((batch, n), np.int32)), Dev(np.zeros(batch, np.int32))
    check(getattr(lib, f"jb_{t}getrf_batched")(n, dA.ptr, n, n * n, dP.ptr, dI.ptr, batch), "getrf_batched")
    check(getattr(lib, f"jb_{t}getrs_batched")(n, nrhs, dA.ptr, n, n * n, dP.ptr, dB.ptr, n, n * nrhs, batch), "getrs_batched")
    check(lib.jb_sync(), "sync")
    LUo, pivo, Xo, infoo = oracle_gesv_batch(orc, t, A, B)
    np.testing.assert_array_equal(dP.get(), pivo)
    np.testing.assert_array_equal(dA.get(), LUo)
    np.testing.assert_array_equal(dB.get(), Xo)
    # gesv_batched with nrhs > 1 routes through the same two kernels
    dA2, dB2 = Dev(A.copy()), Dev(B.copy())
    check(getattr(lib, f"jb_{t}gesv_batched")(n, nrhs, dA2.ptr, n, n * n, dP.ptr, dB2.ptr, n, n * nrhs, dI.ptr, batch), "gesv_batched")
    check(lib.jb_sync(), "sync")
    np.testing.assert_array_equal(dB2.get(), Xo)


def test_dsolve_batched_solve_only_contract(orc, rng):
    """The contract solveLinearSystem exposes (Matrix.hs:513-521): A and B are inputs, only X is produced."""
    lib.jb_init(-1)
    batch, n = 500, 32
    A = batch_rand(rng, "d", batch, n); b = batch_rand(rng, "d", batch, n, 1)
    dA, dB, dX, dI = Dev(A.copy()), Dev(b.copy()), Dev(np.zeros_like(b)), Dev(np.zeros(batch, np.int32))
    check(lib.jb_dsolve_batched(n, 1, dA.ptr, n, n * n, dB.ptr, n, n, dX.ptr, n, n, dI.ptr, batch), "dsolve_batched")
    check(lib.jb_sync(), "sync")
    _, _, Xo, infoo = oracle_gesv_batch(orc, "d", A, b)
    np.testing.assert_array_equal(dX.get(), Xo)
    np.testing.assert_array_equal(dA.get(), A)        # inputs untouched
    np.testing.assert_array_equal(dB.get(), b)


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n", [32, 9])
def test_potrf_potrs_batched(orc, rng, t, uplo, n):
    lib.jb_init(-1)
    batch, nrhs = 64, 2
    M = batch_rand(rng, t, batch, n)
    S = np.stack([(m.T @ m.T.conj().T + n * np.eye(n)).T for m in M]).astype(NP[t])   # [b][col][row]
    S[5, 3, 3] = -1.0                                    # one system that is not positive definite
    B = batch_rand(rng, t, batch, n, nrhs)
    dS, dB, dI = Dev(S.copy()), Dev(B.copy()), Dev(np.zeros(batch, np.int32))
    check(getattr(lib, f"jb_{t}potrf_batched")(uplo.encode(), n, dS.ptr, n, n * n, dI.ptr, batch), "potrf_batched")
    check(getattr(lib, f"jb_{t}potrs_batched")(uplo.encode(), n, nrhs, dS.ptr, n, n * n, dB.ptr, n, n * nrhs, batch), "potrs_batched")
    check(lib.jb_sync(), "sync")
    F, X, info = dS.get(), dB.get(), dI.get()
    for i in range(batch):
        So, Bo = S[i].copy(), B[i].copy()
        io = orc.potrf(t, 102, uplo, n, So, n)
        assert info[i] == io
        if io:
            continue
        orc.potrs(t, 102, uplo, n, nrhs, So, n, Bo, n)
        tri = np.triu_indices(n) if uplo == "L" else np.tril_indices(n)   # [col][row] storage: transposed view
        if t in "sd":
            np.testing.assert_array_equal(F[i][tri], So[tri])
        else:
            assert fro(F[i][tri] - So[tri]) <= 50 * n * EPS[t] * fro(So[tri])
        assert fro(X[i] - Bo) <= 1e3 * n * EPS[t] * fro(Bo)


@pytest.mark.parametrize("t", "ds")
@pytest.mark.parametrize("nrhs", [1, 3, 4, 5])
def test_chol32_tuned_kernels_bit_exact(orc, rng, nrhs, t):
    """The tuned Double / Float n = 32 lower-storage Cholesky kernels (csrc/batched_chol32.cu; nrhs <= 4, else the generic solve):
    factors bit for bit against the oracle's ?potf2, solutions bit for bit against the generic kernel (same operation
    sequence) and within rounding of the oracle's ?potrs, padded strides between systems,
    the strict upper triangle and the padding untouched, a non-positive pivot and a NaN reported through info."""
    lib.jb_init(-1)
    n, batch = 32, 201
    M = batch_rand(rng, t, batch, n)
    S = np.stack([(m.T @ m + n * np.eye(n)).T for m in M]).astype(NP[t])
    S[7, 11, 11] = -2.0
    S[9, 4, 4] = np.nan
    B = batch_rand(rng, t, batch, n, nrhs)
    sa = n * n + 32
    Sbuf = np.full((batch, sa), np.nan, NP[t]); Sbuf[:, :n * n] = S.reshape(batch, -1)
    dS, dB, dI = Dev(Sbuf.copy()), Dev(B.copy()), Dev(np.full(batch, -7, np.int32))
    check(getattr(lib, f"jb_{t}potrf_batched")(b"L", n, dS.ptr, n, sa, dI.ptr, batch), "potrf_batched")
    check(getattr(lib, f"jb_{t}potrs_batched")(b"L", n, nrhs, dS.ptr, n, sa, dB.ptr, n, n * nrhs, batch), "potrs_batched")
    check(lib.jb_sync(), "sync")
    F, X, info = dS.get(), dB.get(), dI.get()
    lib.jb_set_option(b"chol32.generic", 1)
    try:
        gS, gB, gI = Dev(Sbuf.copy()), Dev(B.copy()), Dev(np.full(batch, -7, np.int32))
        check(getattr(lib, f"jb_{t}potrf_batched")(b"L", n, gS.ptr, n, sa, gI.ptr, batch), "potrf_batched")
        check(getattr(lib, f"jb_{t}potrs_batched")(b"L", n, nrhs, gS.ptr, n, sa, gB.ptr, n, n * nrhs, batch), "potrs_batched")
        check(lib.jb_sync(), "sync")
    finally:
        lib.jb_set_option(b"chol32.generic", -1)
    ok = info == 0
    np.testing.assert_array_equal(info, gI.get())
    np.testing.assert_array_equal(X[ok], gB.get()[ok])
    assert np.isnan(F[:, n * n:]).all()
    low = np.triu_indices(n)                 # [col][row] storage: row >= col
    up = np.tril_indices(n, -1)
    for i in range(batch):
        So, Bo = S[i].copy(), B[i].copy()
        io = orc.potrf(t, 102, "L", n, So, n)
        if i == 9:
            io = 5        # the batched entry points have no LAPACKE NaN screen: ?potf2 reports the NaN pivot of column 5
        assert info[i] == io, (i, info[i], io)
        Fi = F[i, :n * n].reshape(n, n)
        np.testing.assert_array_equal(Fi[up], S[i][up])
        if io:
            continue
        np.testing.assert_array_equal(Fi[low], So[low])
        orc.potrs(t, 102, "L", n, nrhs, So, n, Bo, n)
        assert fro(X[i] - Bo) <= 1e3 * n * EPS[t] * fro(Bo)


@pytest.mark.parametrize("variant", [30, 31, 7, 6, 17, 16, 40, 32, 50])
def test_lu32_fast_kernels_special_values_and_strides(orc, rng, variant):
    """Every shape of the tuned 32x32 Double kernels (jb_set_option lu32.variant) against the oracle, bit for bit, on a
    batch that mixes generic systems with everything the fast path must hand to the exact routine: exact ties (integer
    matrices), a zero column (info > 0), NaN and Inf entries (NaNs never win the pivot search: i?amax), denormal and
    huge pivots outside the fast reciprocal's range — with padded strides, an odd batch, and all three entry points."""
    lib.jb_init(-1)
    n, batch = 32, 333
    A = batch_rand(rng, "d", batch, n); b = batch_rand(rng, "d", batch, n, 1)
    A[3:40] = rng.integers(-2, 3, A[3:40].shape)
    A[41, 5, :] = 0.0                         # column 5 of system 41 is zero: info = 6
    A[42, 7, 11] = np.nan                     # a NaN inside column 7
    A[43, :, 3] = np.nan                      # a NaN row
    A[44, 9, 20] = np.inf
    A[45] *= 1e-300; A[46] *= 1e300           # pivots outside [2^-959, 2^957]
    A[47, 0, :] = [2.0 ** -1074] + [0.0] * 31
    sa, sb = n * n + 64, n + 8                # padded strides between systems
    Abuf = np.full((batch, sa), np.nan); Abuf[:, :n * n] = A.reshape(batch, -1)
    bbuf = np.full((batch, sb), np.nan); bbuf[:, :n] = b.reshape(batch, -1)
    # reference: the oracle's unscreened ?getf2 + ?getrs (the batched entry points have no LAPACKE NaN screen)
    getf2, getrs = orc.lib.oracle_dgetf2, orc.lib.oracle_dgetrs_cm
    getf2.restype = C.c_int; getf2.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    getrs.restype = None; getrs.argtypes = [C.c_char, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    LUo, Xo = A.copy(), b.copy()
    pivo = np.zeros((batch, n), np.int32); infoo = np.zeros(batch, np.int32)
    for i in range(batch):
        infoo[i] = getf2(n, n, LUo[i].ctypes.data, n, pivo[i].ctypes.data)
        if infoo[i] == 0:
            getrs(b"N", n, 1, LUo[i].ctypes.data, n, pivo[i].ctypes.data, Xo[i].ctypes.data, n)
    ok = infoo == 0
    assert infoo[41] == 6 and ok.sum() >= batch - 3
    lib.jb_set_option(b"lu32.variant", variant)
    try:
        dA, dB = Dev(Abuf.copy()), Dev(bbuf.copy())
        dP, dI = Dev(np.zeros((batch, n), np.int32)), Dev(np.full(batch, -7, np.int32))
        check(lib.jb_dgesv_batched(n, 1, dA.ptr, n, sa, dP.ptr, dB.ptr, n, sb, dI.ptr, batch), "gesv_batched"); check(lib.jb_sync(), "sync")
        LU, X = dA.get(), dB.get()
        np.testing.assert_array_equal(dI.get(), infoo)
        np.testing.assert_array_equal(dP.get(), pivo)
        assert_same_bits(LU[:, :n * n], LUo.reshape(batch, -1))
        assert_same_bits(X[ok, :n], Xo[ok].reshape(-1, n))
        assert_same_bits(X[~ok, :n], b[~ok].reshape(-1, n))
        assert np.isnan(LU[:, n * n:]).all() and np.isnan(X[:, n:]).all()                                      # padding untouched
        # getrf only, and the solve-only contract (inputs untouched)
        dA2, dP2, dI2 = Dev(Abuf.copy()), Dev(np.zeros((batch, n), np.int32)), Dev(np.full(batch, -7, np.int32))
        check(lib.jb_dgetrf_batched(n, dA2.ptr, n, sa, dP2.ptr, dI2.ptr, batch), "getrf_batched"); check(lib.jb_sync(), "sync")
        assert_same_bits(dA2.get()[:, :n * n], LUo.reshape(batch, -1))
        np.testing.assert_array_equal(dP2.get(), pivo); np.testing.assert_array_equal(dI2.get(), infoo)
        dA3, dB3, dX3, dI3 = Dev(Abuf.copy()), Dev(bbuf.copy()), Dev(np.zeros((batch, sb))), Dev(np.zeros(batch, np.int32))
        check(lib.jb_dsolve_batched(n, 1, dA3.ptr, n, sa, dB3.ptr, n, sb, dX3.ptr, n, sb, dI3.ptr, batch), "dsolve_batched"); check(lib.jb_sync(), "sync")
        assert_same_bits(dX3.get()[ok, :n], Xo[ok].reshape(-1, n))
        np.testing.assert_array_equal(dA3.get().view(np.int64), Abuf.view(np.int64))
        np.testing.assert_array_equal(dI3.get(), infoo)
    finally:
        lib.jb_set_option(b"lu32.variant", -1)


@pytest.mark.parametrize("t", "scz")
def test_lu32g_fast_kernels_special_values_and_strides(orc, rng, t):
    """The tuned 32x32 Float / Complex Float / Complex Double kernels (csrc/batched_lu32g.cu) against the oracle's
    unscreened ?getf2 + ?getrs, bit for bit, on a batch that mixes generic systems with everything the fast path hands
    to the exact routine: exact ties (integer matrices), a zero column (info > 0), NaN and Inf entries, tiny and huge
    scales — with padded strides between systems and an odd batch, through gesv and getrf."""
    lib.jb_init(-1)
    n, batch = 32, 257
    rt = "s" if t in "sc" else "d"
    A = batch_rand(rng, t, batch, n); b = batch_rand(rng, t, batch, n, 1)
    A[3:40] = rng.integers(-2, 3, A[3:40].shape)
    if t in "cz":
        A[20:40] = A[20:40] + 1j * rng.integers(-2, 3, A[20:40].shape)
    A[41, 5, :] = 0.0                         # column 5 of system 41 is zero: info = 6
    A[42, 7, 11] = np.nan                     # a NaN inside column 7
    A[43, :, 3] = np.nan                      # a NaN row
    A[44, 9, 20] = np.inf
    tiny, huge = (1e-30, 1e30) if rt == "s" else (1e-300, 1e300)
    A[45] *= tiny; A[46] *= huge
    A[47, 0, :] = 0.0; A[47, 0, 0] = np.finfo(NP[rt]).smallest_subnormal
    w = 2 if t in "cz" else 1
    sa, sb = n * n + 16, n + 4                # padded strides between systems (in elements)
    pad = np.nan if t == "s" else complex(np.nan, np.nan)
    Abuf = np.full((batch, sa), pad, NP[t]); Abuf[:, :n * n] = A.reshape(batch, -1)
    bbuf = np.full((batch, sb), pad, NP[t]); bbuf[:, :n] = b.reshape(batch, -1)
    getf2, getrs = getattr(orc.lib, f"oracle_{t}getf2"), getattr(orc.lib, f"oracle_{t}getrs_cm")
    getf2.restype = C.c_int; getf2.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    getrs.restype = None; getrs.argtypes = [C.c_char, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    LUo, Xo = A.copy(), b.copy()
    pivo = np.zeros((batch, n), np.int32); infoo = np.zeros(batch, np.int32)
    for i in range(batch):
        infoo[i] = getf2(n, n, LUo[i].ctypes.data, n, pivo[i].ctypes.data)
        if infoo[i] == 0:
            getrs(b"N", n, 1, LUo[i].ctypes.data, n, pivo[i].ctypes.data, Xo[i].ctypes.data, n)
    ok = infoo == 0
    assert infoo[41] == 6 and ok.sum() >= batch - 3
    real = lambda x: np.ascontiguousarray(x).view(NP[rt])
    dA, dB = Dev(Abuf.copy()), Dev(bbuf.copy())
    dP, dI = Dev(np.zeros((batch, n), np.int32)), Dev(np.full(batch, -7, np.int32))
    check(getattr(lib, f"jb_{t}gesv_batched")(n, 1, dA.ptr, n, sa, dP.ptr, dB.ptr, n, sb, dI.ptr, batch), "gesv_batched"); check(lib.jb_sync(), "sync")
    LU, X = dA.get(), dB.get()
    np.testing.assert_array_equal(dI.get(), infoo)
    np.testing.assert_array_equal(dP.get(), pivo)
    assert_same_bits(real(LU[:, :n * n]), real(LUo.reshape(batch, -1)))
    assert_same_bits(real(X[ok, :n]), real(Xo[ok].reshape(-1, n)))
    assert np.isnan(real(LU[:, n * n:])).all() and np.isnan(real(X[:, n:])).all()          # padding untouched
    dA2, dP2, dI2 = Dev(Abuf.copy()), Dev(np.zeros((batch, n), np.int32)), Dev(np.full(batch, -7, np.int32))
    check(getattr(lib, f"jb_{t}getrf_batched")(n, dA2.ptr, n, sa, dP2.ptr, dI2.ptr, batch), "getrf_batched"); check(lib.jb_sync(), "sync")
    assert_same_bits(real(dA2.get()[:, :n * n]), real(LUo.reshape(batch, -1)))
    np.testing.assert_array_equal(dP2.get(), pivo); np.testing.assert_array_equal(dI2.get(), infoo)
    # the fast path and the exact routine (lu32.v1 = 1) agree bit for bit on the whole batch
    lib.jb_set_option(b"lu32.v1", 1)
    try:
        dA3, dB3 = Dev(Abuf.copy()), Dev(bbuf.copy())
        dP3, dI3 = Dev(np.zeros((batch, n), np.int32)), Dev(np.full(batch, -7, np.int32))
        check(getattr(lib, f"jb_{t}gesv_batched")(n, 1, dA3.ptr, n, sa, dP3.ptr, dB3.ptr, n, sb, dI3.ptr, batch), "gesv_batched"); check(lib.jb_sync(), "sync")
    finally:
        lib.jb_set_option(b"lu32.v1", -1)
    assert_same_bits(real(dA3.get()), real(LU)); assert_same_bits(real(dB3.get()[ok]), real(X[ok]))
    np.testing.assert_array_equal(dP3.get(), pivo); np.testing.assert_array_equal(dI3.get(), infoo)


def test_config3_full_size_properties():
    """configs[2] at full size: 1,000,000 independent 32x32 Double systems from the synthetic
    generator.  Size-independent checks: every info is 0, every ipiv entry is in [k, n],
    the batch residual |A x - b| is at the n*eps level, and the first 2000 systems equal a
    standalone run on that prefix bit for bit (each system independent of its neighbours)."""
    lib.jb_init(-1)
    n, batch = 32, 1_000_000
    dA = lib.jb_malloc(batch * n * n * 8, 0); dB = lib.jb_malloc(batch * n * 8, 0)
    dP = lib.jb_malloc(batch * n * 4, 0); dI = lib.jb_malloc(batch * 4, 0)
    assert dA and dB and dP and dI
    try:
        # the whole batch is one (n) x (batch*n) column-major matrix of U(-1,1)
        check(lib.jb_dfill_uniform(dA, n, n, batch * n, 0, 0, 20261019, 0, 0.0), "fill A")
        check(lib.jb_dfill_uniform(dB, n, n, batch, 0, 0, 20261020, 0, 0.0), "fill b")
        check(lib.jb_dgesv_batched(n, 1, dA, n, n * n, dP, dB, n, n, dI, batch), "gesv_batched")
        check(lib.jb_sync(), "sync")
        info = np.empty(batch, np.int32); piv = np.empty((batch, n), np.int32)
        check(lib.jb_memcpy_d2h(info.ctypes.data, dI, info.nbytes), "d2h")
        check(lib.jb_memcpy_d2h(piv.ctypes.data, dP, piv.nbytes), "d2h")
        assert (info == 0).all()
        assert (piv >= np.arange(1, n + 1)[None, :]).all() and (piv <= n).all()
        X = np.empty((batch, n)); check(lib.jb_memcpy_d2h(X.ctypes.data, dB, X.nbytes), "d2h")
        sel = np.r_[0:2000, batch - 2000:batch]
        worst = 0.0
        for i in sel:
            A = synth.uniform(n, n, 20261019, "d", 0, i * n)
            b = synth.uniform(n, 1, 20261020, "d", 0, i)[:, 0]
            r = np.linalg.norm(A @ X[i] - b) / (np.linalg.norm(A) * np.linalg.norm(X[i]) * n * EPS["d"])
            worst = max(worst, r)
        assert worst <= 10, worst
        # prefix run
        m = 2000
        A0 = synth.uniform(n, m * n, 20261019, "d"); b0 = synth.uniform(n, m, 20261020, "d")
        dA0, dB0 = Dev(np.ascontiguousarray(A0.T)), Dev(np.ascontiguousarray(b0.T))
        dP0, dI0 = Dev(np.zeros((m, n), np.int32)), Dev(np.zeros(m, np.int32))
        check(lib.jb_dgesv_batched(n, 1, dA0.ptr, n, n * n, dP0.ptr, dB0.ptr, n, n, dI0.ptr, m), "gesv prefix")
        check(lib.jb_sync(), "sync")
        np.testing.assert_array_equal(dB0.get(), X[:m])
        np.testing.assert_array_equal(dP0.get(), piv[:m])
    finally:
        for p in (dA, dB, dP, dI):
            lib.jb_free(p)


def test_fast_reciprocal_is_correctly_rounded(rng):
    """The tuned 32x32 kernel's MUFU.RCP64H + Newton reciprocal must equal IEEE 1/x bit for bit on its
    guarded exponent range (outside it the kernel falls back to IEEE division)."""
    lib.jb_init(-1)
    lib.jb_test_rcp_fast.restype = C.c_longlong
    lib.jb_test_rcp_fast.argtypes = [C.c_void_p, C.c_int]
    n = 1 << 22
    mant = rng.uniform(1.0, 2.0, n)
    mant[: n // 4] = 1.0 + rng.integers(0, 64, n // 4) * 2.0 ** -52          # just above powers of two
    mant[n // 4: n // 2] = 2.0 - rng.integers(1, 64, n // 4) * 2.0 ** -52     # significands of all ones
    x = np.ldexp(mant, rng.integers(-900, 900, n)) * rng.choice([-1.0, 1.0], n)
    d = Dev(x)
    assert lib.jb_test_rcp_fast(d.ptr, n) == 0
