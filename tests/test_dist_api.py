"""jb_dist_*: the single-process multi-GPU forms of the hot path behind the C ABI, against OpenBLAS
on the host with north_star's criteria: GEMM relative Frobenius error <= 4*k*eps*|A||B|/|C|; LU pivot vectors
bit-identical to LAPACKE_dgetrf and |PA - LU| / (n*eps*|A|) <= 10; Cholesky / solve residuals at the n*eps level.
The grid uses every GPU the box offers (1x1 on a single-GPU box, where the same code runs without peers)."""
import ctypes as C

import numpy as np
import pytest

from jalla_b200 import _lib
from jalla_b200._lib import lib, check
from jalla_b200.dist import grid_shape
from tests.util import EPS, fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    assert lib.jb_init(-1) == 0
    ngpu = lib.jb_device_count()
    world = max(w for w in (1, 2, 4, 8) if w <= ngpu)
    pr, pc = grid_shape(world)
    h = C.c_void_p()
    check(lib.jb_dist_create(C.byref(h), pr, pc, 128), "jb_dist_create")
    assert lib.jb_dist_ngpus(h) == world
    yield h
    lib.jb_dist_destroy(h)


@pytest.mark.parametrize("ta,tb", [(111, 111), (112, 111), (111, 112), (112, 112)])
def test_dist_dgemm(handle, blas, rng, ta, tb):
    m, n, k = 900, 1100, 1300
    A = np.asfortranarray(rng.uniform(-1, 1, (m, k) if ta == 111 else (k, m)))
    B = np.asfortranarray(rng.uniform(-1, 1, (k, n) if tb == 111 else (n, k)))
    C0 = np.asfortranarray(rng.uniform(-1, 1, (m + 3, n))); C0[m:] = np.nan
    for alpha, beta in ((1.0, 0.0), (0.5, -2.0)):
        Cg, Cr = C0.copy(order="F"), C0.copy(order="F")
        check(lib.jb_dist_dgemm(handle, ta, tb, m, n, k, alpha, A.ctypes.data, A.shape[0], B.ctypes.data, B.shape[0], beta, Cg.ctypes.data, m + 3), "dist_dgemm")
        blas.gemm("d", 102, ta, tb, m, n, k, alpha, A, A.shape[0], B, B.shape[0], beta, Cr, m + 3)
        bound = 4 * k * EPS["d"] * fro(A) * fro(B) / fro(Cr[:m])
        assert fro(Cg[:m] - Cr[:m]) / fro(Cr[:m]) <= bound
        assert np.isnan(Cg[m:]).all()


@pytest.mark.parametrize("n", [100, 1000, 2500])
def test_dist_dgetrf_getrs(handle, blas, rng, n):
    A = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
    B = np.asfortranarray(rng.uniform(-1, 1, (n, 3)))
    LU, piv = A.copy(order="F"), np.zeros(n, np.int32)
    info = lib.jb_dist_dgetrf(handle, n, LU.ctypes.data, n, piv.ctypes.data)
    check(info, "dist_dgetrf"); assert info == 0
    LUr, pivr = A.copy(order="F"), np.zeros(n, np.int32)
    assert blas.getrf("d", 102, n, n, LUr, n, pivr) == 0
    np.testing.assert_array_equal(piv, pivr)                       # bit-identical pivot vector (no near-ties in U(-1,1) data)
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    PA = A.copy()
    for i in range(n):
        p = piv[i] - 1
        if p != i:
            PA[[i, p]] = PA[[p, i]]
    assert fro(PA - L @ U) / (n * EPS["d"] * fro(A)) <= 10
    X = B.copy(order="F")
    check(lib.jb_dist_dgetrs(handle, b"N", n, 3, LU.ctypes.data, n, piv.ctypes.data, X.ctypes.data, n), "dist_dgetrs")
    assert fro(A @ X - B) <= 1e3 * n * EPS["d"] * fro(A) * fro(X)
    # factors not resident (a different host array): they are brought in again
    LU2 = LU.copy(order="F"); X2 = B.copy(order="F")
    check(lib.jb_dist_dgetrs(handle, b"N", n, 3, LU2.ctypes.data, n, piv.ctypes.data, X2.ctypes.data, n), "dist_dgetrs")
    np.testing.assert_array_equal(X2, X)


def test_dist_dgetrf_singular_reports_info(handle, rng):
    n = 700
    A = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
    A[:, 300] = 0.0
    piv = np.zeros(n, np.int32)
    assert lib.jb_dist_dgetrf(handle, n, A.ctypes.data, n, piv.ctypes.data) == 301


@pytest.mark.parametrize("n", [130, 1000, 2300])
def test_dist_dpotrf_potrs(handle, blas, rng, n):
    M = rng.uniform(-1, 1, (n, n))
    S = np.asfortranarray(M @ M.T + n * np.eye(n))
    B = np.asfortranarray(rng.uniform(-1, 1, (n, 2)))
    F = S.copy(order="F")
    info = lib.jb_dist_dpotrf(handle, b"L", n, F.ctypes.data, n)
    check(info, "dist_dpotrf"); assert info == 0
    Fr = S.copy(order="F")
    assert blas.potrf("d", 102, "L", n, Fr, n) == 0
    tl = np.tril_indices(n)
    assert fro(F[tl] - Fr[tl]) <= 50 * n * EPS["d"] * fro(Fr[tl])
    np.testing.assert_array_equal(F[np.triu_indices(n, 1)], S[np.triu_indices(n, 1)])      # strictly upper part untouched
    X = B.copy(order="F")
    check(lib.jb_dist_dpotrs(handle, b"L", n, 2, F.ctypes.data, n, X.ctypes.data, n), "dist_dpotrs")
    assert fro(S @ X - B) <= 1e3 * n * EPS["d"] * fro(S) * fro(X)
    S[n // 2, n // 2] = -1.0                                         # not positive definite: info = first failing minor
    assert lib.jb_dist_dpotrf(handle, b"L", n, S.ctypes.data, n) == n // 2 + 1


def test_dist_dgesv_batched_matches_oracle(handle, orc, rng):
    n, batch = 32, 1001
    A = rng.uniform(-1, 1, (batch, n, n)); b = rng.uniform(-1, 1, (batch, 1, n))
    LUo, Xo = A.copy(), b.copy()
    pivo = np.zeros((batch, n), np.int32)
    for i in range(batch):
        assert orc.gesv("d", 102, n, 1, LUo[i], n, pivo[i], Xo[i], n) == 0
    LU, X, piv, info = A.copy(), b.copy(), np.zeros((batch, n), np.int32), np.full(batch, -1, np.int32)
    check(lib.jb_dist_dgesv_batched(handle, n, 1, LU.ctypes.data, n, n * n, piv.ctypes.data, X.ctypes.data, n, n, info.ctypes.data, batch), "dist batched")
    assert (info == 0).all()
    np.testing.assert_array_equal(piv, pivo)
    np.testing.assert_array_equal(LU, LUo)            # bit-exact, like the single-GPU batched path
    np.testing.assert_array_equal(X, Xo)
