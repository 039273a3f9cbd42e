"""The host-side mirror of Jalla's interface (jalla_b200/matrix.py) — tests written like
the reference's own (tests/Test.hs) plus the solve / invert paths the reference never tests."""
import numpy as np
import pytest

from tests.util import NP, EPS, jalla_arbitrary, rand, fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    from jalla_b200 import matrix as J
    from jalla_b200._lib import lib
    assert lib.jb_init(-1) == 0
    return J


@pytest.mark.parametrize("t", "sdcz")
def test_mm_operators(J, rng, t):
    a = jalla_arbitrary(rng, t, 37, 58); b = jalla_arbitrary(rng, t, 58, 21)
    A, B = J.Matrix.from_numpy(a), J.Matrix.from_numpy(b)
    c = (A @ B).to_numpy()
    ref = a.astype(np.complex128) @ b.astype(np.complex128)
    assert fro(c - ref) / fro(ref) <= 4 * 58 * EPS[t]
    c2 = J.mmT((A, J.Trans), (A, J.NoTrans)).to_numpy()
    ref2 = a.T.astype(np.complex128) @ a.astype(np.complex128)
    assert fro(c2 - ref2) / fro(ref2) <= 4 * 37 * EPS[t]
    with pytest.raises(ValueError):
        A @ A


def test_prop_pseudoInverse_gemm_tail(J, rng):
    """tests/Test.hs:106-112: |I - A A+| < 1e-6.  gesvd is out of scope, so
    the SVD comes from numpy; the two GEMMs of pseudoInverse (Matrix.hs:544-551) run on the GPU."""
    for _ in range(5):
        mat = jalla_arbitrary(rng, "d")
        m, n = mat.shape
        u, s, vt = np.linalg.svd(mat, full_matrices=False)
        V = J.Matrix.from_numpy(np.asfortranarray(vt.T * (1.0 / s)))      # matrixMultDiag (vt,Trans) (1/s)
        U = J.Matrix.from_numpy(np.asfortranarray(u))
        plus = J.mmT((V, J.NoTrans), (U, J.Trans))                        # one ##! with (NoTrans, Trans)
        M = J.Matrix.from_numpy(mat)
        a = np.eye(m) - (M @ plus).to_numpy() if m < n else np.eye(n) - (plus @ M).to_numpy()
        assert fro(a) < 1e-6


@pytest.mark.parametrize("t", "sdcz")
def test_solveLinearSystem_and_invert(J, rng, t):
    n = 96
    a = rand(rng, t, n, n); b = rand(rng, t, n, 4)
    A, B = J.Matrix.from_numpy(a), J.Matrix.from_numpy(b)
    x = J.solveLinearSystem(A, B).to_numpy()
    np.testing.assert_array_equal(A.to_numpy(), a)        # inputs are copied, not clobbered (Matrix.hs:519-520)
    np.testing.assert_array_equal(B.to_numpy(), b)
    assert fro(a.astype(np.complex128) @ x - b) <= 10 * n * EPS[t] * fro(a) * fro(x)
    inv = J.invert(A)
    assert inv is not None
    assert fro(a.astype(np.complex128) @ inv.to_numpy() - np.eye(n)) <= 10 * n * EPS[t] * fro(a) * fro(inv.to_numpy())
    inv2 = J.invert_(A).to_numpy()
    assert fro(inv2 - inv.to_numpy()) <= 100 * n * EPS[t] * fro(a) * fro(inv2) ** 2


def test_error_behaviour(J, rng):
    a = J.Matrix.from_numpy(rand(rng, "d", 5, 4))
    assert J.invert(a) is None                              # non-square -> Nothing (Matrix.hs:540)
    sing = np.ones((6, 6), order="F")
    assert J.invert(J.Matrix.from_numpy(sing)) is None      # singular -> Nothing (Matrix.hs:767-772)
    with pytest.raises(ValueError):
        J.unsafeSolveLinearSystem(a, a)                     # shape mismatch -> error (Matrix.hs:508)
    with pytest.raises(ArithmeticError):
        J.solveLinearSystem(J.Matrix.from_numpy(sing), J.Matrix.from_numpy(np.ones((6, 1), order="F")))


def test_matrixCopy_transpose(J, rng):
    for t in "sdcz":
        a = rand(rng, t, 45, 70)
        A = J.Matrix.from_numpy(a)
        np.testing.assert_array_equal(J.matrixCopy(A, J.NoTrans).to_numpy(), a)
        np.testing.assert_array_equal(J.matrixCopy(A, J.Trans).to_numpy(), a.T)


def test_managed_memory_is_host_dereferenceable(J, rng):
    """Jalla's generic code peeks/pokes matrix pointers on the host.  A buffer from
    jb_malloc(JB_MEM_MANAGED) can be filled and read through a plain host pointer while the C ABI treats it as
    device-resident (no staging copy)."""
    import ctypes as C
    from jalla_b200 import _lib
    from jalla_b200._lib import lib, check, check_last
    n = 96
    bufs = [lib.jb_malloc(n * n * 8, _lib.JB_MEM_MANAGED) for _ in range(3)]
    assert all(bufs)
    try:
        views = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n * n,)) for p in bufs]
        a = rng.uniform(-1, 1, (n, n)); b = rng.uniform(-1, 1, (n, n))
        views[0][:] = np.asfortranarray(a).ravel(order="F")          # host "poke"
        views[1][:] = np.asfortranarray(b).ravel(order="F")
        views[2][:] = np.nan
        lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, bufs[0], n, bufs[1], n, 0.0, bufs[2], n)
        check_last("gemm"); check(lib.jb_sync(), "sync")
        c = views[2].reshape(n, n, order="F")                          # host "peek"
        assert fro(c - a @ b) / fro(a @ b) <= 4 * n * EPS["d"]
        piv = (C.c_int * n)()
        rhs = views[1][:n].copy()
        assert lib.LAPACKE_dgesv(102, n, 1, bufs[0], n, piv, bufs[1], n) == 0
        x = views[1][:n]
        assert fro(a @ x - rhs) <= 1e-9
    finally:
        for p in bufs:
            lib.jb_free(p)


@pytest.mark.parametrize("t", "sdcz")
def test_device_construction_helpers(J, rng, t):
    """idMatrix / fill / setDiag / matrixMultDiag on the device (Matrix.hs:525-526,660-665,1023-1036) and the reference's
    own row-by-row matrixCopy (Matrix.hs:855-878) running on device pointers through cblas_?copy."""
    n = 45
    np.testing.assert_array_equal(J.idMatrix(n, t).to_numpy(), np.eye(n, dtype=NP[t]))
    f = J.fillMatrix((7, 11), 2.5, t)
    np.testing.assert_array_equal(f.to_numpy(), np.full((7, 11), 2.5, dtype=NP[t]))
    J.setDiag(f, 2, [1, 2, 3] + [9] * 20)                      # more values than fit: the surplus is ignored
    ref = np.full((7, 11), 2.5, dtype=NP[t]); ref[np.arange(7), np.arange(7) + 2] = ([1, 2, 3] + [9] * 4)
    np.testing.assert_array_equal(f.to_numpy(), ref)
    a = jalla_arbitrary(rng, t, 33, 58)
    A = J.Matrix.from_numpy(a)
    for tr in (J.NoTrans, J.Trans):
        src = a if tr == J.NoTrans else a.T
        np.testing.assert_array_equal(J.referenceMatrixCopy(A, tr).to_numpy(), src)
        np.testing.assert_array_equal(J.matrixCopy(A, tr).to_numpy(), src)
        d = rng.uniform(-2, 2, src.shape[1]).astype(NP[t])
        got = J.matrixMultDiag((A, tr), d).to_numpy()
        # tests/Test.hs:77-103: multiplying by a diagonal matrix through GEMM equals the column scaling exactly
        D = J.Matrix.from_numpy(np.asfortranarray(np.diag(d)))
        via_gemm = J.mmT((A, tr), (D, J.NoTrans)).to_numpy()
        np.testing.assert_array_equal(got, (src * d[None, :]).astype(NP[t]))
        assert fro(got - via_gemm) <= 4 * EPS[t] * fro(got)


def test_prop_pseudoInverse_on_device(J, rng):
    """tests/Test.hs:106-112 with the whole tail of pseudoInverse (Matrix.hs:544-551) on the device:
    matrixMultDiag (vt,Trans) s' then ##! with (u,Trans); only the SVD (gesvd, out of scope) is numpy's."""
    for _ in range(5):
        mat = jalla_arbitrary(rng, "d")
        m, n = mat.shape
        u, s, vt = np.linalg.svd(mat, full_matrices=False)
        P = J.pseudoInverseFromSVD(J.Matrix.from_numpy(np.asfortranarray(u)), s, J.Matrix.from_numpy(np.asfortranarray(vt)))
        Mx = J.Matrix.from_numpy(mat)
        a = np.eye(m) - (Mx @ P).to_numpy() if m < n else np.eye(n) - (P @ Mx).to_numpy()
        assert np.linalg.norm(a) < 1e-6
