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
