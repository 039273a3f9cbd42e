"""The rows SURVEY.md §8(f) marks "next", through the C ABI with DEVICE pointers and against OpenBLAS on the host:
device-aware level-1 BLAS (what Jalla's matrixCopy / scaleColumn reach: Matrix.hs:855-878,1007-1009), cblas_?trmm /
?symm / ?hemm, LAPACKE_?potri, the device fill / setDiag / idMatrix / matrixMultDiag helpers, and a replay of the exact
symbol sequence the reference issues for solveLinearSystem on a GPU-resident matrix (r strided cblas_dcopy calls per
matrixCopy, then LAPACKE_dgesv: Matrix.hs:513-521)."""
import ctypes as C
import itertools

import numpy as np
import pytest

from jalla_b200 import _lib
from jalla_b200._lib import lib, check, check_last
from tests.util import NP, EPS, rand, fro
from tests.gpu_backend import DevBuf

pytestmark = pytest.mark.gpu
_i, _p, _d, _f = C.c_int, C.c_void_p, C.c_double, C.c_float
REAL = {"s": _f, "d": _d, "c": _f, "z": _d}


def ob(blas, name, res, *args):
    f = getattr(blas.lib, "scipy_" + name)
    f.restype, f.argtypes = res, list(args)
    return f


def vec(rng, t, n, inc):
    """Strided vector storage of n elements with increment |inc| (NaN between the elements)."""
    a = np.full(max(1 + (n - 1) * abs(inc), 1), np.nan, dtype=NP[t])
    v = rng.uniform(-1, 1, n) + (1j * rng.uniform(-1, 1, n) if t in "cz" else 0)
    a[:: abs(inc)][:n] = v.astype(NP[t])
    return a


def scalar(t, v):
    if t in "sd":
        return REAL[t](v), None
    arr = np.array([v], dtype=NP[t])
    return C.c_void_p(arr.ctypes.data), arr


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("incx,incy", [(1, 1), (3, 2), (-2, 1), (1, -3), (7, 7)])
def test_copy_swap_axpy_scal_device(blas, rng, t, incx, incy):
    lib.jb_init(-1)
    n = 1000
    x, y = vec(rng, t, n, incx), vec(rng, t, n, incy)
    for op in ("copy", "swap", "axpy"):
        xr, yr = x.copy(), y.copy()
        dx, dy = DevBuf(x.copy()), DevBuf(y.copy())
        al = 0.75 if t in "sd" else 0.75 - 0.5j
        if op == "axpy":
            a_gpu, k1 = scalar(t, al); a_ref, k2 = scalar(t, al)
            getattr(lib, f"cblas_{t}axpy")(n, a_gpu, dx.ptr, incx, dy.ptr, incy)
            ob(blas, f"cblas_{t}axpy", None, _i, REAL[t] if t in "sd" else _p, _p, _i, _p, _i)(n, a_ref, xr.ctypes.data, incx, yr.ctypes.data, incy)
        else:
            getattr(lib, f"cblas_{t}{op}")(n, dx.ptr, incx, dy.ptr, incy)
            ob(blas, f"cblas_{t}{op}", None, _i, _p, _i, _p, _i)(n, xr.ctypes.data, incx, yr.ctypes.data, incy)
        check_last(op); check(lib.jb_sync(), "sync")
        dx.back(); dy.back()
        sel_x, sel_y = slice(None, None, abs(incx)), slice(None, None, abs(incy))
        if op == "axpy":
            assert fro(dy.arr[sel_y] - yr[sel_y]) <= 4 * EPS[t] * fro(yr[sel_y])
        else:
            np.testing.assert_array_equal(dy.arr[sel_y], yr[sel_y])
        np.testing.assert_array_equal(dx.arr[sel_x], xr[sel_x])
        gaps = np.ones(dy.arr.shape, bool); gaps[sel_y] = False
        assert np.isnan(dy.arr[gaps]).all()                      # nothing between the elements was touched
        dx.free(); dy.free()
    # scal (and the real-scalar complex form)
    dx = DevBuf(x.copy()); xr = x.copy()
    a_gpu, k1 = scalar(t, 0.5 if t in "sd" else 0.5 + 0.25j)
    getattr(lib, f"cblas_{t}scal")(n, a_gpu, dx.ptr, abs(incx))
    ob(blas, f"cblas_{t}scal", None, _i, REAL[t] if t in "sd" else _p, _p, _i)(n, a_gpu, xr.ctypes.data, abs(incx))
    check_last("scal"); check(lib.jb_sync(), "sync"); dx.back()
    assert fro(dx.arr[:: abs(incx)] - xr[:: abs(incx)]) <= 4 * EPS[t] * fro(xr[:: abs(incx)])
    dx.free()
    if t in "cz":
        dx = DevBuf(x.copy()); xr = x.copy()
        nm = f"cblas_{t}{'s' if t == 'c' else 'd'}scal"
        getattr(lib, nm)(n, REAL[t](1.5), dx.ptr, abs(incx)); ob(blas, nm, None, _i, REAL[t], _p, _i)(n, REAL[t](1.5), xr.ctypes.data, abs(incx))
        check_last("rscal"); check(lib.jb_sync(), "sync"); dx.back()
        np.testing.assert_array_equal(dx.arr[:: abs(incx)], xr[:: abs(incx)])
        dx.free()


@pytest.mark.parametrize("t", "sdcz")
def test_reductions_device_and_host(blas, rng, t):
    lib.jb_init(-1)
    n, incx, incy = 100_003, 2, -1
    x, y = vec(rng, t, n, incx), vec(rng, t, n, incy)
    x[2 * 4321] = 7.0                                             # a unique maximum for i?amax
    dx, dy = DevBuf(x.copy()), DevBuf(y.copy())
    r = REAL[t]
    pre = {"s": "s", "d": "d", "c": "sc", "z": "dz"}[t]
    for where, px, py in (("device", dx.ptr, dy.ptr), ("host", x.ctypes.data, y.ctypes.data)):
        f = getattr(lib, f"cblas_{pre}nrm2"); got = f(n, px, incx)
        ref = ob(blas, f"cblas_{pre}nrm2", r, _i, _p, _i)(n, x.ctypes.data, incx)
        assert abs(got - ref) <= 8 * EPS[t] * ref, where
        got = getattr(lib, f"cblas_{pre}asum")(n, px, incx); ref = ob(blas, f"cblas_{pre}asum", r, _i, _p, _i)(n, x.ctypes.data, incx)
        assert abs(got - ref) <= 64 * EPS[t] * ref * (40 if t in "sc" else 1), where
        got = getattr(lib, f"cblas_i{t}amax")(n, px, incx); ref = ob(blas, f"cblas_i{t}amax", C.c_size_t, _i, _p, _i)(n, x.ctypes.data, incx)
        assert got == ref == 4321, where
        if t in "sd":
            got = getattr(lib, f"cblas_{t}dot")(n, px, incx, py, incy)
            ref = ob(blas, f"cblas_{t}dot", r, _i, _p, _i, _p, _i)(n, x.ctypes.data, incx, y.ctypes.data, incy)
            xs, ys = x[::incx][:n].astype(np.float64), y[::1][:n][::-1].astype(np.float64)
            assert abs(got - ref) <= 50 * EPS[t] * float(np.abs(xs * ys).sum()), where
            if t == "s":
                got = lib.cblas_dsdot(n, px, incx, py, incy)
                assert abs(got - float(xs @ ys)) <= 1e-9 * float(np.abs(xs * ys).sum())
        else:
            for nm in ("dotu_sub", "dotc_sub"):
                out_g, out_r = np.zeros(1, NP[t]), np.zeros(1, NP[t])
                getattr(lib, f"cblas_{t}{nm}")(n, px, incx, py, incy, out_g.ctypes.data)
                ob(blas, f"cblas_{t}{nm}", None, _i, _p, _i, _p, _i, _p)(n, x.ctypes.data, incx, y.ctypes.data, incy, out_r.ctypes.data)
                assert abs(out_g[0] - out_r[0]) <= 100 * EPS[t] * n ** 0.5, (where, nm)
        check_last("reductions")
    dx.free(); dy.free()


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("side,uplo,trans,diag", list(itertools.product((141, 142), (121, 122), (111, 112, 113), (131, 132))))
def test_trmm_all_variants(blas, rng, t, side, uplo, trans, diag):
    lib.jb_init(-1)
    m, n = 70, 45
    ka = m if side == 141 else n
    A = rand(rng, t, ka, ka, ka + 1); B = rand(rng, t, m, n, m + 2)
    al = 0.5 if t in "sd" else 0.5 - 0.25j
    for order in (102, 101):
        Bo = B if order == 102 else rand(rng, t, n, m, n + 2)
        ldb = Bo.shape[0]
        dA, dB = DevBuf(A.copy(order="F")), DevBuf(Bo.copy(order="F")); Br = Bo.copy(order="F")
        a1, k1 = scalar(t, al)
        sc = REAL[t] if t in "sd" else _p
        getattr(lib, f"cblas_{t}trmm")(order, side, uplo, trans, diag, m, n, a1, dA.ptr, ka + 1, dB.ptr, ldb)
        ob(blas, f"cblas_{t}trmm", None, _i, _i, _i, _i, _i, _i, _i, sc, _p, _i, _p, _i)(order, side, uplo, trans, diag, m, n, a1, A.ctypes.data, ka + 1, Br.ctypes.data, ldb)
        check_last("trmm"); check(lib.jb_sync(), "sync"); dB.back()
        rows = m if order == 102 else n
        assert fro(dB.arr[:rows] - Br[:rows]) <= 20 * ka * EPS[t] * fro(Br[:rows]), order
        assert np.isnan(dB.arr[rows:]).all()
        dA.free(); dB.free()


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("side,uplo", list(itertools.product((141, 142), (121, 122))))
def test_symm_hemm(blas, rng, t, side, uplo):
    lib.jb_init(-1)
    m, n = 90, 60
    ka = m if side == 141 else n
    al, be = (0.75, -0.5) if t in "sd" else (0.75 - 0.25j, -0.5 + 0.5j)
    for herm in ([False, True] if t in "cz" else [False]):
        nm = f"cblas_{t}{'hemm' if herm else 'symm'}"
        for order in (102, 101):
            A = rand(rng, t, ka, ka, ka + 1)
            B = rand(rng, t, m, n, m + 1) if order == 102 else rand(rng, t, n, m, n + 1)
            C0 = rand(rng, t, m, n, m + 3) if order == 102 else rand(rng, t, n, m, n + 3)
            dA, dB, dC = DevBuf(A.copy(order="F")), DevBuf(B.copy(order="F")), DevBuf(C0.copy(order="F")); Cr = C0.copy(order="F")
            a1, k1 = scalar(t, al); b1, k2 = scalar(t, be)
            sc = REAL[t] if t in "sd" else _p
            getattr(lib, nm)(order, side, uplo, m, n, a1, dA.ptr, ka + 1, dB.ptr, B.shape[0], b1, dC.ptr, C0.shape[0])
            ob(blas, nm, None, _i, _i, _i, _i, _i, sc, _p, _i, _p, _i, sc, _p, _i)(order, side, uplo, m, n, a1, A.ctypes.data, ka + 1, B.ctypes.data, B.shape[0], b1, Cr.ctypes.data, C0.shape[0])
            check_last(nm); check(lib.jb_sync(), "sync"); dC.back()
            rows = m if order == 102 else n
            assert fro(dC.arr[:rows] - Cr[:rows]) <= 20 * ka * EPS[t] * fro(Cr[:rows]), (herm, order)
            assert np.isnan(dC.arr[rows:]).all()
            for b in (dA, dB, dC):
                b.free()


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n", [1, 33, 200])
def test_potri(gpu, blas, rng, t, uplo, n):
    M = rand(rng, t, n, n)
    S = np.asfortranarray(M @ M.conj().T + n * np.eye(n)).astype(NP[t])
    for order in (102, 101):
        F = S.copy(order="F")
        assert blas.potrf(t, order, uplo, n, F, n) == 0
        Fg, Fr = F.copy(order="F"), F.copy(order="F")
        buf = DevBuf(Fg)
        info = getattr(lib, f"LAPACKE_{t}potri")(order, uplo.encode(), n, buf.ptr, n)
        check(lib.jb_sync(), "sync"); buf.back(); buf.free()
        f = ob(blas, f"LAPACKE_{t}potri", _i, _i, C.c_char, _i, _p, _i)
        assert info == f(order, uplo.encode(), n, Fr.ctypes.data, n) == 0
        lower_stored = (uplo == "L") == (order == 102)
        tri = np.tril_indices(n) if lower_stored else np.triu_indices(n)
        other = np.triu_indices(n, 1) if lower_stored else np.tril_indices(n, -1)
        assert fro(Fg[tri] - Fr[tri]) <= 100 * n * EPS[t] * fro(Fr[tri]) * np.linalg.cond(S.astype(np.complex128)) ** 0.5
        np.testing.assert_array_equal(Fg[other], F[other])            # the other triangle is not referenced
    # a singular factor and a NaN are reported like LAPACKE does
    if n > 1:
        F = np.asfortranarray(np.eye(n)).astype(NP[t]); F[n // 2, n // 2] = 0
        buf = DevBuf(F); info = getattr(lib, f"LAPACKE_{t}potri")(102, uplo.encode(), n, buf.ptr, n); buf.free()
        assert info == n // 2 + 1
        F = np.asfortranarray(np.eye(n)).astype(NP[t]); F[0, 0] = np.nan
        buf = DevBuf(F); info = getattr(lib, f"LAPACKE_{t}potri")(102, uplo.encode(), n, buf.ptr, n); buf.free()
        assert info == -4


@pytest.mark.parametrize("t", "sdcz")
def test_laset_set_diag_mult_diag(rng, t):
    lib.jb_init(-1)
    m, n, ld = 37, 52, 40
    buf = DevBuf(np.full((ld, n), np.nan, dtype=NP[t], order="F"))
    if t in "sd":
        check(getattr(lib, f"jb_{t}laset")(m, n, REAL[t](0.0), REAL[t](1.0), buf.ptr, ld), "laset")
    else:
        z = np.array([0, 1], dtype=NP[t])
        check(getattr(lib, f"jb_{t}laset")(m, n, z.ctypes.data, z.ctypes.data + z.itemsize, buf.ptr, ld), "laset")
    check(lib.jb_sync(), "sync"); buf.back()
    np.testing.assert_array_equal(buf.arr[:m], np.eye(m, n, dtype=NP[t]))          # idMatrix = fill 0 >> setDiag 0 (repeat 1)
    assert np.isnan(buf.arr[m:]).all()
    for d in (0, 3, -5):
        vals = (rng.uniform(-1, 1, 100) + (1j if t in "cz" else 0) * rng.uniform(-1, 1, 100)).astype(NP[t])
        expect = buf.arr[:m].copy()
        k = len(np.diagonal(expect, d))
        rr = np.arange(k) + max(-d, 0); cc = np.arange(k) + max(d, 0)
        expect[rr, cc] = vals[:k]
        for src in (vals.ctypes.data, None):
            dv = DevBuf(vals.copy()) if src is None else None
            check(getattr(lib, f"jb_{t}set_diag")(m, n, d, dv.ptr if dv else src, 100, buf.ptr, ld), "set_diag")
            check(lib.jb_sync(), "sync"); buf.back()
            np.testing.assert_array_equal(buf.arr[:m], expect)
            if dv:
                dv.free()
    # matrixMultDiag (Matrix.hs:660-665): op(A) * diag(d)
    A = rand(rng, t, m, n, ld)
    dA = DevBuf(A.copy(order="F"))
    for trans in (111, 112, 113):
        r, c = (m, n) if trans == 111 else (n, m)
        dvec = (rng.uniform(-2, 2, c) + (1j if t in "cz" else 0) * rng.uniform(-2, 2, c)).astype(NP[t])
        out = DevBuf(np.full((r + 1, c), np.nan, dtype=NP[t], order="F"))
        check(getattr(lib, f"jb_{t}mult_diag")(trans, r, c, dA.ptr, ld, dvec.ctypes.data, out.ptr, r + 1), "mult_diag")
        check(lib.jb_sync(), "sync"); out.back()
        opA = A[:m] if trans == 111 else (A[:m].T if trans == 112 else A[:m].conj().T)
        ref = (opA * dvec[None, :]).astype(NP[t])
        assert fro(out.arr[:r] - ref) <= 4 * EPS[t] * fro(ref)
        assert np.isnan(out.arr[r:]).all()
        out.free()
    dA.free(); buf.free()


def test_solveLinearSystem_symbol_sequence_on_device_pointers(blas, rng):
    """What the UNMODIFIED Jalla solveLinearSystem issues for a GPU-resident CMatrix (Matrix.hs:513-521):
    matrixCopy b -> x and matrixCopy a -> a' are zipWithM_ unsafeCopyVector over rowsRef (Matrix.hs:855-871): one
    cblas_dcopy per ROW with the source / destination strides of a column-major matrix (inc = lda), then
    LAPACKE_dgesv(ColumnMajor) with a host ipiv (allocaArray, Matrix.hs:502).  Every pointer below is jb_malloc'd
    device memory; the result must equal OpenBLAS on the host."""
    lib.jb_init(-1)
    n, nrhs = 96, 5
    A = np.asfortranarray(rng.uniform(-1, 1, (n, n))); B = np.asfortranarray(rng.uniform(-1, 1, (n, nrhs)))
    dA, dB = DevBuf(A.copy(order="F")), DevBuf(B.copy(order="F"))
    dA2, dX = DevBuf(np.full((n, n), np.nan, order="F")), DevBuf(np.full((n, nrhs), np.nan, order="F"))
    for i in range(n):                                   # matrixCopy b NoTrans: row i of B -> row i of X
        lib.cblas_dcopy(nrhs, dB.ptr + 8 * i, n, dX.ptr + 8 * i, n)
    for i in range(n):                                   # matrixCopy a NoTrans
        lib.cblas_dcopy(n, dA.ptr + 8 * i, n, dA2.ptr + 8 * i, n)
    check_last("cblas_dcopy")
    ipiv = np.zeros(n, np.int32)
    assert lib.LAPACKE_dgesv(102, n, nrhs, dA2.ptr, n, ipiv.ctypes.data, dX.ptr, n) == 0
    check(lib.jb_sync(), "sync")
    dX.back(); dA.back(); dB.back()
    np.testing.assert_array_equal(dA.arr, A); np.testing.assert_array_equal(dB.arr, B)     # inputs untouched
    Ar, Xr, pr = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
    assert blas.gesv("d", 102, n, nrhs, Ar, n, pr, Xr, n) == 0
    np.testing.assert_array_equal(ipiv, pr)
    assert fro(dX.arr - Xr) <= 100 * n * EPS["d"] * fro(Xr)
    # matrixCopy with Trans copies COLUMNS of the source into rows of the destination (Matrix.hs:867)
    dT = DevBuf(np.full((n, n), np.nan, order="F"))
    for j in range(n):
        lib.cblas_dcopy(n, dA.ptr + 8 * n * j, 1, dT.ptr + 8 * j, n)
    check_last("cblas_dcopy"); check(lib.jb_sync(), "sync"); dT.back()
    np.testing.assert_array_equal(dT.arr, A.T)
    for b in (dA, dB, dA2, dX, dT):
        b.free()
