"""Factorise-and-solve parity: LAPACKE_?{gesv,getrf,getrs,getri,potrf,potrs,posv} of
libjalla_b200 against the CPU oracle (and OpenBLAS) on identical inputs.

The reference's own tests never call this path ("parity unpinned"), so the
pins are north_star's criteria: pivot vectors bit-identical on inputs without near-ties,
|PA-LU| / (n*eps*|A|) <= 10, plus LAPACKE's info conventions (zero pivot -> +k, NaN input
-> -(argument index), bad argument -> -(argument index)) and Jalla's calling convention
(Matrix.hs:495-508: order 102, host ipiv from allocaArray, nrhs = cols of B)."""
import numpy as np
import pytest

from tests.util import NP, EPS, rand, fro

pytestmark = pytest.mark.gpu
TYPES = "sdcz"


def lu_residual(t, A, LU, ipiv):
    m, n = A.shape
    mn = min(m, n)
    L = np.tril(LU[:, :mn], -1) + np.eye(m, mn, dtype=LU.dtype)
    U = np.triu(LU[:mn, :])
    PA = A.copy()
    for i, p in enumerate(ipiv):
        p = int(p) - 1
        if p != i:
            PA[[i, p], :] = PA[[p, i], :]
    return fro(PA - L @ U) / (max(m, n) * EPS[t] * fro(A))


def no_near_ties(orc, t, A, thresh=1e-6):
    """True when at every elimination step the runner-up candidate is clearly below the pivot
    (the hypothesis under which pivots must be bit-identical: north_star)."""
    LU = A.copy(order="F")
    m, n = LU.shape
    for j in range(min(m, n)):
        col = np.abs(LU[j:, j].real) + np.abs(LU[j:, j].imag) if t in "cz" else np.abs(LU[j:, j])
        o = np.sort(col)[::-1]
        if len(o) > 1 and o[0] > 0 and (o[0] - o[1]) <= thresh * o[0]:
            return False
        p = j + int(np.argmax(col))
        LU[[j, p], :] = LU[[p, j], :]
        if LU[j, j] != 0:
            LU[j + 1:, j] /= LU[j, j]
            LU[j + 1:, j + 1:] -= np.outer(LU[j + 1:, j], LU[j, j + 1:])
    return True


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("shape", [(1, 1), (7, 7), (32, 32), (33, 33), (100, 100), (64, 40), (40, 64), (257, 257)])
def test_getrf(gpu, orc, blas, rng, t, shape):
    m, n = shape
    A = rand(rng, t, m, n, m + 3)
    Ag, Ao, Ab = (A.copy(order="F") for _ in range(3))
    pg, po, pb = (np.zeros(min(m, n), np.int32) for _ in range(3))
    ig = gpu.getrf(t, 102, m, n, Ag, m + 3, pg)
    io = orc.getrf(t, 102, m, n, Ao, m + 3, po)
    ib = blas.getrf(t, 102, m, n, Ab, m + 3, pb)
    assert ig == io == ib == 0
    assert np.isnan(Ag[m:, :]).all()
    np.testing.assert_array_equal(po, pb)          # the restatement is pinned to OpenBLAS
    if no_near_ties(orc, t, A[:m, :]):
        np.testing.assert_array_equal(pg, po)      # bit-identical pivots
    assert lu_residual(t, A[:m, :], Ag[:m, :], pg) <= 10
    scale = fro(Ao[:m, :])
    assert fro(Ag[:m, :] - Ao[:m, :]) / scale <= 200 * max(m, n) * EPS[t]


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("n,nrhs", [(5, 1), (64, 3), (200, 1), (512, 1), (512, 64)])
def test_gesv_jalla_convention(gpu, orc, rng, t, n, nrhs):
    """unsafeSolveLinearSystem (Matrix.hs:495-508): order 102, host ipiv, X overwrites B."""
    if t in "cz" and n == 512:
        n = 256
    A = rand(rng, t, n, n); B = rand(rng, t, n, nrhs)
    Ag, Ao, Bg, Bo = A.copy(order="F"), A.copy(order="F"), B.copy(order="F"), B.copy(order="F")
    pg, po = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert gpu.gesv(t, 102, n, nrhs, Ag, n, pg, Bg, n) == 0
    assert orc.gesv(t, 102, n, nrhs, Ao, n, po, Bo, n) == 0
    if no_near_ties(orc, t, A):
        np.testing.assert_array_equal(pg, po)
    assert lu_residual(t, A, Ag, pg) <= 10
    # backward error of the solve: |A x - b| / (|A||x| n eps)
    res = fro(A.astype(np.complex128) @ Bg.astype(np.complex128) - B) / (fro(A) * fro(Bg) * n * EPS[t])
    assert res <= 10


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("trans", ["N", "T", "C"])
def test_getrs_all_trans(gpu, orc, rng, t, trans):
    n, nrhs = 97, 5
    A = rand(rng, t, n, n); B = rand(rng, t, n, nrhs, n + 2)
    LU = A.copy(order="F"); piv = np.zeros(n, np.int32)
    assert orc.getrf(t, 102, n, n, LU, n, piv) == 0
    Bg, Bo = B.copy(order="F"), B.copy(order="F")
    assert gpu.getrs(t, 102, trans, n, nrhs, LU, n, piv, Bg, n + 2) == 0
    assert orc.getrs(t, 102, trans, n, nrhs, LU, n, piv, Bo, n + 2) == 0
    assert fro(Bg[:n] - Bo[:n]) / fro(Bo[:n]) <= 1e3 * n * EPS[t]
    assert np.isnan(Bg[n:, :]).all()


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("n,nrhs", [(1, 1), (33, 2), (700, 1), (1100, 16)])
def test_getrs_fused_and_blocked_paths(gpu, orc, rng, t, n, nrhs):
    """Few right-hand sides take the one-CTA-per-column kernel, the rest the blocked trsm path: both against the oracle."""
    from jalla_b200._lib import lib
    if t in "cz" and n > 700:
        n = 600
    A = rand(rng, t, n, n); B = rand(rng, t, n, nrhs, n + 3)
    LU = A.copy(order="F"); piv = np.zeros(n, np.int32)
    assert orc.getrf(t, 102, n, n, LU, n, piv) == 0
    Bo = B.copy(order="F")
    assert orc.getrs(t, 102, "N", n, nrhs, LU, n, piv, Bo, n + 3) == 0
    for fused in (None, 0):
        Bg = B.copy(order="F")
        if fused is not None:
            lib.jb_set_option(b"getrs.fused_max_nrhs", fused)
        try:
            assert gpu.getrs(t, 102, "N", n, nrhs, LU, n, piv, Bg, n + 3) == 0
        finally:
            lib.jb_set_option(b"getrs.fused_max_nrhs", -1)
        assert fro(Bg[:n] - Bo[:n]) / fro(Bo[:n]) <= 1e3 * n * EPS[t]
        assert np.isnan(Bg[n:, :]).all()


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("n", [1, 31, 128, 300])
def test_getri_invert(gpu, orc, rng, t, n):
    """unsafeInvert (Matrix.hs:763-776): getrf then getri."""
    A = rand(rng, t, n, n)
    Ag = A.copy(order="F"); pg = np.zeros(n, np.int32)
    assert gpu.getrf(t, 102, n, n, Ag, n, pg) == 0
    assert gpu.getri(t, 102, n, Ag, n, pg) == 0
    Ao = A.copy(order="F"); po = np.zeros(n, np.int32)
    assert orc.getrf(t, 102, n, n, Ao, n, po) == 0 and orc.getri(t, 102, n, Ao, n, po) == 0
    I = np.eye(n)
    cond = fro(A) * fro(Ao)
    assert fro(A.astype(np.complex128) @ Ag.astype(np.complex128) - I) <= 10 * n * EPS[t] * cond
    assert fro(Ag - Ao) / fro(Ao) <= 10 * n * EPS[t] * cond


def spd(rng, t, n, ld=None):
    M = rand(rng, t, n, n)
    S = (M @ M.conj().T + n * np.eye(n)).astype(NP[t])
    out = np.full((ld or n, n), np.nan, dtype=NP[t], order="F")
    out[:n, :] = S
    return out


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n", [1, 17, 32, 100, 259])
def test_potrf_potrs(gpu, orc, rng, t, uplo, n):
    ld = n + 1
    S = spd(rng, t, n, ld)
    # poison the triangle LAPACK says is not referenced
    Sp = S.copy(order="F")
    idx = np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)
    Sp[:n, :][idx] = np.nan
    Sg, So = Sp.copy(order="F"), Sp.copy(order="F")
    assert gpu.potrf(t, 102, uplo, n, Sg, ld) == 0
    assert orc.potrf(t, 102, uplo, n, So, ld) == 0
    F = np.tril(np.nan_to_num(Sg[:n, :])) if uplo == "L" else np.triu(np.nan_to_num(Sg[:n, :]))
    rec = F @ F.conj().T if uplo == "L" else F.conj().T @ F
    assert fro(rec - S[:n, :]) / (n * EPS[t] * fro(S[:n, :])) <= 10
    assert np.isnan(Sg[:n, :][idx]).all(), "the other triangle was touched"
    Fo = np.tril(np.nan_to_num(So[:n, :])) if uplo == "L" else np.triu(np.nan_to_num(So[:n, :]))
    assert fro(F - Fo) / fro(Fo) <= 100 * n * EPS[t]
    B = rand(rng, t, n, 4)
    Bg, Bo = B.copy(order="F"), B.copy(order="F")
    assert gpu.potrs(t, 102, uplo, n, 4, Sg, ld, Bg, n) == 0
    assert orc.potrs(t, 102, uplo, n, 4, So, ld, Bo, n) == 0
    assert fro(Bg - Bo) / fro(Bo) <= 1e3 * n * EPS[t]


@pytest.mark.parametrize("t", TYPES)
def test_posv_and_row_major(gpu, orc, rng, t):
    n, nrhs = 75, 3
    S = spd(rng, t, n); B = rand(rng, t, n, nrhs)
    for order in (102, 101):
        for uplo in "LU":
            # row-major storage of S is the column-major storage of S^T; B row-major is nrhs x n storage
            Ss = S.copy(order="F") if order == 102 else np.asfortranarray(S.T)
            Bs = B.copy(order="F") if order == 102 else np.asfortranarray(B.T)
            ldb = n if order == 102 else nrhs
            Sg, So, Bg, Bo = Ss.copy(order="F"), Ss.copy(order="F"), Bs.copy(order="F"), Bs.copy(order="F")
            assert gpu.posv(t, order, uplo, n, nrhs, Sg, n, Bg, ldb) == 0
            assert orc.posv(t, order, uplo, n, nrhs, So, n, Bo, ldb) == 0
            assert fro(Bg - Bo) / fro(Bo) <= 1e3 * n * EPS[t]
            keep = np.tril_indices(n) if (uplo == "L") == (order == 102) else np.triu_indices(n)
            assert fro(Sg[keep] - So[keep]) / fro(So[keep]) <= 100 * n * EPS[t]


@pytest.mark.parametrize("t", TYPES)
def test_row_major_lu(gpu, orc, rng, t):
    m, n, nrhs = 45, 45, 2
    A = rand(rng, t, m, n); B = rand(rng, t, n, nrhs)
    As, Bs = np.asfortranarray(A.T), np.asfortranarray(B.T)
    Ag, Ao, Bg, Bo = As.copy(order="F"), As.copy(order="F"), Bs.copy(order="F"), Bs.copy(order="F")
    pg, po = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert gpu.gesv(t, 101, n, nrhs, Ag, n, pg, Bg, nrhs) == 0
    assert orc.gesv(t, 101, n, nrhs, Ao, n, po, Bo, nrhs) == 0
    np.testing.assert_array_equal(pg, po)
    assert fro(Bg - Bo) / fro(Bo) <= 1e3 * n * EPS[t]
    assert fro(Ag - Ao) / fro(Ao) <= 100 * n * EPS[t]


@pytest.mark.parametrize("t", "sd")
def test_info_codes(gpu, orc, rng, t):
    n = 40
    # exact zero pivot: column 5 zero after elimination -> info = k (1-based), LAPACK continues
    A = rand(rng, t, n, n)
    A[:, 7] = 0
    Ag, Ao = A.copy(order="F"), A.copy(order="F")
    pg, po = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert gpu.getrf(t, 102, n, n, Ag, n, pg) == orc.getrf(t, 102, n, n, Ao, n, po) == 8
    np.testing.assert_array_equal(pg, po)
    # gesv on a singular matrix: info > 0 and B untouched
    B = rand(rng, t, n, 2); Bg = B.copy(order="F")
    assert gpu.gesv(t, 102, n, 2, A.copy(order="F"), n, pg, Bg, n) == 8
    np.testing.assert_array_equal(Bg, B)
    # getri on singular U
    assert gpu.getri(t, 102, n, Ag, n, pg) == orc.getri(t, 102, n, Ao, n, po) == 8
    # NaN screening: A -> -4, B -> -7 for gesv (verified on OpenBLAS)
    An = rand(rng, t, n, n); An[3, 9] = np.nan
    assert gpu.gesv(t, 102, n, 2, An.copy(order="F"), n, pg, B.copy(order="F"), n) == -4
    Bn = B.copy(order="F"); Bn[1, 1] = np.nan
    assert gpu.gesv(t, 102, n, 2, rand(rng, t, n, n), n, pg, Bn, n) == -7
    assert gpu.getrf(t, 102, n, n, An.copy(order="F"), n, pg) == orc.getrf(t, 102, n, n, An.copy(order="F"), n, po) == -4
    # not positive definite: info = order of the failing leading minor
    S = spd(rng, t, n); S[12, 12] = -1.0
    assert gpu.potrf(t, 102, "L", n, S.copy(order="F"), n) == orc.potrf(t, 102, "L", n, S.copy(order="F"), n) == 13
    assert gpu.potrf(t, 102, "U", n, S.copy(order="F"), n) == 13
    # argument errors
    good = rand(rng, t, n, n)
    assert gpu.getrf(t, 100, n, n, good.copy(order="F"), n, pg) == -1
    assert gpu.getrf(t, 102, n, n, good.copy(order="F"), n - 1, pg) == orc.getrf(t, 102, n, n, good.copy(order="F"), n - 1, po) == -5
    assert gpu.gesv(t, 102, n, 2, good.copy(order="F"), n, pg, B.copy(order="F"), n - 1) == -8
    assert gpu.potrf(t, 102, "X", n, good.copy(order="F"), n) == -2
    assert gpu.getrs(t, 102, "Q", n, 2, good, n, pg, B.copy(order="F"), n) == -2


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("order", [102, 101])
def test_gesv_augmented_matrix_path(gpu, orc, rng, t, order):
    """?gesv with <= 8 right-hand sides and 128 <= n <= 2048 factorises [A | B] (the forward substitution rides in getrf's
    launches): factors and pivots identical to the plain getrf + getrs route (gesv.augment = 0), solution against the oracle,
    padded leading dimensions untouched, and a singular matrix returns info > 0 with the factors stored and B untouched."""
    from jalla_b200._lib import lib
    n, nrhs = 300, 3
    A = rand(rng, t, n, n, n + 2); B = rand(rng, t, n, nrhs, n + 3)
    if order == 101:
        A = rand(rng, t, n, n, n); B = rand(rng, t, nrhs, n, nrhs)      # row-major n x nrhs stored as its transpose
    lda, ldb = (n + 2, n + 3) if order == 102 else (n, nrhs)
    res = []
    for aug in (-1, 0):
        lib.jb_set_option(b"gesv.augment", aug)
        try:
            Ag, Bg, pg = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
            assert gpu.gesv(t, order, n, nrhs, Ag, lda, pg, Bg, ldb) == 0
        finally:
            lib.jb_set_option(b"gesv.augment", -1)
        res.append((Ag, Bg, pg))
    np.testing.assert_array_equal(res[0][2], res[1][2])
    if order == 102:
        np.testing.assert_array_equal(res[0][0][:n], res[1][0][:n])            # the same factorisation, bit for bit
        assert np.isnan(res[0][0][n:]).all() and np.isnan(res[0][1][n:]).all()
    Ao, Bo, po = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
    assert orc.gesv(t, order, n, nrhs, Ao, lda, po, Bo, ldb) == 0
    rows = slice(0, n) if order == 102 else slice(None)
    assert fro(res[0][1][rows] - Bo[rows]) / fro(Bo[rows]) <= 1e3 * n * EPS[t]
    # singular: column 100 zero -> info = 101, B untouched, factors as getrf leaves them
    if order == 102:
        As = rand(rng, t, n, n); As[:, 100] = 0
        Ag, Bg, pg = As.copy(order="F"), rand(rng, t, n, 1), np.zeros(n, np.int32)
        B0 = Bg.copy(order="F")
        assert gpu.gesv(t, 102, n, 1, Ag, n, pg, Bg, n) == 101
        np.testing.assert_array_equal(Bg, B0)
        Af, pf = As.copy(order="F"), np.zeros(n, np.int32)
        assert gpu.getrf(t, 102, n, n, Af, n, pf) == 101
        np.testing.assert_array_equal(pg, pf); np.testing.assert_array_equal(Ag, Af)


def test_host_pointer_solve(gpu_host, orc, rng):
    """Stock Jalla `Matrix` lives in pageable host memory: the link-level drop-in stages it."""
    n = 130
    A = rand(rng, "d", n, n); B = rand(rng, "d", n, 2)
    Ag, Ao, Bg, Bo = A.copy(order="F"), A.copy(order="F"), B.copy(order="F"), B.copy(order="F")
    pg, po = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert gpu_host.gesv("d", 102, n, 2, Ag, n, pg, Bg, n) == 0
    assert orc.gesv("d", 102, n, 2, Ao, n, po, Bo, n) == 0
    np.testing.assert_array_equal(pg, po)
    assert fro(Bg - Bo) / fro(Bo) <= 1e3 * n * EPS["d"]


def test_host_pointer_solve_staged(gpu_host, blas, rng):
    """Pageable host operands of 16 MB and more go through the library's own pinned-slot staging (csrc/host_stage.cu:
    several host threads fill / drain the slots, D2H delivered by a drain thread before the call returns): padded leading
    dimensions whose NaN padding must survive, pivots and solution against OpenBLAS, a second call re-using the slots."""
    n, nrhs, lda, ldb = 1800, 1200, 1803, 1801           # A 26 MB, B 17 MB
    for rep in range(2):
        A = rand(rng, "d", n, n, lda); B = rand(rng, "d", n, nrhs, ldb)
        Ag, Ao, Bg, Bo = A.copy(order="F"), A.copy(order="F"), B.copy(order="F"), B.copy(order="F")
        pg, po = np.zeros(n, np.int32), np.zeros(n, np.int32)
        assert gpu_host.gesv("d", 102, n, nrhs, Ag, lda, pg, Bg, ldb) == 0
        assert blas.gesv("d", 102, n, nrhs, Ao, lda, po, Bo, ldb) == 0
        np.testing.assert_array_equal(pg, po)
        assert fro(Bg[:n] - Bo[:n]) / fro(Bo[:n]) <= 1e3 * n * EPS["d"]
        assert fro(Ag[:n] - Ao[:n]) / fro(Ao[:n]) <= 50 * n * EPS["d"]
        assert np.isnan(Ag[n:]).all() and np.isnan(Bg[n:]).all()


def test_concurrent_callers_from_several_os_threads(gpu_host, blas, rng):
    """`foreign import ccall unsafe` may enter the library from several OS threads at once (SURVEY.md §8b "Threading",
    jalla.cabal:95 -threaded): four Python threads (ctypes releases the GIL) each run gesv on their own pageable host
    operands — small ones on the driver's staging, large ones through the shared pinned-slot stager — and every result
    must match OpenBLAS."""
    import threading
    jobs = []
    for i, (n, nrhs) in enumerate([(1500, 1400), (300, 2), (1500, 1400), (700, 1)]):
        A = rand(rng, "d", n, n); B = rand(rng, "d", n, nrhs)
        jobs.append((n, nrhs, A, B))
    out = [None] * len(jobs)

    def run(i):
        n, nrhs, A, B = jobs[i]
        Ag, Bg, pg = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
        info = gpu_host.gesv("d", 102, n, nrhs, Ag, n, pg, Bg, n)
        out[i] = (info, pg, Bg)
    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    for t in th: t.start()
    for t in th: t.join()
    for (n, nrhs, A, B), (info, pg, Bg) in zip(jobs, out):
        Ao, Bo, po = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
        assert info == 0 and blas.gesv("d", 102, n, nrhs, Ao, n, po, Bo, n) == 0
        np.testing.assert_array_equal(pg, po)
        assert fro(Bg - Bo) / fro(Bo) <= 1e3 * n * EPS["d"]


@pytest.mark.parametrize("t", "dz")
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_potrf_two_level(gpu, blas, rng, t, uplo):
    """n large enough for the two-level (NB2 = 256 outer, 32 inner) Cholesky; ragged on both levels."""
    from jalla_b200._lib import lib
    n = 701
    S = spd(rng, t, n)
    Sg, Sb = S.copy(order="F"), S.copy(order="F")
    lib.jb_set_option(b"potrf.nb2", 256)      # force the two-level path at a test-sized n
    try:
        assert gpu.potrf(t, 102, uplo, n, Sg, n) == 0
        Sbad = S.copy(order="F"); Sbad[450, 450] = -5.0
        assert gpu.potrf(t, 102, uplo, n, Sbad, n) == 451
    finally:
        lib.jb_set_option(b"potrf.nb2", 0)
    assert blas.potrf(t, 102, uplo, n, Sb, n) == 0
    tri = np.tril_indices(n) if uplo == "L" else np.triu_indices(n)
    other = np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)
    assert fro(Sg[tri] - Sb[tri]) / fro(Sb[tri]) <= 100 * n * EPS[t]
    np.testing.assert_array_equal(Sg[other], S[other])
    F = np.tril(Sg) if uplo == "L" else np.triu(Sg)
    rec = F @ F.conj().T if uplo == "L" else F.conj().T @ F
    assert fro(rec - S) / (n * EPS[t] * fro(S)) <= 10
    Sbad2 = S.copy(order="F"); Sbad2[450, 450] = -5.0
    assert blas.potrf(t, 102, uplo, n, Sbad2, n) == 451


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("shape", [(300, 300), (513, 200), (200, 513), (1500, 1500)])
def test_getrf_cooperative_panel(gpu, blas, rng, t, shape):
    """Tall panels run on the cooperative multi-CTA panel kernel (one grid barrier per column); forced on at
    small sizes through the option, natural at 1500.  Pivots must equal OpenBLAS's, residual within bound."""
    from jalla_b200._lib import lib
    m, n = shape
    if t in "cz" and m == 1500:
        m = n = 1100
    A = rand(rng, t, m, n, m + 1)
    A[:m, :] += 0.01 * np.arange(m * n).reshape(m, n).astype(NP[t]) / (m * n)     # breaks exact ties
    Ag, Ab = A.copy(order="F"), A.copy(order="F")
    pg, pb = np.zeros(min(m, n), np.int32), np.zeros(min(m, n), np.int32)
    lib.jb_set_option(b"getrf.coop_min_rows", 64)
    try:
        assert gpu.getrf(t, 102, m, n, Ag, m + 1, pg) == 0
    finally:
        lib.jb_set_option(b"getrf.coop_min_rows", 0)
    assert blas.getrf(t, 102, m, n, Ab, m + 1, pb) == 0
    if no_near_ties(None, t, A[:m, :], thresh=1e-5 if t in "sc" else 1e-9):
        np.testing.assert_array_equal(pg, pb)
    assert lu_residual(t, A[:m, :], Ag[:m, :], pg) <= 10
    # singular column and ties behave like the single-CTA kernel
    if t == "d" and m == 300:
        Z = rand(rng, t, m, n); Z[:, 40] = 0
        Zg, Zb = Z.copy(order="F"), Z.copy(order="F")
        lib.jb_set_option(b"getrf.coop_min_rows", 64)
        try:
            ig = gpu.getrf(t, 102, m, n, Zg, m, pg)
        finally:
            lib.jb_set_option(b"getrf.coop_min_rows", 0)
        assert ig == blas.getrf(t, 102, m, n, Zb, m, pb) == 41
        np.testing.assert_array_equal(pg, pb)
        I = np.asfortranarray(rng.integers(-2, 3, (m, n)).astype(np.float64) + 4 * np.eye(m))
        Ig, Ib = I.copy(order="F"), I.copy(order="F")
        lib.jb_set_option(b"getrf.coop_min_rows", 64)
        try:
            ig = gpu.getrf(t, 102, m, n, Ig, m, pg)
        finally:
            lib.jb_set_option(b"getrf.coop_min_rows", 0)
        assert ig == blas.getrf(t, 102, m, n, Ib, m, pb)
        np.testing.assert_array_equal(pg[:32], pb[:32])   # first panel: identical arithmetic, exact ties -> lowest index


@pytest.mark.parametrize("t", TYPES)
def test_getrf_two_level(gpu, blas, rng, t):
    """Outer panels of NB2 columns (forced to 64 here) factorised by the NB=32 level, then one rank-NB2 update."""
    from jalla_b200._lib import lib
    for (m, n) in ((333, 333), (400, 250), (250, 400)):
        A = rand(rng, t, m, n, m + 2)
        Ag, Ab = A.copy(order="F"), A.copy(order="F")
        pg, pb = np.zeros(min(m, n), np.int32), np.zeros(min(m, n), np.int32)
        lib.jb_set_option(b"getrf.nb2", 64); lib.jb_set_option(b"getrf.coop_min_rows", 128)
        try:
            assert gpu.getrf(t, 102, m, n, Ag, m + 2, pg) == 0
        finally:
            lib.jb_set_option(b"getrf.nb2", 0); lib.jb_set_option(b"getrf.coop_min_rows", 0)
        assert blas.getrf(t, 102, m, n, Ab, m + 2, pb) == 0
        if no_near_ties(None, t, A[:m, :], thresh=1e-4 if t in "sc" else 1e-9):
            np.testing.assert_array_equal(pg, pb)
        assert lu_residual(t, A[:m, :], Ag[:m, :], pg) <= 10
        assert np.isnan(Ag[m:, :]).all()


def test_dgesv_large_natural_path(gpu, blas, rng):
    """n = 4200: the default configuration takes the two-level + cooperative-panel path."""
    n = 4200
    A = rand(rng, "d", n, n); B = rand(rng, "d", n, 2)
    Ag, Ab, Bg, Bb = A.copy(order="F"), A.copy(order="F"), B.copy(order="F"), B.copy(order="F")
    pg, pb = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert gpu.gesv("d", 102, n, 2, Ag, n, pg, Bg, n) == 0
    assert blas.gesv("d", 102, n, 2, Ab, n, pb, Bb, n) == 0
    np.testing.assert_array_equal(pg, pb)
    assert fro(A @ Bg - B) / (fro(A) * fro(Bg) * n * EPS["d"]) <= 10
    assert fro(Ag - Ab) / fro(Ab) <= 1e3 * n * EPS["d"]
