"""GEMM parity: libjalla_b200's cblas_?gemm against the CPU oracle on identical inputs.

Restates the reference's own GEMM pins — tests/Test.hs:77-103 (prop_matrixMultDiag1-4:
NN and TN products against a diagonal matrix must be *exact*) with the generator of
Numeric/Jalla/Test.hs:22-41 — and adds the north_star criterion
(rel. Frobenius error <= 4*n*eps*|A||B|/|C|), Jalla's calling convention
(Matrix.hs:477-490: order 102, beta = 0 into uninitialised memory), every transpose
combination, ragged shapes, padded leading dimensions, row-major order and empty shapes."""
import itertools

import numpy as np
import pytest

from tests.util import NP, EPS, rand, jalla_arbitrary, fro, op, gemm_bound

pytestmark = pytest.mark.gpu
TYPES = "sdcz"


@pytest.fixture
def force_dmma():
    """Small shapes normally go to the finer-grained generic kernel; these tests want the tuned DMMA kernels' tile
    edges, so they pin the dispatch."""
    from jalla_b200._lib import lib
    lib.jb_set_option(b"gemm.force_dmma", 1)
    yield
    lib.jb_set_option(b"gemm.force_dmma", 0)


def run_both(gpu, orc, t, ta, tb, m, n, k, alpha, beta, rng, pad=(0, 0, 0), order=102, poison=True):
    """Builds op-shaped operands, runs GPU and oracle, returns (Cgpu, Cref, A, B) logical views."""
    ra, ca = (m, k) if ta == 111 else (k, m)
    rb, cb = (k, n) if tb == 111 else (n, k)
    if order == 101:   # row-major storage == column-major storage of the transpose
        A = rand(rng, t, ca, ra, ca + pad[0]); B = rand(rng, t, cb, rb, cb + pad[1])
        C0 = rand(rng, t, n, m, n + pad[2])
    else:
        A = rand(rng, t, ra, ca, ra + pad[0]); B = rand(rng, t, rb, cb, rb + pad[1])
        C0 = rand(rng, t, m, n, m + pad[2])
    if beta == 0 and poison:
        C0[...] = np.nan        # beta == 0 must overwrite, never read, C (Matrix.hs:458-459)
    Cg, Cr = C0.copy(order="F"), C0.copy(order="F")
    lda, ldb, ldc = A.shape[0], B.shape[0], C0.shape[0]
    gpu.gemm(t, order, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cg, ldc)
    orc.gemm(t, order, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cr, ldc)
    if order == 101:
        return Cg[:n, :].T, Cr[:n, :].T, op(A[:ca, :].T, ta), op(B[:cb, :].T, tb), Cg, Cr
    return Cg[:m, :], Cr[:m, :], op(A[:ra, :], ta), op(B[:rb, :], tb), Cg, Cr


def assert_close(t, k, Cg, Cr, A, B, extra=1.0):
    assert np.isfinite(Cg).all(), "non-finite values leaked into C"
    err = fro(Cg - Cr) / max(fro(Cr), 1e-300)
    assert err <= extra * gemm_bound(t, k, A, B, Cr), (err, gemm_bound(t, k, A, B, Cr))


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("ta,tb", list(itertools.product((111, 112), (111, 112))))
def test_jalla_shapes_all_transposes(gpu, orc, rng, t, ta, tb):
    """Shapes from the reference generator (1..100), Jalla's alpha=1, beta=0."""
    for _ in range(6):
        m, n, k = (int(x) for x in rng.integers(1, 101, 3))
        Cg, Cr, A, B, *_ = run_both(gpu, orc, t, ta, tb, m, n, k, 1.0, 0.0, rng)
        assert_close(t, k, Cg, Cr, A, B)


@pytest.mark.parametrize("t", "cz")
@pytest.mark.parametrize("ta,tb", [(113, 111), (111, 113), (113, 113), (112, 113)])
def test_conj_transpose(gpu, orc, rng, t, ta, tb):
    Cg, Cr, A, B, *_ = run_both(gpu, orc, t, ta, tb, 45, 37, 53, 1.0, 0.0, rng)
    assert_close(t, 53, Cg, Cr, A, B)


@pytest.mark.parametrize("t", TYPES)
def test_alpha_beta_and_padding(gpu, orc, rng, t):
    alpha = 0.75 if t in "sd" else 0.75 - 0.5j
    beta = -1.25 if t in "sd" else -1.25 + 0.25j
    for ta, tb in itertools.product((111, 112), (111, 112)):
        Cg, Cr, A, B, Cgs, Crs = run_both(gpu, orc, t, ta, tb, 70, 33, 91, alpha, beta, rng, pad=(3, 5, 7))
        bound = gemm_bound(t, 91, A, B, Cr) + 4 * EPS[t]
        assert fro(Cg - Cr) / fro(Cr) <= bound
        assert np.isnan(Cgs[70:, :]).all(), "rows beyond m (ldc padding) were written"


@pytest.mark.parametrize("t", TYPES)
def test_row_major(gpu, orc, rng, t):
    for ta, tb in [(111, 111), (112, 111), (111, 112)]:
        Cg, Cr, A, B, *_ = run_both(gpu, orc, t, ta, tb, 41, 67, 29, 1.0, 0.0, rng, order=101, pad=(2, 0, 4))
        assert_close(t, 29, Cg, Cr, A, B)


@pytest.mark.parametrize("t", TYPES)
def test_degenerate_shapes(gpu, orc, rng, t):
    """m, n or k == 0 and alpha == 0 follow BLAS: C <- beta*C, A/B not referenced."""
    for (m, n, k) in [(0, 5, 3), (5, 0, 3), (4, 6, 0)]:
        A = rand(rng, t, max(m, 1), max(k, 1)); B = rand(rng, t, max(k, 1), max(n, 1))
        C0 = rand(rng, t, max(m, 1), max(n, 1))
        Cg, Cr = C0.copy(order="F"), C0.copy(order="F")
        gpu.gemm(t, 102, 111, 111, m, n, k, 1.0, A, A.shape[0], B, B.shape[0], 2.0, Cg, C0.shape[0])
        orc.gemm(t, 102, 111, 111, m, n, k, 1.0, A, A.shape[0], B, B.shape[0], 2.0, Cr, C0.shape[0])
        np.testing.assert_array_equal(Cg, Cr)
    A = rand(rng, t, 9, 7); B = rand(rng, t, 7, 8)
    A[2, 3] = np.nan                      # alpha == 0: NaNs in A must not reach C
    C0 = rand(rng, t, 9, 8); Cg = C0.copy(order="F")
    gpu.gemm(t, 102, 111, 111, 9, 8, 7, 0.0, A, 9, B, 7, 1.0, Cg, 9)
    np.testing.assert_array_equal(Cg, C0)
    gpu.gemm(t, 102, 111, 111, 9, 8, 7, 0.0, A, 9, B, 7, 0.0, Cg, 9)
    assert (Cg == 0).all()


def _diag_matrix(t, rows, cols):
    d = np.zeros((rows, cols), dtype=NP[t], order="F")
    for i in range(min(rows, cols)):
        d[i, i] = i + 1
    return d


def test_config1_512_default_dispatch(gpu, orc, rng):
    """The same 512^3 Double product on whatever kernel the dispatcher picks (the finer-grained one at this size)."""
    Cg, Cr, A, B, *_ = run_both(gpu, orc, "d", 111, 111, 512, 512, 512, 1.0, 0.0, rng)
    assert_close("d", 512, Cg, Cr, A, B)


@pytest.mark.parametrize("t", TYPES)
def test_prop_matrixMultDiag(gpu, rng, t):
    """tests/Test.hs:77-94 restated: scaling the columns of `mat` by [1,2,..] equals
    mat ## dm and (mat,Trans) ##! (dm',NoTrans) — exactly, because each dot product has a
    single non-zero term with an integer factor (a 3M complex scheme fails this; 4M passes)."""
    for _ in range(5):
        mat = jalla_arbitrary(rng, t)
        m, n = mat.shape
        dm = _diag_matrix(t, n, m)
        C1 = np.full((m, m), np.nan, dtype=NP[t], order="F")
        gpu.gemm(t, 102, 111, 111, m, m, n, 1.0, mat, m, dm, n, 0.0, C1, m)
        want = np.zeros((m, m), dtype=NP[t])
        w = min(m, n)
        want[:, :w] = mat[:, :w] * np.arange(1, w + 1, dtype=NP[t])
        np.testing.assert_array_equal(C1, want)
        dm2 = _diag_matrix(t, m, n)
        C2 = np.full((n, n), np.nan, dtype=NP[t], order="F")
        gpu.gemm(t, 102, 112, 111, n, n, m, 1.0, mat, m, dm2, m, 0.0, C2, n)
        want2 = np.zeros((n, n), dtype=NP[t])
        want2[:, :w] = mat.T[:, :w] * np.arange(1, w + 1, dtype=NP[t])
        np.testing.assert_array_equal(C2, want2)


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("ta,tb", list(itertools.product((111, 112), (111, 112))))
def test_config1_512(gpu, orc, rng, t, ta, tb, force_dmma):
    """BASELINE.json configs[0]: 512 x 512 multiply (all four transpose combinations; the
    double NN case is the one the reference config names)."""
    n = 512 if t in "sd" else 256
    Cg, Cr, A, B, *_ = run_both(gpu, orc, t, ta, tb, n, n, n, 1.0, 0.0, rng)
    assert_close(t, n, Cg, Cr, A, B)


@pytest.mark.parametrize("ta,tb", list(itertools.product((111, 112), (111, 112))))
def test_dgemm_ragged_large(gpu, blas, rng, ta, tb, force_dmma):
    """Tile-edge coverage of the DMMA kernel: sizes that are not multiples of the 128x128x16
    tile, even leading dimensions (TMA path) with padding; checked against OpenBLAS."""
    m, n, k = 333, 270, 1031
    ra, rb = (m if ta == 111 else k), (k if tb == 111 else n)
    Cg, Cr, A, B, *_ = run_both(gpu, blas, "d", ta, tb, m, n, k, 1.5, -0.5, rng, pad=(2 + ra % 2, 2 + rb % 2, 3))
    assert_close("d", k, Cg, Cr, A, B, extra=1.0)


def test_dgemm_generic_matches_dmma(gpu, rng, force_dmma):
    """The two device paths (tuned DMMA, generic SIMT) agree to rounding on the same input."""
    from jalla_b200._lib import lib
    m, n, k = 384, 256, 320
    A = rand(rng, "d", m, k); B = rand(rng, "d", k, n)
    C1 = np.full((m, n), np.nan, order="F"); C2 = C1.copy(order="F")
    gpu.gemm("d", 102, 111, 111, m, n, k, 1.0, A, m, B, k, 0.0, C1, m)
    lib.jb_set_option(b"gemm.force_generic", 1)
    try:
        gpu.gemm("d", 102, 111, 111, m, n, k, 1.0, A, m, B, k, 0.0, C2, m)
    finally:
        lib.jb_set_option(b"gemm.force_generic", 0)
    assert fro(C1 - C2) / fro(C2) <= 4 * k * EPS["d"]


@pytest.mark.parametrize("t", TYPES)
def test_host_pointers(gpu_host, orc, rng, t):
    """Pageable host operands (stock Jalla `Matrix`: link-level drop-in) are staged."""
    Cg, Cr, A, B, *_ = run_both(gpu_host, orc, t, 111, 112, 130, 75, 66, 1.0, 0.0, rng, pad=(1, 2, 3))
    assert_close(t, 66, Cg, Cr, A, B)


def test_bad_arguments_are_recorded_not_fatal(gpu, rng):
    from jalla_b200._lib import lib, JB_ERR_ARG
    A = rand(rng, "d", 4, 4)
    C0 = A.copy(order="F")
    lib.jb_last_error()
    lib.cblas_dgemm(102, 111, 111, 4, 4, 4, 1.0, A.ctypes.data, 2, A.ctypes.data, 4, 0.0, C0.ctypes.data, 4)  # lda < m
    assert lib.jb_last_error() == JB_ERR_ARG
    lib.cblas_dgemm(103, 111, 111, 4, 4, 4, 1.0, A.ctypes.data, 4, A.ctypes.data, 4, 0.0, C0.ctypes.data, 4)  # bad order
    assert lib.jb_last_error() == JB_ERR_ARG
    np.testing.assert_array_equal(C0, A)


@pytest.mark.parametrize("ta,tb", [(111, 111), (112, 111), (111, 113), (113, 112)])
def test_zgemm_ragged_large(gpu, blas, rng, ta, tb, force_dmma):
    """Tile-edge coverage of the complex DMMA kernel (64x128x8 tiles) incl. odd leading dimensions."""
    m, n, k = 201, 333, 517
    Cg, Cr, A, B, *_ = run_both(gpu, blas, "z", ta, tb, m, n, k, 0.5 - 1.5j, 1.0 + 0.5j, rng, pad=(1, 3, 2))
    assert_close("z", k, Cg, Cr, A, B)


@pytest.mark.parametrize("ta,tb", list(itertools.product((111, 112), (111, 112))))
def test_sgemm_ragged_large(gpu, blas, rng, ta, tb):
    """Tile-edge coverage of the FP32 SIMT kernel (128x128x8 tiles, vector loads need ld % 4 == 0)."""
    m, n, k = 335, 269, 531
    ra, rb = (m if ta == 111 else k), (k if tb == 111 else n)
    Cg, Cr, A, B, *_ = run_both(gpu, blas, "s", ta, tb, m, n, k, 1.5, -0.5, rng, pad=((-ra) % 4, 4 + (-rb) % 4, 3))
    assert_close("s", k, Cg, Cr, A, B)
    Cg, Cr, A, B, *_ = run_both(gpu, blas, "s", ta, tb, 256, 384, 128, 1.0, 0.0, rng)
    assert_close("s", 128, Cg, Cr, A, B)


@pytest.mark.parametrize("ta,tb,beta", [(111, 111, 0.0), (112, 112, 0.5)])
def test_host_pipelined_large(gpu, gpu_host, rng, ta, tb, beta):
    """All-host operands above the pipelining threshold: row blocks of A / column panels of B stream through the
    GPU (gemm_host_pipelined).  Checked against the device-resident path on the same data."""
    m, n, k = 8192 + 64, 2304, 2816
    ra, ca = (m, k) if ta == 111 else (k, m)
    rb, cb = (k, n) if tb == 111 else (n, k)
    A = rand(rng, "d", ra, ca, ra + 2); B = rand(rng, "d", rb, cb, rb + 4); C0 = rand(rng, "d", m, n, m + 2)
    if beta == 0.0:
        C0[:m, :] = np.nan
    C1, C2 = C0.copy(order="F"), C0.copy(order="F")
    gpu_host.gemm("d", 102, ta, tb, m, n, k, 1.25, A, ra + 2, B, rb + 4, beta, C1, m + 2)
    gpu.gemm("d", 102, ta, tb, m, n, k, 1.25, A, ra + 2, B, rb + 4, beta, C2, m + 2)
    assert np.isfinite(C1[:m]).all() and np.isnan(C1[m:]).all()
    assert fro(C1[:m] - C2[:m]) / fro(C2[:m]) <= 4 * k * EPS["d"]
    # spot check a few entries against a float64 dot product
    for (i, j) in ((0, 0), (m - 1, n - 1), (4100, 1234)):
        a = A[i, :k] if ta == 111 else A[:k, i]
        b = B[:k, j] if tb == 111 else B[j, :k]
        want = 1.25 * float(np.dot(a, b)) + (beta * C0[i, j] if beta else 0.0)
        assert abs(C1[i, j] - want) <= 1e-9 * k


@pytest.mark.parametrize("t,ta,tb,order", [("d", 111, 111, 102), ("d", 112, 111, 102), ("d", 111, 112, 101),
                                           ("z", 111, 113, 102), ("z", 112, 111, 101), ("s", 111, 111, 102), ("s", 112, 112, 101),
                                           ("c", 111, 111, 102), ("c", 113, 112, 102), ("c", 112, 113, 101)])
def test_headline_kernels_many_tiles_vs_openblas(gpu, blas, rng, t, ta, tb, order):
    """The persistent tile loop of the tuned kernels at bench-like scale — more than 1000 C tiles per launch (several
    tiles per CTA, band rasterisation wrapping, K = 4096 through every ring stage many times), both storage orders,
    beta != 0 with a non-trivial alpha, padded leading dimensions — full matrix against OpenBLAS on the host with
    north_star's Frobenius bound.  (The N = 16384 / 8192 configurations themselves are column-checked inside bench.py.)"""
    m, n, k = (4224, 4352, 4096) if t in "ds" else (2944, 4224, 2048)      # 33 x 34 (d, s) / 46 x 33 (z) tiles
    alpha, beta = (1.25, -0.75) if t in "ds" else (1.25 - 0.5j, -0.75 + 0.25j)
    pad = (4, 8, 4) if t == "s" else (2, 4, 2)
    Cg, Cr, A, B, *_ = run_both(gpu, blas, t, ta, tb, m, n, k, alpha, beta, rng, pad=pad, order=order)
    assert_close(t, k, Cg, Cr, A, B)


@pytest.mark.parametrize("ta,tb,order,m,n,k", [(111, 111, 102, 2304, 1408, 1000), (111, 112, 102, 2560, 1536, 520),
                                               (111, 111, 101, 1408, 2304, 776), (111, 111, 102, 9216, 4480, 264)])
def test_sgemm_256x128_tile_vs_openblas(gpu, blas, rng, ta, tb, order, m, n, k):
    """The 256 x 128 tile of the TMA-staged sgemm (A MN-major; forced by sgemm.tile = 256 below its size threshold, taken by
    size in the last case): ragged M, N and K edges, padded leading dimensions, alpha and beta != 0, a row-major caller
    (order 101 keeps the transpose flags and swaps the operands, so the 256-row side is then n)."""
    from jalla_b200._lib import lib
    lib.jb_set_option(b"sgemm.tile", 256 if m < 9000 else -1)
    try:
        Cg, Cr, A, B, *_ = run_both(gpu, blas, "s", ta, tb, m, n, k, 1.25, -0.75, rng, pad=(4, 8, 4), order=order)
    finally:
        lib.jb_set_option(b"sgemm.tile", -1)
    assert_close("s", k, Cg, Cr, A, B)


@pytest.mark.parametrize("ta,tb", list(itertools.product((111, 112, 113), (111, 112, 113))))
def test_cgemm_tma_all_ops_ragged(gpu, blas, rng, force_dmma, ta, tb):
    """The TMA-staged FFMA2 cgemm (128 x 128 complex tile) for every op(A), op(B) pair: ragged M, N and K edges (TMA zero
    fill, predicated stores), even padded leading dimensions, complex alpha and beta."""
    m, n, k = 300, 282, 203
    Cg, Cr, A, B, *_ = run_both(gpu, blas, "c", ta, tb, m, n, k, 0.75 - 0.5j, -1.25 + 0.25j, rng, pad=(2, 4, 6))
    assert_close("c", k, Cg, Cr, A, B)
    Cg, Cr, A, B, *_ = run_both(gpu, blas, "c", ta, tb, m, n, k, 1.0, 0.0, rng, pad=(0, 0, 0), order=101)
    assert_close("c", k, Cg, Cr, A, B)
