"""cblas_?trsm (the level-3 sibling the factorisations and the distributed Cholesky use) against OpenBLAS:
every side / uplo / trans / diag combination, both orders, blocked sizes."""
import itertools

import numpy as np
import pytest

from tests.util import NP, EPS, rand, fro

pytestmark = pytest.mark.gpu


def tri_matrix(rng, t, n, ld):
    A = rand(rng, t, n, n, ld)
    A[:n, :] += (n * np.eye(n)).astype(NP[t])      # well conditioned triangles
    return A


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("side,uplo,trans,diag", list(itertools.product((141, 142), (121, 122), (111, 112, 113), (131, 132))))
def test_trsm_all_variants(gpu, blas, rng, t, side, uplo, trans, diag):
    m, n = 77, 53
    ka = m if side == 141 else n
    A = tri_matrix(rng, t, ka, ka + 1)
    B = rand(rng, t, m, n, m + 2)
    alpha = 0.5 if t in "sd" else 0.5 - 0.25j
    Bg, Bb = B.copy(order="F"), B.copy(order="F")
    gpu.trsm(t, 102, side, uplo, trans, diag, m, n, alpha, A, ka + 1, Bg, m + 2)
    blas.trsm(t, 102, side, uplo, trans, diag, m, n, alpha, A, ka + 1, Bb, m + 2)
    assert fro(Bg[:m] - Bb[:m]) / fro(Bb[:m]) <= 200 * max(m, n) * EPS[t]
    assert np.isnan(Bg[m:]).all()


@pytest.mark.parametrize("t", "dz")
def test_trsm_row_major_and_large(gpu, blas, rng, t):
    m, n = 300, 130
    for side, uplo, trans in [(142, 122, 113), (141, 121, 111), (142, 121, 111), (141, 122, 112)]:
        ka = m if side == 141 else n
        A = tri_matrix(rng, t, ka, ka)
        for order in (102, 101):
            B = rand(rng, t, m, n) if order == 102 else rand(rng, t, n, m)
            ldb = B.shape[0]
            Bg, Bb = B.copy(order="F"), B.copy(order="F")
            gpu.trsm(t, order, side, uplo, trans, 131, m, n, 1.0, A, ka, Bg, ldb)
            blas.trsm(t, order, side, uplo, trans, 131, m, n, 1.0, A, ka, Bb, ldb)
            assert fro(Bg - Bb) / fro(Bb) <= 500 * max(m, n) * EPS[t], (side, uplo, trans, order)


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("side,uplo,trans,diag", list(itertools.product((141, 142), (121, 122), (111, 112, 113), (131, 132))))
def test_trsm_gemm_form_all_variants(gpu, blas, rng, t, side, uplo, trans, diag):
    """Wide solves take the GEMM form (inverted 64 x 64 diagonal blocks, csrc/trsm_inv.cu): triangle of order >= 128 against
    >= 256 right-hand sides / rows, ragged last block, padded leading dimensions (NaN padding must stay NaN)."""
    m, n = (200, 300) if side == 141 else (300, 200)
    ka = m if side == 141 else n
    A = tri_matrix(rng, t, ka, ka + 2)
    B = rand(rng, t, m, n, m + 2)
    alpha = 0.5 if t in "sd" else 0.5 - 0.25j
    Bg, Bb = B.copy(order="F"), B.copy(order="F")
    gpu.trsm(t, 102, side, uplo, trans, diag, m, n, alpha, A, ka + 2, Bg, m + 2)
    blas.trsm(t, 102, side, uplo, trans, diag, m, n, alpha, A, ka + 2, Bb, m + 2)
    assert fro(Bg[:m] - Bb[:m]) / fro(Bb[:m]) <= 200 * max(m, n) * EPS[t]
    assert np.isnan(Bg[m:]).all()


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("nrhs", [1, 3, 4, 7])
@pytest.mark.parametrize("uplo,trans,diag", list(itertools.product((121, 122), (111, 112, 113), (131, 132))))
def test_trsm_few_right_hand_sides(gpu, blas, rng, t, nrhs, uplo, trans, diag):
    """Left solves of order >= 256 against at most eight right-hand sides take the bandwidth-bound sweep (csrc/trsm_inv.cu,
    trsv_step_kernel: one launch per 64 rows on the inverted diagonal blocks): ragged last block (300 = 4 x 64 + 44),
    padded leading dimensions whose NaN padding must survive, every uplo / trans / diag, against OpenBLAS."""
    m = 300
    A = tri_matrix(rng, t, m, m + 2)
    B = rand(rng, t, m, nrhs, m + 3)
    alpha = 0.5 if t in "sd" else 0.5 - 0.25j
    Bg, Bb = B.copy(order="F"), B.copy(order="F")
    gpu.trsm(t, 102, 141, uplo, trans, diag, m, nrhs, alpha, A, m + 2, Bg, m + 3)
    blas.trsm(t, 102, 141, uplo, trans, diag, m, nrhs, alpha, A, m + 2, Bb, m + 3)
    assert fro(Bg[:m] - Bb[:m]) / fro(Bb[:m]) <= 200 * m * EPS[t]
    assert np.isnan(Bg[m:]).all()


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("uplo,trans", list(itertools.product((121, 122), (111, 112, 113))))
def test_trsv_persistent_sweep_several_blocks_per_cta(gpu, blas, rng, t, uplo, trans):
    """One right-hand side takes the persistent pipelined sweep (trsv_pipe_kernel): order 1100 = 17 x 64 + 12 with the
    grid capped at 5 CTAs, so that every CTA walks several row blocks in sweep order, and uncapped; against OpenBLAS and
    against the stepped sweep (trsv.pipe = 0)."""
    from jalla_b200._lib import lib
    m = 1100
    A = tri_matrix(rng, t, m, m + 1)
    B = rand(rng, t, m, 1, m + 5)
    Bb = B.copy(order="F")
    blas.trsm(t, 102, 141, uplo, trans, 131, m, 1, 1.0, A, m + 1, Bb, m + 5)
    outs = []
    for opts in ({"trsv.pipe_ctas": 5}, {}, {"trsv.pipe": 0}):
        for k, v in opts.items():
            lib.jb_set_option(k.encode(), v)
        try:
            Bg = B.copy(order="F")
            gpu.trsm(t, 102, 141, uplo, trans, 131, m, 1, 1.0, A, m + 1, Bg, m + 5)
        finally:
            for k in opts:
                lib.jb_set_option(k.encode(), -1)
        assert fro(Bg[:m] - Bb[:m]) / fro(Bb[:m]) <= 200 * m * EPS[t]
        assert np.isnan(Bg[m:]).all()
        outs.append(Bg[:m].copy())
    np.testing.assert_array_equal(outs[0], outs[1])      # the grid size does not change a single operation


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("nrhs", [2, 3, 5, 8])
@pytest.mark.parametrize("uplo,trans", list(itertools.product((121, 122), (111, 113))))
def test_trsv_persistent_sweep_several_right_hand_sides(gpu, blas, rng, t, nrhs, uplo, trans):
    """The persistent sweep with 2, 4 or 8 columns per launch (Complex Double beyond four columns falls back to the stepped
    sweep): order 700, grid capped at 4 CTAs and uncapped, and the stepped sweep (trsv.pipe = 0); padded leading dimensions,
    against OpenBLAS."""
    from jalla_b200._lib import lib
    m = 700
    A = tri_matrix(rng, t, m, m + 1)
    B = rand(rng, t, m, nrhs, m + 5)
    Bb = B.copy(order="F")
    blas.trsm(t, 102, 141, uplo, trans, 131, m, nrhs, 1.0, A, m + 1, Bb, m + 5)
    outs = []
    for opts in ({"trsv.pipe_ctas": 4}, {}, {"trsv.pipe": 0}):
        for k, v in opts.items():
            lib.jb_set_option(k.encode(), v)
        try:
            Bg = B.copy(order="F")
            gpu.trsm(t, 102, 141, uplo, trans, 131, m, nrhs, 1.0, A, m + 1, Bg, m + 5)
        finally:
            for k in opts:
                lib.jb_set_option(k.encode(), -1)
        assert fro(Bg[:m] - Bb[:m]) / fro(Bb[:m]) <= 200 * m * EPS[t]
        assert np.isnan(Bg[m:]).all()
        outs.append(Bg[:m].copy())
    np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.parametrize("t", "dc")
@pytest.mark.parametrize("side,uplo,trans", list(itertools.product((141, 142), (121, 122), (111, 113))))
def test_trsm_recursive_halving(gpu, blas, rng, t, side, uplo, trans):
    """Triangles above 512 are halved recursively (one GEMM couples the halves) before the 64-row sweep runs: order 1100
    (halves 576 + 524, the second halved again), ragged, every direction of the sweep."""
    m, n = (1100, 260) if side == 141 else (260, 1100)
    ka = m if side == 141 else n
    A = tri_matrix(rng, t, ka, ka)
    B = rand(rng, t, m, n, m + 2)
    Bg, Bb = B.copy(order="F"), B.copy(order="F")
    gpu.trsm(t, 102, side, uplo, trans, 131, m, n, 1.0, A, ka, Bg, m + 2)
    blas.trsm(t, 102, side, uplo, trans, 131, m, n, 1.0, A, ka, Bb, m + 2)
    assert fro(Bg[:m] - Bb[:m]) / fro(Bb[:m]) <= 200 * max(m, n) * EPS[t]
    assert np.isnan(Bg[m:]).all()


@pytest.mark.parametrize("t", "sdcz")
@pytest.mark.parametrize("uplo", [121, 122])
@pytest.mark.parametrize("trans", [111, 112])
def test_syrk_herk(gpu, blas, rng, t, uplo, trans):
    """cblas_?syrk and cblas_?herk against OpenBLAS: one triangle written, the other untouched, both orders."""
    n, k = 150, 90
    for herm in ([False, True] if t in "cz" else [False]):
        tr = trans if not (herm and trans == 112) else 113
        for order in (102, 101):
            rows = n if (tr == 111) == (order == 102) else k
            cols = k if (tr == 111) == (order == 102) else n
            A = rand(rng, t, rows, cols, rows + 2)
            C0 = rand(rng, t, n, n, n + 1)
            alpha, beta = (0.75, -0.5) if (herm or t in "sd") else (0.75 - 0.25j, -0.5 + 0.5j)
            Cg, Cb = C0.copy(order="F"), C0.copy(order="F")
            gpu.syrk(t, order, uplo, tr, n, k, alpha, A, rows + 2, beta, Cg, n + 1, herm=herm)
            blas.syrk(t, order, uplo, tr, n, k, alpha, A, rows + 2, beta, Cb, n + 1, herm=herm)
            assert fro(Cg[:n] - Cb[:n]) / fro(Cb[:n]) <= 50 * k * EPS[t], (herm, order)
            # the triangle that is not referenced keeps its input bits; padding row untouched
            lower_stored = (uplo == 122) == (order == 102)
            other = np.triu_indices(n, 1) if lower_stored else np.tril_indices(n, -1)
            np.testing.assert_array_equal(Cg[:n][other], C0[:n][other])
            assert np.isnan(Cg[n:]).all()
