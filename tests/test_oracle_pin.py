"""Pins the CPU oracle (oracle/jalla_oracle.c) before anything trusts it:
  * against the committed golden vectors (tests/golden/, generated from the OpenBLAS Jalla would link);
  * live against that OpenBLAS and the second, unprefixed OpenBLAS build, on fresh random inputs;
  * against the reference's own GEMM pins restated (tests/Test.hs:77-103 exactness on diagonal products).
Runs on CPU (no GPU needed)."""
import os

import numpy as np
import pytest

from oracle import cpu_ref
from tests.util import NP, EPS, rand, jalla_arbitrary, fro

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "jalla_path_golden.npz"))
TYPES = "sdcz"


def close(t, a, b, factor):
    return fro(a - b) <= factor * EPS[t] * max(fro(b), 1e-300)


@pytest.mark.parametrize("t", TYPES)
def test_golden_gemm(orc, t):
    m, n, k = 23, 17, 31
    for ta in (111, 112):
        for tb in (111, 112):
            key = f"{t}gemm_{ta}_{tb}"
            A, B, want = GOLD[key + "_A"], GOLD[key + "_B"], GOLD[key + "_C"]
            Cm = np.full((m, n), np.nan, NP[t], order="F")
            orc.gemm(t, 102, ta, tb, m, n, k, 1.0, np.asfortranarray(A), A.shape[0], np.asfortranarray(B), B.shape[0], 0.0, Cm, m)
            assert close(t, Cm, want, 4 * k)
    A, B, C0, want = (np.asfortranarray(GOLD[f"{t}gemm_ab_{s}"]) for s in ("A", "B", "C0", "C"))
    Cm = C0.copy(order="F")
    orc.gemm(t, 102, 111, 111, m, n, k, 0.5, A, m, B, k, -2.0, Cm, m)
    assert close(t, Cm, want, 4 * k)


@pytest.mark.parametrize("t", TYPES)
def test_golden_lu_chain(orc, t):
    n, nrhs = 24, 3
    A, B = np.asfortranarray(GOLD[f"{t}gesv_A"]), np.asfortranarray(GOLD[f"{t}gesv_B"])
    LU, X, piv = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
    assert orc.gesv(t, 102, n, nrhs, LU, n, piv, X, n) == 0
    np.testing.assert_array_equal(piv, GOLD[f"{t}gesv_ipiv"])           # pivots bit-identical
    assert close(t, LU, GOLD[f"{t}gesv_LU"], 50 * n)
    assert close(t, X, GOLD[f"{t}gesv_X"], 1e3 * n)
    inv = LU.copy(order="F")
    assert orc.getri(t, 102, n, inv, n, piv) == 0
    assert close(t, inv, GOLD[f"{t}getri_inv"], 1e3 * n)
    XT = B.copy(order="F")
    assert orc.getrs(t, 102, "T", n, nrhs, LU, n, piv, XT, n) == 0
    assert close(t, XT, GOLD[f"{t}getrs_T_X"], 1e3 * n)


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_golden_cholesky(orc, t, uplo):
    n, nrhs = 24, 3
    S, B = np.asfortranarray(GOLD[f"{t}spd_S"]), np.asfortranarray(GOLD[f"{t}gesv_B"])
    F, Y = S.copy(order="F"), B.copy(order="F")
    assert orc.potrf(t, 102, uplo, n, F, n) == 0
    assert orc.potrs(t, 102, uplo, n, nrhs, F, n, Y, n) == 0
    tri = np.tril_indices(n) if uplo == "L" else np.triu_indices(n)
    assert close(t, F[tri], GOLD[f"{t}potrf_{uplo}_F"][tri], 50 * n)
    other = np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)
    np.testing.assert_array_equal(F[other], S[other])                   # other triangle untouched
    assert close(t, Y, GOLD[f"{t}potrs_{uplo}_X"], 1e3 * n)


def test_golden_info_conventions(orc):
    rng = np.random.default_rng(1)
    A = rand(rng, "d", 12, 12); A[:, 4] = 0
    piv = np.zeros(12, np.int32)
    assert orc.getrf("d", 102, 12, 12, A, 12, piv) == int(GOLD["d_info_zero_pivot"]) == 5
    An = rand(rng, "d", 12, 12); An[2, 3] = np.nan
    assert orc.gesv("d", 102, 12, 1, An, 12, piv, rand(rng, "d", 12, 1), 12) == int(GOLD["d_info_nan_A_gesv"]) == -4
    Bn = rand(rng, "d", 12, 1); Bn[5, 0] = np.nan
    assert orc.gesv("d", 102, 12, 1, rand(rng, "d", 12, 12), 12, piv, Bn, 12) == int(GOLD["d_info_nan_B_gesv"]) == -7
    S = np.asfortranarray(GOLD["dspd_S"]).copy(order="F"); S[7, 7] = -1.0
    assert orc.potrf("d", 102, "L", 24, S, 24) == int(GOLD["d_info_not_pd"]) == 8
    T = np.asfortranarray(GOLD["d_tie_A"]).copy(order="F"); pt = np.zeros(16, np.int32)
    assert orc.getrf("d", 102, 16, 16, T, 16, pt) == int(GOLD["d_tie_info"])
    np.testing.assert_array_equal(pt, GOLD["d_tie_ipiv"])               # ties: lowest index wins, like i?amax
    assert close("d", T, GOLD["d_tie_LU"], 100)


@pytest.mark.parametrize("t", TYPES)
def test_live_against_openblas(orc, blas, rng, t):
    """Fresh random inputs, every routine, both libraries in the same process."""
    for n in (1, 2, 33, 90):
        A, B = rand(rng, t, n, n, n + 2), rand(rng, t, n, 2, n + 1)
        Ao, Ab, Bo, Bb = A.copy(order="F"), A.copy(order="F"), B.copy(order="F"), B.copy(order="F")
        po, pb = np.zeros(n, np.int32), np.zeros(n, np.int32)
        assert orc.gesv(t, 102, n, 2, Ao, n + 2, po, Bo, n + 1) == blas.gesv(t, 102, n, 2, Ab, n + 2, pb, Bb, n + 1) == 0
        np.testing.assert_array_equal(po, pb)
        assert close(t, Ao[:n], Ab[:n], 100 * n) and close(t, Bo[:n], Bb[:n], 1e4 * n)
        assert orc.getri(t, 102, n, Ao, n + 2, po) == blas.getri(t, 102, n, Ab, n + 2, pb) == 0
        assert close(t, Ao[:n], Ab[:n], 1e4 * n)
    for (m, n) in ((40, 25), (25, 40)):     # rectangular getrf
        A = rand(rng, t, m, n)
        Ao, Ab = A.copy(order="F"), A.copy(order="F")
        po, pb = np.zeros(min(m, n), np.int32), np.zeros(min(m, n), np.int32)
        assert orc.getrf(t, 102, m, n, Ao, m, po) == blas.getrf(t, 102, m, n, Ab, m, pb) == 0
        np.testing.assert_array_equal(po, pb)
        assert close(t, Ao, Ab, 100 * max(m, n))


def test_second_openblas_agrees(orc, rng):
    try:
        ob2 = cpu_ref.openblas_second()
    except Exception as e:  # the second opinion is optional
        pytest.skip(str(e))
    n = 64
    A = rand(rng, "d", n, n)
    Ao, A2 = A.copy(order="F"), A.copy(order="F")
    po, p2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert orc.getrf("d", 102, n, n, Ao, n, po) == ob2.getrf("d", 102, n, n, A2, n, p2) == 0
    np.testing.assert_array_equal(po, p2)


@pytest.mark.parametrize("t", TYPES)
def test_row_major_and_transposes_vs_openblas(orc, blas, rng, t):
    m, n, k = 13, 9, 11
    for order in (101, 102):
        for ta in (111, 112, 113):
            for tb in (111, 112, 113):
                ra, ca = (m, k) if (ta == 111) == (order == 102) else (k, m)
                rb, cb = (k, n) if (tb == 111) == (order == 102) else (n, k)
                A, B = rand(rng, t, ra, ca), rand(rng, t, rb, cb)
                rc, cc = (m, n) if order == 102 else (n, m)
                C0 = rand(rng, t, rc, cc)
                Co, Cb = C0.copy(order="F"), C0.copy(order="F")
                orc.gemm(t, order, ta, tb, m, n, k, 1.5, A, ra, B, rb, 0.5, Co, rc)
                blas.gemm(t, order, ta, tb, m, n, k, 1.5, A, ra, B, rb, 0.5, Cb, rc)
                assert close(t, Co, Cb, 8 * k), (order, ta, tb)


@pytest.mark.parametrize("t", TYPES)
def test_prop_matrixMultDiag_on_oracle(orc, rng, t):
    """tests/Test.hs:77-94: mat ## diag([1..]) scales columns exactly."""
    for _ in range(10):
        mat = jalla_arbitrary(rng, t)
        m, n = mat.shape
        dm = np.zeros((n, m), NP[t], order="F")
        for i in range(min(m, n)):
            dm[i, i] = i + 1
        Cm = np.full((m, m), np.nan, NP[t], order="F")
        orc.gemm(t, 102, 111, 111, m, m, n, 1.0, mat, m, dm, n, 0.0, Cm, m)
        want = np.zeros((m, m), NP[t]); w = min(m, n)
        want[:, :w] = mat[:, :w] * np.arange(1, w + 1, dtype=NP[t])
        np.testing.assert_array_equal(Cm, want)
