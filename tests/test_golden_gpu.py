"""The CUDA path against the committed golden vectors (tests/golden/, OpenBLAS through Jalla's C symbols)."""
import os

import numpy as np
import pytest

from tests.util import NP, EPS, fro

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "jalla_path_golden.npz"))


def close(t, a, b, factor):
    return fro(a - b) <= factor * EPS[t] * max(fro(b), 1e-300)


@pytest.mark.parametrize("t", "sdcz")
def test_golden_gemm(gpu, t):
    m, n, k = 23, 17, 31
    for ta in (111, 112):
        for tb in (111, 112):
            key = f"{t}gemm_{ta}_{tb}"
            A, B = np.asfortranarray(GOLD[key + "_A"]), np.asfortranarray(GOLD[key + "_B"])
            Cm = np.full((m, n), np.nan, NP[t], order="F")
            gpu.gemm(t, 102, ta, tb, m, n, k, 1.0, A, A.shape[0], B, B.shape[0], 0.0, Cm, m)
            assert close(t, Cm, GOLD[key + "_C"], 4 * k)
    A, B, C0 = (np.asfortranarray(GOLD[f"{t}gemm_ab_{s}"]) for s in ("A", "B", "C0"))
    Cm = C0.copy(order="F")
    gpu.gemm(t, 102, 111, 111, m, n, k, 0.5, A, m, B, k, -2.0, Cm, m)
    assert close(t, Cm, GOLD[f"{t}gemm_ab_C"], 4 * k)


@pytest.mark.parametrize("t", "sdcz")
def test_golden_factorisations(gpu, t):
    n, nrhs = 24, 3
    A, B = np.asfortranarray(GOLD[f"{t}gesv_A"]), np.asfortranarray(GOLD[f"{t}gesv_B"])
    LU, X, piv = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
    assert gpu.gesv(t, 102, n, nrhs, LU, n, piv, X, n) == 0
    np.testing.assert_array_equal(piv, GOLD[f"{t}gesv_ipiv"])
    assert close(t, LU, GOLD[f"{t}gesv_LU"], 50 * n) and close(t, X, GOLD[f"{t}gesv_X"], 1e3 * n)
    inv = LU.copy(order="F")
    assert gpu.getri(t, 102, n, inv, n, piv) == 0
    assert close(t, inv, GOLD[f"{t}getri_inv"], 1e4 * n)
    XT = B.copy(order="F")
    assert gpu.getrs(t, 102, "T", n, nrhs, LU, n, piv, XT, n) == 0
    assert close(t, XT, GOLD[f"{t}getrs_T_X"], 1e3 * n)
    S = np.asfortranarray(GOLD[f"{t}spd_S"])
    for uplo in "LU":
        F, Y = S.copy(order="F"), B.copy(order="F")
        assert gpu.posv(t, 102, uplo, n, nrhs, F, n, Y, n) == 0
        tri = np.tril_indices(n) if uplo == "L" else np.triu_indices(n)
        assert close(t, F[tri], GOLD[f"{t}potrf_{uplo}_F"][tri], 50 * n)
        assert close(t, Y, GOLD[f"{t}potrs_{uplo}_X"], 1e3 * n)


def test_golden_ties_and_info(gpu):
    T = np.asfortranarray(GOLD["d_tie_A"]).copy(order="F"); pt = np.zeros(16, np.int32)
    assert gpu.getrf("d", 102, 16, 16, T, 16, pt) == int(GOLD["d_tie_info"])
    np.testing.assert_array_equal(pt, GOLD["d_tie_ipiv"])
    S = np.asfortranarray(GOLD["dspd_S"]).copy(order="F"); S[7, 7] = -1.0
    assert gpu.potrf("d", 102, "L", 24, S, 24) == int(GOLD["d_info_not_pd"])
