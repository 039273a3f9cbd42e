#!/usr/bin/env python3
"""Generates tests/golden/jalla_path_golden.npz — small input/output vectors for the Jalla hot path,
produced by the third-party library Jalla links (jalla.cabal:75 `lapacke cblas blas`), here the
scipy-bundled OpenBLAS 0.3.31.dev (LP64), called through the exact C symbols Jalla's FFI imports
(cblas_?gemm, LAPACKE_?{gesv,getrf,getri,getrs,potrf,potrs}) with Jalla's argument convention
(Numeric/Jalla/Matrix.hs:482,503,766,769: order 102, lda = rows).  The reference tree itself holds no
golden vectors for this path, so these pin the oracle restatement and the CUDA path.

Run in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_ref  # noqa: E402

NP = cpu_ref.NP
rng = np.random.default_rng(20261019)
ob = cpu_ref.openblas()
out = {"_library": np.array(ob.info.get("config", "OpenBLAS"))}


def rnd(t, r, c):
    v = rng.uniform(-1, 1, (r, c))
    if t in "cz":
        v = v + 1j * rng.uniform(-1, 1, (r, c))
    return np.asfortranarray(v.astype(NP[t]))


for t in "sdcz":
    # gemm: every NoTrans/Trans combination on a ragged shape, Jalla's alpha=1, beta=0, plus a general alpha/beta
    m, n, k = 23, 17, 31
    for ta in (111, 112):
        for tb in (111, 112):
            A = rnd(t, *((m, k) if ta == 111 else (k, m))); B = rnd(t, *((k, n) if tb == 111 else (n, k)))
            Cm = np.zeros((m, n), NP[t], order="F")
            ob.gemm(t, 102, ta, tb, m, n, k, 1.0, A, A.shape[0], B, B.shape[0], 0.0, Cm, m)
            key = f"{t}gemm_{ta}_{tb}"
            out[key + "_A"], out[key + "_B"], out[key + "_C"] = A, B, Cm
    A, B, C0 = rnd(t, m, k), rnd(t, k, n), rnd(t, m, n)
    Cm = C0.copy(order="F")
    ob.gemm(t, 102, 111, 111, m, n, k, 0.5, A, m, B, k, -2.0, Cm, m)
    out[f"{t}gemm_ab_A"], out[f"{t}gemm_ab_B"], out[f"{t}gemm_ab_C0"], out[f"{t}gemm_ab_C"] = A, B, C0, Cm
    # gesv / getrf / getri / getrs
    n, nrhs = 24, 3
    A, B = rnd(t, n, n), rnd(t, n, nrhs)
    LU, X, piv = A.copy(order="F"), B.copy(order="F"), np.zeros(n, np.int32)
    assert ob.gesv(t, 102, n, nrhs, LU, n, piv, X, n) == 0
    out[f"{t}gesv_A"], out[f"{t}gesv_B"], out[f"{t}gesv_LU"], out[f"{t}gesv_ipiv"], out[f"{t}gesv_X"] = A, B, LU, piv, X
    inv = LU.copy(order="F")
    assert ob.getri(t, 102, n, inv, n, piv) == 0
    out[f"{t}getri_inv"] = inv
    XT = B.copy(order="F")
    assert ob.getrs(t, 102, "T", n, nrhs, LU, n, piv, XT, n) == 0
    out[f"{t}getrs_T_X"] = XT
    # potrf / potrs, both triangles
    M = rnd(t, n, n)
    S = np.asfortranarray((M @ M.conj().T + n * np.eye(n)).astype(NP[t]))
    out[f"{t}spd_S"] = S
    for u in "LU":
        F, Y = S.copy(order="F"), B.copy(order="F")
        assert ob.potrf(t, 102, u, n, F, n) == 0
        assert ob.potrs(t, 102, u, n, nrhs, F, n, Y, n) == 0
        out[f"{t}potrf_{u}_F"], out[f"{t}potrs_{u}_X"] = F, Y
# info conventions (double): zero pivot column, NaN screening, not positive definite
A = rnd("d", 12, 12); A[:, 4] = 0
piv = np.zeros(12, np.int32)
out["d_info_zero_pivot"] = np.array(ob.getrf("d", 102, 12, 12, A.copy(order="F"), 12, piv))
An = rnd("d", 12, 12); An[2, 3] = np.nan
out["d_info_nan_A_gesv"] = np.array(ob.gesv("d", 102, 12, 1, An.copy(order="F"), 12, piv, rnd("d", 12, 1), 12))
Bn = rnd("d", 12, 1); Bn[5, 0] = np.nan
out["d_info_nan_B_gesv"] = np.array(ob.gesv("d", 102, 12, 1, rnd("d", 12, 12), 12, piv, Bn, 12))
S = out["dspd_S"].copy(order="F"); S[7, 7] = -1.0
out["d_info_not_pd"] = np.array(ob.potrf("d", 102, "L", 24, S, 24))
# pivot tie-break: lowest index wins (integer matrix with repeated magnitudes)
T = np.asfortranarray(rng.integers(-2, 3, (16, 16)).astype(np.float64) + 3 * np.eye(16))
LUt, pt = T.copy(order="F"), np.zeros(16, np.int32)
out["d_tie_A"] = T
out["d_tie_info"] = np.array(ob.getrf("d", 102, 16, 16, LUt, 16, pt))
out["d_tie_ipiv"], out["d_tie_LU"] = pt, LUt
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "jalla_path_golden.npz"), **out)
print("wrote", len(out), "arrays;", ob.info)
