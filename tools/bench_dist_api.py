#!/usr/bin/env python3
"""Times the single-process multi-GPU C ABI (jb_dist_dgetrf / dpotrf / dgemm) on every GPU of the box with pinned host
operands; checks pivots / residuals against LAPACKE on the host at the smaller sizes.  One JSON line per measurement.
    python tools/bench_dist_api.py [--sizes 16384 32768] [--nb 512] [--gpus N]"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jalla_b200 import _lib, synth
from jalla_b200._lib import lib, check
from jalla_b200.dist import grid_shape

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[16384, 32768])
ap.add_argument("--nb", type=int, default=512)
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--check-upto", type=int, default=16384)
args = ap.parse_args()
check(lib.jb_init(0), "init")
world = args.gpus or max(w for w in (1, 2, 4, 8) if w <= lib.jb_device_count())
pr, pc = grid_shape(world)
h = C.c_void_p()
check(lib.jb_dist_create(C.byref(h), pr, pc, args.nb), "create")
PEAK = 37.07


def pinned(nbytes):
    p = lib.jb_malloc(nbytes, _lib.JB_MEM_PINNED); assert p
    return p


for n in args.sizes:
    pA = pinned(n * n * 8); A = np.ctypeslib.as_array(C.cast(pA, C.POINTER(C.c_double)), shape=(n * n,)).reshape(n, n).T     # column-major view
    dA = lib.jb_malloc(n * n * 8, 0)
    for kind in ("getrf", "potrf"):
        if kind == "getrf":
            check(lib.jb_dfill_uniform(dA, n, n, n, 0, 0, 20261019, 0, 0.0), "fill")
        else:
            check(lib.jb_dfill_uniform(dA, n, n, n, 0, 0, 20261022, 1, float(n)), "fill")
        check(lib.jb_memcpy_d2h(pA, dA, n * n * 8), "d2h")
        A0 = A.copy(order="F") if n <= args.check_upto else None
        piv = np.zeros(n, np.int32)
        for rep in range(2):
            if rep:
                check(lib.jb_memcpy_d2h(pA, dA, n * n * 8), "d2h")
            t0 = time.perf_counter()
            info = lib.jb_dist_dgetrf(h, n, pA, n, piv.ctypes.data) if kind == "getrf" else lib.jb_dist_dpotrf(h, b"L", n, pA, n)
            dt = (time.perf_counter() - t0) * 1e3
            check(info, kind); assert info == 0, info
        ms = [lib.jb_dist_last_ms(h, i) for i in range(3)]
        flop = (2.0 / 3.0 if kind == "getrf" else 1.0 / 3.0) * n ** 3
        rec = {"what": f"jb_dist_d{kind}", "n": n, "nb": args.nb, "n_gpus": world, "call_ms": dt, "scatter_ms": ms[0], "factor_ms": ms[1], "gather_ms": ms[2],
               "factor_tflops": flop / ms[1] / 1e9, "frac_of_dmma_peak": flop / ms[1] / 1e9 / (PEAK * world)}
        # solve with the resident factors and check the residual against the original matrix (regenerated row by row)
        b = synth.uniform(n, 1, 777, "d")[:, 0].copy(); x = b.copy()
        t0 = time.perf_counter()
        rc = lib.jb_dist_dgetrs(h, b"N", n, 1, pA, n, piv.ctypes.data, x.ctypes.data, n) if kind == "getrf" else lib.jb_dist_dpotrs(h, b"L", n, 1, pA, n, x.ctypes.data, n)
        rec["solve_ms"] = (time.perf_counter() - t0) * 1e3
        check(rc, "solve")
        if A0 is not None:
            if kind == "potrf":
                A0 = np.tril(A0) + np.tril(A0, -1).T
            r = np.linalg.norm(A0 @ x - b) / (np.linalg.norm(A0) * np.linalg.norm(x) * n * np.finfo(np.float64).eps)
            rec["solve_residual_over_n_eps"] = float(r)
            if kind == "getrf":
                from oracle import cpu_ref
                ob = cpu_ref.openblas(); cpu_ref.openblas_set_threads(os.cpu_count())
                Ar, pr_ = A0.copy(order="F"), np.zeros(n, np.int32)
                t0 = time.perf_counter(); assert ob.getrf("d", 102, n, n, Ar, n, pr_) == 0
                rec["host_lapacke_dgetrf_ms"] = (time.perf_counter() - t0) * 1e3
                rec["pivots_identical_to_lapacke"] = bool((piv == pr_).all())
        print(json.dumps(rec), flush=True)
    lib.jb_free(dA); lib.jb_free(pA)
lib.jb_dist_destroy(h)
