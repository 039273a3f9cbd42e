#!/usr/bin/env python3
"""bench.py — headline benchmark of the jalla_b200 hot path (contract: see DESIGN.md §Measurement).

Metric (BASELINE.json): FP64 GEMM TFLOP/s at N=16384 (+ batched 32x32 LU solves/s), each
as a fraction of roofline, at 1/2/4/8 B200.  One "step" = one pass of the hot path over
one batch of synthetic input:
  * primary   — configs[1]: C = A*B, Double, N=16384, NN, alpha=1, beta=0, column-major
                 (at N>1 the SAME product, 2D-sharded Pr x Pc: strong scaling);
  * secondary — configs[2]: 1,000,000 independent 32x32 Double systems, getrf+getrs
                 (batch split across ranks, no collective), reported under "batched_lu".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
N>1 is launched by torchrun (one rank per GPU).  `--impl reference` times the reference's
own CPU path — the C boundary Jalla calls, resolved in the OpenBLAS this image ships — on
the host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GEMM = 16384
LU_N, LU_BATCH = 32, 1_000_000
LU_BYTES_PER_SYSTEM = 17024          # A in 8192 + b in 256 + LU out 8192 + ipiv out 128 + x out 256
SEED = 20261019
DMMA_PEAK_TFLOPS = 37.07             # measured: tools/peaks.cu on this pool's B200 (profiles/peaks_r1.jsonl)
FFMA_PEAK_TFLOPS = 72.35             # measured: same run (full-precision FP32 FMA, CUDA cores)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons, pw = [], [], set(), []
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """Jalla's CPU path at the C boundary it calls (cblas_dgemm / LAPACKE_dgetrf+dgetrs), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import cpu_ref
    ob = cpu_ref.openblas()
    cores = os.cpu_count() or 1
    cpu_ref.openblas_set_threads(cores)
    n = N_GEMM                                # the configuration itself: one cblas_dgemm N=16384 per step (~9 s on 16 cores)
    from jalla_b200 import synth
    A = synth.uniform(n, n, SEED, "d"); B = synth.uniform(n, n, SEED + 1, "d")
    Cm = np.empty((n, n), order="F")
    flop = 2.0 * n ** 3
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ob.gemm("d", 102, 111, 111, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = flop / (ms * 1e-3) / 1e12
    lu = cpu_batched_lu(ob, sample=200000, threads=cores)
    sample = f"one cblas_dgemm N={n} (the whole configs[1] product) per step, OpenBLAS {ob.info.get('core')} x{cores} threads"
    line = {
        "impl": "reference", "metric": "dgemm_fp64_n16384_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: FP64 dgemm N=16384 square, C=A*B (NN, alpha=1, beta=0), column-major ld=rows",
                   "n": n, "inputs": "counter-based U(-1,1) keyed on global (i,j) (numpy twin of the device generator)"},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "reference", "sample": sample,
                         "blas": ob.info.get("config")},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "batched_lu": lu, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_batched_lu(ob, sample, threads):
    """LAPACKE_dgetrf + LAPACKE_dgetrs per 32x32 system, OpenBLAS single-threaded per call and the
    batch loop split over host threads (ctypes releases the GIL)."""
    import numpy as np
    from oracle import cpu_ref
    from jalla_b200 import synth
    n = LU_N
    A = np.ascontiguousarray(synth.uniform(n, sample * n, SEED, "d").T).reshape(sample, n, n)
    b = np.ascontiguousarray(synth.uniform(n, sample, SEED + 1, "d").T)
    piv = np.zeros((sample, n), np.int32)
    cpu_ref.openblas_set_threads(1)
    a0, p0, b0 = A[0].copy(), piv[0].copy(), b[0].copy()          # binds the ctypes signatures (and warms the library up)
    ob.getrf("d", 102, n, n, a0, n, p0); ob.getrs("d", 102, "N", n, 1, a0, n, p0, b0, n)
    f_getrf = ob._cache[("getrf", "d")]; f_getrs = ob._cache[("getrs", "d")]
    drv = cpu_ref.oracle().lib.oracle_run_batched_lu_procs   # fork-parallel C loop (oracle/cpu_batch_driver.c)
    drv.restype = C.c_int
    drv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    t0 = time.perf_counter()
    bad = drv(C.cast(f_getrf, C.c_void_p), C.cast(f_getrs, C.c_void_p), n, A.ctypes.data, piv.ctypes.data, b.ctypes.data, sample, threads)
    dt = time.perf_counter() - t0
    assert bad == 0
    cpu_ref.openblas_set_threads(threads)
    return {"metric": "batched_lu_32x32_solves_per_s", "value": sample / dt, "unit": "solves/s", "cores": threads,
            "sample": f"{sample} of the 1,000,000 systems, LAPACKE_dgetrf+dgetrs per system, C loop split over forked processes (one per core), OpenBLAS threads=1"}


def small_config(torch, lib, check, stream):
    """BASELINE.json configs[0]: Double 512 x 512 multiply and LU solve (gesv, nrhs = 1) — device-resident on the GPU
    (launch-latency bound there), beside the same calls in OpenBLAS on the host."""
    import numpy as np
    from oracle import cpu_ref
    n = 512
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); B = torch.empty_like(A); Cm = torch.empty_like(A)
    A0 = torch.empty_like(A); b0 = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(b0)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, SEED, 0, 0.0), "fill")
    check(lib.jb_dfill_uniform(B.data_ptr(), n, n, n, 0, 0, SEED + 1, 0, 0.0), "fill")
    check(lib.jb_dfill_uniform(b0.data_ptr(), n, n, 1, 0, 0, SEED + 2, 0, 0.0), "fill")
    piv = (C.c_int * n)()

    def ev(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_gemm = ev(lambda: lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, A0.data_ptr(), n, B.data_ptr(), n, 0.0, Cm.data_ptr(), n), 50)

    def solve():
        A.copy_(A0); b.copy_(b0)
        assert lib.LAPACKE_dgesv(102, n, 1, A.data_ptr(), n, piv, b.data_ptr(), n) == 0
    t_copy = ev(lambda: (A.copy_(A0), b.copy_(b0)), 20)
    t_gesv = ev(solve, 20) - t_copy
    ob = cpu_ref.openblas()
    Ah = A0.cpu().numpy().reshape(n, n).T.copy(order="F"); Bh = B.cpu().numpy().reshape(n, n).T.copy(order="F")
    Ch = np.empty((n, n), order="F"); bh = b0.cpu().numpy().copy(); ph = np.zeros(n, np.int32)

    def best(f, reps):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
        return min(ts) * 1e3
    c_gemm = best(lambda: ob.gemm("d", 102, 111, 111, n, n, n, 1.0, Ah, n, Bh, n, 0.0, Ch, n), 20)
    c_gesv = best(lambda: ob.gesv("d", 102, n, 1, Ah.copy(order="F"), n, ph, bh.copy(), n), 10)
    return {"workload": "configs[0]: Double 512x512 multiply + LU solve (gesv nrhs=1), device-resident vs OpenBLAS host",
            "gpu_dgemm_ms": t_gemm, "gpu_dgesv_ms": t_gesv, "cpu_dgemm_ms": c_gemm, "cpu_dgesv_ms": c_gesv,
            "gpu_dgemm_tflops": 2.0 * n ** 3 / (t_gemm * 1e-3) / 1e12}



def verify_columns(np, torch, gemm, ncols):
    """Column-sampled full check of the distributed product: `ncols` random columns of this rank's C block, every row,
    against numpy's dgemm on the host (A row block and the sampled B columns copied back from the device; the panels a
    rank pulled from its peers are themselves compared with the regenerated inputs on sampled rows / columns).
    Returns (relative Frobenius error, north_star bound 4*n*eps*|A||B|/|C|)."""
    from jalla_b200 import synth
    P, n, t = gemm.plan, gemm.n, gemm.t
    assert gemm.mode in ("pull", "local"), "verify_columns reads the n x nb column block of the pull / local layouts"
    mb, nb = P["mb"], P["nb"]
    dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[t]
    rng = np.random.default_rng(1000 + gemm.grid.rank)
    cols = np.sort(rng.choice(nb, size=min(ncols, nb), replace=False))
    A = np.empty((mb, n), dtype=dt, order="F")
    gemm.check(gemm.lib.jb_memcpy_d2h(A.ctypes.data, gemm.A_ptr, A.nbytes), "verify d2h A")
    Bs = np.empty((n, len(cols)), dtype=dt, order="F")
    Cs = np.empty((mb, len(cols)), dtype=dt, order="F")
    es = A.itemsize
    col = np.empty(n, dtype=dt)
    for q, j in enumerate(cols):       # the B column block is n x nb, column-major, ld = n
        gemm.check(gemm.lib.jb_memcpy_d2h(col.ctypes.data, gemm.B_ptr + int(j) * n * es, n * es), "verify d2h B")
        Bs[:, q] = col
        gemm.check(gemm.lib.jb_memcpy_d2h(Cs[:, q].ctypes.data, gemm.C.data_ptr() + int(j) * mb * es, mb * es), "verify d2h C")
    # the operands on the device are the generator's (covers the peer pulls): sampled rows of A, all sampled columns of B
    for i in rng.choice(mb, size=4, replace=False):
        np.testing.assert_array_equal(A[i], synth.uniform(1, n, gemm.seed, t, P["row0"] + int(i), 0)[0])
    for q, j in enumerate(cols[:8]):
        np.testing.assert_array_equal(Bs[:, q], synth.uniform(n, 1, gemm.seed + 1, t, 0, P["col0"] + int(j))[:, 0])
    ref = A.astype(np.complex128 if t in "cz" else np.float64) @ Bs.astype(np.complex128 if t in "cz" else np.float64)
    err = float(np.linalg.norm(Cs - ref) / np.linalg.norm(ref))
    eps = float(np.finfo(np.float32 if t in "sc" else np.float64).eps)
    bound = 4 * n * eps * float(np.linalg.norm(A)) * float(np.linalg.norm(Bs)) / float(np.linalg.norm(ref))
    return err, bound


def factorisations(torch, lib, check, stream, hbm_gbs):
    """Large single-GPU LAPACKE calls on device-resident Double data (CUDA events around the call, copies subtracted):
    dgetrf / dpotrf / dgesv (nrhs = 1, the Jalla call) at 8192 and 16384, as fractions of the measured DMMA peak."""
    out = {}
    for n in (8192, 16384):
        A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
        b0 = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(b0)
        piv = torch.empty(n, dtype=torch.int32, device="cuda")

        def ev(fn, reps=2):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        t_copy = ev(lambda: A.copy_(A0))
        check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, SEED, 0, 0.0), "fill")
        check(lib.jb_dfill_uniform(b0.data_ptr(), n, n, 1, 0, 0, SEED + 2, 0, 0.0), "fill")

        def getrf():
            A.copy_(A0); assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0

        def gesv():
            A.copy_(A0); b.copy_(b0); assert lib.LAPACKE_dgesv(102, n, 1, A.data_ptr(), n, piv.data_ptr(), b.data_ptr(), n) == 0
        t_getrf, t_gesv = ev(getrf) - t_copy, ev(gesv) - t_copy
        # the solve alone on the factors gesv left in A (one right-hand side: the bandwidth-bound sweep of csrc/trsm_inv.cu)
        t_bcopy = ev(lambda: b.copy_(b0), reps=5)

        def getrs():
            b.copy_(b0); assert lib.LAPACKE_dgetrs(102, b"N", n, 1, A.data_ptr(), n, piv.data_ptr(), b.data_ptr(), n) == 0
        t_getrs = ev(getrs, reps=5) - t_bcopy
        check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, SEED + 3, 1, float(n)), "fill spd")

        def potrf():
            A.copy_(A0); assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
        t_potrf = ev(potrf) - t_copy
        f_lu, f_ch = 2.0 / 3.0 * n ** 3, n ** 3 / 3.0
        out[f"n{n}"] = {"dgetrf_ms": t_getrf, "dgetrf_tflops": f_lu / t_getrf / 1e9, "dgetrf_frac_of_dmma_peak": f_lu / t_getrf / 1e9 / DMMA_PEAK_TFLOPS,
                        "dgesv_nrhs1_ms": t_gesv, "dgetrs_nrhs1_ms": t_getrs, "dgetrs_nrhs1_frac_of_hbm": n * n * 8.0 / t_getrs / 1e6 / hbm_gbs,
                        "dpotrf_ms": t_potrf, "dpotrf_tflops": f_ch / t_potrf / 1e9,
                        "dpotrf_frac_of_dmma_peak": f_ch / t_potrf / 1e9 / DMMA_PEAK_TFLOPS}
        if n == 8192:
            # invert (Matrix.hs:537-540: getrf + getri) and the level-3 triangular solve both run on the GEMM-form trsm
            check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, SEED, 0, 0.0), "fill")

            def invert():
                A.copy_(A0)
                assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
                assert lib.LAPACKE_dgetri(102, n, A.data_ptr(), n, piv.data_ptr()) == 0
            t_inv = ev(invert, reps=1) - t_copy
            B = torch.empty(n * n, dtype=torch.float64, device="cuda")
            check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, SEED + 3, 1, float(n)), "fill")

            def trsm():
                B.copy_(A0); lib.cblas_dtrsm(102, 141, 122, 111, 131, n, n, 1.0, A0.data_ptr(), n, B.data_ptr(), n)
            t_trsm = ev(trsm) - t_copy
            out[f"n{n}"].update({"dgetrf_dgetri_ms": t_inv, "dgetrf_dgetri_tflops": 2.0 * n ** 3 / t_inv / 1e9,
                                 "dtrsm_LLNN_nrhs8192_ms": t_trsm, "dtrsm_tflops": float(n) ** 3 / t_trsm / 1e9})
            del B
        del A0, A, b0, b, piv
        torch.cuda.empty_cache()
    return out


def batched_cholesky(torch, lib, check, stream, hbm_gbs, batch=LU_BATCH, n=LU_N):
    """The SPD sibling of configs[2]: jb_dpotrf_batched + jb_dpotrs_batched (nrhs = 1) on `batch` 32 x 32 Double systems, lower
    storage; algorithmic bytes = the stored triangle in and out (potrf), the triangle in and b in / x out (potrs)."""
    G = torch.empty(batch * n * n, dtype=torch.float64, device="cuda")
    check(lib.jb_dfill_uniform(G.data_ptr(), n, n, batch * n, 0, 0, SEED + 5, 0, 0.0), "fill")
    Gv = G.view(batch, n, n)
    S0 = (torch.bmm(Gv, Gv.transpose(1, 2)) / n + torch.eye(n, dtype=torch.float64, device="cuda")).contiguous().view(-1)
    del G, Gv
    A = torch.empty_like(S0)
    b0 = torch.empty(batch * n, dtype=torch.float64, device="cuda"); b = torch.empty_like(b0)
    check(lib.jb_dfill_uniform(b0.data_ptr(), n, n, batch, 0, 0, SEED + 6, 0, 0.0), "fill")
    info = torch.empty(batch, dtype=torch.int32, device="cuda")
    t_f = t_s = 0.0
    reps = 3
    for r in range(reps + 1):
        A.copy_(S0); b.copy_(b0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        check(lib.jb_dpotrf_batched(b"L", n, A.data_ptr(), n, n * n, info.data_ptr(), batch), "potrf_batched")
        e1.record(stream)
        check(lib.jb_dpotrs_batched(b"L", n, 1, A.data_ptr(), n, n * n, b.data_ptr(), n, n, batch), "potrs_batched")
        e2.record(stream); torch.cuda.synchronize()
        if r:
            t_f += e0.elapsed_time(e1) / reps; t_s += e1.elapsed_time(e2) / reps
    assert int(info.abs().max().item()) == 0
    tri = n * (n + 1) // 2 * 8
    by_f, by_s = 2 * tri, tri + 2 * n * 8
    return {"workload": f"{batch} SPD 32x32 Double systems, jb_dpotrf_batched + jb_dpotrs_batched (nrhs = 1), lower storage",
            "potrf_ms": t_f, "potrs_ms": t_s, "solves_per_s": batch / ((t_f + t_s) * 1e-3),
            "potrf_frac_of_hbm": batch * by_f / t_f / 1e6 / hbm_gbs, "potrs_frac_of_hbm": batch * by_s / t_s / 1e6 / hbm_gbs,
            "algorithmic_bytes_per_system": {"potrf": by_f, "potrs": by_s}}


def batched_lu_other_types(torch, lib, check, stream, hbm_gbs, batch=LU_BATCH, n=LU_N):
    """configs[2]'s call for the other element types: jb_{s,c,z}gesv_batched (getrf + getrs, nrhs = 1) on `batch` 32 x 32
    systems (csrc/batched_lu32g.cu); algorithmic bytes = A in, b in, LU out, ipiv out, x out."""
    out = {}
    for t, tdt, esz, w in (("s", torch.float32, 4, 1), ("c", torch.float32, 8, 2), ("z", torch.float64, 16, 2)):
        A0 = torch.empty(batch * n * n * w, dtype=tdt, device="cuda"); b0 = torch.empty(batch * n * w, dtype=tdt, device="cuda")
        fill = getattr(lib, f"jb_{t}fill_uniform")
        if t == "s":
            check(fill(A0.data_ptr(), n, n, batch * n, 0, 0, SEED + 7, 0, 0.0), "fill"); check(fill(b0.data_ptr(), n, n, batch, 0, 0, SEED + 8, 0, 0.0), "fill")
        else:
            check(fill(A0.data_ptr(), n, n, batch * n, 0, 0, SEED + 7), "fill"); check(fill(b0.data_ptr(), n, n, batch, 0, 0, SEED + 8), "fill")
        A, b = torch.empty_like(A0), torch.empty_like(b0)
        piv = torch.empty(batch * n, dtype=torch.int32, device="cuda"); info = torch.empty(batch, dtype=torch.int32, device="cuda")
        gesv = getattr(lib, f"jb_{t}gesv_batched")
        ms, reps = 0.0, 3
        for r in range(reps + 1):
            A.copy_(A0); b.copy_(b0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            check(gesv(n, 1, A.data_ptr(), n, n * n, piv.data_ptr(), b.data_ptr(), n, n, info.data_ptr(), batch), "gesv_batched")
            e1.record(stream); torch.cuda.synchronize()
            if r:
                ms += e0.elapsed_time(e1) / reps
        assert int(info.abs().max().item()) == 0
        by = 2 * n * n * esz + 2 * n * esz + 4 * n
        out[f"{t}gesv_batched"] = {"ms": ms, "solves_per_s": batch / (ms * 1e-3), "algorithmic_bytes_per_system": by,
                                   "frac_of_hbm": batch * by / ms / 1e6 / hbm_gbs}
        del A0, b0, A, b, piv, info
        torch.cuda.empty_cache()
    out["workload"] = f"{batch} systems of order 32, getrf + getrs with one right-hand side, Float / Complex Float / Complex Double"
    return out


def cholesky_config3(torch, lib, grid, dist, n=65536, nb=1024):
    """BASELINE.json configs[3]: dpotrf + dpotrs (nrhs = 1) of an N x N SPD Double matrix, 2-D block-cyclic over the
    ranks with NCCL panel broadcasts (jalla_b200/dist_chol.py); one warm-up pass, one timed pass, residual of the solve."""
    from jalla_b200.dist_chol import DistCholesky, DeviceOps
    ops = DeviceOps()
    ch = DistCholesky(grid, n, nb, ops)
    best = None
    for rep in range(2):
        ch.fill_spd(); ch.info = 0
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        info = ch.potrf(lookahead=True)
        e1.record()
        b = ops.empty(n); ops.fill_uniform(b, 0, n, n, 1, 0, 0, 777)
        x = b.clone()
        ch.potrs(x, 1)
        e2.record(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = (t[0].item(), t[1].item())
    res = ch.residual(x, b, 1)
    ops.finish()
    flop = n ** 3 / 3.0
    world = grid.world
    return {"workload": "configs[3]: dpotrf + dpotrs (nrhs=1), N=65536 SPD Double, 2-D block-cyclic over the ranks, NCCL panel broadcast",
            "n": n, "tile": nb, "grid": f"{grid.pr}x{grid.pc}", "n_gpus": world, "info": int(info), "potrf_ms": best[0], "potrs_ms": best[1],
            "tflops": flop / (best[0] * 1e-3) / 1e12, "frac_of_dmma_peak": flop / (best[0] * 1e-3) / 1e12 / (DMMA_PEAK_TFLOPS * world),
            "solve_residual_over_n_eps": float(res)}


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="jalla_b200")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the zgemm / sgemm side measurements")
    ap.add_argument("--only", default="", help="gemm|lu: restrict the timed work (profiling aid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from jalla_b200 import _lib
    from jalla_b200._lib import lib, check, check_last

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: jalla_b200 has no CPU path"}))
        return 1
    torch.cuda.set_device(local_rank)
    check(lib.jb_init(local_rank), "jb_init")
    stream = torch.cuda.current_stream()
    lib.jb_set_stream(C.c_void_p(stream.cuda_stream))
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout at communicator creation: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from jalla_b200 import dist as jdist
    grid = jdist.ProcessGrid(world, rank, dist)

    # ---------------------------------------------------------------- primary: dgemm N=16384
    N = N_GEMM
    gemm = jdist.ShardedGemm(grid, N, seed=SEED)       # owns this rank's blocks of A, B, C (device resident)
    flop = 2.0 * N ** 3

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.jb_kernel_launches()
    res = {}
    if args.only in ("", "gemm"):
        for _ in range(args.warmup):
            gemm.step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.jb_kernel_launches()
        t_wall0 = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            gemm.step()
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1) / max(args.steps, 1)
        gemm_launches = lib.jb_kernel_launches() - l0
        kern_ms = gemm.kernel_ms(reps=3)                 # the local GEMM kernel alone, CUDA events
        if dist is not None:
            tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        res["gemm"] = (ms, gemm_launches, kern_ms, t_wall0, t_wall1)
        err, bound = verify_columns(np, torch, gemm, max(64 // world, 8))
        v = torch.tensor([err, err / bound], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        res["verify"] = {"what": f"{max(64 // world, 8)} random columns x all rows of every rank's C block vs numpy dgemm on the host; "
                                 "pulled operands vs the regenerated inputs", "columns_total": max(64 // world, 8) * world,
                         "rel_frobenius_err_max": float(v[0].item()), "err_over_bound_max": float(v[1].item()),
                         "bound": "4*n*eps*|A||B|/|C| (north_star)", "ok": bool(v[1].item() <= 1.0)}
        assert res["verify"]["ok"], res["verify"]
    # ---------------------------------------------------------------- secondary: batched LU
    lu = None
    if args.only in ("", "lu"):
        lu = jdist.ShardedBatchedLU(grid, LU_N, LU_BATCH, seed=SEED)
        for _ in range(max(args.warmup, 1)):
            lu.step()
        barrier()
        # the factorisation overwrites its input: each step restores A,b (untimed) and brackets only the
        # solve kernel with CUDA events on the launching stream; value = batch / mean kernel time, max over ranks
        lu_kern_ms = lu.kernel_ms(reps=max(args.steps, 1))
        lu_ms = lu_kern_ms
        if dist is not None:
            tms = torch.tensor([lu_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            lu_ms = float(tms.item())
        res["lu"] = (lu_ms, lu_kern_ms)
    total_launches = lib.jb_kernel_launches() - launches0
    sampler.stop()

    hbm_peak, hbm_src = measured_peaks()
    out = {}
    if "gemm" in res:
        ms, gemm_launches, kern_ms, tw0, tw1 = res["gemm"]
        value = flop / (ms * 1e-3) / 1e12
        local_flop = flop / world
        ach = local_flop / (kern_ms * 1e-3) / 1e12
        out.update({
            "metric": "dgemm_fp64_n16384_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: FP64 dgemm N=16384 square, C=A*B (NN, alpha=1, beta=0), column-major ld=rows",
                       "n": N, "grid": f"{grid.pr}x{grid.pc}", "l2": "operands 2 GiB each >> 126 MB L2, no flush needed",
                       "inputs": "counter-based U(-1,1) keyed on global (i,j), generated on device"},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": ach / DMMA_PEAK_TFLOPS,
                         # dram__bytes_read+write of one N=16384 launch (profiles/ncu_dgemm_r1a.txt); only meaningful at N=1
                         "traffic": 86.117e9 if world == 1 else None,
                         "peak_source": "measured FP64 DMMA issue-rate peak (tools/peaks.cu, profiles/peaks_r1.jsonl); "
                                        "MEASURED_PEAKS.json holds no FP64 figure",
                         "kernel": "dgemm_dmma_kernel", "kernel_ms": kern_ms,
                         "algorithmic_flops_per_launch": local_flop},
            "gpu_launches": int(total_launches),
            "clocks": sampler.summary(tw0, tw1),
            "verify": res["verify"],
        })
    if "lu" in res:
        lu_ms, lu_kern_ms = res["lu"]
        per_rank = lu.local_batch
        ach = per_rank * LU_BYTES_PER_SYSTEM / (lu_kern_ms * 1e-3) / 1e9
        out["batched_lu"] = {
            "metric": "batched_lu_32x32_solves_per_s", "value": LU_BATCH / (lu_ms * 1e-3), "unit": "solves/s",
            "ms_per_step": lu_ms, "scaling": "strong", "dtype": "f64",
            "config": {"workload": "configs[2]: 1,000,000 independent 32x32 Double systems, getrf+getrs (jb_dgesv_batched), "
                                   "batch split across ranks, no collective", "flush": "8.4 GB working set >> L2"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         # dram__bytes_read + dram__bytes_write of the shipped kernel: 4.988 GB per 296,000 systems
                         # (profiles/ncu_lu32q_r2.txt) scaled to this launch = the algorithmic bytes
                         "traffic": per_rank * 16850.0, "peak_source": hbm_src, "kernel": "lu32q_kernel<1,1,16> (pair-lookahead)",
                         "kernel_ms": lu_kern_ms, "algorithmic_bytes_per_launch": per_rank * LU_BYTES_PER_SYSTEM,
                         "frac_of_nominal_8TBs": ach / 8000.0},
        }
        if "gemm" not in res:
            out.update({"metric": "batched_lu_32x32_solves_per_s", "value": out["batched_lu"]["value"], "unit": "solves/s",
                        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": lu_ms,
                        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                        "config": out["batched_lu"]["config"], "roofline": out["batched_lu"]["roofline"],
                        "gpu_launches": int(total_launches)})
    # ---------------------------------------------------------------- configs[4]: zgemm N=8192, sgemm N=16384 (same 2-D sharding)
    if args.only == "" and not args.no_other:
        other = {}
        if lu is not None:
            del lu
        torch.cuda.empty_cache()
        for t, n, peak, bound in (("z", 8192, DMMA_PEAK_TFLOPS, "FP64 DMMA"), ("s", 16384, FFMA_PEAK_TFLOPS, "FP32 FFMA (no TF32)"),
                                  ("c", 8192, FFMA_PEAK_TFLOPS, "FP32 FFMA, 4M complex (no TF32)")):
            g2 = jdist.ShardedGemm(grid, n, seed=SEED, t=t)
            for _ in range(2):
                g2.step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(2):
                g2.step()
            e1.record(stream)
            barrier()
            ms_o = e0.elapsed_time(e1) / 2
            if dist is not None:
                tms = torch.tensor([ms_o], device="cuda", dtype=torch.float64)
                dist.all_reduce(tms, op=dist.ReduceOp.MAX)
                ms_o = float(tms.item())
            err_o, bound_o = verify_columns(np, torch, g2, 8)
            assert err_o <= bound_o, (t, err_o, bound_o)
            tf = g2.flop / (ms_o * 1e-3) / 1e12
            other[f"{t}gemm_n{n}"] = {"tflops": tf, "ms": ms_o, "flop": g2.flop, "peak_tflops": peak * world, "frac": tf / (peak * world),
                                      "pipe": bound, "n_gpus": world, "verify_rel_frobenius_err": err_o, "verify_bound": bound_o}
            g2.close()
            del g2
            torch.cuda.empty_cache()
        out["other_configs"] = other
    # ---------------------------------------------------------------- configs[0]: the reference's own 512 x 512 case
    if world == 1 and args.only == "" and not args.no_other:
        out["config0_512"] = small_config(torch, lib, check, stream)
    # ---------------------------------------------------------------- e2e through the C ABI with host buffers
    if "gemm" in res and not args.no_e2e:
        out["e2e"] = gemm.e2e(steps=min(args.steps, 2))
        if world == 1:
            # the link-level drop-in hands over what stock Jalla allocates: pageable memory (mallocForeignPtrArray)
            n = N_GEMM
            Ah = np.empty((n, n), order="F"); Bh = np.empty((n, n), order="F"); Ch = np.empty((n, n), order="F")
            check(lib.jb_memcpy_d2h(Ah.ctypes.data, gemm.A_ptr, Ah.nbytes), "d2h"); check(lib.jb_memcpy_d2h(Bh.ctypes.data, gemm.B_ptr, Bh.nbytes), "d2h")
            dts = []
            for _ in range(2):       # the first call also allocates the library's pinned staging slots and starts its copy threads
                Ch = np.empty((n, n), order="F")          # fresh untouched pages, as matrixAlloc' hands them to matrixMult
                t0 = time.perf_counter()
                lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, Ah.ctypes.data, n, Bh.ctypes.data, n, 0.0, Ch.ctypes.data, n)
                dts.append(time.perf_counter() - t0)
                check_last("pageable e2e")
            dt = dts[1]
            out["e2e"]["pageable_host"] = {"value": 2.0 * n ** 3 / dt / 1e12, "unit": "TFLOP/s", "ms_per_step": dt * 1e3, "first_call_ms": dts[0] * 1e3,
                                           "path": "cblas_dgemm(pageable numpy A,B,C): the library's own staging through pinned slots filled by host threads (csrc/host_stage.cu)"}
            del Ah, Bh, Ch
    if "lu" in res and not args.no_e2e and "batched_lu" in out:
        lu2 = jdist.ShardedBatchedLU(grid, LU_N, LU_BATCH, seed=SEED)
        out["batched_lu"]["e2e"] = lu2.e2e(steps=2)
        del lu2
        torch.cuda.empty_cache()
    if world == 1 and args.only == "" and not args.no_other:
        out["factorisations"] = factorisations(torch, lib, check, stream, measured_peaks()[0])
        out["batched_cholesky"] = batched_cholesky(torch, lib, check, stream, measured_peaks()[0])
        out["batched_lu_other_types"] = batched_lu_other_types(torch, lib, check, stream, measured_peaks()[0])
    if world == 8 and args.only == "" and not args.no_other:
        gemm.close()
        del gemm
        torch.cuda.empty_cache()
        ch = cholesky_config3(torch, lib, grid, dist)
        if rank == 0:
            out["dpotrf_n65536"] = ch
    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    if world == 1 and not args.no_cpu and "gemm" in res:
        from oracle import cpu_ref
        ob = cpu_ref.openblas()
        cores = os.cpu_count() or 1
        cpu_ref.openblas_set_threads(cores)
        from jalla_b200 import synth
        n = 8192
        A = synth.uniform(n, n, SEED, "d"); B = synth.uniform(n, n, SEED + 1, "d"); Cm = np.empty((n, n), order="F")
        ob.gemm("d", 102, 111, 111, 1024, 1024, 1024, 1.0, A, n, B, n, 0.0, Cm, n)
        t0 = time.perf_counter()
        ob.gemm("d", 102, 111, 111, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {
            "value": 2.0 * n ** 3 / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference",
            "sample": f"one cblas_dgemm N={n} (1/8 of the step's flops) through the C boundary Jalla calls, "
                      f"OpenBLAS {ob.info.get('core')} {cores} threads", "blas": ob.info.get("config"),
            "batched_lu": cpu_batched_lu(ob, sample=200000, threads=cores)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        try:
            gemm.close()
        except NameError:
            pass
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
