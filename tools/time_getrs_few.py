"""LAPACKE_dgetrs / dpotrs with one and four right-hand sides on device-resident factors: the bandwidth-bound sweep
(trsv.few, csrc/trsm_inv.cu) against the blocked path and the one-CTA-per-column fused kernel (n <= 4096)."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]:
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); piv = torch.empty(n, dtype=torch.int32, device="cuda")
    check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    A0 = A.clone()
    assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    for nrhs in (1, 4, 8):
        b0 = torch.empty(n * nrhs, dtype=torch.float64, device="cuda"); check(lib.jb_dfill_uniform(b0.data_ptr(), n, n, nrhs, 0, 0, 5, 0, 0.0), "fill")
        b = torch.empty_like(b0)
        tc = ev(lambda: b.copy_(b0))
        for name, opts in (("sweep", {"getrs.fused_max_n": 1}), ("sweep, plain launches", {"getrs.fused_max_n": 1, "trsv.pdl": 0}), ("fused (n <= 4096)", {"trsv.few": 0}), ("blocked", {"trsv.few": 0, "getrs.fused_max_n": 1}), ("default", {})):
            if name.startswith("fused") and n > 4096: continue
            for k, v in opts.items(): lib.jb_set_option(k.encode(), v)
            def run(): b.copy_(b0); assert lib.LAPACKE_dgetrs(102, b"N", n, nrhs, A.data_ptr(), n, piv.data_ptr(), b.data_ptr(), n) == 0
            t = ev(run) - tc
            for k in opts: lib.jb_set_option(k.encode(), -1)
            # residual of A x = b
            x = b.view(nrhs, n).T; r = A0.view(n, n).T @ x - b0.view(nrhs, n).T
            res = float(r.norm() / (A0.view(n, n).norm() * x.norm() * n * 2.2e-16))
            print(json.dumps({"op": "dgetrs", "n": n, "nrhs": nrhs, "path": name, "ms": round(t, 3), "GBps_of_factor_bytes": round(n * n * 8 / t / 1e6), "residual_over_n_eps": round(res, 4)}), flush=True)
