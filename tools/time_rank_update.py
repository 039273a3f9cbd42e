"""Rank-K update C (m x m) -= A (m x K) * B (K x m) on the DMMA kernel for the K of the blocked factorisations:
beta = 1 (what LU / Cholesky issue) next to beta = 0 (no C read) — shows the epilogue's share."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
m = int(sys.argv[1]) if len(sys.argv) > 1 else 15872
Cm = torch.zeros(m * m, dtype=torch.float64, device="cuda")
for K in (128, 256, 512, 1024, 2048):
    A = torch.rand(m * K, dtype=torch.float64, device="cuda"); B = torch.rand(K * m, dtype=torch.float64, device="cuda")
    out = []
    for beta in (1.0, 0.0):
        f = lambda: lib.cblas_dgemm(102, 111, 111, m, m, K, C.c_double(-1.0), C.c_void_p(A.data_ptr()), m, C.c_void_p(B.data_ptr()), K,
                                    C.c_double(beta), C.c_void_p(Cm.data_ptr()), m)
        t = ev(f, 5)
        out.append("beta=%g %.3f ms %.1f TFLOP/s" % (beta, t, 2.0 * m * m * K / t / 1e9))
    print("m=%d K=%d  " % (m, K) + " | ".join(out), flush=True)
