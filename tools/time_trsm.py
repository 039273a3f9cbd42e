"""cblas_dtrsm / LAPACKE_dgetri / dgetrs timings with the GEMM-form triangular solve on and off (trsm.inv)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n, nrhs in ((4096, 4096), (8192, 8192), (8192, 512), (512, 16384)):
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); B0 = torch.empty(n * nrhs, dtype=torch.float64, device="cuda"); B = torch.empty_like(B0)
    check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    check(lib.jb_dfill_uniform(B0.data_ptr(), n, n, nrhs, 0, 0, 3, 0, 0.0), "fill")
    tc = ev(lambda: B.copy_(B0))
    out = []
    for inv in (1, 0):
        lib.jb_set_option(b"trsm.inv", inv)
        def run(): B.copy_(B0); lib.cblas_dtrsm(102, 141, 122, 111, 131, n, nrhs, 1.0, A.data_ptr(), n, B.data_ptr(), n)
        t = ev(run) - tc
        out.append("trsm.inv=%d %.3f ms (%.1f TFLOP/s)" % (inv, t, n * n * nrhs / t / 1e9))
    print("dtrsm L,L,N,N n=%d nrhs=%d  " % (n, nrhs) + " | ".join(out), flush=True)
n = 4096
A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0); piv = torch.empty(n, dtype=torch.int32, device="cuda")
check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
for inv in (1, 0):
    lib.jb_set_option(b"trsm.inv", inv)
    def run():
        A.copy_(A0)
        assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
        assert lib.LAPACKE_dgetri(102, n, A.data_ptr(), n, piv.data_ptr()) == 0
    print("dgetrf+dgetri n=%d trsm.inv=%d: %.2f ms" % (n, inv, ev(run)), flush=True)
