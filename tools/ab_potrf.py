import ctypes as C, sys, os
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
for n in (8192, 16384):
    A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
    tc = ev(lambda: A.copy_(A0), 3)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    def potrf(): A.copy_(A0); assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
    for bal in (1, 0, 1, 0):
        for rsv in (-1, 16, 24):
            lib.jb_set_option(b"gemm.tri_balance", bal); lib.jb_set_option(b"potrf.reserve_sms", rsv)
            print(f"n={n} tri_balance={bal} reserve={rsv}: {ev(potrf, 4) - tc:.2f} ms", flush=True)
