"""LAPACKE_dgetrs with 1 / 4 / 8 right-hand sides on device-resident factors: the persistent pipelined sweep (trsv_pipe_kernel,
csrc/trsm_inv.cu) against the stepped sweep (one launch per 64 rows).  Usage: time_trsv_pipe.py [n ...]"""
import ctypes as C, sys, os
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
for n in [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096, 16384]:
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); piv = torch.empty(n, dtype=torch.int32, device="cuda")
    check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    for nrhs in (1, 4, 8):
        b0 = torch.empty(n * nrhs, dtype=torch.float64, device="cuda"); check(lib.jb_dfill_uniform(b0.data_ptr(), n, n, nrhs, 0, 0, 5, 0, 0.0), "fill")
        b = torch.empty_like(b0)
        tc = ev(lambda: b.copy_(b0))
        res = {}
        for name, opts in (("default (persistent sweep)", {}), ("stepped (trsv.pipe=0)", {"trsv.pipe": 0})):
            for k, v in opts.items(): lib.jb_set_option(k.encode(), v)
            def run(): b.copy_(b0); assert lib.LAPACKE_dgetrs(102, b"N", n, nrhs, A.data_ptr(), n, piv.data_ptr(), b.data_ptr(), n) == 0
            res[name] = round(ev(run) - tc, 4)
            for k in opts: lib.jb_set_option(k.encode(), -1)
        print(n, nrhs, res, flush=True)
