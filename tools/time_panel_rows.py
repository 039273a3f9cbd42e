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
for n in (1024, 2048, 4096, 8192):
    A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    tc = ev(lambda: A.copy_(A0), 5)
    out = {}
    for name, opts in (("default", {}), ("panel_rows=512", {"getrf.panel_rows": 512}), ("panel_rows=128", {"getrf.panel_rows": 128})):
        for k, v in opts.items(): lib.jb_set_option(k.encode(), v)
        def f(): A.copy_(A0); assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
        out[name] = round(ev(f, 5) - tc, 3)
        for k in opts: lib.jb_set_option(k.encode(), -1)
    print(n, out, flush=True)
