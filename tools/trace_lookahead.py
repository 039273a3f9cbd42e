"""Per-step timeline of the look-ahead dgetrf / dpotrf (options getrf.trace / potrf.trace, printed by the library to stderr).
Usage: trace_lookahead.py n [getrf|potrf]"""
import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
what = sys.argv[2] if len(sys.argv) > 2 else "getrf"
A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
piv = torch.empty(n, dtype=torch.int32, device="cuda")
if what == "getrf":
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    run = lambda: lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr())
else:
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    run = lambda: lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n)
A.copy_(A0); assert run() == 0; torch.cuda.synchronize()
lib.jb_set_option((what + ".trace").encode(), 1)
A.copy_(A0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); assert run() == 0; e1.record(); torch.cuda.synchronize()
print("%s n=%d traced call %.2f ms" % (what, n, e0.elapsed_time(e1)), file=sys.stderr)
