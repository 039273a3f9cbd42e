"""dgetrf / dpotrf on device-resident data at large n (CUDA events, GPU only)."""
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
for n in [int(x) for x in sys.argv[1:]] or [16384]:
    A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    tc = ev(lambda: A.copy_(A0), 3)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    def getrf(): A.copy_(A0); assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    out = []
    for la in (1, 0):
        lib.jb_set_option(b"getrf.lookahead", la)
        t = ev(getrf, 3) - tc
        out.append("getrf lookahead=%d %.2f ms (%.1f TFLOP/s)" % (la, t, 2 / 3 * n ** 3 / t / 1e9))
    lib.jb_set_option(b"getrf.lookahead", -1)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    def potrf(): A.copy_(A0); assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
    for la in (1, 0):
        lib.jb_set_option(b"potrf.lookahead", la)
        t = ev(potrf, 3) - tc
        out.append("potrf lookahead=%d %.2f ms (%.1f TFLOP/s)" % (la, t, 1 / 3 * n ** 3 / t / 1e9))
    lib.jb_set_option(b"potrf.lookahead", -1)
    print("n=%d  " % n + " | ".join(out), flush=True)
