"""dpotrf / dgetrf time against the outer block width (options potrf.nb2 / getrf.nb2), look-ahead on."""
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
for n in [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192, 16384]:
    A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    tc = ev(lambda: A.copy_(A0), 3)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    def potrf(): A.copy_(A0); assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
    out = []
    for nb2 in (128, 256, 384, 512, 1024):
        lib.jb_set_option(b"potrf.nb2", nb2)
        out.append("%d: %.2f" % (nb2, ev(potrf, 3) - tc))
    lib.jb_set_option(b"potrf.nb2", 0)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    def getrf(): A.copy_(A0); assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    out2 = []
    for nb2 in (128, 256, 384, 512, 1024):
        lib.jb_set_option(b"getrf.nb2", nb2)
        out2.append("%d: %.2f" % (nb2, ev(getrf, 3) - tc))
    lib.jb_set_option(b"getrf.nb2", 0)
    print("n=%d potrf ms by nb2  %s | getrf ms by nb2  %s" % (n, "  ".join(out), "  ".join(out2)), flush=True)
