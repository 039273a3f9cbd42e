"""One device-resident LAPACKE_dgesv (n from argv, default 512, nrhs = 1) for a per-kernel launch list under ncu,
and its wall / event time without a profiler when run with --time."""
import ctypes as C, sys, time
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 512
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); b0 = torch.empty(n, dtype=torch.float64, device="cuda")
check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
check(lib.jb_dfill_uniform(b0.data_ptr(), n, n, 1, 0, 0, 2, 0, 0.0), "fill")
A, b = A0.clone(), b0.clone()
piv = (C.c_int * n)()
def solve():
    A.copy_(A0); b.copy_(b0)
    assert lib.LAPACKE_dgesv(102, n, 1, A.data_ptr(), n, piv, b.data_ptr(), n) == 0
solve(); torch.cuda.synchronize()
if "--time" in sys.argv:
    for _ in range(5): solve()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): solve()
    torch.cuda.synchronize()
    print("dgesv n=%d nrhs=1 device-resident: %.3f ms per call (wall, incl. two tiny copies)" % (n, (time.perf_counter() - t0) / 50 * 1e3))
