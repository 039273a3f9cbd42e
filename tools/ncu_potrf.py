"""One dpotrf (n from argv; further argv: option=value pairs for jb_set_option) for a per-kernel launch list under ncu."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
check(lib.jb_init(0), "init")
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    lib.jb_set_option(k.encode(), int(v))
A = torch.empty(n * n, dtype=torch.float64, device="cuda")
check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")   # symmetric, diagonally dominant
assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
torch.cuda.synchronize()
