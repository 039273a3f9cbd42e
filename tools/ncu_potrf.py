"""One dpotrf (n from argv) for a per-kernel launch list under ncu."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
check(lib.jb_init(0), "init")
A = torch.empty(n * n, dtype=torch.float64, device="cuda")
check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")   # symmetric, diagonally dominant
assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
torch.cuda.synchronize()
