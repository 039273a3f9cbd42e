"""One dgetrf (n from argv, default 512) for a per-kernel launch list under ncu."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
check(lib.jb_init(0), "init")
A = torch.empty(n * n, dtype=torch.float64, device="cuda")
check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
piv = torch.empty(n, dtype=torch.int32, device="cuda")
assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
torch.cuda.synchronize()
