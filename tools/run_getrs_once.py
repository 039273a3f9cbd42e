import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
check(lib.jb_init(0), "init")
A = torch.empty(n * n, dtype=torch.float64, device="cuda"); piv = torch.empty(n, dtype=torch.int32, device="cuda")
check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
b = torch.empty(n, dtype=torch.float64, device="cuda"); check(lib.jb_dfill_uniform(b.data_ptr(), n, n, 1, 0, 0, 5, 0, 0.0), "fill")
torch.cuda.synchronize()
lib.jb_set_option(b"marker", 1)
assert lib.LAPACKE_dgetrs(102, b"N", n, 1, A.data_ptr(), n, piv.data_ptr(), b.data_ptr(), n) == 0
torch.cuda.synchronize()
