"""One LAPACKE_dgetrf + a few LAPACKE_dgetrs (nrhs = 1) at n (default 8192) on device data: the command ncu wraps to capture trsv_step_kernel."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = torch.empty(n * n, dtype=torch.float64, device="cuda"); piv = torch.empty(n, dtype=torch.int32, device="cuda"); b = torch.empty(n, dtype=torch.float64, device="cuda")
check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill"); check(lib.jb_dfill_uniform(b.data_ptr(), n, n, 1, 0, 0, 5, 0, 0.0), "fill")
assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
for _ in range(2):
    assert lib.LAPACKE_dgetrs(102, b"N", n, 1, A.data_ptr(), n, piv.data_ptr(), b.data_ptr(), n) == 0
torch.cuda.synchronize()
print("ok", float(b.abs().max()))
