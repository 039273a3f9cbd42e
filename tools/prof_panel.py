"""Per-phase cycle sums of the register LU panel kernel (debug option getrf.panel_prof)."""
import ctypes as C, sys
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
for n in [int(x) for x in sys.argv[1:]] or [512, 4096]:
    A = torch.empty(n * n, dtype=torch.float64, device="cuda")
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    for rep in range(2):
        check(lib.jb_dfill_uniform(A.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
        lib.jb_set_option(b"getrf.panel_prof", rep)
        assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    lib.jb_set_option(b"getrf.panel_prof", 0)
torch.cuda.synchronize()
