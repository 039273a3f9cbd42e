"""Times the tuned 32x32 batched LU variants (jb_set_option lu32.variant) on 1e6 systems."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
from jalla_b200 import dist as jdist
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
grid = jdist.ProcessGrid(1, 0, None)
lu = jdist.ShardedBatchedLU(grid, 32, 1_000_000, seed=20261019)
for v in [int(x) for x in (sys.argv[1:] or ["0", "2", "-1"])]:
    if v < 0:
        lib.jb_set_option(b"lu32.v1", 1)
    else:
        lib.jb_set_option(b"lu32.v1", 0); lib.jb_set_option(b"lu32.variant", v)
    lu.step(); torch.cuda.synchronize()
    ms = lu.kernel_ms(reps=5)
    print(f"variant {v}: {ms:.3f} ms  {1e6/ms*1e3/1e6:.1f} M solves/s  {1e6*17024/ms/1e6:.0f} GB/s", flush=True)
