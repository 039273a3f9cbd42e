"""Times the 32x32 batched LU variants (jb_set_option lu32.variant) on 1e6 systems and checks every variant's LU, ipiv
and x bit for bit against the exact one-warp routine (lu32.v1 = 1).  Usage: time_lu32.py [variants...] (-1 = v1)."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
from jalla_b200 import dist as jdist
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
grid = jdist.ProcessGrid(1, 0, None)
batch = int(os.environ.get("LU32_BATCH", 1_000_000))
lu = jdist.ShardedBatchedLU(grid, 32, batch, seed=20261019)
lib.jb_set_option(b"lu32.v1", 1)
lu.step(); torch.cuda.synchronize()
ref = (lu.A.clone(), lu.b.clone(), lu.ipiv.clone(), lu.info.clone())
for v in (sys.argv[1:] or ["0", "3", "4", "5", "6", "7", "8", "-1"]):
    lim = -1
    if ":" in v:                     # "7:8" = variant 7 with at most 8 CTAs per SM
        v, lim = v.split(":"); lim = int(lim)
    v = int(v)
    lib.jb_set_option(b"lu32.ctas_per_sm", lim)
    if v < 0:
        lib.jb_set_option(b"lu32.v1", 1)
    else:
        lib.jb_set_option(b"lu32.v1", 0); lib.jb_set_option(b"lu32.variant", v)
    lu.ipiv.zero_(); lu.info.fill_(-5)
    lu.step(); torch.cuda.synchronize()
    same = [bool(torch.equal(x, y)) for x, y in zip((lu.A, lu.b, lu.ipiv, lu.info), ref)]
    ms = lu.kernel_ms(reps=5)
    print(json.dumps({"variant": v, "ctas_per_sm": lim, "ms": round(ms, 4), "Msolves_per_s": round(batch / ms / 1e3, 1), "GBps": round(batch * 17024 / ms / 1e6),
                      "frac_of_6538.9": round(batch * 17024 / ms / 1e6 / 6538.9, 3), "bit_identical_LU_x_ipiv_info": same}), flush=True)
