"""sgemm A/B timing: the TMA-staged FFMA2 kernel with the 128 x 128 and the 256 x 128 tile, and the round-1 SIMT kernel.
python tools/cmp_sgemm.py [n ...]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
from jalla_b200 import dist as jdist
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
sizes = [int(x) for x in sys.argv[1:]] or [8192, 16384]
for name, simt, tile in (("tma 256x128", 0, 256), ("tma 128x128", 0, 128), ("simt", 1, -1)):
    lib.jb_set_option(b"sgemm.simt", simt); lib.jb_set_option(b"sgemm.tile", tile)
    for n in sizes:
        ms, fl = jdist.time_square_gemm("s", n, reps=3)
        print(f"sgemm {name} n={n}: {ms:.3f} ms  {fl/ms/1e9:.2f} TFLOP/s  ({fl/ms/1e9/72.35:.3f} of the FFMA peak)", flush=True)
