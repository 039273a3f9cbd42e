import ctypes as C, sys, os
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
from jalla_b200 import dist as jdist
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
for simt in (0, 1):
    lib.jb_set_option(b"sgemm.simt", simt)
    for n in (8192, 16384):
        ms, fl = jdist.time_square_gemm("s", n, reps=3)
        print(f"sgemm simt={simt} n={n}: {ms:.3f} ms  {fl/ms/1e9:.2f} TFLOP/s", flush=True)
