"""Runs one square GEMM of the given type/size a few times (profiling aid): python tools/run_gemm.py z 8192 2"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
from jalla_b200 import dist as jdist
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
t, n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
ms, fl = jdist.time_square_gemm(t, n, reps=reps)
print(f"{t}gemm n={n}: {ms:.3f} ms  {fl/ms/1e9:.2f} TFLOP/s")
