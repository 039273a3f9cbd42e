"""cblas_dgemm N=16384 from pageable numpy operands (what a stock Jalla Matrix is): wall time per call for a given number of
host copy threads (option host.copy_threads, read when the staging pool is first used).  Usage: pageable_e2e.py THREADS [N]"""
import ctypes as C, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jalla_b200._lib import lib, check
thr = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
check(lib.jb_init(0), "init")
if thr >= 0: lib.jb_set_option(b"host.copy_threads", thr)
else: lib.jb_set_option(b"host.stage", 0)
rng = np.random.default_rng(1)
A = np.asfortranarray(rng.uniform(-1, 1, (n, n))); B = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
out = []
for rep in range(3):
    Cm = np.empty((n, n), order="F")          # fresh, untouched pages: what matrixAlloc' hands to matrixMult
    t0 = time.perf_counter()
    lib.cblas_dgemm(102, 111, 111, n, n, n, C.c_double(1.0), A.ctypes.data_as(C.c_void_p), n, B.ctypes.data_as(C.c_void_p), n, C.c_double(0.0),
                    Cm.ctypes.data_as(C.c_void_p), n)
    out.append(time.perf_counter() - t0)
j = 12345
ref = A @ B[:, j]
err = float(np.linalg.norm(Cm[:, j] - ref) / np.linalg.norm(ref))
print(json.dumps({"n": n, "host_copy_threads": thr if thr >= 0 else "driver staging", "ms": [round(x * 1e3, 1) for x in out], "tflops_best": round(2 * n ** 3 / min(out) / 1e12, 2),
                  "col_rel_err": err, "cores": os.cpu_count()}), flush=True)
