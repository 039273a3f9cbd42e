"""cblas_dsyrk (lower, NoTrans, alpha = -1, beta = 1) = the trailing update of the blocked Cholesky, alone on the GPU."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ld = 16384
Cm = torch.zeros(ld * ld, dtype=torch.float64, device="cuda"); A = torch.empty(ld * 1024, dtype=torch.float64, device="cuda")
check(lib.jb_dfill_uniform(A.data_ptr(), ld, ld, 1024, 0, 0, 3, 0, 0.0), "fill")
for n in (15360, 12288, 8192, 4096):
    for k in (256, 512, 1024):
        for cap in (0, 124):
            lib.jb_set_option(b"gemm.max_ctas", cap if cap else -1)
            t = ev(lambda: lib.cblas_dsyrk(102, 122, 111, n, k, C.c_double(-1.0), A.data_ptr(), ld, C.c_double(1.0), Cm.data_ptr(), ld))
            print(json.dumps({"op": "dsyrk L,N beta=1", "n": n, "k": k, "max_ctas": cap or 148, "ms": round(t, 3), "tflops": round(n * (n + 1) * k / t / 1e9, 2),
                              "frac_of_37.07": round(n * (n + 1) * k / t / 1e9 / 37.07, 3)}), flush=True)
lib.jb_set_option(b"gemm.max_ctas", -1)
