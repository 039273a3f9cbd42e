"""cblas_dtrsm (Left, Lower, NoTrans, Unit) for the U12 shapes of the blocked LU, across trsm.rec_base; also laswp on the same columns."""
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
for n, nrhs in ((512, 15360), (512, 7680), (256, 7680), (128, 7680), (1024, 15360)):
    A = torch.empty(ld * n, dtype=torch.float64, device="cuda"); B0 = torch.empty(ld * nrhs, dtype=torch.float64, device="cuda"); B = torch.empty_like(B0)
    check(lib.jb_dfill_uniform(A.data_ptr(), ld, ld, n, 0, 0, 2, 0, 0.0), "fill")
    check(lib.jb_dfill_uniform(B0.data_ptr(), ld, ld, nrhs, 0, 0, 3, 0, 0.0), "fill")
    tc = ev(lambda: B.copy_(B0))
    ref = None
    for rb in (512, 256, 128):
        lib.jb_set_option(b"trsm.rec_base", rb)
        def run(): B.copy_(B0); lib.cblas_dtrsm(102, 141, 122, 111, 132, n, nrhs, 1.0, A.data_ptr(), ld, B.data_ptr(), ld)
        t = ev(run) - tc
        X = B.view(nrhs, ld)[:, :n]
        if ref is None: ref = X.clone()
        print(json.dumps({"op": "dtrsm LLNU", "n": n, "nrhs": nrhs, "ld": ld, "rec_base": rb, "ms": round(t, 3), "tflops": round(n * n * nrhs / t / 1e9, 2),
                          "max_rel_diff_vs_512": float((X - ref).abs().max() / ref.abs().max())}), flush=True)
    lib.jb_set_option(b"trsm.rec_base", -1)
