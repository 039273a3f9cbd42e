"""Times LAPACKE entry points on device-resident data (CUDA events) next to OpenBLAS on the host."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jalla_b200._lib import lib, check
from oracle import cpu_ref
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
ob = cpu_ref.openblas()
cpu_ref.openblas_set_threads(os.cpu_count())
def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
sizes = [int(x) for x in sys.argv[1:]] or [512, 2048, 8192]
for n in sizes:
    A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
    S0 = torch.empty_like(A0)
    B0 = torch.empty(n, dtype=torch.float64, device="cuda"); B = torch.empty_like(B0)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    check(lib.jb_dfill_uniform(S0.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    check(lib.jb_dfill_uniform(B0.data_ptr(), n, n, 1, 0, 0, 3, 0, 0.0), "fill")
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    reps = 3 if n >= 4096 else 10
    def getrf(): A.copy_(A0); assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    def gesv(): A.copy_(A0); B.copy_(B0); assert lib.LAPACKE_dgesv(102, n, 1, A.data_ptr(), n, piv.data_ptr(), B.data_ptr(), n) == 0
    def potrf(): A.copy_(S0); assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
    def copy(): A.copy_(A0)
    t_copy = ev(copy, reps)
    res = {k: ev(f, reps) - t_copy for k, f in (("getrf", getrf), ("gesv", gesv), ("potrf", potrf))}
    # host
    Ah = A0.cpu().numpy().reshape(n, n).T.copy(order="F"); Sh = S0.cpu().numpy().reshape(n, n).T.copy(order="F")
    ph = np.zeros(n, np.int32)
    def cpu(f):
        t0 = time.perf_counter(); f(); return (time.perf_counter() - t0) * 1e3
    c_getrf = min(cpu(lambda: ob.getrf("d", 102, n, n, Ah.copy(order="F"), n, ph)) for _ in range(2))
    c_potrf = min(cpu(lambda: ob.potrf("d", 102, "L", n, Sh.copy(order="F"), n)) for _ in range(2))
    print(f"n={n}: GPU getrf {res['getrf']:.2f} ms ({2/3*n**3/res['getrf']/1e9:.1f} GF/s)  gesv {res['gesv']:.2f} ms  potrf {res['potrf']:.2f} ms ({n**3/3/res['potrf']/1e9:.1f} GF/s) | CPU getrf {c_getrf:.2f} ms potrf {c_potrf:.2f} ms (incl. copy)", flush=True)
