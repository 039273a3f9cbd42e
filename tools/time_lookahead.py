"""dgetrf / dpotrf at large n under the look-ahead options of round 2c (timed SM cap, 512-row cooperative LU panel).
Every configuration's factors and pivots are compared bit for bit with the serial (look-ahead off) run.
Usage: time_lookahead.py [n ...]   (CUDA events, GPU only)"""
import ctypes as C, sys, json
sys.path.insert(0, "/root/repo")
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))


def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def setopts(o, reset=False):
    for k, v in o.items(): lib.jb_set_option(k.encode(), -1 if reset else v)


GETRF = [("look-ahead (default)", {})]
POTRF = [("serial", {"potrf.lookahead": 0}), ("look-ahead (default)", {})] + [("look-ahead, reserve %d, chain %d%%" % (r, c), {"potrf.reserve_sms": r, "potrf.chain_pct": c})
         for r in (24, 32, 48, 64, 96) for c in (140, 250)]
for n in [int(x) for x in sys.argv[1:]] or [8192, 16384]:
    A0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(A0)
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    tc = ev(lambda: A.copy_(A0), 3)
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 1, 0, 0.0), "fill")
    def getrf(): A.copy_(A0); assert lib.LAPACKE_dgetrf(102, n, n, A.data_ptr(), n, piv.data_ptr()) == 0
    ref = None
    for name, o in GETRF:
        setopts(o)
        t = ev(getrf, 3) - tc
        if ref is None: ref = (A.clone(), piv.clone())
        same = [bool(torch.equal(piv, ref[1])), float((A - ref[0]).abs().max() / ref[0].abs().max())]
        setopts(o, True)
        print(json.dumps({"op": "dgetrf", "n": n, "config": name, "ms": round(t, 2), "tflops": round(2 / 3 * n ** 3 / t / 1e9, 2),
                          "frac_of_37.07": round(2 / 3 * n ** 3 / t / 1e9 / 37.07, 3), "vs_serial_[pivots_equal,]max_rel_diff": same}), flush=True)
    del ref
    check(lib.jb_dfill_uniform(A0.data_ptr(), n, n, n, 0, 0, 2, 1, float(n)), "fill")
    def potrf(): A.copy_(A0); assert lib.LAPACKE_dpotrf(102, b"L", n, A.data_ptr(), n) == 0
    ref = None
    for name, o in POTRF:
        setopts(o)
        t = ev(potrf, 3) - tc
        if ref is None: ref = torch.tril(A.view(n, n).T).clone()
        same = float((torch.tril(A.view(n, n).T) - ref).abs().max() / ref.abs().max())
        setopts(o, True)
        print(json.dumps({"op": "dpotrf", "n": n, "config": name, "ms": round(t, 2), "tflops": round(1 / 3 * n ** 3 / t / 1e9, 2),
                          "frac_of_37.07": round(1 / 3 * n ** 3 / t / 1e9 / 37.07, 3), "vs_serial_[pivots_equal,]max_rel_diff": same}), flush=True)
    del ref, A, A0
    torch.cuda.empty_cache()
