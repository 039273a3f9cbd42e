"""Roofline numbers of the batched entry points other than the headline dgesv n = 32: algorithmic bytes per system (inputs read
once + outputs written once) / CUDA-event time / measured HBM copy bandwidth (MEASURED_PEAKS.json).  10^6 systems each."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
check(lib.jb_init(0), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
HBM = 6538.9
batch = int(os.environ.get("BATCH", 1_000_000))
for kv in os.environ.get("JB_OPTIONS", "").split(","):      # e.g. JB_OPTIONS=lu32g.shape=24
    if "=" in kv:
        lib.jb_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1]))
DT = {"s": (torch.float32, 4, 1), "d": (torch.float64, 8, 1), "c": (torch.float32, 8, 2), "z": (torch.float64, 16, 2)}

def ev(fn, prep, reps=3):
    prep(); fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        prep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps

def fill(t, buf, rows, cols, seed, sym=0, diag=0.0):
    f = getattr(lib, f"jb_{t}fill_uniform")
    if t in "sd":
        check(f(buf.data_ptr(), rows, rows, cols, 0, 0, seed, sym, float(diag)), "fill")
    else:
        check(f(buf.data_ptr(), rows, rows, cols, 0, 0, seed), "fill")

CASES = (("d", 32), ("s", 32), ("z", 32), ("c", 32), ("d", 16), ("d", 24))
if os.environ.get("CASES"):              # e.g. CASES=s32,z32,c32
    CASES = tuple((c[0], int(c[1:])) for c in os.environ["CASES"].split(","))
for t, n in CASES:
    tdt, esz, w = DT[t]
    A0 = torch.empty(batch * n * n * w, dtype=tdt, device="cuda"); b0 = torch.empty(batch * n * w, dtype=tdt, device="cuda")
    A, b = torch.empty_like(A0), torch.empty_like(b0)
    piv = torch.empty(batch * n, dtype=torch.int32, device="cuda"); info = torch.empty(batch, dtype=torch.int32, device="cuda")
    fill(t, A0, n, batch * n, 11); fill(t, b0, n, batch, 12)
    prep = lambda: (A.copy_(A0), b.copy_(b0))
    gesv = getattr(lib, f"jb_{t}gesv_batched"); getrf = getattr(lib, f"jb_{t}getrf_batched"); getrs = getattr(lib, f"jb_{t}getrs_batched")
    rows = []
    ms = ev(lambda: check(gesv(n, 1, A.data_ptr(), n, n * n, piv.data_ptr(), b.data_ptr(), n, n, info.data_ptr(), batch), "gesv"), prep)
    rows.append((f"{t}gesv_batched", ms, 2 * n * n * esz + 2 * n * esz + 4 * n))
    ms = ev(lambda: check(getrf(n, A.data_ptr(), n, n * n, piv.data_ptr(), info.data_ptr(), batch), "getrf"), prep)
    rows.append((f"{t}getrf_batched", ms, 2 * n * n * esz + 4 * n))
    check(getrf(n, A.data_ptr(), n, n * n, piv.data_ptr(), info.data_ptr(), batch), "getrf"); torch.cuda.synchronize()
    ms = ev(lambda: check(getrs(n, 1, A.data_ptr(), n, n * n, piv.data_ptr(), b.data_ptr(), n, n, batch), "getrs"), lambda: b.copy_(b0))
    rows.append((f"{t}getrs_batched", ms, n * n * esz + 2 * n * esz + 4 * n))
    if t in "sd" and n == 32:
        # SPD systems: A = G + n I with G symmetric U(-1,1) inside every 32 x 32 block is not what the generator makes for a
        # batch (it keys on global indices), so build them as B B^T / n + I per block on the device
        G = A0.view(batch, n, n)
        S = torch.bmm(G, G.transpose(1, 2)) / n + torch.eye(n, dtype=tdt, device="cuda")
        S0 = S.contiguous().view(-1)
        potrf = getattr(lib, f"jb_{t}potrf_batched"); potrs = getattr(lib, f"jb_{t}potrs_batched")
        ms = ev(lambda: check(potrf(b"L", n, A.data_ptr(), n, n * n, info.data_ptr(), batch), "potrf"), lambda: A.copy_(S0))
        assert int(info.abs().max().item()) == 0
        rows.append((f"{t}potrf_batched", ms, 2 * (n * (n + 1) // 2) * esz))
        ms = ev(lambda: check(potrs(b"L", n, 1, A.data_ptr(), n, n * n, b.data_ptr(), n, n, batch), "potrs"), lambda: b.copy_(b0))
        rows.append((f"{t}potrs_batched", ms, (n * (n + 1) // 2) * esz + 2 * n * esz))
    for name, ms, bytes_per in rows:
        gbs = batch * bytes_per / ms / 1e6
        print(json.dumps({"entry": name, "n": n, "batch": batch, "ms": round(ms, 3), "Msystems_per_s": round(batch / ms / 1e3, 1), "algorithmic_bytes_per_system": bytes_per,
                          "GBps": round(gbs), "frac_of_6538.9": round(gbs / HBM, 3)}), flush=True)
    del A0, b0, A, b, piv, info
    torch.cuda.empty_cache()
