"""dpotrf at a size whose panel grids exceed the SMs left free during look-ahead: the factors with look-ahead on, off, and
from the separate potf2 / trsm kernels must agree (no torch: ctypes + numpy only)."""
import ctypes as C, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from jalla_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
check(lib.jb_init(0), "init")
nbytes = n * n * 8
dA0 = lib.jb_malloc(nbytes, 0); dA = lib.jb_malloc(nbytes, 0)
assert dA0 and dA
check(lib.jb_dfill_uniform(C.c_void_p(dA0), n, n, n, 0, 0, 2, 1, float(n)), "fill")
res = {}
for name, opts in (("lookahead", {}), ("serial", {b"potrf.lookahead": 0}), ("old kernels", {b"potrf.variant": 1})):
    for k, v in opts.items(): lib.jb_set_option(k, v)
    check(lib.jb_memcpy_d2d(C.c_void_p(dA), C.c_void_p(dA0), nbytes), "copy")
    t0 = time.perf_counter()
    info = lib.LAPACKE_dpotrf(102, b"L", n, C.c_void_p(dA), n)
    lib.jb_sync()
    dt = time.perf_counter() - t0
    for k in opts: lib.jb_set_option(k, -1)
    h = np.empty((n, n), dtype=np.float64, order="F")
    check(lib.jb_memcpy_d2h(h.ctypes.data_as(C.c_void_p), C.c_void_p(dA), nbytes), "d2h")
    res[name] = np.tril(h)
    print("%-12s info=%d  %.1f ms" % (name, info, dt * 1e3), flush=True)
ref = res["old kernels"]
for name in ("lookahead", "serial"):
    d = np.abs(res[name] - ref).max() / np.abs(ref).max()
    print("%-12s max |L - L_old| / max |L| = %.2e  %s" % (name, d, "OK" if d < 1e-12 else "MISMATCH"), flush=True)
