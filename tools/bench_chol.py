#!/usr/bin/env python3
"""BASELINE.json configs[3]: blocked Cholesky potrf + potrs on an N x N SPD Double matrix, 2-D block-cyclic over
the launched ranks (torchrun, one rank per GPU; also runs single-GPU).  Prints one JSON line (rank 0).
    python -m torch.distributed.run --nproc-per-node 8 tools/bench_chol.py --size 65536 --tile 512"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jalla_b200._lib import lib, check
from jalla_b200.dist import ProcessGrid
from jalla_b200.dist_chol import DistCholesky, DeviceOps

ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=65536)
ap.add_argument("--tile", dest="nb", type=int, default=512)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--no-lookahead", action="store_true")
args = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
check(lib.jb_init(lr), "init")
lib.jb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
grid = ProcessGrid(world, rank, dist)
ops = DeviceOps()
ch = DistCholesky(grid, args.n, args.nb, ops)
best = None
for rep in range(args.reps + 1):          # first pass is the warm-up
    ch.fill_spd(); ch.info = 0
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    info = ch.potrf(lookahead=not args.no_lookahead)
    e1.record()
    b = ops.empty(args.n); ops.fill_uniform(b, 0, args.n, args.n, 1, 0, 0, 777)
    x = b.clone()
    ch.potrs(x, 1)
    e2.record(); torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], device="cuda", dtype=torch.float64)
    if dist is not None: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rep > 0 and (best is None or t[0].item() < best[0]):
        best = (t[0].item(), t[1].item())
res = ch.residual(x, b, 1)
ops.finish()
if rank == 0:
    n = args.n
    flop = n ** 3 / 3.0
    print(json.dumps({"metric": "dpotrf_tflops", "n": n, "nb": args.nb, "n_gpus": world, "grid": f"{grid.pr}x{grid.pc}", "info": info,
                      "potrf_ms": best[0], "potrs_ms": best[1], "value": flop / (best[0] * 1e-3) / 1e12, "unit": "TFLOP/s",
                      "frac_of_dmma_peak": flop / (best[0] * 1e-3) / 1e12 / (37.07 * world),
                      "solve_residual_over_n_eps": res, "launches": int(lib.jb_kernel_launches())}), flush=True)
if dist is not None:
    dist.destroy_process_group()
