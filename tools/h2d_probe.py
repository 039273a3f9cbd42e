"""Aggregate pinned host<->device bandwidth with every rank copying at once (torchrun): what bounds the multi-GPU e2e figure.
Each rank moves 1 GiB H2D and 0.5 GiB D2H per repetition, alone (rank 0 first) and then all ranks together."""
import os, json, time
import torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
nb_in, nb_out = 1 << 30, 1 << 29
h_in = torch.empty(nb_in, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(nb_out, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(nb_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nb_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one(reps=3, duplex=True):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if duplex:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    return (time.perf_counter() - t0) / reps
one(1)
res = {}
if world > 1:
    # rank 0 alone
    t = torch.zeros(1, device="cuda")
    if rank == 0:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); res["one_rank_alone_GBps_h2d_plus_d2h"] = round(3 * (nb_in + nb_out) / (time.perf_counter() - t0) / 1e9, 1)
    dist.barrier()
dt = one(3, True); res["all_ranks_duplex_aggregate_GBps"] = round(world * (nb_in + nb_out) / dt / 1e9, 1)
dt = one(3, False); res["all_ranks_h2d_only_aggregate_GBps"] = round(world * nb_in / dt / 1e9, 1)
if rank == 0:
    res["ranks"] = world
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
