"""Host-side logic of the multi-GPU path on CPU: process-grid shapes, batch slicing, the SUMMA panel
schedule (gemm_plan) and — with torch.distributed over gloo at world_size 2 and 4 — the broadcast/compute
choreography of summa_step, with numpy standing in for the device GEMM (test-only stand-in; the product's
local GEMM is cblas_dgemm in libjalla_b200.so)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grid_shapes_and_batch_split():
    from jalla_b200.dist import grid_shape, split_batch
    assert [grid_shape(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    for batch in (0, 1, 7, 1_000_000):
        for world in (1, 2, 3, 8):
            parts = [split_batch(batch, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == batch
            pos = 0
            for s, c in parts:
                assert s == pos
                pos += c
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_gemm_plan_covers_k_exactly():
    from jalla_b200.dist import gemm_plan, grid_shape
    n = 64
    for world in (1, 2, 4, 8):
        pr, pc = grid_shape(world)
        for r in range(world):
            P = gemm_plan(n, pr, pc, r // pc, r % pc)
            ks = sorted((g["k0"], g["k1"]) for g in P["gemms"])
            assert ks[0][0] == 0 and ks[-1][1] == n and all(a[1] == b[0] for a, b in zip(ks, ks[1:]))
            assert P["gemms"][0]["beta"] == 0.0 and all(g["beta"] == 1.0 for g in P["gemms"][1:])
            for g in P["gemms"]:   # the A panels named cover the K range of the step
                assert g["a_panels"][0] * P["kc"] <= g["k0"] and (g["a_panels"][-1] + 1) * P["kc"] >= g["k1"]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from jalla_b200 import synth
    from jalla_b200.dist import ProcessGrid, gemm_plan, summa_step
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = ProcessGrid(world, rank, dist)
        P = gemm_plan(n, grid.pr, grid.pc, grid.row, grid.col)
        mb, nb, kr, kc = P["mb"], P["nb"], P["kr"], P["kc"]
        # column-major blocks as flat tensors; only the owned block is filled
        a_blocks = [torch.zeros(mb * kc, dtype=torch.float64) for _ in range(grid.pc)]
        b_blocks = [torch.zeros(kr * nb, dtype=torch.float64) for _ in range(grid.pr)]
        a_blocks[grid.col].copy_(torch.from_numpy(synth.uniform(mb, kc, 7, "d", P["row0"], grid.col * kc).T.copy().ravel()))
        b_blocks[grid.row].copy_(torch.from_numpy(synth.uniform(kr, nb, 8, "d", grid.row * kr, P["col0"]).T.copy().ravel()))
        Cacc = np.zeros((mb, nb))

        def local_gemm(g):
            nonlocal Cacc
            A_row = np.concatenate([b.numpy().reshape(kc, mb).T for b in a_blocks], axis=1)   # mb x n
            Bp = b_blocks[g["p"]].numpy().reshape(nb, kr).T                                       # kr x nb
            Cacc = g["beta"] * Cacc + A_row[:, g["k0"]:g["k1"]] @ Bp

        summa_step(P, grid, a_blocks, b_blocks, local_gemm)
        A = synth.uniform(n, n, 7, "d"); B = synth.uniform(n, n, 8, "d")
        want = (A @ B)[P["row0"]:P["row0"] + mb, P["col0"]:P["col0"] + nb]
        q.put((rank, float(np.abs(Cacc - want).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_summa_schedule_gloo(world):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 48, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world))
    assert max(e for _, e in res) < 1e-12
