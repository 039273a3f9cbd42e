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
        for r, owned_first in [(r, o) for r in range(world) for o in (False, True)]:
            P = gemm_plan(n, pr, pc, r // pc, r % pc, chunks=8, owned_first=owned_first)
            ks = sorted((g["k0"], g["k1"]) for g in P["panels"])
            assert ks[0][0] == 0 and ks[-1][1] == n and all(a[1] == b[0] for a, b in zip(ks, ks[1:]))
            if not owned_first:      # the collective path needs the same (ascending) K order on every rank
                assert ks == [(g["k0"], g["k1"]) for g in P["panels"]]
            else:                    # the pull path starts with what the rank owns
                cls = [0 if (g["a_mine"] and g["b_mine"]) else (1 if (g["a_mine"] or g["b_mine"]) else 2) for g in P["panels"]]
                assert cls == sorted(cls)
            assert P["panels"][0]["beta"] == 0.0 and all(g["beta"] == 1.0 for g in P["panels"][1:])
            for g in P["panels"]:   # a panel never straddles a granule, and its owners sit in my row / column
                assert g["k0"] // P["g"] == (g["k1"] - 1) // P["g"]
                assert g["a_src"] // pc == r // pc and g["b_src"] % pc == r % pc
                assert g["a_mine"] == (g["a_src"] == r) and g["b_mine"] == (g["b_src"] == r)
            # skewed ownership: every rank owns one granule of BOTH operands, and it comes first on the pull path
            both = [g for g in P["panels"] if g["a_mine"] and g["b_mine"]]
            assert sum(g["w"] for g in both) == P["g"]
            if owned_first and world > 1:
                assert P["panels"][0]["a_mine"] and P["panels"][0]["b_mine"]
        # every (row block, granule) of A and (granule, column block) of B has exactly one owner
        owners_a, owners_b = {}, {}
        for r in range(world):
            P = gemm_plan(n, pr, pc, r // pc, r % pc, chunks=1)
            for g in P["panels"]:
                if g["a_mine"]:
                    owners_a.setdefault((r // pc, g["k0"]), []).append(r)
                if g["b_mine"]:
                    owners_b.setdefault((g["k0"], r % pc), []).append(r)
        assert all(len(v) == 1 for v in owners_a.values()) and len(owners_a) == pr * pc
        assert all(len(v) == 1 for v in owners_b.values()) and len(owners_b) == pc * pc


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
        P = gemm_plan(n, grid.pr, grid.pc, grid.row, grid.col, chunks=6)
        mb, nb = P["mb"], P["nb"]
        # column-major panels as flat tensors; only the owned panels are filled
        a_panels = [torch.zeros(mb * p["w"], dtype=torch.float64) for p in P["panels"]]
        b_panels = [torch.zeros(p["w"] * nb, dtype=torch.float64) for p in P["panels"]]
        for p in P["panels"]:
            if p["a_mine"]:
                a_panels[p["t"]].copy_(torch.from_numpy(synth.uniform(mb, p["w"], 7, "d", P["row0"], p["k0"]).T.copy().ravel()))
            if p["b_mine"]:
                b_panels[p["t"]].copy_(torch.from_numpy(synth.uniform(p["w"], nb, 8, "d", p["k0"], P["col0"]).T.copy().ravel()))
        Cacc = np.zeros((mb, nb))

        def local_gemm(p):
            nonlocal Cacc
            Ap = a_panels[p["t"]].numpy().reshape(p["w"], mb).T      # mb x w
            Bp = b_panels[p["t"]].numpy().reshape(nb, p["w"]).T      # w x nb
            Cacc = p["beta"] * Cacc + Ap @ Bp

        summa_step(P, grid, a_panels, b_panels, local_gemm)
        A = synth.uniform(n, n, 7, "d"); B = synth.uniform(n, n, 8, "d")
        want = (A @ B)[P["row0"]:P["row0"] + mb, P["col0"]:P["col0"] + nb]
        q.put((rank, float(np.abs(Cacc - want).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_summa_schedule_gloo(world):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 64, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world))
    assert max(e for _, e in res) < 1e-12


# ------------------------------------------------------------------ distributed Cholesky choreography
class NumpyOps:
    """CPU stand-in for jalla_b200.dist_chol.DeviceOps (test-only): same interface on CPU torch tensors."""

    def __init__(self):
        import torch
        self.torch = torch

    def empty(self, nelem):
        return self.torch.full((max(int(nelem), 1),), float("nan"), dtype=self.torch.float64)

    def zeros(self, nelem):
        return self.torch.zeros(max(int(nelem), 1), dtype=self.torch.float64)

    @staticmethod
    def _v(t, off, ld, m, n):
        a = t.numpy()
        return np.lib.stride_tricks.as_strided(a[off:], shape=(m, n), strides=(8, 8 * ld)) if m * n else np.zeros((m, n))

    def fill_spd_tile(self, t, off, ld, nb, I, J, seed, diag):
        from jalla_b200 import synth
        self._v(t, off, ld, nb, nb)[...] = synth.uniform(nb, nb, seed, "d", I * nb, J * nb, sym=True, diag=float(diag))

    def potrf(self, t, off, ld, nb):
        v = self._v(t, off, ld, nb, nb)
        try:
            L = np.linalg.cholesky(np.tril(v) + np.tril(v, -1).T)
        except np.linalg.LinAlgError:
            return 1
        v[np.tril_indices(nb)] = L[np.tril_indices(nb)]
        return 0

    def trsm_right_lt(self, m, nb, L, offl, ldl, B, offb, ldb):
        Lm = np.tril(self._v(L, offl, ldl, nb, nb)); Bv = self._v(B, offb, ldb, m, nb)
        Bv[...] = np.linalg.solve(Lm, Bv.T).T

    def trsm_left(self, trans, nb, nrhs, L, offl, ldl, B, offb, ldb):
        Lm = np.tril(self._v(L, offl, ldl, nb, nb)); Bv = self._v(B, offb, ldb, nb, nrhs)
        Bv[...] = np.linalg.solve(Lm if trans == 111 else Lm.T, Bv)

    def gemm(self, ta, tb, m, n, k, alpha, A, offa, lda, B, offb, ldb, beta, Cm, offc, ldc):
        Av = self._v(A, offa, lda, m, k) if ta == 111 else self._v(A, offa, lda, k, m).T
        Bv = self._v(B, offb, ldb, k, n) if tb == 111 else self._v(B, offb, ldb, n, k).T
        Cv = self._v(Cm, offc, ldc, m, n)
        Cv[...] = alpha * (Av @ Bv) + (beta * Cv if beta != 0 else 0.0)

    def copy(self, m, n, A, offa, lda, B, offb, ldb):
        self._v(B, offb, ldb, m, n)[...] = self._v(A, offa, lda, m, n)

    def finish(self):
        pass


def _chol_worker(rank, world, port, n, nb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from jalla_b200 import synth
    from jalla_b200.dist import ProcessGrid
    from jalla_b200.dist_chol import DistCholesky
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = ProcessGrid(world, rank, dist if world > 1 else None)
        ch = DistCholesky(grid, n, nb, NumpyOps(), seed=11)
        ch.fill_spd()
        info = ch.potrf()
        A = synth.uniform(n, n, 11, "d", sym=True, diag=float(n))
        Lref = np.linalg.cholesky(A)
        ops = ch.ops
        err = 0.0
        for j in range(ch.Tc):
            for i in range(ch.Tr):
                I, J = i * grid.pr + grid.row, j * grid.pc + grid.col
                if I >= J:
                    tile = ops._v(ch.A, ch.tile_off(i, j), ch.ld, nb, nb)
                    ref = Lref[I * nb:(I + 1) * nb, J * nb:(J + 1) * nb]
                    err = max(err, float(np.abs((np.tril(tile) if I == J else tile) - ref).max()))
        b = synth.uniform(n, 1, 12, "d")[:, 0]
        x = torch.from_numpy(b.copy())
        ch.potrs(x, 1)
        xerr = float(np.abs(x.numpy() - np.linalg.solve(A, b)).max())
        res = ch.residual(x, torch.from_numpy(b.copy()), 1)
        q.put((rank, info, err, xerr, res))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nb", [(1, 48, 16), (2, 80, 16), (4, 96, 16), (8, 176, 16)])
def test_dist_cholesky_choreography_gloo(world, n, nb):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chol_worker, args=(r, world, port, n, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, info, err, xerr, r in res:
        assert info == 0 and err < 1e-11 and xerr < 1e-11 and r < 10, (rank, info, err, xerr, r)
