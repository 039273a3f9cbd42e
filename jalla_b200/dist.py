"""Multi-GPU partitioning of the hot path, one process per GPU.

Only where the work shards naturally:
  * batched small systems — contiguous batch slices per rank, NO collective;
  * one large GEMM — 2-D block distribution over a Pr x Pc process grid; rank (i,j) owns
    A[i-th row block, j-th K block], B[i-th K block, j-th column block] and computes
    C[i,j].  The K panels it does not own arrive by NCCL broadcast along its process row
    (A) and process column (B), issued up front and consumed panel by panel so the
    transfers hide under the local DMMA GEMMs (SUMMA with all panels in flight; on
    NVSwitch every peer is one hop, so the grid is chosen for compute balance only).
torch / torch.distributed are plumbing here (device buffers, streams, NCCL); every flop is
in libjalla_b200.so.  The host logic is rank-count agnostic and is exercised on CPU with
gloo at world_size 2 (tests/test_dist_cpu.py) through the `backend` hooks below.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List, Optional, Tuple

import numpy as np


def grid_shape(world: int) -> Tuple[int, int]:
    """Pr x Pc with Pr <= Pc, as square as possible: 1x1, 1x2, 2x2, 2x4 (BASELINE.json configs[1])."""
    pr = 1
    for d in range(1, int(world ** 0.5) + 1):
        if world % d == 0:
            pr = d
    return pr, world // pr


def split_batch(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [start, start+count) of a batch for `rank`; remainders go to the low ranks."""
    base, rem = divmod(batch, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


class ProcessGrid:
    def __init__(self, world: int, rank: int, dist=None):
        self.world, self.rank, self.dist = world, rank, dist
        self.pr, self.pc = grid_shape(world)
        self.row, self.col = rank // self.pc, rank % self.pc
        self.row_group = self.col_group = None
        self.row_ranks = [self.row * self.pc + q for q in range(self.pc)]
        self.col_ranks = [p * self.pc + self.col for p in range(self.pr)]
        if dist is not None and world > 1:
            # every rank must create every group, in the same order
            for r in range(self.pr):
                g = dist.new_group([r * self.pc + q for q in range(self.pc)])
                if r == self.row:
                    self.row_group = g
            for c in range(self.pc):
                g = dist.new_group([p * self.pc + c for p in range(self.pr)])
                if c == self.col:
                    self.col_group = g


def gemm_plan(n: int, pr: int, pc: int, row: int, col: int, chunks: int = 16):
    """Static description of rank (row,col)'s share of C = A*B (all sizes in elements).
    K is cut into panels that never straddle an owner's K block; panel [k0,k1) of A lives on process column k0//kc
    and of B on process row k0//kr.  Each panel is one broadcast and one local GEMM,
    so only the first panel's transfer is exposed; everything else hides under the DMMA GEMMs."""
    assert n % pr == 0 and n % pc == 0, "N must divide the process grid"
    mb, nb, kr, kc = n // pr, n // pc, n // pr, n // pc
    # panels may not straddle an ownership boundary (multiples of kr for B, of kc for A).  The first ownership
    # granule is cut geometrically (g/8, g/8, g/4, g/2) so the exposed first transfer is small; the rest are whole
    # granules, which keeps the local GEMMs long (few C read-modify-write sweeps, few partial waves).
    g = min(kr, kc)
    assert kr % g == 0 and kc % g == 0
    widths = []
    if chunks > 1 and pr * pc > 1 and g % 8 == 0:
        widths += [g // 8, g // 8, g // 4, g // 2]
    else:
        widths += [g]
    widths += [g] * (n // g - 1)
    panels, k0 = [], 0
    for t, w in enumerate(widths):
        panels.append({"t": t, "k0": k0, "k1": k0 + w, "w": w, "a_src": row * pc + k0 // kc, "b_src": (k0 // kr) * pc + col,
                       "a_mine": k0 // kc == col, "b_mine": k0 // kr == row, "beta": 0.0 if t == 0 else 1.0})
        k0 += w
    w = g
    return {"mb": mb, "nb": nb, "kr": kr, "kc": kc, "w": w, "row0": row * mb, "col0": col * nb, "panels": panels,
            "gemms": panels}


def summa_step(plan, grid: ProcessGrid, a_panels, b_panels, local_gemm):
    """One distributed product: post every panel broadcast (A panels along the process row, B panels along
    the process column, all asynchronous, in K order), then run the local GEMMs in K order, each waiting only
    for the two panels it consumes.  Backend agnostic (NCCL on GPUs; gloo in the CPU tests): `a_panels[t]` /
    `b_panels[t]` are the tensors panels land in (the owner's slots already hold its data), `local_gemm(p)`
    runs one entry of plan["panels"]."""
    dist = grid.dist
    a_work, b_work = {}, {}
    for p in plan["panels"]:
        t = p["t"]
        a_work[t] = dist.broadcast(a_panels[t], src=p["a_src"], group=grid.row_group, async_op=True) if grid.pc > 1 else None
        b_work[t] = dist.broadcast(b_panels[t], src=p["b_src"], group=grid.col_group, async_op=True) if grid.pr > 1 else None
    for p in plan["panels"]:
        t = p["t"]
        if a_work[t] is not None:
            a_work[t].wait()
        if b_work[t] is not None:
            b_work[t].wait()
        local_gemm(p)


class ShardedGemm:
    """configs[1]: one FP64 N x N x N product, 2-D sharded; a step = panel broadcasts + local GEMMs."""

    _TORCH = {"s": "float32", "d": "float64", "c": "complex64", "z": "complex128"}

    def __init__(self, grid: ProcessGrid, n: int, seed: int, t: str = "d"):
        import torch
        from ._lib import lib, check
        self.torch, self.lib, self.check = torch, lib, check
        self.grid, self.n, self.seed, self.t = grid, n, seed, t
        # measured on 2 and 8 B200s (N=16384): 8 GPUs 214.7 -> 259.2 TFLOP/s with 16 SMs left free during the four
        # GEMMs of the first K granule; 2 GPUs 61.4 -> 68.6 with 8
        env = __import__("os").environ
        self.reserve_sms = int(env.get("JB_RESERVE_SMS", "8" if grid.world <= 2 else "16"))
        self.capped_panels = int(env.get("JB_CAPPED_PANELS", "4"))
        self.esz = {"s": 4, "d": 8, "c": 8, "z": 16}[t]
        self.plan = gemm_plan(n, grid.pr, grid.pc, grid.row, grid.col, chunks=1 if grid.world == 1 else 16)
        P = self.plan
        dev = torch.device("cuda", torch.cuda.current_device())
        f64 = getattr(torch, self._TORCH[t])
        mb, nb, w = P["mb"], P["nb"], P["w"]
        self._gemm = getattr(lib, f"cblas_{t}gemm")
        if t in "sd":
            self._one, self._scal = 1.0, None
        else:   # complex scalars travel by pointer (BlasOps.hs:156,180): [1, 0, 1] -> alpha, beta=0, beta=1
            self._scal = np.array([1.0, 0.0, 1.0], dtype=np.complex64 if t == "c" else np.complex128)
        # column-major blocks held as flat device buffers: A row block mb x n (ld = mb; a K panel is a contiguous
        # slice), B as one w x nb buffer per K panel (ld = w)
        self.A_row = torch.empty(mb * n, dtype=f64, device=dev)
        self.a_panels = [self.A_row[p["k0"] * mb:p["k1"] * mb] for p in P["panels"]]
        self.b_panels = [torch.empty(p["w"] * nb, dtype=f64, device=dev) for p in P["panels"]]
        self.C = torch.empty(mb * nb, dtype=f64, device=dev)
        # generate only the panels this rank OWNS; the rest arrives by broadcast
        fill = getattr(lib, f"jb_{t}fill_uniform")
        extra = (0, 0.0) if t in "sd" else ()
        for p in P["panels"]:
            if p["a_mine"]:
                check(fill(self.a_panels[p["t"]].data_ptr(), mb, mb, p["w"], P["row0"], p["k0"], seed, *extra), "fill A")
            if p["b_mine"]:
                check(fill(self.b_panels[p["t"]].data_ptr(), p["w"], p["w"], nb, p["k0"], P["col0"], seed + 1, *extra), "fill B")
        torch.cuda.synchronize()

    @property
    def flop(self) -> float:
        return (8.0 if self.t in "cz" else 2.0) * float(self.n) ** 3

    def _local_gemm(self, p):
        P = self.plan
        if self._scal is None:
            al, be = 1.0, p["beta"]
        else:
            base, isz = self._scal.ctypes.data, self._scal.itemsize
            al, be = base, base + (2 * isz if p["beta"] else isz)
        self._gemm(102, 111, 111, P["mb"], P["nb"], p["w"], al, self.a_panels[p["t"]].data_ptr(), P["mb"],
                   self.b_panels[p["t"]].data_ptr(), p["w"], be, self.C.data_ptr(), P["mb"])

    def step(self):
        P, grid = self.plan, self.grid
        if grid.world == 1:
            for p in P["panels"]:
                self._local_gemm(p)
            return
        # The persistent DMMA GEMM owns every SM it runs on, so NCCL's broadcast kernels could only run in the gaps
        # between GEMM launches.  The GEMMs of the first K granule (a quarter of the work or less) therefore leave a
        # few SMs free; all panels land while they run, and the remaining, larger GEMMs use the whole GPU again.
        nsm = self.torch.cuda.get_device_properties(self.torch.cuda.current_device()).multi_processor_count
        n_capped = min(self.capped_panels, len(P["panels"]) - 1)

        def gemm(p):
            self.lib.jb_set_option(b"gemm.max_ctas", max(nsm - self.reserve_sms, 1) if p["t"] < n_capped else 0)
            self._local_gemm(p)
        try:
            summa_step(P, grid, self.a_panels, self.b_panels, gemm)
        finally:
            self.lib.jb_set_option(b"gemm.max_ctas", 0)

    def kernel_ms(self, reps=3) -> float:
        """Average duration of this rank's local GEMM work (all K panels) with CUDA events, no collectives."""
        torch = self.torch
        s = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for g in self.plan["panels"]:
            self._local_gemm(g)
        torch.cuda.synchronize()
        e0.record(s)
        for _ in range(reps):
            for g in self.plan["panels"]:
                self._local_gemm(g)
        e1.record(s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def verify_sample(self, samples=6):
        """Spot-check C against float128-free numpy dots of regenerated rows/columns (guards the timed path)."""
        from . import synth
        torch, P = self.torch, self.plan
        rng = np.random.default_rng(self.grid.rank)
        Cm = self.C.view(P["nb"], P["mb"])      # [col][row]
        eps = np.finfo(np.float32 if self.t in "sc" else np.float64).eps
        for _ in range(samples):
            i, j = int(rng.integers(0, P["mb"])), int(rng.integers(0, P["nb"]))
            a = synth.uniform(1, self.n, self.seed, self.t, P["row0"] + i, 0)[0]
            b = synth.uniform(self.n, 1, self.seed + 1, self.t, 0, P["col0"] + j)[:, 0]
            want = complex(np.dot(a.astype(np.complex128), b.astype(np.complex128)))
            got = complex(Cm[j, i].item())
            tol = 4 * self.n * eps * float(np.linalg.norm(a) * np.linalg.norm(b))
            if not abs(got - want) <= tol:
                raise AssertionError(f"{self.t}gemm verification failed at C[{i},{j}]: got {got}, want {want}, tol {tol}")

    def e2e(self, steps=2):
        """Same metric through the reference-facing C ABI with HOST buffers: each step copies this rank's
        operands host->device, multiplies and copies its C block back (copies inside the timed region)."""
        torch, lib, P, grid = self.torch, self.lib, self.plan, self.grid
        n = self.n
        if grid.world == 1:
            from . import _lib
            nbytes = n * n * 8
            hA, hB, hC = (lib.jb_malloc(nbytes, _lib.JB_MEM_PINNED) for _ in range(3))
            if not (hA and hB and hC):
                return {"value": None, "unit": "TFLOP/s", "error": "pinned allocation failed"}
            try:
                self.check(lib.jb_memcpy_d2h(hA, self.A_row.data_ptr(), nbytes), "d2h A")
                self.check(lib.jb_memcpy_d2h(hB, self.b_panels[0].data_ptr(), nbytes), "d2h B")
                lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n)   # warm-up (pool, streams)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    lib.cblas_dgemm(102, 111, 111, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n)   # synchronous for host outputs
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / steps
                from ._lib import check_last
                check_last("e2e dgemm")
                # spot check the host result
                c_host = np.ctypeslib.as_array(C.cast(hC, C.POINTER(C.c_double)), shape=(n * n,))
                from . import synth
                for (i, j) in ((0, 0), (n - 1, n - 1), (1234, 4321)):
                    a = synth.uniform(1, n, self.seed, "d", i, 0)[0]; b = synth.uniform(n, 1, self.seed + 1, "d", 0, j)[:, 0]
                    assert abs(c_host[i + j * n] - float(np.dot(a, b))) <= 1e-9 * n, "e2e result mismatch"
            finally:
                for h in (hA, hB, hC):
                    lib.jb_free(h)
            return {"value": 2.0 * n ** 3 / dt / 1e12, "unit": "TFLOP/s", "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes,
                    "path": "cblas_dgemm(host pinned A,B,C): panel-pipelined H2D / DMMA / D2H inside the call"}
        # multi-GPU: host shards -> device, distributed step, C block -> host
        own_a = [self.a_panels[p["t"]] for p in P["panels"] if p["a_mine"]]
        own_b = [self.b_panels[p["t"]] for p in P["panels"] if p["b_mine"]]
        hA = [torch.empty(t.numel(), dtype=torch.float64, pin_memory=True).copy_(t) for t in own_a]
        hB = [torch.empty(t.numel(), dtype=torch.float64, pin_memory=True).copy_(t) for t in own_b]
        hC = torch.empty(self.C.numel(), dtype=torch.float64, pin_memory=True)
        dist = grid.dist

        def one():
            for d, h in zip(own_a + own_b, hA + hB):
                d.copy_(h, non_blocking=True)
            self.step()
            hC.copy_(self.C, non_blocking=True)

        one(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        return {"value": 2.0 * n ** 3 / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
                "h2d_bytes_per_step": sum(t.numel() for t in hA + hB) * 8 * grid.world, "d2h_bytes_per_step": hC.numel() * 8 * grid.world,
                "path": "pinned host shards -> H2D -> NCCL panel broadcasts + cblas_dgemm panels -> D2H of the C block"}


class ShardedBatchedLU:
    """configs[2]: `batch` independent n x n Double systems, getrf+getrs (jb_dgesv_batched); each rank owns a
    contiguous slice, no communication.  The factorisation overwrites its input, so every step first restores
    A and b from pristine copies; only the solve kernel is inside the CUDA-event bracket."""

    def __init__(self, grid: ProcessGrid, n: int, batch: int, seed: int):
        import torch
        from ._lib import lib, check
        self.torch, self.lib, self.check = torch, lib, check
        self.n = n
        self.start, self.local_batch = split_batch(batch, grid.world, grid.rank)
        dev = torch.device("cuda", torch.cuda.current_device())
        nb = self.local_batch
        self.A0 = torch.empty(nb * n * n, dtype=torch.float64, device=dev)
        self.b0 = torch.empty(nb * n, dtype=torch.float64, device=dev)
        self.A, self.b = torch.empty_like(self.A0), torch.empty_like(self.b0)
        self.ipiv = torch.empty(nb * n, dtype=torch.int32, device=dev)
        self.info = torch.empty(nb, dtype=torch.int32, device=dev)
        # the whole batch is one n x (batch*n) column-major U(-1,1) matrix: system i = columns [i*n, (i+1)*n)
        check(lib.jb_dfill_uniform(self.A0.data_ptr(), n, n, nb * n, 0, self.start * n, seed, 0, 0.0), "fill A")
        check(lib.jb_dfill_uniform(self.b0.data_ptr(), n, n, nb, 0, self.start, seed + 1, 0, 0.0), "fill b")
        self.ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(64)]
        self.k = 0

    def step(self):
        n, nb, lib = self.n, self.local_batch, self.lib
        self.A.copy_(self.A0); self.b.copy_(self.b0)
        self.check(lib.jb_dgesv_batched(n, 1, self.A.data_ptr(), n, n * n, self.ipiv.data_ptr(), self.b.data_ptr(), n, n,
                                        self.info.data_ptr(), nb), "gesv_batched")

    def kernel_ms(self, reps=5) -> float:
        torch = self.torch
        s = torch.cuda.current_stream()
        tot = 0.0
        for _ in range(reps):
            self.A.copy_(self.A0); self.b.copy_(self.b0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            self.check(self.lib.jb_dgesv_batched(self.n, 1, self.A.data_ptr(), self.n, self.n * self.n, self.ipiv.data_ptr(),
                                                 self.b.data_ptr(), self.n, self.n, self.info.data_ptr(), self.local_batch), "gesv_batched")
            e1.record(s)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        assert int(self.info.abs().max().item()) == 0, "a synthetic system reported info != 0"
        return tot / reps


def time_square_gemm(t: str, n: int, reps: int = 2, seed: int = 20261019):
    """configs[4]-style single-GPU timing of one n x n x n product of element type t ('s','d','c','z'):
    NN, alpha = 1, beta = 0, inputs from the device generator.  Returns (ms, real flop count)."""
    import torch
    from ._lib import lib, check, check_last
    esz = {"s": 4, "d": 8, "c": 8, "z": 16}[t]
    bufs = [torch.empty(n * n * esz, dtype=torch.uint8, device="cuda") for _ in range(3)]
    fill = getattr(lib, f"jb_{t}fill_uniform")
    for i, b in enumerate(bufs[:2]):
        if t in "sd":
            check(fill(b.data_ptr(), n, n, n, 0, 0, seed + i, 0, 0.0), "fill")
        else:
            check(fill(b.data_ptr(), n, n, n, 0, 0, seed + i), "fill")
    gemm = getattr(lib, f"cblas_{t}gemm")
    if t in "sd":
        al, be, keep = 1.0, 0.0, None
    else:
        keep = np.array([1.0, 0.0], dtype=np.complex64 if t == "c" else np.complex128)
        al, be = keep.ctypes.data, keep.ctypes.data + keep.itemsize

    def run():
        gemm(102, 111, 111, n, n, n, al, bufs[0].data_ptr(), n, bufs[1].data_ptr(), n, be, bufs[2].data_ptr(), n)

    run(); torch.cuda.synchronize(); check_last("gemm")
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        run()
    e1.record(s); torch.cuda.synchronize()
    flop = (8.0 if t in "cz" else 2.0) * n ** 3
    return e0.elapsed_time(e1) / reps, flop
