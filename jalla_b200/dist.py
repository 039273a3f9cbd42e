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


def gemm_plan(n: int, pr: int, pc: int, row: int, col: int, chunks: int = 16, owned_first: bool = False, split_owned: bool = False):
    """Static description of rank (row,col)'s share of C = A*B (all sizes in elements).

    K is cut into pc granules of g = n/pc columns (pr divides pc on the grids used: 1x2, 2x2, 2x4).  Ownership is SKEWED
    like Cannon's initial alignment, so that every rank holds one granule of BOTH operands and can start multiplying
    without waiting for anybody: with r = pc/pr,
        A[row block i, granule q]   lives on process column (q - i*r) mod pc   -> rank (i,j) owns A granule (j + i*r) mod pc,
        B[granule q, column block j] lives on process row   ((q - j) mod pc) // r -> rank (i,j) owns B granules (j + i*r + 0..r-1) mod pc.
    A panel is one transfer and one local GEMM.  owned_first (the pull path): the rank's own granule first, then those it
    holds one operand of, then the rest; collectives (the NCCL / gloo path) need the same ascending order on every rank.
    The first granule is cut finer only where that helps: geometrically (g/8, g/8, g/4, g/2) when the first product has to
    wait for a transfer, in four equal parts when split_owned (the end-to-end path streams the own operands from the host
    in that order).  The K order differs per rank: C agrees to rounding, not bit for bit."""
    assert n % pr == 0 and n % pc == 0 and pc % pr == 0, "N must divide the process grid (and Pr must divide Pc)"
    mb, nb = n // pr, n // pc
    g, r, G = n // pc, pc // pr, pc
    a_col = lambda i, q: (q - i * r) % pc                 # process column owning A[i, q]
    b_row = lambda q, j: ((q - j) % pc) // r              # process row owning B[q, j]
    granules = []
    for q in range(G):
        a_mine, b_mine = a_col(row, q) == col, b_row(q, col) == row
        granules.append((0 if (a_mine and b_mine) else (1 if (a_mine or b_mine) else 2), q))
    multi = pr * pc > 1
    if owned_first and chunks > 1 and multi:
        granules.sort()
    else:
        granules = [(1 if multi else 0, q) for _, q in granules]
    panels = []
    for idx, (cls, q) in enumerate(granules):
        k0 = q * g
        widths = [g]
        if idx == 0 and chunks > 1 and multi and g % 8 == 0:
            if cls == 0 and split_owned:
                widths = [g // 4] * 4
            elif cls != 0:
                widths = [g // 8, g // 8, g // 4, g // 2]
        for w in widths:
            t = len(panels)
            a_src, b_src = row * pc + a_col(row, q), b_row(q, col) * pc + col
            panels.append({"t": t, "k0": k0, "k1": k0 + w, "w": w, "a_src": a_src, "b_src": b_src,
                           "a_mine": a_src == row * pc + col, "b_mine": b_src == row * pc + col, "beta": 0.0 if t == 0 else 1.0})
            k0 += w
    return {"mb": mb, "nb": nb, "g": g, "w": g, "row0": row * mb, "col0": col * nb, "panels": panels, "gemms": panels}


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
        self.mode = env.get("JB_SUMMA", "pull") if grid.world > 1 else "local"
        self.plan = gemm_plan(n, grid.pr, grid.pc, grid.row, grid.col, chunks=1 if grid.world == 1 else 16, owned_first=self.mode == "pull")
        # same buffers, finer first granule: the end-to-end path streams the own operands from the host in this order
        self.plan_e2e = gemm_plan(n, grid.pr, grid.pc, grid.row, grid.col, chunks=1 if grid.world == 1 else 16, owned_first=self.mode == "pull", split_owned=True)
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
        # slice) and B column block n x nb (ld = n; a K panel is a row range, i.e. a strided sub-matrix) — layouts that
        # do not depend on how a rank cuts K into panels, because peers read each other's blocks.
        # With more than one rank the two slabs come from jb_malloc (plain cudaMalloc, exportable over CUDA IPC): the
        # ranks of a process row map each other's A slab, those of a process column each other's B slab, and every
        # panel a rank does not own is PULLED from its owner by a copy-engine transfer over NVLink — no collective
        # kernel competes with the persistent GEMM for SMs (round 1 reserved 8-16 SMs for NCCL's broadcast kernels).
        elt = self.esz
        if self.mode == "pull":
            from . import _lib
            self._slabs = [lib.jb_malloc(mb * n * elt, _lib.JB_MEM_DEVICE), lib.jb_malloc(n * nb * elt, _lib.JB_MEM_DEVICE)]
            assert all(self._slabs), "device allocation failed"
            self.A_ptr, self.B_ptr = self._slabs
            self.A_row = None
        else:
            self.A_row = torch.empty(mb * n, dtype=f64, device=dev)
            self._B_slab = torch.empty(n * nb, dtype=f64, device=dev)
            self.A_ptr, self.B_ptr = self.A_row.data_ptr(), self._B_slab.data_ptr()
        self.a_ptrs = [self.A_ptr + p["k0"] * mb * elt for p in P["panels"]]
        self.b_ptrs = [self.B_ptr + p["k0"] * elt for p in P["panels"]]
        self.ldb = n
        if self.mode != "pull":
            self.a_panels = [self.A_row[p["k0"] * mb:p["k1"] * mb] for p in P["panels"]]
            self.b_panels = None      # (the collective path keeps its own contiguous panel tensors, see step())
        self.C = torch.empty(mb * nb, dtype=f64, device=dev)
        # generate only the panels this rank OWNS; the rest arrives from its owner
        fill = getattr(lib, f"jb_{t}fill_uniform")
        extra = (0, 0.0) if t in "sd" else ()
        for p in P["panels"]:
            if p["a_mine"]:
                check(fill(self.a_ptrs[p["t"]], mb, mb, p["w"], P["row0"], p["k0"], seed, *extra), "fill A")
            if p["b_mine"]:
                check(fill(self.b_ptrs[p["t"]], n, p["w"], nb, p["k0"], P["col0"], seed + 1, *extra), "fill B")
        torch.cuda.synchronize()
        if self.mode == "pull":
            self._open_peers()

    def _open_peers(self):
        """Exchange the CUDA IPC handles of the two slabs and map the peers this rank pulls from."""
        torch, lib, grid = self.torch, self.lib, self.grid
        mine = (C.c_char * 128)()
        self.check(lib.jb_ipc_export(self.A_ptr, C.addressof(mine)), "ipc export A")
        self.check(lib.jb_ipc_export(self.B_ptr, C.addressof(mine) + 64), "ipc export B")
        t = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(t) for _ in range(grid.world)]
        grid.dist.all_gather(allh, t)
        self.peer_a, self.peer_b, self._mapped = {}, {}, []
        for r in sorted({p["a_src"] for p in self.plan["panels"]} | {p["b_src"] for p in self.plan["panels"]}):
            if r == grid.rank:
                continue
            raw = bytes(allh[r].cpu().numpy().tobytes())
            if r in grid.row_ranks:
                ptr = lib.jb_ipc_open(raw[:64]); assert ptr, lib.jb_last_error_string()
                self.peer_a[r] = ptr; self._mapped.append(ptr)
            if r in grid.col_ranks:
                ptr = lib.jb_ipc_open(raw[64:]); assert ptr, lib.jb_last_error_string()
                self.peer_b[r] = ptr; self._mapped.append(ptr)
        self.s_a, self.s_b = torch.cuda.Stream(), torch.cuda.Stream()
        self.ev_a = [torch.cuda.Event() for _ in self.plan_e2e["panels"]]
        self.ev_b = [torch.cuda.Event() for _ in self.plan_e2e["panels"]]
        self.ev_done = torch.cuda.Event()
        self.ev_done.record(torch.cuda.current_stream())
        grid.dist.barrier()

    def close(self):
        if getattr(self, "mode", "") == "pull":
            self.torch.cuda.synchronize()
            self.grid.dist.barrier()            # nobody is still reading this rank's slabs
            for ptr in self._mapped:
                self.lib.jb_ipc_close(ptr)
            self._mapped = []
            self.grid.dist.barrier()
            for ptr in self._slabs:
                self.lib.jb_free(ptr)
            self._slabs = []
            self.mode = "closed"

    @property
    def flop(self) -> float:
        return (8.0 if self.t in "cz" else 2.0) * float(self.n) ** 3

    def _local_gemm(self, p, P=None):
        P = P or self.plan
        if self._scal is None:
            al, be = 1.0, p["beta"]
        else:
            base, isz = self._scal.ctypes.data, self._scal.itemsize
            al, be = base, base + (2 * isz if p["beta"] else isz)
        if getattr(self, "ldb_panels", False):       # collective path: contiguous panel tensors
            a_ptr, b_ptr, ldb = self.a_ptrs[p["t"]], self.b_ptrs[p["t"]], p["w"]
        else:
            a_ptr, b_ptr, ldb = self.A_ptr + p["k0"] * P["mb"] * self.esz, self.B_ptr + p["k0"] * self.esz, self.ldb
        self._gemm(102, 111, 111, P["mb"], P["nb"], p["w"], al, a_ptr, P["mb"], b_ptr, ldb, be, self.C.data_ptr(), P["mb"])

    def step(self):
        P, grid = self.plan, self.grid
        if grid.world == 1:
            for p in P["panels"]:
                self._local_gemm(p)
            return
        if self.mode == "pull":
            return self._step_pull()
        # NCCL fallback (JB_SUMMA=nccl): the persistent DMMA GEMM owns every SM it runs on, so NCCL's broadcast kernels
        # could only run in the gaps between GEMM launches.  The GEMMs of the first K granule therefore leave a few SMs
        # free; all panels land while they run, and the remaining, larger GEMMs use the whole GPU again.
        nsm = self.torch.cuda.get_device_properties(self.torch.cuda.current_device()).multi_processor_count
        n_capped = min(self.capped_panels, len(P["panels"]) - 1)

        def gemm(p):
            self.lib.jb_set_option(b"gemm.max_ctas", max(nsm - self.reserve_sms, 1) if p["t"] < n_capped else 0)
            self._local_gemm(p)
        if self.b_panels is None:     # contiguous receive buffers for the broadcasts; owned panels are packed into them
            f64 = getattr(self.torch, self._TORCH[self.t])
            self.b_panels = [self.torch.empty(p["w"] * P["nb"], dtype=f64, device=self.C.device) for p in P["panels"]]
            for p in P["panels"]:
                if p["b_mine"]:
                    self.check(self.lib.jb_memcpy2d_async(self.b_panels[p["t"]].data_ptr(), p["w"] * self.esz, self.b_ptrs[p["t"]], self.n * self.esz,
                                                          p["w"] * self.esz, P["nb"], self.torch.cuda.current_stream().cuda_stream), "pack B")
            self.b_ptrs = [t.data_ptr() for t in self.b_panels]
            self.ldb_panels = True
        try:
            summa_step(P, grid, self.a_panels, self.b_panels, gemm)
        finally:
            self.lib.jb_set_option(b"gemm.max_ctas", 0)

    def _step_pull(self, own_ready=None, peers_ready=None, plan=None):
        """One distributed product with copy-engine panel pulls: every panel this rank does not own is copied from its
        owner's slab (peer-mapped, NVLink) on one of two side streams, in K order; the local GEMM of a panel waits only
        for its own two copies.  The operands are inputs (nobody writes them during the product), so no rank has to
        wait for another: the only cross-rank ordering is the barrier around the timed region."""
        torch, lib, P = self.torch, self.lib, plan or self.plan
        mb, nb, elt = P["mb"], P["nb"], self.esz
        cur = torch.cuda.current_stream()
        self.s_a.wait_event(self.ev_done); self.s_b.wait_event(self.ev_done)   # the previous product has read its panels
        if peers_ready is not None:      # end-to-end path: the owners' operands are still arriving from their hosts
            self.s_a.wait_event(peers_ready); self.s_b.wait_event(peers_ready)
        for p in P["panels"]:
            t = p["t"]
            if not p["a_mine"]:
                off = p["k0"] * mb * elt
                self.check(lib.jb_memcpy_async(self.A_ptr + off, self.peer_a[p["a_src"]] + off, p["w"] * mb * elt, self.s_a.cuda_stream), "pull A")
                self.ev_a[t].record(self.s_a)
            if not p["b_mine"]:
                off, pitch = p["k0"] * elt, self.n * elt
                self.check(lib.jb_memcpy2d_async(self.B_ptr + off, pitch, self.peer_b[p["b_src"]] + off, pitch, p["w"] * elt, nb, self.s_b.cuda_stream), "pull B")
                self.ev_b[t].record(self.s_b)
        for p in P["panels"]:
            t = p["t"]
            if not p["a_mine"]:
                cur.wait_event(self.ev_a[t])
            if not p["b_mine"]:
                cur.wait_event(self.ev_b[t])
            if own_ready is not None and (p["a_mine"] or p["b_mine"]):
                cur.wait_event(own_ready[t])
            self._local_gemm(p, P)
        self.ev_done.record(cur)

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
                self.check(lib.jb_memcpy_d2h(hA, self.A_ptr, nbytes), "d2h A")
                self.check(lib.jb_memcpy_d2h(hB, self.B_ptr, nbytes), "d2h B")
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
        # multi-GPU: each rank's own blocks of A and B start in pinned host memory; per step they go host -> device
        # (copy engine), a barrier tells the peers they may pull, the distributed product runs, and the C block goes
        # device -> host.  A second barrier after the product keeps the next step's H2D from overwriting panels a peer
        # is still pulling.
        P = self.plan_e2e
        mb, nb, elt = P["mb"], P["nb"], self.esz
        # pinned host copies of the panels this rank owns, packed in the order the plan consumes them
        a_host, b_host, a_bytes, b_bytes = {}, {}, 0, 0
        for p in P["panels"]:
            if p["a_mine"]:
                a_host[p["t"]] = a_bytes; a_bytes += p["w"] * mb * elt
            if p["b_mine"]:
                b_host[p["t"]] = b_bytes; b_bytes += p["w"] * nb * elt
        c_bytes = mb * nb * elt
        hA = torch.empty(max(a_bytes, 16), dtype=torch.uint8, pin_memory=True); hB = torch.empty(max(b_bytes, 16), dtype=torch.uint8, pin_memory=True)
        hC = torch.empty(c_bytes, dtype=torch.uint8, pin_memory=True)
        cs = torch.cuda.current_stream().cuda_stream
        for p in P["panels"]:
            if p["a_mine"]:
                self.check(lib.jb_memcpy_async(hA.data_ptr() + a_host[p["t"]], self.A_ptr + p["k0"] * mb * elt, p["w"] * mb * elt, cs), "d2h A")
            if p["b_mine"]:      # compact host panel (w x nb, ld = w)
                self.check(lib.jb_memcpy2d_async(hB.data_ptr() + b_host[p["t"]], p["w"] * elt, self.B_ptr + p["k0"] * elt, n * elt, p["w"] * elt, nb, cs), "d2h B")
        torch.cuda.synchronize()
        dist = grid.dist
        cur = torch.cuda.current_stream()
        s_in, s_out, s_bar = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        ev_own = [torch.cuda.Event() for _ in P["panels"]]
        ev_all_in, ev_ready, ev_c = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        token = torch.zeros(1, device="cuda")

        def one():
            # own operands host -> device in the order the plan consumes them (the own granule first, in four parts): a
            # local product starts as soon as its slices are resident
            s_in.wait_stream(cur)
            for p in P["panels"]:
                if p["a_mine"]:
                    self.check(lib.jb_memcpy_async(self.A_ptr + p["k0"] * mb * elt, hA.data_ptr() + a_host[p["t"]], p["w"] * mb * elt, s_in.cuda_stream), "h2d A")
                if p["b_mine"]:
                    self.check(lib.jb_memcpy2d_async(self.B_ptr + p["k0"] * elt, n * elt, hB.data_ptr() + b_host[p["t"]], p["w"] * elt,
                                                     p["w"] * elt, nb, s_in.cuda_stream), "h2d B")
                ev_own[p["t"]].record(s_in)
            ev_all_in.record(s_in)
            s_bar.wait_event(ev_all_in)
            with torch.cuda.stream(s_bar):
                dist.all_reduce(token)             # every rank's operands are resident: peers may pull
                ev_ready.record(s_bar)
            self._step_pull(own_ready=ev_own, peers_ready=ev_ready, plan=P)
            ev_c.record(cur)
            s_out.wait_event(ev_c)
            self.check(lib.jb_memcpy_async(hC.data_ptr(), self.C.data_ptr(), c_bytes, s_out.cuda_stream), "d2h C")
            dist.all_reduce(token)                 # every rank has pulled what it needs from this rank
            cur.wait_stream(s_out)

        one(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        return {"value": self.flop / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
                "h2d_bytes_per_step": (a_bytes + b_bytes) * grid.world, "d2h_bytes_per_step": c_bytes * grid.world,
                "path": "pinned host blocks -> H2D in consumption order (products on owned panels start as they land) -> barrier -> "
                        "copy-engine panel pulls over NVLink + cblas_dgemm panels -> D2H of the C block"}


class ShardedBatchedLU:
    """configs[2]: `batch` independent n x n Double systems, getrf+getrs (jb_dgesv_batched); each rank owns a
    contiguous slice, no communication.  The factorisation overwrites its input, so every step first restores
    A and b from pristine copies; only the solve kernel is inside the CUDA-event bracket."""

    def __init__(self, grid: ProcessGrid, n: int, batch: int, seed: int):
        import torch
        from ._lib import lib, check
        self.torch, self.lib, self.check = torch, lib, check
        self.n, self.grid = n, grid
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


def _batched_lu_e2e(self, steps=2, chunk=31_250):
    """configs[2] end to end: the systems start and end in pinned HOST memory.  The batch moves in chunks of `chunk`
    systems through three streams (H2D / jb_dgesv_batched / D2H of LU, ipiv and x), so the copies of neighbouring
    chunks overlap the kernel; PCIe carries 2 x 8.5 GB per 10^6 systems and is what bounds this figure.  What is not
    overlapped is the first chunk's H2D and the last chunk's D2H, so chunks are small (256 MB: 4.7 ms each way; with
    1 GB chunks the two ends were 38 ms of a 194 ms step)."""
    torch, lib, n = self.torch, self.lib, self.n
    nb = self.local_batch
    hA = torch.empty(nb * n * n, dtype=torch.float64, pin_memory=True); hb = torch.empty(nb * n, dtype=torch.float64, pin_memory=True)
    hP = torch.empty(nb * n, dtype=torch.int32, pin_memory=True)
    hA_out = torch.empty(nb * n * n, dtype=torch.float64, pin_memory=True); hx = torch.empty(nb * n, dtype=torch.float64, pin_memory=True)
    s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    nch = (nb + chunk - 1) // chunk
    dA = [torch.empty(chunk * n * n, dtype=torch.float64, device="cuda") for _ in range(2)]
    db = [torch.empty(chunk * n, dtype=torch.float64, device="cuda") for _ in range(2)]
    dP = [torch.empty(chunk * n, dtype=torch.int32, device="cuda") for _ in range(2)]
    dI = torch.empty(nb, dtype=torch.int32, device="cuda")
    ev_in = [torch.cuda.Event() for _ in range(nch)]; ev_k = [torch.cuda.Event() for _ in range(nch)]; ev_out = [torch.cuda.Event() for _ in range(nch)]
    saved = lib.jb_get_stream()

    def one():
        for c in range(nch):
            lo, cnt, q = c * chunk, min(chunk, nb - c * chunk), c & 1
            if c >= 2:
                s_in.wait_event(ev_out[c - 2])                      # the buffer pair is free again
            with torch.cuda.stream(s_in):
                dA[q][:cnt * n * n].copy_(hA[lo * n * n:(lo + cnt) * n * n], non_blocking=True)
                db[q][:cnt * n].copy_(hb[lo * n:(lo + cnt) * n], non_blocking=True)
                ev_in[c].record(s_in)
            s_k.wait_event(ev_in[c])
            lib.jb_set_stream(C.c_void_p(s_k.cuda_stream))
            self.check(lib.jb_dgesv_batched(n, 1, dA[q].data_ptr(), n, n * n, dP[q].data_ptr(), db[q].data_ptr(), n, n,
                                            dI.data_ptr() + 4 * lo, cnt), "gesv_batched")
            ev_k[c].record(s_k)
            s_out.wait_event(ev_k[c])
            with torch.cuda.stream(s_out):
                hA_out[lo * n * n:(lo + cnt) * n * n].copy_(dA[q][:cnt * n * n], non_blocking=True)
                hx[lo * n:(lo + cnt) * n].copy_(db[q][:cnt * n], non_blocking=True)
                hP[lo * n:(lo + cnt) * n].copy_(dP[q][:cnt * n], non_blocking=True)
                ev_out[c].record(s_out)
        s_out.synchronize()

    try:
        # the device generator's systems, staged once into the pinned buffers (untimed)
        for c in range(nch):
            lo, cnt = c * chunk, min(chunk, nb - c * chunk)
            hA[lo * n * n:(lo + cnt) * n * n].copy_(self.A0[lo * n * n:(lo + cnt) * n * n]); hb[lo * n:(lo + cnt) * n].copy_(self.b0[lo * n:(lo + cnt) * n])
        torch.cuda.synchronize()
        one()                                        # warm-up
        x_first = hx[:n].clone()
        if self.grid.dist is not None:
            self.grid.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    finally:
        lib.jb_set_stream(C.c_void_p(saved))
    tms = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if self.grid.dist is not None:
        self.grid.dist.all_reduce(tms, op=self.grid.dist.ReduceOp.MAX)
    dt = float(tms.item())
    total = nb * self.grid.world if self.grid.dist is not None else nb
    return {"value": total / dt, "unit": "solves/s", "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": total * (n * n + n) * 8, "d2h_bytes_per_step": total * ((n * n + n) * 8 + n * 4),
            "path": f"pinned host systems -> H2D -> jb_dgesv_batched -> D2H of LU, ipiv, x; chunks of {chunk} systems on three streams",
            "finite_first_solution": bool(torch.isfinite(x_first).all().item())}


ShardedBatchedLU.e2e = _batched_lu_e2e


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
