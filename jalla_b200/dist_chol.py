"""Distributed blocked Cholesky (potrf + potrs) on a 2-D block-cyclic layout — BASELINE.json configs[3]
(N = 65536 Double SPD over 8 GPUs, NCCL panel broadcast).  One process per GPU.

Layout: square tiles of `nb`; tile (I, J) lives on rank (I mod Pr, J mod Pc) of the Pr x Pc process grid, each
rank storing its tiles as one column-major local matrix (local tile (i, j) <-> global (i*Pr + r, j*Pc + c)).
Right-looking step K (the LAPACK ?potrf recurrence, /root/reference/Numeric/Jalla/Foreign/LAPACKE.chs:733-736
binds the single-GPU entry point this generalises):
  1. the owner factorises the diagonal tile            (LAPACKE_dpotrf, device)
  2. it is broadcast down process column K mod Pc      (NCCL)
  3. that column's ranks solve their panel tiles       (cblas_dtrsm  Right/Lower/Trans)
  4. the solved panel is broadcast along process rows  (NCCL), and its transpose-role tiles down columns
  5. every rank updates its trailing tiles             (cblas_dgemm N,T per local tile column; DMMA kernel)
All arithmetic is in libjalla_b200.so; this file is choreography.  The same code runs on CPU tensors with
gloo and a numpy `ops` stand-in in tests/test_dist_cpu.py (world sizes 2 and 4), which is how the index
logic is validated without GPUs.  potrs keeps the right-hand side replicated and walks the tile column with
one small reduction + broadcast per step (latency bound; nrhs is 1 in the BASELINE config).
"""
from __future__ import annotations

from typing import List

import numpy as np

from .dist import ProcessGrid


class DeviceOps:
    """Local tile operations through the C ABI on CUDA tensors (flat float64 buffers + element offsets)."""

    def __init__(self):
        import torch
        from ._lib import lib, check, check_last
        self.torch, self.lib, self.check, self.check_last = torch, lib, check, check_last

    def empty(self, nelem):
        return self.torch.empty(max(int(nelem), 1), dtype=self.torch.float64, device="cuda")

    def zeros(self, nelem):
        return self.torch.zeros(max(int(nelem), 1), dtype=self.torch.float64, device="cuda")

    @staticmethod
    def _p(t, off):
        return t.data_ptr() + 8 * int(off)

    def fill_spd_tile(self, t, off, ld, nb, I, J, seed, diag):
        self.check(self.lib.jb_dfill_uniform(self._p(t, off), ld, nb, nb, I * nb, J * nb, seed, 1, float(diag)), "fill")

    def fill_uniform(self, t, off, ld, m, n, row0, col0, seed):
        self.check(self.lib.jb_dfill_uniform(self._p(t, off), ld, m, n, row0, col0, seed, 0, 0.0), "fill")

    def potrf(self, t, off, ld, nb):
        return int(self.lib.LAPACKE_dpotrf(102, b"L", nb, self._p(t, off), ld))

    def trsm_right_lt(self, m, nb, L, offl, ldl, B, offb, ldb):          # B <- B * L^{-T}
        self.lib.cblas_dtrsm(102, 142, 122, 112, 131, m, nb, 1.0, self._p(L, offl), ldl, self._p(B, offb), ldb)

    def trsm_left(self, trans, nb, nrhs, L, offl, ldl, B, offb, ldb):    # B <- op(L)^{-1} B
        self.lib.cblas_dtrsm(102, 141, 122, trans, 131, nb, nrhs, 1.0, self._p(L, offl), ldl, self._p(B, offb), ldb)

    def gemm(self, ta, tb, m, n, k, alpha, A, offa, lda, B, offb, ldb, beta, Cm, offc, ldc):
        self.lib.cblas_dgemm(102, ta, tb, m, n, k, alpha, self._p(A, offa), lda, self._p(B, offb), ldb, beta, self._p(Cm, offc), ldc)

    def copy(self, m, n, A, offa, lda, B, offb, ldb):
        self.check(self.lib.jb_dlacpy(111, m, n, self._p(A, offa), lda, self._p(B, offb), ldb), "lacpy")

    def finish(self):
        self.torch.cuda.synchronize()
        self.check_last("dist_chol")


class DistCholesky:
    def __init__(self, grid: ProcessGrid, n: int, nb: int, ops, seed: int = 20261019):
        assert n % nb == 0, "N must be a multiple of the tile size"
        self.g, self.n, self.nb, self.ops, self.seed = grid, n, nb, ops, seed
        self.T = n // nb
        g = grid
        self.Tr = len(range(g.row, self.T, g.pr))        # local tile rows
        self.Tc = len(range(g.col, self.T, g.pc))        # local tile cols
        self.ld = max(self.Tr * nb, 1)
        self.A = ops.empty(self.ld * max(self.Tc * nb, 1))
        self.D = ops.empty(nb * nb)
        self.info = 0

    # ---- index helpers (pure functions of K and the grid: every rank can evaluate them for any rank)
    def first_row_above(self, r, I):
        """first local tile-row index i on process row r with global tile row i*Pr + r > I"""
        return max(0, -(-(I + 1 - r) // self.g.pr))

    def first_row_from(self, r, I):
        """first local tile-row index with global tile row >= I"""
        return max(0, -(-(I - r) // self.g.pr))

    def first_col_above(self, c, J):
        return max(0, -(-(J + 1 - c) // self.g.pc))

    def local_rows(self, r):
        return len(range(r, self.T, self.g.pr))

    def tile_off(self, i, j):
        return i * self.nb + j * self.nb * self.ld

    # ---- synthetic SPD input of configs[3]: a_ij = a_ji ~ U(-1,1) keyed on (min,max), a_ii += N
    def fill_spd(self):
        g, nb = self.g, self.nb
        for j in range(self.Tc):
            for i in range(self.Tr):
                self.ops.fill_spd_tile(self.A, self.tile_off(i, j), self.ld, nb, i * g.pr + g.row, j * g.pc + g.col, self.seed, self.n)

    def _bcast(self, t, src, group):
        if group is not None:
            self.g.dist.broadcast(t, src=src, group=group)

    def potrf(self) -> int:
        g, nb, ops, T = self.g, self.nb, self.ops, self.T
        dist = g.dist
        for K in range(T):
            kr, kc = K % g.pr, K % g.pc
            lkr, lk = K // g.pr, K // g.pc
            owner = kr * g.pc + kc
            # 1. diagonal tile
            if g.rank == owner:
                info = ops.potrf(self.A, self.tile_off(lkr, lk), self.ld, nb)
                if info and not self.info:
                    self.info = K * nb + info if info > 0 else info
                ops.copy(nb, nb, self.A, self.tile_off(lkr, lk), self.ld, self.D, 0, nb)
            # 2. + 3. process column kc: receive the diagonal factor, solve the panel below it
            i_start = self.first_row_above(g.row, K)
            m1 = (self.Tr - i_start) * nb
            if g.col == kc:
                if g.pr > 1:
                    self._bcast(self.D, owner, g.col_group)
                if m1 > 0:
                    ops.trsm_right_lt(m1, nb, self.D, 0, nb, self.A, self.tile_off(i_start, lk), self.ld)
            if K == T - 1:
                break
            # 4. row broadcast of the solved panel (packed m1 x nb)
            Wrow = ops.empty(max(m1, 1) * nb)
            if g.col == kc and m1 > 0:
                ops.copy(m1, nb, self.A, self.tile_off(i_start, lk), self.ld, Wrow, 0, m1)
            if g.pc > 1 and m1 > 0:
                self._bcast(Wrow, g.row * g.pc + kc, g.row_group)
            # 5. transpose-role tiles: L_JK for my local tile columns J > K, gathered down my process column
            j_start = self.first_col_above(g.col, K)
            n1 = (self.Tc - j_start) * nb
            Wcol = ops.empty(max(n1, 1) * nb)
            for rp in range(g.pr):
                tiles = [J for J in range(K + 1, T) if J % g.pc == g.col and J % g.pr == rp]
                if not tiles:
                    continue
                i_start_rp = self.first_row_above(rp, K)
                m1_rp = (self.local_rows(rp) - i_start_rp) * nb
                if g.pr == 1:
                    for J in tiles:
                        ops.copy(nb, nb, Wrow, (J // g.pr - i_start_rp) * nb, m1_rp, Wcol, (J // g.pc - j_start) * nb, n1)
                    continue
                buf = ops.empty(len(tiles) * nb * nb)
                ldbuf = len(tiles) * nb
                if g.row == rp:
                    for q, J in enumerate(tiles):
                        ops.copy(nb, nb, Wrow, (J // g.pr - i_start_rp) * nb, m1_rp, buf, q * nb, ldbuf)
                self._bcast(buf, rp * g.pc + g.col, g.col_group)
                for q, J in enumerate(tiles):
                    ops.copy(nb, nb, buf, q * nb, ldbuf, Wcol, (J // g.pc - j_start) * nb, n1)
            # 6. trailing update, one GEMM per local tile column (rows from the diagonal tile down)
            for j in range(j_start, self.Tc):
                J = j * g.pc + g.col
                i0 = self.first_row_from(g.row, J)
                mrows = (self.Tr - i0) * nb
                if mrows <= 0:
                    continue
                ops.gemm(111, 112, mrows, nb, nb, -1.0, Wrow, (i0 - i_start) * nb, m1, Wcol, (j - j_start) * nb, n1, 1.0,
                         self.A, self.tile_off(i0, j), self.ld)
        # agree on info
        if dist is not None and g.world > 1:
            import torch
            t = torch.tensor([self.info if self.info > 0 else 2 ** 31 - 1], dtype=torch.int64, device=self.A.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            v = int(t.item())
            self.info = 0 if v == 2 ** 31 - 1 else v
        return self.info

    def potrs(self, x, nrhs: int = 1):
        """Solves A X = B in place on the replicated n x nrhs tensor `x` (column-major, ld = n) from the factor.
        Forward: every rank keeps the partial sums of its own tiles' contributions in local-row layout (the ranks of
        process column 0 start from b), so one sum over the process row yields b_K - sum_{J<K} L_KJ y_J.  Backward:
        the ranks of process column K mod Pc form y_K - sum_{I>K} L_IK^T x_I with one sum down the column."""
        g, nb, ops, T, n = self.g, self.nb, self.ops, self.T, self.n
        dist = g.dist
        multi = dist is not None and g.world > 1
        zero = ops.zeros(nb * nrhs)
        accl = ops.zeros(self.ld * nrhs)
        if g.col == 0:
            for i in range(self.Tr):
                ops.copy(nb, nrhs, x, (i * g.pr + g.row) * nb, n, accl, i * nb, self.ld)
        seg = ops.empty(nb * nrhs)
        # ---- forward: L y = b
        for K in range(T):
            kr, kc = K % g.pr, K % g.pc
            lkr, lk = K // g.pr, K // g.pc
            owner = kr * g.pc + kc
            if g.row == kr:
                ops.copy(nb, nrhs, accl, lkr * nb, self.ld, seg, 0, nb)
                if multi and g.pc > 1:
                    dist.all_reduce(seg, group=g.row_group)
            if g.rank == owner:
                ops.trsm_left(111, nb, nrhs, self.A, self.tile_off(lkr, lk), self.ld, seg, 0, nb)
            if multi:
                dist.broadcast(seg, src=owner)
            ops.copy(nb, nrhs, seg, 0, nb, x, K * nb, n)
            i_start = self.first_row_above(g.row, K)
            m1 = (self.Tr - i_start) * nb
            if g.col == kc and m1 > 0:                       # accl[rows > K] -= L_IK * y_K
                ops.gemm(111, 111, m1, nrhs, nb, -1.0, self.A, self.tile_off(i_start, lk), self.ld, seg, 0, nb, 1.0, accl, i_start * nb, self.ld)
        # ---- backward: L^T x = y
        xl = ops.zeros(self.ld * nrhs)                      # x restricted to my local tile rows
        for K in range(T - 1, -1, -1):
            kr, kc = K % g.pr, K % g.pc
            lkr, lk = K // g.pr, K // g.pc
            owner = kr * g.pc + kc
            i_start = self.first_row_above(g.row, K)
            m1 = (self.Tr - i_start) * nb
            if g.col == kc:
                if g.rank == owner:
                    ops.copy(nb, nrhs, x, K * nb, n, seg, 0, nb)
                else:
                    ops.copy(nb, nrhs, zero, 0, nb, seg, 0, nb)
                if m1 > 0:                                   # seg -= sum over my rows I > K of L_IK^T x_I
                    ops.gemm(112, 111, nb, nrhs, m1, -1.0, self.A, self.tile_off(i_start, lk), self.ld, xl, i_start * nb, self.ld, 1.0, seg, 0, nb)
                if multi and g.pr > 1:
                    dist.all_reduce(seg, group=g.col_group)
            if g.rank == owner:
                ops.trsm_left(112, nb, nrhs, self.A, self.tile_off(lkr, lk), self.ld, seg, 0, nb)
            if multi:
                dist.broadcast(seg, src=owner)
            ops.copy(nb, nrhs, seg, 0, nb, x, K * nb, n)
            if g.row == kr:
                ops.copy(nb, nrhs, seg, 0, nb, xl, lkr * nb, self.ld)
        return x

    # ---- residual of the solve against the regenerated input: |b - A x| / (|A|_F |x| n eps) without ever holding A
    def residual(self, x, b, nrhs: int = 1) -> float:
        import torch
        g, nb, ops, T, n = self.g, self.nb, self.ops, self.T, self.n
        dist = g.dist
        tile = ops.empty(nb * nb)
        y = ops.zeros(n * nrhs)
        for I in range(T):
            for J in range(T):
                if (I * T + J) % g.world != g.rank:
                    continue
                ops.fill_spd_tile(tile, 0, nb, nb, I, J, self.seed, n)
                ops.gemm(111, 111, nb, nrhs, nb, 1.0, tile, 0, nb, x, J * nb, n, 1.0, y, I * nb, n)
        if dist is not None and g.world > 1:
            dist.all_reduce(y)
        r = (b[: n * nrhs] - y[: n * nrhs]).norm().item()
        anorm = (float(n) ** 3 + float(n) ** 2 / 3.0) ** 0.5     # |A|_F of the generator: n on the diagonal + U(-1,1)
        return r / (anorm * x[: n * nrhs].norm().item() * n * np.finfo(np.float64).eps)
