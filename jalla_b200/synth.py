"""numpy twin of the counter-based synthetic-input generator in csrc/aux_kernels.cu
(jb_?fill_uniform): element (i, j) is a pure function of (seed, global i, global j), so
any sharding of a matrix regenerates identical data."""
from __future__ import annotations

import numpy as np

_M1, _M2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)
_G1, _G2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F)


def _mix64(x):
    x = x ^ (x >> np.uint64(30)); x = x * _M1
    x = x ^ (x >> np.uint64(27)); x = x * _M2
    return x ^ (x >> np.uint64(31))


def _key2(seed, i, j):
    with np.errstate(over="ignore"):
        x = _mix64(np.uint64(seed) + i * _G1)
        return _mix64(x + j * _G2)


def uniform(m, n, seed, dtype="d", row0=0, col0=0, sym=False, diag=0.0):
    """m x n matrix (Fortran order) identical to jb_?fill_uniform on the device."""
    i = (np.arange(m, dtype=np.uint64) + np.uint64(row0))[:, None]
    j = (np.arange(n, dtype=np.uint64) + np.uint64(col0))[None, :]
    i, j = np.broadcast_arrays(i, j)

    def ud(h):
        return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0

    def uf(h):
        return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0) * np.float32(2.0)
                - np.float32(1.0)).astype(np.float32)

    if dtype in ("d", "s"):
        ki, kj = (np.minimum(i, j), np.maximum(i, j)) if sym else (i, j)
        h = _key2(seed, ki, kj)
        a = ud(h) if dtype == "d" else uf(h)
        if sym:
            a = a + np.where(i == j, a.dtype.type(diag), a.dtype.type(0))
        return np.asfortranarray(a)
    with np.errstate(over="ignore"):
        hr, hi = _key2(seed, i, np.uint64(2) * j), _key2(seed, i, np.uint64(2) * j + np.uint64(1))
    if dtype == "z":
        return np.asfortranarray(ud(hr) + 1j * ud(hi))
    return np.asfortranarray((uf(hr) + 1j * uf(hi)).astype(np.complex64))
