"""Column sharding across GPUs — the B200 counterpart of the reference's `@distributed (+)` over
`Iterators.partition(𝓘, ceil(|𝓘|/nworkers()))` (src/synthesis.jl:16,23-27).

Columns are independent, so a shard is a contiguous column range and there is no data-path collective.
Unlike the reference's equal-count chunks, ranges are balanced by predicted cost (edge columns of a chain are
0.15x an interior column, corner columns of a grid 0.04x — SURVEY §8e).  Results concatenate by column; an
all-gather is only needed when every rank must hold the whole Φ (device-resident consumers)."""
from __future__ import annotations

import numpy as np


def predicted_cost(n, sum_v, T):
    """SURVEY §8d algorithmic flops of one column: n^2 * sum_t(nx_t+nu_t) + (7/3) T n^3."""
    n = np.asarray(n, dtype=np.float64)
    return n * n * np.asarray(sum_v, dtype=np.float64) + (7.0 / 3.0) * T * n ** 3


def balanced_ranges(cost, world: int):
    """Contiguous ranges [b[r], b[r+1]) over len(cost) columns with near-equal total cost."""
    cost = np.asarray(cost, dtype=np.float64)
    ncols = cost.size
    if world <= 1 or ncols == 0:
        return np.array([0] + [ncols] * max(world, 1), dtype=np.int64)
    pre = np.concatenate([[0.0], np.cumsum(np.maximum(cost, 0.0))])
    total = pre[-1]
    if total <= 0:
        return equal_ranges(ncols, world)
    targets = total * np.arange(1, world) / world
    cuts = np.searchsorted(pre, targets, side="left")
    b = np.concatenate([[0], cuts, [ncols]]).astype(np.int64)
    return np.maximum.accumulate(b)


def equal_ranges(ncols: int, world: int):
    """The reference's equal-count chunks (src/synthesis.jl:16)."""
    chunk = -(-ncols // max(world, 1))
    return np.minimum(np.arange(world + 1, dtype=np.int64) * chunk, ncols)


def chain_cost_model(Nx: int, d: int, T: int, alpha: float):
    """Predicted per-column cost of the tri-diagonal chain without running kernel 1: n = states within d+1 of j."""
    j = np.arange(Nx)
    n = np.minimum(j, d + 1) + np.minimum(Nx - 1 - j, d + 1) + 1
    sum_v = np.zeros(Nx)
    for t in range(1, T + 1):
        k = int(np.floor(alpha * (t - 1)))
        kx, ku = min(d, k), min(d + 1, k)
        sum_v += np.minimum(j, kx) + np.minimum(Nx - 1 - j, kx) + 1
        sum_v += (np.minimum(j, ku) + np.minimum(Nx - 1 - j, ku) + 1) / 3.0
    return predicted_cost(n, sum_v, T)


def assemble_columns(parts, ranges, shape):
    """Concatenate per-rank CSC pieces (each full-width with only its own columns filled) by column —
    what the reference's `(+)` reduction of disjoint-column sparse matrices amounts to (src/synthesis.jl:24)."""
    import scipy.sparse as sp
    indptr = np.zeros(shape[1] + 1, dtype=np.int64)
    idx, dat = [], []
    for r, M in enumerate(parts):
        a, b = int(ranges[r]), int(ranges[r + 1])
        M = sp.csc_matrix(M)
        lo, hi = M.indptr[a], M.indptr[b]
        cnt = np.diff(M.indptr[a:b + 1])
        indptr[a + 1:b + 1] = cnt
        idx.append(M.indices[lo:hi]); dat.append(M.data[lo:hi])
    indptr = np.cumsum(indptr)
    return sp.csc_matrix((np.concatenate(dat) if dat else np.zeros(0), np.concatenate(idx) if idx else np.zeros(0, np.int64), indptr), shape=shape)
