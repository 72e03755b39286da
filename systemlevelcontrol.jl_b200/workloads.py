"""Synthetic plants and (d,T)-localized supports of BASELINE.json's configs, generated directly in Julia
`SparseMatrixCSC` layout (1-based int64) without forming sparse matrix powers.

The README builds the supports with sparse Boolean matrix powers (README.md:52-54):
    𝓢ₓ[t] = (A .≠ 0)^min(d,  ⌊α(t-1)⌋) .≠ 0            𝓢ᵤ[t] = (B₂' .≠ 0)*(A .≠ 0)^min(d+1,⌊α(t-1)⌋) .≠ 0
For the tri-diagonal chain (README.md:45-47) these are bands, written here in closed form so that the
Nx = 100 000 benchmark inputs take seconds to build; tests/test_workloads.py checks them entry-for-entry
against the matrix-power construction of the oracle at small sizes."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .synthesis import JuliaCSC


def chain_plant(Nx: int, off: float = 0.2):
    """README.md:45-47: A = I + off*superdiag - off*subdiag, B1 = I, B2 = I[:, {6n+1, 6n+2}] (1-based)."""
    A = sp.diags([np.ones(Nx), off * np.ones(Nx - 1), -off * np.ones(Nx - 1)], [0, 1, -1], format="csc", dtype=np.float64)
    st = np.arange(Nx)
    act = st[(st % 6) < 2]
    B2 = sp.csc_matrix((np.ones(act.size), act, np.arange(act.size + 1)), shape=(Nx, act.size))
    B1 = sp.identity(Nx, format="csc", dtype=np.float64)
    return A, B1, B2


def _band_csc(Nx, nrows, radius, col_lo, col_hi, row_map=None, row_ok=None, window=False):
    """CSC (Julia layout) whose column j in [col_lo,col_hi) holds the states within `radius` of j (optionally
    filtered by row_ok(state) and renamed by row_map(state)); other columns are empty.  window=True returns only the
    columns [col_lo, col_hi) (a matrix with col_hi - col_lo columns — the slice S[:, col_lo+1:col_hi] in Julia)."""
    j = np.arange(col_lo, col_hi, dtype=np.int64)
    offs = np.arange(-radius, radius + 1, dtype=np.int64)
    rows = j[:, None] + offs[None, :]
    valid = (rows >= 0) & (rows < Nx)
    if row_ok is not None:
        valid &= row_ok(rows)
    if window:
        cnt = valid.sum(axis=1).astype(np.int64)
        Nx = col_hi - col_lo
    else:
        cnt = np.zeros(Nx, np.int64)
        cnt[col_lo:col_hi] = valid.sum(axis=1)
    colptr = np.empty(Nx + 1, np.int64)
    colptr[0] = 1
    np.cumsum(cnt, out=colptr[1:])
    colptr[1:] += 1
    rv = rows[valid]
    if row_map is not None:
        rv = row_map(rv)
    rowval = np.ascontiguousarray(rv + 1, dtype=np.int64)
    return JuliaCSC(nrows, Nx, colptr, rowval, np.ones(rowval.size, np.uint8))


def chain_supports(Nx: int, d: int, T: int, alpha: float, col_lo: int = 0, col_hi: int | None = None, window: bool = False):
    """(Sx, Su): two length-T lists of JuliaCSC Bool matrices for the chain plant; only columns
    [col_lo, col_hi) are populated (a rank's column slab plus the halo it needs).  window=True: the matrices hold just
    those columns (col_hi - col_lo columns, rows still numbered in the Nx-state plant)."""
    col_hi = Nx if col_hi is None else col_hi
    Nu = int(((np.arange(Nx) % 6) < 2).sum())
    Sx, Su = [], []
    cache = {}
    for t in range(1, T + 1):
        k = int(np.floor(alpha * (t - 1)))
        kx, ku = min(d, k), min(d + 1, k)
        if ("x", kx) not in cache:
            cache[("x", kx)] = _band_csc(Nx, Nx, kx, col_lo, col_hi, window=window)
        if ("u", ku) not in cache:
            cache[("u", ku)] = _band_csc(Nx, Nu, ku, col_lo, col_hi, row_map=lambda s: 2 * (s // 6) + (s % 6),
                                         row_ok=lambda s: (s % 6) < 2, window=window)
        Sx.append(cache[("x", kx)])
        Su.append(cache[("u", ku)])
    return Sx, Su


def grid_plant(nr: int, nc: int, a: float = 0.1):
    """BASELINE config #4 (parameters fixed in DESIGN.md): row-major nr x nc grid, A = I + a*(5-point adjacency),
    every node actuated (B2 = I) — sparser actuation makes unactuated columns infeasible."""
    Nx = nr * nc
    idx = np.arange(Nx).reshape(nr, nc)
    pairs = [(idx[:, :-1].ravel(), idx[:, 1:].ravel()), (idx[:-1, :].ravel(), idx[1:, :].ravel())]
    r = np.concatenate([p for p, q in pairs] + [q for p, q in pairs])
    c = np.concatenate([q for p, q in pairs] + [p for p, q in pairs])
    A = (sp.identity(Nx, format="csc") + a * sp.csc_matrix((np.ones(r.size), (r, c)), shape=(Nx, Nx))).tocsc()
    A.sort_indices()
    I = sp.identity(Nx, format="csc", dtype=np.float64)
    return A, I, I.copy()


def grid_supports(nr: int, nc: int, d: int, T: int, alpha: float):
    """Supports of the fully actuated grid: L1 balls of radius min(d,k) (x) and min(d+1,k) (u) around each node."""
    Nx = nr * nc
    rr, cc = np.divmod(np.arange(Nx, dtype=np.int64), nc)

    def ball(radius):
        offs = [(dr, dc) for dr in range(-radius, radius + 1) for dc in range(-(radius - abs(dr)), radius - abs(dr) + 1)]
        offs.sort(key=lambda o: o[0] * nc + o[1])                       # ascending row index inside a column
        dr = np.array([o[0] for o in offs], np.int64)
        dc = np.array([o[1] for o in offs], np.int64)
        R = rr[:, None] + dr[None, :]
        Cc = cc[:, None] + dc[None, :]
        valid = (R >= 0) & (R < nr) & (Cc >= 0) & (Cc < nc)
        colptr = np.empty(Nx + 1, np.int64)
        colptr[0] = 1
        np.cumsum(valid.sum(axis=1), out=colptr[1:])
        colptr[1:] += 1
        rowval = np.ascontiguousarray((R * nc + Cc)[valid] + 1, dtype=np.int64)
        return JuliaCSC(Nx, Nx, colptr, rowval, np.ones(rowval.size, np.uint8))

    cache = {}
    Sx, Su = [], []
    for t in range(1, T + 1):
        k = int(np.floor(alpha * (t - 1)))
        for lst, rad in ((Sx, min(d, k)), (Su, min(d + 1, k))):
            if rad not in cache:
                cache[rad] = ball(rad)
            lst.append(cache[rad])
    return Sx, Su
