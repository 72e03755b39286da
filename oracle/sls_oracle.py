"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A literal numpy/scipy restatement of the column-wise localized H2 state-feedback
SLS synthesis of aaltoKEPO/SystemLevelControl.jl (v1.1.0-DEV):

    SLS_H2                     <- /root/reference/src/synthesis.jl:11-32
    _SLS_H2 (per-column loop)  <- /root/reference/src/synthesis.jl:34-72
    sparsity_dim_reduction     <- /root/reference/src/reduction.jl:11-27
    objective (squared H2)     <- /root/reference/src/synthesis.jl:76-86
    Plant(A,B1,B2) defaults    <- /root/reference/src/types/GeneralizedPlant.jl:105-110
    README plant / supports    <- /root/reference/README.md:43-54
    README closed-loop sim     <- /root/reference/README.md:62-71

PARITY STATUS
  * index sets / masks (integer work): PINNED by the reference's own known-answer test
    test/reduction_test.jl:10-24 (tests/test_oracle.py::test_reduction_known_answer).
  * Phi values (floating point): **parity unpinned**.  The reference never tests SLS_H2
    (test/runtests.jl:10-13) and its arithmetic lives in un-vendored third-party packages
    (JuMP 1.10.0, MathOptInterface 1.14.1, Ipopt.jl 1.2.1 / Ipopt 3.14.10, MUMPS 5.5.1;
    Manifest.toml:173-183,202-206,249-253,265-269) that are not runnable here (no Julia).
    The per-column problem is a strictly convex equality-constrained QP, so its minimiser is
    unique; this oracle computes that exact minimiser (what Ipopt converges to, to its
    tol=1e-8) with a dense SVD-based minimum-norm / KKT solve.  It is anchored by
    reference-independent invariants (achievability residuals, support containment, the
    README closed-loop identity) checked in tests/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "chain_plant", "grid_plant", "make_supports", "Plant", "sparsity_dim_reduction",
    "column_problem", "solve_column", "SLS_H2", "SLS_H2_general", "solve_group_general", "cost_blocks", "simulate_closed_loop",
    "achievability_residual",
]


# --------------------------------------------------------------------------------------
# plants and supports (README.md:43-54)
# --------------------------------------------------------------------------------------
def chain_plant(Nx: int, off: float = 0.2, jitter: float = 0.0, seed: int = 0):
    """README chain: A = I + off*superdiag - off*subdiag (README.md:45), B1 = I (:46),
    B2 = I[:, {6n+1, 6n+2}] clipped to Nx (:47; 1-based).  `jitter` perturbs the
    off-diagonals by (1 + jitter*xi), xi~U(-1,1), pattern unchanged."""
    up = off * np.ones(Nx - 1)
    lo = -off * np.ones(Nx - 1)
    if jitter:
        rng = np.random.default_rng(seed)
        up = up * (1 + jitter * rng.uniform(-1, 1, Nx - 1))
        lo = lo * (1 + jitter * rng.uniform(-1, 1, Nx - 1))
    A = sp.diags([np.ones(Nx), up, lo], [0, 1, -1], format="csc", dtype=np.float64)
    cols = []
    n = 0
    while True:
        a, b = 6 * n + 1, 6 * n + 2          # 1-based actuator states
        if a > Nx:
            break
        cols.append(a - 1)
        if b <= Nx:
            cols.append(b - 1)
        n += 1
    B2 = sp.identity(Nx, format="csc", dtype=np.float64)[:, cols].tocsc()
    B1 = sp.identity(Nx, format="csc", dtype=np.float64)
    return A, B1, B2


def grid_plant(nr: int, nc: int, a: float = 0.1, actuation: str = "checker"):
    """2-D grid (BASELINE config #4, parameters fixed in SURVEY §8d): row-major nodes,
    A = I + a*(5-point adjacency), B2 = selection of checkerboard nodes ((r+c) even)."""
    Nx = nr * nc
    idx = np.arange(Nx).reshape(nr, nc)
    rows, cols = [], []
    for (sl_a, sl_b) in (((slice(None), slice(0, -1)), (slice(None), slice(1, None))),
                         ((slice(0, -1), slice(None)), (slice(1, None), slice(None)))):
        p, q = idx[sl_a].ravel(), idx[sl_b].ravel()
        rows += [p, q]
        cols += [q, p]
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    Adj = sp.csc_matrix((np.ones(rows.size), (rows, cols)), shape=(Nx, Nx))
    A = (sp.identity(Nx, format="csc") + a * Adj).tocsc()
    A.sort_indices()
    rr, cc = np.divmod(np.arange(Nx), nc)
    if actuation == "checker":
        sel = np.nonzero((rr + cc) % 2 == 0)[0]
    elif actuation == "full":
        sel = np.arange(Nx)
    else:
        raise ValueError(actuation)
    B2 = sp.identity(Nx, format="csc", dtype=np.float64)[:, sel].tocsc()
    B1 = sp.identity(Nx, format="csc", dtype=np.float64)
    return A, B1, B2


def _pat(M) -> sp.csc_matrix:
    """`M .≠ 0` on a SparseMatrixCSC: Bool pattern of the stored entries that are nonzero
    (stored zeros give `false` and are dropped)."""
    M = sp.csc_matrix(M)
    P = sp.csc_matrix((M.data != 0, M.indices.copy(), M.indptr.copy()), shape=M.shape)
    P.eliminate_zeros()
    P.sort_indices()
    return P.astype(np.int64)


def _pat_power(P: sp.csc_matrix, p: int) -> sp.csc_matrix:
    """`(A .≠ 0)^p` for integer-valued p; p = 0 gives the identity.
    Only the pattern of the result is used (`.≠ 0` follows), so entries are clamped to 1 to
    avoid integer overflow of the path counts."""
    n = P.shape[0]
    R = sp.identity(n, format="csc", dtype=np.int64)
    for _ in range(int(p)):
        R = (R @ P).tocsc()
        R.data[:] = 1
    return R


def make_supports(A, B2, d: int, T: int, alpha: float):
    """README.md:53-54:
        Sx[t] = (A .≠ 0)^min(d,   floor(alpha*(t-1))) .≠ 0
        Su[t] = (B2' .≠ 0)*(A .≠ 0)^min(d+1, floor(alpha*(t-1))) .≠ 0      t = 1..T
    Returns two lists of T boolean CSC matrices with sorted indices, all stored values True."""
    pA = _pat(A)
    pB2t = _pat(sp.csc_matrix(B2).T)
    Sx, Su = [], []
    cache = {}

    def power(p):
        if p not in cache:
            cache[p] = _pat_power(pA, p)
        return cache[p]

    for t in range(1, T + 1):
        k = int(np.floor(alpha * (t - 1)))
        X = power(min(d, k)).copy()
        U = (pB2t @ power(min(d + 1, k))).tocsc()
        for M, out in ((X, Sx), (U, Su)):
            M = sp.csc_matrix(M)
            M.eliminate_zeros()
            M.sort_indices()
            out.append(sp.csc_matrix((np.ones(M.nnz, dtype=bool), M.indices, M.indptr), shape=M.shape))
    return Sx, Su


class Plant:
    """Minimal stand-in for GeneralizedPlant{Float64,StateFeedback}.  3-argument form (GeneralizedPlant.jl:105-110):
    C1 = [I;0], D12 = [0;I], D11 = 0, C2 = I.  6-argument form (GeneralizedPlant.jl:101-103): explicit C1 (Nz x Nx),
    D11 (Nz x Nw, 0 -> zeros), D12 (Nz x Nu)."""

    def __init__(self, A, B1, B2, C1=None, D11=None, D12=None):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.B1 = sp.csc_matrix(B1, dtype=np.float64)
        self.B2 = sp.csc_matrix(B2, dtype=np.float64)
        self.Nx = self.A.shape[0]
        self.Nw = self.B1.shape[1]
        self.Nu = self.B2.shape[1]
        if C1 is None:
            CD = sp.identity(self.Nx + self.Nu, format="csc", dtype=np.float64)
            C1, D12 = CD[:, :self.Nx], CD[:, self.Nx:]
        self.C1 = sp.csc_matrix(C1, dtype=np.float64)
        self.D12 = sp.csc_matrix(D12, dtype=np.float64)
        self.Nz = self.C1.shape[0]
        self.D11 = sp.csc_matrix((self.Nz, self.Nw)) if (D11 is None or np.isscalar(D11)) else sp.csc_matrix(D11, dtype=np.float64)
        for M in (self.A, self.B1, self.B2, self.C1, self.D11, self.D12):
            M.sort_indices()
        self.Ny = self.Nx


# --------------------------------------------------------------------------------------
# gather: sparsity_dim_reduction (reduction.jl:11-27)
# --------------------------------------------------------------------------------------
def _unique_first(v):
    _, first = np.unique(v, return_index=True)
    return v[np.sort(first)]


def support_products(P: Plant, S):
    """`𝓢ⱼ[end]*(P.A.≠0)` for 𝓢ⱼ in (𝓢ₓ, 𝓢ᵤ) (reduction.jl:14).  The reference recomputes
    this global product for every column; it does not depend on the column, so it can be
    hoisted by callers looping over columns."""
    pA = _pat(P.A)
    out = []
    for Sj in S:
        M = (sp.csc_matrix(Sj[-1]).astype(np.int64) @ pA).tocsc()
        M.sort_indices()
        out.append(M)
    return out


def sparsity_dim_reduction(P: Plant, cj, S, _prod=None):
    """reduction.jl:11-27 (StateFeedback branch).  `cj` is a 0-based integer vector here.
    Returns (A_t, B1_t, B2_t), I_tilde (dense n x q), ii_x (bool n), s_x, s_u (0-based int64).
      s = unique(findnz(M[:,cj])[1])  — CSC order (column-major over cj, rows ascending
      inside a column), first occurrences kept (reduction.jl:14)."""
    cj = np.asarray(cj, dtype=np.int64)
    prods = _prod if _prod is not None else support_products(P, S)
    ss = []
    for M in prods:
        rows = [M.indices[M.indptr[c]:M.indptr[c + 1]] for c in cj]
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        ss.append(_unique_first(rows.astype(np.int64)))
    sx, su = ss
    # view(P, (sx, [sx; Nx+su]), (sx, cj, su))  (reduction.jl:15, GeneralizedPlant.jl:266-285)
    At = P.A[sx, :][:, sx].toarray()
    B1t = P.B1[sx, :][:, cj].toarray()
    B2t = P.B2[sx, :][:, su].toarray()
    P._last_cost_blocks = cost_blocks(P, cj, sx, su)
    iix = np.isin(sx, cj)                                          # reduction.jl:22
    n, q = sx.size, cj.size
    It = np.zeros((n, q))
    sel = np.nonzero(iix)[0]
    It[sel, np.arange(sel.size)] = 1.0                             # reduction.jl:23
    return (At, B1t, B2t), It, iix, sx, su


def sparsity_dim_reduction_of(mats, cj, S):
    """reduction.jl:16-18, the OutputFeedback branch:  s_x, s_u, s_y = (unique(findnz(𝓢ⱼ[end][:,cⱼ])[1]) for 𝓢ⱼ in 𝓢)
    — stored rows (findnz lists stored entries, whatever their value) of the listed columns of each family's last tap,
    first occurrences in CSC order; no product with the pattern of A.  `mats` = dict of the nine plant matrices
    (A, B1, B2, C1, D11, D12, C2, D21, D22; scipy sparse), cj 0-based.  Returns (sub, I_tilde, ii_x, s_x, s_u, s_y) where
    sub = view(P, (s_x, [s_x; Nx+s_u], s_y), (s_x, c_j, s_u)) materialised block by block
    (GeneralizedPlant.jl:266-285, OutputFeedback rows I[3] for C2/D21/D22) and I_tilde = [I(n)[:,ii_x] 0] (reduction.jl:22-23)."""
    cj = np.asarray(cj, dtype=np.int64)
    ss = []
    for Sj in S:
        M = sp.csc_matrix(Sj[-1])
        M.sort_indices()
        rows = [M.indices[M.indptr[c]:M.indptr[c + 1]] for c in cj]
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        ss.append(_unique_first(rows.astype(np.int64)))
    sx, su, sy = ss
    Nx = mats["A"].shape[0]
    zr = np.concatenate([sx, Nx + su])
    cut = lambda M, r, c: sp.csc_matrix(M)[r, :][:, c].toarray()
    sub = dict(A=cut(mats["A"], sx, sx), B1=cut(mats["B1"], sx, cj), B2=cut(mats["B2"], sx, su),
               C1=cut(mats["C1"], zr, sx), D11=cut(mats["D11"], zr, cj), D12=cut(mats["D12"], zr, su),
               C2=cut(mats["C2"], sy, sx), D21=cut(mats["D21"], sy, cj), D22=cut(mats["D22"], sy, su))
    iix = np.isin(sx, cj)
    n, q = sx.size, cj.size
    It = np.zeros((n, q))
    sel = np.nonzero(iix)[0]
    It[sel, np.arange(sel.size)] = 1.0
    return sub, It, iix, sx, su, sy


def cost_blocks(P: Plant, cj, sx, su):
    """[C1~ D12~] and D11~ of the reduced plant: rows [s_x; Nx+s_u] of C1, D12, D11 (reduction.jl:15,
    GeneralizedPlant.jl:266-271), columns s_x / s_u / c_j.  Returns (L, D11t) dense."""
    rows = np.concatenate([np.asarray(sx, np.int64), P.Nx + np.asarray(su, np.int64)])
    if rows.size and rows.max() >= P.Nz:
        raise ValueError("the reduced plant selects rows [s_x; Nx+s_u] of C1/D12/D11: needs Nz >= Nx+Nu rows")
    C1t = P.C1[rows, :][:, sx].toarray()
    D12t = P.D12[rows, :][:, su].toarray()
    D11t = P.D11[rows, :][:, np.asarray(cj, np.int64)].toarray()
    return np.hstack([C1t, D12t]), D11t


# --------------------------------------------------------------------------------------
# per-column problem (synthesis.jl:46-62) and its exact minimiser
# --------------------------------------------------------------------------------------
def column_masks(S, sx, su, j):
    """Free-variable masks 𝓢ₓ[t][sₓ,cⱼ] == 1, 𝓢ᵤ[t][sᵤ,cⱼ] == 1 (synthesis.jl:58-59), q = 1."""
    Sx, Su = S
    T = len(Sx)
    Mx = np.zeros((T, sx.size), dtype=bool)
    Mu = np.zeros((T, su.size), dtype=bool)
    for t in range(T):
        for Sm, s, M in ((Sx[t], sx, Mx), (Su[t], su, Mu)):
            rows = Sm.indices[Sm.indptr[j]:Sm.indptr[j + 1]]
            vals = Sm.data[Sm.indptr[j]:Sm.indptr[j + 1]]
            rows = rows[vals != 0]
            M[t] = np.isin(s, rows)
    return Mx, Mu


def column_problem(At, B2t, It, Mx, Mu):
    """Stack the equality constraints of synthesis.jl:53-55 over the free (in-mask) entries
    only (off-mask entries are fixed to 0.0, synthesis.jl:57-60).  q = 1.
    Unknown vector v = [x_1[free]; u_1[free]; x_2[free]; u_2[free]; ...].
    Returns G (rows = (T+1)*n), h, and (xoff, uoff) index maps (-1 = fixed)."""
    T, n = Mx.shape
    m = Mu.shape[1]
    xoff = -np.ones((T, n), dtype=np.int64)
    uoff = -np.ones((T, m), dtype=np.int64)
    nv = 0
    for t in range(T):
        k = int(Mx[t].sum()); xoff[t, Mx[t]] = nv + np.arange(k); nv += k
        k = int(Mu[t].sum()); uoff[t, Mu[t]] = nv + np.arange(k); nv += k
    G = np.zeros(((T + 1) * n, nv))
    h = np.zeros((T + 1) * n)
    # X[1] == I~                                   (synthesis.jl:53)
    for i in range(n):
        if xoff[0, i] >= 0:
            G[i, xoff[0, i]] = 1.0
    h[:n] = It[:, 0]
    # X[t+1] == A~ X[t] + B2~ U[t], t = 1..T-1     (synthesis.jl:54)
    # 0      == A~ X[T] + B2~ U[T]                 (synthesis.jl:55)
    for t in range(T):
        r0 = (t + 1) * n
        fx = np.nonzero(Mx[t])[0]
        fu = np.nonzero(Mu[t])[0]
        G[r0:r0 + n, xoff[t, fx]] -= At[:, fx]
        G[r0:r0 + n, uoff[t, fu]] -= B2t[:, fu]
        if t + 1 < T:
            for i in np.nonzero(Mx[t + 1])[0]:
                G[r0 + i, xoff[t + 1, i]] += 1.0
    return G, h, (xoff, uoff)


def solve_column(At, B2t, It, Mx, Mu, rcond=1e-11):
    """Exact minimiser of  sum_t ||X[t]||^2 + ||U[t]||^2  s.t. G v = h  (synthesis.jl:50-55 with
    [C1~ D12~] = I, B1~[ii_x,:] = 1, D11~ = 0 — the 3-arg Plant, GeneralizedPlant.jl:105-110).
    G is rank-deficient but consistent => SVD-based minimum-norm solution.
    Returns X (T x n), U (T x m), info dict(residual, rank, rows, cost)."""
    T, n = Mx.shape
    m = Mu.shape[1]
    G, h, (xoff, uoff) = column_problem(At, B2t, It, Mx, Mu)
    v, _, rank, sv = np.linalg.lstsq(G, h, rcond=rcond)
    res = float(np.linalg.norm(G @ v - h)) if G.shape[1] else float(np.linalg.norm(h))
    X = np.zeros((T, n)); U = np.zeros((T, m))
    X[xoff >= 0] = v[xoff[xoff >= 0]]
    U[uoff >= 0] = v[uoff[uoff >= 0]]
    nz_rows = int((np.abs(G).sum(axis=1) > 0).sum())
    return X, U, dict(residual=res, rank=int(rank), rows=nz_rows, cost=float(v @ v),
                      smin=float(sv[rank - 1]) if rank else 0.0, smax=float(sv[0]) if sv.size else 0.0)


def solve_group_general(At, B2t, It, Mxs, Mus, L, R, D, rcond=1e-11):
    """Exact minimiser of the reference's general per-group problem (synthesis.jl:46-62, 76-77):
        min  sum_t || L [X_t; U_t] R + D ||_F^2     X_t (n x q), U_t (m x q), off-mask entries fixed to 0
        s.t. X_1 = I~,  X_{t+1} = A~ X_t + B2~ U_t,  0 = A~ X_T + B2~ U_T          (column by column)
    with L = [C1~ D12~], R = B1~[ii_x,:] (q x q), D = D11~.  Mxs/Mus: per-column masks (lists of T x n / T x m).
    Null-space method on the stacked system; minimum-norm among minimisers when the cost is not strictly convex.
    Returns X (T x n x q), U (T x m x q), info."""
    q = len(Mxs)
    T, n = Mxs[0].shape
    m = Mus[0].shape[1]
    Gs, hs, offs, nvs = [], [], [], []
    for c in range(q):
        G, h, (xoff, uoff) = column_problem(At, B2t, It[:, [c]], Mxs[c], Mus[c])
        Gs.append(G); hs.append(h); offs.append((xoff, uoff)); nvs.append(G.shape[1])
    start = np.concatenate([[0], np.cumsum(nvs)])
    nv = int(start[-1])
    G = np.zeros((sum(g.shape[0] for g in Gs), nv)); h = np.concatenate(hs)
    r = 0
    for c in range(q):
        G[r:r + Gs[c].shape[0], start[c]:start[c + 1]] = Gs[c]; r += Gs[c].shape[0]
    # cost rows: for every tap t, vec(L V_t R + D)  =  sum_c' (R[c',c] * L[:, var]) ...
    nzr = L.shape[0]
    F = np.zeros((T * nzr * q, nv)); f0 = np.zeros(T * nzr * q)
    for t in range(T):
        for cout in range(q):                                    # column of the (Nz x q) block L V R + D
            rows = slice((t * q + cout) * nzr, (t * q + cout + 1) * nzr)
            f0[rows] = D[:, cout] if D.size else 0.0
            for cin in range(q):
                rr = R[cin, cout]
                if rr == 0.0:
                    continue
                xoff, uoff = offs[cin]
                fx = np.nonzero(xoff[t] >= 0)[0]; fu = np.nonzero(uoff[t] >= 0)[0]
                F[rows, start[cin] + xoff[t, fx]] += rr * L[:, fx]
                F[rows, start[cin] + uoff[t, fu]] += rr * L[:, n + fu]
    vp, _, rank, sv = np.linalg.lstsq(G, h, rcond=rcond)
    res = float(np.linalg.norm(G @ vp - h))
    U_, s_, Vt = np.linalg.svd(G, full_matrices=True)
    rk = int((s_ > rcond * (s_[0] if s_.size else 1.0)).sum())
    N = Vt[rk:].T                                                # null space of G
    if N.shape[1]:
        z = np.linalg.lstsq(F @ N, -(F @ vp + f0), rcond=rcond)[0]
        v = vp + N @ z
    else:
        v = vp
    X = np.zeros((T, n, q)); U = np.zeros((T, m, q))
    for c in range(q):
        xoff, uoff = offs[c]
        vc = v[start[c]:start[c + 1]]
        X[:, :, c][xoff >= 0] = vc[xoff[xoff >= 0]]
        U[:, :, c][uoff >= 0] = vc[uoff[uoff >= 0]]
    return X, U, dict(residual=res, rank=rk, cost=float(np.sum((F @ v + f0) ** 2)))


def SLS_H2_general(P: Plant, S, I=None):
    """synthesis.jl:11-72 in full generality for state-feedback plants: column groups of any size (the 𝓘 keyword,
    synthesis.jl:15), general [C1 D12], B1 and D11 in the cost (synthesis.jl:42-43,50,76-77).  Dense, slow: small cases.
    Scatter ordering (the `# TODO: Verify dimensions` of synthesis.jl:64): `vec(repeat(s_x, q, 1)')` cycles through s_x
    fastest and `vec(repeat(c_j, 1, n)')` repeats each group column n times, which is exactly the column-major order of
    `vec(value.(Phi~_x[t]))` (n x q) — entry (i, c) of the local solution lands at (s_x[i], c_j[c]), as below."""
    Sx, Su = S
    T = len(Sx)
    groups = [[j] for j in range(P.Nx)] if I is None else [list(map(int, np.atleast_1d(g))) for g in I]
    prods = support_products(P, S)
    ex = [([], [], []) for _ in range(T)]
    eu = [([], [], []) for _ in range(T)]
    for cj in groups:
        (At, B1t, B2t), It, iix, sx, su = sparsity_dim_reduction(P, cj, S, _prod=prods)
        L, D = P._last_cost_blocks
        R = B1t[iix, :]                                          # synthesis.jl:42
        if R.shape[0] != len(cj):
            raise ValueError("B1~[ii_x,:] is not square: a group column is missing from its own s_x")
        Mxs, Mus = zip(*[column_masks(S, sx, su, j) for j in cj])
        X, U, _ = solve_group_general(At, B2t, It, list(Mxs), list(Mus), L, R, D)
        for c, j in enumerate(cj):
            for t in range(T):
                fx = np.nonzero(Mxs[c][t])[0]
                ex[t][0].append(sx[fx]); ex[t][1].append(np.full(fx.size, j)); ex[t][2].append(X[t, fx, c])
                fu = np.nonzero(Mus[c][t])[0]
                eu[t][0].append(su[fu]); eu[t][1].append(np.full(fu.size, j)); eu[t][2].append(U[t, fu, c])
    return _build_taps(ex, (P.Nx, P.Nx)), _build_taps(eu, (P.Nu, P.Nx))


def _build_taps(e, shape):
    out = []
    for r, c, v in e:
        r = np.concatenate(r) if r else np.zeros(0, np.int64)
        c = np.concatenate(c) if c else np.zeros(0, np.int64)
        v = np.concatenate(v) if v else np.zeros(0)
        order = np.lexsort((r, c))
        r, c, v = r[order], c[order], v[order]
        indptr = np.zeros(shape[1] + 1, dtype=np.int64)
        np.add.at(indptr, c + 1, 1)
        indptr = np.cumsum(indptr)
        out.append(sp.csc_matrix((v, r.astype(np.int64), indptr), shape=shape))
    return out


def SLS_H2(P: Plant, S, I=None, return_info=False):
    """synthesis.jl:11-32 + :34-72 for single-column groups (q = 1).  `I` is a list of 0-based
    columns (default all, synthesis.jl:15 — note: 1:P.Nx, not P.Nw).
    Returns (Phi_x, Phi_u): two T-lists of CSC matrices (Nx x Nx, Nu x Nx) whose stored pattern
    is the full mask (explicit zeros kept; the reference's `+=` at synthesis.jl:67 would drop
    exact 0.0 — compare values, not stored-zero patterns)."""
    Sx, Su = S
    T = len(Sx)
    cols = list(range(P.Nx)) if I is None else [int(c) for c in I]
    prods = support_products(P, S)
    ex = [([], [], []) for _ in range(T)]
    eu = [([], [], []) for _ in range(T)]
    infos = {}
    for j in cols:
        (At, _B1t, B2t), It, iix, sx, su = sparsity_dim_reduction(P, [j], S, _prod=prods)
        Mx, Mu = column_masks(S, sx, su, j)
        X, U, info = solve_column(At, B2t, It, Mx, Mu)
        infos[j] = dict(info, n=sx.size, m=su.size)
        for t in range(T):
            fx = np.nonzero(Mx[t])[0]
            ex[t][0].append(sx[fx]); ex[t][1].append(np.full(fx.size, j)); ex[t][2].append(X[t, fx])
            fu = np.nonzero(Mu[t])[0]
            eu[t][0].append(su[fu]); eu[t][1].append(np.full(fu.size, j)); eu[t][2].append(U[t, fu])

    def build(e, shape):
        out = []
        for r, c, v in e:
            r = np.concatenate(r) if r else np.zeros(0, np.int64)
            c = np.concatenate(c) if c else np.zeros(0, np.int64)
            v = np.concatenate(v) if v else np.zeros(0)
            # build CSC by hand so that explicit zeros are kept and rows stay ascending
            order = np.lexsort((r, c))
            r, c, v = r[order], c[order], v[order]
            indptr = np.zeros(shape[1] + 1, dtype=np.int64)
            np.add.at(indptr, c + 1, 1)
            indptr = np.cumsum(indptr)
            out.append(sp.csc_matrix((v, r.astype(np.int64), indptr), shape=shape))
        return out

    Phix = build(ex, (P.Nx, P.Nx))
    Phiu = build(eu, (P.Nu, P.Nx))
    if return_info:
        return Phix, Phiu, infos
    return Phix, Phiu


# --------------------------------------------------------------------------------------
# invariants and the README closed-loop simulation
# --------------------------------------------------------------------------------------
def achievability_residual(P: Plant, Phix, Phiu, cols=None):
    """max-abs residual of Φx[1]=I, Φx[t+1]=AΦx[t]+B2Φu[t], AΦx[T]+B2Φu[T]=0 on the assembled
    global matrices, restricted to the synthesised columns."""
    T = len(Phix)
    cols = np.arange(P.Nx) if cols is None else np.asarray(cols)
    I = sp.identity(P.Nx, format="csc")
    r = abs((Phix[0] - I)[:, cols]).max() if cols.size else 0.0
    for t in range(T):
        nxt = Phix[t + 1] if t + 1 < T else sp.csc_matrix((P.Nx, P.Nx))
        R = (nxt - P.A @ Phix[t] - P.B2 @ Phiu[t])[:, cols]
        r = max(r, abs(R).max() if R.nnz else 0.0)
    return float(r)


def simulate_closed_loop(P: Plant, Phix, Phiu, w, steps: int):
    """README.md:62-71.  `w` is an (Nw x steps) array (column t-1 = w(t), 1-based t).
    Returns x (Nx x steps), u (Nu x steps), beta (Nx x steps)."""
    T = len(Phix)
    x = np.zeros((P.Nx, steps)); beta = np.zeros((P.Nx, steps)); u = np.zeros((P.Nu, steps))
    what = lambda k: x[:, k] - beta[:, k]
    for t in range(1, steps):                      # t is the README's 1-based t
        b = np.zeros(P.Nx)
        for tau in range(1, min(t, T - 1) + 1):
            b += Phix[tau] @ what(t - tau)        # Φx[τ+1] * ŵ(t+1-τ)  (0-based: Phix[tau], col t-tau)
        beta[:, t] = b
        uu = np.zeros(P.Nu)
        for tau in range(1, min(t, T) + 1):
            uu += Phiu[tau - 1] @ what(t - tau)
        u[:, t - 1] = uu
        x[:, t] = P.A @ x[:, t - 1] + P.B1 @ w[:, t - 1] + P.B2 @ u[:, t - 1]
    return x, u, beta
