"""Host-side mirror of the reference's interface for the column-wise localized H2 synthesis path.

    Plant(A, B1, B2)                      <- src/types/GeneralizedPlant.jl:105-110,190 (3-argument ctor)
    SLS_H2(P, [Sx, Su]; I=None)           <- src/synthesis.jl:11-32  (`SLS_𝓗₂`)
    sparsity_dim_reduction(P, cj, S)      <- src/reduction.jl:11-27

Same names, argument meaning and error behaviour as the Julia functions; sparse matrices are
scipy CSC on the Python side and are handed to libslsb200 in Julia's `SparseMatrixCSC{Tv,Int64}`
field layout (1-based int64 colptr/rowval), which is what the Julia shim passes zero-copy.
All arithmetic happens in the CUDA library (csrc/); nothing here computes Φ on the CPU.
Python convention: column indices are 0-based (Julia's are 1-based)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _capi
from ._capi import SlsOptions, SlsTiming, f64p, i32p, i64p, ptr, ptr_array, u8p

__all__ = ["Plant", "JuliaCSC", "SLS_H2", "sparsity_dim_reduction", "Synthesizer"]


class JuliaCSC:
    """Field-for-field image of a Julia `SparseMatrixCSC{Tv,Int64}`: m, n, colptr, rowval, nzval."""

    __slots__ = ("m", "n", "colptr", "rowval", "nzval")

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n = int(m), int(n)
        self.colptr, self.rowval, self.nzval = colptr, rowval, nzval

    @classmethod
    def from_scipy(cls, M, dtype):
        M = sp.csc_matrix(M)
        if not M.has_sorted_indices:
            M = M.copy()
            M.sort_indices()
        colptr = M.indptr.astype(np.int64) + 1
        rowval = M.indices.astype(np.int64) + 1
        if dtype == np.bool_:
            nzval = (M.data != 0).astype(np.uint8)
        else:
            nzval = np.ascontiguousarray(M.data, dtype=dtype)
        return cls(M.shape[0], M.shape[1], colptr, rowval, nzval)

    def to_scipy(self):
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))

    @property
    def nnz(self):
        return int(self.colptr[-1] - 1)


def _as_julia(M, dtype):
    return M if isinstance(M, JuliaCSC) else JuliaCSC.from_scipy(M, dtype)


class Plant:
    """`Plant(A, B₁, B₂)`: state-feedback generalized plant with the LQR-style defaults of the
    3-argument constructor — C₁ = [I;0], D₁₂ = [0;I], D₁₁ = 0, C₂ = I (GeneralizedPlant.jl:105-110) — or the
    6-argument `Plant(A, B₁, B₂, C₁, D₁₁, D₁₂)` (GeneralizedPlant.jl:101-103; D₁₁ = 0 gives zeros of the right size).
    Dimension checks follow validate_GeneralizedPlant (GeneralizedPlant.jl:291-311): a mismatch raises
    ValueError (the reference throws ErrorException)."""

    def __init__(self, A, B1, B2, C1=None, D11=None, D12=None, C2=None, D21=None, D22=None):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.B1 = sp.csc_matrix(B1, dtype=np.float64)
        self.B2 = sp.csc_matrix(B2, dtype=np.float64)
        if self.A.shape[0] != self.A.shape[1]:
            raise ValueError("A must be a square matrix")
        if self.B1.shape[0] != self.A.shape[0] or self.B2.shape[0] != self.A.shape[0]:
            raise ValueError("B1 and B2 must have the same number of rows as A")
        self.Nx = self.A.shape[0]
        self.Nw = self.B1.shape[1]
        self.Nu = self.B2.shape[1]
        if (C1 is None) != (D12 is None):
            raise ValueError("C1 and D12 must be given together")
        if C1 is None:
            CD = sp.identity(self.Nx + self.Nu, format="csc", dtype=np.float64)
            C1, D12 = CD[:, :self.Nx], CD[:, self.Nx:]
        self.C1 = sp.csc_matrix(C1, dtype=np.float64)
        self.D12 = sp.csc_matrix(D12, dtype=np.float64)
        self.Nz = self.C1.shape[0]
        if self.C1.shape[1] != self.Nx or self.D12.shape != (self.Nz, self.Nu):
            raise ValueError("C1 must be Nz x Nx and D12 must be Nz x Nu")
        if D11 is None or (np.isscalar(D11) and D11 == 0):
            self.D11 = sp.csc_matrix((self.Nz, self.Nw), dtype=np.float64)
        else:
            self.D11 = sp.csc_matrix(D11, dtype=np.float64)
            if self.D11.shape != (self.Nz, self.Nw):
                raise ValueError("D11 must be Nz x Nw")
        for M in (self.A, self.B1, self.B2, self.C1, self.D11, self.D12):
            M.sort_indices()
        # 9-argument form (GeneralizedPlant.jl:69-99): C₂ == I with an empty / zero D₂₁ is state feedback, anything else is
        # an OutputFeedback plant — which SLS_H2 does not synthesise (synthesis.jl:13 falls through) but
        # sparsity_dim_reduction reduces (reduction.jl:16-18)
        def _isempty(M):                                          # Julia isempty: no elements (not "no stored entries")
            return M is not None and not np.isscalar(M) and int(np.prod(np.shape(M) if not sp.issparse(M) else M.shape)) == 0
        d21_zero = D21 is None or (np.isscalar(D21) and D21 == 0) or _isempty(D21)     # isempty(D₂₁) || D₂₁ == 0
        c2_eye = C2 is None or (sp.issparse(C2) and C2.shape == (self.Nx, self.Nx)
                                and abs(sp.csc_matrix(C2) - sp.identity(self.Nx, format="csc")).nnz == 0)
        if c2_eye and d21_zero:
            self.Ny = self.Nx
            self.feedback = "StateFeedback"
        else:
            self.C2 = sp.identity(self.Nx, format="csc", dtype=np.float64) if C2 is None else sp.csc_matrix(C2, dtype=np.float64)
            self.Ny = self.C2.shape[0]
            self.D21 = sp.csc_matrix((self.Ny, self.Nw), dtype=np.float64) if d21_zero else sp.csc_matrix(D21, dtype=np.float64)
            self.D22 = sp.csc_matrix((self.Ny, self.Nu), dtype=np.float64) if (D22 is None or (np.isscalar(D22) and D22 == 0) or _isempty(D22)) \
                else sp.csc_matrix(D22, dtype=np.float64)
            if self.C2.shape[1] != self.Nx:
                raise ValueError("C2 must have Nx columns")
            if self.D21.shape != (self.Ny, self.Nw) or self.D22.shape != (self.Ny, self.Nu):
                raise ValueError("D21 must be Ny x Nw and D22 must be Ny x Nu")
            for M in (self.C2, self.D21, self.D22):
                M.sort_indices()
            self.feedback = "OutputFeedback"

    def diagonal_cost(self):
        """(wx, wu, why): the diagonal weights when [C₁ D₁₂] = blockdiag(diag(wx), diag(wu)) with positive entries,
        D₁₁ = 0 and B₁ is diagonal with a nonzero diagonal — the cost family libslsb200 solves (sls_set_weights);
        wx = wu = None for the identity.  Otherwise (None, None, reason)."""
        Nx, Nu = self.Nx, self.Nu
        if self.B1.shape != (Nx, Nx):
            return None, None, "B1 is not square (Nw != Nx)"
        B1 = self.B1.tocoo()
        if (B1.row != B1.col).any() and abs(B1.data[B1.row != B1.col]).max() > 0:
            return None, None, "B1 is not diagonal: R = B1~[ii_x,:] couples the columns of a group in the cost"
        if (self.B1.diagonal() == 0).any():
            return None, None, "B1 has a zero on its diagonal: the cost of that column vanishes"
        if self.D11.nnz and abs(self.D11.data).max() > 0:
            return None, None, "D11 != 0 adds a linear term to the cost"
        if self.Nz != Nx + Nu:
            return None, None, "Nz != Nx + Nu"
        wx = self.C1[:Nx, :].diagonal(); wu = self.D12[Nx:, :].diagonal()
        Cd = sp.vstack([sp.diags(wx), sp.csc_matrix((Nu, Nx))]).tocsc()
        Dd = sp.vstack([sp.csc_matrix((Nx, Nu)), sp.diags(wu)]).tocsc()
        if abs(self.C1 - Cd).nnz and abs(self.C1 - Cd).max() > 0 or abs(self.D12 - Dd).nnz and abs(self.D12 - Dd).max() > 0:
            return None, None, "[C1 D12] is not diagonal"
        if (wx <= 0).any() or (wu <= 0).any():
            return None, None, "a diagonal cost weight is not positive"
        if (wx == 1).all() and (wu == 1).all():
            return None, None, None
        return np.ascontiguousarray(wx, dtype=np.float64), np.ascontiguousarray(wu, dtype=np.float64), None

    def _b1_is_identity(self):
        if self.B1.shape != (self.Nx, self.Nx):
            return False
        return (abs(self.B1 - sp.identity(self.Nx, format="csc")).nnz == 0)


def _groups_decouple(P, Sx, Su, groups):
    """A column group cⱼ with q > 1 shares sₓ = the union of its columns' index sets (reduction.jl:14), so every column
    gets the union's constraint rows.  With a diagonal B₁ the cost is separable, and the group problem equals the
    single-column problems exactly when the extra rows are empty for every member column j: everything A and B₂ reach
    from column j's own supports already lies inside j's own sₓ.  Checked here on the patterns (host, integer work)."""
    pA = (P.A != 0).astype(np.int64).tocsc()
    pB = (P.B2 != 0).astype(np.int64).tocsc()
    own = ((sp.csc_matrix(Sx[-1]) != 0).astype(np.int64) @ pA).tocsc()       # column j: its own s_x
    reach = None
    for M, pat in ((Sx, pA), (Su, pB)):
        for St in M:
            Rm = (pat @ (sp.csc_matrix(St) != 0).astype(np.int64)).tocsc()
            reach = Rm if reach is None else reach + Rm
    reach = reach.tocsc()
    for g in groups:
        if len(g) < 2:
            continue
        for j in g:
            rows = reach.indices[reach.indptr[j]:reach.indptr[j + 1]]
            if not np.isin(rows, own.indices[own.indptr[j]:own.indptr[j + 1]]).all():
                return False
    return True


class Synthesizer:
    """One libslsb200 handle (one GPU).  Thin, stateful wrapper used by SLS_H2, the tests and bench.py."""

    def __init__(self, device: int = 0, rank_tol: float = 0.0, feas_tol: float = 0.0, flags: int = 0, api=None):
        self.api = api if api is not None else _capi.default_api()
        self.h = C.c_void_p()
        opts = SlsOptions(device, flags, rank_tol, feas_tol)
        rc = self.api.sls_create(C.byref(opts), C.byref(self.h))
        if rc != 0:
            msg = self.api.sls_last_error(None)
            raise _capi.SlsError(rc, msg.decode() if msg else "")
        self.Nx = self.Nu = self.T = 0
        self.ncols = 0
        self._keep = []

    def close(self):
        if self.h:
            self.api.sls_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        self.api.check(self.h, rc)

    # -- inputs ---------------------------------------------------------------------------------
    def set_weights(self, wx, wu, Nx, Nu):
        """Diagonal cost weights (sls_set_weights); call before set_plant.  (None, None) clears them."""
        wx = None if wx is None else np.ascontiguousarray(wx, dtype=np.float64)
        wu = None if wu is None else np.ascontiguousarray(wu, dtype=np.float64)
        self._ck(self.api.sls_set_weights(self.h, Nx, Nu, ptr(wx, f64p), ptr(wu, f64p)))

    def set_plant(self, A, B2):
        A = _as_julia(A, np.float64)
        B2 = _as_julia(B2, np.float64)
        self.Nx, self.Nu = A.n, B2.n
        self._ck(self.api.sls_set_plant(self.h, A.n, B2.n, ptr(A.colptr, i64p), ptr(A.rowval, i64p), ptr(A.nzval, f64p),
                                        ptr(B2.colptr, i64p), ptr(B2.rowval, i64p), ptr(B2.nzval, f64p)))

    def set_cost(self, C1, D12, D11=None, B1=None):
        """General cost matrices (sls_set_cost); call after set_plant.  D11 None = 0, B1 None = I; all None clears."""
        if C1 is None and D12 is None and D11 is None and B1 is None:
            n = ptr(None, i64p); z = ptr(None, f64p)
            self._ck(self.api.sls_set_cost(self.h, 0, 0, n, n, z, n, n, z, n, n, z, n, n, z))
            return
        mats = [None if M is None else _as_julia(M, np.float64) for M in (C1, D12, D11, B1)]
        Nz = mats[0].m if mats[0] is not None else (mats[2].m if mats[2] is not None else 0)
        Nw = mats[2].n if mats[2] is not None else (mats[3].n if mats[3] is not None else self.Nx)
        args = []
        for M in mats:
            args += [ptr(None if M is None else M.colptr, i64p), ptr(None if M is None else M.rowval, i64p),
                     ptr(None if M is None else M.nzval, f64p)]
        self._keep_cost = mats
        self._ck(self.api.sls_set_cost(self.h, Nz, Nw, *args))

    def set_groups(self, grp_ptr):
        """Column groups over the planned column list (sls_set_groups): 0-based offsets, None clears."""
        if grp_ptr is None:
            self._ck(self.api.sls_set_groups(self.h, 0, ptr(None, i64p)))
            return
        g = np.ascontiguousarray(grp_ptr, dtype=np.int64)
        self._ck(self.api.sls_set_groups(self.h, g.size - 1, ptr(g, i64p)))

    def set_supports(self, Sx, Su, pass_nzval=True):
        Sx = [_as_julia(M, np.bool_) for M in Sx]
        Su = [_as_julia(M, np.bool_) for M in Su]
        if len(Sx) != len(Su):
            raise ValueError("Sx and Su must have the same number of taps")
        self.T = len(Sx)
        args = []
        for S in (Sx, Su):
            args.append(ptr_array([M.colptr for M in S], i64p))
            args.append(ptr_array([M.rowval for M in S], i64p))
            args.append(ptr_array([M.nzval for M in S], u8p) if pass_nzval else None)
        self._ck(self.api.sls_set_supports(self.h, self.T, *args))

    def make_supports(self, d, T, alpha):
        """README.md:52-54 on the device (no host-side supports, no PCIe transfer).  Returns (nnz_x[T], nnz_u[T])."""
        nnz_x = np.zeros(T, np.int64)
        nnz_u = np.zeros(T, np.int64)
        self._ck(self.api.sls_make_supports(self.h, d, T, float(alpha), ptr(nnz_x, i64p), ptr(nnz_u, i64p)))
        self.T = T
        self._supp_nnz = (nnz_x, nnz_u)
        return nnz_x, nnz_u

    def get_supports(self):
        """Device-generated supports back on the host: (Sx, Su) lists of JuliaCSC (Bool, all true)."""
        out = []
        for which, nrow in ((0, self.Nx), (1, self.Nu)):
            taps = []
            for t in range(self.T):
                nnz = int(self._supp_nnz[which][t])
                cp = np.zeros(self.Nx + 1, np.int64); rv = np.zeros(max(nnz, 1), np.int64); nz = np.zeros(max(nnz, 1), np.uint8)
                self._ck(self.api.sls_get_supports(self.h, which, t, ptr(cp, i64p), ptr(rv, i64p), ptr(nz, u8p)))
                taps.append(JuliaCSC(nrow, self.Nx, cp, rv[:nnz], nz[:nnz]))
            out.append(taps)
        return out

    # -- kernel 1 -------------------------------------------------------------------------------
    def plan(self, cols=None, sizes=True):
        """cols: 0-based disturbance columns (None = all).  Returns (nnz_x[T], nnz_u[T], n[ncols], m[ncols]).
        sizes=False: n / m are not fetched (NULL n_out / m_out: the per-column sizes then stay on the device)."""
        T = self.T
        nnz_x = np.zeros(T, np.int64)
        nnz_u = np.zeros(T, np.int64)
        if cols is None:
            ncols = self.Nx
            c1 = None
        else:
            c1 = np.ascontiguousarray(np.asarray(cols, dtype=np.int64) + 1)
            ncols = c1.size
        n = np.zeros(max(ncols, 1) if sizes else 1, np.int32)
        m = np.zeros(max(ncols, 1) if sizes else 1, np.int32)
        self._ck(self.api.sls_plan(self.h, ptr(c1, i64p), ncols, ptr(nnz_x, i64p), ptr(nnz_u, i64p),
                                   ptr(n, i32p) if sizes else None, ptr(m, i32p) if sizes else None))
        self.ncols = ncols
        self.nnz_x, self.nnz_u = nnz_x, nnz_u
        self.n, self.m = (n[:ncols], m[:ncols]) if sizes else (None, None)
        return nnz_x, nnz_u, self.n, self.m

    def reduce(self, cj):
        """sparsity_dim_reduction index sets for the 0-based column group cj: (sx, su, iix), 0-based."""
        cj1 = np.ascontiguousarray(np.asarray(cj, dtype=np.int64) + 1)
        sx = np.zeros(self.Nx, np.int64)
        su = np.zeros(max(self.Nu, 1), np.int64)
        iix = np.zeros(self.Nx, np.uint8)
        nsx = C.c_int64()
        nsu = C.c_int64()
        self._ck(self.api.sls_reduce(self.h, ptr(cj1, i64p), cj1.size, ptr(sx, i64p), sx.size, C.byref(nsx),
                                     ptr(su, i64p), su.size, C.byref(nsu), ptr(iix, u8p)))
        return sx[:nsx.value] - 1, su[:nsu.value] - 1, iix[:nsx.value].astype(bool)

    def reduce_of(self, last_taps, cj):
        """Output-feedback branch of sparsity_dim_reduction (reduction.jl:16-18) through sls_reduce_of: `last_taps` = the last
        tap 𝓢ⱼ[end] of every support family; returns ([s_x, s_u, s_y, ...], ii_x), 0-based."""
        cj1 = np.ascontiguousarray(np.asarray(cj, dtype=np.int64) + 1)
        fams = [_as_julia(M, np.bool_) for M in last_taps]
        ncols = fams[0].n
        if any(F.n != ncols for F in fams):
            raise ValueError("the support families must have the same number of columns")
        nrows = np.array([F.m for F in fams], dtype=np.int64)
        outs = [np.zeros(max(F.m, 1), np.int64) for F in fams]
        caps = np.array([o.size for o in outs], dtype=np.int64)
        ns = np.zeros(len(fams), np.int64)
        iix = np.zeros(max(fams[0].m, 1), np.uint8)
        self._ck(self.api.sls_reduce_of(self.h, len(fams), ptr_array([F.colptr for F in fams], i64p),
                                        ptr_array([F.rowval for F in fams], i64p), ptr(nrows, i64p), ncols,
                                        ptr(cj1, i64p), cj1.size, ptr_array(outs, i64p), ptr(caps, i64p), ptr(ns, i64p),
                                        ptr(iix, u8p)))
        return [o[:int(k)] - 1 for o, k in zip(outs, ns)], iix[:int(ns[0])].astype(bool)

    def local_problem(self, k):
        n, m, T = int(self.n[k]), int(self.m[k]), self.T
        At = np.zeros((n, n), order="F")
        B2t = np.zeros((n, m), order="F")
        Mx = np.zeros((T, n), np.uint8)
        Mu = np.zeros((T, m), np.uint8)
        self._ck(self.api.sls_local_problem(self.h, k, ptr(At, f64p), ptr(B2t, f64p), ptr(Mx, u8p), ptr(Mu, u8p)))
        return At, B2t, Mx.astype(bool), Mu.astype(bool)

    # -- kernels 2 + 3 --------------------------------------------------------------------------
    def alloc_outputs(self, pinned=False):
        """Caller-owned output buffers in Julia layout, sized by the last plan()."""
        T, Nx = self.T, self.Nx

        def buf(n, dt):
            if pinned:
                import torch
                t = torch.empty(max(int(n), 1), dtype={np.int64: torch.int64, np.float64: torch.float64}[dt]).pin_memory()
                self._keep.append(t)
                return t.numpy()[:int(n)]
            return np.empty(int(n), dt)

        out = {}
        for name, nnz in (("x", self.nnz_x), ("u", self.nnz_u)):
            out[name] = dict(colptr=[buf(Nx + 1, np.int64) for _ in range(T)],
                             rowval=[buf(nnz[t], np.int64) for t in range(T)],
                             nzval=[buf(nnz[t], np.float64) for t in range(T)])
        out["status"] = np.zeros(max(self.ncols, 1), np.int32)
        out["resid"] = np.zeros(max(self.ncols, 1), np.float64)
        return out

    def solve(self, out=None):
        out = out if out is not None else self.alloc_outputs()
        args = []
        for name in ("x", "u"):
            args.append(ptr_array(out[name]["colptr"], i64p))
            args.append(ptr_array(out[name]["rowval"], i64p))
            args.append(ptr_array(out[name]["nzval"], f64p))
        self._ck(self.api.sls_solve(self.h, *args, ptr(out["status"], i32p), ptr(out["resid"], f64p)))
        return out

    def solve_status(self):
        """sls_solve with every CSC output left out (NULL): runs kernels 2+3 and returns (status, resid) only."""
        status = np.zeros(max(self.ncols, 1), np.int32)
        resid = np.zeros(max(self.ncols, 1), np.float64)
        self._ck(self.api.sls_solve(self.h, None, None, None, None, None, None, ptr(status, i32p), ptr(resid, f64p)))
        return status[:self.ncols], resid[:self.ncols]

    def synthesize(self, Sx, Su, pass_nzval=True, out=None, pinned=False, col_range=None):
        """One-shot pipelined `set_supports + plan(all columns) + solve` (sls_synthesize): H2D of the next column
        chunk, kernels 1-3 and D2H of the previous chunk overlap.  Output taps are allocated at the support's size
        (the stored pattern of Phi[t] is a subset of S[t]'s) and trimmed to the returned counts.
        col_range = (c0, c1): 0-based half-open contiguous column range (one GPU's shard); None = all columns."""
        Sx = [_as_julia(M, np.bool_) for M in Sx]
        Su = [_as_julia(M, np.bool_) for M in Su]
        if len(Sx) != len(Su):
            raise ValueError("Sx and Su must have the same number of taps")
        self.T = T = len(Sx)
        if out is None:
            self.nnz_x = np.array([M.nnz for M in Sx], np.int64)
            self.nnz_u = np.array([M.nnz for M in Su], np.int64)
            self.ncols = self.Nx
            out = self.alloc_outputs(pinned=pinned)
        c0, c1 = (0, self.Nx) if col_range is None else (int(col_range[0]), int(col_range[1]))
        cap_x = np.array([a.size for a in out["x"]["rowval"]], np.int64)
        cap_u = np.array([a.size for a in out["u"]["rowval"]], np.int64)
        args = []
        for S in (Sx, Su):
            args.append(ptr_array([M.colptr for M in S], i64p))
            args.append(ptr_array([M.rowval for M in S], i64p))
            args.append(ptr_array([M.nzval for M in S], u8p) if pass_nzval else None)
        oargs = []
        for name in ("x", "u"):
            oargs.append(ptr_array(out[name]["colptr"], i64p))
            oargs.append(ptr_array(out[name]["rowval"], i64p))
            oargs.append(ptr_array(out[name]["nzval"], f64p))
        nnz_x = np.zeros(T, np.int64); nnz_u = np.zeros(T, np.int64)
        self._ck(self.api.sls_synthesize(self.h, T, *args, c0 + 1, c1 - c0, ptr(cap_x, i64p), ptr(cap_u, i64p), *oargs,
                                         ptr(nnz_x, i64p), ptr(nnz_u, i64p), ptr(out["status"], i32p), ptr(out["resid"], f64p)))
        self.nnz_x, self.nnz_u, self.ncols = nnz_x, nnz_u, c1 - c0
        for name, nnz in (("x", nnz_x), ("u", nnz_u)):
            out[name]["rowval"] = [a[:int(nnz[t])] for t, a in enumerate(out[name]["rowval"])]
            out[name]["nzval"] = [a[:int(nnz[t])] for t, a in enumerate(out[name]["nzval"])]
        return out

    def synthesize_values_only(self, Sx, Su, out, pass_nzval=True, col_range=None):
        """sls_synthesize with NULL rowval outputs (colptr and nzval only).  Returns (nnz_x, nnz_u): where
        nnz_x[t] == nnz(Sx[t][:, range]) the pattern of Phi_x[t] is exactly Sx[t]'s and its rowval can be aliased."""
        Sx = [_as_julia(M, np.bool_) for M in Sx]
        Su = [_as_julia(M, np.bool_) for M in Su]
        T = len(Sx)
        c0, c1 = (0, self.Nx) if col_range is None else (int(col_range[0]), int(col_range[1]))
        cap_x = np.array([a.size for a in out["x"]["nzval"]], np.int64)
        cap_u = np.array([a.size for a in out["u"]["nzval"]], np.int64)
        args = []
        for S in (Sx, Su):
            args.append(ptr_array([M.colptr for M in S], i64p))
            args.append(ptr_array([M.rowval for M in S], i64p))
            args.append(ptr_array([M.nzval for M in S], u8p) if pass_nzval else None)
        oargs = []
        for name in ("x", "u"):
            oargs += [ptr_array(out[name]["colptr"], i64p), None, ptr_array(out[name]["nzval"], f64p)]
        nnz_x = np.zeros(T, np.int64); nnz_u = np.zeros(T, np.int64)
        self._ck(self.api.sls_synthesize(self.h, T, *args, c0 + 1, c1 - c0, ptr(cap_x, i64p), ptr(cap_u, i64p), *oargs,
                                         ptr(nnz_x, i64p), ptr(nnz_u, i64p), ptr(out["status"], i32p), ptr(out["resid"], f64p)))
        return nnz_x, nnz_u

    def synthesize_window(self, Sx, Su, win, col_range, out=None, pass_nzval=True, pinned=False, values_only=False,
                          row_off=(0, 0)):
        """Shard form (sls_synthesize_window): `win` = (w0, w1) 0-based half-open support window that covers the column
        range `col_range` = (c0, c1) and its halo.  Sx / Su are either full-width JuliaCSC matrices (their colptr is passed
        as a view starting at w0 — the `pointer(S.colptr, w0+1)` form) or window slices with w1-w0 columns.  Outputs are
        the column slab: colptr arrays of c1-c0+1 entries.  values_only: NULL rowval outputs (the caller aliases
        S[t].rowval when the returned counts equal nnz(S[t][:, range])).  row_off = (x, u): the supports' rows are numbered
        in a larger plant of which the handle's plant is the window starting at those offsets (win / col_range stay in
        the handle's own column numbering); output rows come back in the larger plant's numbering."""
        w0, w1 = int(win[0]), int(win[1])
        c0, c1 = int(col_range[0]), int(col_range[1])
        Sx = [_as_julia(M, np.bool_) for M in Sx]
        Su = [_as_julia(M, np.bool_) for M in Su]
        self.T = T = len(Sx)

        def cp_view(M):
            if M.n == w1 - w0:
                return M.colptr
            assert M.n == self.Nx, "supports must be full-width or exactly the window"
            return M.colptr[w0:w1 + 1]
        cps = [[cp_view(M) for M in S] for S in (Sx, Su)]
        if out is None:
            self.nnz_x = np.array([int(cp[c1 - w0] - cp[c0 - w0]) for cp in cps[0]], np.int64)
            self.nnz_u = np.array([int(cp[c1 - w0] - cp[c0 - w0]) for cp in cps[1]], np.int64)
            self.ncols = c1 - c0
            Nx_keep = self.Nx
            try:
                self.Nx = c1 - c0                        # slab colptr arrays: ncols + 1 entries
                out = self.alloc_outputs(pinned=pinned)
            finally:
                self.Nx = Nx_keep
        cap_x = np.array([a.size for a in out["x"]["nzval"]], np.int64)
        cap_u = np.array([a.size for a in out["u"]["nzval"]], np.int64)
        args = []
        for S, cp in ((Sx, cps[0]), (Su, cps[1])):
            args.append(ptr_array(cp, i64p))
            args.append(ptr_array([M.rowval for M in S], i64p))
            args.append(ptr_array([M.nzval for M in S], u8p) if pass_nzval else None)
        oargs = []
        for name in ("x", "u"):
            oargs.append(ptr_array(out[name]["colptr"], i64p))
            oargs.append(None if values_only else ptr_array(out[name]["rowval"], i64p))
            oargs.append(ptr_array(out[name]["nzval"], f64p))
        nnz_x = np.zeros(T, np.int64); nnz_u = np.zeros(T, np.int64)
        self._ck(self.api.sls_synthesize_window(self.h, T, w0 + 1, w1 - w0, int(row_off[0]), int(row_off[1]), *args, c0 + 1, c1 - c0, ptr(cap_x, i64p), ptr(cap_u, i64p),
                                                *oargs, ptr(nnz_x, i64p), ptr(nnz_u, i64p), ptr(out["status"], i32p), ptr(out["resid"], f64p)))
        self.nnz_x, self.nnz_u, self.ncols = nnz_x, nnz_u, c1 - c0
        return out, nnz_x, nnz_u

    def host_register(self, arr):
        """Page-lock a caller-owned numpy array (sls_host_register)."""
        self._ck(self.api.sls_host_register(self.h, C.c_void_p(arr.ctypes.data), arr.nbytes))

    def host_unregister(self, arr):
        self._ck(self.api.sls_host_unregister(self.h, C.c_void_p(arr.ctypes.data)))

    def solve_device(self):
        self._ck(self.api.sls_solve_device(self.h))

    def simulate(self, dist_col, dist_time, steps, dense=False):
        """Batched closed-loop FIR simulation (README.md:62-71) of the controller of the last solve.
        dist_col: 0-based disturbed states j_s; dist_time: 1-based impulse times t_s (README's t).
        Returns dict(xnorm[nscen, steps], unorm[nscen, steps][, x[nscen, steps, Nx], u[nscen, steps, Nu]])."""
        dc = np.ascontiguousarray(np.asarray(dist_col, dtype=np.int64) + 1)
        dt = np.ascontiguousarray(np.asarray(dist_time, dtype=np.int64))
        ns = dc.size
        out = dict(xnorm=np.zeros((ns, steps)), unorm=np.zeros((ns, steps)))
        if dense:
            out["x"] = np.zeros((ns, steps, self.Nx))
            out["u"] = np.zeros((ns, steps, max(self.Nu, 1)))
        self._ck(self.api.sls_simulate(self.h, ns, steps, ptr(dc, i64p), ptr(dt, i64p), ptr(out["xnorm"], f64p),
                                       ptr(out["unorm"], f64p), ptr(out.get("x"), f64p), ptr(out.get("u"), f64p)))
        return out

    def mark(self, which):
        self._ck(self.api.sls_mark(self.h, which))

    def elapsed_ms(self):
        ms = C.c_float()
        self._ck(self.api.sls_elapsed_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def timing(self):
        t = SlsTiming()
        self._ck(self.api.sls_get_timing(self.h, C.byref(t)))
        return t.as_dict()

    def outputs_to_scipy(self, out):
        T, Nx, Nu = self.T, self.Nx, self.Nu
        Phix = [JuliaCSC(Nx, Nx, out["x"]["colptr"][t], out["x"]["rowval"][t], out["x"]["nzval"][t]).to_scipy() for t in range(T)]
        Phiu = [JuliaCSC(Nu, Nx, out["u"]["colptr"][t], out["u"]["rowval"][t], out["u"]["nzval"][t]).to_scipy() for t in range(T)]
        return Phix, Phiu


def _check_plant(P):
    if getattr(P, "feedback", None) != "StateFeedback":
        return False
    return True


def _impulse_pairing_matches(P, Sx, groups):
    """The reference builds I~ = I[:, ii_x] (reduction.jl:22-23): its c-th column is the impulse of the c-th member of the
    group IN THE ORDER OF s_x (first occurrence over the group's columns, reduction.jl:14), while the c-th local column is
    scattered to c_j[c] (synthesis.jl:65-66).  The device pairs every column with its own impulse; that equals the
    reference exactly when the members' first occurrences in s_x follow the group's order — checked here on the patterns."""
    if all(len(g) < 2 for g in groups):
        return True
    pA = (P.A != 0).astype(np.int64).tocsc()
    own = ((sp.csc_matrix(Sx[-1]) != 0).astype(np.int64) @ pA).tocsc()       # column j: its s_x (ascending rows)
    own.sort_indices()
    for g in groups:
        if len(g) < 2:
            continue
        rows = [own.indices[own.indptr[c]:own.indptr[c + 1]] for c in g]
        keys = []
        for c in g:
            first = next((i for i, r in enumerate(rows) if np.searchsorted(r, c) < r.size and r[np.searchsorted(r, c)] == c), None)
            if first is None:
                return False                                                 # a member outside s_x: R is not square in Julia
            keys.append((first, c))
        if any(b <= a for a, b in zip(keys, keys[1:])):
            return False
    return True


def _route(P, S, I):
    """Host-side dispatch of SLS_H2: which kernel family solves (P, 𝓘).  Returns (wx, wu, general, cols, grp_ptr):
    diagonal weights for sls_set_weights (or None), whether the general kernel 2g is needed (sls_set_cost), the planned
    column list (None = all) and the group offsets for sls_set_groups (None = single columns)."""
    wx, wu, why = P.diagonal_cost()
    general = why is not None            # general [C1 D12] / D11 / B1: kernel 2g (sls_set_cost)
    Sx, Su = S
    cols = None
    grp_ptr = None
    if I is not None:
        groups = [[int(c) for c in np.atleast_1d(np.asarray(g))] for g in I]
        cols = [c for g in groups for c in g]
        if any(len(g) != 1 for g in groups):
            if len(set(cols)) != len(cols):
                raise ValueError("a column appears in more than one group")
            if not _impulse_pairing_matches(P, Sx, groups):
                raise ValueError("a group's columns are not in the order in which they first occur in its index set s_x: the "
                                 "reference would pair the impulses I~ with other columns of the group (reduction.jl:14,"
                                 "22-23); list the group's columns in ascending order")
            if general or not _groups_decouple(P, Sx, Su, groups):
                grp_ptr = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(np.int64)
    if grp_ptr is not None and wx is not None:
        general, wx, wu = True, None, None       # coupled groups with diagonal weights: the weights travel as C1 / D12
    if general and P.Nw < P.Nx and (cols is None or max(cols) >= P.Nw):
        raise ValueError("B1 / D11 have fewer columns than the disturbance columns asked for")
    return wx, wu, general, cols, grp_ptr


def SLS_H2(P, S, I=None, *, device: int = 0, return_info: bool = False, api=None):
    """`SLS_𝓗₂(P, 𝓢; 𝓘=nothing)` (src/synthesis.jl:11-32).

    P: Plant (state feedback).  S = [Sx, Su]: two length-T lists of Bool sparse matrices (Nx x Nx, Nu x Nx).
    I: optional list of column groups of 0-based columns (`[[j] for j in ...]` is the reference's default shape,
       synthesis.jl:15; plain ints are accepted too).  Groups with q > 1 ascending columns are solved as one problem
       on the union index set (kernel 2g) unless they provably decouple into their single columns.
    Cost: the identity and diagonal [C1 D12] with a diagonal B1 run on the DMMA kernels; any other C1, D12, D11, B1
       (6-argument Plant) runs on the general kernel (sls_set_cost).
    Returns (Phi_x, Phi_u): two length-T lists of scipy CSC matrices, the layout `eachcol(Φ)` gives in Julia
    (synthesis.jl:27).  The stored pattern is the full in-support mask (explicit zeros kept).
    Like the reference, a plant that is not state-feedback returns None (synthesis.jl:13 falls through)."""
    if not _check_plant(P):
        return None
    wx, wu, general, cols, grp_ptr = _route(P, S, I)
    Sx, Su = S
    syn = Synthesizer(device=device, api=api)
    try:
        if wx is not None:
            syn.set_weights(wx, wu, P.Nx, P.Nu)
        syn.set_plant(P.A, P.B2)
        if general:
            syn.set_cost(P.C1, P.D12, P.D11 if P.D11.nnz else None, None if P._b1_is_identity() else P.B1)
        if grp_ptr is not None:
            syn.set_groups(grp_ptr)
        if cols is None and not return_info and not general:
            out = syn.synthesize(Sx, Su)              # the default 𝓘 = 1..Nx: one pipelined call
            Phix, Phiu = syn.outputs_to_scipy(out)
            info = None
        else:
            syn.set_supports(Sx, Su)
            syn.plan(cols)
            out = syn.solve()
            Phix, Phiu = syn.outputs_to_scipy(out)
            info = dict(status=out["status"][:syn.ncols].copy(), resid=out["resid"][:syn.ncols].copy(),
                        n=syn.n.copy(), m=syn.m.copy(), timing=syn.timing())
    finally:
        syn.close()
    if return_info:
        return Phix, Phiu, info
    return Phix, Phiu


def sparsity_dim_reduction(P, cj, S, *, device: int = 0, api=None):
    """`sparsity_dim_reduction(P, cⱼ, 𝓢)` (src/reduction.jl:11-27), state-feedback branch.
    Returns (P̃, Ĩ, iiₓ, sₓ, sᵤ) with P̃ = Plant(A[sₓ,sₓ], B₁[sₓ,cⱼ], B₂[sₓ,sᵤ]) (the sub-plant the reference's
    `view` describes, GeneralizedPlant.jl:266-285), Ĩ = I(|sₓ|)[:, iiₓ] padded to |cⱼ| columns, and 0-based
    index vectors.  The index sets come from kernel 1 on the GPU."""
    if getattr(P, "feedback", None) == "OutputFeedback":
        return _sparsity_dim_reduction_of(P, cj, S, device=device, api=api)
    Sx, Su = S
    syn = Synthesizer(device=device, api=api)
    try:
        syn.set_plant(P.A, P.B2)
        syn.set_supports([Sx[-1]], [Su[-1]])      # only 𝓢ⱼ[end] is read (reduction.jl:14)
        sx, su, iix = syn.reduce(cj)
    finally:
        syn.close()
    cj = np.asarray(cj, dtype=np.int64)
    Pt = Plant(P.A[sx, :][:, sx], P.B1[sx, :][:, cj], P.B2[sx, :][:, su])
    n, q = sx.size, cj.size
    It = np.zeros((n, q))
    sel = np.nonzero(iix)[0]
    It[sel, np.arange(sel.size)] = 1.0
    return Pt, It, iix, sx, su


def _sparsity_dim_reduction_of(P, cj, S, *, device: int = 0, api=None):
    """The OutputFeedback branch (src/reduction.jl:16-18): 𝓢 = [𝓢ₓ, 𝓢ᵤ, 𝓢ᵧ],
    sₓ, sᵤ, sᵧ = (unique(findnz(𝓢ⱼ[end][:,cⱼ])[1]) for 𝓢ⱼ in 𝓢) and P̃ = view(P, (sₓ, [sₓ;Nx+sᵤ], sᵧ), (sₓ, cⱼ, sᵤ)).
    Returns the reference's five values (P̃, Ĩ, iiₓ, sₓ, sᵤ) — sᵧ is not returned there either; it is kept on P̃.s_y.
    The index sets come from the GPU (sls_reduce_of); P̃ is materialised block by block as the view describes
    (GeneralizedPlant.jl:266-285)."""
    if len(S) != 3:
        raise ValueError("an OutputFeedback plant is reduced with three support families [Sx, Su, Sy] (reduction.jl:17)")
    syn = Synthesizer(device=device, api=api)
    try:
        (sx, su, sy), iix = syn.reduce_of([Sj[-1] for Sj in S], cj)
    finally:
        syn.close()
    cj = np.asarray(cj, dtype=np.int64)
    zr = np.concatenate([sx, P.Nx + su])
    if zr.size and zr.max() >= P.Nz:
        raise ValueError("the view selects rows [s_x; Nx+s_u] of C1/D11/D12: needs Nz >= Nx+Nu rows")
    cut = lambda M, r, c: M[r, :][:, c]
    Pt = Plant(cut(P.A, sx, sx), cut(P.B1, sx, cj), cut(P.B2, sx, su),
               cut(P.C1, zr, sx), cut(P.D11, zr, cj), cut(P.D12, zr, su),
               cut(P.C2, sy, sx), cut(P.D21, sy, cj), cut(P.D22, sy, su))
    Pt.s_y = sy
    n, q = sx.size, cj.size
    It = np.zeros((n, q))
    sel = np.nonzero(iix)[0]
    It[sel, np.arange(sel.size)] = 1.0                       # [I(n)[:,ii_x] 0]  (reduction.jl:22-23)
    return Pt, It, iix, sx, su
