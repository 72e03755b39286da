"""Parity checks shared by the GPU tests (CUDA library, `-m gpu`) and the kernel-logic tests that run the
same kernel sources on the test-only SIMT emulator (`-m "not gpu"`).  Every function takes the bound
library (`api`) explicitly and compares against the CPU oracle / golden vectors."""
import numpy as np
import pytest
import scipy.sparse as sp

import slc_b200
from slc_b200 import _capi
from conftest import PHI_FLOOR, PHI_RTOL, assert_phi_matches, golden_taps, load_golden
from oracle import sls_oracle as O


def chain(Nx, jitter=0.0, d=9, T=29, alpha=1.5):
    A, B1, B2 = O.chain_plant(Nx, jitter=jitter, seed=0)
    return A, B1, B2, O.make_supports(A, B2, d, T, alpha)


def restrict_cols(taps, cols, ncol):
    """golden taps (all columns) -> taps whose columns outside `cols` are emptied"""
    out = []
    keep = np.zeros(ncol, bool); keep[cols] = True
    for indptr, indices, data in taps:
        cnt = np.diff(indptr) * keep
        sel = np.repeat(keep, np.diff(indptr))
        out.append((np.r_[0, np.cumsum(cnt)].astype(np.int64), indices[sel], data[sel]))
    return out


def check_golden(api, name, plant, cols=None, rtol=PHI_RTOL, floor=0.0):
    g = load_golden(name)
    A, B1, B2 = plant
    S = O.make_supports(A, B2, int(g["d"]), int(g["T"]), float(g["alpha"]))
    P = slc_b200.Plant(A, B1, B2)
    gcols = g["cols"]
    use = gcols if cols is None else np.asarray(cols)
    I = None if (cols is None and gcols.size == A.shape[0]) else [[int(c)] for c in use]
    Phix, Phiu, info = slc_b200.SLS_H2(P, S, I=I, return_info=True, api=api)
    tx, tu = golden_taps(g, "x"), golden_taps(g, "u")
    if cols is not None:
        tx, tu = restrict_cols(tx, use, A.shape[0]), restrict_cols(tu, use, A.shape[0])
    wx = assert_phi_matches(Phix, tx, "Phix", rtol, floor)
    wu = assert_phi_matches(Phiu, tu, "Phiu", rtol, floor)
    # n, m per column are integer outputs of kernel 1: exact
    pos = {int(c): i for i, c in enumerate(gcols)}
    idx = [pos[int(c)] for c in use]
    assert np.array_equal(info["n"], g["n"][idx]) and np.array_equal(info["m"], g["m"][idx])
    assert not (info["status"] & _capi.STATUS_RESIDUAL).any()
    assert info["resid"].max() < 1e-12
    return max(wx, wu), info


def check_reduction_fixture(api):
    """/root/reference/test/reduction_test.jl:10-24 through sls_reduce (kernel 1)."""
    A, B1, B2 = O.chain_plant(59)
    pA = O._pat(A)
    Sx = [sp.csc_matrix(O._pat_power(pA, 9) != 0)]
    Su = [sp.csc_matrix((O._pat(sp.csc_matrix(B2).T) @ O._pat_power(pA, 9)) != 0)]
    P = slc_b200.Plant(A, B1, B2)
    Pt, It, iix, sx, su = slc_b200.sparsity_dim_reduction(P, np.arange(20), [Sx, Su], api=api)
    assert np.array_equal(sx, np.arange(30)) and np.array_equal(su, np.arange(10))
    assert np.array_equal(iix, np.r_[np.ones(20, bool), np.zeros(10, bool)])
    assert np.array_equal(It, np.eye(30)[:, :20])
    assert (Pt.A != A[:30, :30]).nnz == 0 and (Pt.B1 != B1[:30, :20]).nnz == 0 and (Pt.B2 != B2[:30, :10]).nnz == 0
    # first-occurrence order for an unsorted group (reduction.jl:14: unique(findnz(...)[1]))
    cj = np.array([40, 3])
    _, _, iix2, sx2, su2 = slc_b200.sparsity_dim_reduction(P, cj, [Sx, Su], api=api)
    (_, _, _), _, oiix, osx, osu = O.sparsity_dim_reduction(O.Plant(A, B1, B2), cj, [Sx, Su])
    assert np.array_equal(sx2, osx) and np.array_equal(su2, osu) and np.array_equal(iix2, oiix)
    assert sx2[0] == 30 and not np.all(np.diff(sx2) > 0)          # genuinely unsorted


def of_fixture(Nx=59):
    """An OutputFeedback plant on the README chain (every third state measured, D21 != 0) and single-tap supports of
    radius 9 for the three families [Sx, Su, Sy] the reference's OF branch iterates over (reduction.jl:17)."""
    A, B1, B2 = O.chain_plant(Nx)
    Nu = B2.shape[1]
    C2 = sp.identity(Nx, format="csc")[np.arange(0, Nx, 3), :]
    Ny = C2.shape[0]
    D21 = sp.csc_matrix(0.1 * np.ones((Ny, Nx)))
    CD = sp.identity(Nx + Nu, format="csc")
    mats = dict(A=A, B1=B1, B2=B2, C1=CD[:, :Nx], D11=sp.csc_matrix((Nx + Nu, Nx)), D12=CD[:, Nx:], C2=C2, D21=D21,
                D22=sp.csc_matrix((Ny, Nu)))
    pA = O._pat(A)
    P9 = O._pat_power(pA, 9)
    Sx = [sp.csc_matrix(P9 != 0)]
    Su = [sp.csc_matrix((O._pat(sp.csc_matrix(B2).T) @ P9) != 0)]
    Sy = [sp.csc_matrix((O._pat(C2) @ P9) != 0)]
    return mats, [Sx, Su, Sy]


def check_reduction_of(api):
    """the OutputFeedback branch of sparsity_dim_reduction (reduction.jl:16-18) through sls_reduce_of
    (sls_union_first_kernel) against the oracle's restatement — index sets in first-occurrence order, ii_x, I~, the
    nine blocks of the view — plus its error paths."""
    mats, S = of_fixture()
    P = slc_b200.Plant(mats["A"], mats["B1"], mats["B2"], mats["C1"], mats["D11"], mats["D12"], mats["C2"], mats["D21"], mats["D22"])
    assert P.feedback == "OutputFeedback" and P.Ny == 20
    assert slc_b200.SLS_H2(P, S[:2], api=api) is None          # synthesis.jl:13: the guard falls through, nothing is returned
    for cj in (np.arange(20), np.array([30, 5, 31]), np.array([58]), np.array([7, 7, 2])):
        Pt, It, iix, sx, su = slc_b200.sparsity_dim_reduction(P, cj, S, api=api)
        sub, oIt, oiix, osx, osu, osy = O.sparsity_dim_reduction_of(mats, cj, S)
        assert np.array_equal(sx, osx) and np.array_equal(su, osu) and np.array_equal(Pt.s_y, osy)
        assert np.array_equal(iix, oiix) and np.array_equal(It, oIt)
        for name in ("A", "B1", "B2", "C1", "D11", "D12"):
            assert np.array_equal(getattr(Pt, name).toarray(), sub[name]), name
        if Pt.feedback == "OutputFeedback":
            for name in ("C2", "D21", "D22"):
                assert np.array_equal(getattr(Pt, name).toarray(), sub[name]), name
    # the family sizes are what the branch is for: no product with pat(A), so s_x of column 30 is the radius-9 band
    _, _, _, sx, su = slc_b200.sparsity_dim_reduction(P, np.array([30, 5, 31]), S, api=api)
    assert np.array_equal(sx, np.r_[21:40, 0:15, 40])
    # error paths: column out of range, malformed rows (descending inside a column), too small a capacity
    syn = slc_b200.Synthesizer(api=api)
    try:
        with pytest.raises(_capi.SlsError):
            syn.reduce_of([S[0][-1], S[1][-1], S[2][-1]], [59])
        bad = slc_b200.JuliaCSC.from_scipy(S[0][-1], np.bool_)
        bad.rowval = bad.rowval.copy()
        bad.rowval[:2] = bad.rowval[:2][::-1]
        with pytest.raises(_capi.SlsError):
            syn.reduce_of([bad], [0])
        (s0,), _ = syn.reduce_of([S[0][-1]], [0])               # the handle stays usable after an error
        assert np.array_equal(s0, np.arange(10))
    finally:
        syn.close()


def check_local_blocks(api, A, B1, B2, S, cols):
    """kernel 1 outputs (index sets, masks, local A~/B2~) bit-exact against the oracle's gather."""
    Po = O.Plant(A, B1, B2)
    prods = O.support_products(Po, S)
    syn = slc_b200.Synthesizer(api=api)
    try:
        syn.set_plant(A, B2)
        syn.set_supports(*S)
        _, _, n, m = syn.plan(cols)
        for k, j in enumerate(cols):
            (At, _, B2t), It, iix, sx, su = O.sparsity_dim_reduction(Po, [j], S, _prod=prods)
            Mx, Mu = O.column_masks(S, sx, su, j)
            gsx, gsu, giix = syn.reduce([j])
            assert np.array_equal(gsx, sx) and np.array_equal(gsu, su) and np.array_equal(giix, iix)
            assert n[k] == sx.size and m[k] == su.size
            dAt, dB2t, dMx, dMu = syn.local_problem(k)
            assert np.array_equal(dAt, At) and np.array_equal(dB2t, B2t)
            assert np.array_equal(dMx, Mx) and np.array_equal(dMu, Mu)
    finally:
        syn.close()


def check_against_oracle(api, A, B1, B2, S, cols, rtol=PHI_RTOL, floor=0.0):
    """Φ columns from the library vs the oracle's exact minimiser on the same columns."""
    Po = O.Plant(A, B1, B2)
    Ox, Ou, infos = O.SLS_H2(Po, S, I=cols, return_info=True)
    Phix, Phiu, info = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), S, I=[[int(c)] for c in cols], return_info=True, api=api)
    worst = 0.0
    for name, G, H in (("Phix", Phix, Ox), ("Phiu", Phiu, Ou)):
        taps = [(np.asarray(M.indptr, np.int64), np.asarray(M.indices, np.int64), M.data) for M in H]
        worst = max(worst, assert_phi_matches(G, taps, name, rtol, floor))
    return worst, info, infos


def check_error_paths(api):
    A, B1, B2, S = chain(30, d=3, T=6)
    syn = slc_b200.Synthesizer(api=api)
    try:
        with np.testing.assert_raises(_capi.SlsError):
            syn.plan([0])                                         # nothing set yet
        syn.set_plant(A, B2)
        with np.testing.assert_raises(_capi.SlsError):
            syn.plan([0])                                         # supports missing
        syn.set_supports(*S)
        for bad in ([30], [-1], [3, 3]):
            with np.testing.assert_raises(_capi.SlsError):
                syn.plan(bad)
        with np.testing.assert_raises(_capi.SlsError):
            syn.solve_device()                                    # failed plan leaves no valid plan
        # malformed CSC
        J = slc_b200.JuliaCSC.from_scipy(A, np.float64)
        J.colptr = J.colptr.copy(); J.colptr[0] = 0
        with np.testing.assert_raises(_capi.SlsError):
            syn.set_plant(J, B2)
        # empty column list is legal: all outputs empty
        syn.set_plant(A, B2); syn.set_supports(*S)
        nnz_x, nnz_u, n, m = syn.plan([])
        assert nnz_x.sum() == 0 and nnz_u.sum() == 0
        out = syn.solve()
        assert all((cp == 1).all() for cp in out["x"]["colptr"] + out["u"]["colptr"])
    finally:
        syn.close()


def check_unsorted_plan(api):
    """A column list in arbitrary order (kernel 3's non-contiguous store path) gives the same Φ as the sorted list, bit for bit."""
    A, B1, B2, S = chain(40)
    P = slc_b200.Plant(A, B1, B2)
    cols = [35, 3, 20, 21, 0, 39, 7]
    X1, U1 = slc_b200.SLS_H2(P, S, I=[[c] for c in cols], api=api)
    X2, U2 = slc_b200.SLS_H2(P, S, I=[[c] for c in sorted(cols)], api=api)
    for a, b in zip(X1 + U1, X2 + U2):
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)


def check_masks_and_flags(api):
    """stored-false support entries, stored zeros in A, non-nested supports, an infeasible column."""
    A, B1, B2, S = chain(59)
    Sx, Su = [M.copy() for M in S[0]], [M.copy() for M in S[1]]
    # a stored `false` inside 𝓢ₓ[12][:,20] fixes that entry (synthesis.jl:58: `.≠ 1`)
    M = Sx[11]; e = M.indptr[20] + 1; M.data[e] = False
    # non-nested: drop one row of tap 15 entirely for column 21
    M = Sx[14].tolil(); M[24, 21] = False; Sx[14] = sp.csc_matrix(M.tocsc()); Sx[14].eliminate_zeros()
    # explicitly stored 0.0 in A: `A .≠ 0` drops it (reduction.jl:14)
    Az = A.tocoo()
    A2 = sp.csc_matrix((np.r_[Az.data, 0.0], (np.r_[Az.row, 10], np.r_[Az.col, 30])), shape=A.shape)
    A2.sort_indices()
    assert A2.nnz == A.nnz + 1
    cols = [0, 19, 20, 21, 22, 30, 58]
    worst, info, infos = check_against_oracle(api, A2, B1, B2, [Sx, Su], cols)
    assert max(i["residual"] for i in infos.values()) < 1e-12        # the modified problems stay feasible
    assert not (info["status"] & _capi.STATUS_RESIDUAL).any()
    # infeasible: remove every actuator support => residual flag, like the oracle's residual
    Su0 = [sp.csc_matrix(M.shape, dtype=bool) for M in Su]
    Phix, Phiu, info = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), [Sx, Su0], I=[[20]], return_info=True, api=api)
    assert info["status"][0] & _capi.STATUS_RESIDUAL and info["resid"][0] > 1e-3
    return worst


def check_gather_paths(api):
    """Kernel 1's fast paths (index sets / masks of banded plants from the end rows of each support segment) against the
    general paths and the oracle, bit for bit, on inputs that mix them inside one plan and one column:
    a gap in a LAST-tap segment (1a: sorted-union path for the three columns that read it, bitmap path for the rest),
    a gap in an intermediate tap (1b: that support column on the per-entry list, the other 2T-1 on the lane-per-tap path),
    a stored `false` inside a run, a stored `false` at the END of a run (the end-row test alone would accept it), and a
    long-range coupling in A that spreads a column's segments over more than 1 024 rows (bitmap path refused)."""
    A, B1, B2, S = chain(80)
    Sx, Su = [M.copy() for M in S[0]], [M.copy() for M in S[1]]
    T = len(Sx)
    def drop(M, r, c):
        L = M.tolil(); L[r, c] = False; M2 = sp.csc_matrix(L.tocsc()); M2.eliminate_zeros(); M2.sort_indices(); return M2
    Sx[T - 1] = drop(Sx[T - 1], 41, 40)                         # gap inside the last-tap segment of column 40
    Sx[12] = drop(Sx[12], 52, 50)                               # gap in an intermediate tap of column 50
    M = Sx[20]; M.data[M.indptr[30] + 2] = False                # stored false inside the run of column 30
    M = Sx[21]; M.data[M.indptr[31 + 1] - 1] = False            # stored false at the end of the run of column 31
    M = Su[15]; M.data[M.indptr[33]] = False                    # and at the start of an actuator run
    cols = [0, 29, 30, 31, 33, 39, 40, 41, 49, 50, 51, 79]
    check_local_blocks(api, A, B1, B2, [Sx, Su], cols)
    worst, info, infos = check_against_oracle(api, A, B1, B2, [Sx, Su], cols)
    ok = [k for k, j in enumerate(cols) if infos[j]["residual"] < 1e-10]
    assert len(ok) >= len(cols) - 3                              # dropping a support row can make a column infeasible
    assert not (info["status"][ok] & _capi.STATUS_RESIDUAL).any()
    # long-range coupling: column 1300 of A also drives state 60, so its index set is two runs 1 240 rows apart
    Nx = 1400
    A2, B1b, B2b = O.chain_plant(Nx)
    A2 = A2.tolil(); A2[60, 1300] = 0.2; A2 = sp.csc_matrix(A2.tocsc()); A2.sort_indices()
    S2 = O.make_supports(A2, B2b, 4, 8, 1.5)
    check_local_blocks(api, A2, B1b, B2b, S2, [59, 60, 61, 1299, 1300, 1301])
    return worst


def check_large_local_problems(api):
    """Local state dimensions above 64 (sls_solve_huge.cu): a fully actuated 15 x 15 grid with d = 5 (centre column n = 85,
    m = 113: the n > 64 kernel beside the DMMA kernels of the smaller edge / corner columns in one plan) and d = 6 (n = 113,
    m = 141: input masks of five words, every class on the n > 64 kernel).  Index sets / masks / local blocks bit-exact, Phi
    against the oracle's exact minimiser."""
    A, B1, B2 = O.grid_plant(15, 15, actuation="full")
    worst = 0.0
    for d, T, cols, nexp in ((5, 8, [112, 0, 7, 16], 85), (6, 8, [112, 0], 113)):
        S = O.make_supports(A, B2, d, T, 1.5)
        check_local_blocks(api, A, B1, B2, S, cols)
        w, info, infos = check_against_oracle(api, A, B1, B2, S, cols, floor=PHI_FLOOR)
        assert int(info["n"][0]) == nexp, info["n"]
        assert max(i["residual"] for i in infos.values()) < 1e-12
        assert not (info["status"] & _capi.STATUS_RESIDUAL).any() and float(info["resid"].max()) < 1e-13
        worst = max(worst, w)
    return worst


def check_simulation(api, Nx, scenarios, steps=120, tol=1e-13):
    """sls_simulate (windowed, on device) against the oracle's dense README recursion, scenario by scenario."""
    A, B1, B2, S = chain(Nx)
    syn = slc_b200.Synthesizer(api=api)
    try:
        syn.set_plant(A, B2)
        syn.set_supports(*S)
        syn.plan(None)
        out = syn.solve()
        Phix, Phiu = syn.outputs_to_scipy(out)
        sim = syn.simulate([j for j, _ in scenarios], [t for _, t in scenarios], steps, dense=True)
        tm = syn.timing()
    finally:
        syn.close()
    Po = O.Plant(A, B1, B2)
    T = len(Phix)
    worst = 0.0
    for s, (j, ts) in enumerate(scenarios):
        w = np.zeros((Nx, steps)); w[j, ts - 1] = 1.0
        x, u, beta = O.simulate_closed_loop(Po, Phix, Phiu, w, steps)
        worst = max(worst, np.abs(sim["x"][s].T - x).max(), np.abs(sim["u"][s].T - u).max())
        assert np.abs(sim["xnorm"][s] - np.linalg.norm(x, axis=0)).max() < tol
        assert np.abs(sim["unorm"][s] - np.linalg.norm(u, axis=0)).max() < tol
        # closed-loop identity: x(ts+k) = Phi_x[k][:, j], zero after the FIR horizon
        for k in range(1, T + 1):
            if ts + k <= steps:
                assert np.abs(sim["x"][s][ts + k - 1] - Phix[k - 1][:, j].toarray().ravel()).max() < 1e-12
        if ts + T + 1 <= steps:
            assert np.abs(sim["x"][s][ts + T:]).max() < 1e-12
    assert worst < tol, worst
    return worst, tm


def check_device_supports(api, plant, d, T, alpha):
    """sls_make_supports (device) == the README's Boolean matrix powers (oracle.make_supports), entry for entry;
    and a synthesis on device-generated supports == the synthesis on host-supplied ones, bit for bit."""
    A, B1, B2 = plant
    Ox, Ou = O.make_supports(A, B2, d, T, alpha)
    syn = slc_b200.Synthesizer(api=api)
    try:
        syn.set_plant(A, B2)
        nnz_x, nnz_u = syn.make_supports(d, T, alpha)
        Sx, Su = syn.get_supports()
        for t in range(T):
            for J, M in ((Sx[t], Ox[t]), (Su[t], Ou[t])):
                M = M.tocsc(); M.sort_indices()
                assert np.array_equal(J.colptr - 1, M.indptr) and np.array_equal(J.rowval - 1, M.indices) and J.nzval.all()
        assert np.array_equal(nnz_x, [M.nnz for M in Ox]) and np.array_equal(nnz_u, [M.nnz for M in Ou])
        syn.plan(None)
        a = syn.solve()
        syn.set_supports(Ox, Ou)
        syn.plan(None)
        b = syn.solve()
        for name in ("x", "u"):
            for t in range(T):
                assert np.array_equal(a[name]["rowval"][t], b[name]["rowval"][t])
                assert np.array_equal(a[name]["nzval"][t], b[name]["nzval"][t])
    finally:
        syn.close()


def check_pipelined(api, plant, S, chunks, pass_nzval=True):
    """sls_synthesize (one-shot, column-chunked pipeline) == sls_set_supports + sls_plan + sls_solve, bit for bit."""
    import os
    A, B1, B2 = plant
    syn = slc_b200.Synthesizer(api=api)
    old = os.environ.get("SLS_PIPE_CHUNKS")
    try:
        syn.set_plant(A, B2)
        syn.set_supports(*S, pass_nzval=pass_nzval)
        syn.plan(None)
        ref = syn.solve()
        nnz_ref = (syn.nnz_x.copy(), syn.nnz_u.copy())
        for K in chunks:
            os.environ["SLS_PIPE_CHUNKS"] = str(K)
            out = syn.synthesize(*S, pass_nzval=pass_nzval)
            assert np.array_equal(syn.nnz_x, nnz_ref[0]) and np.array_equal(syn.nnz_u, nnz_ref[1]), K
            for fam in ("x", "u"):
                for t in range(syn.T):
                    for part in ("colptr", "rowval", "nzval"):
                        assert np.array_equal(out[fam][part][t], ref[fam][part][t]), (K, fam, t, part)
            assert np.array_equal(out["status"][:syn.Nx], ref["status"][:syn.Nx])
            assert np.array_equal(out["resid"][:syn.Nx], ref["resid"][:syn.Nx])
        # a contiguous column range (one GPU's shard) == the three-call path on that column list
        c0, c1 = syn.Nx // 3, (2 * syn.Nx) // 3 + 1
        syn.set_supports(*S, pass_nzval=pass_nzval)
        syn.plan(np.arange(c0, c1))
        ref2 = syn.solve()
        os.environ["SLS_PIPE_CHUNKS"] = str(chunks[-1])
        out2 = syn.synthesize(*S, pass_nzval=pass_nzval, col_range=(c0, c1))
        for fam in ("x", "u"):
            for t in range(syn.T):
                for part in ("colptr", "rowval", "nzval"):
                    assert np.array_equal(out2[fam][part][t], ref2[fam][part][t]), ("range", fam, t, part)
        assert np.array_equal(out2["status"][:c1 - c0], ref2["status"][:c1 - c0])
        assert np.array_equal(out2["resid"][:c1 - c0], ref2["resid"][:c1 - c0])
        # too-small output capacity is an argument error, not a crash
        syn.nnz_x, syn.nnz_u, syn.ncols = nnz_ref[0], nnz_ref[1], syn.Nx
        small = syn.alloc_outputs()
        small["x"]["rowval"][-1] = small["x"]["rowval"][-1][:-1]
        small["x"]["nzval"][-1] = small["x"]["nzval"][-1][:-1]
        with pytest.raises(_capi.SlsError) as ei:
            syn.synthesize(*S, out=small)
        assert ei.value.code == -4
        # a broken colptr in a later column chunk is reported by the feeder thread as a layout error (-2), not a crash
        os.environ["SLS_PIPE_CHUNKS"] = "3"
        Sx_bad = [slc_b200.JuliaCSC.from_scipy(sp.csc_matrix(M), np.bool_) for M in S[0]]
        cp = Sx_bad[1].colptr.copy(); j = (3 * syn.Nx) // 4; cp[j] = cp[j + 1] + 1
        Sx_bad[1] = slc_b200.JuliaCSC(Sx_bad[1].m, Sx_bad[1].n, cp, Sx_bad[1].rowval, Sx_bad[1].nzval)
        with pytest.raises(_capi.SlsError) as ei:
            syn.synthesize(Sx_bad, S[1])
        assert ei.value.code == -2
        # the handle is still usable afterwards
        out = syn.synthesize(*S)
        assert np.array_equal(out["x"]["nzval"][0], ref["x"]["nzval"][0])
    finally:
        if old is None:
            os.environ.pop("SLS_PIPE_CHUNKS", None)
        else:
            os.environ["SLS_PIPE_CHUNKS"] = old
        syn.close()


def check_pipelined_rowval_sources(api, Nx=24, d=3, T=8):
    """sls_synthesize fills a tap's rowval output from the caller's own S[t].rowval when the stored-entry count over a column
    chunk equals the support's (stored rows are a subset of the support column, so equal counts mean equal rows) and from the
    device otherwise.  Supports with stored-false entries and with rows outside s_x take the device branch for the taps they
    touch; every output must equal the three-call path bit for bit, whatever the source, and flags bit 3 turns the host
    branch off."""
    import os
    A, B1, B2, S = chain(Nx, d=d, T=T)
    Sx, Su = [sp.csc_matrix(M).copy() for M in S[0]], [sp.csc_matrix(M).copy() for M in S[1]]
    for M in Sx + Su:
        M.sort_indices()
    M = Sx[3]; M.data[M.indptr[5] + 1] = False                      # stored false: the entry is fixed to 0, not stored in Phi
    M = Su[2]; M.data[M.indptr[7]] = False
    L = Sx[4].tolil(); L[min(Nx - 1, 10 + d + 4), 10] = True; Sx[4] = sp.csc_matrix(L.tocsc()); Sx[4].sort_indices()   # a row outside s_x
    touched = {("x", 3), ("u", 2), ("x", 4)}
    check_pipelined(api, (A, B1, B2), [Sx, Su], chunks=[1, 3])        # bit for bit against the three-call path
    old = os.environ.get("SLS_PIPE_CHUNKS")
    os.environ["SLS_PIPE_CHUNKS"] = "1"
    try:
        for flags in (0, 8):
            syn = slc_b200.Synthesizer(api=api, flags=flags)
            try:
                syn.set_plant(A, B2)
                out = syn.synthesize(Sx, Su)
                hcb = syn.timing()["host_copied_bytes"]
                expect = 0 if flags else 8 * sum(out[f]["rowval"][t].size for f in ("x", "u") for t in range(T) if (f, t) not in touched)
                assert hcb == expect, (flags, hcb, expect)
                for f, t in touched:                                  # the touched taps really store fewer entries than their support
                    assert out[f]["rowval"][t].size < (Sx if f == "x" else Su)[t].nnz
                if flags == 0:
                    ref = out
                else:
                    for f in ("x", "u"):
                        for t in range(T):
                            for part in ("colptr", "rowval", "nzval"):
                                assert np.array_equal(out[f][part][t], ref[f][part][t]), (f, t, part)
            finally:
                syn.close()
    finally:
        if old is None:
            os.environ.pop("SLS_PIPE_CHUNKS", None)
        else:
            os.environ["SLS_PIPE_CHUNKS"] = old


def check_window(api, plant, S, ranges, halo=1, chunks=(1, 3), sub_reach=None):
    """sls_synthesize_window (support window in, column slab out) == the three-call path on the same columns, bit for
    bit — for both ways of passing a window (a colptr view into the full matrices; rebased slices), and with the
    rowval outputs left out.  Also: the handle then refuses plans that reach outside the uploaded window, a window that
    does not cover the halo is an argument error, malformed support rows are caught by the device-side check."""
    import os
    A, B1, B2 = plant
    Nx = A.shape[0]
    syn = slc_b200.Synthesizer(api=api)
    old = os.environ.get("SLS_PIPE_CHUNKS")
    try:
        syn.set_plant(A, B2)
        Sxj = [slc_b200.JuliaCSC.from_scipy(sp.csc_matrix(M), np.bool_) for M in S[0]]
        Suj = [slc_b200.JuliaCSC.from_scipy(sp.csc_matrix(M), np.bool_) for M in S[1]]
        T = len(Sxj)
        for (c0, c1) in ranges:
            w0, w1 = max(0, c0 - halo), min(Nx, c1 + halo)
            syn.set_supports(*S)
            syn.plan(np.arange(c0, c1))
            ref = syn.solve()
            for K in chunks:
                os.environ["SLS_PIPE_CHUNKS"] = str(K)
                for form in ("view", "slice", "values_only"):
                    if form == "slice":
                        def sl(M):
                            lo, hi = int(M.colptr[w0] - 1), int(M.colptr[w1] - 1)
                            return slc_b200.JuliaCSC(M.m, w1 - w0, np.ascontiguousarray(M.colptr[w0:w1 + 1] - lo),
                                                     np.ascontiguousarray(M.rowval[lo:hi]), np.ascontiguousarray(M.nzval[lo:hi]))
                        X, U = [sl(M) for M in Sxj], [sl(M) for M in Suj]
                    else:
                        X, U = Sxj, Suj
                    out, nnz_x, nnz_u = syn.synthesize_window(X, U, (w0, w1), (c0, c1), values_only=(form == "values_only"))
                    for fam, nnz in (("x", nnz_x), ("u", nnz_u)):
                        for t in range(T):
                            full_cp = ref[fam]["colptr"][t]
                            lo, hi = int(full_cp[c0] - 1), int(full_cp[c1] - 1)
                            assert nnz[t] == hi - lo
                            assert np.array_equal(out[fam]["colptr"][t], full_cp[c0:c1 + 1] - lo), (form, K, fam, t)
                            assert np.array_equal(out[fam]["nzval"][t][:hi - lo], ref[fam]["nzval"][t][lo:hi]), (form, K, fam, t)
                            if form != "values_only":
                                assert np.array_equal(out[fam]["rowval"][t][:hi - lo], ref[fam]["rowval"][t][lo:hi]), (form, K, fam, t)
                    assert np.array_equal(out["status"][:c1 - c0], ref["status"][:c1 - c0])
                    assert np.array_equal(out["resid"][:c1 - c0], ref["resid"][:c1 - c0])
        # sub-plant shard: the handle holds P[pw, pw] only, the supports keep the large plant's row numbers
        if sub_reach is not None:
            sub = slc_b200.Synthesizer(api=api)
            try:
                for (c0, c1) in ranges:
                    p0, p1 = max(0, c0 - sub_reach), min(Nx, c1 + sub_reach)
                    ucols = np.unique(sp.csc_matrix(B2)[p0:p1, :].nonzero()[1])
                    assert ucols.size == 0 or np.array_equal(ucols, np.arange(ucols[0], ucols[-1] + 1))
                    u0 = int(ucols[0]) if ucols.size else 0
                    sub.set_plant(sp.csc_matrix(A)[p0:p1, :][:, p0:p1], sp.csc_matrix(B2)[p0:p1, :][:, ucols])
                    w0, w1 = max(0, c0 - halo), min(Nx, c1 + halo)

                    def sl(M):
                        lo, hi = int(M.colptr[w0] - 1), int(M.colptr[w1] - 1)
                        return slc_b200.JuliaCSC(M.m, w1 - w0, np.ascontiguousarray(M.colptr[w0:w1 + 1] - lo),
                                                 np.ascontiguousarray(M.rowval[lo:hi]), np.ascontiguousarray(M.nzval[lo:hi]))
                    syn.set_supports(*S)
                    syn.plan(np.arange(c0, c1))
                    ref = syn.solve()
                    out, nnz_x, nnz_u = sub.synthesize_window([sl(M) for M in Sxj], [sl(M) for M in Suj], (w0 - p0, w1 - p0),
                                                              (c0 - p0, c1 - p0), row_off=(p0, u0))
                    for fam, nnz in (("x", nnz_x), ("u", nnz_u)):
                        for t in range(T):
                            full_cp = ref[fam]["colptr"][t]
                            lo, hi = int(full_cp[c0] - 1), int(full_cp[c1] - 1)
                            assert nnz[t] == hi - lo
                            assert np.array_equal(out[fam]["colptr"][t], full_cp[c0:c1 + 1] - lo), ("sub", fam, t)
                            assert np.array_equal(out[fam]["rowval"][t][:hi - lo], ref[fam]["rowval"][t][lo:hi]), ("sub", fam, t)
                            assert np.array_equal(out[fam]["nzval"][t][:hi - lo], ref[fam]["nzval"][t][lo:hi]), ("sub", fam, t)
                    assert np.array_equal(out["resid"][:c1 - c0], ref["resid"][:c1 - c0])
                # a sub-plant that is too small for the supports is a layout error (rows fall outside the handle's plant)
                c0, c1 = ranges[len(ranges) // 2]
                p0, p1 = max(0, c0 - halo), min(Nx, c1 + halo)
                w0, w1 = p0, p1                                 # (sl() slices the supports to [w0, w1))
                if p0 > 0:
                    ucols = np.unique(sp.csc_matrix(B2)[p0:p1, :].nonzero()[1])
                    sub.set_plant(sp.csc_matrix(A)[p0:p1, :][:, p0:p1], sp.csc_matrix(B2)[p0:p1, :][:, ucols])
                    with pytest.raises(_capi.SlsError) as ei:
                        sub.synthesize_window([sl(M) for M in Sxj], [sl(M) for M in Suj], (0, p1 - p0), (c0 - p0, c1 - p0),
                                              row_off=(p0, int(ucols[0])))
                    assert ei.value.code == -2
            finally:
                sub.close()
        # the window of the last call is all the device holds: planning a column outside it is refused, not garbage
        c0, c1 = ranges[-1]
        syn.synthesize_window(Sxj, Suj, (max(0, c0 - halo), min(Nx, c1 + halo)), (c0, c1))
        far = 0 if c0 > halo + 1 else Nx - 1
        if not (max(0, c0 - halo) <= far < min(Nx, c1 + halo)):
            with pytest.raises(_capi.SlsError) as ei:
                syn.plan([far])
            assert ei.value.code == -3
        # a window without the halo
        if c0 > 0:
            with pytest.raises(_capi.SlsError) as ei:
                syn.synthesize_window(Sxj, Suj, (c0, min(Nx, c1 + halo)), (c0, c1))
            assert ei.value.code == -3
        # malformed rows (descending inside a column) are caught on the device, in both paths
        bad = [slc_b200.JuliaCSC(M.m, M.n, M.colptr, M.rowval.copy(), M.nzval) for M in Sxj]
        M = bad[T - 1]; e = int(M.colptr[c0] - 1)
        assert M.colptr[c0 + 1] - M.colptr[c0] >= 2
        M.rowval[e], M.rowval[e + 1] = M.rowval[e + 1], M.rowval[e]
        with pytest.raises(_capi.SlsError) as ei:
            syn.synthesize_window(bad, Suj, (max(0, c0 - halo), min(Nx, c1 + halo)), (c0, c1))
        assert ei.value.code == -2
        with pytest.raises(_capi.SlsError) as ei:
            syn.set_supports(bad, Suj)
        assert ei.value.code == -2
        M.rowval[e] = Nx + 5                                    # out of range
        with pytest.raises(_capi.SlsError) as ei:
            syn.set_supports(bad, Suj)
        assert ei.value.code == -2
        # the handle is still usable
        syn.set_supports(*S)
        syn.plan(np.arange(c0, c1))
        again = syn.solve()
        assert np.array_equal(again["x"]["nzval"][T - 1], ref["x"]["nzval"][T - 1])
        # duplicates in a reduction group are first occurrences (reduction.jl:14), not an error
        a = syn.reduce([c0, c0 + 1, c0])
        b = syn.reduce([c0, c0 + 1])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    finally:
        if old is None:
            os.environ.pop("SLS_PIPE_CHUNKS", None)
        else:
            os.environ["SLS_PIPE_CHUNKS"] = old
        syn.close()


def check_weighted_cost(api, cols=(3, 14, 39)):
    """6-argument Plant with diagonal [C1 D12] (+ a diagonal B1 and a q = 2 group) through SLS_H2 -> sls_set_weights,
    against the oracle's general restatement of synthesis.jl:42-43,50,76-77."""
    Nx, d, T = 40, 6, 20
    A, B1, B2 = O.chain_plant(Nx)
    S = O.make_supports(A, B2, d, T, 1.5)
    rng = np.random.default_rng(0)
    Nu = B2.shape[1]
    wx, wu = rng.uniform(0.5, 2, Nx), rng.uniform(0.5, 2, Nu)
    C1 = sp.vstack([sp.diags(wx), sp.csc_matrix((Nu, Nx))]).tocsc()
    D12 = sp.vstack([sp.csc_matrix((Nx, Nu)), sp.diags(wu)]).tocsc()
    B1d = sp.diags(np.linspace(0.5, 2, Nx)).tocsc()
    cols = list(cols)
    groups = [cols[:2], cols[2:]]
    Ox, Ou = O.SLS_H2_general(O.Plant(A, B1d, B2, C1, 0, D12), S, I=groups)
    Gx, Gu = slc_b200.SLS_H2(slc_b200.Plant(A, B1d, B2, C1, 0, D12), S, I=groups, api=api)
    worst = 0.0
    for G, H in zip(Gx + Gu, Ox + Ou):
        assert np.array_equal(G.indptr, H.indptr) and np.array_equal(G.indices, H.indices)
        nb = np.linalg.norm(H.data)
        if nb > 0:
            worst = max(worst, np.linalg.norm(G.data - H.data) / nb)
    assert worst < PHI_RTOL, worst
    # all columns, one-shot pipelined path with weights == three-call path with weights
    syn = slc_b200.Synthesizer(api=api)
    try:
        syn.set_weights(wx, wu, Nx, Nu)
        syn.set_plant(A, B2)
        syn.set_supports(*S)
        syn.plan(None)
        ref = syn.solve()
        out = syn.synthesize(*S)
        for fam in ("x", "u"):
            for t in range(T):
                assert np.array_equal(out[fam]["nzval"][t], ref[fam]["nzval"][t])
        # clearing the weights needs the plant again
        syn.set_weights(None, None, Nx, Nu)
        with pytest.raises(_capi.SlsError):
            syn.set_supports(*S)
    finally:
        syn.close()
    return worst


def _general_plant(Nx, seed, dense_cost=True, b1_band=True, d11=True):
    """Chain plant with a non-diagonal cost: banded SPD-ish C1 / D12 blocks (rows [x; u] as the reduced plant selects
    rows [s_x; Nx+s_u], reduction.jl:15), a tridiagonal B1 and a sparse D11."""
    A, B1, B2 = O.chain_plant(Nx)
    Nu = B2.shape[1]
    rng = np.random.default_rng(seed)
    if dense_cost:
        Qh = sp.diags([rng.uniform(1.0, 2.0, Nx), rng.uniform(-0.4, 0.4, Nx - 1), rng.uniform(-0.3, 0.3, Nx - 1)], [0, 1, -1])
        Rh = sp.diags([rng.uniform(1.0, 2.0, Nu), rng.uniform(-0.3, 0.3, Nu - 1)], [0, 1])
        Xu = sp.random(Nx, Nu, density=0.05, random_state=seed, data_rvs=lambda k: rng.uniform(-0.2, 0.2, k))
        C1 = sp.vstack([Qh, 0.1 * Xu.T]).tocsc()                # state rows, then input rows
        D12 = sp.vstack([Xu, Rh]).tocsc()
    else:
        CD = sp.identity(Nx + Nu, format="csc"); C1, D12 = CD[:, :Nx], CD[:, Nx:]
    B1g = sp.diags([np.linspace(0.8, 1.5, Nx), rng.uniform(-0.5, 0.5, Nx - 1), rng.uniform(-0.5, 0.5, Nx - 1)], [0, 1, -1]).tocsc() if b1_band else B1
    D11 = sp.random(Nx + Nu, Nx, density=0.04, random_state=seed + 1, data_rvs=lambda k: rng.uniform(-0.3, 0.3, k)).tocsc() if d11 else 0
    return A, B1g, B2, C1, D11, D12


def check_general_cost(api, Nx=30, d=6, T=12, cases=None):
    """SURVEY §8 row f3 on the device: non-diagonal [C1 D12], D11 != 0, non-diagonal B1 and coupled q = 2 / q = 3 groups
    through SLS_H2 -> sls_set_cost / sls_set_groups -> kernel 2g, against the oracle's null-space restatement of
    synthesis.jl:42-43,50,76-77 (SLS_H2_general).  Patterns bit-exact, values to PHI_RTOL per tap."""
    worst = 0.0
    cases = cases or [
        dict(kw=dict(dense_cost=True, b1_band=False, d11=False), I=[[3], [14], [29]]),          # dense cost, single columns
        dict(kw=dict(dense_cost=True, b1_band=False, d11=True), I=[[0], [15]]),                  # + linear term D11
        dict(kw=dict(dense_cost=False, b1_band=True, d11=False), I=[[4, 5], [20, 22]]),          # identity cost, B1 couples a group
        dict(kw=dict(dense_cost=True, b1_band=True, d11=True), I=[[10, 11, 12], [0, 1], [29]]),  # everything at once
    ]
    for case in cases:
        A, B1, B2, C1, D11, D12 = _general_plant(Nx, 3, **case["kw"])
        S = O.make_supports(A, B2, d, T, 1.5)
        Ox, Ou = O.SLS_H2_general(O.Plant(A, B1, B2, C1, D11, D12), S, I=case["I"])
        Gx, Gu, info = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2, C1, D11, D12), S, I=case["I"], api=api, return_info=True)
        assert (info["status"] & ~2 == 0).all(), info["status"]
        assert info["resid"].max() < 1e-12, info["resid"]
        for G, H in zip(Gx + Gu, Ox + Ou):
            assert np.array_equal(G.indptr, H.indptr) and np.array_equal(G.indices, H.indices)
            nb = np.linalg.norm(H.data)
            if nb > 0:
                worst = max(worst, np.linalg.norm(G.data - H.data) / nb)
        assert worst < PHI_RTOL, (case, worst)
    return worst


def check_coupled_groups_identity_cost(api, Nx=30, d=6, T=12):
    """A q = 2 group whose union index set adds non-empty constraint rows to a member (supports not closed under the
    plant's pattern: the case _groups_decouple refuses) with the README cost: kernel 2g vs the oracle."""
    A, B1, B2 = O.chain_plant(Nx)
    S = O.make_supports(A, B2, d, T, 1.5)
    groups = [[2, 9], [15, 16], [25]]
    Ox, Ou = O.SLS_H2_general(O.Plant(A, B1, B2), S, I=groups)
    Gx, Gu, info = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), S, I=groups, api=api, return_info=True)
    worst = 0.0
    for G, H in zip(Gx + Gu, Ox + Ou):
        assert np.array_equal(G.indptr, H.indptr) and np.array_equal(G.indices, H.indices)
        nb = np.linalg.norm(H.data)
        if nb > 0:
            worst = max(worst, np.linalg.norm(G.data - H.data) / nb)
    assert worst < PHI_RTOL, worst
    return worst, info


def check_general_error_paths(api):
    """Refusals and flags of the general path: one-shot call refused, weights + cost refused, groups that do not cover the
    plan, a singular cost (flag, not garbage), a group too large for shared memory."""
    Nx, d, T = 30, 6, 12
    A, B1, B2 = O.chain_plant(Nx)
    Nu = B2.shape[1]
    S = O.make_supports(A, B2, d, T, 1.5)
    CD = sp.identity(Nx + Nu, format="csc")
    syn = slc_b200.Synthesizer(api=api)
    try:
        with np.testing.assert_raises(_capi.SlsError):
            syn.set_cost(CD[:, :Nx], CD[:, Nx:])                  # plant first
        syn.set_plant(A, B2)
        syn.set_cost(CD[:, :Nx], CD[:, Nx:])
        with np.testing.assert_raises(_capi.SlsError):
            syn.synthesize(*S)                                    # the pipelined one-shot path has no general kernel
        syn.set_supports(*S); syn.plan([3, 4, 10])
        syn.set_groups([0, 2])                                    # does not cover the three planned columns
        with np.testing.assert_raises(_capi.SlsError):
            syn.solve()
        syn.set_groups([0, 2, 3])
        ident = syn.solve()
        assert (ident["status"][:3] & ~2 == 0).all()
        # identity installed as a "general" cost == the DMMA kernels' answer (different algorithms: to rounding)
        syn.set_cost(None, None); syn.set_groups(None)
        ref = syn.solve()
        for fam in ("x", "u"):
            for t in range(T):
                np.testing.assert_allclose(ident[fam]["nzval"][t], ref[fam]["nzval"][t], rtol=0, atol=1e-11)
        # a cost that ignores the inputs entirely: singular on the free u entries -> flagged
        Z12 = sp.csc_matrix((Nx + Nu, Nu))
        syn.set_cost(CD[:, :Nx], Z12); syn.plan([10])
        out = syn.solve()
        assert out["status"][0] & 8, out["status"][:1]
        # all columns as ONE group: q(n+m) far beyond shared memory
        syn.set_cost(None, None); syn.plan(None); syn.set_groups([0, Nx])
        with np.testing.assert_raises(_capi.SlsError):
            syn.solve()
        syn.set_groups(None)
        # weights and a general cost exclude each other
        syn.set_weights(np.full(Nx, 2.0), None, Nx, Nu); syn.set_plant(A, B2)
        with np.testing.assert_raises(_capi.SlsError):
            syn.set_cost(CD[:, :Nx], CD[:, Nx:])
    finally:
        syn.close()
    with np.testing.assert_raises(ValueError):
        slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), S, I=[[5, 3]], api=api)       # 3 occurs in s_x before 5: reference pairing differs


def random_banded_plant(Nx, bw, seed, act_every=3):
    """Random banded A (bandwidth bw, unsymmetric values, stable-ish scaling) and a random column-selection-like B2 with
    random gains: exercises non-tridiagonal index sets, several nonzeros per B2 column and non-unit actuator gains."""
    rng = np.random.default_rng(seed)
    diags, offs = [], []
    for k in range(-bw, bw + 1):
        v = rng.uniform(-0.3, 0.3, Nx - abs(k)) if k else rng.uniform(0.6, 1.1, Nx)
        diags.append(v); offs.append(k)
    A = sp.diags(diags, offs, format="csc", dtype=np.float64)
    cols = np.sort(rng.choice(Nx, size=max(Nx // act_every, 2), replace=False))
    B2 = sp.lil_matrix((Nx, cols.size))
    for l, c in enumerate(cols):
        B2[c, l] = rng.uniform(0.5, 1.5)
        if c + 1 < Nx and rng.random() < 0.5:
            B2[c + 1, l] = rng.uniform(-0.5, 0.5)                 # an actuator touching two states
    return A, sp.identity(Nx, format="csc"), sp.csc_matrix(B2)


def check_random_banded(api, seeds=(0, 1, 2), Nx=48, bw=2, d=4, T=12, cols=(0, 5, 23, 24, 47)):
    """Library vs oracle on random banded plants — well-conditioned, ill-conditioned and infeasible columns mixed.
    Contract checked: (1) a column the library does NOT flag SLS_STATUS_RESIDUAL matches the oracle to PHI_RTOL;
    (2) every column the oracle solves (consistent: oracle residual < 1e-10, cond(G) < 1e6) is NOT flagged and meets
    PHI_RTOL — the refinement sweeps of kernel 2 lift the eps cond(G)^2 limit of the normal-equations recursion to
    ~eps cond(G); (3) columns whose localisation is infeasible (oracle residual > 1e-6) are flagged.
    Returns (worst unflagged error, #unflagged, #flagged)."""
    worst, nclean, nflag = 0.0, 0, 0
    for seed in seeds:
        A, B1, B2 = random_banded_plant(Nx, bw, seed)
        S = O.make_supports(A, B2, d, T, 1.5)
        Ox, Ou, infos = O.SLS_H2(O.Plant(A, B1, B2), S, I=list(cols), return_info=True)
        Px, Pu, info = slc_b200.SLS_H2(slc_b200.Plant(A, B1, B2), S, I=[[int(c)] for c in cols], return_info=True, api=api)
        for k, c in enumerate(cols):
            err = 0.0
            for G, H in zip(Px + Pu, Ox + Ou):
                assert np.array_equal(G.indptr, H.indptr) and np.array_equal(G.indices, H.indices)
                a, b = G[:, c].toarray().ravel(), H[:, c].toarray().ravel()
                nb = np.linalg.norm(b)
                if nb > 0:
                    err = max(err, np.linalg.norm(a - b) / nb)
            flagged = bool(info["status"][k] & _capi.STATUS_RESIDUAL)
            ores, cond = infos[c]["residual"], infos[c]["smax"] / max(infos[c]["smin"], 1e-300)
            if not flagged:
                assert ores < 1e-9 and err < PHI_RTOL, (seed, c, err, ores, cond, info["resid"][k])
                worst = max(worst, err); nclean += 1
            else:
                nflag += 1
                assert not (ores < 1e-10 and cond < 1e6), (seed, c, err, info["resid"][k], ores, cond)
            if ores > 1e-6:
                assert flagged, (seed, c, ores, info["resid"][k])
    return worst, nclean, nflag
