"""Parity tests proper: the CUDA library (through the C ABI) against the CPU oracle, the committed golden
vectors, and — at BASELINE.json's full sizes — size-independent properties.  Run on the B200 box:
    python -m pytest tests -m gpu
Bar: index sets / patterns bit-exact; Φ within 1e-9 relative Frobenius error per tap (north_star)."""
import numpy as np
import pytest
import scipy.sparse as sp

import parity_cases as pc
import slc_b200
from slc_b200 import _capi
from oracle import sls_oracle as O
from conftest import PHI_FLOOR

pytestmark = pytest.mark.gpu


def test_cuda_library_is_the_one_loaded(cuda_api):
    assert cuda_api.path.endswith("systemlevelcontrol.jl_b200/lib/libslsb200.so")
    assert cuda_api.sls_device_count() >= 1


def test_reduction_fixture(cuda_api):
    pc.check_reduction_fixture(cuda_api)


def test_reduction_output_feedback_branch(cuda_api):
    pc.check_reduction_of(cuda_api)


def test_config1_readme_chain_vs_golden(cuda_api):
    worst, info = pc.check_golden(cuda_api, "readme_chain_59", O.chain_plant(59))
    assert worst < 1e-10
    # rank-deficient but consistent constraints are the normal case
    assert (info["status"] & _capi.STATUS_RANK_SKIPPED).all()


def test_local_blocks_bit_exact(cuda_api):
    A, B1, B2, S = pc.chain(59)
    pc.check_local_blocks(cuda_api, A, B1, B2, S, list(range(0, 59, 4)) + [58])


def test_jitter_chain_vs_golden(cuda_api):
    worst, _ = pc.check_golden(cuda_api, "jitter_chain_300_subset", O.chain_plant(300, jitter=0.1, seed=0))
    assert worst < 1e-10


def test_grid_9x9_vs_golden(cuda_api):
    worst, _ = pc.check_golden(cuda_api, "grid_9x9_full_d2_T8", O.grid_plant(9, 9, actuation="full"), floor=PHI_FLOOR)
    assert worst < 1e-10


def test_grid_13x13_wide_solver_vs_golden(cuda_api):
    """n up to 61, m up to 85: the 64-bit-mask / N=64 size classes of kernel 2 (BASELINE config #4 shapes)."""
    worst, info = pc.check_golden(cuda_api, "grid_13x13_full_d4_T20_subset", O.grid_plant(13, 13, actuation="full"), floor=PHI_FLOOR)
    assert info["n"].max() == 61 and info["m"].max() == 85
    assert worst < 1e-9


def test_config4_grid_properties(cuda_api):
    """BASELINE config #4 at 60 x 60 (same d=4, T=20 local problems as 300 x 300): invariants on all columns and
    oracle parity on a few."""
    from slc_b200 import workloads as W
    nr = nc = 60
    A, B1, B2 = W.grid_plant(nr, nc)
    Sx, Su = W.grid_supports(nr, nc, 4, 20, 1.5)
    S = [[M.to_scipy().astype(bool) for M in Sx], [M.to_scipy().astype(bool) for M in Su]]
    P = slc_b200.Plant(A, B1, B2)
    Phix, Phiu, info = slc_b200.SLS_H2(P, S, return_info=True, api=cuda_api)
    assert not (info["status"] & _capi.STATUS_RESIDUAL).any() and info["resid"].max() < 1e-12
    Nx = nr * nc
    I = sp.identity(Nx, format="csc")
    r = abs(Phix[0] - I).max()
    for t in range(20):
        nxt = Phix[t + 1] if t + 1 < 20 else sp.csc_matrix((Nx, Nx))
        R = nxt - A @ Phix[t] - B2 @ Phiu[t]
        r = max(r, abs(R).max() if R.nnz else 0.0)
        assert np.array_equal(Phix[t].indices, S[0][t].indices) and np.array_equal(Phiu[t].indices, S[1][t].indices)
    assert r < 1e-12
    cols = [0, 59, 1830, 1831, 3599]
    worst, _, _ = pc.check_against_oracle(cuda_api, A, B1, B2, S, cols, floor=PHI_FLOOR)
    assert worst < 1e-9


def test_config4_grid_300x300_full_size(cuda_api):
    """BASELINE config #4 at its stated size (Nx = 90 000, d = 4, T = 20, B2 = I): invariants on all 90 000 columns
    (no residual flag, pattern == support pattern bit-exact, global achievability, translation invariance of interior
    columns bit for bit) and oracle parity on 6 columns (corner, edge, near-edge, interior)."""
    from slc_b200 import workloads as W
    nr = nc = 300
    Nx = nr * nc
    A, B1, B2 = W.grid_plant(nr, nc)
    Sx, Su = W.grid_supports(nr, nc, 4, 20, 1.5)
    syn = slc_b200.Synthesizer(api=cuda_api)
    try:
        syn.set_plant(A, B2); syn.set_supports(Sx, Su)
        _, _, n, m = syn.plan(None)
        out = syn.solve()
        assert n.max() == 61 and m.max() == 85
        assert not (out["status"][:Nx] & _capi.STATUS_RESIDUAL).any() and out["resid"][:Nx].max() < 1e-12
        Phix, Phiu = syn.outputs_to_scipy(out)
    finally:
        syn.close()
    for t in range(20):
        assert np.array_equal(out["x"]["colptr"][t], Sx[t].colptr) and np.array_equal(out["x"]["rowval"][t], Sx[t].rowval)
        assert np.array_equal(out["u"]["colptr"][t], Su[t].colptr) and np.array_equal(out["u"]["rowval"][t], Su[t].rowval)
    I = sp.identity(Nx, format="csc")
    r = abs(Phix[0] - I).max()
    for t in range(20):
        nxt = Phix[t + 1] if t + 1 < 20 else sp.csc_matrix((Nx, Nx))
        R = nxt - A @ Phix[t] - B2 @ Phiu[t]
        r = max(r, abs(R).max() if R.nnz else 0.0)
    assert r < 1e-12
    # interior columns are translates of each other: identical values, bit for bit
    j0 = 150 * nc + 150
    for t in (0, 1, 3, 19):
        M = Phix[t]
        a = M.data[M.indptr[j0]:M.indptr[j0 + 1]]
        for j in (j0 + 1, j0 + nc, 40 * nc + 77, 250 * nc + 13):
            assert np.array_equal(a, M.data[M.indptr[j]:M.indptr[j + 1]])
    # oracle on a few columns of the full-size problem (the oracle's gather is Θ(1) per column once the product is hoisted)
    S = [[M.to_scipy().astype(bool) for M in Sx], [M.to_scipy().astype(bool) for M in Su]]
    cols = [0, nc - 1, 3 * nc + 2, j0, j0 + 7 * nc + 5, Nx - 1]
    Ox, Ou = O.SLS_H2(O.Plant(A, B1, B2), S, I=cols)
    for t in range(20):
        for G, H in ((Phix[t], Ox[t]), (Phiu[t], Ou[t])):
            nmax = max(np.linalg.norm(M_[:, cols].toarray()) for M_ in (Ox if G is Phix[t] else Ou))
            a = G[:, cols].toarray(); b = H[:, cols].toarray()
            assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b) + PHI_FLOOR * nmax


def test_error_paths(cuda_api):
    pc.check_error_paths(cuda_api)


def test_gather_fast_and_general_paths(cuda_api):
    """kernel 1: run fast paths vs sorted-union / per-entry paths mixed in one plan, bit-exact vs the oracle's gather."""
    assert pc.check_gather_paths(cuda_api) < 1e-9


def test_large_local_problems_n_above_64(cuda_api):
    """64 < n <= 128: the one-CTA-per-column SIMT kernel (sls_solve_huge.cu), alone and beside the DMMA kernels."""
    assert pc.check_large_local_problems(cuda_api) < 1e-10


def test_masks_and_flags(cuda_api):
    pc.check_masks_and_flags(cuda_api)
    pc.check_unsorted_plan(cuda_api)


def test_config2_chain_5000(cuda_api):
    """BASELINE config #2: oracle parity on a column sample + invariants on all 5 000 columns."""
    A, B1, B2, S = pc.chain(5000)
    cols = [0, 1, 2, 5, 9, 10, 11, 2497, 2498, 2499, 2500, 2501, 2502, 4988, 4990, 4995, 4998, 4999]
    worst, info, _ = pc.check_against_oracle(cuda_api, A, B1, B2, S, cols)
    assert worst < 1e-10
    P = slc_b200.Plant(A, B1, B2)
    Phix, Phiu, info = slc_b200.SLS_H2(P, S, return_info=True, api=cuda_api)
    check_global_properties(A, B2, S, Phix, Phiu, info)


def check_global_properties(A, B2, S, Phix, Phiu, info, period=6, margin=40):
    Nx = A.shape[0]
    T = len(Phix)
    assert not (info["status"] & _capi.STATUS_RESIDUAL).any()
    assert info["resid"].max() < 1e-12
    # stored pattern == mask pattern == 𝓢[t] (nested supports) — bit-exact
    for t in range(T):
        assert np.array_equal(Phix[t].indptr, S[0][t].indptr) and np.array_equal(Phix[t].indices, S[0][t].indices)
        assert np.array_equal(Phiu[t].indptr, S[1][t].indptr) and np.array_equal(Phiu[t].indices, S[1][t].indices)
    # achievability on the assembled global matrices (synthesis.jl:53-55)
    I = sp.identity(Nx, format="csc")
    r = abs(Phix[0] - I).max()
    for t in range(T):
        nxt = Phix[t + 1] if t + 1 < T else sp.csc_matrix((Nx, Nx))
        R = nxt - A @ Phix[t] - B2 @ Phiu[t]
        r = max(r, abs(R).max() if R.nnz else 0.0)
    assert r < 1e-12
    # translation invariance of the chain: interior columns j and j+6 hold bit-identical values
    for t in (0, 1, 5, T - 1):
        M = Phix[t]
        cnt = np.diff(M.indptr)
        j0 = np.arange(margin, Nx - margin - period)
        same = cnt[j0] == cnt[j0 + period]
        assert same.all()
        for j in j0[:: max(1, j0.size // 500)]:
            a = M.data[M.indptr[j]:M.indptr[j + 1]]
            b = M.data[M.indptr[j + period]:M.indptr[j + period + 1]]
            assert np.array_equal(a, b)


def test_config3_chain_100k_properties(cuda_api):
    """BASELINE config #3 (headline size): properties that do not need the oracle at full size, plus oracle
    parity on a handful of columns, plus the README closed-loop identity on the assembled Φ."""
    Nx = 100_000
    A, B1, B2, S = pc.chain(Nx)
    P = slc_b200.Plant(A, B1, B2)
    Phix, Phiu, info = slc_b200.SLS_H2(P, S, return_info=True, api=cuda_api)
    check_global_properties(A, B2, S, Phix, Phiu, info)
    # oracle on a few columns of the full-size problem
    Po = O.Plant(A, B1, B2)
    cols = [0, 7, 49_999, 50_000, 50_003, 99_999]
    Ox, Ou = O.SLS_H2(Po, S, I=cols)
    for t in range(len(Phix)):
        for G, H in ((Phix[t], Ox[t]), (Phiu[t], Ou[t])):
            a = G[:, cols].toarray(); b = H[:, cols].toarray()
            nb = np.linalg.norm(b)
            assert np.linalg.norm(a - b) <= 1e-9 * max(nb, 1e-300)
    # closed-loop identity (README.md:62-71) around one impulse, simulated on the host with sparse mat-vecs
    j, t0, steps = 50_003, 5, 5 + 29 + 6
    T = len(Phix)
    x = np.zeros((Nx, steps + 1)); beta = np.zeros((Nx, steps + 1))
    for t in range(1, steps):
        what = x - beta
        b = np.zeros(Nx); uu = np.zeros(B2.shape[1])
        for tau in range(1, min(t, T - 1) + 1):
            b += Phix[tau] @ what[:, t - tau]
        for tau in range(1, min(t, T) + 1):
            uu += Phiu[tau - 1] @ what[:, t - tau]
        beta[:, t] = b
        w = np.zeros(Nx)
        if t == t0:
            w[j] = 1.0
        x[:, t] = A @ x[:, t - 1] + w + B2 @ uu
    for k in range(1, T + 1):
        assert np.abs(x[:, t0 + k - 1] - Phix[k - 1][:, j].toarray().ravel()).max() < 1e-12
    assert np.abs(x[:, t0 + T:steps]).max() < 1e-12


def test_sharded_equals_single_bit_for_bit(cuda_api):
    """Column shards are independent: any partition of the columns gives bit-identical Φ (multi-GPU contract)."""
    A, B1, B2, S = pc.chain(600)
    P = slc_b200.Plant(A, B1, B2)
    Fx, Fu = slc_b200.SLS_H2(P, S, api=cuda_api)
    cuts = [0, 211, 400, 600]
    for a, b in zip(cuts[:-1], cuts[1:]):
        Px, Pu = slc_b200.SLS_H2(P, S, I=[[j] for j in range(a, b)], api=cuda_api)
        for t in range(len(Fx)):
            for F, Q in ((Fx[t], Px[t]), (Fu[t], Pu[t])):
                assert np.array_equal(F[:, a:b].indices, Q[:, a:b].indices)
                assert np.array_equal(F[:, a:b].data, Q[:, a:b].data)
                assert Q[:, :a].nnz == 0 and Q[:, b:].nnz == 0


def test_closed_loop_simulation_vs_oracle(cuda_api):
    """README.md:62-71 through sls_simulate: README scenario (j=30, t=50, 250 steps) + edge/late scenarios."""
    worst, tm = pc.check_simulation(cuda_api, 59, [(29, 50), (0, 1), (58, 200), (30, 249)], steps=250)
    rng = np.random.default_rng(0)
    scen = [(int(j), int(t)) for j, t in zip(rng.integers(0, 400, 24), rng.integers(1, 100, 24))]
    worst, tm = pc.check_simulation(cuda_api, 400, scen, steps=130)
    assert tm["sim_window"] == 39 and tm["sim_uwindow"] <= 20


def test_config5_simulation_100k_4096_scenarios(cuda_api):
    """BASELINE config #5: Nx = 100 000 controller, 4 096 random impulse scenarios, 250 steps (seed 0).  Checked
    through the closed-loop identity on norms: ||x(ts+k)|| = ||Phi_x[k][:,j]||, 0 after the FIR horizon."""
    Nx, ns, steps = 100_000, 4096, 250
    from slc_b200 import workloads as W
    A, B1, B2 = W.chain_plant(Nx)
    Sx, Su = W.chain_supports(Nx, 9, 29, 1.5)
    syn = slc_b200.Synthesizer(api=cuda_api)
    try:
        syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan(None)
        out = syn.solve()
        rng = np.random.default_rng(0)
        js = rng.integers(0, Nx, ns); ts = rng.integers(1, steps - 29, ns)
        sim = syn.simulate(js, ts, steps)
        tm = syn.timing()
    finally:
        syn.close()
    T = 29
    colnorm = np.zeros((T, Nx))
    for t in range(T):
        cp = out["x"]["colptr"][t] - 1
        sq = np.add.reduceat(np.r_[out["x"]["nzval"][t] ** 2, 0.0], cp[:-1])
        sq[cp[:-1] == cp[1:]] = 0.0
        colnorm[t] = np.sqrt(sq)
    for s in range(0, ns, 7):
        j, t0 = int(js[s]), int(ts[s])
        got = sim["xnorm"][s]
        assert np.abs(got[:t0]).max() == 0.0
        assert np.abs(got[t0:t0 + T] - colnorm[:, j]).max() < 1e-12
        assert np.abs(got[t0 + T:]).max() < 1e-12
    print("config5 sim_ms", tm["sim_ms"])


def test_device_generated_supports(cuda_api):
    """README.md:52-54 on the device, against the matrix-power construction."""
    pc.check_device_supports(cuda_api, O.chain_plant(59), 9, 29, 1.5)
    pc.check_device_supports(cuda_api, O.chain_plant(300, jitter=0.1), 5, 12, 2.0)
    pc.check_device_supports(cuda_api, O.grid_plant(11, 9, actuation="full"), 3, 10, 1.5)
    A, B1, B2 = O.grid_plant(8, 8, actuation="checker")          # B2 is a strict selection: Nu < Nx
    pc.check_device_supports(cuda_api, (A, B1, B2), 2, 7, 1.0)


def test_pipelined_one_shot_equals_three_call_small(cuda_api):
    A, B1, B2, S = pc.chain(59)
    pc.check_pipelined(cuda_api, (A, B1, B2), S, chunks=[1, 3, 7])
    pc.check_pipelined(cuda_api, (A, B1, B2), S, chunks=[4], pass_nzval=False)


def test_pipelined_rowval_from_host_or_device(cuda_api):
    """sls_synthesize: rowval outputs filled from the caller's S[t].rowval (pattern proven equal) or from kernel 3's output
    (stored-false entries, rows outside s_x) — bit for bit the three-call result either way; flags bit 3 = device only."""
    pc.check_pipelined_rowval_sources(cuda_api, Nx=59, d=9, T=29)
    pc.check_pipelined_rowval_sources(cuda_api)


def test_pipelined_one_shot_config3_100k_bit_for_bit(cuda_api):
    """BASELINE config #3 through sls_synthesize (default 8 column chunks, pinned buffers, copies overlapped with the
    kernels) == the three-call path, bit for bit; plus a 2-shard split of the columns."""
    from slc_b200 import workloads as W
    Nx = 100_000
    A, B1, B2 = W.chain_plant(Nx)
    Sx, Su = W.chain_supports(Nx, 9, 29, 1.5)
    syn = slc_b200.Synthesizer(api=cuda_api)
    try:
        syn.set_plant(A, B2)
        syn.set_supports(Sx, Su)
        syn.plan(None)
        ref = syn.solve()
        out = syn.synthesize(Sx, Su, pinned=True)
        for fam in ("x", "u"):
            for t in range(29):
                for part in ("colptr", "rowval", "nzval"):
                    assert np.array_equal(out[fam][part][t], ref[fam][part][t]), (fam, t, part)
        assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["resid"], ref["resid"])
        assert out["resid"].max() < 1e-12
        # two shards: the entries of every tap are the concatenation of the shards' entries
        h0 = syn.synthesize(Sx, Su, col_range=(0, Nx // 2))
        h1 = syn.synthesize(Sx, Su, col_range=(Nx // 2, Nx))
        for fam in ("x", "u"):
            for t in (0, 7, 28):
                assert np.array_equal(np.concatenate([h0[fam]["nzval"][t], h1[fam]["nzval"][t]]), ref[fam]["nzval"][t])
                assert np.array_equal(np.concatenate([h0[fam]["rowval"][t], h1[fam]["rowval"][t]]), ref[fam]["rowval"][t])
    finally:
        syn.close()


def test_window_shard_form(cuda_api):
    """sls_synthesize_window: support window in, column slab out (the per-GPU shape; no array grows with Nx)."""
    A, B1, B2, S = pc.chain(600)
    pc.check_window(cuda_api, (A, B1, B2), S, ranges=[(0, 200), (199, 417), (417, 600)], chunks=(1, 4), sub_reach=11)


def test_weighted_cost_and_column_groups_vs_general_oracle(cuda_api):
    """6-argument Plant with diagonal cost weights, diagonal B1, q = 2 group."""
    assert pc.check_weighted_cost(cuda_api, cols=(3, 14, 39, 0, 20, 27)) < 1e-10


def test_general_cost_and_coupled_groups_vs_general_oracle(cuda_api):
    """SURVEY §8f-3 in full: non-diagonal [C1 D12], D11 != 0, non-diagonal B1, coupled q = 2 / q = 3 groups
    (synthesis.jl:42-43,50,76-77; reduction.jl:14-23) on kernel 2g, against SLS_H2_general."""
    assert pc.check_coupled_groups_identity_cost(cuda_api)[0] < 1e-10
    assert pc.check_general_cost(cuda_api) < 1e-10
    assert pc.check_general_cost(cuda_api, Nx=60, d=9, T=29, cases=[
        dict(kw=dict(dense_cost=True, b1_band=True, d11=True), I=[[0, 1], [29, 30], [31], [58, 59]]),
        dict(kw=dict(dense_cost=True, b1_band=False, d11=False), I=[[j] for j in range(0, 60, 7)])]) < 1e-10
    pc.check_general_error_paths(cuda_api)


def test_random_banded_plants_vs_oracle(cuda_api):
    """Pentadiagonal random plants with two-state actuators and random gains (not a chain, not a grid): the accuracy /
    flagging contract on well-conditioned, ill-conditioned and infeasible columns (parity_cases.check_random_banded)."""
    worst, nclean, nflag = pc.check_random_banded(cuda_api, seeds=tuple(range(12)))
    assert nclean >= 30 and nflag >= 8, (nclean, nflag)       # flagged = infeasible or cond(G) >= 1e6 only
    assert worst < 1e-9, worst
    # wider band, n up to 37: the refinement sweeps of the one-CTA-per-column kernel (sls_solve_wide.cu)
    worst, nclean, nflag = pc.check_random_banded(cuda_api, seeds=(0, 1, 2, 3), Nx=80, bw=3, d=5, T=9, cols=(0, 11, 40, 41, 79))
    assert nclean >= 8 and worst < 1e-9, (worst, nclean, nflag)
