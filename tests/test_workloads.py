"""Closed-form workload generators == the README's matrix-power construction (oracle.make_supports)."""
import numpy as np

import slc_b200
from slc_b200 import workloads as W
from oracle import sls_oracle as O


def same(J, M):
    M = M.tocsc(); M.sort_indices()
    return (np.array_equal(J.colptr - 1, M.indptr) and np.array_equal(J.rowval - 1, M.indices)
            and J.nzval.all() and (J.m, J.n) == M.shape)


def test_chain_generators_match_matrix_powers():
    for Nx in (59, 64, 200):
        A, B1, B2 = W.chain_plant(Nx)
        Ao, B1o, B2o = O.chain_plant(Nx)
        assert (A != Ao).nnz == 0 and (B2 != B2o).nnz == 0
        Sx, Su = W.chain_supports(Nx, 9, 29, 1.5)
        Ox, Ou = O.make_supports(Ao, B2o, 9, 29, 1.5)
        for t in range(29):
            assert same(Sx[t], Ox[t]) and same(Su[t], Ou[t]), (Nx, t)


def test_chain_slab_is_a_column_slice():
    Sx, Su = W.chain_supports(200, 9, 29, 1.5)
    sx, su = W.chain_supports(200, 9, 29, 1.5, 50, 120)
    for F, P in ((Sx[-1], sx[-1]), (Su[7], su[7])):
        f, p = F.to_scipy(), P.to_scipy()
        assert (f[:, 50:120] != p[:, 50:120]).nnz == 0 and p[:, :50].nnz == 0 and p[:, 120:].nnz == 0


def test_grid_generators_match_matrix_powers():
    A, B1, B2 = W.grid_plant(7, 9)
    Ao, _, B2o = O.grid_plant(7, 9, actuation="full")
    assert (A != Ao).nnz == 0 and (B2 != B2o).nnz == 0
    Sx, Su = W.grid_supports(7, 9, 3, 8, 1.5)
    Ox, Ou = O.make_supports(Ao, B2o, 3, 8, 1.5)
    for t in range(8):
        assert same(Sx[t], Ox[t]) and same(Su[t], Ou[t]), t


def test_chain_supports_window_form_is_a_column_slice():
    """window=True returns exactly the columns [lo, hi) of the full matrices (rows still numbered in the full plant):
    what a rank passes to sls_synthesize_window."""
    Nx, d, T, alpha, lo, hi = 60, 4, 9, 1.5, 17, 41
    Fx, Fu = W.chain_supports(Nx, d, T, alpha)
    Wx, Wu = W.chain_supports(Nx, d, T, alpha, lo, hi, window=True)
    for F, V in zip(Fx + Fu, Wx + Wu):
        assert V.n == hi - lo and V.m == F.m
        a, b = int(F.colptr[lo] - 1), int(F.colptr[hi] - 1)
        assert np.array_equal(V.colptr, F.colptr[lo:hi + 1] - a)
        assert np.array_equal(V.rowval, F.rowval[a:b]) and V.nzval.all()


def test_sub_plant_of_a_chain_is_a_chain():
    """bench.ChainShard relies on it: with the window start a multiple of 6, A[pw,pw] / B2[pw,uw] of the chain are the
    chain plant of |pw| states (same values, same actuator pattern and numbering up to the offsets)."""
    Nx, p0, p1 = 120, 36, 97
    A, _, B2 = W.chain_plant(Nx)
    a, _, b2 = W.chain_plant(p1 - p0)
    u0 = 2 * (p0 // 6)
    assert abs(A[p0:p1, p0:p1] - a).nnz == 0
    sub = B2[p0:p1, u0:u0 + b2.shape[1]]
    assert abs(sub - b2).nnz == 0
    assert B2[p0:p1, :u0].nnz == 0 and B2[p0:p1, u0 + b2.shape[1]:].nnz == 0
