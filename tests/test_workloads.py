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
