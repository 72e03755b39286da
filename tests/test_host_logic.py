"""Host-side mirror of the reference interface: Plant validation, layout conversion, guards, sharding."""
import numpy as np
import pytest
import scipy.sparse as sp

import slc_b200
from slc_b200 import JuliaCSC, Plant
from oracle import sls_oracle as O


def test_plant_defaults_and_dims():
    # GeneralizedPlant.jl:105-110 — Nz = Nx+Nu, Ny = Nx for the 3-argument constructor
    A, B1, B2 = O.chain_plant(59)
    P = Plant(A, B1, B2)
    assert (P.Nx, P.Nu, P.Nw, P.Nz, P.Ny) == (59, 20, 59, 79, 59)
    assert P.feedback == "StateFeedback" and P._b1_is_identity()


def test_plant_dimension_mismatch_raises():
    # test/types_GeneralizedPlant_test.jl:123-130 (@test_throws ErrorException)
    A = sp.identity(4, format="csc")
    with pytest.raises(ValueError):
        Plant(sp.csc_matrix(np.ones((4, 3))), A, A)
    with pytest.raises(ValueError):
        Plant(A, sp.identity(3, format="csc"), A)
    with pytest.raises(ValueError):
        Plant(A, A, sp.csc_matrix(np.ones((5, 2))))


def test_julia_csc_roundtrip_is_one_based_int64():
    M = sp.random(30, 20, density=0.2, format="csc", random_state=0)
    J = JuliaCSC.from_scipy(M, np.float64)
    assert J.colptr.dtype == np.int64 and J.rowval.dtype == np.int64 and J.colptr[0] == 1
    assert J.rowval.min() >= 1 and J.rowval.max() <= 30 and J.nnz == M.nnz
    assert (J.to_scipy() != M).nnz == 0
    Jb = JuliaCSC.from_scipy(M != 0, np.bool_)
    assert Jb.nzval.dtype == np.uint8 and Jb.nzval.all()


def test_non_state_feedback_returns_none_like_reference():
    # src/synthesis.jl:13 — the `if` falls through and the function returns nothing
    class OF:
        feedback = "OutputFeedback"
    assert slc_b200.SLS_H2(OF(), [[], []]) is None


def test_cost_and_group_routing():
    """Which kernel family SLS_H2 picks: identity / diagonal costs -> DMMA kernels (+ sls_set_weights);
    non-diagonal [C1 D12], D11 != 0, non-diagonal B1 -> kernel 2g (sls_set_cost); q > 1 groups -> kernel 2g unless they
    provably decouple.  Host logic only."""
    from slc_b200.synthesis import _route
    A, B1, B2 = O.chain_plant(12)
    Nx, Nu = 12, B2.shape[1]
    S = O.make_supports(A, B2, 3, 6, 1.5)
    assert _route(Plant(A, B1, B2), S, None) == (None, None, False, None, None)
    B1c = sp.csc_matrix(B1 + sp.diags([0.1 * np.ones(Nx - 1)], [1]))          # R = B1~[ii_x,:] couples group columns
    assert _route(Plant(A, B1c, B2), S, None)[2] is True
    C1 = sp.vstack([sp.identity(Nx) + sp.diags([0.1 * np.ones(Nx - 1)], [1]), sp.csc_matrix((Nu, Nx))])
    D12 = sp.vstack([sp.csc_matrix((Nx, Nu)), sp.identity(Nu)])
    assert _route(Plant(A, B1, B2, C1, 0, D12), S, None)[2] is True
    D11 = sp.csc_matrix(np.ones((Nx + Nu, Nx)))
    assert _route(Plant(A, B1, B2, sp.vstack([sp.identity(Nx), sp.csc_matrix((Nu, Nx))]), D11, D12), S, None)[2] is True
    # diagonal weights stay on the DMMA kernels ...
    wx = np.linspace(0.5, 2, Nx)
    Cw = sp.vstack([sp.diags(wx), sp.csc_matrix((Nu, Nx))])
    r = _route(Plant(A, B1, B2, Cw, 0, D12), S, [[3], [4]])
    assert r[2] is False and np.array_equal(r[0], wx) and r[3] == [3, 4] and r[4] is None
    # ... unless a coupled group needs kernel 2g: the weights then travel inside C1 / D12
    wide = sp.csc_matrix(O._pat_power(O._pat(A), 3) != 0)
    Sw = [[wide, sp.identity(12, format="csc", dtype=bool)], [sp.csc_matrix((Nu, 12), dtype=bool)] * 2]
    r = _route(Plant(A, B1, B2, Cw, 0, D12), Sw, [[0, 1], [5]])
    assert r[0] is None and r[2] is True and r[3] == [0, 1, 5] and r[4].tolist() == [0, 2, 3]
    r = _route(Plant(A, B1, B2), Sw, [[0, 1]])
    assert r[2] is False and r[4].tolist() == [0, 2]                         # identity cost, coupled group: groups only
    with pytest.raises(ValueError):
        _route(Plant(A, B1, B2), S, [[4, 3]])                                 # 3 is first seen among column 4's rows, before 4
    assert _route(Plant(A, B1, B2), S, [[11, 0]])[3] == [11, 0]               # far apart: first occurrences follow the group order
    with pytest.raises(ValueError):
        _route(Plant(A, B1, B2), S, [[3, 4], [4]])                            # a column in two groups


def test_diagonal_cost_classification():
    A, B1, B2 = O.chain_plant(12)
    Nx, Nu = 12, B2.shape[1]
    assert Plant(A, B1, B2).diagonal_cost() == (None, None, None)
    assert Plant(A, 2.0 * B1, B2).diagonal_cost() == (None, None, None)       # diagonal B1: same minimiser
    wx, wu = np.linspace(0.5, 2, Nx), np.linspace(1, 3, Nu)
    C1 = sp.vstack([sp.diags(wx), sp.csc_matrix((Nu, Nx))]); D12 = sp.vstack([sp.csc_matrix((Nx, Nu)), sp.diags(wu)])
    gx, gu, why = Plant(A, B1, B2, C1, 0, D12).diagonal_cost()
    assert why is None and np.array_equal(gx, wx) and np.array_equal(gu, wu)
    P6 = Plant(A, B1, B2, C1, 0, D12)
    assert (P6.Nz, P6.D11.shape) == (Nx + Nu, (Nx + Nu, Nx))
    with pytest.raises(ValueError):
        Plant(A, B1, B2, C1, 0, sp.csc_matrix((Nx + Nu + 1, Nu)))


def test_group_decoupling_check():
    from slc_b200.synthesis import _groups_decouple
    A, B1, B2, S = (*O.chain_plant(30), None)
    S = O.make_supports(A, B2, 4, 10, 1.5)
    P = Plant(A, B1, B2)
    assert _groups_decouple(P, S[0], S[1], [[0, 1, 2], [10, 25]])             # README-style supports are closed
    wide = sp.csc_matrix(O._pat_power(O._pat(A), 3) != 0)                     # a wide tap before a narrow last tap is not
    Sx = [wide, sp.identity(30, format="csc", dtype=bool)]
    Su = [sp.csc_matrix((B2.shape[1], 30), dtype=bool)] * 2
    assert not _groups_decouple(P, Sx, Su, [[3, 4]])
    assert _groups_decouple(P, Sx, Su, [[3], [4]])                            # single-column groups never couple
