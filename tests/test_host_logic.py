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


def test_unsupported_shapes_raise():
    A, B1, B2 = O.chain_plant(12)
    P = Plant(A, 2.0 * B1, B2)
    with pytest.raises(NotImplementedError):
        slc_b200.SLS_H2(P, [[], []])
    P = Plant(A, B1, B2)
    with pytest.raises(NotImplementedError):
        slc_b200.SLS_H2(P, [[sp.identity(12, format="csc")], [sp.csc_matrix((B2.shape[1], 12))]], I=[[0, 1]])
