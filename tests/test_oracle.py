"""The CPU oracle against (i) the reference's only hot-path known-answer test, (ii) reference-independent
invariants, (iii) the committed golden vectors.  No GPU, no CUDA library."""
import numpy as np
import scipy.sparse as sp

from conftest import assert_phi_matches, golden_taps, load_golden
from oracle import sls_oracle as O


def readme():
    A, B1, B2 = O.chain_plant(59)
    return O.Plant(A, B1, B2), A, B1, B2


def test_readme_plant_matches_reference_script():
    # README.md:43-47: Nx,Nu = 59,20; B2 columns (1:2) .+ 6(0:9)'
    P, A, B1, B2 = readme()
    assert (P.Nx, P.Nu, P.Nw, P.Nz) == (59, 20, 59, 79)
    act = [r for n in range(10) for r in (6 * n, 6 * n + 1)]
    assert np.array_equal(B2.indices, act) and np.all(B2.data == 1.0)
    assert A[3, 4] == 0.2 and A[4, 3] == -0.2 and A[5, 5] == 1.0


def test_reduction_known_answer():
    """/root/reference/test/reduction_test.jl:10-24 — the reference's only fixture on this path."""
    P, A, B1, B2 = readme()
    pA = O._pat(A)
    Sx = [sp.csc_matrix(O._pat_power(pA, 9) != 0)]
    Su = [sp.csc_matrix((O._pat(sp.csc_matrix(B2).T) @ O._pat_power(pA, 9)) != 0)]
    cj = np.arange(20)
    (At, B1t, B2t), It, iix, sx, su = O.sparsity_dim_reduction(P, cj, [Sx, Su])
    assert np.array_equal(sx, np.arange(30))                       # sₓ = 1:30
    assert np.array_equal(su, np.arange(10))                       # sᵤ = 1:10
    assert np.array_equal(iix, np.r_[np.ones(20, bool), np.zeros(10, bool)])
    assert np.array_equal(It, np.eye(30)[:, :20])                  # Ĩ = I(30)[:,1:20]
    assert np.array_equal(At, A[:30, :30].toarray())               # P̃ == Plant(A[sₓ,sₓ], B₁[sₓ,cⱼ], B₂[sₓ,sᵤ])
    assert np.array_equal(B1t, B1[:30, :20].toarray())
    assert np.array_equal(B2t, B2[:30, :10].toarray())


def test_supports_readme_counts():
    P, A, B1, B2 = readme()
    Sx, Su = O.make_supports(A, B2, 9, 29, 1.5)
    assert sum(M.nnz for M in Sx) == 26413 and sum(M.nnz for M in Su) == 9616      # SURVEY §8 header
    assert Sx[0].nnz == 59 and (Sx[0] != sp.identity(59, dtype=bool, format="csc")).nnz == 0   # ^0 = identity


def test_invariants_and_survey_numbers():
    P, A, B1, B2 = readme()
    S = O.make_supports(A, B2, 9, 29, 1.5)
    Phix, Phiu, infos = O.SLS_H2(P, S, return_info=True)
    # achievability on the assembled global matrices
    assert O.achievability_residual(P, Phix, Phiu) < 1e-13
    # support containment
    for t in range(29):
        assert (Phix[t] != 0).astype(int).multiply(1 - S[0][t].astype(int).toarray()).sum() == 0
        assert (Phiu[t] != 0).astype(int).multiply(1 - S[1][t].astype(int).toarray()).sum() == 0
    # BASELINE.md §4 numbers
    cost = sum(i["cost"] for i in infos.values())
    assert abs(cost - 893.3262819769567) < 1e-9
    for j, c in ((0, 1.739859170172), (2, 18.761369694391), (29, 18.830871965169), (33, 23.545630212737), (58, 22.362423145148)):
        assert abs(infos[j]["cost"] - c) < 1e-10
    assert (infos[0]["rows"], infos[0]["rank"]) == (290, 288)
    assert (infos[29]["rows"], infos[29]["rank"]) == (550, 547)
    assert (infos[29]["n"], infos[29]["m"]) == (21, 8)
    # closed-loop identity (README.md:62-71): x(50+k) = Φx[k][:,30]
    w = np.zeros((59, 250)); w[29, 49] = 1.0
    x, u, beta = O.simulate_closed_loop(P, Phix, Phiu, w, 250)
    for k in range(1, 30):
        assert np.abs(x[:, 49 + k] - Phix[k - 1][:, 29].toarray().ravel()).max() < 1e-13
    assert np.abs(x[:, 49 + 30:]).max() < 1e-13


def test_golden_vectors_reproduce():
    g = load_golden("readme_chain_59")
    P, A, B1, B2 = readme()
    S = O.make_supports(A, B2, int(g["d"]), int(g["T"]), float(g["alpha"]))
    Phix, Phiu = O.SLS_H2(P, S)
    assert_phi_matches(Phix, golden_taps(g, "x"), "Phix", rtol=1e-12)
    assert_phi_matches(Phiu, golden_taps(g, "u"), "Phiu", rtol=1e-12)


def test_grid_infeasible_actuation_is_detected():
    """checkerboard actuation is infeasible at unactuated nodes (bipartite invariant)."""
    A, B1, B2 = O.grid_plant(9, 9, actuation="checker")
    P = O.Plant(A, B1, B2)
    S = O.make_supports(A, B2, 3, 10, 1.5)
    _, _, infos = O.SLS_H2(P, S, I=[40, 41], return_info=True)
    assert infos[40]["residual"] < 1e-12 and infos[41]["residual"] > 1e-3
