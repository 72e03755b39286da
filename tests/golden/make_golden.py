"""Generates tests/golden/*.npz with the CPU oracle (oracle/sls_oracle.py).

The reference itself cannot run in this environment (no Julia; its arithmetic lives in JuMP/Ipopt,
see oracle header), and none of its tests pin Φ values, so these vectors are ORACLE-derived: the exact
minimiser of each column QP by dense SVD.  They freeze the oracle's answers so that a later change to
either side is caught.  The only reference-derived fixture is the reduction known-answer test
(test/reduction_test.jl:10-24), hard-coded in tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import sls_oracle as O  # noqa: E402


def pack(Phix, Phiu):
    d = {}
    for name, Phi in (("x", Phix), ("u", Phiu)):
        d[f"{name}_indptr"] = np.stack([M.indptr.astype(np.int64) for M in Phi])
        d[f"{name}_indices"] = np.concatenate([M.indices.astype(np.int64) for M in Phi])
        d[f"{name}_data"] = np.concatenate([M.data for M in Phi])
    return d


def case(name, A, B1, B2, d, T, alpha, cols=None):
    P = O.Plant(A, B1, B2)
    S = O.make_supports(A, B2, d, T, alpha)
    Phix, Phiu, infos = O.SLS_H2(P, S, I=cols, return_info=True)
    cl = sorted(infos)
    out = pack(Phix, Phiu)
    out.update(cols=np.array(cl, np.int64), d=d, T=T, alpha=alpha,
               cost=np.array([infos[c]["cost"] for c in cl]), rank=np.array([infos[c]["rank"] for c in cl]),
               rows=np.array([infos[c]["rows"] for c in cl]), resid=np.array([infos[c]["residual"] for c in cl]),
               n=np.array([infos[c]["n"] for c in cl]), m=np.array([infos[c]["m"] for c in cl]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "cols", len(cl), "total cost", out["cost"].sum(), "max resid", out["resid"].max())


if __name__ == "__main__":
    A, B1, B2 = O.chain_plant(59)
    case("readme_chain_59", A, B1, B2, 9, 29, 1.5)
    A, B1, B2 = O.chain_plant(300, jitter=0.1, seed=0)
    case("jitter_chain_300_subset", A, B1, B2, 9, 29, 1.5, cols=[0, 1, 2, 3, 7, 12, 100, 101, 102, 103, 104, 105, 150, 290, 297, 298, 299])
    A, B1, B2 = O.grid_plant(9, 9, a=0.1, actuation="full")
    case("grid_9x9_full_d2_T8", A, B1, B2, 2, 8, 1.5)
    A, B1, B2 = O.grid_plant(13, 13, a=0.1, actuation="full")
    case("grid_13x13_full_d4_T20_subset", A, B1, B2, 4, 20, 1.5, cols=[0, 6, 14, 84, 85, 168])
