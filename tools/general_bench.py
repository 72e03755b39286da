"""Times kernel 2g (general cost / column groups, sls_solve_general.cu) on the README chain: `ncols` columns with a banded
non-diagonal [C1 D12], as single columns and as q = 2 groups.  Usage: python tools/general_bench.py [Nx] [ncols]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import slc_b200
from slc_b200 import workloads as W

Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
A, B1, B2 = W.chain_plant(Nx)
Nu = B2.shape[1]
Sx, Su = W.chain_supports(Nx, 9, 29, 1.5)
rng = np.random.default_rng(0)
Qh = sp.diags([rng.uniform(1, 2, Nx), rng.uniform(-0.4, 0.4, Nx - 1), rng.uniform(-0.3, 0.3, Nx - 1)], [0, 1, -1])
Rh = sp.diags([rng.uniform(1, 2, Nu), rng.uniform(-0.3, 0.3, Nu - 1)], [0, 1])
C1 = sp.vstack([Qh, sp.csc_matrix((Nu, Nx))]).tocsc()
D12 = sp.vstack([sp.csc_matrix((Nx, Nu)), Rh]).tocsc()
from slc_b200 import _capi
api = _capi.CApi(os.environ['SLSB200_LIB']) if os.environ.get('SLSB200_LIB') else None
syn = slc_b200.Synthesizer(api=api)
syn.set_plant(A, B2)
syn.set_supports(Sx, Su)
cols = np.arange(Nx // 2 - ncols // 2, Nx // 2 - ncols // 2 + ncols)
for label, cost, grp in (("DMMA kernels, identity cost", False, None), ("kernel 2g, banded cost, q = 1", True, None),
                         ("kernel 2g, banded cost, q = 2", True, np.arange(0, ncols + 1, 2))):
    syn.set_cost(C1, D12) if cost else syn.set_cost(None, None)
    syn.plan(cols, sizes=False)
    syn.set_groups(grp)
    for rep in range(2):
        syn.solve_device()
    tm = syn.timing()
    st, rs = syn.solve_status()
    print(f"{label}: {ncols} columns, solve {tm['solve_ms']:.2f} ms = {ncols / tm['solve_ms'] * 1e3:.0f} columns/s, "
          f"flagged {int((st & ~2 != 0).sum())}, max resid {rs.max():.2e}", flush=True)
syn.close()
