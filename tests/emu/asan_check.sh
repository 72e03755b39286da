#!/bin/sh
# TEST-ONLY: AddressSanitizer pass over the kernel sources on the SIMT emulator (compute-sanitizer is closed on the
# GPU pool).  Builds csrc/*.cu with g++ -fsanitize=address -DSLS_EMU and runs config #1-like chain cases (all size
# classes + simulation) and an n = 61 grid column.  Expected output: "chain ok ...", "grid ok ..." and no ASan report.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${TMPDIR:-/tmp}/sls_asan; mkdir -p "$OUT"
for f in sls_api sls_gather sls_solve sls_solve_wide sls_solve_huge sls_solve_general sls_scatter sls_sim sls_supports; do
  g++ -O1 -g -std=c++17 -DSLS_EMU -fPIC -w -fsanitize=address -fno-omit-frame-pointer -x c++ -I "$ROOT/tests/emu" \
      -c "$ROOT/systemlevelcontrol.jl_b200/csrc/$f.cu" -o "$OUT/$f.o" &
done; wait
g++ -shared -fsanitize=address -o "$OUT/libslsb200_asan.so" "$OUT"/*.o
cat > "$OUT/run.py" <<PY
import sys
sys.path.insert(0, "$ROOT")
import slc_b200
from slc_b200 import workloads as W, _capi
api = _capi.CApi("$OUT/libslsb200_asan.so")
A, B1, B2 = W.chain_plant(40); Sx, Su = W.chain_supports(40, 9, 29, 1.5)
syn = slc_b200.Synthesizer(api=api); syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan(None); out = syn.solve()
syn.simulate([20, 0, 39], [5, 3, 20], 40)
print("chain ok", out["resid"][:40].max())
import os, numpy as np
os.environ["SLS_PIPE_CHUNKS"] = "3"
o2 = syn.synthesize(Sx, Su)                                   # pipelined one-shot path (feeder thread, per-chunk plans)
assert all(np.array_equal(o2["x"]["nzval"][t], out["x"]["nzval"][t]) for t in range(29))
o3 = syn.synthesize(Sx, Su, col_range=(10, 31))
syn.set_weights(np.linspace(0.5, 2, 40), np.linspace(1, 2, B2.shape[1]), 40, B2.shape[1]); syn.set_plant(A, B2)
o4 = syn.synthesize(Sx, Su)                                   # cost weights: scaled plant + back-scaling kernel
syn.set_weights(None, None, 40, B2.shape[1]); syn.set_plant(A, B2); syn.make_supports(9, 29, 1.5); syn.plan(None); syn.solve()
print("pipelined / weighted / device supports ok")
syn.set_supports(Sx, Su); syn.plan(None); full = syn.solve()
o5, nx5, nu5 = syn.synthesize_window(Sx, Su, (9, 32), (10, 31))                     # shard form: window in, slab out
o6, _, _ = syn.synthesize_window(Sx, Su, (9, 32), (10, 31), values_only=True)
syn.set_supports(Sx, Su); syn.plan([35, 3, 20, 21, 0, 39, 7]); syn.solve()                                    # unsorted list: kernel 3's non-contiguous stores
print("window / unsorted ok"); syn.close()
import scipy.sparse as sp
sub = slc_b200.Synthesizer(api=api); sub.set_plant(sp.csc_matrix(A)[6:40, 6:40], sp.csc_matrix(B2)[6:40, 2:])   # sub-plant shard with row offsets
WS = W.chain_supports(40, 9, 29, 1.5, 15, 31, window=True)
sub.synthesize_window(WS[0], WS[1], (15 - 6, 31 - 6), (16 - 6, 30 - 6), row_off=(6, 2)); print("sub-plant ok"); sub.close()
os.environ["SLS_SIM_GLOBAL"] = "1"
syn = slc_b200.Synthesizer(api=api); syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan(None); syn.solve_device()
syn.simulate([20, 0, 39], [5, 3, 20], 40); print("sim (L2 scratch mode) ok"); syn.close()
A, B1, B2 = W.grid_plant(13, 13); Sx, Su = W.grid_supports(13, 13, 4, 20, 1.5)
syn = slc_b200.Synthesizer(api=api); syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan([0, 6, 84]); out = syn.solve()
print("grid ok (CTA kernel, n = 21 / 36 / 61)", out["resid"][:3].max(), syn.n); syn.close()
A, B1, B2 = W.grid_plant(15, 15)
for d_ in (5, 6):                                             # n = 85 beside the DMMA kernels; n = 113, m = 141: every class on the n > 64 kernel
    Sx, Su = W.grid_supports(15, 15, d_, 6, 1.5)
    syn = slc_b200.Synthesizer(api=api); syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan([112, 0, 7]); out = syn.solve()
    print("grid ok (n > 64 kernel, d =", d_, ")", out["resid"][:3].max(), syn.n); syn.close()
sys.path.insert(0, "$ROOT/tests")
import parity_cases as pc
print("general cost / groups ok (kernel 2g)", pc.check_general_cost(api, cases=[
    dict(kw=dict(dense_cost=True, b1_band=True, d11=True), I=[[0, 1], [10, 11, 12], [29]])]))
pc.check_general_error_paths(api); print("general error paths ok")
PY
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python "$OUT/run.py"
