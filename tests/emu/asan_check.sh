#!/bin/sh
# TEST-ONLY: AddressSanitizer pass over the kernel sources on the SIMT emulator (compute-sanitizer is closed on the
# GPU pool).  Builds csrc/*.cu with g++ -fsanitize=address -DSLS_EMU and runs config #1-like chain cases (all size
# classes + simulation) and an n = 61 grid column.  Expected output: "chain ok ...", "grid ok ..." and no ASan report.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${TMPDIR:-/tmp}/sls_asan; mkdir -p "$OUT"
for f in sls_api sls_gather sls_solve sls_scatter sls_sim sls_supports; do
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
print("pipelined / weighted / device supports ok"); syn.close()
A, B1, B2 = W.grid_plant(13, 13); Sx, Su = W.grid_supports(13, 13, 4, 20, 1.5)
syn = slc_b200.Synthesizer(api=api); syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan([0, 84]); out = syn.solve()
print("grid ok", out["resid"][:2].max(), syn.n); syn.close()
PY
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python "$OUT/run.py"
