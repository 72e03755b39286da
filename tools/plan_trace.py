"""Host-side timing of one plan (kernel 1) for a shard with an explicit column list and for the all-columns plan:
    SLS_PLAN_TRACE=1 prints  to-sync1 / sync1 / host->fill / fill..sync2 / sync2 / tail  per sls_plan call (ms).
Run on the GPU box:  python tools/plan_trace.py"""
import os, sys
os.environ["SLS_PLAN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, slc_b200
syn = slc_b200.Synthesizer()
sh = bench.ChainShard(syn, 200000, 0, 100000, with_e2e=False)
for i in range(4):
    ms = sh.device_step()
    print("step", ms, {k: syn.timing()[k] for k in ("gather_ms", "gather_gpu_ms", "solve_ms")}, flush=True)
sh2 = bench.ChainShard(syn, 100000, 0, 100000, with_e2e=False)
for i in range(3):
    ms = sh2.device_step()
    print("full step", ms, {k: syn.timing()[k] for k in ("gather_ms", "gather_gpu_ms", "solve_ms")}, flush=True)
