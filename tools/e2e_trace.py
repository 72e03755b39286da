#!/usr/bin/env python
"""e2e_trace.py — the host drop-in call of bench.py (`sls_set_plant` + `sls_synthesize_window`, pinned host buffers) alone,
with the library's per-chunk pipeline trace (SLS_PIPE_TRACE=1: completion times of copy-in / plan / scatter / solve / copy-out
of every chunk on the device clock).  Chunking is chosen through the environment (SLS_PIPE_CHUNKS, SLS_PIPE_FRACS).

    SLS_PIPE_TRACE=1 python tools/e2e_trace.py [--reps 5] [--nx 100000]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nx", type=int, default=100_000)
    ap.add_argument("--values-only", action="store_true")
    a = ap.parse_args()
    import torch
    import slc_b200
    import bench
    torch.cuda.set_device(0)
    syn = slc_b200.Synthesizer(device=0)
    shard = bench.ChainShard(syn, a.nx, 0, a.nx, with_e2e=True)
    out = shard.alloc_slab(values_only=a.values_only)
    trace = os.environ.pop("SLS_PIPE_TRACE", None)
    for _ in range(2):
        shard.e2e_step(out, values_only=a.values_only)
    torch.cuda.synchronize()
    ms = [shard.e2e_step(out, values_only=a.values_only)[0] for _ in range(a.reps)]
    assert out["resid"].max() < 1e-9
    import time
    sp, sw = [], []
    for _ in range(3):
        t0 = time.perf_counter(); syn.set_plant(shard.Aj, shard.B2j); t1 = time.perf_counter()
        syn.synthesize_window(shard.Sxw, shard.Suw, (shard.w0 - shard.p0, shard.w1 - shard.p0), (shard.c0 - shard.p0, shard.c1 - shard.p0),
                              out=out, values_only=a.values_only, row_off=(shard.p0, shard.u0))
        t2 = time.perf_counter()
        sp.append(1e3 * (t1 - t0)); sw.append(1e3 * (t2 - t1))
    if trace:
        os.environ["SLS_PIPE_TRACE"] = trace
        shard.e2e_step(out, values_only=a.values_only)
    print(f"host wall: sls_set_plant {min(sp):.2f} ms, sls_synthesize_window {min(sw):.2f} ms (incl. the ctypes wrapper)", flush=True)
    print(f"e2e chunks={os.environ.get('SLS_PIPE_CHUNKS', 'default')} fracs={os.environ.get('SLS_PIPE_FRACS', '-')} "
          f"host_copy_threads={os.environ.get('SLS_HOST_COPY_THREADS', 'default')} nt_stores={0 if os.environ.get('SLS_HOST_COPY_NO_NT') else 1} rows_from_device={os.environ.get('SLS_PIPE_ROWS_FROM_DEVICE', '0')} "
          f"values_only={a.values_only}: mean {np.mean(ms):.2f} ms, min {np.min(ms):.2f} ms over {a.reps} calls", flush=True)
    syn.close()


if __name__ == "__main__":
    main()
