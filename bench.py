#!/usr/bin/env python
"""bench.py — localized H2 SLS synthesis throughput (columns/s) on B200, BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [...]                         # CPU arm: restated oracle on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU

Headline workload (config.workload): BASELINE config #3 — tri-diagonal chain, 100 000 disturbance columns per GPU
(chain of N x 100 000 states), Nu = Nx/3, d = 9, T = 29, alpha = 1.5 (README.md:43-54 scaled).  A "step" is one full
pass of the hot path over the rank's column shard: kernel 1 (gather) -> kernel 2 (batched FP64 solve) -> kernel 3
(scatter).  Columns are independent: every rank owns a contiguous column range and holds only the sub-plant and the
support window that range needs (`sls_synthesize_window`); there is no data-path collective ("scaling": "weak").

 value   : columns/s with plant and supports already resident in HBM (device time, CUDA events on the library's stream,
           max over ranks).
 e2e     : the same through the reference-facing call with HOST buffers (pinned): sls_set_plant + sls_synthesize_window
           (H2D of the support window, kernels 1-3, every Φ tap's column slab in Julia CSC layout back in the caller's arrays:
           colptr and values by D2H copies, rowval by D2H or — for taps whose pattern the device proved equal to the
           support's — from the caller's own S[t].rowval by host threads), pipelined.
 roofline: kernel 2 (dominant) against the FP64 DMMA peak measured on this device by sls_fp64_peak
           (MEASURED_PEAKS.json has no FP64 entry); roofline_hbm: kernels 1+3 against the measured copy peak.
 strong  : the same hot path on Nx = 100 000 columns TOTAL, cost-balanced contiguous ranges over the N GPUs
           (sharding.balanced_ranges) — the split north_star names (src/synthesis.jl:16,23-27).
 config4 / config5 (N = 1): BASELINE configs #4 (300x300 grid) and #5 (4 096 x 250 FIR simulation) with their rooflines.
 cpu_baseline: the CPU oracle (a port — the Julia/JuMP/Ipopt reference cannot run here) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX_PER_GPU = 100_000
D, T, ALPHA = 9, 29, 1.5
PLANT_REACH = 12          # sub-plant margin around a rank's column range: every s_x lies within d+1 = 10 states of its column
METRIC = "Phi synthesis throughput, localized H2 state feedback (chain Nx=100k/GPU, d=9, T=29)"
UNIT = "columns/s"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed ncu --set full captures
    (profiles/ncu_traffic.json: value + the summary file it was read from).  None when there is no capture."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = tab.get(key)
        return (float(e["bytes"]), e["source"]) if e else (None, None)
    except Exception:
        return None, None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=NX_PER_GPU, help="columns per GPU (default: the headline 100 000)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="columns in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config4 / config5 objects of the N=1 line")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--workload", default="chain", choices=["chain", "grid"],
                    help="chain = BASELINE config #3 (the headline, default); grid = only the config #4 object (300x300)")
    ap.add_argument("--grid-side", type=int, default=300)
    ap.add_argument("--allgather", action="store_true",
                    help="N>1: also time an NCCL all-gather of the device-resident Phi value slabs (only needed when every "
                         "GPU must hold the whole Phi; not part of value/e2e)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle (oracle/sls_oracle.py), multi-process over the host cores
# ---------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(nx):
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import sls_oracle as O
    A, B1, B2 = O.chain_plant(nx)
    P = O.Plant(A, B1, B2)
    S = O.make_supports(A, B2, D, T, ALPHA)
    _W.update(O=O, P=P, S=S, prods=O.support_products(P, S))


def _cpu_cols(cols):
    O, P, S = _W["O"], _W["P"], _W["S"]
    acc = 0.0
    for j in cols:
        (At, _, B2t), It, iix, sx, su = O.sparsity_dim_reduction(P, [j], S, _prod=_W["prods"])
        Mx, Mu = O.column_masks(S, sx, su, j)
        X, U, info = O.solve_column(At, B2t, It, Mx, Mu)
        acc += info["cost"]
    return acc


class CpuArm:
    """Times the oracle's per-column path (gather + exact QP solve) on interior+edge columns of a chain.
    The oracle is Θ(1) per column once the global support product is hoisted (the Julia reference recomputes
    it per column and is Θ(Nx) per column), so a 2 000-state chain gives the same per-column work
    as the 100 000-state one; the sample plant size is stated in `sample`."""

    def __init__(self, nx_sample=2000):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        self.nx = nx_sample
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(nx_sample,))
        self.pool.map(_cpu_cols, [[10]] * self.cores)        # warm every worker

    def run(self, ncols):
        cols = np.linspace(0, self.nx - 1, ncols).astype(int).tolist()
        chunks = [cols[i::self.cores * 4] for i in range(self.cores * 4)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_cols, [c for c in chunks if c])
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args, rank):
    if rank != 0:
        return
    arm = CpuArm()
    per_step = args.cpu_sample or 32 * arm.cores
    for _ in range(max(args.warmup, 1)):
        arm.run(arm.cores * 2)
    times = [arm.run(per_step) for _ in range(args.steps)]
    arm.close()
    total = float(np.sum(times))
    value = per_step * args.steps / total
    sample = (f"{per_step} columns/step x {args.steps} steps, evenly spaced over a {arm.nx}-state chain (same d,T,alpha; "
              f"per-column work identical to Nx=100k), CPU oracle (dense SVD min-norm per column), {arm.cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"chain Nx={args.nx}/GPU ({args.nx * args.gpus} total), Nu=Nx/3, d={D}, T={T}, alpha={ALPHA} (BASELINE config #3)",
                   "note": "Julia reference (JuMP+Ipopt) not runnable here: no julia in the image; CPU oracle port timed instead"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().strip().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
_pinned = {}
_keep = []


def pin(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    _keep.append(t)
    return t.numpy()


def pinned_csc(M):
    """the same JuliaCSC over page-locked copies of its arrays (taps that share an array in the generator share the copy)"""
    import slc_b200
    out = []
    for arr in (M.colptr, M.rowval, M.nzval):
        if id(arr) not in _pinned:
            _pinned[id(arr)] = (arr, pin(arr))
        out.append(_pinned[id(arr)][1])
    return slc_b200.JuliaCSC(M.m, M.n, *out)


class ChainShard:
    """One rank's share [c0, c1) of a chain of nx_total states: the sub-plant P[p0:p1, p0:p1] it needs (p0 a multiple of 6
    so that the actuator pattern and numbering line up), its supports in the sub-plant's own numbering (device-resident
    path) and as a column window with the global row numbers (host drop-in path, sls_synthesize_window + row offsets)."""

    def __init__(self, syn, nx_total, c0, c1, with_e2e=True):
        import slc_b200
        from slc_b200 import workloads as W
        self.syn, self.nx_total, self.c0, self.c1 = syn, nx_total, c0, c1
        self.p0 = p0 = max(0, ((c0 - PLANT_REACH) // 6) * 6)
        self.p1 = p1 = min(nx_total, c1 + PLANT_REACH)
        self.u0 = 2 * (p0 // 6)
        self.w0, self.w1 = max(0, c0 - 1), min(nx_total, c1 + 1)      # kernel 1 reads 𝓢[T][:,k] for k in nz(A[:,j]): j±1
        A, B1, B2 = W.chain_plant(p1 - p0)
        self.Aj = pinned_csc(slc_b200.JuliaCSC.from_scipy(A, np.float64))
        self.B2j = pinned_csc(slc_b200.JuliaCSC.from_scipy(B2, np.float64))
        Sx, Su = W.chain_supports(p1 - p0, D, T, ALPHA, self.w0 - p0, self.w1 - p0)       # local numbering, full local width
        self.Sx, self.Su = [pinned_csc(M) for M in Sx], [pinned_csc(M) for M in Su]
        self.cols = None if (c0 == p0 and c1 == p1) else np.arange(c0 - p0, c1 - p0)
        self.ncols = c1 - c0
        self.h2d = self.d2h = 0
        if with_e2e:
            Sxw, Suw = W.chain_supports(nx_total, D, T, ALPHA, self.w0, self.w1, window=True)  # global rows, window columns
            self.Sxw, self.Suw = [pinned_csc(M) for M in Sxw], [pinned_csc(M) for M in Suw]
            # taps that share an array in the generator are still uploaded once PER TAP by the ABI: count every tap
            self.h2d = sum(M.colptr.nbytes + M.rowval.nbytes + M.nzval.nbytes for M in self.Sxw + self.Suw + [self.Aj, self.B2j])
        syn.set_plant(self.Aj, self.B2j)
        syn.set_supports(self.Sx, self.Su)

    def device_step(self):
        syn = self.syn
        syn.mark(0)
        syn.plan(self.cols, sizes=False)
        syn.solve_device()
        syn.mark(1)
        return syn.elapsed_ms()

    def alloc_slab(self, pinned=True, values_only=False):
        nnz_x = [int(M.colptr[self.c1 - self.w0] - M.colptr[self.c0 - self.w0]) for M in self.Sxw]
        nnz_u = [int(M.colptr[self.c1 - self.w0] - M.colptr[self.c0 - self.w0]) for M in self.Suw]
        mk = (lambda n, dt: pin(np.zeros(max(n, 1), dt))[:n]) if pinned else (lambda n, dt: np.zeros(max(n, 1), dt)[:n])
        out = {}
        for name, nnz in (("x", nnz_x), ("u", nnz_u)):
            out[name] = dict(colptr=[mk(self.ncols + 1, np.int64) for _ in range(T)],
                             rowval=[mk(0 if values_only else n, np.int64) for n in nnz],
                             nzval=[mk(n, np.float64) for n in nnz])
        out["status"] = np.zeros(self.ncols, np.int32)
        out["resid"] = np.zeros(self.ncols, np.float64)
        self.d2h_full = sum(a.nbytes for name in ("x", "u") for k in ("colptr", "rowval", "nzval") for a in out[name][k]) + \
            out["status"].nbytes + out["resid"].nbytes
        return out

    def e2e_step(self, out, values_only=False, Sx=None, Su=None, plant=None):
        syn = self.syn
        A, B2 = plant if plant is not None else (self.Aj, self.B2j)
        syn.mark(0)
        syn.set_plant(A, B2)
        _, nx_, nu_ = syn.synthesize_window(Sx or self.Sxw, Su or self.Suw, (self.w0 - self.p0, self.w1 - self.p0),
                                            (self.c0 - self.p0, self.c1 - self.p0), out=out, values_only=values_only,
                                            row_off=(self.p0, self.u0))
        syn.mark(1)
        return syn.elapsed_ms(), nx_, nu_


def grid_config4(syn_factory, side, steps, warmup, peaks_fp64):
    """BASELINE config #4: side x side grid (5-point stencil A, B2 = I), d = 4, T = 20 — the large-block size classes."""
    import slc_b200
    from slc_b200 import workloads as W
    d4, T4 = 4, 20
    A, B1, B2 = W.grid_plant(side, side)
    Sx, Su = W.grid_supports(side, side, d4, T4, ALPHA)
    syn = syn_factory()
    try:
        syn.set_plant(slc_b200.JuliaCSC.from_scipy(A, np.float64), slc_b200.JuliaCSC.from_scipy(B2, np.float64))
        syn.set_supports(Sx, Su)

        def step():
            syn.mark(0); syn.plan(None, sizes=False); syn.solve_device(); syn.mark(1)
            return syn.elapsed_ms()
        for _ in range(warmup):
            step()
        ms, solve, launches = [], [], 0
        for _ in range(steps):
            ms.append(step())
            tm = syn.timing()
            solve.append(tm["solve_ms"]); launches += tm["gather_launches"] + tm["solve_launches"] + tm["scatter_launches"]
        tm = syn.timing()
        status, resid = syn.solve_status()
        assert resid.max() < 1e-12 and not (status & 1).any(), "config4 failed its feasibility check"
    finally:
        syn.close()
    Nx = side * side
    solve_ms = float(np.mean(solve))
    ach = tm["alg_flops"] / (solve_ms * 1e-3) / 1e12
    traffic, src = ncu_traffic(f"grid_{side}x{side}:solve")
    return {"workload": f"grid {side}x{side} Nx={Nx}, 5-point A = I + 0.1 adj, B2 = I, d={d4}, T={T4}, alpha={ALPHA} (BASELINE config #4; "
                        "interior column n=61, m=85: the N=64 size class)",
            "metric": "Phi synthesis throughput, localized H2 state feedback (grid)", "unit": UNIT,
            "value": Nx / (float(np.mean(ms)) * 1e-3), "ms_per_step": float(np.mean(ms)), "steps": steps, "warmup": warmup,
            "kernel_ms": {"gather": tm["gather_ms"], "gather_gpu": tm["gather_gpu_ms"], "solve": solve_ms, "scatter": tm["scatter_ms"]},
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peaks_fp64["dmma"], "unit": "TFLOP/s", "frac": ach / peaks_fp64["dmma"],
                         "frac_of_dfma_peak": ach / peaks_fp64["dfma"], "traffic": traffic, "traffic_source": src,
                         "kernel": "sls_solve kernels of all size classes of one step", "flops_per_launch": tm["alg_flops"],
                         "peak_source": "sls_fp64_peak DMMA chain, measured in this run"},
            "resid_max": float(resid.max()), "gpu_launches": int(launches)}, launches


def sim_config5(syn, nx, hbm_peak, hbm_src):
    """BASELINE config #5 on the handle's resident Nx = 100k controller (plan(None) + solve_device must have run)."""
    ns, steps = 4096, 250
    rng = np.random.default_rng(0)
    js = rng.integers(0, nx, ns); ts = rng.integers(1, steps - T, ns)
    syn.simulate(js, ts, steps)                                  # warm-up: the handle's scratch (1.4 GB) is allocated once
    t0 = time.perf_counter()
    sim = syn.simulate(js, ts, steps)
    wall = time.perf_counter() - t0
    tm = syn.timing()
    # closed-loop identity on a sample: zero before the impulse and after the FIR horizon
    for s in range(0, ns, 97):
        t0_ = int(ts[s])
        assert np.abs(sim["xnorm"][s][:t0_]).max() == 0.0 and np.abs(sim["xnorm"][s][t0_ + T:]).max() < 1e-12
    traffic, src = ncu_traffic("chain_100000:sim")
    ach = tm["sim_bytes"] / (tm["sim_ms"] * 1e-3) / 1e9 if tm["sim_ms"] > 0 else None
    return {"workload": f"closed-loop FIR simulation (README.md:62-71) of the Nx={nx} controller, {ns} impulse scenarios "
                        f"(j~U, t~U{{1..{steps - T}}}, numpy seed 0), {steps} steps (BASELINE config #5)",
            "metric": "scenario-steps/s", "value": ns * steps / (tm["sim_ms"] * 1e-3), "sim_kernel_ms": tm["sim_ms"],
            "window_and_pack_ms": tm["sim_pack_ms"], "call_wall_ms": 1e3 * wall,
            "window_states": tm["sim_window"], "window_inputs": tm["sim_uwindow"],
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": (ach / hbm_peak) if ach else None,
                         "traffic": traffic, "traffic_source": src,
                         "kernel": "sls_sim_cta_kernel (one CTA per scenario; the row-ordered window of the 2 x 148 resident scenarios stays in L2)",
                         "bytes_per_launch": tm["sim_bytes"],
                         "bytes_definition": "packed window Phi entries (8 B value + 2 B local row) the recursion multiplies, summed over "
                                             "scenarios, steps and taps (sls_timing.sim_bytes)",
                         "peak_source": hbm_src},
            "gpu_launches": 5}


def run_ours(args, rank, world, local_rank):
    import torch
    import slc_b200
    from slc_b200 import sharding

    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None or x is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")

    # host threads the library may use for host-side copies, per rank (its own default assumes it owns the box)
    os.environ.setdefault("SLS_HOST_COPY_THREADS", str(max(2, min(8, (os.cpu_count() or 8) // max(world, 1) // 2))))
    syn = slc_b200.Synthesizer(device=local_rank)
    import ctypes as C
    dfma, dmma = C.c_double(), C.c_double()
    syn.api.sls_fp64_peak(local_rank, 4000, C.byref(dfma), C.byref(dmma))
    fp64 = {"dfma": dfma.value, "dmma": dmma.value}

    if args.workload == "grid":                                # stand-alone config #4 line
        if rank == 0:
            obj, _ = grid_config4(lambda: slc_b200.Synthesizer(device=local_rank), args.grid_side, max(2, min(args.steps, 5)),
                                  max(1, min(args.warmup, 2)), fp64)
            obj.update(n_gpus=1, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                       config={"workload": obj["workload"]})
            print(json.dumps(obj), flush=True)
        syn.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- headline: 100 000 columns per GPU (weak)
    def measure(shard, steps, warmup, e2e_reps):
        for _ in range(warmup):
            shard.device_step()
        barrier()
        step_ms, k_ms, launches = [], {"gather_ms": [], "gather_gpu_ms": [], "solve_ms": [], "scatter_ms": []}, 0
        for _ in range(steps):
            step_ms.append(shard.device_step())
            tm = syn.timing()
            for k in k_ms:
                k_ms[k].append(tm[k])
            launches += tm["gather_launches"] + tm["solve_launches"] + tm["scatter_launches"]
        barrier()
        tm = syn.timing()
        res = {"total_ms": float(np.sum(step_ms)), "k_ms": {k: float(np.mean(v)) for k, v in k_ms.items()}, "launches": launches, "tm": tm,
               "e2e_ms": None}
        if os.environ.get("BENCH_DEBUG"):
            print(f"[rank {rank}] steps {[round(x, 2) for x in step_ms]} kernels {res['k_ms']}", file=sys.stderr, flush=True)
        if e2e_reps:
            out = shard.alloc_slab()
            for _ in range(2):
                shard.e2e_step(out)
            barrier()
            e2e = [shard.e2e_step(out)[0] for _ in range(e2e_reps)]
            barrier()
            assert out["resid"].max() < 1e-9, "e2e result failed its own feasibility check"
            # rowval outputs the library filled host-side from the caller's own S[t].rowval (pattern proven equal on the
            # device) never crossed PCIe: they are not device-to-host bytes
            hcb = int(syn.timing()["host_copied_bytes"])
            res.update(e2e_ms=float(np.mean(e2e)), out=out, d2h=shard.d2h_full - hcb, host_copied=hcb)
        return res

    nx_local = args.nx
    Nx = nx_local * world
    shard = ChainShard(syn, Nx, rank * nx_local, (rank + 1) * nx_local, with_e2e=not args.no_e2e)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e2e_reps = 0 if args.no_e2e else max(2, min(args.steps, 5))
    main = measure(shard, args.steps, args.warmup, e2e_reps)
    clocks = sampler.stop() if sampler else None
    tm = main["tm"]

    # ---------------------------------------------------------------- variants of the host path (N = 1 only)
    variants = {}
    if world == 1 and not args.no_e2e:
        out = main["out"]
        ref_vals = [a.copy() for a in (out["x"]["nzval"][0], out["x"]["nzval"][T - 1], out["u"]["nzval"][T - 1])]

        def same_values(o):
            return all(np.array_equal(a, b) for a, b in zip(ref_vals, (o["x"]["nzval"][0], o["x"]["nzval"][T - 1], o["u"]["nzval"][T - 1])))
        # (1) values only: nnz(Phi[t]) == nnz(S[t][:, range]) proves pattern(Phi[t]) == pattern(S[t]); the host aliases S[t].rowval
        outv = shard.alloc_slab(values_only=True)
        shard.e2e_step(outv, values_only=True)
        r = [shard.e2e_step(outv, values_only=True) for _ in range(3)]
        pattern_is_support = all(int(r[-1][1][t]) == outv["x"]["nzval"][t].size for t in range(T)) and \
            all(int(r[-1][2][t]) == outv["u"]["nzval"][t].size for t in range(T))
        if pattern_is_support and same_values(outv):
            variants["e2e_values_only"] = {"value": shard.ncols / (np.mean([x[0] for x in r]) * 1e-3), "unit": UNIT,
                                           "ms_per_step": float(np.mean([x[0] for x in r])), "d2h_bytes_per_step": int(shard.d2h_full),
                                           "what": "sls_synthesize_window with NULL rowval outputs: the returned nnz(Phi[t]) == nnz(S[t]) proves "
                                                   "the stored pattern is S[t]'s, so the host aliases its own S[t].rowval (0.53 GB less D2H); "
                                                   "the documented Julia fast path (integration/SLSB200.jl), not the headline"}
        # (1b) every rowval output copied back from the device (flags bit 3 / SLS_PIPE_ROWS_FROM_DEVICE): the call as it was
        #      before the library learnt to fill pattern-equal taps from the caller's own S[t].rowval
        os.environ["SLS_PIPE_ROWS_FROM_DEVICE"] = "1"
        try:
            shard.e2e_step(out)
            rd = float(np.mean([shard.e2e_step(out)[0] for _ in range(3)]))
            assert int(syn.timing()["host_copied_bytes"]) == 0 and same_values(out)
        finally:
            os.environ.pop("SLS_PIPE_ROWS_FROM_DEVICE", None)
        variants["e2e_rows_from_device"] = {"value": shard.ncols / (rd * 1e-3), "unit": UNIT, "ms_per_step": rd,
                                            "d2h_bytes_per_step": int(main["d2h"] + main["host_copied"]),
                                            "what": "same call, rowval outputs always copied from the device (sls_options.flags bit 3)"}
        # (2) the three-call form of the ABI
        out3 = None
        if shard.cols is None:
            def e2e3():
                syn.mark(0); syn.set_plant(shard.Aj, shard.B2j); syn.set_supports(shard.Sx, shard.Su); syn.plan(None); syn.solve(out3); syn.mark(1)
                return syn.elapsed_ms()
            syn.set_supports(shard.Sx, shard.Su); syn.plan(None); out3 = syn.alloc_outputs(pinned=True)
            e2e3()
            variants["e2e_three_call"] = {"value": shard.ncols / (np.mean([e2e3() for _ in range(3)]) * 1e-3), "unit": UNIT,
                                          "api": "sls_set_plant + sls_set_supports + sls_plan + sls_solve"}
            variants["e2e_three_call"]["ms_per_step"] = 1e3 * shard.ncols / variants["e2e_three_call"]["value"]
            assert np.array_equal(out3["x"]["nzval"][T - 1], ref_vals[1]), "three-call and one-shot results differ"
        # (3) ordinary (pageable) caller buffers, and the same buffers page-locked in place by sls_host_register
        import slc_b200 as S_
        unp = lambda M: S_.JuliaCSC(M.m, M.n, M.colptr.copy(), M.rowval.copy(), M.nzval.copy())
        cache = {}

        def unp_list(L):
            res = []
            for M in L:
                if id(M.rowval) not in cache:
                    cache[id(M.rowval)] = unp(M)
                res.append(cache[id(M.rowval)])
            return res
        Sxp, Sup = unp_list(shard.Sxw), unp_list(shard.Suw)
        plantp = (unp(shard.Aj), unp(shard.B2j))
        outp = shard.alloc_slab(pinned=False)
        shard.e2e_step(outp, Sx=Sxp, Su=Sup, plant=plantp)
        pg = float(np.mean([shard.e2e_step(outp, Sx=Sxp, Su=Sup, plant=plantp)[0] for _ in range(2)]))
        variants["e2e_pageable"] = {"value": shard.ncols / (pg * 1e-3), "unit": UNIT, "ms_per_step": pg,
                                    "what": "same call, ordinary pageable host arrays (what a Julia Vector is before sls_host_register)"}
        if same_values(outp) and hasattr(syn.api, "sls_host_register"):
            regd, t0 = [], time.perf_counter()
            try:
                for M in list(cache.values()) + list(plantp):
                    for arr in (M.colptr, M.rowval, M.nzval):
                        if arr.nbytes:
                            syn.host_register(arr); regd.append(arr)
                for name in ("x", "u"):
                    for k in ("colptr", "rowval", "nzval"):
                        for arr in outp[name][k]:
                            if arr.nbytes:
                                syn.host_register(arr); regd.append(arr)
                reg_ms = 1e3 * (time.perf_counter() - t0)
                shard.e2e_step(outp, Sx=Sxp, Su=Sup, plant=plantp)
                rg = float(np.mean([shard.e2e_step(outp, Sx=Sxp, Su=Sup, plant=plantp)[0] for _ in range(3)]))
                variants["e2e_registered"] = {"value": shard.ncols / (rg * 1e-3), "unit": UNIT, "ms_per_step": rg,
                                              "register_once_ms": reg_ms,
                                              "what": "the same pageable arrays after sls_host_register (one-time cost in register_once_ms)"}
            finally:
                for arr in regd:
                    syn.host_unregister(arr)
        # (4) supports generated on the device instead of shipped over PCIe (not the reference-facing signature)
        if out3 is not None:
            def e2e_ds():
                syn.mark(0); syn.set_plant(shard.Aj, shard.B2j); syn.make_supports(D, T, ALPHA); syn.plan(None); syn.solve(out3); syn.mark(1)
                return syn.elapsed_ms()
            e2e_ds()
            ds = float(np.mean([e2e_ds() for _ in range(3)]))
            variants["e2e_device_supports"] = {"value": shard.ncols / (ds * 1e-3), "unit": UNIT, "ms_per_step": ds,
                                               "what": "three-call form with sls_make_supports(d,T,alpha) on the device instead of sls_set_supports"}
            assert np.array_equal(out3["x"]["nzval"][T - 1], ref_vals[1]), "device-generated supports changed the result"
        syn.set_plant(shard.Aj, shard.B2j)
        syn.set_supports(shard.Sx, shard.Su)

    # N > 1: the values-only form of the host path as well.  The host's aggregate device-to-host ceiling (71-92 GB/s on
    # these boxes whatever the GPU count, profiles/r02_e2e_pipeline_trace.txt) is what holds `e2e` at N = 8; this form
    # moves half the bytes down.
    if world > 1 and not args.no_e2e:
        outv = shard.alloc_slab(values_only=True)
        shard.e2e_step(outv, values_only=True)
        barrier()
        r = [shard.e2e_step(outv, values_only=True) for _ in range(3)]
        barrier()
        ok = all(int(r[-1][1][t]) == outv["x"]["nzval"][t].size for t in range(T)) and \
            all(int(r[-1][2][t]) == outv["u"]["nzval"][t].size for t in range(T)) and float(outv["resid"].max()) < 1e-9
        ok = allsum(0.0 if ok else 1.0) == 0.0
        vo_ms = allmax(float(np.mean([x[0] for x in r])))
        vo_d2h = int(allsum(float(shard.d2h_full)))
        if ok:
            variants["e2e_values_only"] = {"value": nx_local * world / (vo_ms * 1e-3), "unit": UNIT, "ms_per_step": vo_ms,
                                           "d2h_bytes_per_step": vo_d2h,
                                           "what": "sls_synthesize_window with NULL rowval outputs on every rank (nnz(Phi[t]) == nnz(S[t]) checked: "
                                                   "the host aliases its own S[t].rowval); the documented Julia fast path, not the headline"}
        del outv

    # ---------------------------------------------------------------- config #5 on the resident controller, config #4 (N = 1)
    config5 = config4 = None
    extra_launches = 0
    if world == 1 and not args.no_configs and shard.cols is None:
        syn.plan(None); syn.solve_device()
        config5 = sim_config5(syn, nx_local, hbm_peak, hbm_src)
        extra_launches += config5["gpu_launches"]

    ag_ms = None
    if args.allgather and dist is not None:
        class _Dev:                                             # device buffer -> torch tensor without a copy
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        syn.plan(shard.cols); syn.solve_device()
        pz, nnz = C.c_void_p(), C.c_int64()
        syn._ck(syn.api.sls_device_output(syn.h, 0, 0, None, None, C.byref(pz), C.byref(nnz)))
        total = int(tm["nnz_out"])
        mine = torch.as_tensor(_Dev(pz.value, total), device=f"cuda:{local_rank}")
        cap = torch.tensor([total], device="cuda"); dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        send = torch.zeros(int(cap.item()), dtype=torch.float64, device="cuda"); send[:total] = mine
        recv = torch.empty(world * int(cap.item()), dtype=torch.float64, device="cuda")
        for _ in range(2):
            dist.all_gather_into_tensor(recv, send)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_gather_into_tensor(recv, send)
        e1.record(); torch.cuda.synchronize()
        ag_ms = allmax(e0.elapsed_time(e1) / 5)

    # ---------------------------------------------------------------- strong scaling: 100 000 columns TOTAL
    strong = None
    if not args.no_strong:
        nx_s = NX_PER_GPU
        if world == 1 and nx_local == nx_s:
            s_total, s_e2e, s_rng = main["total_ms"] / args.steps, main["e2e_ms"], [0, nx_s]
            s_launch = main["launches"] / args.steps
        else:
            cost = sharding.chain_cost_model(nx_s, D, T, ALPHA)
            s_rng = [int(b) for b in sharding.balanced_ranges(cost, world)]
            sh = ChainShard(syn, nx_s, s_rng[rank], s_rng[rank + 1], with_e2e=not args.no_e2e)
            r = measure(sh, args.steps, args.warmup, 0 if args.no_e2e else 3)
            s_total, s_e2e = r["total_ms"] / args.steps, r["e2e_ms"]
            s_launch = r["launches"] / args.steps
        s_total, s_e2e = allmax(s_total), allmax(s_e2e)
        strong = {"workload": f"chain Nx={nx_s} TOTAL over {world} GPU(s), cost-balanced contiguous column ranges (sharding.balanced_ranges)",
                  "ranges": s_rng, "scaling": "strong", "ms_per_step": s_total, "value": nx_s / (s_total * 1e-3), "unit": UNIT,
                  "launches_per_step": allsum(float(s_launch))}
        if s_e2e is not None:
            strong["e2e"] = {"ms_per_step": s_e2e, "value": nx_s / (s_e2e * 1e-3), "unit": UNIT}

    # ---------------------------------------------------------------- config #5 on N > 1 GPUs: scenarios sharded by owner column
    # sls_simulate only needs the Phi columns of the scenario's window (within 19 states of j on the chain), so rank r takes
    # the scenarios whose disturbance column lies in its column range and synthesises that range plus a halo — no collective.
    if world > 1 and not args.no_configs:
        nx5, ns5, steps5, H5 = NX_PER_GPU, 4096, 250, 24
        rng5 = [int(b) for b in sharding.equal_ranges(nx5, world)]
        a5, b5 = rng5[rank], rng5[rank + 1]
        sh5 = ChainShard(syn, nx5, max(0, a5 - H5), min(nx5, b5 + H5), with_e2e=False)
        syn.plan(sh5.cols); syn.solve_device()
        g = np.random.default_rng(0)
        js = g.integers(0, nx5, ns5); ts = g.integers(1, steps5 - T, ns5)
        mine = (js >= a5) & (js < b5)
        syn.simulate(js[mine] - sh5.p0, ts[mine], steps5)              # warm-up: scratch allocated once
        sim = syn.simulate(js[mine] - sh5.p0, ts[mine], steps5)
        tm5 = syn.timing()
        for k in range(0, int(mine.sum()), 53):
            t0_ = int(ts[mine][k])
            assert np.abs(sim["xnorm"][k][:t0_]).max() == 0.0 and np.abs(sim["xnorm"][k][t0_ + T:]).max() < 1e-12
        sim_ms5 = allmax(float(tm5["sim_ms"])); pack_ms5 = allmax(float(tm5["sim_pack_ms"])); bytes5 = allsum(float(tm5["sim_bytes"]))
        config5 = {"workload": f"closed-loop FIR simulation of the Nx={nx5} controller, {ns5} impulse scenarios (numpy seed 0), {steps5} steps "
                               f"(BASELINE config #5), scenarios sharded over {world} GPUs by owner column (each rank synthesises its range "
                               f"+ {H5} halo columns; no collective)",
                   "metric": "scenario-steps/s", "value": ns5 * steps5 / (sim_ms5 * 1e-3), "sim_kernel_ms": sim_ms5,
                   "window_and_pack_ms": pack_ms5, "scenarios_per_rank_max": int(allmax(float(mine.sum()))),
                   "roofline": {"bound": "l2/hbm", "achieved": bytes5 / (sim_ms5 * 1e-3) / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                                "frac": bytes5 / (sim_ms5 * 1e-3) / 1e9 / (hbm_peak * world), "bytes_per_launch": bytes5,
                                "kernel": "sls_sim_cta_kernel", "peak_source": hbm_src + f" x {world}"}}
        syn.set_plant(shard.Aj, shard.B2j)

    total_ms = allmax(main["total_ms"])
    e2e_ms = allmax(main["e2e_ms"])
    launches_all = int(allsum(float(main["launches"])))
    h2d_all = int(allsum(float(shard.h2d)))
    d2h_all = int(allsum(float(main.get("d2h", 0))))
    syn.close()

    if rank == 0 and world == 1 and not args.no_configs:
        config4, l4 = grid_config4(lambda: slc_b200.Synthesizer(device=local_rank), args.grid_side, 3, 1, fp64)
        extra_launches += l4

    if rank == 0:
        k = main["k_ms"]
        solve_ms = k["solve_ms"]
        gs_ms = k["gather_gpu_ms"] + k["scatter_ms"]
        ach_tf = tm["alg_flops"] / (solve_ms * 1e-3) / 1e12
        bytes13 = tm["alg_bytes"] - 8.0 * tm["nnz_out"]              # B_col minus the nzval half of the writes (kernel 2's)
        ach_gbs = bytes13 / (gs_ms * 1e-3) / 1e9
        ncols_all = nx_local * world
        t_solve, src_solve = ncu_traffic(f"chain_{nx_local}:solve")
        t_gs, src_gs = ncu_traffic(f"chain_{nx_local}:gather_scatter")
        line = {
            "metric": METRIC, "value": ncols_all * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"chain Nx={nx_local}/GPU ({Nx} total), Nu=Nx/3, d={D}, T={T}, alpha={ALPHA} (BASELINE config #3)",
                       "columns_per_gpu": nx_local,
                       "parallelism": f"column-sharded x{world}: each rank holds the sub-plant and support window of its range, no collective",
                       "l2": "inputs (0.57 GB supports) and outputs (1.06 GB) exceed the 126 MB L2; no flush needed",
                       "nnz_out_per_gpu": int(tm["nnz_out"])},
            "kernel_ms": {"gather": k["gather_ms"], "gather_gpu": k["gather_gpu_ms"], "solve": solve_ms, "scatter": k["scatter_ms"]},
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64["dmma"], "unit": "TFLOP/s", "frac": ach_tf / fp64["dmma"],
                         "frac_of_dfma_peak": ach_tf / fp64["dfma"], "dfma_peak": fp64["dfma"],
                         "traffic": t_solve, "traffic_source": src_solve,
                         "kernel": "sls_solve_kernel (kernel 2, all size classes of one step)",
                         "flops_per_launch": tm["alg_flops"],
                         "peak_source": "sls_fp64_peak DMMA (mma.sync m8n8k4 f64) chain, measured in this run; the contraction runs on DMMA",
                         "refined_columns": int(tm["refined"])},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                             "traffic": t_gs, "traffic_source": src_gs,
                             "kernel": "kernels 1a + class lists + 1b + colptr scan (GPU time, CUDA events around the launches) + kernel 3 (rowval "
                                       "scatter; runs alone before kernel 2)",
                             "gather_gpu_ms": k["gather_gpu_ms"], "scatter_ms": k["scatter_ms"],
                             "bytes_per_launch": bytes13, "gpu_ms": gs_ms, "host_round_trip_ms": k["gather_ms"] - k["gather_gpu_ms"],
                             "bytes_definition": "SURVEY 8d B_col summed over the columns minus the 8 B/entry nzval writes (those are kernel 2's)",
                             "peak_source": hbm_src},
            "gpu_launches": launches_all + extra_launches,
            "clocks": clocks,
        }
        if ag_ms is not None:
            line["allgather"] = {"ms": ag_ms, "bytes_per_rank": int(tm["nnz_out"]) * 8, "what": "NCCL all_gather_into_tensor of the F64 value slabs"}
        if e2e_ms is not None:
            line["e2e"] = {"value": ncols_all / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                           "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                           "h2d_bytes_per_rank": int(shard.h2d), "d2h_bytes_per_rank": int(main.get("d2h", 0)),
                           "host_copied_bytes_per_rank": int(main.get("host_copied", 0)),
                           "host_copied_what": "rowval outputs of taps whose stored pattern the device proved equal to the support's "
                                               "(equal stored-entry counts; stored rows are a subset of the support column) are filled "
                                               "from the caller's own S[t].rowval by host threads instead of a device-to-host copy; "
                                               "bit-identical (tests: check_pipelined_rowval_sources); e2e_rows_from_device is the call without it",
                           "api": "sls_set_plant (the rank's sub-plant) + sls_synthesize_window (support window in, Phi column slab out; "
                                  "column-chunked copy/compute pipeline), pinned host buffers, full CSC outputs (colptr + rowval + nzval)"}
            line.update(variants)
        if strong is not None:
            line["strong"] = strong
        if config4 is not None:
            line["config4"] = config4
        if config5 is not None:
            line["config5"] = config5
        if not args.no_cpu_baseline and world == 1:
            arm = CpuArm()
            n = args.cpu_sample or 64 * arm.cores
            arm.run(arm.cores * 2)
            dt = arm.run(n)
            arm.close()
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": arm.cores, "kind": "port",
                                    "sample": f"{n} columns evenly spaced over a {arm.nx}-state chain (same d,T,alpha), CPU oracle, "
                                              f"{arm.cores} processes, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
