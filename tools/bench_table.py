"""Prints the tables of DESIGN.md §8 from committed bench lines:
    python tools/bench_table.py profiles/r02_bench_n1.json profiles/r02_scale_2.json profiles/r02_scale_4.json profiles/r02_scale_8.json"""
import json
import sys

lines = [json.loads(open(f).read().strip().splitlines()[-1]) for f in sys.argv[1:]]
base = lines[0]
print("N | weak ms  Mcol/s  eff | e2e ms  Mcol/s  eff | strong ms  eff | strong e2e ms  eff | config5 sim ms  eff | h2d B/rank")
for d in lines:
    N = d["n_gpus"]; st = d.get("strong", {}); c5 = d.get("config5", {})
    row = [N, "|", round(d["ms_per_step"], 2), round(d["value"] / 1e6, 2), round(d["value"] / base["value"] / N, 3)]
    if "e2e" in d:
        row += ["|", round(d["e2e"]["ms_per_step"], 1), round(d["e2e"]["value"] / 1e6, 2), round(d["e2e"]["value"] / base["e2e"]["value"] / N, 3)]
    if st:
        row += ["|", round(st["ms_per_step"], 2), round(base["strong"]["ms_per_step"] / st["ms_per_step"] / N, 3)]
        if "e2e" in st:
            row += ["|", round(st["e2e"]["ms_per_step"], 1), round(base["strong"]["e2e"]["ms_per_step"] / st["e2e"]["ms_per_step"] / N, 3)]
    if c5:
        row += ["|", round(c5["sim_kernel_ms"], 1), round(base["config5"]["sim_kernel_ms"] / c5["sim_kernel_ms"] / N, 3)]
    row += ["|", d.get("e2e", {}).get("h2d_bytes_per_rank")]
    print(*row)
d = base
print("\nN = 1:", round(d["ms_per_step"], 2), "ms/step", {k: round(v, 3) for k, v in d["kernel_ms"].items()})
for name in ("roofline", "roofline_hbm"):
    r = d[name]
    print(f"{name}: achieved {r['achieved']:.3f} {r['unit']} / peak {r['peak']:.1f} = {r['frac']:.3f}; traffic {r['traffic']}")
for k in ("e2e", "e2e_values_only", "e2e_three_call", "e2e_pageable", "e2e_registered", "e2e_device_supports"):
    if k in d:
        print(k, round(d[k]["ms_per_step"], 2), "ms")
for k in ("config4", "config5"):
    if k in d:
        c = d[k]
        print(k, {kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in c.items() if kk in ("ms_per_step", "sim_kernel_ms", "window_and_pack_ms", "call_wall_ms", "value")},
              "roofline frac", round(c["roofline"]["frac"], 3), "traffic", c["roofline"]["traffic"])
print("cpu_baseline", d.get("cpu_baseline"))
