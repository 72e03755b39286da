"""One block per profiled launch of an .ncu-rep (ncu --set full): duration, launch shape, occupancy limits, DRAM / L2
traffic, issue / pipe utilisation and the stall reasons above 5 % — the summaries committed under profiles/.
    python tools/ncu_kernels_table.py gpurun_out/<file>.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
print(f"# {rep}: ncu --set full --clock-control none --import-source on (one block per profiled launch)")
for r in rows[2:]:
    print("----")
    for k in keys:
        if k in hdr:
            print(f"{k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[hdr.index(h)] or 0)
            except ValueError:
                continue
            st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    tot = sum(v for v, _ in st) or 1.0
    print("stall share (warps stalled per issue-active cycle, normalised): " +
          ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True) if v / tot > 0.04))
