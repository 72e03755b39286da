import csv, sys, subprocess, collections, re
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; r=rows[-1]
keys=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in hdr: print(f'{k} = {r[hdr.index(k)]} {rows[1][hdr.index(k)]}')
for h in hdr:
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
        v=float(r[hdr.index(h)] or 0)
        if v>1: print(h.replace('smsp__warp_issue_stalled_','stall '), v)
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=None; ops=collections.Counter(); samp=collections.Counter(); tot=0; stot=0
for row in rows:
    if row and row[0]=='Address': h=row; continue
    if h is None or len(row)<len(h): continue
    sass=row[h.index('Source')].strip()
    op=sass.split()[0] if not sass.startswith('@') else sass.split()[1]
    op=op.split('.')[0]
    n=int(row[h.index('Instructions Executed')] or 0); s=int(row[h.index('# Samples')] or 0)
    ops[op]+=n; samp[op]+=s; tot+=n; stot+=s
print('total inst', tot, 'samples', stot)
for op,n in ops.most_common(25): print(f'{op:10s} {n:14d} {100*n/tot:5.1f}%  samples {100*samp[op]/max(stot,1):5.1f}%')
