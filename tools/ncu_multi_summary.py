#!/usr/bin/env python
"""Per-kernel key metrics and top stalls of an ncu report holding several kernels:  python tools/ncu_multi_summary.py report.ncu-rep"""
import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.per_cycle_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
stall=[h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('-----')
    for k in keys:
        if k in d: print(f'{k}: {d[k][:60]} {units[hdr.index(k)]}')
    top=sorted(((float(d[h]),h.split('stalled_')[1].split('_per')[0]) for h in stall), reverse=True)[:6]
    print('stalls:', ', '.join(f'{n} {v:.2f}' for v,n in top))
