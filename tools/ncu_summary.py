#!/usr/bin/env python
"""Key metrics of one `ncu --set full` report as CSV (what profiles/*_ncu_*.csv hold):  python tools/ncu_summary.py report.ncu-rep"""
import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys=['gpu__time_duration.sum','sm__cycles_elapsed.avg','launch__registers_per_thread','launch__grid_size','launch__block_size',
'smsp__issue_active.avg.per_cycle_active','sm__inst_executed.sum.per_cycle_elapsed','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed','sm__icc_request_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','sass__inst_executed_local_loads','sm__warps_active.avg.pct_of_peak_sustained_active',
'smsp__sass_inst_executed_op_tmem_ldt.sum','smsp__sass_inst_executed_op_tmem_stt.sum','sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active']
d={h:(u,v) for h,u,v in zip(hdr,units,vals)}
for k in keys:
    if k in d: print(f'{k},{d[k][0]},{d[k][1]}')
for h in hdr:
    if 'issue_stalled' in h and 'per_issue_active' in h: print(f'{h},{d[h][0]},{d[h][1]}')
