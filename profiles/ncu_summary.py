#!/usr/bin/env python
"""Prints the roofline-relevant metrics of an .ncu-rep (run here: `ncu -i ... --page raw --csv` needs no GPU)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
STALL = 'smsp__average_warps_issue_stalled_'
rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:80], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for k in KEYS:
        if k in hdr:
            print(f'  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
    st = [(float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')]
    print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)[:8]))
