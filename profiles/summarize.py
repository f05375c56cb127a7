"""Summarise an .ncu-rep (read here, no GPU needed) into the metrics DESIGN.md quotes.
    python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rN_name_ncu_summary.txt"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== kernel:', r[hdr.index('Kernel Name')][:100], ' grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for k in KEYS:
        if k in hdr:
            print('  %-70s %-14s %s' % (k, units[hdr.index(k)], r[hdr.index(k)]))
    st = [(float(r[i]), k) for i, k in enumerate(hdr) if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and r[i] not in ('', 'n/a')]
    print('  top stalls (warps per issue-active):', ', '.join('%s=%.2f' % (k.split('stalled_')[1].split('_per_')[0], v) for v, k in sorted(st, reverse=True)[:6]))
