"""Prints the handful of ncu metrics we track from a raw-page CSV export (ncu -i X --page raw --csv)."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    for k in keys:
        for h in hdr:
            if h == k:
                print('%-75s %-10s %s' % (h, units[hdr.index(h)], d[h][:90]))
    st = sorted(((float(d[h]), h) for h in stall), reverse=True)[:7]
    for v, h in st:
        print('   stall %-45s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    print()
