#!/usr/bin/env python3
"""Key figures of `ncu -i X.ncu-rep --page raw --csv` output (one row per profiled launch): time, DRAM bytes, pipe
utilisation, occupancy, stall reasons per issue.  usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active']


def main():
    rows = [r for r in csv.reader(l for l in sys.stdin if not l.startswith('=='))]
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    for vals in rows[2:]:
        print('# kernel,%s' % vals[ik].replace(',', ';'))
        for i, h in enumerate(hdr):
            if h in KEEP or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h):
                print('%s,%s,%s' % (h, units[i], vals[i]))


if __name__ == '__main__':
    main()
