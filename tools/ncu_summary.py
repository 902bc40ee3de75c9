#!/usr/bin/env python3
"""Summarises an `ncu --page raw --csv` export: one block of key metrics per profiled launch.
usage: ncu -i X.ncu-rep --page raw --csv > X_raw.csv; python tools/ncu_summary.py X_raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('-----')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-82s %s %s" % (w, r[i][:90], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(r[i].replace(',', '')), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("stall reasons (warps per issue-active cycle): " + ", ".join("%s=%.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
