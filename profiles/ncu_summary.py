#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): per-kernel headline metrics + top stall reasons.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/<name>.md]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f"## {name[:100]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        st = []
        for i, h in enumerate(hdr):
            if 'smsp__average_warp' in h and 'issue_stalled' in h and 'not_issued' not in h and r[i]:
                st.append((h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), float(r[i].replace(',', ''))))
        print("\nstall cycles per issued instruction (top):",
              ', '.join(f"{k} {v:.2f}" for k, v in sorted(st, key=lambda x: -x[1])[:7]), "\n")


if __name__ == '__main__':
    main(sys.argv[1])
