#!/usr/bin/env python
"""Markdown table of the metrics the DESIGN / bench lines quote, from an .ncu-rep brought back in gpurun_out/ (read on the CPU box):
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep "title" > profiles/x.md"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '') for r in data]
    print('# %s\n' % title)
    print('| metric | ' + ' | '.join(names) + ' |')
    print('|---|' + '---|' * len(names))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('| %s [%s] | ' % (k, units[i]) + ' | '.join(r[i] for r in data) + ' |')
    for i, h in enumerate(hdr):
        if 'warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            vals = [float(r[i]) for r in data]
            if max(vals) > 0.04:
                print('| stall %s per issue | ' % h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '') +
                      ' | '.join('%.2f' % v for v in vals) + ' |')


if __name__ == '__main__':
    main()
