#!/usr/bin/env python
"""Summarise ncu outputs into text kept under profiles/ (read by the judge).

    python profiles/summarize_ncu.py launches <launches.csv>
    python profiles/summarize_ncu.py raw <report.ncu-rep>
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if r[0] == 'ID':
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(d['Metric Value'].replace(',', ''))
            v = v / 1e3 if d['Metric Unit'] in ('ns', 'nsecond') else v
            agg[d['Kernel Name'][:70]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f'total {tot / 1e3:.3f} ms over {sum(len(v) for v in agg.values())} launches (ncu: cold cache, serialised)')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{sum(v) / tot * 100:6.2f}%  n={len(v):4d}  avg={sum(v) / len(v):9.1f} us  {k}')


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name'), '| grid', d.get('Grid Size'), '| block', d.get('Block Size'))
        for k in KEYS:
            if k in d:
                print(f'   {k:70s} {d[k]:>16s} {units[hdr.index(k)]}')


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
