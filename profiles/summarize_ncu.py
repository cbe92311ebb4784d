#!/usr/bin/env python
"""Summarise ncu outputs into text kept under profiles/ (read by the judge).

    python profiles/summarize_ncu.py launches <launches.csv>
    python profiles/summarize_ncu.py raw <report.ncu-rep>
    python profiles/summarize_ncu.py traffic <report.ncu-rep> [out.json]   DRAM bytes per launch of every kernel in the
                                         report -> JSON that bench.py reads for `roofline.traffic`
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


def short_name(full):
    """b200d::ncc_kernel<(int)11, (int)4>(...) -> ncc_kernel<11,REFINE> etc."""
    import re
    m = re.search(r'(\w+)<([^>]*)>', full)
    if not m:
        return full.split('(')[0]
    args = [a.replace('(int)', '').strip() for a in m.group(2).split(',')]
    modes = {'0': 'STATS', '1': 'COST', '2': 'CONF', '3': 'SHIFT', '4': 'REFINE', '5': 'SWEEP', '6': 'SWEEP_DUMP', '7': 'VPROP'}
    cands = {'0': 'MAP', '1': 'SHIFT', '2': 'REFINE', '3': 'PLANE'}
    bounds = {'0': 'PATCH', '1': 'IMAGE', '2': 'DEPTH'}
    name = m.group(1)
    if name == 'ncc_kernel':
        return f'ncc_kernel<{args[0]},{modes.get(args[1], args[1])}>'
    if name == 'ncc_sep_kernel':
        return f'ncc_sep_kernel<{args[0]},{modes.get(args[1], args[1])},{"bounded" if args[2] == "1" else "fast"}>'
    if name in ('sample_kernel', 'sample_fast_kernel'):
        return f'{name}<{cands.get(args[0], args[0])},{bounds.get(args[1], args[1])}' + (f',{args[2]}>' if len(args) > 2 else '>')
    return f'{name}<{",".join(args)}>'


def traffic(path, out=None):
    import json
    res = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(res.splitlines()))
    hdr, units = rows[0], rows[1]

    def to_bytes(v, unit):
        v = float(v.replace(',', ''))
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    agg = collections.defaultdict(list)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rd = to_bytes(d['dram__bytes_read.sum'], units[hdr.index('dram__bytes_read.sum')])
        wr = to_bytes(d['dram__bytes_write.sum'], units[hdr.index('dram__bytes_write.sum')])
        agg[short_name(d['Kernel Name'])].append((rd + wr, float(d['gpu__time_duration.sum'].replace(',', '')),
                                                  d.get('Grid Size'), d.get('Block Size')))
    doc = {'report': path, 'how': 'ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, mean over '
                                  'the captured launches of each kernel (ncu flushes caches between passes: cold-cache figures)',
           'kernels': {k: {'launches': len(v), 'dram_bytes_per_launch': sum(x[0] for x in v) / len(v),
                           'duration_us': sum(x[1] for x in v) / len(v), 'grid': v[0][2], 'block': v[0][3]}
                       for k, v in agg.items()}}
    text = json.dumps(doc, indent=1)
    if out:
        open(out, 'w').write(text + '\n')
    print(text)


if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'traffic':
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        {'launches': launches, 'raw': raw}[cmd](sys.argv[2])
