#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/ncu_source_lines.py <report.ncu-rep> <kernel-name-regex> [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name',
                          f'regex:{kern}'], capture_output=True, text=True).stdout
    fname, hdr, lines, first = None, None, [], True
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            if not first and hdr is not None and fname is None:
                break
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr and len(r) >= 9 and r[0] != '':
            try:
                lines.append((fname, int(r[0]), r[1].strip()[:100], int(r[7].replace(',', '')), int(r[6].replace(',', ''))))
            except ValueError:
                pass
    ti, ts = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
    print(f'{kern}: {ti} warp instructions, {ts} stall samples over {len(lines)} source lines')
    print('--- by instructions executed')
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f'{l[0]:20s} L{l[1]:<4d} inst {100 * l[3] / ti:5.1f}%  samp {100 * l[4] / ts:5.1f}%  {l[2]}')
    print('--- by stall samples')
    for l in sorted(lines, key=lambda l: -l[4])[:top]:
        print(f'{l[0]:20s} L{l[1]:<4d} inst {100 * l[3] / ti:5.1f}%  samp {100 * l[4] / ts:5.1f}%  {l[2]}')


if __name__ == '__main__':
    main()
