#!/usr/bin/env python3
"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py X_cuda_sass.csv [top_n]"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    cur, f, data = None, None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            f = r[1].split('/')[-1]
            continue
        if len(r) >= 2 and r[0] == 'Function Name':
            cur = r[1][:60]
            data.setdefault(cur, [])
            continue
        if cur and len(r) > 10 and r[0].isdigit() and r[2] == '-':
            data[cur].append((f, r))
    for k, v in data.items():
        tot = sum(int(r[4]) for _, r in v)
        inst = sum(int(r[7]) for _, r in v)
        print("=====", k, "samples", tot, "warp-instructions", inst)
        for f, r in sorted(v, key=lambda x: -int(x[1][4]))[:top]:
            print("%-16s %5s %5.1f%% inst=%10s  %s" % (f, r[0], 100 * int(r[4]) / max(tot, 1), r[7], r[1][:120]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
