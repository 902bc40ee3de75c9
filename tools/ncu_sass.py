#!/usr/bin/env python3
"""SASS-level stall listing from `ncu --page source --csv --print-source cuda,sass` (hot instructions in address order).
usage: python tools/ncu_sass.py X_cs.csv [min_exec] [min_pct]"""
import collections
import csv
import sys


def main(path, min_exec=1500000, min_pct=0.0):
    rows = list(csv.reader(open(path)))
    hdr = [r for r in rows if len(r) > 10 and r[0] == 'Line No'][0]
    cols = {n: i for i, n in enumerate(hdr)}
    out, cur, seen = [], None, set()
    for r in rows:
        if len(r) > 10 and r[0].isdigit() and r[2] == '-':
            cur = r[0]
        elif len(r) > 10 and r[2].startswith('0x'):
            a = int(r[2], 16)
            if a in seen:
                continue
            seen.add(a)
            out.append((a, cur, r))
    out.sort()
    tot = sum(int(o[2][4]) for o in out)
    names = ['stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_math', 'stall_branch_resolving', 'stall_selected',
             'stall_mio', 'stall_lg', 'stall_barrier', 'stall_no_inst', 'stall_dispatch', 'stall_not_selected']
    agg = collections.Counter()
    hot = 0
    for a, l, r in out:
        for n in names:
            try:
                agg[n] += int(r[cols[n]])
            except ValueError:
                pass
        if int(r[7]) >= min_exec:
            hot += 1
            pct = 100 * int(r[4]) / max(tot, 1)
            if pct >= min_pct:
                st = " ".join("%s=%s" % (n[6:9], r[cols[n]]) for n in names if r[cols[n]] not in ('0', ''))
                print("%05x L%-4s %6d %4.1f%%  %-58s %s" % (a & 0xfffff, l, int(r[4]), pct, r[3].strip()[:58], st))
    print("total samples", tot, "hot instructions", hot)
    print(", ".join("%s=%.1f%%" % (n[6:], 100 * v / max(sum(agg.values()), 1)) for n, v in agg.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1500000, float(sys.argv[3]) if len(sys.argv) > 3 else 0.0)
