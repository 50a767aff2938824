#!/usr/bin/env python3
"""Aggregates the ncu source page per CUDA source line (stall samples, executed warp instructions, top stall
reasons).  Input: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:KERNEL > file.csv
Usage: ncu_source_hot.py file.csv [top]"""
import csv, sys


def num(v):
    try:
        return float(v)
    except ValueError:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, per = None, None, []
for r in rows:
    if len(r) == 2:
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    s, i = num(r[6]), num(r[7])
    if s == 0 and i == 0:
        continue
    st = sorted(((num(r[k]), hdr[k][6:]) for k in range(31, 48)), reverse=True)[:3]
    per.append((s, i, cur, int(r[0]), r[1].strip()[:90], " ".join("%s=%d" % (n, v) for v, n in st if v)))
ts, ti = sum(p[0] for p in per), sum(p[1] for p in per)
print("total stall samples %d, warp instructions %d" % (ts, ti))
for s, i, f, ln, src, st in sorted(per, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s   [%s]" % (100 * s / max(ts, 1), 100 * i / max(ti, 1), f, ln, src, st))
