#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an ncu report:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.cuda.csv ; python profiles/src_hot.py X.cuda.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, out, nfun = None, [], 0
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Function Name':
        nfun += 1
        continue
    if len(r) > 8 and r[0].isdigit():
        inst = int(r[7]) if r[7].isdigit() else 0
        samp = int(r[4]) if r[4].isdigit() else 0
        out.append((inst, samp, cur, int(r[0]), r[1][:120]))
tot = sum(o[0] for o in out) or 1
ts = sum(o[1] for o in out) or 1
print(f'total warp instructions {tot}, stall samples {ts}')
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0] / tot * 100:5.1f}% inst {o[1] / ts * 100:5.1f}% smp  {o[2]}:{o[3]}  {o[4]}")
