#!/usr/bin/env python
"""Where a kernel loses lanes: per source line, warp instructions x (32 - average active threads), from the source page of an
ncu report (the .cuda.csv that profiles/src_hot.py also reads).
    python profiles/lane_loss.py X.cuda.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, out = None, []
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 9 and r[0].isdigit():
        inst = int(r[7]) if r[7].isdigit() else 0
        th = int(r[8]) if r[8].isdigit() else 0
        out.append((inst, th, cur, int(r[0]), r[1][:100]))
tot = sum(o[0] for o in out) or 1
tt = sum(o[1] for o in out)
tl = sum(max(o[0] * 32 - o[1], 0) for o in out) or 1
print(f'warp instructions {tot}, average active threads {tt / tot:.2f} of 32, lost lane slots {100 * tl / (tot * 32):.1f} %')
for o in sorted(out, key=lambda o: -(o[0] * 32 - o[1]))[:top]:
    print(f"{100 * (o[0] * 32 - o[1]) / tl:5.1f}% of the loss  {o[1] / max(o[0], 1):5.1f} threads  {100 * o[0] / tot:4.1f}% inst  {o[2]}:{o[3]}  {o[4]}")
