#!/usr/bin/env python
"""Dynamic SASS opcode histogram of one kernel from an ncu report (executed warp instructions per opcode) and the proof
lines for tcgen05 / TMA / packed FP32:  ncu -i X.ncu-rep --page source --print-source sass --csv > X.sass.csv ;
python profiles/sass_hist.py X.sass.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data, name = None, [], "?"
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) > 5:
        data.append(r)
ie = hdr.index("Instructions Executed")
tot = collections.Counter()
static = collections.Counter()
for r in data:
    toks = r[1].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    n = int(r[ie]) if r[ie].isdigit() else 0
    tot[op] += n
    static[op] += 1
total = sum(tot.values()) or 1
print(f"kernel: {name[:120]}")
print(f"static SASS instructions {sum(static.values())}, executed warp instructions {total}")
for op, n in tot.most_common(45):
    print(f"  {op:28s} {n:14d}  {100.0 * n / total:5.1f} %   (static {static[op]})")
special = [op for op in static if any(k in op for k in ("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "FFMA2", "HMMA", "UTCBAR", "SYNCS"))]
if special:
    print("tensor-core / TMEM / TMA / packed-FP32 opcodes present (static counts): " + ", ".join(f"{op} x{static[op]}" for op in sorted(special)))
