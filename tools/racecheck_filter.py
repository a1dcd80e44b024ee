"""stdin: compute-sanitizer racecheck output in --racecheck-report hazard mode; stdout: the same without the records on the
control word the step loop polls on purpose (shared 0x400-0x403) and without host back traces.
    compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 0 python tools/race_probe.py 2>&1 | python tools/racecheck_filter.py"""
import sys, re
rec = []
def flush():
    if rec and "hazard detected" in rec[0] and not re.search(r"__shared__ 0x40[0-3] ", rec[0]):
        for l in rec:
            if "Host Frame" in l or "Saved host" in l: continue
            sys.stdout.write(re.sub(r"void fg::k_rollout<[^>]*>\([^)]*\)", "K", l))
for line in sys.stdin:
    if line.startswith("========= Error") or line.startswith("========= Warning"):
        flush(); rec = [line]
    elif line.startswith("========="):
        rec.append(line)
    else:
        flush(); rec = []
        sys.stdout.write(line)
flush()
