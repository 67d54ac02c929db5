"""Dynamic instruction mix of the LMM Euler kernel's rate loop from an ncu source-page export.

usage: ncu -i prof.ncu-rep --page source --csv > src.csv ; python profiles/tools/hot_path.py src.csv PATHS RATE_STEPS_PER_PATH [--list]
Prints the warp-level instruction count per rate-step by opcode (instructions executed at least 0.3 x per chunk) and the stall samples.
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
paths, per_path = float(sys.argv[2]), float(sys.argv[3])
hdr = rows[1]
ia, isrc, ie, iss = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
rs = paths / 32 * per_path
data = []
for r in rows[2:]:
    try:
        data.append((r[ia], r[isrc], int(r[ie]), int(r[iss])))
    except (ValueError, IndexError):
        pass
print("kernel:", rows[0][1][:100])
print("warp instructions per rate-step (all): %.1f" % (sum(d[2] for d in data) / rs))
hot = [d for d in data if d[2] >= 0.3 * rs / 2]
print("hot-path instructions per rate-step: %.1f (%d SASS lines)" % (sum(d[2] for d in hot) / rs, len(hot)))
cat, stall = collections.Counter(), collections.Counter()
for a, s, e, st in hot:
    tok = s.split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    cat[op] += e / rs
    stall[op] += st
tot_st = sum(d[3] for d in data)
fp64 = sum(v for k, v in cat.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
print("FP64 per rate-step: %.1f   other: %.1f" % (fp64, sum(cat.values()) - fp64))
for k, v in cat.most_common():
    print("%-8s %6.2f   stall samples %5.1f %%" % (k, v, 100.0 * stall[k] / tot_st))
if "--list" in sys.argv:
    base = None
    for a, s, e, st in data:
        x = int(a, 16) if a.startswith("0x") else int(a)
        base = x if base is None else base
        if e >= 0.3 * rs / 2:
            print("%05x %5.2f %6d  %s" % (x - base, e / rs * 2, st, s))
