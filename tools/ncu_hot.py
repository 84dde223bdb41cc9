"""Summarise an `ncu --page source --csv` dump: share of executed instructions and stall samples per block of
SASS instructions.  Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [block]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "Source" in r and "Address" in r)
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= max(iA, iE, iS) or not r[iE].isdigit():
        if data:
            break
        continue
    data.append(r)
tot = sum(int(r[iE]) for r in data)
stot = sum(int(r[iS]) for r in data)
print("total warp-inst", tot, "sass lines", len(data), "samples", stot)
for b in range(0, len(data), blk):
    e = sum(int(r[iE]) for r in data[b:b + blk])
    s = sum(int(r[iS]) for r in data[b:b + blk])
    ops = {}
    for r in data[b:b + blk]:
        t = r[iA].split()
        op = t[1] if t[0].startswith("@") else t[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[iE])
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print(f"{b:5d} inst {100 * e / tot:5.1f}%  stall-samples {100 * s / max(stot, 1):5.1f}%", top)
