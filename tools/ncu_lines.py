"""Per CUDA source line: share of executed warp instructions and of stall samples.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv; python tools/ncu_lines.py cs.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= max(iE, iS):
        continue
    if r[0].isdigit():  # a CUDA line: aggregated numbers of its SASS
        try:
            out.append((cur_file, int(r[0]), r[1].strip(), int(r[iE]), int(r[iS])))
        except ValueError:
            pass
tot = sum(o[3] for o in out)
stot = sum(o[4] for o in out)
print("total warp-inst", tot, "samples", stot)
agg = {}
for f, ln, src, e, s in out:
    k = (f, ln)
    a = agg.setdefault(k, [src, 0, 0])
    a[1] += e
    a[2] += s
for (f, ln), (src, e, s) in sorted(agg.items()):
    if 100 * e / tot >= thr or 100 * s / max(stot, 1) >= thr:
        print(f"{f}:{ln:<5d} inst {100 * e / tot:5.2f}%  stall {100 * s / max(stot, 1):5.2f}%  {src[:100]}")
