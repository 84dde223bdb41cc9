"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of one device-resident bench step
(between two rows_ek_power_kernel launches) with their share, then all captured launches by kernel.
Usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
iK, iV, iU = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
SC = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
L = []
for r in rows[1:]:
    name = r[iK].split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()
    L.append((name, float(r[iV].replace(",", "")) * SC.get(r[iU], 1e-3)))
starts = [i for i, (n, _) in enumerate(L) if n == "rows_ek_power_kernel"]
print("# per-launch device time (cold-cache, serialised: compare SHARES, not absolutes).  First: one device-resident step")
steps = [(a, b) for a, b in zip(starts, starts[1:]) if any(n.startswith("pipeline_fast_kernel") for n, _ in L[a:b]) and b - a < 12]
if steps:
    a, b = steps[len(steps) // 2]
    tot = sum(t for _, t in L[a:b])
    for n, t in L[a:b]:
        print(f"{n:62s} {t:9.1f} us  {100 * t / tot:5.1f}%")
    print(f"{'step total':62s} {tot:9.1f} us   ({len(steps)} resident steps captured)")
agg = collections.OrderedDict()
for n, t in L:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# all {len(L)} captured launches, by kernel")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{100 * t / tot:6.2f}%  n={c:4d}  avg {t / c:9.1f} us  {n[:110]}")
