"""Summarise one kernel of an .ncu-rep: duration, DRAM traffic, IPC, stall mix, instruction mix per (thread, tile).
Usage: python tools/ncu_summary.py X.ncu-rep [units_per_launch]   (units: thread-tiles, default 80000*16)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 80000 * 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
def f(k):
    try:
        return float(d[k].replace(",", ""))
    except Exception:
        return float("nan")
print("kernel", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
units = dict(zip(rows[0], rows[1]))
SC = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
def scaled(k):
    return f(k) * SC.get(units.get(k, ""), float("nan"))
print("duration_ms %.3f  dram_read_GB %.3f dram_write_GB %.4f  dram_pct %.1f  ipc %.2f  regs %s  smem_dyn_KB %s" % (
    scaled("gpu__time_duration.sum"), scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum"),
    f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f("sm__inst_executed.avg.per_cycle_active"),
    d.get("launch__registers_per_thread"), d.get("launch__shared_mem_per_block_dynamic")))
st = {k[33:]: f(k) for k in d if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
tot = sum(v for v in st.values() if v == v)
print("stalls% " + " ".join(f"{k}:{100 * v / tot:.0f}" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v / tot > 0.02))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Source" in r and "Address" in r)
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > iE and r[iE].isdigit()]
ops = collections.Counter()
for r in data:
    t = r[iA].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += int(r[iE])
tot = sum(ops.values())
print("instr per unit %.1f:" % (tot / per), " ".join(f"{op}:{n / per:.0f}" for op, n in ops.most_common(26)))
stot = sum(int(r[iS]) for r in data)
blk = 50
for b in range(0, len(data), blk):
    e = sum(int(r[iE]) for r in data[b:b + blk])
    s_ = sum(int(r[iS]) for r in data[b:b + blk])
    if e / per > 8 or s_ / stot > 0.03:
        print(f"  sass {b:5d}: instr/unit {e / per:6.1f}  stall {100 * s_ / stot:5.1f}%")
