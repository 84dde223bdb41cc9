"""Per-source-line cost of one kernel: joins the SASS page of an .ncu-rep (instructions executed / stall samples per
instruction) with the line table of the same kernel in the built library (nvdisasm -g), instruction by instruction.
Usage: python tools/ncu_lines.py X.ncu-rep <cubin name, e.g. pipeline_fast_f32b> <kernel substring> [units_per_launch]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, unit, kern = sys.argv[1:4]
per = float(sys.argv[4]) if len(sys.argv) > 4 else 80000 * 16
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", unit + ".sm_100a.cubin", os.path.join(root, "echopype_b200", "libepb200.so")], cwd=tmp, capture_output=True)
cub = os.path.join(tmp, unit + ".sm_100a.cubin")
dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Source" in r and "Address" in r)
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > iE and r[iE].isdigit()]
print(f"# {len(data)} profiled instructions, {len(lines)} disassembled")
if len(data) != len(lines):
    print("# WARNING: instruction counts differ (profile and library are different builds); lines are approximate")
ex, st = collections.Counter(), collections.Counter()
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else None
    ex[key] += int(r[iE])
    st[key] += int(r[iS])
tot_s = sum(st.values())
print("# file:line  instr/unit  stall%")
for key, n in sorted(ex.items(), key=lambda kv: (kv[0] or ("", 0))):
    if n / per >= 2 or st[key] / tot_s > 0.01:
        print(f"{key[0] if key else '?'}:{key[1] if key else 0:5d}  {n / per:7.1f}  {100 * st[key] / tot_s:5.1f}")
