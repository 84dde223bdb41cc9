"""SASS mnemonic counts per kernel of the built library (evidence for the Blackwell-specific paths).
Usage: python tools/sass_counts.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "echopype_b200", "libepb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
WANT = ["UBLKCP", "SYNCS", "UBLKPF", "FFMA2", "FADD2", "FMUL2", "LDGSTS", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "LDS.128", "LDG.E.128", "STG.E", "ATOMS", "REDUX", "VOTE",
        "REDG", "DADD", "F2F", "PRMT", "VIMNMX3", "SHFL", "BAR", "STL", "LDL", "UTCHMMA", "UTCQMMA", "HMMA"]
KEEP = ["pipeline_fast_kernelILi5ELi2ELb1ELb0ELb0ELb0", "pipeline_fast_kernelILi5ELi2ELb1ELb0ELb1ELb0", "pipeline_fast_kernelILi5ELi2ELb1ELb1ELb0ELb0",
        "pipeline_fast_kernelILi5ELi2ELb1ELb0ELb0ELb1", "pipeline_fast_kernelILi1ELi4ELb1ELb0ELb1ELb0", "bin_reduce_staged_kernelIfLb0", "transient_strip_kernelILb0ELb0",
        "transient_strip_kernelILb0ELb1", "impulse_fused_kernel", "impulse_mask_wide_kernel", "pulse_fft_kernelILi4ELi4", "sv_power_vec4ILb1ELb1Ef", "sv_complex_kernel",
        "noise_estimate_kernel", "coarsen_kernel", "apply_mask_kernel", "attenuated_ping_kernel", "attenuated_limits_kernel"]
print("# SASS evidence (cuobjdump -sass echopype_b200/libepb200.so; tools/sass_counts.py): counts of the mnemonics that show the")
print("# Blackwell-specific paths each kernel uses.  UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, UBLKPF = TMA L2 prefetch,")
print("# FFMA2 / FADD2 / FMUL2 = packed FP32 (two lanes per instruction), LDGSTS = cp.async, MUFU.EX2 / LG2 = single-MUFU exp2 / log2.")
print("# No UTC*MMA / HMMA / TMEM instructions anywhere: the path has no dense contraction (north_star).")
print("# architectures in the library:", archs)
cur, cnt, tot = None, collections.Counter(), 0
res = {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if cur:
            res[cur] = (tot, cnt)
        cur, cnt, tot = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        tot += 1
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + ".") or op.startswith(w):
                cnt[w] += 1
                break
if cur:
    res[cur] = (tot, cnt)
anymma = sum(c["UTCHMMA"] + c["UTCQMMA"] + c["HMMA"] for _, c in res.values())
for k in KEEP:
    for name, (t, c) in res.items():
        if k in name:
            short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d+", "", name)
            print(f"{k}: {t} SASS instructions; " + ", ".join(f"{w} {c[w]}" for w in WANT if c[w]))
            break
print(f"# tensor-core mnemonics in the whole library: {anymma}; kernels in the library: {len(res)}")
