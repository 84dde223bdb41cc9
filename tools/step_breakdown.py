"""Device-time breakdown of one FusedPlan.run() step (CUDA events around every library call, warm)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from echopype_b200 import _lib, pipeline, synth

ed = synth.make_ek60(4, 100000, 4096, seed=2000, device=True, nan_tail=0.005)
plan = pipeline.FusedPlan(ed, ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s")
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
orig = _lib.call
events = []
def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    events.append((name, e0, e1))
_lib.call = timed
import echopype_b200.kernels as K
K._lib.call = timed
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = {}
N = 10
s0.record()
for _ in range(N):
    plan.run()
s1.record()
torch.cuda.synchronize()
for n, a, b in events:
    tot[n] = tot.get(n, 0.0) + a.elapsed_time(b) / N
print(json.dumps({"step_ms": s0.elapsed_time(s1) / N, "calls_ms": {k: round(v, 4) for k, v in tot.items()}, "sum_calls_ms": round(sum(tot.values()), 4)}))
