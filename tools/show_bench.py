"""Print the key numbers of a bench.py JSON line (file argument or stdin)."""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.strip().splitlines() if l.startswith("{")][-1])
e = d.get("e2e", {})
print("value %.4g %s  ms/step %.4f  roofline %.3f  n_gpus %s" % (d["value"], d["unit"], d["ms_per_step"], (d.get("roofline") or {}).get("frac", float("nan")), d.get("n_gpus")))
print("e2e %.4g  ms/step %.2f  h2d_only %s" % (e.get("value", float("nan")), e.get("ms_per_step", float("nan")), json.dumps((e.get("h2d_only") or {}).get("GBps_per_rank"))))
if "e2e_raw_counts" in d:
    print("e2e_raw_counts %.4g" % d["e2e_raw_counts"]["value"])
print("config:", {k: v for k, v in d.get("config", {}).items() if k in ("verified", "e2e_matches_resident_nan_mask", "collectives_ms_per_step")})
print("clocks:", d.get("clocks"), " launches:", d.get("gpu_launches"))
