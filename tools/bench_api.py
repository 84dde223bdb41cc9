"""Wall-clock (CUDA-synchronised) time of every public call of the path on a device-resident synthetic volume
(default: cfg2, 4 x 100 000 x 4096).  Finds host-side or kernel-side outliers that the per-kernel benchmarks do not cover.
Usage: python tools/bench_api.py [--C 4 --P 100000 --R 4096] [--only name,...]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import echopype_b200 as ep  # noqa: E402
from echopype_b200 import synth  # noqa: E402


def timed(name, fn, n, reps=2):
    best, out = None, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(json.dumps({"call": name, "ms": round(best * 1e3, 2), "Gsamples_s": round(n / best / 1e9, 2)}), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=4)
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--R", type=int, default=4096)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    only = set(a.only.split(",")) if a.only else None
    n = a.C * a.P * a.R
    ed = synth.make_ek60(C=a.C, P=a.P, R=a.R, device=True)

    def want(k):
        return only is None or k in only

    ds = timed("calibrate.compute_Sv", lambda: ep.calibrate.compute_Sv(ed), n)
    if want("TS"):
        timed("calibrate.compute_TS", lambda: ep.calibrate.compute_TS(ed), n)
    if want("noise"):
        timed("clean.estimate_background_noise(5, 30)", lambda: ep.clean.estimate_background_noise(ds, 5, 30), n)
        den = timed("clean.remove_background_noise(5, 30)", lambda: ep.clean.remove_background_noise(ds, 5, 30), n)
        del den
    if want("mvbs"):
        timed("commongrid.compute_MVBS(20m, 20s)", lambda: ep.commongrid.compute_MVBS(ds, range_bin="20m", ping_time_bin="20s"), n)
        timed("commongrid.compute_MVBS_index_binning(10, 100)", lambda: ep.commongrid.compute_MVBS_index_binning(ds, range_sample_num=100, ping_num=10), n)
    if want("fused"):
        timed("pipeline.compute_Sv_clean_MVBS", lambda: ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s"), n)
    if want("depth"):
        dsd = timed("consolidate.add_depth(depth_offset=5)", lambda: ep.consolidate.add_depth(ds, depth_offset=5.0), n)
        timed("commongrid.compute_MVBS(range_var=depth)", lambda: ep.commongrid.compute_MVBS(dsd, range_var="depth", range_bin="20m", ping_time_bin="20s"), n)
    else:
        dsd = None
    if want("mask"):
        f = list(ds.frequency_nominal.values[:2])
        eq = f"{f[0]}Hz-{f[1]}Hz>5dB"
        m = timed("mask.frequency_differencing", lambda: ep.mask.frequency_differencing(source_Sv=ds, freqABEq=eq), n)
        timed("mask.apply_mask", lambda: ep.mask.apply_mask(source_ds=ds, var_name="Sv", mask=m), n)
        del m
    if want("impulse"):
        timed("clean.mask_impulse_noise(index binning)", lambda: ep.clean.mask_impulse_noise(ds, "5m", 2, "10.0dB", "echo_range", use_index_binning=True), n)
        timed("clean.mask_impulse_noise(default: depth values)", lambda: ep.clean.mask_impulse_noise(ds, "5m", 2, "10.0dB", "echo_range"), n)
    if want("transient"):
        timed("clean.mask_transient_noise(index binning)", lambda: ep.clean.mask_transient_noise(ds, "nanmean", "10m", 25, "250.0m", "12.0dB", "echo_range", use_index_binning=True), n)
        timed("clean.mask_transient_noise(default: depth values)", lambda: ep.clean.mask_transient_noise(ds, "nanmean", "10m", 25, "250.0m", "12.0dB", "echo_range"), n)
    if want("attenuated"):
        timed("clean.mask_attenuated_signal(400m, 500m, 15, 8dB)", lambda: ep.clean.mask_attenuated_signal(ds, "400.0m", "500.0m", 15, "8.0dB", "echo_range"), n)


if __name__ == "__main__":
    main()
