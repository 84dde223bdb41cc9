"""Per-kernel micro-benchmarks (CUDA events on torch's current stream, which is the stream the library
launches on).  Usage: python tools/bench_kernels.py [--which sv,svr,...] [--C 4 --P 100000 --R 4096]"""

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ctypes  # noqa: E402

import torch  # noqa: E402

from echopype_b200 import _lib  # noqa: E402
from echopype_b200.device import ptr, stream  # noqa: E402

from echopype_b200 import kernels, synth  # noqa: E402
from echopype_b200.calibrate.calibrate_ek import CalibrateEK60  # noqa: E402

PEAK = 6547.8
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def report(name, ms, ms_min, nbytes, nsamp):
    print(json.dumps({"kernel": name, "ms_median": round(ms, 4), "ms_min": round(ms_min, 4), "GBps": round(nbytes / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(nbytes / ms / 1e6 / PEAK, 3), "Gsamples_s": round(nsamp / ms / 1e6, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=4)
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--R", type=int, default=4096)
    ap.add_argument("--which", default="copy,sv,svr,svrm,noise,bins,pipe")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bb-pings", type=int, default=2000)
    a = ap.parse_args()
    C, P, R = a.C, a.P, a.R
    which = a.which.split(",")
    n = C * P * R
    ed = synth.make_ek60(C, P, R, device=True)
    cal = CalibrateEK60(ed)
    prm, _ = cal._power_params("Sv")
    rows = kernels.rows_ek_power(C, P, R, kernels.SONAR_EX60, "Sv", prm, cal._is_gpt())
    x = ed["Sonar/Beam_group1"]["backscatter_r"].data
    out = torch.empty_like(x)
    rng = torch.empty_like(x)
    if "copy" in which:
        ms, mn = timeit(lambda: out.copy_(x), a.iters)
        report("torch_copy", ms, mn, 8 * n, n)
    if "sv" in which:
        ms, mn = timeit(lambda: kernels.sv_power(x, rows, C, P, R, want_range=False, out=out), a.iters)
        report("sv_power", ms, mn, 8 * n, n)
    if "svr" in which:
        ms, mn = timeit(lambda: kernels.sv_power(x, rows, C, P, R, want_range=True, out=out, rng=rng), a.iters)
        report("sv_power+range", ms, mn, 12 * n, n)
    if "svrm" in which:
        ms, mn = timeit(lambda: kernels.sv_power(x, rows, C, P, R, want_range=True, want_minmax=True, out=out, rng=rng), a.iters)
        report("sv_power+range+minmax", ms, mn, 12 * n, n)
    if {"noise", "bins", "pipe", "pipe16", "masks", "pipex"} & set(which):
        extra(a, which, C, P, R, n, x, rows, out, rng, ed)
    if "pulse" in which or "k2" in which:
        del x, out, rng, ed
        torch.cuda.empty_cache()
        if "k2" in which:
            complex_cw(a)
            torch.cuda.empty_cache()
        if "pulse" in which:
            pulse(a)


def pulse(a):
    """cfg3-shaped EK80 broadband volume (6 ch x pings x 8192 complex samples x 4 beams), K3 alone."""
    import numpy as np

    from echopype_b200.calibrate.calibrate_ek import CalibrateEK80

    C, P, R, B = 6, a.bb_pings, 8192, 4
    ed = synth.make_ek80(C=C, P=P, R=R, B=B, mode="BB", encode="complex", device=True, nan_tail=0.005, seed=3000)
    cal = CalibrateEK80(ed, waveform_mode="BB", encode_mode="complex")
    holder = {}

    def run():
        holder["ds"] = cal._cal_complex_samples("Sv")

    run()  # host-side replica synthesis etc. happen on every call; time the kernel through the C ABI instead
    beam = ed["Sonar/Beam_group1"]
    re, im = beam["backscatter_r"].data, beam["backscatter_i"].data
    chan = np.asarray(beam["channel"].values)
    tx = [cal._tx[c] for c in chan]
    M = max(len(t) for t in tx)
    out = torch.empty((C, P, R), dtype=torch.float32, device=re.device)
    rng = torch.empty_like(out)
    ms, mn = timeit(lambda: kernels.pulse_compress_sv(re, im, tx, cal.rows, C, P, R, B, want_range=True), a.iters)
    n = C * P * R
    Mpad = (M + 7) // 8 * 8
    flops = 8.0 * Mpad * n
    print(json.dumps({"kernel": "pulse_compress_sv (K3)", "taps": M, "ms_median": round(ms, 3), "ms_min": round(mn, 3),
                      "GBps": round(40 * n / ms / 1e6, 1), "frac_of_measured_hbm": round(40 * n / ms / 1e6 / PEAK, 3),
                      "TFLOPs_fp32": round(flops / ms / 1e9, 2), "frac_of_fp32_peak_74.4": round(flops / ms / 1e9 / 74.4, 3),
                      "Gsamples_s": round(n / ms / 1e6, 2)}), flush=True)
    _ = out, rng


def complex_cw(a):
    """EK80 CW complex samples (K2): 6 ch x pings x 4096 samples x 4 beams, 36 algorithmic bytes per sample (40 with range)."""
    from echopype_b200.calibrate.calibrate_ek import CalibrateEK80

    C, P, R, B = 6, 8 * a.bb_pings, 4096, 4
    ed = synth.make_ek80(C=C, P=P, R=R, B=B, mode="CW", encode="complex", device=True, nan_tail=0.005, seed=3100)
    cal = CalibrateEK80(ed, waveform_mode="CW", encode_mode="complex")
    cal._cal_complex_samples("Sv")
    beam = ed["Sonar/Beam_group1"]
    re, im = beam["backscatter_r"].data, beam["backscatter_i"].data
    n = C * P * R
    ms, mn = timeit(lambda: kernels.sv_complex(re, im, cal.rows, C, P, R, B, want_range=True), a.iters)
    report("sv_complex (K2, 4 beams, + range)", ms, mn, 40 * n, n)
    ms, mn = timeit(lambda: kernels.sv_complex(re, im, cal.rows, C, P, R, B, want_range=False), a.iters)
    report("sv_complex (K2, 4 beams)", ms, mn, 36 * n, n)


def extra(a, which, C, P, R, n, x, rows, out, rng, ed):
    import numpy as np

    from echopype_b200.commongrid.utils import assign_bins, ping_time_edges, range_edges
    from echopype_b200.device import ParamPack

    dev = x.device
    kernels.sv_power(x, rows, C, P, R, want_range=True, out=out, rng=rng)
    pack = ParamPack(C, P, dev)
    alpha = pack.cp(np.array([0.0027, 0.0098, 0.0375, 0.0527, 0.0187, 0.0809][:C]))
    if "noise" in which:
        holder = {}

        def est():
            holder["n"] = kernels.noise_estimate(out, rng, alpha, pack, C, P, R, 5, 30)

        ms, mn = timeit(est, a.iters)
        report("noise_estimate(5x30)", ms, mn, 8 * n, n)
        sn, sc = torch.empty_like(x), torch.empty_like(x)
        nz = holder["n"]

        def app():
            _lib.call("epb_noise_apply", ptr(out), ptr(rng), alpha, ptr(nz), ptr(sn), ptr(sc), None, C, P, R, 5,
                      ctypes.c_float(3.0), stream())

        ms, mn = timeit(app, a.iters)
        report("noise_apply", ms, mn, 16 * n, n)
        del sn, sc
    pt = ed["Sonar/Beam_group1"]["ping_time"].values
    rmax = kernels.range_max(x, rows, C, P, R)
    r_edges = range_edges(rmax, 20.0)
    p_edges = ping_time_edges(pt, "20s")
    xb = torch.from_numpy(assign_bins(pt, p_edges)).to(dev)
    et = torch.from_numpy(r_edges).to(dev)
    nX = len(p_edges) - 1
    acc = kernels.new_acc(C, nX, len(r_edges) - 1, dev)
    if "bins" in which:
        ms, mn = timeit(lambda: kernels.bin_reduce_law(out, rows, xb, et, acc, C, P, R, nX), a.iters)
        report("bin_reduce_law", ms, mn, 4 * n, n)
        ms, mn = timeit(lambda: kernels.bin_reduce(out, rng, xb, et, acc, C, P, R, nX), a.iters)
        report("bin_reduce_generic_f32", ms, mn, 8 * n, n)
    if "pipe" in which:
        ms, mn = timeit(lambda: kernels.pipeline_power_mvbs(x, rows, xb, et, acc, C, P, R, nX, 5, 30), a.iters)
        report("pipeline(noise 5x30 + mvbs)", ms, mn, 4 * n, n)
        ms, mn = timeit(lambda: kernels.pipeline_power_mvbs(x, rows, xb, et, acc, C, P, R, nX, 0, 0), a.iters)
        report("pipeline(sv->mvbs)", ms, mn, 4 * n, n)
    if "masks" in which:  # clean.mask_impulse_noise / mask_transient_noise (index binning); 5 algorithmic bytes per sample
        nsamp = np.full(C, 27, dtype=np.int64)
        mbuf = (torch.empty((C, P, R), dtype=torch.uint8, device=dev), torch.empty((C, P, -(-R // 27)), dtype=torch.float32, device=dev))
        ms, mn = timeit(lambda: kernels.impulse_noise_mask(out, nsamp, C, P, R, 2, 10.0, out=mbuf), max(3, a.iters // 2))
        report("mask_impulse_noise(5m, 2 pings)", ms, mn, 5 * n, n)
        kernels.sv_power(x, rows, C, P, R, want_range=True, out=out, rng=rng)
        edges = np.arange(0.0, float(rng.nan_to_num(nan=0.0).max()) + 5.0, 5.0)
        ms, mn = timeit(lambda: kernels.impulse_noise_mask_depth(out, rng, edges, C, P, R, 2, 10.0), max(3, a.iters // 2))
        report("mask_impulse_noise(5m, depth-value bins)", ms, mn, 9 * n, n)
        ms, mn = timeit(lambda: kernels.impulse_noise_mask_depth(out, rng, edges, C, P, R, 2, 10.0, want_upsampled=False), max(3, a.iters // 2))
        report("mask_impulse_noise(5m, depth-value bins, single pass: no upsampled array)", ms, mn, 9 * n, n)
        nsamp = np.full(C, 53, dtype=np.int64)
        tbuf = (mbuf[0], torch.empty((C, P, R, 2), dtype=torch.float32, device=dev))
        ms, mn = timeit(lambda: kernels.transient_noise_mask(out, nsamp, C, P, R, 1300, 25, 12.0, out=tbuf), max(3, a.iters // 2))
        report("mask_transient_noise(10m, 25 pings)", ms, mn, 5 * n, n)
        dmin, dmax = float(rng.nan_to_num(nan=1e30).min()), float(rng.nan_to_num(nan=0.0).max())
        ms, mn = timeit(lambda: kernels.transient_noise_mask_depth(out, rng, C, P, R, dmin, dmax, 10.0, 250.0, 25, 12.0), 3)
        report("mask_transient_noise(10m, 25 pings, depth-value windows)", ms, mn, 9 * n, n)
    if "pipex" in which:  # cost of the optional outputs of the fused launch
        nz = torch.empty((C, -(-P // 5)), dtype=torch.float32, device=dev)
        rm = torch.empty(1, dtype=torch.float64, device=dev)
        for name, kw in (("none", {}), ("noise_out", {"noise_out": nz}), ("rmax_out", {"rmax_out": rm}), ("both", {"noise_out": nz, "rmax_out": rm})):
            ms, mn = timeit(lambda: kernels.pipeline_power_mvbs(x, rows, xb, et, acc, C, P, R, nX, 5, 30, **kw), a.iters)
            report(f"pipeline(noise 5x30 + mvbs) [{name}]", ms, mn, 4 * n, n)
    if "pipe16" in which:  # the same chain on int16 raw power counts (2 algorithmic bytes per sample)
        q = kernels.synth_fill_i16((C, P, R), seed=1001)
        ms, mn = timeit(lambda: kernels.pipeline_power_mvbs_i16(q, out.view(-1), rows, xb, et, acc, C, P, R, nX, 5, 30), a.iters)
        report("pipeline_i16(noise 5x30 + mvbs)", ms, mn, 2 * n, n)
        ms, mn = timeit(lambda: kernels.pipeline_power_mvbs_i16(q, out.view(-1), rows, xb, et, acc, C, P, R, nX, 0, 0), a.iters)
        report("pipeline_i16(sv->mvbs)", ms, mn, 2 * n, n)
        ms, mn = timeit(lambda: kernels.ingest_power_i16(q, out=out), a.iters)
        report("ingest_power_i16", ms, mn, 6 * n, n)
        ms, mn = timeit(lambda: kernels.sv_power(q, rows, C, P, R, want_range=False, out=out), a.iters)
        report("sv_power_i16 (K1 on raw counts)", ms, mn, 6 * n, n)


if __name__ == "__main__":
    main()
