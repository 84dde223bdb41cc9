"""Generate golden vectors by EXECUTING the reference's own functions from /root/reference.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/reference_vectors.npz and tests/golden/reference_cases.json (committed).

The reference package cannot be imported (xarray, dask, flox, ... are not installed), so the
numpy/scipy-only functions are lifted out of their modules with ``ast`` and executed unmodified in
a namespace that provides numpy, scipy.signal, re, pandas and a 3-line ``xr.DataArray`` stand-in
that only records its arguments.  Nothing from the reference is copied into this repository - only
the numeric outputs are stored.
"""

import ast
import json
import os
import re
from collections import defaultdict
from functools import partial

import numpy as np
import pandas as pd
from scipy import signal

REF = "/root/reference/echopype"
HERE = os.path.dirname(os.path.abspath(__file__))


class _DA:
    def __init__(self, data=None, coords=None, dims=None, **kw):
        self.data = np.asarray(data)
        self.coords = coords


class _XR:
    DataArray = _DA
    Dataset = dict


def lift(path, names, extra=None):
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    ns = {"np": np, "signal": signal, "re": re, "pd": pd, "xr": _XR, "partial": partial,
          "defaultdict": defaultdict, "Dict": dict, "Union": None, "Literal": None, "Optional": None}
    ns.update(extra or {})
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(ast.fix_missing_locations(mod), path, "exec"), ns)
    return ns


def main():
    out = {}
    cases = {}

    # ---- utils/uwa.py (numpy only) -------------------------------------------------------------
    uwa = lift("utils/uwa.py", {"calc_sound_speed", "calc_absorption"})
    T = np.array([-1.0, 4.0, 10.0, 19.5, 20.0, 27.0])
    S = np.array([0.5, 27.9, 30.0, 33.0, 35.0, 38.0])
    P = np.array([0.0, 10.0, 59.0, 100.0, 1000.0, 3500.0])
    f = np.array([18e3, 38e3, 70e3, 120e3, 200e3, 333e3, 455e3, 769e3])
    out["uwa_T"], out["uwa_S"], out["uwa_P"], out["uwa_f"] = T, S, P, f
    out["uwa_c_mackenzie"] = np.array([[[uwa["calc_sound_speed"](t, s, p, "Mackenzie") for p in P] for s in S] for t in T])
    out["uwa_c_azfp"] = np.array([[[uwa["calc_sound_speed"](t, s, p, "AZFP") for p in P] for s in S] for t in T])
    for src in ("AM", "FG", "AZFP"):
        out[f"uwa_abs_{src}"] = np.array(
            [[[[uwa["calc_absorption"](ff, t, s, p, 8.1, None, src) for ff in f] for p in P] for s in S] for t in T]
        )
    out["uwa_abs_FG_c1500_pH78"] = np.array([uwa["calc_absorption"](ff, 8.0, 33.0, 50.0, 7.8, 1500.0, "FG") for ff in f])

    # ---- utils/compute.py ----------------------------------------------------------------------
    cmp_ = lift("utils/compute.py", {"_log2lin", "_lin2log"})
    x = np.linspace(-160, 10, 41)
    out["dB_x"], out["log2lin"] = x, cmp_["_log2lin"](x)
    out["lin2log"] = cmp_["_lin2log"](cmp_["_log2lin"](x))

    # ---- clean/utils.py extract_dB ; commongrid/utils.py _parse_x_bin, ping_time_bin parsing ----
    cu = lift("clean/utils.py", {"extract_dB"})
    cases["extract_dB"] = {s: cu["extract_dB"](s) for s in ["3.0dB", "-125dB", "+7.5db", "0dB", "10.dB", "3DB"]}
    bad = {}
    for s in ["3.0", "dB", "3.0 dB", "abc", ".5dB", "3.0dBm"]:
        try:
            cu["extract_dB"](s)
            bad[s] = "ok"
        except Exception as e:  # noqa
            bad[s] = type(e).__name__
    cases["extract_dB_errors"] = bad
    gu = lift("commongrid/utils.py", {"_parse_x_bin", "ping_time_bin_parsing_and_conversion"})
    cases["parse_x_bin"] = {
        "range_bin": {s: gu["_parse_x_bin"](s, "range_bin") for s in ["10m", "0.2m", "20 m", "5M", " 7.5m ", "1,5m".replace(",", ".")]},
        "dist_bin": {s: gu["_parse_x_bin"](s, "dist_bin") for s in ["0.5nmi", "2nmi", "1 NMI"]},
    }
    perr = {}
    for s, lab in [("10km", "range_bin"), ("10", "range_bin"), ("10m", "invalid_label"), ("5m", "dist_bin")]:
        try:
            gu["_parse_x_bin"](s, lab)
            perr[f"{s}|{lab}"] = "ok"
        except Exception as e:  # noqa
            perr[f"{s}|{lab}"] = [type(e).__name__, str(e)]
    cases["parse_x_bin_errors"] = perr
    cases["ping_time_bin"] = {s: list(gu["ping_time_bin_parsing_and_conversion"](s)) for s in ["20s", "1min", "2h", "500ms", "1D", "90s"]}

    # ---- calibrate/ek80_complex.py : replica, filter/decimate, tau_eff, norm, per-channel conv ---
    ek = lift(
        "calibrate/ek80_complex.py",
        {"tapered_chirp", "filter_decimate_chirp", "get_tau_effective", "get_norm_fac", "_convolve_per_channel"},
    )
    rng = np.random.default_rng(80)
    fs = 1.5e6
    chirp_cases = [
        # (tau, slope, f0, f1, drop_last)
        (2.048e-3, 0.05, 34e3, 45e3, False),
        (2.048e-3, 0.05, 34e3, 45e3, True),
        (1.024e-3, 0.0117, 90e3, 170e3, False),
        (0.512e-3, 0.2, 38e3, 38e3, False),  # CW tone
    ]
    wbt = (rng.standard_normal(64) + 1j * rng.standard_normal(64)) / 16
    pcf = (rng.standard_normal(32) + 1j * rng.standard_normal(32)) / 8
    out["ek80_wbt_fil"], out["ek80_pc_fil"] = wbt, pcf
    ytx = {}
    for i, (tau, slope, f0, f1, drop) in enumerate(chirp_cases):
        y, t = ek["tapered_chirp"](fs, np.array([tau]), np.array([slope]), np.array([f0]), np.array([f1]), drop)
        out[f"chirp{i}_y"], out[f"chirp{i}_t"] = y, t
        coeff = {"wbt_fil": wbt, "wbt_decifac": 6, "pc_fil": pcf, "pc_decifac": 2}
        yd, td = ek["filter_decimate_chirp"](coeff, y, fs)
        out[f"chirp{i}_ydeci"], out[f"chirp{i}_tdeci"] = yd, td
        ytx[f"ch{i}"] = yd
    cases["chirp_cases"] = [list(map(float, c[:4])) + [bool(c[4])] for c in chirp_cases]
    fs_deci = {k: fs / 12 for k in ytx}
    for mode in ("BB", "CW"):
        te = ek["get_tau_effective"](ytx, fs_deci, mode, channel=list(ytx), ping_time=None)
        out[f"tau_eff_{mode}"] = te.data
    out["norm_fac"] = ek["get_norm_fac"](ytx).data

    # per-channel convolution of a (range_sample, channel) slab, as compress_pulse drives it
    class _Ch:
        def __init__(self, v):
            self.values = v

    R, chans = 257, ["ch0", "ch2"]
    slab = rng.standard_normal((R, 2)) + 1j * rng.standard_normal((R, 2))
    slab[200:, :] = 0.0  # NaN-padding already zeroed
    replica = {c: np.flipud(np.conj(ytx[c])) for c in chans}
    out["conv_slab"] = slab
    out["conv_out"] = ek["_convolve_per_channel"](slab, replica, [_Ch(c) for c in chans])
    out["conv_zero_out"] = ek["_convolve_per_channel"](np.zeros((8, 2), complex), replica, [_Ch(c) for c in chans])
    cases["conv_channels"] = chans

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_cases.json"), "w") as fh:
        json.dump(cases, fh, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", {k: (v.shape if hasattr(v, "shape") else None) for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
