"""Pin the noise-mask oracles (impulse, transient, attenuated signal) to EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_masks.py
Writes tests/golden/mask_vectors.npz (committed).

How: like make_golden_calibrate.py - ``tests/golden/xrlite.py`` stands in for xarray, and the reference's own
functions are lifted with ``ast`` out of /root/reference/echopype/clean/api.py (mask_transient_noise,
mask_impulse_noise, mask_attenuated_signal) and clean/utils.py (pool_Sv, index_binning_pool_Sv,
index_binning_downsample_upsample_along_depth, echopy_impulse_noise_mask, echopy_attenuated_signal_mask, extract_dB)
and executed UNMODIFIED (type annotations dropped).  Third-party pieces that are not installed here:
  * dask_image.ndfilters.generic_filter -> scipy.ndimage.generic_filter (dask-image applies exactly that function per
    chunk with the halo the footprint needs; one chunk here), wrapped so that ``.compute()`` works;
  * flox (downsample_upsample_along_depth, the use_index_binning=False path of mask_impulse_noise) is NOT emulated: that
    path stays a restatement in oracle/clean.py, pinned only through echopy_impulse_noise_mask.
Nothing from the reference is copied into this repository - only numeric inputs / outputs are stored.
"""

import ast
import os
import re
import sys
import types
import warnings
from functools import partial

import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/echopype"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import xrlite  # noqa: E402

DIMS3 = ("channel", "ping_time", "range_sample")


def _lift(path, names, ns):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.returns = None
            node.decorator_list = []
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


class _Lazy:
    def __init__(self, a):
        self._a = a

    def compute(self):
        return self._a


def reference_functions():
    dask_image = types.ModuleType("dask_image")
    dask_image.ndfilters = types.ModuleType("dask_image.ndfilters")
    dask_image.ndfilters.generic_filter = lambda a, function, size, mode: _Lazy(
        ndimage.generic_filter(np.asarray(a), function=function, size=size, mode=mode))

    class _Logger:
        def warning(self, *a, **k):
            pass

        info = warning

    ns = {"np": np, "xr": xrlite, "re": re, "partial": partial, "dask_image": dask_image, "logger": _Logger(), "Callable": None}
    _lift("utils/compute.py", {"_log2lin", "_lin2log"}, ns)
    _lift("commongrid/utils.py", {"_parse_x_bin"}, ns)
    _lift("clean/utils.py", {"extract_dB", "pool_Sv", "index_binning_pool_Sv", "index_binning_downsample_upsample_along_depth",
                             "echopy_impulse_noise_mask", "echopy_attenuated_signal_mask"}, ns)
    ns["downsample_upsample_along_depth"] = None  # flox path: not executed here
    _lift("clean/api.py", {"mask_transient_noise", "mask_impulse_noise", "mask_attenuated_signal"}, ns)
    return ns


def synthetic_volume(seed, C, P, R, dz, nan_frac=0.02, nan_tail_ping=None, offsets=True):
    """Sv (float32 values held as float64), depth (C, P, R): depth increases along range_sample with a per-channel sample
    spacing and a small per-ping offset (a heaving platform), a scattering layer, a few NaN samples, optionally one
    NaN-padded ping."""
    rng = np.random.default_rng(seed)
    n = np.arange(R)
    depth = np.empty((C, P, R))
    for c in range(C):
        off = (0.3 * np.sin(np.arange(P) * 0.7 + c))[:, None] if offsets else 0.0
        depth[c] = 2.0 + off + (dz * (1.0 + 0.25 * c)) * n[None, :]
    depth = depth.astype(np.float32).astype(np.float64)
    Sv = -70.0 + 6.0 * rng.standard_normal((C, P, R)) - 0.02 * depth
    Sv += 12.0 * np.exp(-0.5 * ((depth - 0.6 * depth.max()) / (0.08 * depth.max())) ** 2)  # a layer
    weak = rng.random((C, P)) < 0.25  # attenuated pings
    Sv[weak] -= 9.0
    spikes = rng.random((C, P, R)) < 0.03
    Sv[spikes] += 25.0
    Sv[rng.random((C, P, R)) < nan_frac] = np.nan
    if nan_tail_ping is not None:
        Sv[:, nan_tail_ping, R // 2:] = np.nan
    Sv = Sv.astype(np.float32).astype(np.float64)
    return Sv, depth


def to_ds(Sv, depth, range_var="depth"):
    C, P, R = Sv.shape
    ds = xrlite.Dataset(coords={"channel": ("channel", np.array([f"ch{c}" for c in range(C)])),
                                "ping_time": ("ping_time", np.datetime64("2024-01-01T00:00:00", "ns") + np.arange(P) * np.timedelta64(1, "s")),
                                "range_sample": ("range_sample", np.arange(R))})
    ds["Sv"] = (DIMS3, Sv.copy())
    ds[range_var] = (DIMS3, depth.copy())
    return ds


def canon(da):
    return np.asarray(da.transpose(*[d for d in DIMS3 if d in da.dims]).values)


def main():
    warnings.simplefilter("ignore", RuntimeWarning)
    ref = reference_functions()
    out = {}

    # ---- attenuated signal --------------------------------------------------------------------------------------
    cases_att = {
        # key: (seed, C, P, R, dz, nan_tail_ping, kwargs)
        "att_a": (11, 2, 40, 96, 0.5, None, dict(upper_limit_sl="20.0m", lower_limit_sl="35.0m", num_side_pings=4,
                                                 attenuation_signal_threshold="-3.0dB")),
        "att_b": (12, 3, 31, 64, 0.8, 7, dict(upper_limit_sl="12.5m", lower_limit_sl="40m", num_side_pings=3,
                                              attenuation_signal_threshold="-2.0dB")),
        "att_default_thr": (13, 1, 36, 80, 0.4, None, dict(upper_limit_sl="10m", lower_limit_sl="20m", num_side_pings=15,
                                                           attenuation_signal_threshold="8.0dB")),
        "att_outside": (14, 2, 12, 32, 0.5, None, dict(upper_limit_sl="400.0m", lower_limit_sl="500.0m", num_side_pings=2,
                                                       attenuation_signal_threshold="8.0dB")),
        "att_echo_range": (15, 2, 25, 48, 0.6, 3, dict(upper_limit_sl="8m", lower_limit_sl="9.9m", num_side_pings=1,
                                                       attenuation_signal_threshold="-1.0dB", range_var="echo_range")),
    }
    for key, (seed, C, P, R, dz, tail, kw) in cases_att.items():
        Sv, depth = synthetic_volume(seed, C, P, R, dz, nan_tail_ping=tail)
        if tail is not None:
            depth[:, tail, R // 2:] = np.nan  # NaN-padded range rows: np.argmin lands on the first NaN
        rv = kw.get("range_var", "depth")
        m = ref["mask_attenuated_signal"](to_ds(Sv, depth, rv), **kw)
        out[f"{key}__Sv"], out[f"{key}__range"], out[f"{key}__mask"] = Sv.astype(np.float32), depth.astype(np.float32), canon(m).astype(bool)
        out[f"{key}__kw"] = np.array(repr(kw))
        print(key, "masked pings:", int(canon(m)[:, :, 0].sum()), "of", C * P)
    # the per-channel function itself, incl. an even / odd window and a zero-width layer
    Sv, depth = synthetic_volume(21, 1, 30, 50, 0.5)
    for i, (u, l, n, t) in enumerate([(5.0, 12.0, 3, -2.5), (5.0, 12.5, 2, 0.5), (9.0, 9.1, 2, 0.0), (3.0, 26.0, 0, 1.0)]):
        out[f"att_fn{i}__mask"] = ref["echopy_attenuated_signal_mask"](Sv[0], depth[0], u, l, n, t)
        out[f"att_fn{i}__args"] = np.array([u, l, n, t])
    out["att_fn__Sv"], out["att_fn__range"] = Sv.astype(np.float32), depth.astype(np.float32)

    # ---- impulse noise: the ping comparison itself and the index-binning API path --------------------------------
    Sv, depth = synthetic_volume(31, 2, 20, 60, 0.5, offsets=False)
    out["imp__Sv"], out["imp__range"] = Sv.astype(np.float32), depth.astype(np.float32)
    for i, (k, t) in enumerate([(1, 10.0), (2, 6.0), (5, 3.0)]):
        out[f"imp_fn{i}__mask"] = ref["echopy_impulse_noise_mask"](Sv[0].T, k, t)  # (range_sample, ping_time)
        out[f"imp_fn{i}__args"] = np.array([k, t])
    for i, kw in enumerate([dict(depth_bin="2.2m", num_side_pings=2, impulse_noise_threshold="10.0dB"),
                            dict(depth_bin="3.3m", num_side_pings=1, impulse_noise_threshold="6.0dB")]):
        m = ref["mask_impulse_noise"](to_ds(Sv, depth), use_index_binning=True, **kw)
        out[f"imp_api{i}__mask"] = canon(m).astype(bool)
        out[f"imp_api{i}__kw"] = np.array(repr(kw))
        up = ref["index_binning_downsample_upsample_along_depth"](to_ds(Sv, depth), ref["_parse_x_bin"](kw["depth_bin"], "range_bin"), "depth")
        out[f"imp_api{i}__upsampled"] = canon(up)

    # ---- transient noise: depth-value windows (pool_Sv) and index windows (generic_filter), nanmean / nanmedian ---
    Sv, depth = synthetic_volume(41, 2, 14, 36, 0.5)
    out["tr__Sv"], out["tr__range"] = Sv.astype(np.float32), depth.astype(np.float32)
    for i, kw in enumerate([dict(func="nanmean", depth_bin="2m", num_side_pings=2, exclude_above="6.0m", transient_noise_threshold="6.0dB"),
                            dict(func="nanmedian", depth_bin="1.5m", num_side_pings=3, exclude_above="4.0m", transient_noise_threshold="8.0dB")]):
        ds = to_ds(Sv, depth)
        m = ref["mask_transient_noise"](ds, **kw)
        fn = np.nanmean if kw["func"] == "nanmean" else np.nanmedian
        pooled = ref["pool_Sv"](ds, fn, ref["_parse_x_bin"](kw["depth_bin"], "range_bin"), kw["num_side_pings"],
                                ref["_parse_x_bin"](kw["exclude_above"], "range_bin"), "depth")
        out[f"tr_depth{i}__mask"], out[f"tr_depth{i}__pooled"] = canon(m).astype(bool), canon(pooled)
        out[f"tr_depth{i}__kw"] = np.array(repr(kw))
    Sv, depth = synthetic_volume(42, 2, 16, 40, 0.5, offsets=False)
    out["tri__Sv"], out["tri__range"] = Sv.astype(np.float32), depth.astype(np.float32)
    for i, kw in enumerate([dict(func="nanmean", depth_bin="1.6m", num_side_pings=2, exclude_above="5.0m", transient_noise_threshold="6.0dB"),
                            dict(func="nanmedian", depth_bin="1.2m", num_side_pings=1, exclude_above="3.0m", transient_noise_threshold="8.0dB")]):
        ds = to_ds(Sv, depth)
        m = ref["mask_transient_noise"](ds, use_index_binning=True, chunk_dict={}, **kw)
        fn = np.nanmean if kw["func"] == "nanmean" else np.nanmedian
        pooled = ref["index_binning_pool_Sv"](ds, fn, ref["_parse_x_bin"](kw["depth_bin"], "range_bin"), kw["num_side_pings"],
                                              ref["_parse_x_bin"](kw["exclude_above"], "range_bin"), "depth", {})
        out[f"tr_index{i}__mask"], out[f"tr_index{i}__pooled"] = canon(m).astype(bool), canon(pooled)
        out[f"tr_index{i}__kw"] = np.array(repr(kw))

    np.savez_compressed(os.path.join(HERE, "mask_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
