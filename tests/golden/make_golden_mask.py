"""Pin mask.frequency_differencing to EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_mask.py
Writes tests/golden/freqdiff_vectors.npz (committed).

``frequency_differencing`` (mask/api.py:467-676) and its helpers ``_parse_freq_diff_eq`` / ``_check_freq_diff_source_Sv``
(mask/freq_diff.py) are lifted with ``ast`` and executed unmodified over tests/golden/xrlite.py; ``validate_source``
(fsspec path handling) is the identity for an in-memory Dataset, as in the reference.  Only numeric inputs / outputs are
stored.  ``apply_mask`` is a ``where`` over broadcast masks (mask/api.py:395-438); its argument checks need the whole
mask module (flox, dask, the seafloor / shoal detectors) and stay restated.
"""

import ast
import datetime
import operator as op
import os
import re
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/echopype"
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
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    assert all(n in ns for n in names), [n for n in names if n not in ns]


def reference_frequency_differencing():
    dask = types.ModuleType("dask")
    dask.array = types.ModuleType("dask.array")
    dask.array.Array = type("Array", (), {})
    ns = {"np": np, "xr": xrlite, "re": re, "op": op, "dask": dask, "datetime": datetime, "sys": sys,
          "validate_source": lambda ds, storage_options: (ds, None), "List": None, "Optional": None, "Union": None}
    _lift("mask/freq_diff.py", {"_parse_freq_diff_eq", "_check_freq_diff_source_Sv"}, ns)
    _lift("mask/api.py", {"str2ops", "frequency_differencing"}, ns)
    return ns["frequency_differencing"]


def reference_apply_mask():
    """apply_mask and its four helpers (mask/api.py:41-464), executed unmodified; provenance helpers are no-ops"""
    import pathlib

    ns = {"np": np, "xr": xrlite, "pathlib": pathlib, "datetime": datetime, "sys": sys, "List": None, "Optional": None, "Union": None,
          "validate_source": lambda obj, storage_options: (obj, None), "echopype_prov_attrs": lambda **k: {},
          "insert_input_processing_level": lambda ds, input_ds=None: ds}
    _lift("mask/api.py", {"_check_mask_dim_alignment", "_validate_and_collect_mask_input", "_check_var_name_fill_value",
                          "_variable_prov_attrs", "apply_mask"}, ns)
    return ns["apply_mask"]


def main():
    warnings.simplefilter("ignore", RuntimeWarning)
    if not hasattr(xrlite.DataArray, "variable"):  # frequency_differencing looks at Sv.variable._data to tell dask from numpy
        xrlite.DataArray.variable = property(lambda self: self)
    fd = reference_frequency_differencing()
    rng = np.random.default_rng(23)
    C, P, R = 4, 9, 11
    Sv = (np.round((-70 + 4 * rng.standard_normal((C, P, R))) * 2) / 2).astype(np.float32).astype(np.float64)  # 0.5 dB grid: exact ties occur
    Sv[rng.random((C, P, R)) < 0.05] = np.nan
    Sv[1, 3, 4:] = np.nan
    freqs = np.array([18e3, 38e3, 120e3, 200e3])
    chan = np.array(["chA 18", "chB 38", "chC 120", "chD 200"])
    ds = xrlite.Dataset(coords={"channel": ("channel", chan), "ping_time": ("ping_time", np.datetime64("2024-01-01T00:00:00", "ns") + np.arange(P) * np.timedelta64(1, "s")),
                                "range_sample": ("range_sample", np.arange(R))})
    ds["Sv"] = (DIMS3, Sv)
    ds["frequency_nominal"] = (("channel",), freqs)
    out = {"Sv": Sv.astype(np.float32), "frequency_nominal": freqs, "channel": chan}
    cases = {
        "f_gt": dict(freqABEq="38000.0Hz-120000.0Hz>2.5dB"),
        "f_le": dict(freqABEq="200000Hz - 18000Hz<=1.0dB"),
        "f_eq": dict(freqABEq="18kHz-38kHz==0.0dB"),
        "f_ge": dict(freqABEq="120.0kHz - 38.0kHz >= 0.5 dB"),
        "c_lt": dict(chanABEq='"chD 200" - "chA 18" < 4.0dB'),
        "c_ge": dict(chanABEq='"chB 38"-"chC 120">=3dB'),
    }
    for key, kw in cases.items():
        m = fd(ds, storage_options={}, **kw)
        assert tuple(m.dims) == ("ping_time", "range_sample"), m.dims
        out[f"{key}__mask"] = np.asarray(m.values).astype(bool)
        out[f"{key}__kw"] = np.array(repr(kw))
        out[f"{key}__operation"] = np.array(m.attrs["history"].split("Operation: ")[1])
        print(key, int(out[f"{key}__mask"].sum()), "of", P * R, "|", out[f"{key}__operation"])
    # what the reference does with equations it does not accept (a negative right-hand side is one of them)
    bad = {}
    for i, kw in enumerate([dict(freqABEq="200000Hz - 18000Hz<=-1.0dB"), dict(chanABEq='"chB 38"-"chC 120">-3dB'), dict(freqABEq="38kHz-38kHz>1dB"),
                            dict(freqABEq="38kHz-120kHz=>1dB"), dict(freqABEq="38kHz+120kHz>1dB"), dict(), dict(freqABEq="38kHz-120kHz>1dB", chanABEq='"a"-"b">1dB'),
                            dict(freqABEq="38kHz-70kHz>1dB"), dict(chanABEq='"chB 38"-"nope">1dB')]):
        try:
            fd(ds, storage_options={}, **kw)
            bad[i] = (repr(kw), "ok", "")
        except Exception as e:  # noqa
            bad[i] = (repr(kw), type(e).__name__, str(e))
    out["bad__cases"] = np.array([list(v) for v in bad.values()])
    # ---- apply_mask: values for several mask / fill combinations, and what it rejects --------------------------------------
    am = reference_apply_mask()
    coords3 = {k: ds._coords[k].data for k in DIMS3}

    def da(a, dims, attrs=None):
        return xrlite.DataArray(np.asarray(a), {d: coords3[d] for d in dims}, dims, attrs=attrs)

    m2 = rng.random((P, R)) < 0.6
    m3 = (rng.random((C, P, R)) < 0.7).astype(np.float64)
    mfd = fd(ds, storage_options={}, chanABEq='"chA 18" - "chB 38" > 1.0dB')
    fill_arr = rng.normal(-90.0, 1.0, (P, R)).astype(np.float32).astype(np.float64)
    out["am__m2"], out["am__m3"], out["am__mfd"], out["am__fill_arr"] = m2, m3, np.asarray(mfd.values).astype(bool), fill_arr
    am_cases = {
        "am_single_nan": (mfd, {}),
        "am_list_fill": ([mfd, da(m2, DIMS3[1:]), da(m3, DIMS3)], dict(fill_value=-999.0)),
        "am_int_fill": (da(m3, DIMS3), dict(fill_value=0)),
        # an array fill needs the channel dimension in the final mask: xr.where broadcasts the plain fill array positionally
        "am_fill_array": ([da(m3, DIMS3), mfd], dict(fill_value=xrlite.DataArray(fill_arr[None], {"ping_time": coords3["ping_time"], "range_sample": coords3["range_sample"]},
                                                                                 ("channel", "ping_time", "range_sample")))),
    }
    for key, (mask, kw) in am_cases.items():
        res = am(ds, mask, var_name="Sv", storage_options_ds={}, storage_options_mask={}, **kw)
        assert tuple(res["Sv"].dims) == DIMS3
        out[f"{key}__Sv"] = np.asarray(res["Sv"].values, dtype=np.float64)
        a = res["Sv"].attrs
        out[f"{key}__long_name"], out[f"{key}__actual_range"] = np.array(a["long_name"]), np.asarray(a["actual_range"], dtype=np.float64)
        out[f"{key}__mask_type"] = np.array(a.get("mask_type", ""))
        print(key, "kept", int(np.isfinite(out[f"{key}__Sv"]).sum()), a["actual_range"], a.get("mask_type"))
    nanmask = np.ones((P, R))
    nanmask[1, 2] = np.nan
    bad_am = []
    for label, mask, kw in [
        ("transposed", da(m3.transpose(0, 2, 1), ("channel", "range_sample", "ping_time")), {}),
        ("nan", da(nanmask, DIMS3[1:]), {}),
        ("two", da(np.full((P, R), 2.0), DIMS3[1:]), {}),
        ("ints_in_list", [da(np.ones((P, R), bool), DIMS3[1:]), da(np.arange(P * R).reshape(P, R), DIMS3[1:])], {}),
        ("bad_dim_name", xrlite.DataArray(np.ones((P, R), bool), {"range_sample": coords3["range_sample"]}, ("time", "range_sample")), {}),
        ("one_dim", da(np.ones(P, bool), ("ping_time",)), {}),
        ("depth_dim", xrlite.DataArray(np.ones((P, R), bool), {"ping_time": coords3["ping_time"]}, ("ping_time", "depth")), {}),
        ("short", xrlite.DataArray(np.ones((P, R - 1), bool), {"ping_time": coords3["ping_time"]}, ("ping_time", "range_sample")), {}),
        ("no_var", da(np.ones((P, R), bool), DIMS3[1:]), dict(var_name="Sv_corrected")),
        ("fill_str", da(np.ones((P, R), bool), DIMS3[1:]), dict(fill_value="nan")),
        ("fill_shape", da(np.ones((P, R), bool), DIMS3[1:]), dict(fill_value=xrlite.DataArray(fill_arr[:, :-1], None, ("ping_time", "range_sample")))),
    ]:
        kw = dict(dict(var_name="Sv"), **kw)
        try:
            am(ds, mask, storage_options_ds={}, storage_options_mask={}, **kw)
            bad_am.append([label, "ok", ""])
        except Exception as e:  # noqa
            bad_am.append([label, type(e).__name__, str(e)])
    out["am_bad__cases"] = np.array(bad_am)
    np.savez_compressed(os.path.join(HERE, "freqdiff_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
