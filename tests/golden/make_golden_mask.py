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
    np.savez_compressed(os.path.join(HERE, "freqdiff_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
