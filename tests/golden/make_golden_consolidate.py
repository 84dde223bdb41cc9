"""Pin consolidate.add_depth to EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_consolidate.py
Writes tests/golden/consolidate_vectors.npz (committed).

With tests/golden/xrlite.py registered as xarray and the namespace skeleton of make_golden_calibrate.py, the reference's
own modules ``utils/align.py`` and ``consolidate/ek_depth_utils.py`` are IMPORTED unmodified from /root/reference, and
``add_depth`` (consolidate/api.py:66-247; its module imports the zarr / datatree based EchoData) is lifted with ``ast`` and
executed unmodified (decorator dropped, ``open_source`` = identity).  scipy (``interp1d`` behind ``DataArray.interp``,
``Rotation``) is installed and used as the reference uses it.  Only numeric inputs / outputs are stored.
"""

import ast
import datetime
import importlib
import os
import sys
import types
import warnings
from numbers import Number

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/echopype"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden_calibrate as mgc  # noqa: E402
import xrlite  # noqa: E402

DIMS3 = ("channel", "ping_time", "range_sample")


def reference_functions():
    mods, EchoData = mgc.install_reference()
    pkg = types.ModuleType("echopype.consolidate")
    pkg.__path__ = [os.path.join(REF, "consolidate")]
    sys.modules["echopype.consolidate"] = pkg
    eku = importlib.import_module("echopype.consolidate.ek_depth_utils")

    class _Logger:
        def warning(self, *a, **k):
            pass

    ns = {"np": np, "xr": xrlite, "Number": Number, "datetime": datetime, "sys": sys, "logger": _Logger(),
          "open_source": lambda obj, kind, opts: obj, "align_to_ping_time": mods["utils.align"].align_to_ping_time,
          "ek_use_platform_vertical_offsets": eku.ek_use_platform_vertical_offsets,
          "ek_use_platform_angles": eku.ek_use_platform_angles, "ek_use_beam_angles": eku.ek_use_beam_angles,
          "Union": None, "Optional": None, "pathlib": None, "EchoData": EchoData}
    tree = ast.parse(open(os.path.join(REF, "consolidate/api.py")).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "add_depth":
            node.returns = None
            node.decorator_list = []
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), "consolidate/api.py", "exec"), ns)
    return ns["add_depth"], eku, mods["utils.align"].align_to_ping_time, EchoData


def main():
    warnings.simplefilter("ignore", RuntimeWarning)
    add_depth, eku, align, EchoData = reference_functions()
    rs = np.random.default_rng(17)
    C, P, R = 3, 13, 9
    t0 = np.datetime64("2024-05-01T00:00:00", "ns")
    pt = t0 + (np.arange(P) * 10**9 + rs.integers(0, 10**8, P)).astype("timedelta64[ns]")
    er = (np.array([0.19, 0.095, 0.38])[:, None, None] * np.arange(R)[None, None, :] * np.ones((1, P, 1))).astype(np.float32).astype(np.float64)
    er[0, 4, 5:] = np.nan
    er[2, 0, :] = np.nan
    chan = np.array([f"ch{c}" for c in range(C)])
    out = {"echo_range": er, "ping_time": pt.astype(np.int64), "channel": chan}

    def ds_sv():
        ds = xrlite.Dataset(coords={"channel": ("channel", chan), "ping_time": ("ping_time", pt), "range_sample": ("range_sample", np.arange(R))})
        ds["echo_range"] = (DIMS3, er.copy())
        ds["Sv"] = (DIMS3, np.zeros((C, P, R)))
        return ds

    def series(values, times, dim="time3"):
        return xrlite.DataArray(np.asarray(values, dtype=np.float64), {dim: times}, (dim,))

    def depth_of(**kw):
        return np.asarray(add_depth(ds_sv(), **kw)["depth"].transpose(*DIMS3).values)

    # numbers / upward
    out["numbers__depth"] = depth_of(depth_offset=7.5, tilt=12.0)
    out["upward__depth"] = depth_of(depth_offset=250.0, tilt=3.0, downward=False)
    out["plain__depth"] = depth_of()
    # time series on their own clock (nearest, extrapolated), on the ping clock (rename path), and a single value
    t3 = pt[::4] + np.timedelta64(300, "ms")
    off, tl = 5 + rs.random(len(t3)), 10 * rs.random(len(t3))
    out["series__t3"], out["series__off"], out["series__tilt"] = t3.astype(np.int64), off, tl
    out["series__depth"] = depth_of(depth_offset=series(off, t3), tilt=series(tl, t3))
    out["series__off_aligned"] = np.asarray(align(series(off, t3), "time3", ds_sv()["ping_time"]).values)
    offp = 3 + rs.random(P)
    out["onping__off"] = offp
    out["onping__depth"] = depth_of(depth_offset=series(offp, pt, "time_x"))
    out["single__depth"] = depth_of(depth_offset=series([4.25], pt[3:4]), tilt=series([20.0], pt[5:6]))
    # Platform group: vertical offsets and pitch / roll
    t2 = pt[::5] - np.timedelta64(200, "ms")
    n2 = len(t2)
    plat = xrlite.Dataset(coords={"time2": ("time2", t2)})
    vals = {"water_level": rs.random(n2), "vertical_offset": rs.random(n2) - 0.5, "transducer_offset_z": 4 + rs.random(n2),
            "pitch": 6 * rs.random(n2) - 3, "roll": 8 * rs.random(n2) - 4}
    for k, v in vals.items():
        plat[k] = (("time2",), v)
        out[f"platform__{k}"] = v
    out["platform__t2"] = t2.astype(np.int64)
    beam = xrlite.Dataset(coords={"channel": ("channel", chan)})
    bx, by, bz = np.array([0.1, 0.0, 0.0]), np.array([0.0, 0.2, 0.0]), np.array([0.99, 0.97, 0.0])
    for n, v in zip("xyz", (bx, by, bz)):
        beam[f"beam_direction_{n}"] = (("channel",), v)
        out[f"beam__{n}"] = v
    sonar = xrlite.Dataset(attrs={"sonar_model": "EK60"})
    ed = EchoData("EK60", {"Platform": plat, "Sonar/Beam_group1": beam, "Sonar": sonar})
    out["platform__depth"] = depth_of(echodata=ed, use_platform_vertical_offsets=True, use_platform_angles=True)
    out["platform_offsets_only__depth"] = depth_of(echodata=ed, use_platform_vertical_offsets=True)
    out["beam__depth"] = depth_of(echodata=ed, use_beam_angles=True)
    out["platform__transducer_depth"] = np.asarray(eku.ek_use_platform_vertical_offsets(plat, ds_sv()["ping_time"]).values)
    out["platform__scaling"] = np.asarray(eku.ek_use_platform_angles(plat, ds_sv()["ping_time"]).values)
    out["beam__scaling"] = np.asarray(eku.ek_use_beam_angles(beam).values)
    np.savez_compressed(os.path.join(HERE, "consolidate_vectors.npz"), **out)
    print("wrote", len(out), "arrays;", {k: v.shape for k, v in out.items() if k.endswith("__depth")})


if __name__ == "__main__":
    main()
