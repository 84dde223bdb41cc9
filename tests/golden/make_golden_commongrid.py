"""Pin compute_MVBS_index_binning (K7) to EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_commongrid.py
Writes tests/golden/commongrid_vectors.npz (committed): the index-binning cases (``ib_*``) and the bin grids that
``compute_MVBS`` builds around the flox group-by (``grid_*``, see ``mvbs_grid_cases``).

``compute_MVBS_index_binning`` (commongrid/api.py:195-266) is pure xarray (``coarsen(boundary="pad")`` mean / min): it is
lifted with ``ast`` and executed UNMODIFIED over tests/golden/xrlite.py (the decorator, attribute and provenance helpers
are replaced by no-ops: they do not touch values).  The outputs include the COORDINATES the coarsening produces:
``ping_time`` is the mean time of the pings of a tile (coord_func="mean", NaT padding skipped), not its first ping.
One statement cannot be reproduced faithfully: the function re-labels ``range_sample`` to 0..n-1 (api.py:226-230) BEFORE it
assigns the coarsened ``echo_range`` (:231-237), whose ``range_sample`` labels are still the tile means (4.5, 14.5, ...).
``Dataset.__setitem__`` of xarray aligns the new variable to the dataset's indexes, which xrlite refuses to emulate
(NotImplementedError).  The generator therefore records the right-hand side of that assignment - the coarsened minimum the
code comment describes ("binned echo_range (use first value in each average bin)") - as ``echo_range_coarsened``; what
xarray's alignment then makes of it is documented as a caveat in INTEGRATION.md, not asserted.
``compute_MVBS`` / ``compute_NASC`` group through flox, a third-party dependency that is absent here: they stay pinned
by the reference's brute-force known-answer tests restated in tests/test_oracle_golden.py.
Nothing from the reference is copied into this repository - only numeric inputs / outputs are stored.
"""

import ast
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/echopype"
sys.path.insert(0, HERE)

import xrlite  # noqa: E402

DIMS3 = ("channel", "ping_time", "range_sample")


def reference_index_binning():
    ns = {"np": np, "xr": xrlite, "_set_MVBS_attrs": lambda ds: None, "echopype_prov_attrs": lambda **k: {},
          "insert_input_processing_level": lambda ds, input_ds=None: ds}
    tree = ast.parse(open(os.path.join(REF, "commongrid/api.py")).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "compute_MVBS_index_binning":
            node.returns = None
            node.decorator_list = []
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), "commongrid/api.py", "exec"), ns)
    return ns["compute_MVBS_index_binning"]


UNALIGNED = {}
_orig_setitem = xrlite.Dataset.__setitem__


def _recording_setitem(self, k, v):
    try:
        _orig_setitem(self, k, v)
    except NotImplementedError:
        UNALIGNED[k] = v  # the value as computed, before xarray's index alignment


def volume(seed, C, P, R, dt_s):
    rng = np.random.default_rng(seed)
    Sv = (-75.0 + 8.0 * rng.standard_normal((C, P, R))).astype(np.float32).astype(np.float64)
    Sv[rng.random((C, P, R)) < 0.03] = np.nan
    er = np.broadcast_to((0.19 * (1 + 0.5 * np.arange(C)))[:, None, None] * np.arange(R)[None, None, :], (C, P, R)).astype(np.float32).astype(np.float64)
    Sv[0, P // 2, R // 3:] = np.nan  # a NaN-padded ping
    er[0, P // 2, R // 3:] = np.nan
    Sv[C - 1, 1, :] = np.nan  # a dropped ping
    er[C - 1, 1, :] = np.nan
    # irregular ping times (the tile mean differs from the tile's first ping and from its midpoint)
    t = np.datetime64("2024-03-01T12:00:00", "ns") + np.cumsum(rng.integers(1, 4, P) * int(dt_s * 1e9)).astype("timedelta64[ns]")
    return Sv, er, t


def main():
    warnings.simplefilter("ignore", RuntimeWarning)
    fn = reference_index_binning()
    out = {}
    for key, (seed, C, P, R, dt, rsn, pn) in {
        "ib_a": (1, 2, 23, 37, 1.0, 10, 5),
        "ib_b": (2, 3, 16, 64, 0.5, 8, 4),      # exact tiles
        "ib_c": (3, 1, 7, 20, 2.0, 100, 100),   # the defaults on a volume smaller than one tile
        "ib_d": (4, 2, 31, 50, 1.0, 1, 3),      # no range averaging
    }.items():
        Sv, er, t = volume(seed, C, P, R, dt)
        ds = xrlite.Dataset(coords={"channel": ("channel", np.array([f"ch{c}" for c in range(C)])), "ping_time": ("ping_time", t),
                                    "range_sample": ("range_sample", np.arange(R))})
        ds["Sv"] = (DIMS3, Sv.copy())
        ds["echo_range"] = (DIMS3, er.copy())
        ds["frequency_nominal"] = (("channel",), 38e3 * (1 + np.arange(C)))
        UNALIGNED.clear()
        xrlite.Dataset.__setitem__ = _recording_setitem
        try:
            res = fn(ds, range_sample_num=rsn, ping_num=pn)
        finally:
            xrlite.Dataset.__setitem__ = _orig_setitem
        er_out = UNALIGNED["echo_range"] if "echo_range" in UNALIGNED else res["echo_range"]
        out[f"{key}__echo_range_was_aligned_by_label"] = np.array("echo_range" not in UNALIGNED)
        out[f"{key}__Sv_in"], out[f"{key}__echo_range_in"], out[f"{key}__ping_time_in"] = Sv.astype(np.float32), er.astype(np.float32), t.astype("datetime64[ns]").astype(np.int64)
        out[f"{key}__args"] = np.array([rsn, pn])
        out[f"{key}__Sv"] = np.asarray(res["Sv"].transpose(*DIMS3).values)
        out[f"{key}__echo_range_coarsened"] = np.asarray(er_out.transpose(*DIMS3).values)
        out[f"{key}__echo_range_coarsened_range_sample_labels"] = np.asarray(er_out["range_sample"].values, dtype=np.float64)
        out[f"{key}__ping_time"] = np.asarray(res["ping_time"].values).astype("datetime64[ns]").astype(np.int64)
        out[f"{key}__range_sample"] = np.asarray(res["range_sample"].values)
        out[f"{key}__actual_range"] = np.asarray(res["Sv"].attrs["actual_range"], dtype=np.float64)
        print(key, out[f"{key}__Sv"].shape, out[f"{key}__actual_range"], out[f"{key}__ping_time"][:2] - out[f"{key}__ping_time_in"][0])
    np.savez_compressed(os.path.join(HERE, "commongrid_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


# ---- compute_MVBS: everything AROUND the flox group-by, executed unmodified --------------------------------------------------
def mvbs_grid_cases():
    """compute_MVBS (commongrid/api.py:31-191) with compute_raw_MVBS replaced by a recorder: _setup_and_validate, the range
    grid (np.arange(0, max + bin, bin) from range_var.max(skipna=True) or from range_var_max + 1e-8), the ping grid (pandas
    resample index plus one closing edge), _convert_bins_to_interval_index(closed=...), the output coordinates (interval
    left ends) and the attribute strings all run as they are; only the group-by itself (flox) does not."""
    import pandas as pd

    captured = {}

    def fake_raw_MVBS(ds_Sv, range_interval, ping_interval, range_var="echo_range", **kw):
        captured["range_interval"], captured["ping_interval"], captured["kw"] = range_interval, ping_interval, kw
        C = ds_Sv["Sv"].shape[0]
        out = xrlite.Dataset(coords={"channel": ("channel", np.asarray(ds_Sv["channel"].values)),
                                     "ping_time_bins": ("ping_time_bins", np.array(list(ping_interval), dtype=object)),
                                     f"{range_var}_bins": (f"{range_var}_bins", np.array(list(range_interval), dtype=object))})
        out["Sv"] = (("channel", "ping_time_bins", f"{range_var}_bins"), np.zeros((C, len(ping_interval), len(range_interval))))
        return out

    ns = {"np": np, "pd": pd, "xr": xrlite, "re": __import__("re"), "compute_raw_MVBS": fake_raw_MVBS, "echopype_prov_attrs": lambda **k: {},
          "insert_input_processing_level": lambda ds, input_ds=None: ds, "Literal": None, "Union": None, "Optional": None, "List": None,
          "POSITION_VARIABLES": ["latitude", "longitude"], "xarray_reduce": None}
    utree = ast.parse(open(os.path.join(REF, "commongrid/utils.py")).read())
    want = {"_setup_and_validate", "_parse_x_bin", "_convert_bins_to_interval_index", "_get_reduced_positions", "_set_MVBS_attrs", "_set_var_attrs",
            "ping_time_bin_parsing_and_conversion"}
    for node in utree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), "commongrid/utils.py", "exec"), ns)
    tree = ast.parse(open(os.path.join(REF, "commongrid/api.py")).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "compute_MVBS":
            node.returns = None
            node.decorator_list = []
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), "commongrid/api.py", "exec"), ns)
    fn = ns["compute_MVBS"]

    rng = np.random.default_rng(5)
    out = {}
    settings = {
        # key: (t0, ping intervals [s], R, dz, kwargs)
        "grid_a": ("2024-03-01T23:59:31", rng.integers(1, 4, 60), 50, 0.9, dict(range_bin="20m", ping_time_bin="20s")),
        "grid_b": ("2024-03-01T00:00:00", np.full(40, 0.5), 33, 0.31, dict(range_bin="2.5m", ping_time_bin="5s", closed="right")),
        "grid_c": ("2024-03-01T12:00:07.25", rng.integers(1, 90, 30), 20, 1.0, dict(range_bin="10m", ping_time_bin="1min", range_var_max="30m")),
        "grid_d": ("2024-03-01T06:30:00", np.full(25, 7.0), 64, 0.19, dict(range_bin="5m", ping_time_bin="2h", range_var="depth")),
        "grid_e": ("2024-03-01T00:00:09.9", np.full(12, 1.0), 16, 0.5, dict(range_bin="0.5m", ping_time_bin="500ms")),
        "grid_f": ("2024-03-01T10:10:10", np.full(10, 3600.0), 10, 2.0, dict(range_bin="20m", ping_time_bin="1D", range_var_max="18m", closed="right")),
    }
    for key, (t0, dts, R, dz, kw) in settings.items():
        P = len(dts)
        t = np.datetime64(t0, "ns") + np.cumsum(np.asarray(dts, dtype=np.float64) * 1e9).astype(np.int64).astype("timedelta64[ns]")
        C = 2
        rv = kw.get("range_var", "echo_range")
        er = (dz * np.arange(R)[None, None, :] * np.array([1.0, 1.07])[:, None, None] * np.ones((1, P, 1))).astype(np.float32).astype(np.float64)
        er[0, P // 2, R // 2:] = np.nan
        ds = xrlite.Dataset(coords={"channel": ("channel", np.array(["ch0", "ch1"])), "ping_time": ("ping_time", t), "range_sample": ("range_sample", np.arange(R))})
        ds["Sv"] = (DIMS3, np.zeros((C, P, R)))
        ds[rv] = (DIMS3, er)
        ds["frequency_nominal"] = (("channel",), np.array([38e3, 120e3]))
        res = fn(ds, **kw)
        ri, pi = captured["range_interval"], captured["ping_interval"]
        out[f"{key}__ping_time_in"] = t.astype(np.int64)
        out[f"{key}__range_in"] = er.astype(np.float32)
        out[f"{key}__kw"] = np.array(repr(kw))
        out[f"{key}__range_edges"] = np.append(np.asarray(ri.left, dtype=np.float64), float(ri.right[-1]))
        out[f"{key}__ping_edges"] = np.append(np.asarray(pi.left.values, dtype="datetime64[ns]").astype(np.int64), np.datetime64(pi.right[-1], "ns").astype(np.int64))
        out[f"{key}__closed"] = np.array([str(ri.closed), str(pi.closed)])
        out[f"{key}__out_ping_time"] = np.asarray(res["ping_time"].values).astype("datetime64[ns]").astype(np.int64)
        out[f"{key}__out_range"] = np.asarray(res[rv].values, dtype=np.float64)
        a = res["Sv"].attrs
        out[f"{key}__cell_methods"] = np.array(a["cell_methods"])
        out[f"{key}__range_meter_interval"], out[f"{key}__ping_time_interval"] = np.array(a["range_meter_interval"]), np.array(a["ping_time_interval"])
        print(key, "range edges", len(out[f"{key}__range_edges"]), "ping edges", len(out[f"{key}__ping_edges"]), str(ri.closed), "|", a["cell_methods"][:70])
    # what the argument checks raise
    bad = []
    ds_ok = ds
    for label, kw in [("range_var", dict(range_var="range")), ("missing_depth", dict(range_var="depth")), ("range_bin_type", dict(range_bin=10)),
                      ("range_bin_unit", dict(range_bin="10km")), ("range_bin_nounit", dict(range_bin="10")), ("closed", dict(closed="both")),
                      ("ping_time_bin_type", dict(ping_time_bin=10)), ("reindex", dict(method="cohorts", reindex=True))]:
        try:
            fn(ds_ok if label != "missing_depth" else ds_ok.drop_vars("depth", errors="ignore"), **kw)
            bad.append([label, "ok", ""])
        except Exception as e:  # noqa
            bad.append([label, type(e).__name__, str(e)])
    out["grid_bad__cases"] = np.array(bad)
    return out


if __name__ == "__main__":
    main()
    extra = mvbs_grid_cases()
    path = os.path.join(HERE, "commongrid_vectors.npz")
    merged = dict(np.load(path))
    merged.update(extra)
    np.savez_compressed(path, **merged)
    print("added", len(extra), "grid arrays")
