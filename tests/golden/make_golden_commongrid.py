"""Pin compute_MVBS_index_binning (K7) to EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_commongrid.py
Writes tests/golden/commongrid_vectors.npz (committed).

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


if __name__ == "__main__":
    main()
