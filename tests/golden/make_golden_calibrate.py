"""Pin the compute_Sv / compute_TS / noise-removal arithmetic to EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_calibrate.py
Writes tests/golden/calibrate_vectors.npz (committed).

How: ``tests/golden/xrlite.py`` (a small labelled-array container) is registered as ``xarray``, a
package skeleton for ``echopype`` is put into ``sys.modules`` (namespace only: ``__path__`` pointing at
/root/reference/echopype, no ``__init__`` executed, so dask / flox / zarr / netCDF4 are never
imported), and then the reference's own modules are imported UNMODIFIED from /root/reference:

    echopype/calibrate/range.py, calibrate_base.py, calibrate_ek.py, calibrate_azfp.py,
    cal_params.py, env_params.py, ek80_complex.py, ecs.py, echopype/utils/uwa.py, utils/align.py,
    utils/log.py, utils/compute.py, echodata/simrad.py, clean/utils.py

The calibration classes ``CalibrateEK60 / CalibrateEK80 / CalibrateAZFP`` are instantiated exactly as
``calibrate/api.py:64-87`` does and ``compute_Sv() / compute_TS()`` are called on small synthetic
EchoData objects (the generators of echopype_b200.synth, host arrays).  ``estimate_background_noise``
and ``remove_background_noise`` (clean/api.py:362-511) are lifted with ``ast`` (their module imports
dask-image based maskers) and executed unmodified as well.  Nothing from the reference is copied into
this repository - only numeric inputs / outputs are stored.
"""

import ast
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/echopype"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import xrlite  # noqa: E402

DIMS3 = ("channel", "ping_time", "range_sample")


def install_reference():
    """namespace skeleton + stubs; returns the imported reference modules"""
    sys.modules["xarray"] = xrlite
    if "dask" not in sys.modules:  # utils/compute.py names dask.array.Array in type annotations only
        dask, dask_array = types.ModuleType("dask"), types.ModuleType("dask.array")
        dask_array.Array = type("Array", (), {})
        dask.array = dask_array
        sys.modules["dask"], sys.modules["dask.array"] = dask, dask_array

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    pkg("echopype", REF)
    for sub in ("calibrate", "utils", "echodata", "convert", "clean", "commongrid"):
        setattr(sys.modules["echopype"], sub, pkg(f"echopype.{sub}", os.path.join(REF, sub)))

    class EchoData:  # the reference's EchoData is a zarr/datatree container; only item access is used
        def __init__(self, sonar_model, groups, source_file=None):
            self.sonar_model, self._g, self.source_file, self.converted_raw_path = sonar_model, groups, source_file, None

        def __getitem__(self, k):
            return self._g[k]

    sys.modules["echopype.echodata"].EchoData = EchoData
    ed_mod = types.ModuleType("echopype.echodata.echodata")
    ed_mod.EchoData = EchoData
    sys.modules["echopype.echodata.echodata"] = ed_mod

    # three string constants of convert/set_groups_ek80.py (read from the file, which imports the parsers)
    consts = types.ModuleType("echopype.convert.set_groups_ek80")
    tree = ast.parse(open(os.path.join(REF, "convert/set_groups_ek80.py")).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Constant) and node.targets[0].id in ("DECIMATION", "FILTER_IMAG", "FILTER_REAL"):
            setattr(consts, node.targets[0].id, node.value.value)
    sys.modules["echopype.convert.set_groups_ek80"] = consts

    mods = {}
    for name in ("utils.log", "utils.uwa", "utils.align", "utils.compute", "echodata.simrad", "calibrate.cal_params",
                 "calibrate.env_params", "calibrate.range", "calibrate.ecs", "calibrate.ek80_complex",
                 "calibrate.calibrate_base", "calibrate.calibrate_ek", "calibrate.calibrate_azfp"):
        mods[name] = importlib.import_module("echopype." + name)
    import logging

    logging.getLogger("echopype").setLevel(logging.ERROR)
    return mods, EchoData


def lift_noise(mods):
    """estimate_background_noise / remove_background_noise from clean/api.py, executed unmodified (decorator,
    attrs and provenance helpers replaced by no-ops: they do not touch values)."""
    src = open(os.path.join(REF, "clean/api.py")).read()
    tree = ast.parse(src)
    cu_tree = ast.parse(open(os.path.join(REF, "clean/utils.py")).read())
    ns = {"np": np, "xr": xrlite, "_log2lin": mods["utils.compute"]._log2lin, "_lin2log": mods["utils.compute"]._lin2log,
          "add_remove_background_noise_attrs": lambda da, *a, **k: da, "echopype_prov_attrs": lambda **k: {},
          "insert_input_processing_level": lambda ds, input_ds=None: ds}
    import re

    ns["re"] = re
    for node in cu_tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "extract_dB":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "clean/utils.py", "exec"), ns)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("estimate_background_noise", "remove_background_noise"):
            node.decorator_list = []
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), "clean/api.py", "exec"), ns)
    return ns["estimate_background_noise"], ns["remove_background_noise"]


def to_xr(ds):
    """echopype_b200.dataset.Dataset (host arrays) -> xrlite.Dataset"""
    out = xrlite.Dataset()
    for k, v in ds.coords.items():
        out._coords[k] = xrlite._Coord(tuple(v.dims), np.asarray(v.values))
    for name in ds.data_vars:
        da = ds[name]
        out[name] = (tuple(da.dims), np.asarray(da.values))
    return out


def to_ref_echodata(ed, EchoData):
    return EchoData(ed.sonar_model, {g: to_xr(ed[g]) for g in ed.group_paths}, source_file="synthetic")


def canon(da, dims=DIMS3):
    """values in canonical dim order, broadcast to the dims the variable actually has"""
    have = [d for d in dims if d in da.dims]
    extra = [d for d in da.dims if d not in dims]
    assert not extra, (da.name, da.dims)
    return np.asarray(da.transpose(*have).values), have


def dump_case(out, key, ds, cal_type, inputs):
    v, dims = canon(ds[cal_type])
    assert dims == list(DIMS3)
    out[f"{key}/{cal_type}"] = v
    er = canon(ds["echo_range"])[0]
    if f"{key}/echo_range" not in out:
        out[f"{key}/echo_range"] = er
    elif not np.array_equal(out[f"{key}/echo_range"], er, equal_nan=True):  # AZFP: the TS range differs (range.py:77-80)
        out[f"{key}/echo_range_{cal_type}"] = er
    for p in ("sound_speed", "sound_absorption", "gain_correction", "sa_correction", "equivalent_beam_angle", "tau_effective",
              "temperature", "salinity", "pressure", "pH", "impedance_transceiver", "impedance_transducer",
              "receiver_sampling_frequency", "angle_offset_alongship", "beamwidth_alongship", "EL", "DS", "TVR", "VTX0", "Sv_offset"):
        if p in ds:
            val = ds[p]
            if val.dtype.kind in "fiu":
                arr, d = canon(val, ("channel", "ping_time"))
                out[f"{key}/param/{p}"] = arr.astype(np.float64)
                out[f"{key}/paramdims/{p}"] = np.array(",".join(d))
    for k, v in inputs.items():
        out[f"{key}/in/{k}"] = v


def main():
    mods, EchoData = install_reference()
    import calibrate_cases as cc

    cls_of = {"ek60": mods["calibrate.calibrate_ek"].CalibrateEK60, "ek80": mods["calibrate.calibrate_ek"].CalibrateEK80,
              "azfp": mods["calibrate.calibrate_azfp"].CalibrateAZFP}
    rng_mod = mods["calibrate.range"]
    est_noise, rm_noise = lift_noise(mods)
    out = {}

    def run(cls, ed, cal_type, **kw):
        # the call sequence of calibrate/api.py:64-87 (_compute_cal_ds)
        args = dict(env_params=None, cal_params=None, ecs_file=None, waveform_mode=None, encode_mode=None,
                    drop_last_hanning_zero=False, slice_dict={})
        args.update({k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
        obj = cls(to_ref_echodata(ed, EchoData), **args)
        obj._check_echodata_backscatter_size()
        return obj.compute_Sv() if cal_type == "Sv" else obj.compute_TS()

    for key, (maker, kw, calkw) in cc.CASES.items():
        ed = cc.build(key)
        inputs = {}
        if key not in cc.INPUT_OF:
            b = ed["Sonar/Beam_group1"]
            inputs["backscatter_r"] = np.asarray(b["backscatter_r"].values)
            if "backscatter_i" in b:
                inputs["backscatter_i"] = np.asarray(b["backscatter_i"].values)
        for ct in cc.cal_types(key):
            ds = run(cls_of[maker], ed, ct, **calkw)
            dump_case(out, key, ds, ct, inputs)
        if key == "noise":  # estimate / remove background noise on the reference's own Sv dataset
            for tag, (pn, rn, nmax, snr) in cc.NOISE_ARGS.items():
                out[f"noise/{tag}/est"] = canon(est_noise(ds, pn, rn, background_noise_max=nmax))[0]
                res = rm_noise(ds.copy(), pn, rn, background_noise_max=nmax, SNR_threshold=snr)
                out[f"noise/{tag}/Sv_noise"] = canon(res["Sv_noise"])[0]
                out[f"noise/{tag}/Sv_corrected"] = canon(res["Sv_corrected"])[0]

    # ---- the range functions on their own (range.py:98-201), EK80 with a GPT channel -------------------------------
    red = to_ref_echodata(cc.build("ek80_cw_power"), EchoData)
    beam, vend = red["Sonar/Beam_group1"], red["Vendor_specific"]
    r = rng_mod.compute_range_EK("EK80", beam, {"sound_speed": 1481.0})
    out["range/ek80_range"] = canon(r)[0]
    out["range/ek80_tvg"] = canon(rng_mod.range_mod_TVG_EK("EK80", beam, vend, r.copy(), 1481.0))[0]

    np.savez_compressed(cc.VECTORS, **{k.replace("/", "__"): v for k, v in out.items()})
    print("wrote", cc.VECTORS, len(out), "arrays,", os.path.getsize(cc.VECTORS) // 1024, "KiB")


if __name__ == "__main__":
    main()
