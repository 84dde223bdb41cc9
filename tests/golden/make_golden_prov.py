"""Processing-level / provenance attributes of the reference, from EXECUTED reference code.

Run in the builder container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_prov.py
Writes tests/golden/prov_cases.json (committed).

``echopype/utils/prov.py`` is IMPORTED unmodified (namespace skeleton of make_golden_calibrate.py, xrlite as xarray, a stub
``_echopype_version`` module): ``add_processing_level`` wraps a function that returns a prepared Dataset, for every
processing-level code the path uses (and invalid ones), with / without valid ``latitude`` / ``longitude`` and with /
without ``input_processing_level``; ``insert_input_processing_level`` and ``echopype_prov_attrs`` are called as they are.
"""

import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden_calibrate as mgc  # noqa: E402
import xrlite  # noqa: E402

CODES = ["L2A", "L2B", "L3A", "L3B", "L4", "L*A", "L*B", "L2*", "L3*", "L5", "L**", "2A", "L*C"]
LATLON = ["valid", "all_nan", "absent", "lat_only"]
INPUT_LEVEL = [None, "Level 2A", "Level 2B", "Level 3B"]


def make_ds(latlon, input_level):
    ds = xrlite.Dataset(coords={"ping_time": ("ping_time", np.arange(4))})
    ds["Sv"] = (("ping_time",), np.zeros(4))
    if latlon in ("valid", "all_nan", "lat_only"):
        ds["latitude"] = (("ping_time",), np.full(4, np.nan) if latlon == "all_nan" else np.array([44.0, np.nan, 44.2, 44.3]))
    if latlon in ("valid", "all_nan"):
        ds["longitude"] = (("ping_time",), np.full(4, np.nan) if latlon == "all_nan" else np.array([-124.0, -124.1, np.nan, -124.3]))
    if input_level is not None:
        ds.attrs["input_processing_level"] = input_level
    ds.attrs["keep"] = "me"
    return ds


_STATE = {}


def produce():  # module level: the decorator tells functions from methods by the qualified name
    return make_ds(_STATE["ll"], _STATE["lev"])


def produce_number():
    return 3


def main():
    mgc.install_reference()
    ver = types.ModuleType("_echopype_version")
    ver.version = "0.0.0+reference"
    sys.modules["_echopype_version"] = ver
    prov = importlib.import_module("echopype.utils.prov")
    cases = []
    for code in CODES:
        for ll in LATLON:
            for lev in INPUT_LEVEL:
                rec = {"code": code, "latlon": ll, "input_level": lev}
                _STATE["ll"], _STATE["lev"] = ll, lev
                try:
                    out = prov.add_processing_level(code)(produce)()
                    rec["attrs"] = {k: v for k, v in out.attrs.items()}
                except Exception as e:  # noqa
                    rec["error"] = [type(e).__name__, str(e)]
                cases.append(rec)
    ins = []
    for in_attrs in [{}, {"processing_level": "Level 2A"}, {"processing_level": "Level 3B", "other": 1}]:
        src = xrlite.Dataset(attrs=dict(in_attrs))
        out = prov.insert_input_processing_level(xrlite.Dataset(attrs={"a": 1}), input_ds=src)
        ins.append({"input_attrs": in_attrs, "attrs": dict(out.attrs)})
    pa = prov.echopype_prov_attrs(process_type="processing")
    with open(os.path.join(HERE, "prov_cases.json"), "w") as fh:
        json.dump({"add_processing_level": cases, "insert_input_processing_level": ins, "prov_attr_keys": sorted(pa),
                   "not_a_dataset_error": _not_a_dataset(prov)}, fh, indent=1, sort_keys=True)
    print("wrote", len(cases), "decorator cases;", sum("error" in c for c in cases), "errors")


def _not_a_dataset(prov):
    try:
        prov.add_processing_level("L2A")(produce_number)()
    except Exception as e:  # noqa
        return [type(e).__name__, str(e)]
    return None


if __name__ == "__main__":
    main()
