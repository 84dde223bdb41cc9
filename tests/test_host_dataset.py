"""Host-side container tests: the xarray interop (as_dataset / Dataset.to_xarray) exercised with tests/golden/xrlite.py
standing in for xarray (same constructor and accessor surface; the real package is not installed in the build image), and
the Dataset surface the reference's documented chain uses (rename_vars / where / swap_dims / drop_vars / isel,
tests/utils/test_processinglevels_integration.py:103-141)."""

import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import xrlite  # noqa: E402

from echopype_b200.dataset import DataArray, Dataset, EchoData, as_dataset  # noqa: E402

DIMS = ("channel", "ping_time", "range_sample")


def _xr_ds(C=2, P=5, R=7):
    rng = np.random.default_rng(0)
    ds = xrlite.Dataset(coords={"channel": ("channel", np.array([f"ch{c}" for c in range(C)])),
                                "ping_time": ("ping_time", np.datetime64("2024-01-01T00:00:00", "ns") + np.arange(P) * np.timedelta64(1, "s")),
                                "range_sample": ("range_sample", np.arange(R))},
                        attrs={"processing_level": "Level 2A"})
    ds["Sv"] = (DIMS, rng.normal(-70, 5, (C, P, R)))
    ds["echo_range"] = (DIMS, np.broadcast_to(0.2 * np.arange(R), (C, P, R)).copy())
    ds["frequency_nominal"] = (("channel",), np.array([18e3, 38e3][:C]))
    ds["Sv"].attrs["units"] = "dB"
    return ds


def test_as_dataset_from_xarray_like_and_back(monkeypatch):
    x = _xr_ds()
    ds = as_dataset(x)
    assert isinstance(ds, Dataset) and as_dataset(ds) is ds
    assert set(ds.data_vars) == {"Sv", "echo_range", "frequency_nominal"} and set(ds.coords) == set(DIMS)
    assert tuple(ds["Sv"].dims) == DIMS and ds["Sv"].attrs["units"] == "dB" and ds.attrs["processing_level"] == "Level 2A"
    np.testing.assert_array_equal(ds["Sv"].values, x["Sv"].values)
    np.testing.assert_array_equal(ds["ping_time"].values, x["ping_time"].values)
    # and back, with the stand-in registered as xarray
    monkeypatch.setitem(sys.modules, "xarray", xrlite)
    back = ds.to_xarray()
    assert isinstance(back, xrlite.Dataset) and set(back.data_vars) == set(x.data_vars)
    for k in x.data_vars:
        assert tuple(back[k].dims) == tuple(x[k].dims)
        np.testing.assert_array_equal(back[k].values, x[k].values)
    np.testing.assert_array_equal(back["channel"].values, x["channel"].values)
    assert back.attrs["processing_level"] == "Level 2A"


def test_as_dataset_rejects_other_types_and_accepts_mappings():
    with pytest.raises(TypeError):
        as_dataset(np.zeros(3))
    ds = as_dataset({"Sv": (DIMS, np.zeros((1, 2, 3)))})
    assert ds["Sv"].shape == (1, 2, 3)
    ed = EchoData("EK60", {"Sonar/Beam_group1": _xr_ds()})
    assert isinstance(ed["Sonar/Beam_group1"], Dataset) and len(ed["Platform"].data_vars) == 0
    with pytest.raises(KeyError):
        ed["Environment"]


def test_dataset_surface_of_the_documented_chain():
    ds = as_dataset(_xr_ds())
    ds["Sv_corrected"] = DataArray(ds["Sv"].values + 1.0, DIMS)
    # rename Sv_corrected -> Sv after dropping the original (test_processinglevels_integration.py:117-120)
    ren = ds.drop_vars("Sv").rename_vars({"Sv_corrected": "Sv"})
    assert "Sv_corrected" not in ren and np.allclose(ren["Sv"].values, ds["Sv"].values + 1.0)
    assert "Sv" in ds and "Sv_corrected" in ds  # the source is untouched
    with pytest.raises(ValueError):
        ds.drop_vars("nope")
    assert "Sv" in ds.drop_vars("nope", errors="ignore")
    # where: NaN outside the condition, dims kept
    w = ds.where(ds["Sv"] > -70.0)
    assert np.array_equal(np.isnan(w["Sv"].values), ~(ds["Sv"].values > -70.0))
    # swap_dims channel -> frequency_nominal (consolidate.swap_dims_channel_frequency uses it)
    sw = ds.swap_dims({"channel": "frequency_nominal"})
    assert sw["Sv"].dims[0] == "frequency_nominal" and "frequency_nominal" in sw.coords
    # isel keeps the coordinates aligned
    sub = ds.isel(ping_time=slice(1, 4), range_sample=slice(0, 3))
    assert sub["Sv"].shape == (2, 3, 3) and len(sub["ping_time"].values) == 3
    np.testing.assert_array_equal(sub["Sv"].values, ds["Sv"].values[:, 1:4, :3])


# ---- processing-level attributes against the EXECUTED reference decorator (tests/golden/make_golden_prov.py) ------------
def _prov_ds(latlon, input_level):
    ds = Dataset({"Sv": (("ping_time",), np.zeros(4))}, coords={"ping_time": np.arange(4)})
    if latlon in ("valid", "all_nan", "lat_only"):
        ds["latitude"] = (("ping_time",), np.full(4, np.nan) if latlon == "all_nan" else np.array([44.0, np.nan, 44.2, 44.3]))
    if latlon in ("valid", "all_nan"):
        ds["longitude"] = (("ping_time",), np.full(4, np.nan) if latlon == "all_nan" else np.array([-124.0, -124.1, np.nan, -124.3]))
    if input_level is not None:
        ds.attrs["input_processing_level"] = input_level
    ds.attrs["keep"] = "me"
    return ds


_PROV_STATE = {}


def produce():  # module level and named like the generator's function: the error messages quote the qualified name
    return _prov_ds(_PROV_STATE["ll"], _PROV_STATE["lev"])


def produce_number():
    return 3


def test_processing_level_decorator_equals_reference():
    import json

    from echopype_b200.utils import prov

    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prov_cases.json")))
    n_ok = n_err = 0
    for c in cases["add_processing_level"]:
        _PROV_STATE["ll"], _PROV_STATE["lev"] = c["latlon"], c["input_level"]
        if "error" in c:
            etype = {"ValueError": ValueError, "RuntimeError": RuntimeError}[c["error"][0]]
            with pytest.raises(etype) as ei:
                prov.add_processing_level(c["code"])(produce)()
            assert str(ei.value) == c["error"][1], c
            n_err += 1
        else:
            out = prov.add_processing_level(c["code"])(produce)()
            assert dict(out.attrs) == c["attrs"], c
            n_ok += 1
    assert n_ok == 140 and n_err == 68
    for c in cases["insert_input_processing_level"]:
        out = prov.insert_input_processing_level(Dataset(attrs={"a": 1}), input_ds=Dataset(attrs=dict(c["input_attrs"])))
        assert dict(out.attrs) == c["attrs"]
    assert sorted(prov.echopype_prov_attrs(process_type="processing")) == cases["prov_attr_keys"]
    with pytest.raises(RuntimeError) as ei:
        prov.add_processing_level("L2A")(produce_number)()
    assert str(ei.value) == cases["not_a_dataset_error"][1]
