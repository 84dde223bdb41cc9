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
