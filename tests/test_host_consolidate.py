"""CPU: host-side pieces of consolidate.add_depth against the oracle (scipy interp1d / Rotation, as the reference uses)."""

import numpy as np
import pytest

from oracle import consolidate as ocons


def test_align_nearest_matches_scipy_interp1d():
    from echopype_b200.consolidate.api import align_to_ping_time

    rs = np.random.default_rng(0)
    t = np.sort(rs.integers(0, 10**12, 37)).astype(np.int64)
    v = rs.standard_normal(37)
    p = np.sort(np.concatenate([rs.integers(-10**11, 11 * 10**11, 200), t[:5], (t[3:8] + t[4:9]) // 2])).astype(np.int64)
    got = align_to_ping_time(v, t.astype("datetime64[ns]"), p.astype("datetime64[ns]"))
    np.testing.assert_array_equal(got, ocons.align_nearest(v, t, p))
    np.testing.assert_array_equal(align_to_ping_time(v, t.astype("datetime64[ns]"), t.astype("datetime64[ns]")), v)
    np.testing.assert_array_equal(align_to_ping_time([2.5], t[:1].astype("datetime64[ns]"), p.astype("datetime64[ns]")), np.full(len(p), 2.5))
    assert np.isnan(align_to_ping_time([], t[:0].astype("datetime64[ns]"), p.astype("datetime64[ns]"))).all()


def test_platform_and_beam_scaling_closed_forms():
    pitch, roll = np.array([3.0, -2.0, 0.0, 10.0]), np.array([4.0, 1.0, 0.0, -7.0])
    np.testing.assert_allclose(ocons.platform_angle_scaling(pitch, roll), np.cos(np.deg2rad(pitch)) * np.cos(np.deg2rad(roll)), rtol=1e-14)
    sc = ocons.beam_angle_scaling([0.1, 0.0, 0.0], [0.0, 3.0, 0.0], [0.99, 4.0, 0.0])
    np.testing.assert_allclose(sc[:2], [0.99 / np.hypot(0.1, 0.99), 0.8])
    assert np.isnan(sc[2])


def test_add_depth_broadcasting():
    er = np.arange(2 * 3 * 4, dtype=np.float64).reshape(2, 3, 4)
    d = ocons.add_depth(er, np.array([1.0, 2.0, 3.0]), np.array([0.5, 1.0, 2.0]), downward=False)
    np.testing.assert_allclose(d[1, 2], 3.0 - er[1, 2] * 2.0)
    d = ocons.add_depth(er, 4.0, np.array([0.5, 2.0]), per_channel=True)
    np.testing.assert_allclose(d[1, 0], 4.0 + er[1, 0] * 2.0)


def test_dataset_rename_swap_where_and_law_rules():
    """Dataset surface the documented processing chain uses (rename_vars: tests/utils/test_processinglevels_integration.py:119)
    and the rule that ends index-space binning when "Sv" is replaced by an untrusted array."""
    import echopype_b200 as ep

    C, P, R = 2, 3, 4
    Sv = np.arange(C * P * R, dtype=np.float32).reshape(C, P, R)
    ds = ep.Dataset(
        {"Sv": ep.DataArray(Sv, ("channel", "ping_time", "range_sample"), law={"kind": "derived"}),
         "Sv_corrected": ep.DataArray(Sv + 1, ("channel", "ping_time", "range_sample"), law={"kind": "derived"}),
         "echo_range": ep.DataArray(Sv * 0.1, ("channel", "ping_time", "range_sample"), law={"kind": "echo_range", "rows": object(), "minmax": None}),
         "frequency_nominal": (("channel",), np.array([18e3, 38e3]))},
        coords={"channel": np.array(["a", "b"], dtype=object), "ping_time": np.arange(P), "range_sample": np.arange(R)},
    )
    out = ds.rename_vars(name_dict={"Sv": "Sv_raw", "Sv_corrected": "Sv"})
    assert set(out.data_vars) == {"Sv_raw", "Sv", "echo_range", "frequency_nominal"}
    np.testing.assert_array_equal(out["Sv"].values, Sv + 1)
    assert out["echo_range"].law["rows"] is not None  # Sv_corrected is trusted: the row table stays
    assert list(out.frequency_nominal.values) == [18e3, 38e3]  # attribute access like xarray
    with pytest.raises(ValueError):
        ds.rename_vars({"nope": "x"})
    # a user array in the place of Sv ends the guarantee
    ds2 = ds.copy()
    ds2["Sv"] = (("channel", "ping_time", "range_sample"), Sv * 2)
    assert ds2["echo_range"].law is None or ds2["echo_range"].law.get("rows") is None
    assert ds["echo_range"].law["rows"] is not None  # the original is untouched
    # swap_dims / rename / where
    sw = ds.swap_dims({"channel": "frequency_nominal"})
    assert sw["Sv"].dims == ("frequency_nominal", "ping_time", "range_sample") and list(sw["frequency_nominal"].values) == [18e3, 38e3]
    rn = ds.rename({"ping_time": "time"})
    assert rn["Sv"].dims == ("channel", "time", "range_sample") and "time" in rn.coords
    w = ds["Sv"].where(ds["Sv"] > 5.0)
    assert np.isnan(w.values[0, 0, 0]) and w.values[1, 2, 3] == Sv[1, 2, 3]
    w2 = ds.where(ds["Sv"] > 5.0, other=-999.0)
    assert w2["Sv"].values[0, 0, 0] == -999.0 and w2["frequency_nominal"].values[0] == 18e3
    assert w2["echo_range"].law is None or w2["echo_range"].law.get("rows") is None


def test_index_binning_ping_time_is_tile_mean():
    from echopype_b200.commongrid.api import _coarsen_time_mean

    t = np.datetime64("2020-01-01T00:00:00", "ns") + (np.array([0, 1, 2, 4, 10, 11, 30]) * 10**9).astype("timedelta64[ns]")
    got = _coarsen_time_mean(t, 3)
    want = np.datetime64("2020-01-01T00:00:00", "ns") + (np.array([1.0, 25 / 3, 30.0]) * 1e9).astype(np.int64).astype("timedelta64[ns]")
    np.testing.assert_array_equal(got, want)
