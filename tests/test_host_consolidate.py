"""CPU: host-side pieces of consolidate.add_depth against the oracle (scipy interp1d / Rotation, as the reference uses)."""

import numpy as np

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
