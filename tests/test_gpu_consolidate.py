"""consolidate.add_depth on the device against the oracle, and compute_MVBS(range_var="depth") on the resulting law
(index-space binning of depth = offset[p] + echo_range * cos(tilt[p]))."""

import numpy as np
import pytest

import oracle_glue as og
from oracle import commongrid as ogrid
from oracle import consolidate as ocons

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import echopype_b200 as ep

    return ep


def _ns(t):
    return np.asarray(t).astype("datetime64[ns]").astype(np.int64)


def _series(ep, pt, values, every):
    t = pt[::every]
    return ep.DataArray(np.asarray(values, dtype=np.float64)[: len(t)], dims=("time3",), coords={"time3": t})


@pytest.mark.parametrize("mode", ["numbers", "series", "platform", "beam", "upward"])
def test_add_depth_matches_oracle(ep, mode):
    from echopype_b200 import synth
    from echopype_b200.dataset import Dataset

    C, P, R = 3, 41, 516
    ed = synth.make_ek60(C, P, R, seed=12, nan_tail=0.2)
    ds = ep.calibrate.compute_Sv(ed)
    ref = og.ek60(ed, "Sv")
    pt = ds["ping_time"].values
    rs = np.random.default_rng(4)
    kw, want = {}, None
    if mode == "numbers":
        kw = dict(depth_offset=7.5, tilt=12.0)
        want = ocons.add_depth(ref["echo_range"], 7.5, np.cos(np.deg2rad(12.0)))
    elif mode == "upward":
        kw = dict(depth_offset=250.0, tilt=3.0, downward=False)
        want = ocons.add_depth(ref["echo_range"], 250.0, np.cos(np.deg2rad(3.0)), downward=False)
    elif mode == "series":
        off = 5 + rs.random(P)
        tl = 10 * rs.random(P)
        t3 = pt[::4]
        kw = dict(depth_offset=_series(ep, pt, off, 4), tilt=_series(ep, pt, tl, 4))
        want = ocons.add_depth(ref["echo_range"], ocons.align_nearest(off[: len(t3)], _ns(t3), _ns(pt)),
                               np.cos(np.deg2rad(ocons.align_nearest(tl[: len(t3)], _ns(t3), _ns(pt)))))
    elif mode == "platform":
        t2 = pt[::5]
        n2 = len(t2)
        plat = Dataset(
            {"water_level": (("time2",), rs.random(n2)), "vertical_offset": (("time2",), rs.random(n2) - 0.5),
             "transducer_offset_z": (("time2",), 4 + rs.random(n2)), "pitch": (("time2",), 6 * rs.random(n2) - 3),
             "roll": (("time2",), 8 * rs.random(n2) - 4)},
            coords={"time2": t2},
        )
        ed["Platform"] = plat
        kw = dict(echodata=ed, use_platform_vertical_offsets=True, use_platform_angles=True)
        td = plat["transducer_offset_z"].values - (plat["water_level"].values + plat["vertical_offset"].values)
        sc = ocons.platform_angle_scaling(plat["pitch"].values, plat["roll"].values)
        want = ocons.add_depth(ref["echo_range"], ocons.align_nearest(td, _ns(t2), _ns(pt)), ocons.align_nearest(sc, _ns(t2), _ns(pt)))
    else:
        beam = ed["Sonar/Beam_group1"]
        x, y, z = np.array([0.1, 0.0, 0.0]), np.array([0.0, 0.2, 0.0]), np.array([0.99, 0.97, 0.0])
        for n, v in zip("xyz", (x, y, z)):
            beam[f"beam_direction_{n}"] = (("channel",), v)
        kw = dict(echodata=ed, use_beam_angles=True)
        want = ocons.add_depth(ref["echo_range"], 0.0, ocons.beam_angle_scaling(x, y, z), per_channel=True)
    out = ep.consolidate.add_depth(ds, **kw)
    got = out["depth"].values.astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=2.5e-7, atol=1e-6, equal_nan=True)
    assert "`depth` calculated using: Sv `echo_range`" in out["depth"].attrs["history"]
    assert out.attrs.get("processing_level", None) in (None, "Level 2A")


def test_add_depth_argument_errors(ep):
    from echopype_b200 import synth

    ed = synth.make_ek60(2, 8, 64)
    ds = ep.calibrate.compute_Sv(ed)
    with pytest.raises(ValueError, match="then `echodata` cannot be `None`"):
        ep.consolidate.add_depth(ds, use_platform_angles=True)
    with pytest.raises(NotImplementedError, match="both platform and beam angles"):
        ep.consolidate.add_depth(ds, echodata=ed, use_platform_angles=True, use_beam_angles=True)


def test_mvbs_on_depth_uses_exact_law(ep):
    """MVBS binned on depth = offset[p] + echo_range cos(tilt[p]) against the oracle's float64 binning of the same depth."""
    from echopype_b200 import synth

    C, P, R = 3, 90, 1000
    ed = synth.make_ek60(C, P, R, seed=14, nan_tail=0.1)
    ds = ep.calibrate.compute_Sv(ed)
    pt = ds["ping_time"].values
    rs = np.random.default_rng(8)
    off, tl = 3 + 2 * rs.random(P), 15 * rs.random(P)
    ds = ep.consolidate.add_depth(ds, depth_offset=ep.DataArray(off, dims=("ping_time",), coords={"ping_time": pt}),
                                  tilt=ep.DataArray(tl, dims=("ping_time",), coords={"ping_time": pt}))
    assert getattr(ds["depth"], "law", None) is not None and ds["depth"].law["kind"] == "depth"
    mv = ep.commongrid.compute_MVBS(ds, range_var="depth", range_bin="10m", ping_time_bin="15s")
    ref = og.ek60(ed, "Sv")
    depth = ocons.add_depth(ref["echo_range"], off, np.cos(np.deg2rad(tl)))
    # the reference takes the bin grid from nanmax(depth) in float64; the product from its float32 depth array
    want = ogrid.compute_MVBS(ref["out"], depth, _ns(pt), range_bin="10m", ping_time_bin="15s")
    got = mv["Sv"].values
    nR = min(got.shape[2], want["Sv"].shape[2])
    assert abs(got.shape[2] - want["Sv"].shape[2]) <= 1
    g, w = got[:, :, :nR], want["Sv"][:, :, :nR]
    assert np.array_equal(np.isnan(g), np.isnan(w))
    ok = ~np.isnan(w)
    assert np.abs(g[ok] - w[ok]).max() <= 1e-4
