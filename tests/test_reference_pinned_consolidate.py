"""consolidate.add_depth: the oracle, the product's host-side alignment / scaling functions AND the CUDA result against
outputs of the reference's own code (utils/align.py and consolidate/ek_depth_utils.py imported unmodified, add_depth
lifted from consolidate/api.py:66-247 and executed by tests/golden/make_golden_consolidate.py)."""

import os

import numpy as np
import pytest

from oracle import consolidate as ocons

HERE = os.path.dirname(os.path.abspath(__file__))
DIMS = ("channel", "ping_time", "range_sample")


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(HERE, "golden", "consolidate_vectors.npz"))


def _check(got, want, rtol=1e-14):
    assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=rtol, atol=0, equal_nan=True)


# ---- CPU: oracle == reference --------------------------------------------------------------------------------------------
def test_oracle_equals_reference_add_depth(vec):
    er, pt = vec["echo_range"], vec["ping_time"]
    _check(ocons.add_depth(er), vec["plain__depth"])
    _check(ocons.add_depth(er, 7.5, np.cos(np.deg2rad(12.0))), vec["numbers__depth"])
    _check(ocons.add_depth(er, 250.0, np.cos(np.deg2rad(3.0)), downward=False), vec["upward__depth"])
    t3 = vec["series__t3"]
    off = ocons.align_nearest(vec["series__off"], t3, pt)
    np.testing.assert_array_equal(off, vec["series__off_aligned"])
    _check(ocons.add_depth(er, off, np.cos(np.deg2rad(ocons.align_nearest(vec["series__tilt"], t3, pt)))), vec["series__depth"])
    _check(ocons.add_depth(er, ocons.align_nearest(vec["onping__off"], pt, pt)), vec["onping__depth"])
    _check(ocons.add_depth(er, ocons.align_nearest([4.25], pt[3:4], pt), np.cos(np.deg2rad(ocons.align_nearest([20.0], pt[5:6], pt)))),
           vec["single__depth"])
    t2 = vec["platform__t2"]
    td = vec["platform__transducer_offset_z"] - (vec["platform__water_level"] + vec["platform__vertical_offset"])
    sc = ocons.platform_angle_scaling(vec["platform__pitch"], vec["platform__roll"])
    np.testing.assert_allclose(ocons.align_nearest(td, t2, pt), vec["platform__transducer_depth"], rtol=1e-15)
    np.testing.assert_allclose(ocons.align_nearest(sc, t2, pt), vec["platform__scaling"], rtol=1e-15)
    _check(ocons.add_depth(er, ocons.align_nearest(td, t2, pt), ocons.align_nearest(sc, t2, pt)), vec["platform__depth"])
    _check(ocons.add_depth(er, ocons.align_nearest(td, t2, pt)), vec["platform_offsets_only__depth"])
    bs = ocons.beam_angle_scaling(vec["beam__x"], vec["beam__y"], vec["beam__z"])
    np.testing.assert_allclose(bs, vec["beam__scaling"], rtol=1e-15, equal_nan=True)
    _check(ocons.add_depth(er, 0.0, bs, per_channel=True), vec["beam__depth"])


# ---- CPU: the product's host-side functions == reference ----------------------------------------------------------------
def _platform(vec):
    from echopype_b200.dataset import Dataset

    return Dataset({k: (("time2",), vec[f"platform__{k}"]) for k in ("water_level", "vertical_offset", "transducer_offset_z", "pitch", "roll")},
                   coords={"time2": vec["platform__t2"].astype("datetime64[ns]")})


def _beam(vec):
    from echopype_b200.dataset import Dataset

    return Dataset({f"beam_direction_{n}": (("channel",), vec[f"beam__{n}"]) for n in "xyz"}, coords={"channel": np.array(vec["channel"], dtype=object)})


def test_host_alignment_and_scaling_equal_reference(vec):
    from echopype_b200.consolidate import api as capi

    pt = vec["ping_time"].astype("datetime64[ns]")
    got = capi.align_to_ping_time(vec["series__off"], vec["series__t3"].astype("datetime64[ns]"), pt)
    np.testing.assert_array_equal(got, vec["series__off_aligned"])
    np.testing.assert_allclose(capi.ek_use_platform_vertical_offsets(_platform(vec), pt), vec["platform__transducer_depth"], rtol=1e-15)
    # cos(pitch) cos(roll) against scipy's rotation matrix element: equal to a few ulp
    np.testing.assert_allclose(capi.ek_use_platform_angles(_platform(vec), pt), vec["platform__scaling"], rtol=1e-14)
    np.testing.assert_allclose(capi.ek_use_beam_angles(_beam(vec)), vec["beam__scaling"], rtol=1e-15, equal_nan=True)


# ---- GPU: the product == reference ----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import echopype_b200 as ep

    return ep


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plain", "numbers", "upward", "series", "onping", "single", "platform", "platform_offsets_only", "beam"])
def test_cuda_add_depth_equals_reference(ep, vec, mode):
    from echopype_b200.dataset import Dataset, EchoData

    er = vec["echo_range"]
    C, P, R = er.shape
    pt = vec["ping_time"].astype("datetime64[ns]")
    ds = Dataset({"echo_range": (DIMS, er.astype(np.float32)), "Sv": (DIMS, np.zeros((C, P, R), np.float32))},
                 coords={"channel": np.array(vec["channel"], dtype=object), "ping_time": pt, "range_sample": np.arange(R)})

    def series(values, t_ns, dim="time3"):
        return ep.DataArray(np.asarray(values, dtype=np.float64), dims=(dim,), coords={dim: np.asarray(t_ns).astype("datetime64[ns]")})

    ed = EchoData("EK60", {"Platform": _platform(vec), "Sonar/Beam_group1": _beam(vec)})
    kw = {
        "plain": {},
        "numbers": dict(depth_offset=7.5, tilt=12.0),
        "upward": dict(depth_offset=250.0, tilt=3.0, downward=False),
        "series": dict(depth_offset=series(vec["series__off"], vec["series__t3"]), tilt=series(vec["series__tilt"], vec["series__t3"])),
        "onping": dict(depth_offset=series(vec["onping__off"], vec["ping_time"], "time_x")),
        "single": dict(depth_offset=series([4.25], vec["ping_time"][3:4]), tilt=series([20.0], vec["ping_time"][5:6])),
        "platform": dict(echodata=ed, use_platform_vertical_offsets=True, use_platform_angles=True),
        "platform_offsets_only": dict(echodata=ed, use_platform_vertical_offsets=True),
        "beam": dict(echodata=ed, use_beam_angles=True),
    }[mode]
    got = np.asarray(ep.consolidate.add_depth(ds, **kw)["depth"].values, dtype=np.float64)
    want = vec[f"{mode}__depth"]
    assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=2.5e-7, atol=1e-6, equal_nan=True)  # float32 depth against the float64 reference


# ---- mask.frequency_differencing against the executed reference function (tests/golden/make_golden_mask.py) ------------
FD_KEYS = ["f_gt", "f_le", "f_eq", "f_ge", "c_lt", "c_ge"]


@pytest.fixture(scope="module")
def fvec():
    return np.load(os.path.join(HERE, "golden", "freqdiff_vectors.npz"))


def _fd_dataset(fvec):
    from echopype_b200.dataset import Dataset

    Sv = fvec["Sv"]
    C, P, R = Sv.shape
    return Dataset({"Sv": (DIMS, Sv), "frequency_nominal": (("channel",), fvec["frequency_nominal"])},
                   coords={"channel": np.array(fvec["channel"], dtype=object),
                           "ping_time": np.datetime64("2024-01-01") + np.arange(P) * np.timedelta64(1, "s"), "range_sample": np.arange(R)})


def _fd_kw(fvec, key):
    import ast

    return ast.literal_eval(str(fvec[f"{key}__kw"]))


@pytest.mark.parametrize("key", FD_KEYS)
def test_oracle_and_host_parser_equal_reference_frequency_differencing(fvec, key):
    import re

    from echopype_b200.mask.freq_diff import _parse_freq_diff_eq
    from oracle import mask as omask

    kw = _fd_kw(fvec, key)
    freqAB, chanAB, operator, diff = _parse_freq_diff_eq(kw.get("freqABEq"), kw.get("chanABEq"))
    chans, freqs = list(fvec["channel"]), list(fvec["frequency_nominal"])
    a, b = ([freqs.index(f) for f in freqAB] if freqAB is not None else [chans.index(c) for c in chanAB])
    # the operation string of the reference's history attribute names the same channels, operator and threshold
    m = re.fullmatch(r"Sv\['(.+)'\] - Sv\['(.+)'\] (\S+) (\S+)", str(fvec[f"{key}__operation"]))
    assert (m.group(1), m.group(2), m.group(3), float(m.group(4))) == (chans[a], chans[b], operator, diff)
    got = omask.frequency_differencing(np.asarray(fvec["Sv"], dtype=np.float64), a, b, operator, diff)
    assert np.array_equal(got, fvec[f"{key}__mask"])


def test_host_frequency_differencing_rejects_what_the_reference_rejects(fvec):
    """Same exception type and message, before any device work (a negative right-hand side is not accepted by the
    reference's equation pattern: its operator group swallows the sign)."""
    import ast

    import echopype_b200 as ep

    ds = _fd_dataset(fvec)
    for kw_repr, etype, msg in fvec["bad__cases"]:
        kw = ast.literal_eval(str(kw_repr))
        assert str(etype) != "ok"
        with pytest.raises({"ValueError": ValueError, "TypeError": TypeError}[str(etype)]) as ei:
            ep.mask.frequency_differencing(source_Sv=ds, **kw)
        assert str(ei.value) == str(msg), (kw, str(ei.value), str(msg))


@pytest.mark.gpu
@pytest.mark.parametrize("key", FD_KEYS)
def test_cuda_frequency_differencing_equals_reference(ep, fvec, key):
    got = ep.mask.frequency_differencing(source_Sv=_fd_dataset(fvec), **_fd_kw(fvec, key))
    assert tuple(got.dims) == ("ping_time", "range_sample")
    assert np.array_equal(np.asarray(got.values).astype(bool), fvec[f"{key}__mask"])  # float32 differences of 0.5 dB steps are exact


# ---- mask.apply_mask against the executed reference function (tests/golden/make_golden_mask.py) -------------------------
def test_oracle_equals_reference_apply_mask(fvec):
    from oracle import mask as omask

    Sv = np.asarray(fvec["Sv"], dtype=np.float64)
    mfd, m2, m3, fill = fvec["am__mfd"], fvec["am__m2"], fvec["am__m3"], fvec["am__fill_arr"]
    for key, masks, fv in [("am_single_nan", [mfd], np.nan), ("am_list_fill", [mfd, m2, m3], -999.0), ("am_int_fill", [m3], 0),
                           ("am_fill_array", [m3, mfd], fill)]:
        got = omask.apply_mask(Sv, masks, fv)
        want = fvec[f"{key}__Sv"]
        assert np.array_equal(np.isnan(got), np.isnan(want)), key
        np.testing.assert_array_equal(np.nan_to_num(got, nan=1.0), np.nan_to_num(want, nan=1.0))
        lo, hi = fvec[f"{key}__actual_range"]
        assert [round(float(np.nanmin(got)), 2), round(float(np.nanmax(got)), 2)] == [lo, hi]
        assert str(fvec[f"{key}__long_name"]) == "Volume backscattering strength, masked (Sv re 1 m-1)"
    # the attributes of the FIRST mask travel to the masked variable (mask/api.py:290-297)
    assert str(fvec["am_single_nan__mask_type"]) == "frequency differencing" and str(fvec["am_int_fill__mask_type"]) == ""


def test_host_apply_mask_rejects_what_the_reference_rejects(fvec):
    """Every mask the reference's own apply_mask refused when it was executed (NaN entries, values other than 0 / 1, dimension
    sets outside the allowed list, dimensions that do not match the source) is refused here with the same exception type and
    message, before any device work.  Three recorded cases are deliberately different (known deviations):
    a (channel, range_sample, ping_time) mask is aligned by NAME here instead of failing the positional shape check, a
    missing var_name raises the helper's ValueError instead of xarray's KeyError, and the two shape checks that need the
    device-side shapes are exercised in the GPU test."""
    import echopype_b200 as ep

    ds = _fd_dataset(fvec)
    C, P, R = fvec["Sv"].shape
    nanmask = np.ones((P, R))
    nanmask[1, 2] = np.nan
    ours = {
        "nan": ep.DataArray(nanmask, ("ping_time", "range_sample")),
        "two": ep.DataArray(np.full((P, R), 2.0), ("ping_time", "range_sample")),
        "ints_in_list": [ep.DataArray(np.ones((P, R), bool), ("ping_time", "range_sample")),
                         ep.DataArray(np.arange(P * R).reshape(P, R), ("ping_time", "range_sample"))],
        "bad_dim_name": ep.DataArray(np.ones((P, R), bool), ("time", "range_sample")),
        "one_dim": ep.DataArray(np.ones(P, bool), ("ping_time",)),
        "depth_dim": ep.DataArray(np.ones((P, R), bool), ("ping_time", "depth")),
    }
    seen = set()
    for label, etype, msg in fvec["am_bad__cases"]:
        label, etype, msg = str(label), str(etype), str(msg)
        assert etype != "ok", label
        if label not in ours:
            continue
        seen.add(label)
        with pytest.raises({"ValueError": ValueError, "TypeError": TypeError}[etype]) as ei:
            ep.mask.apply_mask(ds, ours[label])
        if label == "depth_dim":  # the message prints two Python sets, whose element order is not fixed
            assert str(ei.value).startswith("The dimensions of mask: (") and str(ei.value).endswith("when not considering 'channel'.")
            assert msg.startswith("The dimensions of mask: (") and msg.endswith("when not considering 'channel'.")
        else:
            assert str(ei.value) == msg, (label, str(ei.value), msg)
    assert seen == set(ours)
    with pytest.raises(TypeError) as ei:
        ep.mask.apply_mask(ds, ep.DataArray(np.ones((P, R), bool), ("ping_time", "range_sample")), fill_value="nan")
    assert str(ei.value) == [str(m) for l, t, m in fvec["am_bad__cases"] if str(l) == "fill_str"][0]
