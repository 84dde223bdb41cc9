"""mask.frequency_differencing / mask.apply_mask: parsing and validation on CPU (messages as in the reference's tests),
the reference's documented known answer, and GPU parity against the oracle."""

import numpy as np
import pytest

from oracle import mask as omask


def test_parse_freq_diff_eq():
    from echopype_b200.mask.freq_diff import _parse_freq_diff_eq

    assert _parse_freq_diff_eq('38.0kHz - 120.0kHz >= 10.0dB', None) == [[38000.0, 120000.0], None, ">=", 10.0]
    assert _parse_freq_diff_eq("1 Hz-2MHz<5dB", None) == [[1.0, 2e6], None, "<", 5.0]
    assert _parse_freq_diff_eq(None, '"chan1" - "chan2">=10.0dB') == [None, ["chan1", "chan2"], ">=", 10.0]
    with pytest.raises(ValueError, match="Either freqAB or chanAB must be given!"):
        _parse_freq_diff_eq(None, None)
    with pytest.raises(ValueError, match="Only one of freqAB or chanAB should be given, but not both!"):
        _parse_freq_diff_eq("1Hz-2Hz>1dB", '"a"-"b">1dB')
    with pytest.raises(ValueError, match="Invalid operator!"):
        _parse_freq_diff_eq("1Hz - 2Hz != 1dB", None)
    with pytest.raises(ValueError, match="freqAB must be a list of length 2 with unique elements!"):
        _parse_freq_diff_eq("2Hz - 2Hz > 1dB", None)
    with pytest.raises(ValueError, match="chanAB must be a list of length 2 with unique elements!"):
        _parse_freq_diff_eq(None, '"a" - "a" > 1dB')
    with pytest.raises(TypeError, match="Invalid freqAB Equation!"):
        _parse_freq_diff_eq("garbage", None)
    with pytest.raises(TypeError, match="Invalid chanAB Equation!"):
        _parse_freq_diff_eq(None, "garbage")


def test_oracle_matches_reference_docstring_example():
    """echopype/mask/api.py:533-560: rows 2..4 True."""
    n = 5
    Sv = np.stack([np.arange(n**2).reshape(n, n), np.identity(n)]).astype(np.float64)
    m = omask.frequency_differencing(Sv, 0, 1, ">=", 10.0)
    want = np.zeros((n, n), bool)
    want[2:] = True
    np.testing.assert_array_equal(m, want)


def _mock_ds(ep, C=3, P=37, R=516, seed=0, on_device=False):
    rs = np.random.default_rng(seed)
    Sv = (-70 + 15 * rs.standard_normal((C, P, R))).astype(np.float32)
    Sv[rs.random((C, P, R)) < 0.05] = np.nan
    ds = ep.Dataset(
        {"Sv": (("channel", "ping_time", "range_sample"), Sv), "frequency_nominal": (("channel",), np.array([18e3, 38e3, 120e3][:C]))},
        coords={"channel": np.array(["chan1", "chan2", "chan3"][:C], dtype=object),
                "ping_time": np.datetime64("2020-01-01", "ns") + np.arange(P).astype("timedelta64[s]"), "range_sample": np.arange(R)},
    )
    return ds, Sv


@pytest.mark.gpu
@pytest.mark.parametrize("eq,kind", [('"chan1" - "chan3" >= 3.0dB', "chan"), ("38.0kHz - 18kHz<2.5dB", "freq"), ('"chan2" - "chan1" == 0dB', "chan")])
def test_frequency_differencing_gpu(eq, kind):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import echopype_b200 as ep

    ds, Sv = _mock_ds(ep)
    m = ep.mask.frequency_differencing(ds, freqABEq=eq if kind == "freq" else None, chanABEq=eq if kind == "chan" else None)
    assert m.name == "mask" and m.dims == ("ping_time", "range_sample")
    a, b, op, diff = {'"chan1" - "chan3" >= 3.0dB': (0, 2, ">=", 3.0), "38.0kHz - 18kHz<2.5dB": (1, 0, "<", 2.5), '"chan2" - "chan1" == 0dB': (1, 0, "==", 0.0)}[eq]
    want = omask.frequency_differencing(Sv.astype(np.float32).astype(np.float64), a, b, op, diff)
    np.testing.assert_array_equal(m.values.astype(bool), want)
    assert m.attrs["mask_type"] == "frequency differencing"
    assert f"Operation: Sv['chan{a + 1}'] - Sv['chan{b + 1}'] {op} {diff}" in m.attrs["history"]
    with pytest.raises(ValueError, match="not in the channel coordinate"):
        ep.mask.frequency_differencing(ds, chanABEq='"chan1" - "nope" > 1dB')
    with pytest.raises(ValueError, match="not in the frequency_nominal variable"):
        ep.mask.frequency_differencing(ds, freqABEq="38kHz - 70kHz > 1dB")


@pytest.mark.gpu
def test_apply_mask_gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import echopype_b200 as ep

    ds, Sv = _mock_ds(ep, seed=3)
    rs = np.random.default_rng(1)
    m2 = rs.random(Sv.shape[1:]) < 0.6            # (ping_time, range_sample), host bool
    m3 = (rs.random(Sv.shape) < 0.7).astype(np.float64)  # (channel, ping, range) float 0 / 1 (NaN entries are rejected, below)
    mfd = ep.mask.frequency_differencing(ds, chanABEq='"chan1" - "chan2" > 1.0dB')  # device uint8
    masks = [mfd, ep.DataArray(m2, ("ping_time", "range_sample")), ep.DataArray(m3, ("channel", "ping_time", "range_sample"))]
    out = ep.mask.apply_mask(ds, masks, fill_value=-999.0)
    want = omask.apply_mask(Sv, [omask.frequency_differencing(Sv, 0, 1, ">", 1.0), m2, m3], -999.0)
    got = out["Sv"].values.astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(np.nan_to_num(got, nan=1.0), np.nan_to_num(want, nan=1.0))
    assert out["Sv"].attrs["long_name"] == "Volume backscattering strength, masked (Sv re 1 m-1)"
    assert out["Sv"].attrs["mask_type"] == "frequency differencing"
    assert out.attrs["mask_function"] == "mask.apply_mask"
    nanfill = ep.mask.apply_mask(ds, mfd)
    w2 = omask.apply_mask(Sv, [omask.frequency_differencing(Sv, 0, 1, ">", 1.0)])
    g2 = nanfill["Sv"].values
    assert np.array_equal(np.isnan(g2), np.isnan(w2))
    ar = nanfill["Sv"].attrs["actual_range"]
    assert ar == [round(float(np.nanmin(w2)), 2), round(float(np.nanmax(w2)), 2)]
    with pytest.raises(ValueError, match="is not of the same shape"):
        ep.mask.apply_mask(ds, ep.DataArray(m2[:, :-1], ("ping_time", "range_sample")))
    # fill_value given as an array of the shape of one channel (mask/api.py:233-246), broadcast over channel
    fill = rs.normal(-90.0, 1.0, size=Sv.shape[1:]).astype(np.float32)
    got3 = ep.mask.apply_mask(ds, masks[:2], fill_value=ep.DataArray(fill[None], ("channel", "ping_time", "range_sample")))["Sv"].values
    want3 = omask.apply_mask(Sv, [omask.frequency_differencing(Sv, 0, 1, ">", 1.0), m2], fill.astype(np.float64))
    assert np.array_equal(np.isnan(got3), np.isnan(want3))
    np.testing.assert_array_equal(np.nan_to_num(got3.astype(np.float64), nan=1.0), np.nan_to_num(want3, nan=1.0))
    with pytest.raises(ValueError, match="If fill_value is an array it must be of the same shape as Sv!"):
        ep.mask.apply_mask(ds, mfd, fill_value=ep.DataArray(fill[:, :-1], ("ping_time", "range_sample")))
    with pytest.raises(TypeError, match="The input fill_value must be of type int, float, or xr.DataArray!"):
        ep.mask.apply_mask(ds, mfd, fill_value="nan")


def test_apply_mask_rejects_what_the_reference_rejects():
    """mask/api.py:131-160 (_validate_and_collect_mask_input) and :41-71 (_check_mask_dim_alignment): same exception types
    and messages, raised before any device work."""
    import echopype_b200 as ep

    C, P, R = 2, 4, 5
    dims = ("channel", "ping_time", "range_sample")
    ds = ep.Dataset({"Sv": (dims, np.zeros((C, P, R), np.float32))},
                    coords={"channel": np.array(["a", "b"], dtype=object), "ping_time": np.arange(P), "range_sample": np.arange(R)})
    nanmask = np.ones((P, R))
    nanmask[1, 2] = np.nan
    with pytest.raises(TypeError, match="Mask cannot contain NaN"):
        ep.mask.apply_mask(ds, ep.DataArray(nanmask, ("ping_time", "range_sample")))
    with pytest.raises(TypeError, match=r"Mask must be boolean \(True/False or 1/0\)"):
        ep.mask.apply_mask(ds, ep.DataArray(np.full((P, R), 2.0), ("ping_time", "range_sample")))
    with pytest.raises(TypeError, match=r"Mask must be boolean \(True/False or 1/0\)"):
        ep.mask.apply_mask(ds, [ep.DataArray(np.ones((P, R), bool), ("ping_time", "range_sample")),
                                ep.DataArray(np.arange(P * R).reshape(P, R), ("ping_time", "range_sample"))])
    with pytest.raises(ValueError, match="Masks must have one of the following dimensions"):
        ep.mask.apply_mask(ds, ep.DataArray(np.ones((P, R), bool), ("time", "range_sample")))
    with pytest.raises(ValueError, match="Masks must have one of the following dimensions"):
        ep.mask.apply_mask(ds, ep.DataArray(np.ones(P, bool), ("ping_time",)))
    with pytest.raises(ValueError, match="do not match the dimensions of source"):
        ep.mask.apply_mask(ds, ep.DataArray(np.ones((P, R), bool), ("ping_time", "depth")))
    with pytest.raises(ValueError, match="The Dataset source_ds does not contain the variable var_name!"):
        ep.mask.apply_mask(ds, ep.DataArray(np.ones((P, R), bool), ("ping_time", "range_sample")), var_name="Sv_corrected")
