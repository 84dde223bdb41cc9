"""Parity at the BASELINE.json sizes.  The full volumes are generated on the device (Philox, echopype_b200.synth), run
through the product path once, and checked against the float64 oracle on ping SLICES copied back from the very same
device volume: the noise tiles (ping_num) and the ping bins are aligned with the slice boundaries, so the bins of the
slice in the full-size result must equal the oracle's result on the slice alone.  Plus size-independent properties
(member counts in closed form)."""

import numpy as np
import pytest

import oracle_glue as og
from oracle import clean as oclean
from oracle import commongrid as ogrid

pytestmark = pytest.mark.gpu
ATOL = 1e-4


@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, _ = torch.cuda.mem_get_info()
    if free < 120e9:
        pytest.skip("needs a B200-class device (HBM)")
    import echopype_b200 as ep

    return ep


def _ns(t):
    return np.asarray(t).astype("datetime64[ns]").astype(np.int64)


def _slice_ed(make, ed, p0, p1, **kw):
    """Host EchoData holding pings [p0, p1) of the device volume `ed` (same parameters, same ping times)."""
    x = ed["Sonar/Beam_group1"]["backscatter_r"].data[:, p0:p1].cpu().numpy()
    return make(P=p1 - p0, ping_offset=p0, backscatter=x, **kw)


def _check_slice(ds_full, ed_s, kind, p0, p1, pn, rn, bin_pings, extra=None):
    """MVBS bins [p0/bin_pings, p1/bin_pings) of the full-size result vs the oracle chain on the slice."""
    if kind == "ek60":
        ref = og.ek60(ed_s, "Sv")
    elif kind == "azfp":
        ref = og.azfp(ed_s, "Sv", 30.0, 50.0)
    else:
        ref = og.ek80(ed_s, "Sv", "CW", "power")
    Sv, rng = ref["out"], ref["echo_range"]
    marg = np.zeros(0)
    if pn:
        nz = oclean.remove_background_noise(Sv, rng, ref["sound_absorption"], pn, rn, None, "3.0dB")
        with np.errstate(all="ignore"):
            lin = 10 ** (Sv / 10) - 10 ** (nz["Sv_noise"] / 10)
            c0 = 10 * np.log10(np.where(lin > 0, lin, np.nan))
            marg = np.nan_to_num(np.abs(c0 - nz["Sv_noise"] - 3.0), nan=np.inf)
        Sv = nz["Sv_corrected"]
    pt = ed_s["Sonar/Beam_group1"]["ping_time"].values
    mv = ogrid.compute_MVBS(Sv, rng, _ns(pt), range_bin="20m", ping_time_bin=f"{bin_pings}s")
    b0 = p0 // bin_pings
    got = ds_full["Sv"].values[:, b0 : b0 + mv["Sv"].shape[1], :]
    want = mv["Sv"]
    nR = min(got.shape[2], want.shape[2])  # the slice may not reach the volume's maximum range (NaN tails)
    got, want = got[:, :, :nR], want[:, :, :nR]
    # bins holding a sample within 1e-3 dB of the strict SNR threshold may differ by that sample
    marg_bins = np.zeros(want.shape, bool)
    if marg.size and (marg < 1e-3).any():
        xc = ogrid.bin_codes(_ns(pt), mv["p_edges"], "left")
        rc = ogrid.bin_codes(rng, mv["r_edges"], "left")
        c, p, n = np.nonzero(marg < 1e-3)
        ok = (xc[p] >= 0) & (rc[c, p, n] >= 0) & (rc[c, p, n] < nR)
        marg_bins[c[ok], xc[p][ok], rc[c, p, n][ok]] = True
    strict = ~marg_bins
    assert np.array_equal(np.isnan(got[strict]), np.isnan(want[strict]))
    ok = strict & ~np.isnan(want)
    d = np.abs(got[ok] - want[ok])
    assert ok.sum() > 0.95 * (~np.isnan(want)).sum()
    assert d.max() <= ATOL, f"max |dMVBS| = {d.max():.3e} dB over {ok.sum()} bins"
    return float(d.max())


def test_cfg2_ek60_full_size(ep):
    """cfg2: EK60 Sv -> remove_noise(5, 30) -> MVBS(20 m, 20 s) on 4 x 100 000 x 4096."""
    import functools

    from echopype_b200 import synth

    C, P, R = 4, 100_000, 4096
    ed = synth.make_ek60(C, P, R, seed=2000, device=True, nan_tail=0.005)
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s", finalize=True)
    assert ds["Sv"].values.shape[:2] == (C, P // 20)
    make = functools.partial(synth.make_ek60, C=C, R=R, seed=2000)
    for p0 in (0, 43_700, 99_900):
        ed_s = _slice_ed(make, ed, p0, p0 + 100)
        _check_slice(ds, ed_s, "ek60", p0, p0 + 100, 5, 30, 20)
    # size-independent property: members per (channel, ping bin, range bin) = 20 pings x samples whose range falls in
    # the bin, minus NaN-padded samples; total members over the grid in closed form
    raw = ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s", finalize=False)
    acc = raw.attrs["acc"]
    members = float((acc[..., 1] + acc[..., 2]).sum())
    x = ed["Sonar/Beam_group1"]["backscatter_r"].data
    import torch

    nan_total = int(torch.isnan(x).sum())
    edges = np.asarray(ds["echo_range"].values)
    a = 2.56e-4 * 1500.0 / 2
    n_in = int(np.sum((np.arange(R) * a >= edges[0]) & (np.arange(R) * a < edges[-1] + 20.0)))
    assert members == float(C * P * n_in - nan_total)


def test_cfg4_azfp_full_size(ep):
    """cfg4: AZFP compute_Sv + compute_MVBS on 4 x 200 000 x 2048 (no noise removal)."""
    import functools

    from echopype_b200 import synth

    C, P, R = 4, 200_000, 2048
    ed = synth.make_azfp(C, P, R, seed=4000, device=True)
    kw = {"env_params": {"salinity": 30.0, "pressure": 50.0}}
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, range_bin="20m", ping_time_bin="20s", **kw)
    assert ds["Sv"].values.shape[:2] == (C, P // 20)

    def make(P, ping_offset, backscatter):
        e = synth.make_azfp(C, P, R, seed=4000, ping_offset=ping_offset)
        e["Sonar/Beam_group1"]["backscatter_r"] = (("channel", "ping_time", "range_sample"), backscatter)
        return e

    for p0 in (0, 123_460, 199_900):
        ed_s = _slice_ed(make, ed, p0, p0 + 100)
        _check_slice(ds, ed_s, "azfp", p0, p0 + 100, None, None, 20)
    _ = functools


def test_cfg5_ek80_cw_full_size(ep):
    """cfg5: EK80 CW power (one GPT channel: double TVG offset) Sv -> noise -> MVBS on 6 x 1 000 000 x 4096 (98 GB)."""
    from echopype_b200 import synth

    C, P, R = 6, 1_000_000, 4096
    ed = synth.make_ek80(C=C, P=P, R=R, mode="CW", encode="power", device=True, gpt_channel=1, nan_tail=0.005, seed=5000)
    kw = {"waveform_mode": "CW", "encode_mode": "power"}
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s", **kw)
    assert ds["Sv"].values.shape[:2] == (C, P // 20)

    def make(P, ping_offset, backscatter):
        e = synth.make_ek80(C=C, P=P, R=R, mode="CW", encode="power", gpt_channel=1, nan_tail=0.0, seed=5000, ping_offset=ping_offset)
        e["Sonar/Beam_group1"]["backscatter_r"] = (("channel", "ping_time", "range_sample"), backscatter)
        return e

    for p0 in (0, 512_340, 999_900):
        ed_s = _slice_ed(make, ed, p0, p0 + 100)
        _check_slice(ds, ed_s, "ek80", p0, p0 + 100, 5, 30, 20)


def test_cfg3_ek80_bb_full_size(ep):
    """cfg3: EK80 broadband pulse-compressed compute_Sv on 6 x 50 000 x 8192 complex samples x 4 beams (79 GB of input).
    Ping slices of the device volume go through the oracle (scipy complex128 convolution per beam + the Sv epilogue)."""
    import torch

    from echopype_b200 import synth

    C, P, R, B = 6, 50_000, 8192, 4
    ed = synth.make_ek80(C=C, P=P, R=R, B=B, mode="BB", encode="complex", device=True, nan_tail=0.005, seed=3000)
    ds = ep.calibrate.compute_Sv(ed, waveform_mode="BB", encode_mode="complex")
    sv = ds["Sv"].data
    assert tuple(sv.shape) == (C, P, R)
    beam = ed["Sonar/Beam_group1"]
    dims = ("channel", "ping_time", "range_sample", "beam")
    for p0 in (0, 24_998, 49_997):
        p1 = p0 + 3
        e = synth.make_ek80(C=C, P=p1 - p0, R=R, B=B, mode="BB", encode="complex", nan_tail=0.0, seed=3000, ping_offset=p0)
        e["Sonar/Beam_group1"]["backscatter_r"] = (dims, beam["backscatter_r"].data[:, p0:p1].cpu().numpy())
        e["Sonar/Beam_group1"]["backscatter_i"] = (dims, beam["backscatter_i"].data[:, p0:p1].cpu().numpy())
        want = og.ek80(e, "Sv", "BB", "complex")
        got = sv[:, p0:p1].cpu().numpy().astype(np.float64)
        # the north-star tolerance (1e-4 dB) outside nulls of the matched-filter output; RMS-relative amplitude inside
        m, frac_null, worst = og.compare_bb_db(got, want["out"], want["prx"], ATOL, "cfg3 Sv")
        assert frac_null < 0.02, frac_null
        ok = ~np.isnan(got)
        assert np.median(np.abs(got[ok] - want["out"][ok])) < 1e-5
    # size-independent property: the NaN tail of a padded ping is NaN in Sv, everything before it is defined beyond
    # the TVG guard (R' > 0)
    nan_rows = torch.isnan(beam["backscatter_r"].data[..., 0]).any(dim=2)
    assert 0.002 < float(nan_rows.float().mean()) < 0.01
    assert bool((torch.isnan(sv).any(dim=2) | ~nan_rows).all())
    del ds, sv
    torch.cuda.empty_cache()
