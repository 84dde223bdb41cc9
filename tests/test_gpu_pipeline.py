"""GPU parity of the fused pipeline (power -> Sv -> remove_background_noise -> MVBS in one kernel) against
(a) the float64 oracle chain and (b) the three separate API calls.  MVBS bins that contain a sample within 1e-3 dB
of the SNR threshold may differ by that sample's share of the bin mean (it may fall on either side of the strict
'>' test, clean/api.py:487); all other bins agree within 1e-4 dB."""

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
    import echopype_b200 as ep

    return ep


def _ns(t):
    return np.asarray(t).astype("datetime64[ns]").astype(np.int64)


def _oracle_chain(ed, kind, pn, rn, nmax, snr, rb, tb, closed="left", skipna=True, env=None):
    ref = og.ek60(ed, "Sv") if kind == "ek60" else (og.azfp(ed, "Sv", 30.0, 50.0) if kind == "azfp" else og.ek80(ed, "Sv", "CW", "power"))
    Sv, rng = ref["out"], ref["echo_range"]
    pt = ed["Sonar/Beam_group1"]["ping_time"].values
    margin = None
    if pn:
        nz = oclean.remove_background_noise(Sv, rng, ref["sound_absorption"], pn, rn, nmax, snr)
        with np.errstate(all="ignore"):
            lin = 10 ** (Sv / 10) - 10 ** (nz["Sv_noise"] / 10)
            c0 = 10 * np.log10(np.where(lin > 0, lin, np.nan))
            margin = np.fmin(np.nan_to_num(np.abs(c0 - nz["Sv_noise"] - float(snr[:-2])), nan=np.inf),
                             np.nan_to_num(np.abs(Sv - nz["Sv_noise"]), nan=np.inf))
        ref.update(nz)
        Svb = nz["Sv_corrected"]
    else:
        Svb = Sv
    mv = ogrid.compute_MVBS(Svb, rng, _ns(pt), range_bin=rb, ping_time_bin=tb, closed=closed, skipna=skipna)
    # bins holding a marginal sample
    marg_bins = np.zeros(mv["Sv"].shape, bool)
    if margin is not None and (margin < 1e-3).any():
        xc = ogrid.bin_codes(_ns(pt), mv["p_edges"], closed)
        rc = ogrid.bin_codes(rng, mv["r_edges"], closed)
        c, p, n = np.nonzero(margin < 1e-3)
        ok = (xc[p] >= 0) & (rc[c, p, n] >= 0)
        marg_bins[c[ok], xc[p][ok], rc[c, p, n][ok]] = True
    return ref, mv, marg_bins, margin


def _check_mvbs(got, want, marg_bins):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    strict = ~marg_bins
    assert np.array_equal(np.isnan(got[strict]), np.isnan(want[strict])), "MVBS NaN mask differs"
    ok = strict & ~np.isnan(want)
    d = np.abs(got[ok] - want[ok])
    assert d.size == 0 or d.max() <= ATOL, f"max |dMVBS| = {d.max():.3e} dB"
    okm = marg_bins & ~np.isnan(want) & ~np.isnan(got)
    assert (np.abs(got[okm] - want[okm]) <= 0.5).all()
    return float(d.max()) if d.size else 0.0


CASES = [
    # kind, shape, time_varying, ping_num, range_sample_num, noise_max, range_bin, ping_time_bin
    ("ek60", (4, 103, 1000), False, 5, 30, None, "20m", "20s"),
    ("ek60", (2, 60, 2048), True, 5, 30, "-125.0dB", "10m", "7s"),
    ("ek60", (3, 47, 516), False, 10, 20, None, "5m", "1min"),
    ("ek60", (2, 64, 4096), False, 30, 100, None, "50m", "30s"),
    ("ek60", (2, 40, 1000), False, None, None, None, "20m", "20s"),
    ("azfp", (4, 55, 2048), False, None, None, None, "10m", "10s"),
    ("azfp", (2, 40, 512), False, 4, 16, None, "2m", "10s"),
    ("ek80", (3, 36, 1024), False, 6, 40, None, "20m", "12s"),
]


@pytest.mark.parametrize("kind,shape,tv,pn,rn,nmax,rb,tb", CASES)
def test_fused_pipeline_vs_oracle(ep, kind, shape, tv, pn, rn, nmax, rb, tb):
    from echopype_b200 import synth

    kw = {}
    if kind == "ek60":
        ed = synth.make_ek60(*shape, seed=21, nan_tail=0.15, time_varying=tv)
    elif kind == "azfp":
        ed = synth.make_azfp(*shape, seed=22)
        kw = {"env_params": {"salinity": 30.0, "pressure": 50.0}}
    else:
        ed = synth.make_ek80(C=shape[0], P=shape[1], R=shape[2], mode="CW", encode="power", gpt_channel=1, nan_tail=0.15, seed=23)
        kw = {"waveform_mode": "CW", "encode_mode": "power"}
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=pn, range_sample_num=rn, background_noise_max=nmax, SNR_threshold="3.0dB",
                                           range_bin=rb, ping_time_bin=tb, keep=("Sv", "echo_range") + (("Sv_noise", "Sv_corrected") if pn else ()), **kw)
    ref, mv, marg_bins, margin = _oracle_chain(ed, kind, pn, rn, nmax, "3.0dB", rb, tb)
    _check_mvbs(ds["Sv"].values, mv["Sv"], marg_bins)
    np.testing.assert_array_equal(_ns(ds["ping_time"].values), mv["ping_time"])
    np.testing.assert_allclose(ds["echo_range"].values, mv["range"])
    kept = ds.attrs["kept"]
    og.compare_db(kept["Sv"].values, ref["out"], ATOL, "Sv")
    er = kept["echo_range"].values.astype(np.float64)
    assert np.array_equal(np.isnan(er), np.isnan(ref["echo_range"]))
    np.testing.assert_allclose(er, ref["echo_range"], rtol=1.3e-7, equal_nan=True)
    if pn:
        og.compare_db(kept["Sv_noise"].values, ref["Sv_noise"], ATOL, "Sv_noise")
        got_c, want_c = kept["Sv_corrected"].values.astype(np.float64), ref["Sv_corrected"]
        flip = np.isnan(got_c) != np.isnan(want_c)
        assert (margin[flip] < 1e-3).all()
        both = ~np.isnan(got_c) & ~np.isnan(want_c)
        assert np.abs(got_c[both] - want_c[both]).max() <= ATOL
        # the per-tile noise estimate itself
        nz = ds.attrs["noise_estimate"].values
        P = shape[1]
        want_noise = (ref["Sv_noise"] - oclean.transmission_loss(ref["echo_range"], ref["sound_absorption"]))
        for i in range(nz.shape[1]):
            seg = want_noise[:, i * pn:min((i + 1) * pn, P), :]
            with np.errstate(all="ignore"):
                w = np.nanmean(seg.reshape(seg.shape[0], -1), axis=1)
            np.testing.assert_allclose(nz[:, i], w, atol=ATOL, equal_nan=True)


def test_fused_equals_three_calls(ep):
    from echopype_b200 import synth

    ed = synth.make_ek60(3, 80, 1000, seed=5, nan_tail=0.1)
    fused = ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s")
    ds = ep.calibrate.compute_Sv(ed)
    ds = ep.clean.remove_background_noise(ds, ping_num=5, range_sample_num=30)
    ds["Sv"] = ds["Sv_corrected"]
    three = ep.commongrid.compute_MVBS(ds, range_bin="20m", ping_time_bin="20s")
    a, b = fused["Sv"].values, three["Sv"].values
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert np.abs(a[ok] - b[ok]).max() < 5e-3  # threshold-marginal samples may flip between the two float32 formulations
    assert np.median(np.abs(a[ok] - b[ok])) < 1e-5


def test_fused_large_property(ep):
    """Size-independent property at a larger size: constant power => every (ping bin, range bin) of the fused MVBS
    (no noise removal) equals the closed-form Sv averaged over the bin's samples, identical across ping bins."""
    import torch

    from echopype_b200 import synth

    C, P, R = 2, 2000, 4096
    ed = synth.make_ek60(C, P, R, seed=1, nan_tail=0.0)
    beam = ed["Sonar/Beam_group1"]
    beam["backscatter_r"] = (("channel", "ping_time", "range_sample"), torch.full((C, P, R), -70.0, device="cuda"))
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, range_bin="20m", ping_time_bin="100s")
    mv = ds["Sv"].values
    assert np.nanmax(np.abs(mv - mv[:, :1, :])) < 2e-5  # identical pings -> identical ping bins
    ed_small = synth.make_ek60(C, 3, R, seed=1, nan_tail=0.0)
    ed_small["Sonar/Beam_group1"]["backscatter_r"] = (("channel", "ping_time", "range_sample"), np.full((C, 3, R), -70.0, np.float32))
    ref = og.ek60(ed_small, "Sv")
    want = ogrid.compute_MVBS(ref["out"], ref["echo_range"], _ns(ed_small["Sonar/Beam_group1"]["ping_time"].values), "20m", "100s")
    og.compare_db(mv[:, 0, :], want["Sv"][:, 0, :], ATOL, "MVBS")


@pytest.mark.parametrize("chunk", [10, 35, 4096])
def test_streamed_host_volume_equals_resident(ep, chunk):
    """Host-resident volume streamed in (channel, ping-chunk) slabs == the same volume resident on the device
    (accumulators are integer-count exact; float sums differ only by atomics order)."""
    import torch

    from echopype_b200 import synth

    ed = synth.make_ek60(3, 83, 1000, seed=9, nan_tail=0.2)
    kw = dict(ping_num=5, range_sample_num=30, range_bin="20m", ping_time_bin="20s", keep=("Sv_corrected", "echo_range"))
    host = ep.pipeline.compute_Sv_clean_MVBS(ed, chunk_pings=chunk, **kw)
    x = torch.from_numpy(ed["Sonar/Beam_group1"]["backscatter_r"].values).cuda()
    ed_dev = synth.make_ek60(3, 83, 1000, seed=9, backscatter=x)
    dev = ep.pipeline.compute_Sv_clean_MVBS(ed_dev, **kw)
    a, b = host["Sv"].values, dev["Sv"].values
    assert a.shape == b.shape
    np.testing.assert_array_equal(host["echo_range"].values, dev["echo_range"].values)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, atol=2e-5, equal_nan=True)
    for k in ("Sv_corrected", "echo_range"):
        np.testing.assert_array_equal(host.attrs["kept"][k].values, dev.attrs["kept"][k].values)
    np.testing.assert_array_equal(host.attrs["noise_estimate"].values, dev.attrs["noise_estimate"].values)


def test_streamed_short_volume_trims_range_grid(ep):
    """Every ping NaN-padded beyond sample 300: the exact range maximum (and so the number of range bins) is
    smaller than the upper bound the streamed kernel bins against; the grid must be cut back to the exact one."""
    from echopype_b200 import synth

    ed = synth.make_ek60(2, 40, 1000, seed=3, nan_tail=0.0)
    x = ed["Sonar/Beam_group1"]["backscatter_r"].values.copy()
    x[:, :, 300:] = np.nan
    ed = synth.make_ek60(2, 40, 1000, seed=3, backscatter=x)
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, range_bin="10m", ping_time_bin="20s", chunk_pings=16)
    ref, mv, marg_bins, _ = _oracle_chain(ed, "ek60", None, None, None, "3.0dB", "10m", "20s")
    assert ds["Sv"].values.shape == mv["Sv"].shape
    _check_mvbs(ds["Sv"].values, mv["Sv"], marg_bins)


FAST_CASES = [
    # kind, (C, P, R), time_varying, ping_num, range_sample_num, noise_max, range_bin, ping_time_bin, closed
    ("ek60", (4, 103, 1000), False, 5, 30, None, "20m", "20s", "left"),     # partial last tile, R4 not a warp multiple
    ("ek60", (2, 200, 4096), False, 5, 30, None, "20m", "20s", "left"),     # the benchmark shape (1024 threads)
    ("ek60", (2, 97, 2048), False, 5, 30, "-125.0dB", "10m", "7s", "right"),  # ping bins straddle tiles, noise cap
    ("ek60", (3, 64, 512), False, 8, 16, None, "5m", "1min", "left"),       # largest register tile
    ("ek60", (3, 50, 516), False, 1, 20, None, "5m", "3s", "left"),         # single-row tiles
    ("ek60", (2, 61, 2048), True, 5, 30, None, "10m", "7s", "left"),        # irregular (law changes) -> general kernel via gate
    ("ek60", (2, 45, 1000), False, None, None, None, "20m", "20s", "left"),  # no noise removal
    ("azfp", (4, 55, 2048), False, None, None, None, "10m", "10s", "left"),
    ("azfp", (2, 43, 512), False, 4, 16, None, "2m", "10s", "left"),
    ("ek80", (3, 37, 1024), False, 6, 40, None, "20m", "12s", "left"),      # GPT channel with the double TVG offset
    ("ek60", (2, 83, 1024), False, 5, 30, None, "0.3m", "9s", "left"),      # bins narrower than a column group: per-column flush
    ("ek60", (2, 83, 1024), False, 3, 4, None, "1m", "9s", "right"),        # smallest range tile of the fast path; 5-sample bins
    ("ek60", (2, 300, 640), False, 5, 7, None, "50m", "10min", "left"),     # cells longer than the packed counters (flush by rows)
    # ping_num > 8: two sweeps over sub-tiles (the reference's own test setting is ping_num=10, range_sample_num=20)
    ("ek60", (3, 205, 4096), False, 10, 20, None, "20m", "20s", "left"),    # 2 sub-tiles of 5; P % 10 != 0: short last noise tile
    ("ek60", (2, 97, 1024), False, 9, 30, "-125.0dB", "10m", "7s", "left"),  # 5 + 4 rows
    ("ek60", (2, 140, 2048), False, 17, 25, None, "5m", "13s", "right"),    # 3 sub-tiles of 6 (6 + 6 + 5)
    ("ek80", (3, 131, 1024), False, 32, 40, None, "20m", "12s", "left"),    # 4 sub-tiles of 8; P ends inside a sub-tile
    ("ek60", (2, 61, 2048), True, 12, 30, None, "10m", "7s", "left"),       # law changes inside noise tiles -> general kernel via gate
    ("azfp", (2, 143, 512), False, 20, 16, None, "2m", "10s", "left"),      # 3 sub-tiles of 7 (7 + 7 + 6)
    # rows of 4097 .. 8192 samples: four column groups per thread, single-row tiles, the noise tile streams twice
    ("ek60", (2, 103, 8192), False, 5, 30, None, "20m", "20s", "left"),     # cfg3-sized rows; partial last noise tile
    ("ek60", (2, 64, 6000), False, 10, 20, "-125.0dB", "10m", "7s", "right"),  # threads past the row end; ping bins straddle
    ("ek60", (2, 45, 8192), False, None, None, None, "20m", "20s", "left"),  # no noise removal
    ("ek80", (2, 37, 5120), False, 3, 40, None, "20m", "12s", "left"),
    ("ek60", (2, 41, 8192), True, 5, 30, None, "50m", "20s", "left"),       # law changes -> general kernel via gate
]


@pytest.mark.parametrize("kind,shape,tv,pn,rn,nmax,rb,tb,closed", FAST_CASES)
def test_fast_kernel_equals_general_kernel(ep, kind, shape, tv, pn, rn, nmax, rb, tb, closed):
    """The persistent register-resident kernel (pipeline_fast_impl.cuh) against the general kernel (pipeline.cu) on the raw
    accumulators: member / NaN-member counts bit-identical, linear sums within float32 accumulation noise."""
    from echopype_b200 import synth

    kw = {}
    if kind == "ek60":
        ed = synth.make_ek60(*shape, seed=31, nan_tail=0.2, time_varying=tv)
    elif kind == "azfp":
        ed = synth.make_azfp(*shape, seed=32)
        kw = {"env_params": {"salinity": 30.0, "pressure": 50.0}}
    else:
        ed = synth.make_ek80(C=shape[0], P=shape[1], R=shape[2], mode="CW", encode="power", gpt_channel=1, nan_tail=0.2, seed=33)
        kw = {"waveform_mode": "CW", "encode_mode": "power"}
    args = dict(ping_num=pn, range_sample_num=rn, background_noise_max=nmax, range_bin=rb, ping_time_bin=tb, closed=closed,
                finalize=False, **kw)
    a = ep.pipeline.compute_Sv_clean_MVBS(ed, fast=True, **args)
    b = ep.pipeline.compute_Sv_clean_MVBS(ed, fast=False, **args)
    fa, fb = a.attrs["acc"].cpu().numpy(), b.attrs["acc"].cpu().numpy()
    assert fa.shape == fb.shape
    if pn:
        np.testing.assert_allclose(a.attrs["noise_estimate"].values, b.attrs["noise_estimate"].values, atol=2e-5, equal_nan=True)
    total = fa[..., 1] + fa[..., 2]
    np.testing.assert_array_equal(total, fb[..., 1] + fb[..., 2])  # members per bin: exact
    # the SNR test may flip for samples within float rounding of the threshold: allow a handful per grid
    flips = np.abs(fa[..., 1] - fb[..., 1])
    assert flips.sum() <= max(2, 2e-5 * total.sum()), (flips.sum(), total.sum())
    ok = (flips == 0) & (fb[..., 1] > 0)
    np.testing.assert_allclose(fa[..., 0][ok], fb[..., 0][ok], rtol=2e-5)
    assert ok.sum() > 0.9 * (fb[..., 1] > 0).sum()


KEEP_CASES = [
    # kind, (C, P, R), time_varying, ping_num, range_sample_num
    ("ek60", (2, 103, 4096), False, 5, 30),    # benchmark geometry: two column groups per thread
    ("ek60", (3, 47, 1000), False, 8, 16),     # one group, R / 4 not a warp multiple, partial last tile
    ("ek60", (2, 205, 2048), False, 10, 20),   # two sweeps (the reference's own setting)
    ("ek60", (2, 45, 1000), False, None, None),  # no noise removal: Sv / echo_range only
    ("azfp", (2, 43, 512), False, 4, 16),
    ("ek80", (3, 37, 1024), False, 6, 40),     # GPT channel
    ("ek60", (2, 61, 2048), True, 5, 30),      # law changes inside tiles: the gate hands the volume to the general kernel
]


@pytest.mark.parametrize("kind,shape,tv,pn,rn", KEEP_CASES)
def test_fast_kernel_keep_outputs_equal_general_kernel(ep, kind, shape, tv, pn, rn):
    """keep= through the persistent kernel (kKeep instantiations stream Sv / echo_range / Sv_noise / Sv_corrected out of
    the register tile) against the general kernel: echo_range bit-identical, dB arrays within float32 rounding, NaN
    masks identical except for samples within rounding of the SNR threshold."""
    from echopype_b200 import synth

    kw = {}
    if kind == "ek60":
        ed = synth.make_ek60(*shape, seed=41, nan_tail=0.2, time_varying=tv)
    elif kind == "azfp":
        ed = synth.make_azfp(*shape, seed=42)
        kw = {"env_params": {"salinity": 30.0, "pressure": 50.0}}
    else:
        ed = synth.make_ek80(C=shape[0], P=shape[1], R=shape[2], mode="CW", encode="power", gpt_channel=1, nan_tail=0.2, seed=43)
        kw = {"waveform_mode": "CW", "encode_mode": "power"}
    keep = ("Sv", "echo_range") + (("Sv_noise", "Sv_corrected") if pn else ())
    args = dict(ping_num=pn, range_sample_num=rn, range_bin="20m", ping_time_bin="20s", keep=keep, **kw)
    a = ep.pipeline.compute_Sv_clean_MVBS(ed, fast=True, **args)
    b = ep.pipeline.compute_Sv_clean_MVBS(ed, fast=False, **args)
    ka, kb = a.attrs["kept"], b.attrs["kept"]
    np.testing.assert_array_equal(ka["echo_range"].values, kb["echo_range"].values)
    for k in ("Sv",) + (("Sv_noise",) if pn else ()):
        x, y = ka[k].values, kb[k].values
        assert np.array_equal(np.isnan(x), np.isnan(y)), k
        np.testing.assert_allclose(x, y, atol=1e-4, equal_nan=True, err_msg=k)  # a few float32 ulps at |dB| ~ 200
    if pn:
        x, y = ka["Sv_corrected"].values, kb["Sv_corrected"].values
        flip = np.isnan(x) != np.isnan(y)
        assert flip.sum() <= max(2, 2e-5 * x.size), flip.sum()
        both = ~np.isnan(x) & ~np.isnan(y)
        # Sv_corrected = dB(v - noise): cancellation amplifies the float32 rounding of v and noise near the threshold
        # (v / (v - noise) <= 1 / (1 - 10^-0.3) = 2 at the 3 dB SNR threshold)
        assert np.abs(x[both] - y[both]).max() <= 2e-4
    np.testing.assert_allclose(a["Sv"].values, b["Sv"].values, atol=2e-4, equal_nan=True)


@pytest.mark.parametrize("pn,rn", [(5, 30), (10, 20)])
def test_fast_kernel_vs_oracle_benchmark_shape(ep, pn, rn):
    """Benchmark-shaped tile geometry (R = 4096, 20 s bins) through the fast kernel against the oracle: ping_num 5 (u in
    registers) and the reference's own test setting ping_num=10, range_sample_num=20 (two sweeps over sub-tiles)."""
    from echopype_b200 import synth

    ed = synth.make_ek60(2, 120, 4096, seed=77, nan_tail=0.1)
    ds = ep.pipeline.compute_Sv_clean_MVBS(ed, ping_num=pn, range_sample_num=rn, range_bin="20m", ping_time_bin="20s")
    ref, mv, marg_bins, margin = _oracle_chain(ed, "ek60", pn, rn, None, "3.0dB", "20m", "20s")
    _check_mvbs(ds["Sv"].values, mv["Sv"], marg_bins)


@pytest.mark.gpu
@pytest.mark.parametrize("windows", [[(0, 4), (4, 7), (7, 9)], [(0, 3), (3, 3), (3, 6)], [(2, 5), (-1, -2), (6, 8)]])
def test_straddle_pack_unpack_kernels_equal_global_sums(ep, windows):
    """The NCCL path of the straddling-bin exchange (epb_straddle_pack -> all-reduce(sum) -> epb_straddle_unpack) on one
    GPU: the all-reduce is emulated by adding the packed buffers of three fake ranks; every rank must end up with the
    global sums in its first / last ping bin and with the global range maximum."""
    import torch

    from echopype_b200 import _lib
    from echopype_b200.device import ptr, stream
    from echopype_b200.pipeline import straddle_slots

    rng = np.random.default_rng(7)
    C, nR, world = 2, 5, len(windows)
    accs, rmaxs, bufs = [], [], []
    W = 2 * C * nR * 4 + 1
    for r, (lo, hi) in enumerate(windows):
        nXl = max(hi - lo + 1, 1)
        a = rng.random((C, nXl, nR, 4)) if hi >= lo else np.zeros((C, 1, nR, 4))
        accs.append(torch.from_numpy(a).cuda())
        rmaxs.append(torch.tensor([10.0 * r + 1.0, 5.0 - r], dtype=torch.float64, device="cuda"))
        buf = torch.full((world * W,), 7.0, dtype=torch.float64, device="cuda")  # stale contents must be overwritten
        _lib.call("epb_straddle_pack", ptr(accs[r]), C, nXl, nR, ptr(rmaxs[r]), 2, ptr(buf), r, world, int(hi > lo), stream())
        bufs.append(buf)
    total = torch.stack(bufs).sum(0)
    glob = {}  # global bin -> sum over the ranks that hold it
    for r, (lo, hi) in enumerate(windows):
        for b in range(lo, hi + 1):
            glob[b] = glob.get(b, 0) + accs[r][:, b - lo].cpu().numpy()
    idx = [w if w[1] >= w[0] else (-1, -1) for w in windows]
    for r, (lo, hi) in enumerate(windows):
        slots = straddle_slots(idx, lo, hi)
        src = np.zeros((2, world), dtype=np.int32)
        n = [0, 0]
        for slot, lst in slots:
            k = 0 if slot == 0 else 1
            src[k, : len(lst)], n[k] = lst, len(lst)
        acc = accs[r].clone()
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        _lib.call("epb_straddle_unpack", ptr(total), ptr(torch.from_numpy(src).cuda()), n[0], n[1], C, acc.shape[1], nR, world,
                  ptr(acc), ptr(out), stream())
        assert float(out) == 10.0 * (world - 1) + 1.0
        got = acc.cpu().numpy()
        for b in range(lo, hi + 1):
            np.testing.assert_allclose(got[:, b - lo], glob[b], rtol=1e-15)
