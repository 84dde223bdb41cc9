"""GPU parity for clean.remove_background_noise / estimate_background_noise and commongrid.compute_MVBS /
compute_MVBS_index_binning / compute_NASC (public API -> ctypes C-ABI -> sm_100a kernels) against the float64
CPU oracle.

Tolerances (float32 device path vs float64 oracle): 1e-4 dB on every finite value; NaN masks identical,
except that a sample whose oracle SNR margin |Sv_corrected - Sv_noise - SNR| (or |Sv - Sv_noise|) is below
1e-3 dB may fall on either side of the strict '>' test of clean/api.py:485-487."""

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


def _noise_margin(Sv, Sv_noise, snr):
    with np.errstate(all="ignore"):
        lin = 10 ** (Sv / 10) - 10 ** (Sv_noise / 10)
        c0 = 10 * np.log10(np.where(lin > 0, lin, np.nan))
        m1 = np.abs(c0 - Sv_noise - snr)
        m2 = np.abs(Sv - Sv_noise)
    return np.fmin(np.where(np.isnan(m1), np.inf, m1), np.where(np.isnan(m2), np.inf, m2))


def check_corrected(got, want, margin, atol=ATOL):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    flip = np.isnan(got) != np.isnan(want)
    assert (margin[flip] < 1e-3).all(), f"{int((flip & (margin >= 1e-3)).sum())} NaN-mask flips away from the SNR threshold"
    both = ~np.isnan(got) & ~np.isnan(want)
    d = np.abs(got[both] - want[both])
    assert d.size == 0 or d.max() <= atol, f"max |dSv_corrected| = {d.max():.3e}"
    return int(flip.sum())


def _sv_dataset(ep, shape, seed=3, nan_tail=0.2, time_varying=False):
    from echopype_b200 import synth

    ed = synth.make_ek60(*shape, seed=seed, nan_tail=nan_tail, time_varying=time_varying)
    return ed, ep.calibrate.compute_Sv(ed)


@pytest.mark.parametrize("shape,pn,rn,nmax", [((3, 53, 1000), 5, 30, None), ((2, 40, 515), 10, 20, "-125.0dB"), ((1, 7, 64), 3, 200, None),
                                              ((2, 31, 4096), 30, 100, "-200dB")])
def test_remove_background_noise(ep, shape, pn, rn, nmax):
    ed, ds = _sv_dataset(ep, shape, time_varying=(shape[0] == 2))
    out = ep.clean.remove_background_noise(ds, ping_num=pn, range_sample_num=rn, background_noise_max=nmax, SNR_threshold="3.0dB")
    Sv32 = ds["Sv"].values.astype(np.float64)
    er32 = ds["echo_range"].values.astype(np.float64)
    alpha = np.asarray(ds["sound_absorption"].values, dtype=np.float64)
    want = oclean.remove_background_noise(Sv32, er32, alpha, pn, rn, nmax, "3.0dB")
    og.compare_db(out["Sv_noise"].values, want["Sv_noise"], ATOL, "Sv_noise")
    nflip = check_corrected(out["Sv_corrected"].values, want["Sv_corrected"], _noise_margin(Sv32, want["Sv_noise"], 3.0))
    assert nflip <= max(2, Sv32.size // 20000)
    for name, key in (("Sv_noise", "attrs_noise"), ("Sv_corrected", "attrs_corrected")):
        a, w = out[name].attrs, want[key]
        assert a["long_name"] == w["long_name"] and a["units"] == "dB"
        assert a["noise_ping_num"] == pn and a["noise_range_sample_num"] == rn and a["SNR_threshold"] == 3.0
        assert a["noise_max"] == nmax
        np.testing.assert_allclose(a["actual_range"], w["actual_range"], atol=0.011, equal_nan=True)
    assert out.attrs["processing_function"] == "clean.remove_background_noise"
    # estimate_background_noise returns the bare noise array (clean/api.py:433)
    est = ep.clean.estimate_noise(ds, ping_num=pn, range_sample_num=rn, background_noise_max=nmax)
    np.testing.assert_array_equal(est.values, out["Sv_noise"].values)


def test_noise_toy_known_answer(ep):
    """The reference's own known-answer test (tests/clean/test_noise.py:902-987): spikes survive, and exactly 6
    samples of the seed-1 normal background are NaN in the first 50 samples."""
    from echopype_b200.dataset import Dataset

    np.random.seed(1)
    nchan, npings, nrange_samples = 1, 10, 100
    data = np.ones(nrange_samples)
    data[30], data[60] = 100, 100
    data = np.array([data] * npings)
    Sv = np.array([data] * nchan)
    Sv[0, :, :50] = Sv[0, :, :50] + np.random.normal(loc=100, scale=1, size=(npings, 50))  # stats as the reference
    er = np.array([[np.linspace(0, 10, nrange_samples)] * npings] * nchan)
    ds = Dataset(
        {"Sv": (("channel", "ping_time", "range_sample"), Sv), "echo_range": (("channel", "ping_time", "range_sample"), er),
         "sound_absorption": ((), np.asarray(0.001))},
        coords={"channel": np.array(["c0"], dtype=object), "ping_time": np.arange(npings).astype("datetime64[s]").astype("datetime64[ns]"),
                "range_sample": np.arange(nrange_samples)},
    )
    out = ep.clean.remove_background_noise(ds, ping_num=2, range_sample_num=5, SNR_threshold="3.0dB")
    want = oclean.remove_background_noise(Sv.astype(np.float32), er.astype(np.float32), 0.001, 2, 5, None, "3.0dB")
    assert np.array_equal(np.isnan(out["Sv_corrected"].values), np.isnan(want["Sv_corrected"]))
    og.compare_db(out["Sv_corrected"].values, want["Sv_corrected"], 2e-4, "Sv_corrected")


def _mvbs_oracle(Sv, rng, pt, **kw):
    return ogrid.compute_MVBS(np.asarray(Sv, np.float64), np.asarray(rng, np.float64), _ns(pt), **kw)


@pytest.mark.parametrize("closed", ["left", "right"])
@pytest.mark.parametrize("shape,rb,tb", [((3, 200, 1000), "20m", "20s"), ((2, 333, 515), "7.5m", "1min"), ((1, 41, 4096), "50m", "5s")])
def test_compute_MVBS_law_path(ep, shape, rb, tb, closed):
    """echo_range produced by compute_Sv: bins located from the exact float64 law -> membership identical
    with binning the float64 reference range."""
    ed, ds = _sv_dataset(ep, shape, time_varying=(shape[0] == 2))
    got = ep.commongrid.compute_MVBS(ds, range_bin=rb, ping_time_bin=tb, closed=closed)
    ref = og.ek60(ed, "Sv")  # float64 echo_range of the reference
    pt = ds["ping_time"].values
    rmax = float(np.nanmax(ds["echo_range"].values))  # edge construction uses the float32 extrema
    want = _mvbs_oracle(ds["Sv"].values, ref["echo_range"], pt, range_bin=rb, ping_time_bin=tb, closed=closed,
                        range_var_max=None)
    r_edges = np.arange(0, rmax + float(rb[:-1]), float(rb[:-1]))
    assert got["Sv"].shape[1] == len(want["ping_time"])
    nR = min(got["Sv"].shape[2], want["Sv"].shape[2])
    assert abs(got["Sv"].shape[2] - want["Sv"].shape[2]) <= 1 and len(r_edges) - 1 == got["Sv"].shape[2]
    og.compare_db(got["Sv"].values[:, :, :nR], want["Sv"][:, :, :nR], ATOL, "MVBS")
    np.testing.assert_array_equal(_ns(got["ping_time"].values), want["ping_time"])
    np.testing.assert_allclose(got["echo_range"].values[:nR], want["range"][:nR])
    assert got["Sv"].attrs["binning_mode"] == "physical units"
    assert got["Sv"].attrs["ping_time_interval"] == tb
    assert got.attrs["processing_function"] == "commongrid.compute_MVBS"


@pytest.mark.parametrize("skipna", [True, False])
@pytest.mark.parametrize("f64", [True, False])
def test_compute_MVBS_generic_path(ep, skipna, f64):
    """User-supplied (non-law) irregular range arrays, float64 and float32, incl. NaN coordinates and fill_value."""
    from echopype_b200.dataset import Dataset

    g = np.random.default_rng(5)
    C, P, R = 2, 120, 300
    Sv = g.uniform(-90, -40, (C, P, R))
    Sv[g.random((C, P, R)) < 0.05] = np.nan
    rng = np.cumsum(g.uniform(0.05, 0.4, (C, P, R)), axis=2)
    rng[0, 5, 200:] = np.nan
    if not f64:
        rng = rng.astype(np.float32)
    pt = (np.datetime64("2020-01-01T00:00:03", "ns") + (np.cumsum(g.integers(200, 1500, P)) * 1_000_000).astype("timedelta64[ns]"))
    ds = Dataset(
        {"Sv": (("channel", "ping_time", "range_sample"), Sv), "echo_range": (("channel", "ping_time", "range_sample"), rng),
         "frequency_nominal": (("channel",), np.array([38e3, 120e3]))},
        coords={"channel": np.array(["a", "b"], dtype=object), "ping_time": pt, "range_sample": np.arange(R)},
    )
    got = ep.commongrid.compute_MVBS(ds, range_bin="5m", ping_time_bin="10s", skipna=skipna, fill_value=-999.0 if skipna else np.nan)
    want = _mvbs_oracle(Sv.astype(np.float32), rng, pt, range_bin="5m", ping_time_bin="10s", skipna=skipna,
                        fill_value=-999.0 if skipna else np.nan)
    assert got["Sv"].shape == want["Sv"].shape
    og.compare_db(got["Sv"].values, want["Sv"], ATOL, "MVBS")


@pytest.mark.parametrize("case", ["narrow_bins", "non_monotone", "closed_right_wide", "gaps_and_outside"])
def test_compute_MVBS_generic_path_stress(ep, case):
    """The staged generic kernel keeps two adjacent bins per thread in registers and falls back to an exact per-sample path
    when a thread's chunk of samples holds a third bin: bins narrower than the chunk, a range variable that is not
    monotone, samples outside the grid, rows wider than one pass."""
    from echopype_b200.dataset import Dataset

    g = np.random.default_rng({"narrow_bins": 1, "non_monotone": 2, "closed_right_wide": 3, "gaps_and_outside": 4}[case])
    C, P, R = (2, 45, 4096) if case != "closed_right_wide" else (1, 30, 8192)
    Sv = g.uniform(-90, -40, (C, P, R)).astype(np.float32)
    Sv[g.random((C, P, R)) < 0.03] = np.nan
    rng = np.cumsum(g.uniform(0.01, 0.05, (C, P, R)), axis=2).astype(np.float32)  # ~120 m
    kw = dict(range_bin="20m", ping_time_bin="10s")
    if case == "narrow_bins":
        kw["range_bin"] = "0.2m"  # ~7 samples per bin: every 16-sample chunk holds three or more bins
    elif case == "non_monotone":
        rng = g.permuted(rng, axis=2).astype(np.float32)
        kw["range_bin"] = "7m"
    elif case == "closed_right_wide":
        kw.update(range_bin="3m", closed="right")
        rng[0, :, ::97] = np.round(rng[0, :, ::97] / 3.0) * 3.0  # samples exactly on bin edges
    else:
        rng[:, ::3, 1000:1500] = np.nan          # NaN coordinates in the middle of rows
        rng[1, :, :300] -= 5.0                     # negative ranges: below the first edge
        kw["range_var_max"] = "63m"                # samples beyond the grid
    pt = np.datetime64("2020-01-01T00:00:03", "ns") + (np.arange(P) * 1_300_000_000).astype("timedelta64[ns]")
    ds = Dataset(
        {"Sv": (("channel", "ping_time", "range_sample"), Sv), "echo_range": (("channel", "ping_time", "range_sample"), rng),
         "frequency_nominal": (("channel",), np.array([38e3, 120e3][:C]))},
        coords={"channel": np.array(["a", "b"][:C], dtype=object), "ping_time": pt, "range_sample": np.arange(R)},
    )
    got = ep.commongrid.compute_MVBS(ds, **kw)
    want = _mvbs_oracle(Sv, rng, pt, **kw)
    assert got["Sv"].shape == want["Sv"].shape
    og.compare_db(got["Sv"].values, want["Sv"], ATOL, "MVBS " + case)


def test_compute_MVBS_validation(ep):
    ed, ds = _sv_dataset(ep, (1, 10, 64))
    with pytest.raises(ValueError, match="range_var must be one of 'echo_range' or 'depth'."):
        ep.commongrid.compute_MVBS(ds, range_var="foo")
    with pytest.raises(ValueError, match="Input Sv dataset must contain all of"):
        ep.commongrid.compute_MVBS(ds, range_var="depth")
    with pytest.raises(TypeError, match="range_bin must be a string"):
        ep.commongrid.compute_MVBS(ds, range_bin=10)
    with pytest.raises(ValueError, match=r"Range bin must be in meters \(e.g., '10m'\)."):
        ep.commongrid.compute_MVBS(ds, range_bin="10km")
    with pytest.raises(TypeError, match="ping_time_bin must be a string"):
        ep.commongrid.compute_MVBS(ds, ping_time_bin=10)
    with pytest.raises(ValueError, match="is not a valid option. Options are 'left' or 'right'."):
        ep.commongrid.compute_MVBS(ds, closed="both")
    with pytest.raises(ValueError, match="is only allowed when method='map_reduce'"):
        ep.commongrid.compute_MVBS(ds, method="cohorts", reindex=True)
    got = ep.commongrid.compute_MVBS(ds, range_var_max="30m", range_bin="10m")
    np.testing.assert_array_equal(got["echo_range"].values, [0.0, 10.0, 20.0, 30.0])


@pytest.mark.parametrize("shape,pn,rn", [((2, 50, 1000), 10, 100), ((1, 33, 257), 7, 50)])
def test_index_binning(ep, shape, pn, rn):
    ed, ds = _sv_dataset(ep, shape)
    got = ep.commongrid.compute_MVBS_index_binning(ds, range_sample_num=rn, ping_num=pn)
    want = ogrid.compute_MVBS_index_binning(ds["Sv"].values, ds["echo_range"].values, rn, pn)
    og.compare_db(got["Sv"].values, want["Sv"], ATOL, "MVBS")
    np.testing.assert_allclose(got["echo_range"].values, want["echo_range"], rtol=1e-7, equal_nan=True)
    np.testing.assert_allclose(got["Sv"].attrs["actual_range"], want["actual_range"], atol=0.011)
    assert got["Sv"].attrs["binning_mode"] == "sample number"


def test_compute_NASC(ep):
    from echopype_b200.dataset import Dataset

    g = np.random.default_rng(9)
    C, P, R = 2, 90, 400
    Sv = g.uniform(-90, -40, (C, P, R))
    Sv[g.random((C, P, R)) < 0.03] = np.nan
    depth = 3.0 + np.cumsum(np.full((C, P, R), 0.25) + g.uniform(0, 1e-3, (C, P, R)), axis=2)
    lat = 42.0 + np.cumsum(g.uniform(1e-4, 4e-4, P))
    lon = -124.0 + np.cumsum(g.uniform(-1e-4, 3e-4, P))
    lat[7] = np.nan
    pt = np.datetime64("2021-05-05T10:00:00", "ns") + (np.arange(P) * 2_000_000_000).astype("timedelta64[ns]")
    ds = Dataset(
        {"Sv": (("channel", "ping_time", "range_sample"), Sv), "depth": (("channel", "ping_time", "range_sample"), depth),
         "latitude": (("ping_time",), lat), "longitude": (("ping_time",), lon), "frequency_nominal": (("channel",), np.array([38e3, 120e3]))},
        coords={"channel": np.array(["a", "b"], dtype=object), "ping_time": pt, "range_sample": np.arange(R)},
    )
    got = ep.commongrid.compute_NASC(ds, range_bin="10m", dist_bin="0.02nmi")
    want = ogrid.compute_NASC(Sv.astype(np.float32), depth, lat, lon, _ns(pt), "10m", "0.02nmi")
    assert got["NASC"].shape == want["NASC"].shape
    np.testing.assert_allclose(got["NASC"].values, want["NASC"], rtol=3e-5, equal_nan=True)  # 1e-4 dB = 2.3e-5 relative
    np.testing.assert_allclose(got["distance"].values, want["distance"], rtol=1e-9)
    assert got["NASC"].attrs["units"] == "m2 nmi-2"
    assert got.attrs["Conventions"] == "CF-1.7,ACDD-1.3"


@pytest.mark.parametrize("shape,rb,tb,closed", [((3, 97, 1000), "20m", "20s", "left"), ((2, 150, 4096), "10m", "7s", "right"), ((2, 40, 516), "5m", "1min", "left")])
def test_bin_reduce_law_fast_equals_general(ep, shape, rb, tb, closed):
    """compute_MVBS on library-produced echo_range: the persistent kernel (Sv input) against the warp-per-row kernel on the raw
    accumulators: counts identical, sums within float32 accumulation noise."""
    import torch

    from echopype_b200 import kernels, synth
    from echopype_b200.commongrid.utils import assign_bins, ping_time_edges, range_edges

    C, P, R = shape
    ed = synth.make_ek60(C, P, R, seed=41, nan_tail=0.2)
    ds = ep.calibrate.compute_Sv(ed)
    rows = ds["echo_range"].law["rows"]
    Sv_t = ds["Sv"].data
    pt = ds["ping_time"].values
    rmax = kernels.range_max(ed["Sonar/Beam_group1"]["backscatter_r"].data if False else None, rows, C, P, R)
    r_edges = range_edges(rmax, float(rb[:-1]))
    p_edges = ping_time_edges(pt, tb)
    xb = torch.from_numpy(assign_bins(pt, p_edges, closed)).cuda()
    et = torch.from_numpy(r_edges).cuda()
    nX = len(p_edges) - 1
    a = kernels.bin_reduce_law(Sv_t, rows, xb, et, kernels.new_acc(C, nX, len(r_edges) - 1), C, P, R, nX, closed_right=(closed == "right"), fast=True)
    b = kernels.bin_reduce_law(Sv_t, rows, xb, et, kernels.new_acc(C, nX, len(r_edges) - 1), C, P, R, nX, closed_right=(closed == "right"), fast=False)
    fa, fb = a.cpu().numpy(), b.cpu().numpy()
    np.testing.assert_array_equal(fa[..., 1], fb[..., 1])
    np.testing.assert_array_equal(fa[..., 2], fb[..., 2])
    ok = fb[..., 1] > 0
    np.testing.assert_allclose(fa[..., 0][ok], fb[..., 0][ok], rtol=2e-5)
