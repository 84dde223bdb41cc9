"""GPU parity: compute_Sv / compute_TS through the public API (ctypes C-ABI -> sm_100a kernels) against the
float64 CPU oracle on the same synthetic inputs.  Tolerance (north_star): Sv within 1e-4 dB, identical NaN
masks; echo_range within 1 float32 ulp (relative 1.2e-7) of the float64 reference."""

import numpy as np
import pytest

import oracle_glue as og

pytestmark = pytest.mark.gpu

SV_ATOL = 1e-4  # dB, float32 device path vs float64 oracle
RANGE_RTOL = 1.3e-7


@pytest.fixture(scope="module")
def ep():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import echopype_b200 as ep

    return ep


def _check_range(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "echo_range NaN mask differs"
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok])
    assert (err <= RANGE_RTOL * np.abs(want[ok]) + 1e-12).all(), f"echo_range max rel err {np.max(err / np.maximum(np.abs(want[ok]), 1e-30)):.3e}"


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("shape,time_varying", [((4, 64, 1000), False), ((3, 37, 515), True), ((2, 5, 4096), True), ((1, 1, 7), False)])
def test_ek60_power(ep, cal_type, shape, time_varying):
    from echopype_b200 import synth

    C, P, R = shape
    ed = synth.make_ek60(C, P, R, seed=11 + R, nan_tail=0.2, time_varying=time_varying)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed)
    want = og.ek60(ed, cal_type)
    assert ds[cal_type].dims == ("channel", "ping_time", "range_sample")
    assert ds[cal_type].on_device and ds[cal_type].dtype.__str__().endswith("float32")
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])
    if R > 3:
        assert np.isnan(ds[cal_type].values[:, :, :3]).all()  # R' <= 0 for n in {0,1,2} (calibrate_ek.py:107)


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("shape", [(4, 50, 2048), (2, 9, 333)])
def test_azfp(ep, cal_type, shape):
    from echopype_b200 import synth

    ed = synth.make_azfp(*shape, seed=5)
    env = {"salinity": 30.0, "pressure": 50.0}
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, env_params=env)
    want = og.azfp(ed, cal_type, 30.0, 50.0)
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("gpt", [None, 1])
def test_ek80_cw_power(ep, cal_type, gpt):
    from echopype_b200 import synth

    ed = synth.make_ek80(C=3, P=21, R=640, mode="CW", encode="power", gpt_channel=gpt, nan_tail=0.2, seed=77)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, waveform_mode="CW", encode_mode="power")
    want = og.ek80(ed, cal_type, "CW", "power")
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])
    if cal_type == "Sv":
        np.testing.assert_allclose(ds["tau_effective"].values, want["tau_effective"], rtol=1e-12)


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("B", [4, 3, 1])
def test_ek80_cw_complex(ep, cal_type, B):
    from echopype_b200 import synth

    ed = synth.make_ek80(C=2, P=13, R=500, B=B, mode="CW", encode="complex", gpt_channel=None, nan_tail=0.2, seed=78)
    beam = ed["Sonar/Beam_group1"]
    re = beam["backscatter_r"].values.copy()  # single-beam NaNs: nanmean over beams; beam 0 also voids echo_range / Sv
    rs = np.random.default_rng(11)
    for _ in range(60):
        re[rs.integers(2), rs.integers(13), rs.integers(500), rs.integers(B)] = np.nan
    beam["backscatter_r"] = (("channel", "ping_time", "range_sample", "beam"), re)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, waveform_mode="CW", encode_mode="complex")
    want = og.ek80(ed, cal_type, "CW", "complex")
    og.compare_db(ds[cal_type].values, want["out"], SV_ATOL, cal_type)
    _check_range(ds["echo_range"].values, want["echo_range"])


BB_ATOL = 1e-4  # north-star tolerance outside nulls of the matched-filter output (oracle_glue.compare_bb_db)


@pytest.mark.parametrize("cal_type", ["Sv", "TS"])
@pytest.mark.parametrize("B,R,beam_nan", [(4, 512, False), (4, 500, False), (4, 301, True), (3, 256, False), (1, 200, True)])
def test_ek80_bb_pulse_compression(ep, cal_type, B, R, beam_nan):
    """EK80 broadband: matched filter + Sv/TS epilogue (K3) against scipy's convolution in complex128.  beam_nan pokes NaNs
    into single beams so that the per-beam fallback (exact nanmean over beams) runs."""
    from echopype_b200 import synth

    ed = synth.make_ek80(C=2, P=9, R=R, B=B, mode="BB", encode="complex", nan_tail=0.3, seed=91)
    beam = ed["Sonar/Beam_group1"]
    if beam_nan:
        re, im = beam["backscatter_r"].values.copy(), beam["backscatter_i"].values.copy()
        rng = np.random.default_rng(5)
        for _ in range(40):
            c, p, n, b = rng.integers(2), rng.integers(9), rng.integers(R), rng.integers(B)
            re[c, p, n, b] = np.nan
            if rng.random() < 0.5:
                im[c, p, n, b] = np.nan
        dims = ("channel", "ping_time", "range_sample", "beam")
        beam["backscatter_r"] = (dims, re)
        beam["backscatter_i"] = (dims, im)
    fn = ep.calibrate.compute_Sv if cal_type == "Sv" else ep.calibrate.compute_TS
    ds = fn(ed, waveform_mode="BB", encode_mode="complex")
    want = og.ek80(ed, cal_type, "BB", "complex")
    got = ds[cal_type].values.astype(np.float64)
    og.compare_bb_db(got, want["out"], want["prx"], BB_ATOL, cal_type)
    ok = ~np.isnan(got)
    assert np.median(np.abs(got[ok] - want["out"][ok])) < 2e-5
    _check_range(ds["echo_range"].values, want["echo_range"])
    if cal_type == "Sv":
        np.testing.assert_allclose(ds["tau_effective"].values, want["tau_effective"], rtol=1e-12)
    # "FM" is an alias of "BB" (calibrate/api.py:35)
    ds2 = fn(ed, waveform_mode="FM", encode_mode="complex")
    np.testing.assert_array_equal(ds2[cal_type].values, ds[cal_type].values)


@pytest.mark.parametrize("method", ["fft", "direct"])
def test_pulse_compress_kernel_vs_scipy(ep, method):
    """The compressed, normalised, beam-averaged signal itself (pc_out) against scipy.signal.convolve: the overlap-save
    FFT kernel and the direct tap-loop kernel."""
    import torch
    from scipy import signal

    from echopype_b200 import kernels, synth

    C, P, R, B = 2, 5, (777 - 1 if method == "direct" else 9001), 4  # 9001: three FFT segments, the last one short
    rng = np.random.default_rng(3)
    re = rng.standard_normal((C, P, R, B)).astype(np.float32)
    im = rng.standard_normal((C, P, R, B)).astype(np.float32)
    re[0, 1, 500:] = np.nan
    im[0, 1, 500:] = np.nan
    tx = [rng.standard_normal(37) + 1j * rng.standard_normal(37), rng.standard_normal(150) + 1j * rng.standard_normal(150)]
    ed = synth.make_ek80(C=C, P=P, R=R, B=B, mode="BB", encode="complex", nan_tail=0.0, seed=1)
    from echopype_b200.calibrate.calibrate_ek import CalibrateEK80

    cal = CalibrateEK80(ed, waveform_mode="BB", encode_mode="complex")
    cal._cal_complex_samples("Sv")
    re[1, 2, 100:140, 2] = np.nan  # one beam missing on a stretch: the per-beam path (exact nanmean over beams)
    out, _, pc, _ = kernels.pulse_compress_sv(torch.from_numpy(re).cuda(), torch.from_numpy(im).cuda(), tx, cal.rows, C, P, R, B,
                                              want_pc=True, method=method)
    pc = pc.cpu().numpy()
    got = pc[..., 0] + 1j * pc[..., 1]
    x = re.astype(np.float64) + 1j * im.astype(np.float64)
    nan = np.isnan(x)
    xz = np.where(nan, 0, x)
    for c in range(C):
        h = np.flipud(np.conj(tx[c]))
        norm = np.linalg.norm(tx[c]) ** 2
        for p in range(P):
            y = np.stack([signal.convolve(xz[c, p, :, b], h, mode="full")[len(h) - 1:] for b in range(B)], axis=-1) / norm
            y = np.where(nan[c, p], np.nan, y)
            want = np.nanmean(y, axis=-1) if not nan[c, p].all(axis=-1).any() else np.where(nan[c, p].all(axis=-1), np.nan, np.nanmean(np.where(nan[c, p].all(axis=-1)[:, None], 0, y), axis=-1))
            g = got[c, p]
            assert np.array_equal(np.isnan(g), np.isnan(want))
            ok = ~np.isnan(want)
            scale = np.abs(want[ok]).max()
            rms = np.sqrt(np.mean(np.abs(want[ok]) ** 2))
            tol = (1.0e-6 if method == "fft" else 3e-6) * rms * (1 if method == "fft" else scale / rms)
            assert np.abs(g[ok] - want[ok]).max() <= tol, (method, c, p, np.abs(g[ok] - want[ok]).max() / rms)


def test_output_contract(ep):
    from echopype_b200 import synth

    ed = synth.make_ek60(2, 8, 64)
    ds = ep.calibrate.compute_Sv(ed)
    for name in ("Sv", "echo_range", "tau_effective", "frequency_nominal", "sound_speed", "sound_absorption",
                 "sa_correction", "gain_correction", "equivalent_beam_angle", "source_filenames", "water_level"):
        assert name in ds, name
    assert ds["Sv"].attrs["long_name"] == "Volume backscattering strength (Sv re 1 m-1)"
    assert ds["Sv"].attrs["units"] == "dB"
    assert ds["echo_range"].attrs == {"long_name": "Range distance", "units": "m"}
    assert ds.attrs["processing_function"] == "calibrate.compute_Sv"
    ts = ep.calibrate.compute_TS(ed)
    assert "TS" in ts and "tau_effective" not in ts


def test_argument_errors(ep):
    from echopype_b200 import synth

    ed80 = synth.make_ek80(C=1, P=2, R=16, mode="CW", encode="power")
    with pytest.raises(ValueError, match="waveform_mode and encode_mode must be specified"):
        ep.calibrate.compute_Sv(ed80)
    with pytest.raises(ValueError, match="must be recorded as complex samples"):
        ep.calibrate.compute_Sv(ed80, waveform_mode="BB", encode_mode="power")
    ed60 = synth.make_ek60(1, 2, 16)
    with pytest.raises(ValueError, match="assume_single_filter_time can only be used on complex EK80 data."):
        ep.calibrate.compute_Sv(ed60, assume_single_filter_time=True)
    with pytest.raises(ReferenceError):
        ep.calibrate.compute_Sv(synth.make_azfp(1, 2, 16))


# ---- EK80 with several filter_time entries (calibrate/api.py:95-197) --------------------------------------------------
def _multi_filter_ed(second_same=True, mux=False):
    """EK80 BB volume whose Vendor_specific carries two filter sets: at the first ping and at ping 7."""
    from echopype_b200 import synth
    from echopype_b200.dataset import Dataset

    ed = synth.make_ek80(C=2, P=12, R=256, B=4, mode="BB", encode="complex", nan_tail=0.2, seed=17)
    vend, beam = ed["Vendor_specific"], ed["Sonar/Beam_group1"]
    pt = np.asarray(beam["ping_time"].values)
    new = {}
    for name in vend:
        v = vend[name]
        a = np.asarray(v.values)
        if "filter_time" in v.dims:
            ax = list(v.dims).index("filter_time")
            b = a.copy()
            if not second_same and name.startswith("PC_coeffs"):
                b = b * 0.5 + 0.01
            a = np.concatenate([a, b], axis=ax)
        new[name] = (tuple(v.dims), a)
    coords = {k: np.asarray(c.values) for k, c in vend.coords.items() if k != "filter_time"}
    coords["filter_time"] = pt[[1 if mux else 0, 7]]
    ed["Vendor_specific"] = Dataset(new, coords=coords)
    if mux:  # channel-dependent first valid ping (multiplexed transceivers)
        tau = np.asarray(beam["transmit_duration_nominal"].values, dtype=np.float64).copy()
        tau[:, 0] = np.nan
        beam["transmit_duration_nominal"] = (("channel", "ping_time"), tau)
    return ed


def test_ek80_multiple_filter_times_equal_single_when_filters_repeat(ep):
    """The same filter set re-sent mid-file (what combining files produces): the piecewise calibration merged back
    together equals the one-shot calibration, channels in sorted order."""
    from echopype_b200 import synth

    ed1 = synth.make_ek80(C=2, P=12, R=256, B=4, mode="BB", encode="complex", nan_tail=0.2, seed=17)
    want = ep.calibrate.compute_Sv(ed1, waveform_mode="BB", encode_mode="complex")
    got = ep.calibrate.compute_Sv(_multi_filter_ed(), waveform_mode="BB", encode_mode="complex")
    order = np.argsort([str(c) for c in want["channel"].values])
    assert list(got["channel"].values) == [want["channel"].values[i] for i in order]
    np.testing.assert_array_equal(got["ping_time"].values, want["ping_time"].values)
    a, b = got["echo_range"].values, want["echo_range"].values[order]
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))

    def same_sv(a, b):
        # the FFT matched filter packs the short last segments of up to four pings into one transform, so a ping's
        # rounding noise (~1e-7 of its RMS output) depends on which pings share the transform: equal outside nulls only
        # (these pings are shorter than the replica, R = 256 < M = 277: every output is a partial window)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        d = np.abs(a - b)[~np.isnan(a)]
        # a wrong merge (pings or channels swapped) would show up as dB-scale differences
        assert np.median(d) <= 1e-3 and np.quantile(d, 0.9) <= 1e-2, (float(np.median(d)), float(np.quantile(d, 0.9)))

    same_sv(got["Sv"].values, want["Sv"].values[order])
    np.testing.assert_allclose(got["tau_effective"].values, want["tau_effective"].values[order], rtol=0)
    # assume_single_filter_time: one pass with the collapsed filter table, same numbers
    one = ep.calibrate.compute_Sv(_multi_filter_ed(), waveform_mode="BB", encode_mode="complex", assume_single_filter_time=True)
    same_sv(one["Sv"].values, want["Sv"].values[order])


def test_ek80_multiple_filter_times_conflict_and_missing(ep):
    # differing filter sets give per-channel tau_effective values that xr.merge(compat="no_conflicts") refuses
    with pytest.raises(ValueError, match="conflicting values for variable 'tau_effective'"):
        ep.calibrate.compute_Sv(_multi_filter_ed(second_same=False), waveform_mode="BB", encode_mode="complex")
    # assume_single_filter_time picks the filter set AT each channel's first valid ping (ping 1 here)
    ed = _multi_filter_ed(second_same=False, mux=True)
    one = ep.calibrate.compute_Sv(ed, waveform_mode="BB", encode_mode="complex", assume_single_filter_time=True)
    assert np.isnan(one["Sv"].values[:, 0]).all() and not np.isnan(one["Sv"].values[:, 1]).all()
    # ... and fails like Dataset.sel when no filter set was recorded at that ping
    bad = _multi_filter_ed(mux=True)
    bad["Vendor_specific"]._set_coord("filter_time", np.asarray(bad["Sonar/Beam_group1"]["ping_time"].values)[[0, 7]])
    with pytest.raises(KeyError):
        ep.calibrate.compute_Sv(bad, waveform_mode="BB", encode_mode="complex", assume_single_filter_time=True)
    with pytest.raises(ValueError, match="assume_single_filter_time can only be used on complex EK80 data."):
        ep.calibrate.compute_Sv(_multi_filter_ed(), waveform_mode="CW", encode_mode="power", assume_single_filter_time=True)
