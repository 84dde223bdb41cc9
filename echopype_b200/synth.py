"""Synthetic EchoData builders with the value distributions of SURVEY.md 8(d).

``device=False`` builds host (numpy) arrays with numpy's RNG - used by tests and by the CPU baseline;
``device=True`` fills the big backscatter arrays on the GPU with the library's Philox generator
(epb_synth_fill) so that benchmark volumes never cross PCIe.  Parameter groups are always tiny host
arrays, exactly the variables the reference reads (SURVEY.md 8b "Input contract").
"""

import numpy as np

from .dataset import DataArray, Dataset, EchoData

_T0 = np.datetime64("2018-07-01T00:00:00", "ns")
_DB_PER_COUNT = 10.0 * np.log10(2.0) / 256.0  # INDEX2POWER, convert/parse_base.py:24


def ping_times(P, interval_s=1.0, jitter_ms=0, rng=None, offset=0):
    t = _T0 + ((np.arange(P, dtype=np.int64) + offset) * int(round(interval_s * 1e9))).astype("timedelta64[ns]")
    if jitter_ms:
        t = np.sort(t + (rng.integers(jitter_ms, size=P) * 1_000_000).astype("timedelta64[ns]"))
    return t


def _chan_names(prefix, freqs):
    return np.array([f"{prefix} {int(f / 1000)} kHz" for f in freqs], dtype=object)


def _power_host(rng, C, P, R, nan_tail, raw_counts=False):
    qi = rng.integers(-24000, -2000, size=(C, P, R), endpoint=True).astype(np.int16)
    x = (qi.astype(np.float64) * _DB_PER_COUNT).astype(np.float32)  # parse_base.py:302 computes this product in float64
    if nan_tail:
        short = rng.random((C, P)) < nan_tail
        cut = rng.integers(R // 4, R, size=(C, P))
        n = np.arange(R)[None, None, :]
        pad = short[:, :, None] & (n >= cut[:, :, None])
        x[pad] = np.nan
        qi[pad] = -32768  # the padding marker of the raw-count ingest format
    return qi if raw_counts else x


def make_ek60(C=4, P=1000, R=1000, seed=1001, device=False, nan_tail=0.005, ping_interval_s=1.0, time_varying=False, ping_offset=0,
              backscatter=None, raw_counts=False):
    """EK60 CW power volume (cfg1 / cfg2).  ``time_varying`` makes sample_interval / pulse length / env
    parameters change along ping_time to exercise the per-row paths."""
    rng = np.random.default_rng(seed)
    freqs = np.array([18e3, 38e3, 120e3, 200e3, 70e3, 333e3][:C] if C <= 6 else np.linspace(18e3, 333e3, C))
    chan = _chan_names("GPT", freqs)
    pt = ping_times(P, ping_interval_s, offset=ping_offset)
    cyc = lambda a: np.resize(np.asarray(a, dtype=np.float64), C)  # noqa: E731
    dt = np.full((C, P), 2.56e-4)
    tau = np.full((C, P), 1.024e-3)
    if time_varying:
        dt[:, P // 2 :] = 1.28e-4
        tau[:, P // 3 :] = 0.512e-3
        tau[0, 1] = np.nan  # a dropped ping (multiplexed systems): NaN parameters -> NaN row
    if backscatter is not None:  # caller-supplied (C, P, R) float32 volume (host array or device tensor)
        x = backscatter
        assert tuple(x.shape) == (C, P, R)
    elif device:
        from . import kernels

        if raw_counts:  # the same volume as int16 raw power counts (ingest format)
            x = kernels.synth_fill_i16((C, P, R), seed=seed, nan_tail=nan_tail, ping_offset=ping_offset)
        else:
            x = kernels.synth_fill((C, P, R), kind=0, seed=seed, nan_tail=nan_tail, ping_offset=ping_offset)
    else:
        x = _power_host(rng, C, P, R, nan_tail, raw_counts)
    beam = Dataset(
        {
            "backscatter_r": (("channel", "ping_time", "range_sample"), x),
            "sample_interval": (("channel", "ping_time"), dt),
            "transmit_duration_nominal": (("channel", "ping_time"), tau),
            "transmit_power": (("channel", "ping_time"), np.repeat(cyc([2000, 2000, 250, 120, 750, 60])[:, None], P, 1)),
            "frequency_nominal": (("channel",), freqs),
            "equivalent_beam_angle": (("channel",), cyc([-17.0, -20.6, -20.7, -20.5, -20.7, -20.1])),
            # split-beam angle parameters: not used by the power calibration, but the reference's
            # get_cal_params_EK pulls them from the beam group for every EK60 file (cal_params.py:455-460)
            "angle_offset_alongship": (("channel",), cyc([0.05, -0.08, 0.0, 0.02])),
            "angle_offset_athwartship": (("channel",), cyc([-0.03, 0.06, 0.01, 0.0])),
            "angle_sensitivity_alongship": (("channel",), cyc([15.5, 18.0, 23.0, 23.0])),
            "angle_sensitivity_athwartship": (("channel",), cyc([15.5, 18.0, 23.0, 23.0])),
            "beamwidth_twoway_alongship": (("channel",), cyc([10.9, 7.0, 6.8, 6.5])),
            "beamwidth_twoway_athwartship": (("channel",), cyc([10.8, 7.1, 6.7, 6.6])),
        },
        coords={"channel": chan, "ping_time": pt, "range_sample": np.arange(R)},
    )
    t1 = pt[[0, -1]] if (time_varying and P > 1) else pt[:1]
    ss = np.repeat(np.array([[1500.0]]), C, 0) if len(t1) == 1 else np.repeat(np.array([[1495.0, 1505.0]]), C, 0)
    ab = cyc([0.0027, 0.0098, 0.0375, 0.0527, 0.0187, 0.0809])[:, None] * (np.ones((1, len(t1))) if len(t1) == 1 else np.array([[1.0, 1.05]]))
    env = Dataset(
        {"sound_speed_indicative": (("channel", "time1"), ss), "absorption_indicative": (("channel", "time1"), ab)},
        coords={"channel": chan, "time1": t1},
    )
    pl = np.array([0.256e-3, 0.512e-3, 1.024e-3, 2.048e-3, 4.096e-3])
    vend = Dataset(
        {
            "pulse_length": (("channel", "pulse_length_bin"), np.repeat(pl[None, :], C, 0)),
            "gain_correction": (("channel", "pulse_length_bin"), cyc([22.9, 26.0, 27.0, 26.5, 26.8, 25.5])[:, None] + np.array([-0.6, -0.3, 0.0, 0.1, 0.15])[None, :]),
            "sa_correction": (("channel", "pulse_length_bin"), cyc([-0.7, -0.5, -0.3, -0.3, -0.4, -0.2])[:, None] + np.array([0.2, 0.1, 0.0, -0.05, -0.1])[None, :]),
        },
        coords={"channel": chan, "pulse_length_bin": np.arange(5)},
    )
    plat = Dataset({"water_level": ((), np.asarray(0.0))})
    return EchoData("EK60", {"Sonar/Beam_group1": beam, "Environment": env, "Vendor_specific": vend, "Platform": plat}, source_file="synthetic_ek60.raw")


def make_azfp(C=4, P=1000, R=1000, seed=4001, device=False, ping_interval_s=1.0, ping_offset=0, backscatter=None):
    """AZFP counts volume (cfg4)."""
    rng = np.random.default_rng(seed)
    freqs = np.resize(np.array([38e3, 125e3, 200e3, 455e3]), C)
    chan = np.array([f"55030-{int(f / 1000)}-{i + 1}" for i, f in enumerate(freqs)], dtype=object)
    pt = ping_times(P, ping_interval_s, offset=ping_offset)
    if backscatter is not None:
        x = backscatter
        assert tuple(x.shape) == (C, P, R)
    elif device:
        from . import kernels

        x = kernels.synth_fill((C, P, R), kind=1, seed=seed, nan_tail=0.0, ping_offset=ping_offset)
    else:
        x = rng.integers(0, 65536, size=(C, P, R)).astype(np.float32)
    cyc = lambda a: np.resize(np.asarray(a, dtype=np.float64), C)  # noqa: E731
    beam = Dataset(
        {
            "backscatter_r": (("channel", "ping_time", "range_sample"), x),
            "transmit_duration_nominal": (("channel", "ping_time"), np.repeat(cyc([5e-4, 3e-4, 3e-4, 1.5e-4])[:, None], P, 1)),
            "frequency_nominal": (("channel",), freqs),
            "equivalent_beam_angle": (("channel",), cyc([0.0218, 0.0084, 0.0086, 0.0085])),  # linear (steradian)
        },
        coords={"channel": chan, "ping_time": pt, "range_sample": np.arange(R)},
    )
    vend = Dataset(
        {
            "number_of_samples_per_average_bin": (("channel",), cyc([1, 1, 1, 1])),
            "digitization_rate": (("channel",), cyc([64000.0, 64000.0, 64000.0, 64000.0])),
            "lock_out_index": (("channel",), cyc([0, 10, 10, 20])),
            "EL": (("channel",), cyc([142.8, 141.1, 140.6, 139.5])),
            "DS": (("channel",), cyc([0.02333, 0.02359, 0.02301, 0.02256])),
            "TVR": (("channel",), cyc([169.9, 171.9, 174.0, 173.4])),
            "VTX0": (("channel",), cyc([108.5, 107.3, 105.6, 106.2])),
            "Sv_offset": (("channel",), cyc([1.1, 1.3, 1.3, 1.6])),
        },
        coords={"channel": chan},
    )
    env = Dataset({"temperature": (("time1",), np.array([10.0]))}, coords={"time1": pt[:1]})
    return EchoData("AZFP", {"Sonar/Beam_group1": beam, "Environment": env, "Vendor_specific": vend, "Platform": Dataset()}, source_file="synthetic.01A")


def _ek80_filters(C, rng):
    """Synthetic WBT (64 taps, /6) and PC (32 taps, /2) low-pass FIRs -> fs_dec = 125 kHz."""
    def lp(n, fc):
        k = np.arange(n) - (n - 1) / 2
        return np.sinc(2 * fc * k) * np.hanning(n) * 2 * fc

    wbt = np.repeat(lp(64, 1 / 14)[None, None, :], C, 0).astype(np.float64)
    pc = np.repeat(lp(32, 1 / 5)[None, None, :], C, 0).astype(np.float64)
    return wbt, pc


def make_ek80(C=2, P=50, R=512, B=4, seed=3001, mode="BB", encode="complex", device=False, gpt_channel=None, nan_tail=0.02, ping_offset=0,
              backscatter=None):
    """EK80 volume: mode "BB" (complex, pulse compression), "CW" with encode "complex" or "power" (cfg3 / cfg5)."""
    rng = np.random.default_rng(seed)
    freqs = np.resize(np.array([18e3, 38e3, 70e3, 120e3, 200e3, 333e3]), C)
    chan = _chan_names("WBT", freqs)
    pt = ping_times(P, 1.0, offset=ping_offset)
    tau = 2.048e-3 if mode == "BB" else 1.024e-3
    ttype = np.array(["WBT"] * C, dtype=object)
    if gpt_channel is not None:
        ttype[gpt_channel] = "GPT"
    coords = {"channel": chan, "ping_time": pt, "range_sample": np.arange(R)}
    dv = {
        "sample_interval": (("channel", "ping_time"), np.full((C, P), 8e-6 if encode == "complex" else 2.56e-4)),
        "transmit_duration_nominal": (("channel", "ping_time"), np.full((C, P), tau)),
        "transmit_power": (("channel", "ping_time"), np.repeat(np.resize([2000.0, 1000.0, 750.0, 250.0, 105.0, 40.0], C)[:, None], P, 1)),
        "frequency_nominal": (("channel",), freqs),
        "equivalent_beam_angle": (("channel",), np.resize([-17.0, -20.7, -20.7, -20.7, -20.7, -20.1], C)),
        "slope": (("channel", "ping_time"), np.full((C, P), 0.05)),
        "transmit_type": (("channel", "ping_time"), np.full((C, P), "LFM" if mode == "BB" else "CW", dtype=object)),
        "transmit_frequency_start": (("channel", "ping_time"), np.repeat((freqs * (0.88 if mode == "BB" else 1.0))[:, None], P, 1)),
        "transmit_frequency_stop": (("channel", "ping_time"), np.repeat((freqs * (1.12 if mode == "BB" else 1.0))[:, None], P, 1)),
        "angle_offset_alongship": (("channel",), np.resize([0.05, -0.08, 0.0, 0.02], C)),
        "angle_offset_athwartship": (("channel",), np.resize([-0.03, 0.06, 0.01, 0.0], C)),
        "angle_sensitivity_alongship": (("channel",), np.resize([15.5, 18.0, 23.0, 23.0], C)),
        "angle_sensitivity_athwartship": (("channel",), np.resize([15.5, 18.0, 23.0, 23.0], C)),
        "beamwidth_twoway_alongship": (("channel",), np.resize([10.9, 7.0, 6.8, 6.5], C)),
        "beamwidth_twoway_athwartship": (("channel",), np.resize([10.8, 7.1, 6.7, 6.6], C)),
    }
    if encode == "complex":
        coords["beam"] = np.arange(1, B + 1).astype(str)
        if backscatter is not None:  # caller-supplied (re, im) pair of (C, P, R, B) float32 volumes
            re, im = backscatter
            assert tuple(re.shape) == tuple(im.shape) == (C, P, R, B)
        elif device:
            from . import kernels

            re = kernels.synth_fill((C, P, R), kind=2, seed=seed, inner=B, nan_tail=nan_tail, scale=1e-3, ping_offset=ping_offset)
            im = kernels.synth_fill((C, P, R), kind=2, seed=seed + 7, inner=B, nan_tail=0.0, scale=1e-3, ping_offset=ping_offset)
        else:
            re = (rng.standard_normal((C, P, R, B)) * 1e-3).astype(np.float32)
            im = (rng.standard_normal((C, P, R, B)) * 1e-3).astype(np.float32)
            short = rng.random((C, P)) < nan_tail
            cut = rng.integers(R // 4, R, size=(C, P))
            m = short[:, :, None] & (np.arange(R)[None, None, :] >= cut[:, :, None])
            re[m] = np.nan
            im[m] = np.nan
        dv["backscatter_r"] = (("channel", "ping_time", "range_sample", "beam"), re)
        dv["backscatter_i"] = (("channel", "ping_time", "range_sample", "beam"), im)
        descr = "complex_FM" if mode == "BB" else "complex_CW"
    else:
        if backscatter is not None:
            x = backscatter
            assert tuple(x.shape) == (C, P, R)
        elif device:
            from . import kernels

            x = kernels.synth_fill((C, P, R), kind=0, seed=seed, nan_tail=nan_tail, ping_offset=ping_offset)
        else:
            x = _power_host(rng, C, P, R, nan_tail)
        dv["backscatter_r"] = (("channel", "ping_time", "range_sample"), x)
        descr = "power"
    beam = Dataset(dv, coords=coords)
    wbt, pc = _ek80_filters(C, rng)
    pl = np.array([0.256e-3, 0.512e-3, 1.024e-3, 2.048e-3, 4.096e-3])
    vend = Dataset(
        {
            "WBT_coeffs_real": (("channel", "filter_time", "WBT_filter_n"), wbt),
            "WBT_coeffs_imag": (("channel", "filter_time", "WBT_filter_n"), wbt * 0.1),
            "WBT_deci_fac": (("channel", "filter_time"), np.full((C, 1), 6)),
            "PC_coeffs_real": (("channel", "filter_time", "PC_filter_n"), pc),
            "PC_coeffs_imag": (("channel", "filter_time", "PC_filter_n"), pc * -0.05),
            "PC_deci_fac": (("channel", "filter_time"), np.full((C, 1), 2)),
            "transceiver_type": (("channel",), ttype),
            "impedance_transceiver": (("channel",), np.full(C, 5400.0)),
            "receiver_sampling_frequency": (("channel",), np.full(C, 1.5e6)),
            "pulse_length": (("channel", "pulse_length_bin"), np.repeat(pl[None, :], C, 0)),
            "gain_correction": (("channel", "pulse_length_bin"), np.resize([22.9, 26.0, 27.0, 26.5, 26.8, 25.5], C)[:, None] + np.array([-0.6, -0.3, 0.0, 0.1, 0.15])[None, :]),
            "sa_correction": (("channel", "pulse_length_bin"), np.resize([-0.7, -0.5, -0.3, -0.3, -0.4, -0.2], C)[:, None] + np.array([0.2, 0.1, 0.0, -0.05, -0.1])[None, :]),
        },
        coords={"channel": chan, "filter_time": pt[:1], "pulse_length_bin": np.arange(5)},
    )
    env = Dataset(
        {
            "temperature": (("time1",), np.array([8.0])), "salinity": (("time1",), np.array([33.0])),
            "depth": (("time1",), np.array([50.0])), "acidity": (("time1",), np.array([7.9])),
            "sound_speed_indicative": (("time1",), np.array([1481.0])),
        },
        coords={"time1": pt[:1]},
    )
    sonar = Dataset({"waveform_encode_descr": (("beam_group",), np.array([descr], dtype=object))}, coords={"beam_group": np.array(["Beam_group1"], dtype=object)})
    return EchoData(
        "EK80",
        {"Sonar": sonar, "Sonar/Beam_group1": beam, "Environment": env, "Vendor_specific": vend, "Platform": Dataset()},
        source_file="synthetic_ek80.raw",
    )
