"""Test glue: run the CPU oracle (oracle/) on the synthetic EchoData objects of echopype_b200.synth.

The parameter assembly here is an independent restatement of the reference's host logic
(calibrate/env_params.py:160-353, cal_params.py:261-522, calibrate_ek.py:112-151,:507-530,:561-562) in
terms of oracle functions only; nothing from echopype_b200's calibrate package is used, so a bug in the
product's host-side assembly shows up as a parity failure.
"""

import numpy as np

from oracle import calibrate as ocal
from oracle import ek80_signal, seawater

PL_DIMS = ("channel", "pulse_length_bin")


def _v(ds, name):
    return np.asarray(ds[name].values)


def _ns(t):
    return np.asarray(t).astype("datetime64[ns]").astype(np.int64)


def _harmonize_cp(values, time1, ping_time):
    """(C,time1) -> (C,) or (C,P) via the oracle's per-channel harmonisation (env_params.py:24-71)."""
    vals = np.asarray(values, dtype=np.float64)
    out = [ocal.harmonize_time1(vals[c], _ns(time1), _ns(ping_time)) for c in range(vals.shape[0])]
    if all(np.ndim(o) == 0 for o in out):
        return np.asarray(out, dtype=np.float64)
    P = len(ping_time)
    return np.stack([np.full(P, o) if np.ndim(o) == 0 else o for o in out])


def ek60(ed, cal_type):
    beam, env, vend = ed["Sonar/Beam_group1"], ed["Environment"], ed["Vendor_specific"]
    pt = _v(beam, "ping_time")
    tau = _v(beam, "transmit_duration_nominal")
    c = _harmonize_cp(_v(env, "sound_speed_indicative"), _v(env, "time1"), pt)
    alpha = _harmonize_cp(_v(env, "absorption_indicative"), _v(env, "time1"), pt)
    gain = ocal.vend_cal_params_power(tau, _v(vend, "pulse_length"), _v(vend, "gain_correction"))
    sa = ocal.vend_cal_params_power(tau, _v(vend, "pulse_length"), _v(vend, "sa_correction"))
    C, P = tau.shape
    res = ocal.ek_power_cal(
        cal_type, "EK60", _v(beam, "backscatter_r"), _v(beam, "sample_interval"), c, alpha, tau, _v(beam, "transmit_power"),
        _v(beam, "frequency_nominal"), gain, sa, _v(beam, "equivalent_beam_angle"),
        np.repeat(tau[:, :1], P, axis=1),  # tau_effective = nominal duration of ping 0 (calibrate_ek.py:134-151)
        is_gpt=np.ones(C, bool),
    )
    res["sound_absorption"] = alpha
    res["params"] = {"sound_speed": c, "sound_absorption": alpha, "gain_correction": gain, "sa_correction": sa,
                     "tau_effective": tau[:, 0]}
    return res


def azfp(ed, cal_type, salinity, pressure):
    beam, env, vend = ed["Sonar/Beam_group1"], ed["Environment"], ed["Vendor_specific"]
    T = float(np.asarray(_v(env, "temperature")).reshape(-1)[0])
    c = seawater.sound_speed(T, salinity, pressure, "AZFP")
    f = _v(beam, "frequency_nominal").astype(np.float64)
    alpha = seawater.absorption(f, T, salinity, pressure, formula_source="AZFP")
    tau = _v(beam, "transmit_duration_nominal")
    res = ocal.azfp_power_cal(
        cal_type, _v(beam, "backscatter_r"), c, alpha, tau, _v(vend, "number_of_samples_per_average_bin"),
        _v(vend, "digitization_rate"), _v(vend, "lock_out_index"), _v(vend, "EL"), _v(vend, "DS"), _v(vend, "TVR"),
        _v(vend, "VTX0"), _v(beam, "equivalent_beam_angle"), _v(vend, "Sv_offset"),
    )
    res["sound_absorption"] = alpha
    res["params"] = {"sound_speed": c, "sound_absorption": alpha}
    return res


def _ek80_replicas(beam, vend, waveform_mode, drop_last_hanning_zero=False):
    C = len(_v(beam, "channel"))
    fs = _v(vend, "receiver_sampling_frequency").astype(np.float64)
    txs, fs_deci = [], []
    for c in range(C):
        def coeff(name):
            v = _v(vend, f"{name}_coeffs_real")[c, 0] + 1j * _v(vend, f"{name}_coeffs_imag")[c, 0]
            return v[~np.isnan(v)]

        filt = {
            "wbt_fil": coeff("WBT"), "wbt_decifac": int(_v(vend, "WBT_deci_fac")[c, 0]),
            "pc_fil": coeff("PC"), "pc_decifac": int(_v(vend, "PC_deci_fac")[c, 0]),
        }
        tx, t = ek80_signal.transmit_signal(
            waveform_mode, fs[c], float(_v(beam, "transmit_duration_nominal")[c, 0]), float(_v(beam, "slope")[c, 0]),
            float(_v(beam, "transmit_frequency_start")[c, 0]), float(_v(beam, "transmit_frequency_stop")[c, 0]),
            float(_v(beam, "frequency_nominal")[c]), filt, drop_last_hanning_zero,
        )
        txs.append(tx)
        fs_deci.append(1.0 / (t[1] - t[0]))
    return txs, np.asarray(fs_deci)


def ek80(ed, cal_type, waveform_mode, encode_mode, drop_last_hanning_zero=False):
    beam, env, vend = ed["Sonar/Beam_group1"], ed["Environment"], ed["Vendor_specific"]
    C = len(_v(beam, "channel"))
    tau = _v(beam, "transmit_duration_nominal")
    P = tau.shape[1]
    f_nom = _v(beam, "frequency_nominal").astype(np.float64)
    if waveform_mode == "BB":
        f_c = (_v(beam, "transmit_frequency_start") + _v(beam, "transmit_frequency_stop")) / 2  # (C,P)
    else:
        f_c = np.repeat(f_nom[:, None], P, 1)
    T, S = float(_v(env, "temperature")[0]), float(_v(env, "salinity")[0])
    D, pH = float(_v(env, "depth")[0]), float(_v(env, "acidity")[0])
    c = float(_v(env, "sound_speed_indicative")[0])
    alpha = seawater.absorption(f_c, T, S, D, pH, c, "FG")  # env_params.py:322-340: always FG at f_center
    is_gpt = _v(vend, "transceiver_type").astype(str) == "GPT"
    gain = ocal.vend_cal_params_power(tau, _v(vend, "pulse_length"), _v(vend, "gain_correction"))
    sa = ocal.vend_cal_params_power(tau, _v(vend, "pulse_length"), _v(vend, "sa_correction"))
    psi = _v(beam, "equivalent_beam_angle").astype(np.float64)
    if encode_mode == "power":
        txs, fs_deci = _ek80_replicas(beam, vend, "CW")
        te = np.array([ek80_signal.tau_effective(txs[i], fs_deci[i], "CW") for i in range(C)])
        te = np.where(is_gpt, tau[:, 0], te)
        res = ocal.ek_power_cal(
            cal_type, "EK80", _v(beam, "backscatter_r"), _v(beam, "sample_interval"), c, alpha, tau, _v(beam, "transmit_power"),
            f_nom, gain, sa, psi, np.repeat(te[:, None], P, 1), is_gpt=is_gpt,
        )
        res["sound_absorption"] = alpha
        res["tau_effective"] = te
        res["params"] = {"sound_speed": c, "sound_absorption": alpha, "gain_correction": gain, "sa_correction": sa,
                         "tau_effective": te}
        return res
    txs, fs_deci = _ek80_replicas(beam, vend, waveform_mode, drop_last_hanning_zero)
    te = np.array([ek80_signal.tau_effective(txs[i], fs_deci[i], waveform_mode) for i in range(C)])
    te = np.where(is_gpt, tau[:, 0], te)
    if waveform_mode == "BB":
        bw_al = _v(beam, "beamwidth_twoway_alongship")[:, None] * (f_nom[:, None] / f_c)  # cal_params.py:477-483
        bw_at = _v(beam, "beamwidth_twoway_athwartship")[:, None] * (f_nom[:, None] / f_c)
        Bm = ocal.b_theta_phi_m(_v(beam, "angle_offset_alongship")[:, None], _v(beam, "angle_offset_athwartship")[:, None], bw_al, bw_at)
        gain = gain - Bm
        psi = psi[:, None] + 20 * np.log10(f_nom[:, None] / f_c)
    z_er = _v(vend, "impedance_transceiver").astype(np.float64)
    res = ocal.ek80_complex_cal(
        cal_type, waveform_mode, _v(beam, "backscatter_r"), _v(beam, "backscatter_i"), _v(beam, "sample_interval"), c, alpha,
        tau, _v(beam, "transmit_power"), f_c, gain, sa, psi, np.repeat(te[:, None], P, 1), 75.0, z_er, is_gpt=is_gpt,
        chirp=txs if waveform_mode == "BB" else None,
    )
    res["sound_absorption"] = alpha
    res["tau_effective"] = te
    res["chirp"] = txs
    res["params"] = {"sound_speed": c, "sound_absorption": alpha, "gain_correction": gain, "sa_correction": sa,
                     "tau_effective": te, "equivalent_beam_angle": psi}
    return res


def compare_db(got, want, atol, what="Sv"):
    """NaN masks must be identical; finite values within atol (dB).  Returns max |diff|."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), f"{what}: NaN masks differ at {int((gn != wn).sum())} of {got.size} samples"
    gi, wi = np.isinf(got), np.isinf(want)
    assert np.array_equal(gi, wi) and np.array_equal(np.sign(got[gi]), np.sign(want[wi])), f"{what}: inf pattern differs"
    ok = ~(gn | gi)
    d = np.abs(got[ok] - want[ok])
    m = float(d.max()) if d.size else 0.0
    assert m <= atol, f"{what}: max |diff| = {m:.3e} > {atol:g}"
    return m


BB_NULL_DB = 26.0  # a "null" of the matched-filter output: received power more than 26 dB below the ping's mean power
BB_NULL_REL = 5e-6  # amplitude error allowed inside nulls, relative to the ping's RMS matched-filter amplitude (-106 dB)


def compare_bb_db(got, want, prx, atol, what="Sv"):
    """Pulse-compressed Sv / TS.  NaN masks identical.  Outside nulls (received power prx within BB_NULL_DB of the
    ping's mean) |diff| <= atol dB.  Inside nulls a float32 matched filter cannot hold a dB tolerance (the output is a
    cancelling sum of ~300 products; the reference computes it in complex128 and rounds to complex64): there the
    amplitude error must stay below BB_NULL_REL of the ping's RMS amplitude.  Returns (max dB diff outside nulls,
    fraction of samples inside nulls, max RMS-relative amplitude error)."""
    got, want, prx = np.asarray(got, np.float64), np.asarray(want, np.float64), np.asarray(prx, np.float64)
    assert got.shape == want.shape == prx.shape
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), f"{what}: NaN masks differ at {int((gn != wn).sum())} samples"
    ok = ~wn & np.isfinite(want) & (prx > 0)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        mean_p = np.nanmean(np.where(ok, prx, np.nan), axis=-1, keepdims=True)
    rel_p = np.where(ok, prx / mean_p, np.nan)
    strong = ok & (rel_p >= 10 ** (-BB_NULL_DB / 10))
    d = np.abs(got - want)
    m = float(d[strong].max()) if strong.any() else 0.0
    assert m <= atol, f"{what}: max |diff| outside nulls = {m:.3e} dB > {atol:g}"
    amp_err = np.abs(10 ** ((got - want) / 20) - 1) * np.sqrt(rel_p)
    null = ok & ~strong
    worst = float(amp_err[null].max()) if null.any() else 0.0
    assert worst <= BB_NULL_REL, f"{what}: amplitude error {worst:.2e} of the ping RMS > {BB_NULL_REL:g}"
    frac_null = float((ok & ~strong).sum() / max(1, ok.sum()))
    # (short pings, R ~ M, end in a long ramp of partial windows far below the mean; long pings have < 1 % nulls:
    # tests/test_gpu_fullsize.py asserts that for cfg3)
    return m, frac_null, worst
