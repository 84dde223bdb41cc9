"""Oracle: per-sample Sv / TS calibration (TEST INFRASTRUCTURE, see oracle/__init__.py).

float64 numpy restatement of
  echopype/calibrate/range.py:11-201          (echo_range, TVG range modification)
  echopype/calibrate/calibrate_ek.py:79-206   (EK60 / EK80 power samples)
  echopype/calibrate/calibrate_ek.py:456-659  (EK80 complex samples, CW and BB)
  echopype/calibrate/calibrate_azfp.py:49-111 (AZFP counts)
  echopype/calibrate/cal_params.py:261-324    (pulse-length table lookup)
  echopype/calibrate/env_params.py:24-71      (time1 -> ping_time harmonisation)
All arrays are plain numpy with the reference dimension order (channel, ping_time, range_sample
[, beam]).  PINNED: tests/test_reference_pinned.py checks every function here against outputs of the
reference's own calibration classes executed from /root/reference (tests/golden/make_golden_calibrate.py,
tests/golden/calibrate_vectors.npz) to 1e-9 dB.
"""

import numpy as np

from . import ek80_signal


def _cp1(x, C, P):
    """Broadcast a scalar / (C,) / (C,P) parameter to (C,P,1) float64."""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        a = np.full((C, P), float(a))
    elif a.ndim == 1:
        if a.shape[0] != C:
            raise ValueError("1-D parameter must have length = number of channels")
        a = np.repeat(a[:, None], P, axis=1)
    elif a.shape != (C, P):
        a = np.broadcast_to(a, (C, P)).copy()
    return a[:, :, None]


def harmonize_time1(values, time1_ns, ping_time_ns):
    """env_params.py:24-71 + utils/align.py:9-61 for one channel's (time1,) series.

    One time1 value (or one non-NaN value) -> scalar; identical axes -> as is; otherwise linear
    interpolation with linear extrapolation (scipy interp1d fill_value="extrapolate").
    """
    v = np.asarray(values, dtype=np.float64)
    t1 = np.asarray(time1_ns, dtype=np.int64)
    if v.size == 1:
        return float(v.reshape(-1)[0])
    ok = ~np.isnan(v)
    if ok.sum() == 1:
        return float(v[ok][0])
    if ping_time_ns is None:
        raise ValueError("ping_time needs to be provided for comparison or interpolating")
    tp = np.asarray(ping_time_ns, dtype=np.int64)
    v, t1 = v[ok], t1[ok]
    if t1.size == tp.size and np.array_equal(t1, tp):
        return v.copy()
    x = (tp - t1[0]).astype(np.float64)
    xp = (t1 - t1[0]).astype(np.float64)
    out = np.interp(x, xp, v)
    lo, hi = x < xp[0], x > xp[-1]
    if lo.any():
        out[lo] = v[0] + (x[lo] - xp[0]) * (v[1] - v[0]) / (xp[1] - xp[0])
    if hi.any():
        out[hi] = v[-1] + (x[hi] - xp[-1]) * (v[-1] - v[-2]) / (xp[-1] - xp[-2])
    return out


def vend_cal_params_power(tau_nominal, pulse_length, table):
    """cal_params.py:288-316.  argmin_k |tau[c,p] - pulse_length[c,k]| -> table[c,k]; NaN tau -> NaN.

    tau_nominal (C,P); pulse_length, table (C,K) with rows already matched to the same channels.
    """
    tau = np.asarray(tau_nominal, dtype=np.float64)
    pl = np.asarray(pulse_length, dtype=np.float64)
    tb = np.asarray(table, dtype=np.float64)
    isnull = np.isnan(tau)
    d = np.abs(tau[:, :, None] - pl[:, None, :])
    d = np.where(np.isnan(d), np.inf, d)
    idx = np.argmin(d, axis=2)  # first minimum, like xarray idxmin
    idx = np.where(isnull, 0, idx)
    out = np.take_along_axis(tb[:, None, :].repeat(tau.shape[1], axis=1), idx[:, :, None], axis=2)[
        :, :, 0
    ]
    return np.where(isnull, np.nan, out)


# --------------------------------------------------------------------------------------------
# range
# --------------------------------------------------------------------------------------------
def ek_echo_range(n_range_sample, sample_interval, sound_speed, backscatter_r):
    """range.py:138-148.  R = ((n * dt) * c) / 2 ; NaN where backscatter_r (beam 0) is NaN."""
    bs = np.asarray(backscatter_r)
    if bs.ndim == 4:
        bs = bs[..., 0]
    C, P, R = bs.shape
    n = np.arange(n_range_sample, dtype=np.int64)[None, None, :]
    rng = (n * _cp1(sample_interval, C, P)) * _cp1(sound_speed, C, P) / 2
    rng = np.where(np.isnan(bs), np.nan, rng)
    return rng


def ek_tvg_range(sonar_model, echo_range, sample_interval, sound_speed, tau_nominal, is_gpt=None):
    """range.py:160-201 followed by calibrate_ek.py:107 (R' <= 0 -> NaN)."""
    C, P, _ = echo_range.shape
    dt = _cp1(sample_interval, C, P)
    c = _cp1(sound_speed, C, P)
    ex60 = 2 * dt * c / 2
    if sonar_model in ("EK60", "ES70"):
        r = echo_range - ex60
    elif sonar_model in ("EK80", "ES80", "EA640"):
        r = echo_range - c * _cp1(tau_nominal, C, P) / 4
        if is_gpt is not None and np.any(is_gpt):
            g = np.asarray(is_gpt, dtype=bool)
            r[g] = r[g] - ex60[g]
    else:
        raise ValueError("The specified sonar_model is not supported!")
    with np.errstate(invalid="ignore"):
        return np.where(r > 0, r, np.nan)


def azfp_echo_range(n_range_sample, sound_speed, tau_nominal, N, f_dig, L, cal_type):
    """range.py:60-95.  N, f_dig, L per channel (C,); tau (C,P); sound_speed scalar/(P,)/(C,P)."""
    if cal_type is None:
        raise ValueError('cal_type must be "Sv" or "TS"')
    tau = np.asarray(tau_nominal, dtype=np.float64)
    C, P = tau.shape
    c = np.asarray(sound_speed, dtype=np.float64)
    if c.ndim == 1 and c.shape[0] == P and P != C:
        c = np.broadcast_to(c[None, :], (C, P))
    c = _cp1(c, C, P)
    N_ = np.asarray(N)[:, None, None]
    f_ = np.asarray(f_dig, dtype=np.float64)[:, None, None]
    L_ = np.asarray(L)[:, None, None]
    tau_ = tau[:, :, None]
    n = np.arange(n_range_sample, dtype=np.int64)[None, None, :]
    offset = 0 if cal_type == "Sv" else c * tau_ / 4
    rng = c * L_ / (2 * f_) + (c / 4) * (((2 * (n + 1) - 1) * N_ * 1 - 1) / f_ + tau_) - offset
    return np.broadcast_to(rng, (C, P, n_range_sample)).astype(np.float64)


# --------------------------------------------------------------------------------------------
# EK60 / EK80 power samples
# --------------------------------------------------------------------------------------------
def ek_power_cal(
    cal_type,
    sonar_model,
    backscatter_r,
    sample_interval,
    sound_speed,
    sound_absorption,
    tau_nominal,
    transmit_power,
    frequency_nominal,
    gain_correction,
    sa_correction,
    equivalent_beam_angle,
    tau_effective,
    is_gpt=None,
):
    """calibrate_ek.py:79-206.  Returns dict(out=(C,P,R) f64 Sv|TS, echo_range=(C,P,R) f64)."""
    bs = np.asarray(backscatter_r).astype(np.float64)  # numpy>=2: f32 array op f64 -> f64
    C, P, R = bs.shape
    c = _cp1(sound_speed, C, P)
    rng = ek_echo_range(R, sample_interval, sound_speed, backscatter_r)
    rp = ek_tvg_range(sonar_model, rng, sample_interval, sound_speed, tau_nominal, is_gpt)
    with np.errstate(invalid="ignore", divide="ignore"):
        spreading = 20 * np.log10(rp)
        absorb = 2 * _cp1(sound_absorption, C, P) * rp
        wavelength = c / np.asarray(frequency_nominal, dtype=np.float64)[:, None, None]
        G = _cp1(gain_correction, C, P)
        Pt = _cp1(transmit_power, C, P)
        if cal_type == "Sv":
            te = _cp1(tau_effective, C, P)
            CSv = (
                10 * np.log10(Pt)
                + 2 * G
                + _cp1(equivalent_beam_angle, C, P)
                + 10 * np.log10(wavelength**2 * te * c / (32 * np.pi**2))
            )
            out = bs + spreading + absorb - CSv - 2 * _cp1(sa_correction, C, P)
        elif cal_type == "TS":
            CSp = 10 * np.log10(Pt) + 2 * G + 10 * np.log10(wavelength**2 / (16 * np.pi**2))
            out = bs + spreading * 2 + absorb - CSp
        else:
            raise ValueError("cal_type must be Sv or TS")
    return {"out": out, "echo_range": rng}


# --------------------------------------------------------------------------------------------
# AZFP
# --------------------------------------------------------------------------------------------
def azfp_power_cal(
    cal_type,
    backscatter_r,
    sound_speed,
    sound_absorption,
    tau_nominal,
    N,
    f_dig,
    L,
    EL,
    DS,
    TVR,
    VTX0,
    equivalent_beam_angle,
    Sv_offset,
):
    """calibrate_azfp.py:49-111.  Per-channel (C,) cal params; counts (C,P,R)."""
    bs = np.asarray(backscatter_r).astype(np.float64)
    C, P, R = bs.shape
    rng = azfp_echo_range(R, sound_speed, tau_nominal, N, f_dig, L, cal_type)
    ch = lambda v: np.asarray(v, dtype=np.float64).reshape(C, 1, 1)  # noqa: E731
    c = np.asarray(sound_speed, dtype=np.float64)
    if c.ndim == 1 and c.shape[0] == P and P != C:
        c = np.broadcast_to(c[None, :], (C, P))
    c = _cp1(c, C, P)
    alpha = _cp1(sound_absorption, C, P)
    with np.errstate(invalid="ignore", divide="ignore"):
        spreading = 20 * np.log10(rng)
        absorb = 2 * alpha * rng
        SL = ch(TVR) + 20 * np.log10(ch(VTX0))
        a = ch(DS)
        ELv = ch(EL) - 2.5 / a + bs / (26214 * a)
        if cal_type == "Sv":
            out = (
                ELv
                - SL
                + spreading
                + absorb
                - 10
                * np.log10(
                    0.5 * c * np.asarray(tau_nominal, dtype=np.float64)[:, :, None] * ch(equivalent_beam_angle)
                )
                + ch(Sv_offset)
            )
        elif cal_type == "TS":
            out = ELv - SL + 2 * spreading + absorb
        else:
            raise ValueError("cal_type not recognized!")
    return {"out": out, "echo_range": rng}


# --------------------------------------------------------------------------------------------
# EK80 complex samples (CW and BB)
# --------------------------------------------------------------------------------------------
def b_theta_phi_m(angle_offset_along, angle_offset_athwart, beamwidth_along, beamwidth_athwart):
    """calibrate_ek.py:507-530.  NaN -> 0."""
    fa = (np.abs(-np.asarray(angle_offset_along, dtype=np.float64)) / (np.asarray(beamwidth_along) / 2)) ** 2
    ft = (np.abs(-np.asarray(angle_offset_athwart, dtype=np.float64)) / (np.asarray(beamwidth_athwart) / 2)) ** 2
    B = 0.5 * 6.0206 * (fa + ft - 0.18 * fa * ft)
    return np.where(np.isnan(B), 0.0, B)


def prx_from_complex(sig, n_beam, z_et, z_er):
    """calibrate_ek.py:483-490.  sig (C,P,R,B) complex; beam mean skips NaN (xarray default)."""
    cnt = np.sum(~np.isnan(sig), axis=-1)
    s = np.where(np.isnan(sig), 0, sig).sum(axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(cnt > 0, s / np.where(cnt > 0, cnt, 1), np.nan + 0j)
        C, P = sig.shape[:2]
        zer = _cp1(z_er, C, P)
        zet = _cp1(z_et, C, P)
        return n_beam * np.abs(mean) ** 2 / (2 * np.sqrt(2)) ** 2 * (np.abs(zer + zet) / zer) ** 2 / zet


def ek80_complex_cal(
    cal_type,
    waveform_mode,
    backscatter_r,
    backscatter_i,
    sample_interval,
    sound_speed,
    sound_absorption,
    tau_nominal,
    transmit_power,
    freq_center,
    gain_correction,
    sa_correction,
    equivalent_beam_angle,
    tau_effective,
    z_et,
    z_er,
    is_gpt=None,
    chirp=None,
):
    """calibrate_ek.py:532-659.  backscatter_* (C,P,R,B) float; chirp = list of per-channel replicas (BB).

    ``gain_correction`` must already include the BB ``- B_theta_phi_m`` term (calibrate_ek.py:561-562)
    and ``equivalent_beam_angle`` the BB ``+20log10(f_nom/f_c)`` term (cal_params.py:499-503).
    """
    br = np.asarray(backscatter_r).astype(np.float64)
    bi = np.asarray(backscatter_i).astype(np.float64)
    C, P, R, B = br.shape
    c = _cp1(sound_speed, C, P)
    rng = ek_echo_range(R, sample_interval, sound_speed, br)
    rp = ek_tvg_range("EK80", rng, sample_interval, sound_speed, tau_nominal, is_gpt)
    sig = br + 1j * bi
    if waveform_mode == "BB":
        pc = ek80_signal.compress_pulse(sig, chirp)
        norm = np.array([np.linalg.norm(tx) ** 2 for tx in chirp])
        sig = pc / norm[:, None, None, None]
    prx = prx_from_complex(sig, B, z_et, z_er)
    with np.errstate(invalid="ignore", divide="ignore"):
        prx = np.where(prx > 0, prx, np.nan)
        spreading = 20 * np.log10(rp)
        absorb = 2 * _cp1(sound_absorption, C, P) * rp
        wavelength = c / _cp1(freq_center, C, P)
        Pt = _cp1(transmit_power, C, P)
        G = _cp1(gain_correction, C, P)
        if cal_type == "Sv":
            out = (
                10 * np.log10(prx)
                + spreading
                + absorb
                - 10 * np.log10(wavelength**2 * Pt * c / (32 * np.pi**2))
                - 2 * G
                - 10 * np.log10(_cp1(tau_effective, C, P))
                - _cp1(equivalent_beam_angle, C, P)
            )
            if waveform_mode == "CW":
                out = out - 2 * _cp1(sa_correction, C, P)
        elif cal_type == "TS":
            out = (
                10 * np.log10(prx)
                + 2 * spreading
                + absorb
                - 10 * np.log10(wavelength**2 * Pt / (16 * np.pi**2))
                - 2 * G
            )
        else:
            raise ValueError("cal_type must be Sv or TS")
    return {"out": out, "echo_range": rng, "prx": prx}
