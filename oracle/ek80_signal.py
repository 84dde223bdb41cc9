"""Oracle: EK80 transmit replica, effective pulse length, pulse compression.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows echopype/calibrate/ek80_complex.py:
  tapered_chirp :12-52, filter_decimate_chirp :55-80, get_tau_effective :162-208,
  get_transmit_signal :211-282, _convolve_per_channel :285-313, compress_pulse :316-369,
  get_norm_fac :372-391.
Third-party arithmetic: scipy.signal.convolve (scipy 1.18.1 installed here) is used directly, as
the reference does.
"""

import numpy as np
from scipy import signal


def tapered_chirp(fs, tau, slope, f0, f1, drop_last_hanning_zero=False):
    """ek80_complex.py:12-52.  Returns (y, t) with y normalised by max(y)."""
    nsamples = int(np.floor(tau * np.float32(fs)))
    t = np.linspace(0, nsamples - 1, num=nsamples) * 1 / fs
    a = np.pi * (f1 - f0) / tau
    b = 2 * np.pi * f0
    y = np.cos(a * t * t + b * t)
    L = int(np.round(tau * fs * slope * 2.0))
    w = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(0, L, 1) / (L - 1)))
    half = int(len(w) / 2)
    w1 = w[:half]
    w2 = w[half:-1] if drop_last_hanning_zero else w[half:]
    y[: len(w1)] = y[: len(w1)] * w1
    y[len(y) - len(w2) :] = y[len(y) - len(w2) :] * w2
    return y / np.max(y), t


def filter_decimate_chirp(y, fs, wbt_fil, wbt_decifac, pc_fil, pc_decifac):
    """ek80_complex.py:55-80.  Two full convolutions, each followed by integer decimation."""
    y1 = signal.convolve(y, wbt_fil)[0 :: int(wbt_decifac)]
    y2 = signal.convolve(y1, pc_fil)[0 :: int(pc_decifac)]
    t2 = np.arange(y2.size) * 1 / fs * wbt_decifac * pc_decifac
    return y2, t2


def transmit_signal(waveform_mode, fs, tau, slope, f_start, f_stop, f_nominal, filt, drop_last_hanning_zero=False):
    """ek80_complex.py:211-282 for ONE channel with unique tx parameters.

    ``filt`` = dict(wbt_fil, wbt_decifac, pc_fil, pc_decifac).  CW uses f_nominal for both sweep ends.
    """
    if waveform_mode == "CW":
        f_start = f_stop = f_nominal
    y, _ = tapered_chirp(fs, tau, slope, f_start, f_stop, drop_last_hanning_zero)
    return filter_decimate_chirp(y, fs, filt["wbt_fil"], filt["wbt_decifac"], filt["pc_fil"], filt["pc_decifac"])


def tau_effective(ytx, fs_deci, waveform_mode):
    """ek80_complex.py:183-191 for one channel."""
    if waveform_mode == "BB":
        ytxa = signal.convolve(ytx, np.flip(np.conj(ytx))) / np.linalg.norm(ytx) ** 2
        ptxa = np.abs(ytxa) ** 2
    else:
        ptxa = np.abs(ytx) ** 2
    return ptxa.sum() / (ptxa.max() * fs_deci)


def compress_pulse(backscatter, chirp):
    """ek80_complex.py:285-369.  backscatter (C,P,R,B) complex128; chirp = per-channel replica list.

    NaN -> 0 before, per (ping, beam) slab: all-zero slab returned unchanged, otherwise each channel is
    ``convolve(x, flip(conj(tx)), 'full')[M-1:]`` stored as complex64; NaN restored afterwards.
    """
    bs = np.asarray(backscatter)
    C, P, R, B = bs.shape
    replicas = [np.flipud(np.conj(np.asarray(tx))) for tx in chirp]
    nan_mask = np.isnan(bs)
    x0 = np.where(nan_mask, 0.0 + 0.0j, bs)
    pc = np.zeros((C, P, R, B), dtype=np.complex64)
    for p in range(P):
        for b in range(B):
            slab = x0[:, p, :, b]  # (C, R)
            if np.all(slab == 0.0 + 0.0j):
                continue
            for c in range(C):
                rep = replicas[c]
                pc[c, p, :, b] = signal.convolve(slab[c], rep, mode="full")[rep.size - 1 :]
    out = pc.astype(np.complex128)  # xr.where(nan_mask, np.nan, pc) promotes with float64 NaN
    out[nan_mask] = np.nan
    return out
