"""EK80 transmit replica synthesis (host side; per-channel 1-D arrays of 10^2-10^4 samples - stays on
the CPU by design, SURVEY.md 8a #a9) and the pulse-compression entry point (device).

Same names and argument meaning as echopype/calibrate/ek80_complex.py: tapered_chirp :12-52,
filter_decimate_chirp :55-80, get_vend_filter_EK80 :83-127, get_filter_coeff :130-159,
get_tau_effective :162-208, get_transmit_signal :211-282, get_norm_fac :372-391.
"""

from collections import defaultdict
from typing import Dict

import numpy as np
from scipy import signal

from ..dataset import DataArray, Dataset

FILTER_IMAG, FILTER_REAL, DECIMATION = "coeffs_imag", "coeffs_real", "deci_fac"


def _scalar(x):
    return float(np.asarray(x).reshape(-1)[0])


def tapered_chirp(fs, transmit_duration_nominal, slope, transmit_frequency_start, transmit_frequency_stop,
                  drop_last_hanning_zero=False):
    """Chirp replica (Andersen / CRIMAC): cosine sweep with Hann-tapered ends, peak-normalised."""
    tau, f0, f1 = _scalar(transmit_duration_nominal), _scalar(transmit_frequency_start), _scalar(transmit_frequency_stop)
    fs, slope = _scalar(fs), _scalar(slope)
    n_tx = int(np.floor(tau * np.float32(fs)))
    t = np.linspace(0, n_tx - 1, num=n_tx) * 1 / fs
    sweep = np.cos(np.pi * (f1 - f0) / tau * t * t + 2 * np.pi * f0 * t)
    n_win = int(np.round(tau * fs * slope * 2.0))
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(0, n_win, 1) / (n_win - 1)))
    head = hann[: n_win // 2]
    tail = hann[n_win // 2 : -1] if drop_last_hanning_zero else hann[n_win // 2 :]
    sweep[: head.size] *= head
    sweep[n_tx - tail.size :] *= tail
    return sweep / np.max(sweep), t


def filter_decimate_chirp(coeff_ch: Dict, y_ch: np.ndarray, fs: float):
    """WBT filter + decimate, then PC filter + decimate (both full convolutions)."""
    stage1 = signal.convolve(y_ch, coeff_ch["wbt_fil"])[0 :: int(coeff_ch["wbt_decifac"])]
    stage2 = signal.convolve(stage1, coeff_ch["pc_fil"])[0 :: int(coeff_ch["pc_decifac"])]
    t = np.arange(stage2.size) * 1 / fs * coeff_ch["wbt_decifac"] * coeff_ch["pc_decifac"]
    return stage2, t


def get_vend_filter_EK80(vend: Dataset, channel_id, filter_name, param_type):
    names = [f"{filter_name}_{FILTER_IMAG}", f"{filter_name}_{FILTER_REAL}", f"{filter_name}_{DECIMATION}"]
    if not all(v in vend for v in names):
        return None
    ci = int(np.flatnonzero(np.asarray(vend["channel"].values) == channel_id)[0])
    if param_type == "coeff":
        re = np.asarray(vend[names[1]].values)
        im = np.asarray(vend[names[0]].values)
        if re.ndim == 3:  # (channel, filter_time, n): first filter_time only (ek80_complex.py:148-149)
            ax = vend[names[1]].dims.index("filter_time")
            re, im = np.take(re, 0, axis=ax), np.take(im, 0, axis=ax)
        v = re[ci] + 1j * im[ci]
        return v[~np.isnan(v)]
    d = np.asarray(vend[names[2]].values)
    if d.ndim == 2:
        d = np.take(d, 0, axis=vend[names[2]].dims.index("filter_time"))
    return d[ci]


def get_filter_coeff(vend: Dataset) -> Dict:
    coeff = defaultdict(dict)
    for ch_id in vend["channel"].values:
        coeff[ch_id]["wbt_fil"] = get_vend_filter_EK80(vend, ch_id, "WBT", "coeff")
        coeff[ch_id]["pc_fil"] = get_vend_filter_EK80(vend, ch_id, "PC", "coeff")
        coeff[ch_id]["wbt_decifac"] = get_vend_filter_EK80(vend, ch_id, "WBT", "decimation")
        coeff[ch_id]["pc_decifac"] = get_vend_filter_EK80(vend, ch_id, "PC", "decimation")
    return coeff


def get_tau_effective(ytx_dict, fs_deci_dict, waveform_mode, channel, ping_time=None) -> DataArray:
    vals = []
    for ch, ytx in ytx_dict.items():
        if waveform_mode == "BB":
            acorr = signal.convolve(ytx, np.flip(np.conj(ytx))) / np.linalg.norm(ytx) ** 2
            p = np.abs(acorr) ** 2
        elif waveform_mode == "CW":
            p = np.abs(ytx) ** 2
        else:
            raise ValueError("waveform_mode must be 'CW' or 'BB'")
        vals.append(float(p.sum() / (p.max() * _scalar(fs_deci_dict[ch]))))
    return DataArray(np.asarray(vals), dims=["channel"], coords={"channel": np.asarray(getattr(channel, "values", channel))})


def _unique_non_nan(row):
    """np.unique(row) without NaN; rows that hold one value (the normal case) skip the sort."""
    if row.size and row[0] == row[0] and bool(np.all(row == row[0])):
        return row[:1].copy()
    v = np.unique(row)
    return v[~np.isnan(v)]


def get_transmit_signal(beam: Dataset, coeff: Dict, waveform_mode: str, fs, drop_last_hanning_zero: bool = False):
    if waveform_mode == "BB":
        ttype = np.asarray(beam["transmit_type"].values)
        if ttype.size and ttype.flat[0] == "CW" and np.all(ttype == "CW"):
            raise TypeError("File does not contain BB mode complex samples!")
    y_all, y_time_all = {}, {}
    names = ["transmit_duration_nominal", "slope", "transmit_frequency_start", "transmit_frequency_stop"]
    chans = np.asarray(beam["channel"].values)
    tables = {}
    for p in names:
        if waveform_mode == "CW" and p in ("transmit_frequency_start", "transmit_frequency_stop"):
            tables[p] = np.asarray(beam["frequency_nominal"].values, dtype=np.float64).reshape(len(chans), -1)
        else:
            tables[p] = np.asarray(beam[p].transpose("channel", "ping_time").values, dtype=np.float64)
    for ci, ch in enumerate(chans):
        fs_chan = _scalar(fs.sel(channel=ch).values) if isinstance(fs, DataArray) else fs
        tx = {}
        for p in names:
            cw_freq = waveform_mode == "CW" and p in ("transmit_frequency_start", "transmit_frequency_stop")
            tx[p] = np.unique(tables[p][ci]) if cw_freq else _unique_non_nan(tables[p][ci])
            if tx[p].size != 1:
                raise TypeError("File contains changing %s!" % p)
        y_ch, _ = tapered_chirp(fs=fs_chan, drop_last_hanning_zero=drop_last_hanning_zero, **tx)
        y_all[ch], y_time_all[ch] = filter_decimate_chirp(coeff_ch=coeff[ch], y_ch=y_ch, fs=fs_chan)
    return y_all, y_time_all


def get_norm_fac(chirp: Dict) -> DataArray:
    return DataArray(
        np.asarray([np.linalg.norm(tx) ** 2 for tx in chirp.values()]), dims=["channel"], coords={"channel": np.asarray(list(chirp))}
    )
