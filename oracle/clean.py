"""Oracle: De Robertis & Higginbottom background-noise estimate / removal.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows echopype/clean/api.py:362-433
(estimate_background_noise), :436-511 (remove_background_noise) and echopype/clean/utils.py:13-26
(extract_dB), :380-401 (attrs incl. actual_range).  xarray's ``coarsen(..., boundary="pad").mean()``
is restated as NaN-pad + reshape + nanmean; ``reindex(method="ffill")`` as index // ping_num.
Pinned by the reference's known-answer test tests/clean/test_noise.py:902-987 (restated in
tests/test_oracle_golden.py).
"""

import re
import warnings

import numpy as np


def extract_dB(dB_str):
    """clean/utils.py:13-26."""
    if not isinstance(dB_str, str):
        raise TypeError(
            "Decibal input must be a string formatted as `NUMdB` or `NUMdb."
            f"Cannot be of type `{type(dB_str)}`."
        )
    m = re.search(r"^[-+]?\d+\.?\d*(?:dB|db)$", dB_str, flags=re.IGNORECASE)
    if m:
        return float(m.group(0)[:-2])
    raise ValueError("Decibal string must be formatted as 'NUMdB' or `NUMdb")


def log2lin(x):
    """utils/compute.py:13-27."""
    return 10 ** (x / 10)


def lin2log(x):
    """utils/compute.py:29-42."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return 10 * np.log10(x)


def coarsen_pad(a, ping_num, range_sample_num):
    """xarray coarsen(ping_time=ping_num, range_sample=range_sample_num, boundary='pad') windows.

    a (C,P,R) -> (C, nP, ping_num, nR, range_sample_num) view of the NaN-padded array.
    """
    C, P, R = a.shape
    nP = -(-P // ping_num)
    nR = -(-R // range_sample_num)
    pad = np.full((C, nP * ping_num, nR * range_sample_num), np.nan, dtype=np.float64)
    pad[:, :P, :R] = a
    return pad.reshape(C, nP, ping_num, nR, range_sample_num)


def coarsen_mean(a, ping_num, range_sample_num):
    w = coarsen_pad(a, ping_num, range_sample_num)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmean(w, axis=(2, 4))


def coarsen_min(a, ping_num, range_sample_num):
    w = coarsen_pad(a, ping_num, range_sample_num)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmin(w, axis=(2, 4))


def transmission_loss(echo_range, sound_absorption):
    """clean/api.py:397-398.  sound_absorption broadcastable to (C,P,1)."""
    rng = np.asarray(echo_range, dtype=np.float64)
    C, P, _ = rng.shape
    a = np.asarray(sound_absorption, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None, None]
    elif a.ndim == 2:
        a = a[:, :, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        spreading = 20 * np.log10(np.where(rng >= 1, rng, 1))  # NaN range -> 1 -> 0 dB
        absorb = 2 * a * rng
    return spreading + absorb


def estimate_background_noise(Sv, echo_range, sound_absorption, ping_num, range_sample_num, background_noise_max=None):
    """clean/api.py:362-433.  Returns Sv_noise (C,P,R) float64."""
    if background_noise_max is not None:
        background_noise_max = extract_dB(background_noise_max)
    Sv = np.asarray(Sv, dtype=np.float64)
    C, P, R = Sv.shape
    TL = transmission_loss(echo_range, sound_absorption)
    power_cal = log2lin(Sv - TL)
    binned = lin2log(coarsen_mean(power_cal, ping_num, range_sample_num))  # (C, nP, nR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        noise = np.nanmin(binned, axis=2)  # (C, nP); all-NaN -> NaN
    if background_noise_max is not None:
        with np.errstate(invalid="ignore"):
            noise = np.where(noise < background_noise_max, noise, background_noise_max)
    up = noise[:, np.arange(P) // ping_num]  # reindex ffill
    return up[:, :, None] + TL


def remove_background_noise(
    Sv, echo_range, sound_absorption, ping_num, range_sample_num, background_noise_max=None, SNR_threshold="3.0dB"
):
    """clean/api.py:436-511.  Returns dict(Sv_noise, Sv_corrected, attrs_noise, attrs_corrected)."""
    snr = extract_dB(SNR_threshold) if SNR_threshold is not None else None
    Sv = np.asarray(Sv, dtype=np.float64)
    Sv_noise = estimate_background_noise(
        Sv, echo_range, sound_absorption, ping_num, range_sample_num, background_noise_max
    )
    lin = log2lin(Sv) - log2lin(Sv_noise)
    with np.errstate(invalid="ignore"):
        corr = lin2log(np.where(lin > 0, lin, np.nan))
        corr = np.where(corr - Sv_noise > snr, corr, np.nan)
    nm = background_noise_max  # clean/api.py:490-501 passes the caller's string (or None) through

    def _attrs(a, kind):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            return {
                "long_name": f"Volume backscattering strength, {kind} (Sv re 1 m-1)",
                "units": "dB",
                "actual_range": [round(float(np.nanmin(a)), 2), round(float(np.nanmax(a)), 2)],
                "noise_ping_num": ping_num,
                "noise_range_sample_num": range_sample_num,
                "SNR_threshold": snr,
                "noise_max": nm,
            }

    return {
        "Sv_noise": Sv_noise,
        "Sv_corrected": corr,
        "attrs_noise": _attrs(Sv_noise, "noise"),
        "attrs_corrected": _attrs(corr, "corrected"),
    }
