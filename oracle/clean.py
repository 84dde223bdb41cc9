"""Oracle: De Robertis & Higginbottom background-noise estimate / removal.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows echopype/clean/api.py:362-433
(estimate_background_noise), :436-511 (remove_background_noise) and echopype/clean/utils.py:13-26
(extract_dB), :380-401 (attrs incl. actual_range).  xarray's ``coarsen(..., boundary="pad").mean()``
is restated as NaN-pad + reshape + nanmean; ``reindex(method="ffill")`` as index // ping_num.
Pinned by outputs of the reference's own estimate_ / remove_background_noise (tests/golden/make_golden_calibrate.py,
tests/test_reference_pinned.py) and by its known-answer test tests/clean/test_noise.py:902-987 (restated in
tests/test_oracle_golden.py).  The noise masks further down (impulse, transient, attenuated signal) are pinned by outputs of
the reference's own mask functions and workers (tests/golden/make_golden_masks.py, tests/test_reference_pinned_masks.py);
downsample_upsample_along_depth (flox) is the one restated worker.
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


# ---- impulse / transient noise masks with index binning ---------------------------------------
def samples_per_depth_bin(range_var, depth_bin):
    """clean/utils.py:131-133 / 280-282: range samples per depth bin of each channel; range_var (C, P, R)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.ceil(depth_bin / np.nanmean(np.diff(range_var, axis=2), axis=(1, 2))).astype(int)


def index_binning_downsample_upsample(Sv, range_var, depth_bin):
    """clean/utils.py:263-317: per channel coarsen(range_sample=n, boundary="pad").mean(skipna=True) of 10^(Sv/10),
    back to dB, block b re-labelled with range_sample n*b and reindexed onto every range_sample with ffill."""
    C, P, R = Sv.shape
    out = np.empty((C, P, R))
    for c, n in enumerate(samples_per_depth_bin(range_var, depth_bin)):
        nb = -(-R // n)
        lin = np.full((P, nb * n), np.nan)
        lin[:, :R] = log2lin(Sv[c])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            coarse = lin2log(np.nanmean(lin.reshape(P, nb, n), axis=2))
        out[c] = np.repeat(coarse, n, axis=1)[:, :R]
    return out


def echopy_impulse_noise_mask(Sv, num_side_pings, impulse_noise_threshold):
    """clean/utils.py:320-337; Sv is one channel laid out (range_sample, ping_time)."""
    k = num_side_pings
    dummy = np.zeros((Sv.shape[0], k)) * np.nan
    fwd = Sv - np.c_[Sv[:, k:], dummy]
    bwd = Sv - np.c_[dummy, Sv[:, 0:-k]]
    fwd[np.isnan(fwd)] = np.inf
    bwd[np.isnan(bwd)] = np.inf
    return (fwd > impulse_noise_threshold) & (bwd > impulse_noise_threshold)


def mask_impulse_noise_index_binning(Sv, range_var, depth_bin, num_side_pings, impulse_noise_threshold):
    """clean/api.py:169-266 with use_index_binning=True.  Returns (mask, upsampled_Sv), both (C, P, R); the reference's
    apply_ufunc output carries the same values with dims (channel, range_sample, ping_time)."""
    up = index_binning_downsample_upsample(Sv, range_var, depth_bin)
    mask = np.stack([echopy_impulse_noise_mask(up[c].T, num_side_pings, impulse_noise_threshold).T for c in range(Sv.shape[0])])
    return mask, up


def index_binning_pool_Sv(Sv, range_var, depth_bin, num_side_pings, exclude_above, func=np.nanmean):
    """clean/utils.py:109-189 with func = np.nanmean / np.nanmedian (clean/api.py:132-145): per channel scipy.ndimage.generic_filter (what
    dask_image.ndfilters.generic_filter evaluates) of 10^(Sv/10) over [(2 k + 1) pings, (2 n + 1) range samples],
    mode="reflect", on the volume sliced at the first flat index deeper than exclude_above (:141)."""
    from scipy import ndimage

    C, P, R = Sv.shape
    pooled = np.full((C, P, R), np.nan)
    m0 = int(np.argmin(range_var <= exclude_above))
    if m0 >= R:
        return pooled, m0
    for c, n in enumerate(samples_per_depth_bin(range_var, depth_bin)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            filt = ndimage.generic_filter(log2lin(Sv[c][:, m0:]), function=func,
                                          size=[2 * num_side_pings + 1, 2 * n + 1], mode="reflect")
            pooled[c][:, m0:] = lin2log(filt)
    return pooled, m0


def mask_transient_noise_index_binning(Sv, range_var, depth_bin, num_side_pings, exclude_above, transient_noise_threshold,
                                       func=np.nanmean):
    """clean/api.py:30-166 with use_index_binning=True.  Returns (mask, pooled_Sv)."""
    pooled, _ = index_binning_pool_Sv(Sv, range_var, depth_bin, num_side_pings, exclude_above, func)
    with np.errstate(invalid="ignore"):
        return (Sv - pooled) > transient_noise_threshold, pooled


def downsample_upsample_along_depth(Sv, range_var, depth_bin):
    """clean/utils.py:192-260 (use_index_binning=False): depth intervals [e_k, e_k+1) from np.arange(min, max + bin, bin);
    per (channel, ping) nanmean of 10^(Sv/10) over the samples whose depth falls into the interval (flox nanmean, NaN depths
    dropped, empty interval -> NaN); every sample is then given the value of the interval np.digitize assigns it to
    (first sample of each interval as the new range_sample label, reindex with ffill).  Returns (downsampled, upsampled).
    Like the reference it needs depth to increase along range_sample and every interval to occur in every ping."""
    C, P, R = Sv.shape
    edges = np.arange(np.nanmin(range_var), np.nanmax(range_var) + depth_bin, depth_bin)
    lefts = edges[:-1]
    nb = len(lefts)
    lin = log2lin(Sv)
    down = np.full((C, P, nb), np.nan)
    up = np.full((C, P, R), np.nan)
    assign = np.digitize(range_var, lefts)
    for c in range(C):
        for p in range(P):
            d = range_var[c, p]
            for b in range(nb):
                m = (d >= edges[b]) & (d < edges[b + 1]) & ~np.isnan(lin[c, p])
                if m.any():
                    down[c, p, b] = lin2log(lin[c, p][m].mean())
            uniq, first = np.unique(assign[c, p], return_index=True)
            if len(uniq) != nb:
                raise ValueError("conflicting sizes for dimension 'depth_bins'")  # assign_coords in the reference
            j = np.searchsorted(first, np.arange(R), side="right") - 1  # reindex(..., method="ffill")
            up[c, p] = np.where(j >= 0, down[c, p][np.maximum(j, 0)], np.nan)
    return down, up


def mask_impulse_noise_depth_binning(Sv, range_var, depth_bin, num_side_pings, impulse_noise_threshold):
    """clean/api.py:169-266 with use_index_binning=False.  Returns (mask, upsampled_Sv), both (C, P, R)."""
    _, up = downsample_upsample_along_depth(Sv, range_var, depth_bin)
    mask = np.stack([echopy_impulse_noise_mask(up[c].T, num_side_pings, impulse_noise_threshold).T for c in range(Sv.shape[0])])
    return mask, up


def pool_Sv(Sv, range_var, depth_bin, num_side_pings, exclude_above, func=np.nanmean):
    """clean/utils.py:28-105 with func = np.nanmean / np.nanmedian (the use_index_binning=False path of mask_transient_noise): for every
    sample whose depth d keeps d +- depth_bin inside the depth extent of the dataset and below exclude_above and whose
    ping keeps p - k >= 0 and p + k <= n_ping, the nanmean of 10^(Sv/10) over the samples of the same channel with
    d - depth_bin <= depth <= d + depth_bin in the pings p - k .. p + k, in dB; NaN elsewhere."""
    C, P, R = Sv.shape
    dmin, dmax = np.nanmin(range_var), np.nanmax(range_var)
    lin = log2lin(Sv)
    pooled = np.full((C, P, R), np.nan)
    k = num_side_pings
    for c in range(C):
        for p in range(P):
            if not (p - k >= 0 and p + k <= P):
                continue
            q0, q1 = p - k, min(p + k, P - 1)
            dep = range_var[c, q0 : q1 + 1]
            win = lin[c, q0 : q1 + 1]
            for n in range(R):
                d = range_var[c, p, n]
                if not ((d - depth_bin >= dmin) and (d + depth_bin <= dmax) and (d - depth_bin >= exclude_above)):
                    continue
                with np.errstate(invalid="ignore"):
                    m = (d - depth_bin <= dep) & (dep <= d + depth_bin) & ~np.isnan(win)
                if m.any():
                    pooled[c, p, n] = lin2log(func(win[m]))
    return pooled


def mask_transient_noise_depth_binning(Sv, range_var, depth_bin, num_side_pings, exclude_above, transient_noise_threshold,
                                       func=np.nanmean):
    """clean/api.py:30-166 with use_index_binning=False.  Returns (mask, pooled_Sv)."""
    pooled = pool_Sv(Sv, range_var, depth_bin, num_side_pings, exclude_above, func)
    with np.errstate(invalid="ignore"):
        return (Sv - pooled) > transient_noise_threshold, pooled


# ---- attenuated-signal mask (clean/api.py:269-359, clean/utils.py:337-377) ------------------------------------------
def echopy_attenuated_signal_mask(Sv, range_var, upper_limit_sl, lower_limit_sl, num_side_pings, attenuation_signal_threshold):
    """clean/utils.py:337-377 for one channel, Sv and range_var laid out (ping_time, range_sample).  The layer is
    Sv[p, up:lw] with up / lw the samples nearest to the limits (np.argmin: first minimum; the first NaN when the range
    row has one); ping p is assessed when p - n >= 0, p + n <= P - 1 and the layer holds a valid sample; it is masked as
    a whole when its linear-domain median is less than the threshold (dB) above the median of pings p - n .. p + n - 1."""
    P = Sv.shape[0]
    n = num_side_pings
    mask = np.zeros(Sv.shape, dtype=bool)
    for p in range(P):
        up = int(np.argmin(np.abs(range_var[p] - upper_limit_sl)))
        lw = int(np.argmin(np.abs(range_var[p] - lower_limit_sl)))
        if p - n < 0 or p + n > P - 1:
            continue
        layer = Sv[p, up:lw]
        if np.all(np.isnan(layer)):
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            with np.errstate(divide="ignore", invalid="ignore"):
                ping_median = lin2log(np.nanmedian(log2lin(layer)))
                block_median = lin2log(np.nanmedian(log2lin(Sv[p - n : p + n, up:lw])))
                if ping_median - block_median < attenuation_signal_threshold:
                    mask[p, :] = True
    return mask


def attenuated_signal_margin(Sv, range_var, upper_limit_sl, lower_limit_sl, num_side_pings, attenuation_signal_threshold):
    """(ping median - block median) - threshold per ping of one channel (NaN where the ping is not assessed): the tests
    use it to set aside pings that sit on the threshold to within rounding."""
    P = Sv.shape[0]
    n = num_side_pings
    out = np.full(P, np.nan)
    for p in range(n, P - n):
        up = int(np.argmin(np.abs(range_var[p] - upper_limit_sl)))
        lw = int(np.argmin(np.abs(range_var[p] - lower_limit_sl)))
        layer = Sv[p, up:lw]
        if np.all(np.isnan(layer)) or n == 0:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            out[p] = lin2log(np.nanmedian(log2lin(layer))) - lin2log(np.nanmedian(log2lin(Sv[p - n : p + n, up:lw]))) \
                - attenuation_signal_threshold
    return out


def mask_attenuated_signal(Sv, range_var, upper_limit_sl, lower_limit_sl, num_side_pings, attenuation_signal_threshold):
    """clean/api.py:269-359 on (C, P, R) arrays, limits in metres and threshold in dB already parsed: all False when the
    searching range lies outside the extent of range_var (:330-334), else the per-channel mask (apply_ufunc, vectorize)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        if (upper_limit_sl > np.nanmax(range_var)) or (lower_limit_sl < np.nanmin(range_var)):
            return np.zeros(range_var.shape, dtype=bool)
    return np.stack([echopy_attenuated_signal_mask(Sv[c], range_var[c], upper_limit_sl, lower_limit_sl, num_side_pings,
                                                   attenuation_signal_threshold) for c in range(Sv.shape[0])])
