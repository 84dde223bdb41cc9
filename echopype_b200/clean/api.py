"""Background-noise estimate / removal (De Robertis & Higginbottom 2007) with the reference's signatures
(echopype/clean/api.py:362-433 estimate_background_noise, :436-511 remove_background_noise); the array work
runs in two kernels of libepb200 (epb_noise_estimate: tile means + block min, epb_noise_apply: subtraction,
SNR gate and the actual_range extrema) - 16 B read + 8 B written per sample."""

import numpy as np
import torch

from .. import kernels
from ..dataset import DataArray, Dataset, as_dataset
from ..device import ParamPack, require_cuda, to_device_f32
from ..utils.prov import add_processing_level, echopype_prov_attrs, insert_input_processing_level
from ..commongrid.utils import _parse_x_bin
from .utils import extract_dB, noise_attrs

DIMS = ("channel", "ping_time", "range_sample")


def _device_inputs(ds_Sv):
    for v in ("Sv", "echo_range", "sound_absorption"):
        if v not in ds_Sv:
            raise KeyError(v)
    sv = ds_Sv["Sv"]
    if tuple(sv.dims) != DIMS:
        raise ValueError(f"Sv must have dims {DIMS}, got {tuple(sv.dims)}")
    dev = require_cuda()
    C, P, R = sv.shape
    Sv = to_device_f32(sv.data, dev)
    rng = to_device_f32(ds_Sv["echo_range"].data, dev)
    if tuple(rng.shape) != (C, P, R):
        raise ValueError("echo_range must have the shape of Sv")
    pack = ParamPack(C, P, dev)
    alpha = ds_Sv["sound_absorption"]
    a = np.asarray(alpha.values, dtype=np.float64)
    if alpha.dims == ("ping_time", "channel"):
        a = a.T
    return Sv, rng, pack.cp(a), pack, C, P, R


def _estimate(ds_Sv, ping_num, range_sample_num, background_noise_max):
    if not (isinstance(ping_num, (int, np.integer)) and ping_num > 0):
        raise ValueError("ping_num must be a positive integer")
    if not (isinstance(range_sample_num, (int, np.integer)) and range_sample_num > 0):
        raise ValueError("range_sample_num must be a positive integer")
    noise_max = extract_dB(background_noise_max) if background_noise_max is not None else None
    Sv, rng, alpha, pack, C, P, R = _device_inputs(ds_Sv)
    noise = kernels.noise_estimate(Sv, rng, alpha, pack, C, P, R, int(ping_num), int(range_sample_num), noise_max)
    return Sv, rng, alpha, pack, noise, C, P, R


def estimate_background_noise(ds_Sv, ping_num: int, range_sample_num: int, background_noise_max: str = None) -> DataArray:
    """
    Estimate background noise by computing mean calibrated power of a collection of pings.

    Parameters and behaviour as echopype.clean.estimate_background_noise: ``ds_Sv`` holds ``Sv``,
    ``echo_range`` and ``sound_absorption``; returns the noise estimate ``Sv_noise`` (channel, ping_time,
    range_sample), device resident.
    """
    ds_Sv = as_dataset(ds_Sv)
    Sv, rng, alpha, pack, noise, C, P, R = _estimate(ds_Sv, ping_num, range_sample_num, background_noise_max)
    sn, _, _ = kernels.noise_apply(Sv, rng, alpha, pack, noise, C, P, R, int(ping_num), 0.0, want_noise=True, want_corr=False,
                                   want_minmax=False)
    return DataArray(sn, DIMS, coords={d: ds_Sv[d].values for d in DIMS if d in ds_Sv}, name="Sv_noise")


@add_processing_level("L*B")
def remove_background_noise(ds_Sv, ping_num: int, range_sample_num: int, background_noise_max: str = None,
                            SNR_threshold: str = "3.0dB") -> Dataset:
    """
    Remove noise by using estimates of background noise from mean calibrated power of a collection of pings
    (De Robertis & Higginbottom 2007).  Returns the input dataset with ``Sv_corrected`` and ``Sv_noise`` added.
    """
    ds_Sv = as_dataset(ds_Sv)
    if SNR_threshold is None:
        raise TypeError("SNR_threshold must be a string such as '3.0dB'")  # the reference fails on None as well
    snr = extract_dB(SNR_threshold)
    Sv, rng, alpha, pack, noise, C, P, R = _estimate(ds_Sv, ping_num, range_sample_num, background_noise_max)
    sn, sc, mm = kernels.noise_apply(Sv, rng, alpha, pack, noise, C, P, R, int(ping_num), snr)
    lo_n, hi_n, lo_c, hi_c = mm.tolist()
    if lo_n == float("inf"):
        lo_n = hi_n = float("nan")
    if lo_c == float("inf"):
        lo_c = hi_c = float("nan")
    out = ds_Sv.copy()
    out["Sv_noise"] = DataArray(sn, DIMS, attrs=noise_attrs("noise", lo_n, hi_n, ping_num, range_sample_num, snr, background_noise_max))
    trusted = isinstance(ds_Sv["Sv"].law, dict) and ds_Sv["Sv"].law.get("kind") == "derived"
    out["Sv_corrected"] = DataArray(sc, DIMS, attrs=noise_attrs("corrected", lo_c, hi_c, ping_num, range_sample_num, snr, background_noise_max),
                                    law={"kind": "derived"} if trusted else None)  # NaN wherever Sv is
    prov_dict = echopype_prov_attrs(process_type="processing")
    prov_dict["processing_function"] = "clean.remove_background_noise"
    out = out.assign_attrs(prov_dict)
    out = insert_input_processing_level(out, input_ds=ds_Sv)
    return out


# ---- impulse / transient noise masks with index binning ---------------------------------------
def _index_binning_inputs(ds_Sv, range_var, what):
    """Sv and the range variable as device arrays plus the number of range samples per depth bin of each channel
    (clean/utils.py:131-133, 280-282: ceil(depth_bin / nanmean(diff(range_var, axis=range_sample))))."""
    if range_var not in ["echo_range", "depth"]:
        raise ValueError("`range_var` must be either `echo_range` or `depth`.")
    if range_var not in ds_Sv:
        raise ValueError(f"Masking {what} noise requires `{range_var}` data variable in `ds_Sv`.")
    sv = ds_Sv["Sv"]
    if tuple(sv.dims) != DIMS:
        raise ValueError(f"Sv must have dims {DIMS}, got {tuple(sv.dims)}")
    dev = require_cuda()
    C, P, R = sv.shape
    Sv = to_device_f32(sv.data, dev)
    rng = to_device_f32(ds_Sv[range_var].data, dev)
    if tuple(rng.shape) != (C, P, R):
        raise ValueError(f"{range_var} must have the shape of Sv")
    return Sv, rng, C, P, R


def _samples_per_bin(rng, depth_bin, C, P, R):
    mean_diff = kernels.range_diff_mean(rng, C, P, R)
    if not np.all(np.isfinite(mean_diff)) or np.any(mean_diff <= 0):
        raise ValueError("the range variable has no valid, increasing sample spacing in some channel")
    return np.ceil(depth_bin / mean_diff).astype(int)


def _mask_coords(ds_Sv):
    return {d: ds_Sv[d].values for d in DIMS if d in ds_Sv}


def mask_impulse_noise(ds_Sv, depth_bin: str = "5m", num_side_pings: int = 2, impulse_noise_threshold: str = "10.0dB",
                       range_var: str = "depth", use_index_binning: bool = False) -> DataArray:
    """
    Locate and create a mask for impulse noise using a ping-wise two-sided comparison (Ryan et al. 2015; arguments
    as echopype.clean.mask_impulse_noise, clean/api.py:169-266).

    Sv is averaged in the linear domain over depth bins - intervals of ``range_var`` VALUES per ping (the default), or
    with ``use_index_binning=True`` blocks of ``ceil(depth_bin / mean sample spacing)`` range samples per channel -
    forward filled back to every sample and compared with the pings ``num_side_pings`` before and after.  Returns a boolean (uint8 on the device) DataArray with dims
    (channel, ping_time, range_sample); the reference's ``apply_ufunc`` hands the same values back with the last two
    dims swapped.
    """
    ds_Sv = as_dataset(ds_Sv)
    if range_var not in ["echo_range", "depth"]:
        raise ValueError("`range_var` must be either `echo_range` or `depth`.")
    if range_var not in ds_Sv and not use_index_binning:
        raise ValueError(f"Masking impulse noise requires `{range_var}` data variable in `ds_Sv`.")
    thr = extract_dB(impulse_noise_threshold)
    depth_bin = _parse_x_bin(depth_bin, "range_bin")
    if not (isinstance(num_side_pings, (int, np.integer)) and num_side_pings >= 1):
        raise ValueError("num_side_pings must be a positive integer")
    Sv, rng, C, P, R = _index_binning_inputs(ds_Sv, range_var, "impulse")
    if not use_index_binning:
        # clean/utils.py:192-260: intervals of depth VALUES, np.arange(min, max + depth_bin, depth_bin), per ping
        lo, hi, _ = kernels.minmax(rng)
        if not (np.isfinite(lo) and np.isfinite(hi)):
            raise ValueError(f"`{range_var}` has no valid values")
        edges = np.arange(float(lo), float(hi) + depth_bin, depth_bin)
        mask, _, _, _ = kernels.impulse_noise_mask_depth(Sv, rng, edges, C, P, R, int(num_side_pings), thr, want_upsampled=False)
        return DataArray(mask, DIMS, coords=_mask_coords(ds_Sv), name="Sv")
    nsamp = _samples_per_bin(rng, depth_bin, C, P, R)
    mask, _ = kernels.impulse_noise_mask(Sv, nsamp, C, P, R, int(num_side_pings), thr)
    return DataArray(mask, DIMS, coords=_mask_coords(ds_Sv), name="Sv")


def mask_transient_noise(ds_Sv, func: str = "nanmean", depth_bin: str = "10m", num_side_pings: int = 25,
                         exclude_above: str = "250.0m", transient_noise_threshold: str = "12.0dB", range_var: str = "depth",
                         use_index_binning: bool = False, chunk_dict: dict = {}) -> DataArray:
    """
    Locate and create a mask for transient noise using a pooling comparison (Ryan et al. 2015; arguments as
    echopype.clean.mask_transient_noise, clean/api.py:30-166).

    Pooled Sv is the mean or the median (linear domain, ``func="nanmean"`` / ``"nanmedian"``) over ``2 num_side_pings + 1`` pings and, by default, the
    samples within ``depth_bin`` of the sample's own depth (windows of ``range_var`` VALUES, valid where the window stays
    inside the data and below ``exclude_above``); with ``use_index_binning=True`` over ``2 ceil(depth_bin / mean sample
    spacing) + 1`` range samples with reflected borders, below the first sample deeper than ``exclude_above``.  The mask is
    ``Sv - pooled_Sv > transient_noise_threshold``.  ``chunk_dict`` (dask chunking of the reference) is ignored.
    """
    ds_Sv = as_dataset(ds_Sv)
    if range_var not in ["echo_range", "depth"]:
        raise ValueError("`range_var` must be either `echo_range` or `depth`.")
    if range_var not in ds_Sv and not use_index_binning:
        raise ValueError(f"Masking transient noise requires `{range_var}` data variable in `ds_Sv`.")
    if func != "nanmean" and func != "nanmedian":
        raise ValueError(f"Input `func` is `{func}`. `func` must be `nanmean` or `nanmedian`.")
    thr = extract_dB(transient_noise_threshold)
    depth_bin = _parse_x_bin(depth_bin, "range_bin")
    exclude_above = _parse_x_bin(exclude_above, "range_bin")
    median = func == "nanmedian"  # a selection per sample (radix select over the window) instead of window sums
    if not (isinstance(num_side_pings, (int, np.integer)) and num_side_pings >= 0):
        raise ValueError("num_side_pings must be a non-negative integer")
    Sv, rng, C, P, R = _index_binning_inputs(ds_Sv, range_var, "transient")
    if not use_index_binning:
        # clean/utils.py:28-105 pool_Sv: windows of depth VALUES (d +- depth_bin) over pings p - k .. p + k
        lo, hi, _ = kernels.minmax(rng)
        pool = kernels.transient_noise_mask_depth_median if median else kernels.transient_noise_mask_depth
        mask, _ = pool(Sv, rng, C, P, R, lo, hi, depth_bin, exclude_above, int(num_side_pings), thr)
        return DataArray(mask, DIMS, coords=_mask_coords(ds_Sv), name="Sv")
    nsamp = _samples_per_bin(rng, depth_bin, C, P, R)
    # clean/utils.py:141: np.argmin over the WHOLE (channel, ping_time, range_sample) <= mask, i.e. the first flat index
    # that is deeper than exclude_above (0 when every sample is shallower: argmin of an all-True array)
    m0 = kernels.first_not_le(rng, exclude_above)
    m0 = 0 if m0 is None else m0
    if m0 >= R:  # slice(min_range_sample, None) is empty: nothing is pooled, nothing is masked
        m0 = R
    pool = kernels.transient_noise_mask_median if median else kernels.transient_noise_mask
    mask, _ = pool(Sv, nsamp, C, P, R, m0, int(num_side_pings), thr)
    return DataArray(mask, DIMS, coords=_mask_coords(ds_Sv), name="Sv")


def mask_attenuated_signal(ds_Sv, upper_limit_sl: str = "400.0m", lower_limit_sl: str = "500.0m", num_side_pings: int = 15,
                           attenuation_signal_threshold: str = "8.0dB", range_var: str = "depth") -> DataArray:
    """
    Locate attenuated signals and create an attenuated-signal mask (Ryan et al. 2015; arguments as
    echopype.clean.mask_attenuated_signal, clean/api.py:269-359).

    Per channel and ping the median Sv (linear domain) of the scattering layer between ``upper_limit_sl`` and
    ``lower_limit_sl`` - the samples from the one nearest to the upper limit up to, not including, the one nearest to the
    lower limit - is compared with the median over the ``num_side_pings`` pings before and the ``num_side_pings - 1``
    pings after it; the WHOLE ping is masked when it lies less than ``attenuation_signal_threshold`` above the block
    median (so with the default +8 dB every ping that can be assessed and is not 8 dB above its neighbours is masked, as
    in the reference).  Pings within ``num_side_pings`` of either end, and pings without valid samples in the layer, are
    not masked.  Returns a boolean (uint8 on the device) DataArray with dims (channel, ping_time, range_sample); all
    False when the limits lie outside the extent of ``range_var``.
    """
    ds_Sv = as_dataset(ds_Sv)
    if range_var not in ["echo_range", "depth"]:
        raise ValueError("`range_var` must be either `echo_range` or `depth`.")
    if range_var not in ds_Sv:
        raise ValueError(f"Masking attenuated signal requires `{range_var}` data variable in `ds_Sv`.")
    # clean/api.py:318: the reference compares the two limit STRINGS here, before parsing them; kept as it is
    if upper_limit_sl > lower_limit_sl:
        raise ValueError("Minimum range has to be shorter than maximum range")
    thr = extract_dB(attenuation_signal_threshold)
    lower = _parse_x_bin(lower_limit_sl, "range_bin")
    upper = _parse_x_bin(upper_limit_sl, "range_bin")
    if not (isinstance(num_side_pings, (int, np.integer)) and num_side_pings >= 0):
        raise ValueError("num_side_pings must be a non-negative integer")
    Sv, rng, C, P, R = _index_binning_inputs(ds_Sv, range_var, "attenuated signal")
    lo, hi, _ = kernels.minmax(rng)
    if (upper > hi) or (lower < lo):  # searching range outside the echosounder range: nothing is masked
        return DataArray(torch.zeros((C, P, R), dtype=torch.uint8, device=Sv.device), DIMS, coords=_mask_coords(ds_Sv), name=range_var)
    mask, _ = kernels.attenuated_signal_mask(Sv, rng, C, P, R, upper, lower, int(num_side_pings), thr)
    return DataArray(mask, DIMS, coords=_mask_coords(ds_Sv), name="Sv")
