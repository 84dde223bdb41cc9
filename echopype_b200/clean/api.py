"""Background-noise estimate / removal (De Robertis & Higginbottom 2007) with the reference's signatures
(echopype/clean/api.py:362-433 estimate_background_noise, :436-511 remove_background_noise); the array work
runs in two kernels of libepb200 (epb_noise_estimate: tile means + block min, epb_noise_apply: subtraction,
SNR gate and the actual_range extrema) - 16 B read + 8 B written per sample."""

import numpy as np

from .. import kernels
from ..dataset import DataArray, Dataset, as_dataset
from ..device import ParamPack, require_cuda, to_device_f32
from ..utils.prov import add_processing_level, echopype_prov_attrs, insert_input_processing_level
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
    out["Sv_corrected"] = DataArray(sc, DIMS, attrs=noise_attrs("corrected", lo_c, hi_c, ping_num, range_sample_num, snr, background_noise_max))
    prov_dict = echopype_prov_attrs(process_type="processing")
    prov_dict["processing_function"] = "clean.remove_background_noise"
    out = out.assign_attrs(prov_dict)
    out = insert_input_processing_level(out, input_ds=ds_Sv)
    return out
