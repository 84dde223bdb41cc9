"""Public calibration API: same signatures, argument checks, messages and output attributes as
echopype/calibrate/api.py (CALIBRATOR :11-18, _compute_cal :23-246, compute_Sv :249, compute_TS :348)."""

import numpy as np

from ..dataset import Dataset, EchoData
from ..utils.log import _init_logger
from ..utils.prov import echopype_prov_attrs, source_files_vars
from .calibrate_azfp import CalibrateAZFP
from .calibrate_ek import CalibrateEK60, CalibrateEK80

CALIBRATOR = {
    "EK60": CalibrateEK60, "EK80": CalibrateEK80, "AZFP": CalibrateAZFP,
    "ES70": CalibrateEK60, "ES80": CalibrateEK80, "EA640": CalibrateEK80,
}

logger = _init_logger(__name__)


def check_input_args_combination(waveform_mode: str, encode_mode: str, pulse_compression: bool = None) -> None:
    """echodata/simrad.py:12-55."""
    if waveform_mode not in ["CW", "BB"]:
        raise ValueError("The input waveform_mode must be either 'CW' or 'BB'!")
    if encode_mode not in ["complex", "power"]:
        raise ValueError("The input encode_mode must be either 'complex' or 'power'!")
    if (waveform_mode == "BB") and (encode_mode == "power"):
        raise ValueError("Data from broadband ('BB') transmission must be recorded as complex samples")
    if pulse_compression is not None:
        if pulse_compression and ((waveform_mode != "BB") or (encode_mode != "complex")):
            raise RuntimeError("Pulse compression can only be used with waveform_mode='BB' and encode_mode='complex'")


def _compute_cal(cal_type, echodata: EchoData, env_params=None, cal_params=None, ecs_file=None, waveform_mode=None,
                 encode_mode=None, assume_single_filter_time=None, drop_last_hanning_zero=False):
    waveform_mode = "BB" if waveform_mode == "FM" else waveform_mode
    if echodata.sonar_model == "EK80":
        if waveform_mode is None or encode_mode is None:
            raise ValueError("waveform_mode and encode_mode must be specified for EK80 calibration")
        check_input_args_combination(waveform_mode=waveform_mode, encode_mode=encode_mode)
    elif echodata.sonar_model in ("EK60", "AZFP"):
        if waveform_mode is not None and waveform_mode != "CW":
            logger.warning(
                "This sonar model transmits only narrowband signals (waveform_mode='CW'). Calibration will be in CW mode",
            )
        if encode_mode is not None and encode_mode != "power":
            logger.warning(
                "This sonar model only record data as power or power/angle samples "
                "(encode_mode='power'). Calibration will be done on the power samples.",
            )
    if (echodata.sonar_model != "EK80" or encode_mode != "complex") and assume_single_filter_time is not None:
        raise ValueError("assume_single_filter_time can only be used on complex EK80 data.")
    if echodata.sonar_model not in CALIBRATOR:
        raise ValueError(f"Unsupported sonar_model {echodata.sonar_model!r}")

    def _compute_cal_ds(ed, slice_dict):
        cal_obj = CALIBRATOR[ed.sonar_model](
            ed, env_params=env_params, cal_params=cal_params, ecs_file=ecs_file, waveform_mode=waveform_mode,
            encode_mode=encode_mode, drop_last_hanning_zero=drop_last_hanning_zero, slice_dict=slice_dict,
        )
        cal_obj._check_echodata_backscatter_size()
        return cal_obj.compute_Sv() if cal_type == "Sv" else cal_obj.compute_TS()

    vend = echodata["Vendor_specific"] if echodata.sonar_model in ["EK80", "ES80", "EA640"] else None
    if vend is None or "filter_time" not in vend.sizes or vend.sizes["filter_time"] == 1:
        cal_ds = _compute_cal_ds(echodata, {})
    else:  # several filter sets, calibrate/api.py:101-197
        from . import filter_time as ft
        from .calibrate_ek import retrieve_correct_beam_group

        beam_group = retrieve_correct_beam_group(echodata, waveform_mode, encode_mode)
        beam = echodata[beam_group]
        if assume_single_filter_time:
            first_valid = ft.first_valid_filter_time_per_channel(beam)
            ed = ft._with_groups(echodata)
            ed["Vendor_specific"] = ft.collapse_vend(vend, first_valid)
            cal_ds = _compute_cal_ds(ed, {"first_valid_filter_time_per_channel": first_valid})
        else:
            pieces = [
                _compute_cal_ds(ft.piece_echodata(echodata, beam_group, ci, p_idx, fi), {"channel": ci, "filter_time": fi})
                for ci, p_idx, fi in ft.filter_pieces(beam, vend)
            ]
            cal_ds = ft.merge_pieces(pieces)

    # attributes, calibrate/api.py:200-219
    cal_ds["range_sample"].attrs.update({"long_name": "Along-range sample number, base 0"})
    cal_ds["echo_range"].attrs.update({"long_name": "Range distance", "units": "m"})
    cal_ds[cal_type].attrs.update(
        {"long_name": {"Sv": "Volume backscattering strength (Sv re 1 m-1)", "TS": "Target strength (TS re 1 m^2)"}[cal_type], "units": "dB"}
    )
    if echodata.sonar_model == "EK80":
        cal_ds[cal_type].attrs.update({"waveform_mode": waveform_mode, "encode_mode": encode_mode})

    # provenance, calibrate/api.py:221-241
    if echodata.source_file is not None:
        source_file = echodata.source_file
    elif echodata.converted_raw_path is not None:
        source_file = echodata.converted_raw_path
    else:
        source_file = "SOURCE FILE NOT IDENTIFIED"
    prov = echopype_prov_attrs(process_type="processing")
    prov["processing_function"] = f"calibrate.compute_{cal_type}"
    fv = source_files_vars(source_file)
    cal_ds._set_coord("filenames", fv["source_files_coord"]["filenames"])
    cal_ds["source_filenames"] = fv["source_files_var"]["source_filenames"]
    cal_ds.attrs.update(prov)
    if "water_level" in echodata["Platform"]:
        cal_ds["water_level"] = echodata["Platform"]["water_level"]
    return cal_ds


def compute_Sv(echodata: EchoData, **kwargs) -> Dataset:
    """Compute volume backscattering strength (Sv) from raw data.

    Keyword arguments as in the reference: env_params, cal_params, ecs_file, waveform_mode
    ({"CW","BB","FM"}), encode_mode ({"complex","power"}), assume_single_filter_time,
    drop_last_hanning_zero.  Returns a Dataset with ``Sv``, ``echo_range`` (device resident,
    float32; read ``.values`` for a host copy), ``tau_effective``, ``frequency_nominal`` and every
    environmental / calibration parameter used.
    """
    return _compute_cal(cal_type="Sv", echodata=echodata, **kwargs)


def compute_TS(echodata: EchoData, **kwargs):
    """Compute target strength (TS) from raw data (arguments as :func:`compute_Sv`)."""
    return _compute_cal(cal_type="TS", echodata=echodata, **kwargs)
