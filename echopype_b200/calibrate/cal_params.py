"""Host-side calibration-parameter assembly (tiny (channel[, ping_time]) arrays; stays on the CPU and feeds
the per-row coefficient table built on the device).  Same names, argument meaning and errors as
echopype/calibrate/cal_params.py (CAL_PARAMS :6-33, param2da :53-85, sanitize_user_cal_dict :88-166,
_get_interp_da :169-258, get_vend_cal_params_power :261-324, get_cal_params_AZFP :327-362,
get_cal_params_EK :365-522)."""

from typing import Dict, List, Union

import numpy as np

from ..dataset import DataArray, Dataset

CAL_PARAMS = {
    "EK60": (
        "sa_correction", "gain_correction", "equivalent_beam_angle", "angle_offset_alongship",
        "angle_offset_athwartship", "angle_sensitivity_alongship", "angle_sensitivity_athwartship",
        "beamwidth_alongship", "beamwidth_athwartship",
    ),
    "EK80": (
        "sa_correction", "gain_correction", "equivalent_beam_angle", "angle_offset_alongship",
        "angle_offset_athwartship", "angle_sensitivity_alongship", "angle_sensitivity_athwartship",
        "beamwidth_alongship", "beamwidth_athwartship", "impedance_transducer", "impedance_transceiver",
        "receiver_sampling_frequency",
    ),
    "AZFP": ("EL", "DS", "TVR", "VTX0", "equivalent_beam_angle", "Sv_offset"),
}

EK80_DEFAULT_PARAMS = {
    "impedance_transducer": 75,
    "impedance_transceiver": 1000,
    "receiver_sampling_frequency": {
        "default": 1500000, "GPT": 500000, "SBT": 50000, "WBAT": 1500000, "WBT TUBE": 1500000,
        "WBT MINI": 1500000, "WBT": 1500000, "WBT HP": 187500, "WBT LF": 93750,
    },
}


def _chan_values(channel):
    return list(channel.values) if isinstance(channel, DataArray) else list(channel)


def param2da(p_val: Union[int, float, list], channel: Union[list, DataArray]) -> DataArray:
    if not isinstance(p_val, (int, float, list)):
        raise ValueError("'p_val' needs to be one of type int, float, or list")
    ch = _chan_values(channel)
    if isinstance(p_val, list):
        if len(p_val) != len(ch):
            raise ValueError("The lengths of 'p_val' and 'channel' should be identical")
        return DataArray(np.asarray(p_val), dims=["channel"], coords={"channel": np.asarray(ch)})
    return DataArray(np.asarray([p_val] * len(ch)), dims=["channel"], coords={"channel": np.asarray(ch)})


def sanitize_user_cal_dict(sonar_type, user_dict: Dict, channel: Union[List, DataArray]) -> Dict:
    if sonar_type not in ["EK60", "EK80", "AZFP"]:
        raise ValueError("'sonar_type' has to be one of: 'EK60', 'EK80', or 'AZFP'")
    if not isinstance(channel, (list, DataArray)):
        raise ValueError("'channel' has to be a list or an xr.DataArray")
    channel_sorted = sorted(_chan_values(channel))
    out_dict = dict.fromkeys(CAL_PARAMS[sonar_type])
    for p_name, p_val in user_dict.items():
        if p_name not in out_dict:
            continue
        if isinstance(p_val, DataArray):
            if "channel" in p_val.coords:
                if sorted(p_val.coords["channel"].tolist()) != channel_sorted:
                    raise ValueError(f"The 'channel' coordinate of {p_name} has to match that of the data to be calibrated")
            elif "cal_channel_id" in p_val.coords and "cal_frequency" in p_val.coords:
                if sorted(p_val.coords["cal_channel_id"].tolist()) != channel_sorted:
                    raise ValueError(f"The 'cal_channel_id' coordinate of {p_name} has to match that of the data to be calibrated")
            else:
                raise ValueError(
                    f"{p_name} has to either have 'channel' as a coordinate "
                    "or have both 'cal_channel_id' and 'cal_frequency' as coordinates"
                )
            out_dict[p_name] = p_val
        elif isinstance(p_val, (int, float, list)):
            out_dict[p_name] = param2da(p_val, channel)
        else:
            raise ValueError(f"{p_name} has to be a scalar, list, or an xr.DataArray")
    return out_dict


def _get_interp_da(da_param, freq_center: DataArray, alternative, BB_factor=1) -> DataArray:
    """Interpolate a frequency-dependent parameter (dims cal_channel_id, cal_frequency) at freq_center,
    channel by channel; channels without such values use ``alternative`` (cal_params.py:169-258)."""
    channels = freq_center.coords["channel"]
    has_ping = "ping_time" in freq_center.dims
    nP = freq_center.sizes.get("ping_time", 1)
    rows = []
    for ci, ch_id in enumerate(channels):
        fc = np.asarray(freq_center.values[ci], dtype=np.float64).reshape(-1)
        if da_param is not None and "cal_channel_id" in da_param.coords and ch_id in da_param.coords["cal_channel_id"]:
            k = int(np.flatnonzero(da_param.coords["cal_channel_id"] == ch_id)[0])
            tbl = np.asarray(da_param.values, dtype=np.float64)
            tbl = tbl[k] if da_param.dims[0] == "cal_channel_id" else tbl[:, k]
            cf = np.asarray(da_param.coords["cal_frequency"], dtype=np.float64)
            cf = cf[k] if cf.ndim == 2 else cf
            ok = ~np.isnan(tbl)
            if fc.size > 1 and np.all(fc == fc[0]):  # one transmit setting for the whole file: interpolate once
                v = np.full(fc.size, np.interp(fc[:1], cf[ok], tbl[ok], left=np.nan, right=np.nan)[0])
            else:
                v = np.interp(fc, cf[ok], tbl[ok], left=np.nan, right=np.nan)  # xarray interp: NaN outside
            rows.append(v if has_ping else v.reshape(()))
        else:
            bb = BB_factor.values[ci] if isinstance(BB_factor, DataArray) else BB_factor
            if isinstance(alternative, DataArray):
                alt = np.asarray(alternative.sel(channel=ch_id).values * bb, dtype=np.float64).squeeze()
            elif isinstance(alternative, (int, float)):
                alt = np.full(fc.size, alternative, dtype=np.float64).squeeze() * bb
            else:
                raise ValueError("'alternative' has to be of the type int, float, or xr.DataArray")
            alt = np.asarray(alt, dtype=np.float64)
            if alt.size == 1 and has_ping:
                alt = np.full(nP, float(alt.reshape(-1)[0]))
            rows.append(alt)
    param = np.array(rows)
    if has_ping:
        if param.ndim == 1:
            param = param[:, None]
        return DataArray(param, dims=["channel", "ping_time"], coords={"channel": channels, "ping_time": freq_center.coords.get("ping_time", np.arange(nP))})
    return DataArray(param, dims=["channel"], coords={"channel": channels})


def get_vend_cal_params_power(beam: Dataset, vend: Dataset, param: str) -> DataArray:
    """Match transmit_duration_nominal with the allowable pulse_length table: argmin_k |tau - pl[c,k]|
    (first minimum), NaN tau -> NaN; vend rows are matched to beam channels BY NAME."""
    if param not in ["sa_correction", "gain_correction"]:
        raise ValueError(f"Unknown parameter {param}")
    if param not in vend:
        raise ValueError(f"{param} does not exist in the Vendor_specific group!")
    tau_da = beam["transmit_duration_nominal"]
    tau = np.asarray(tau_da.transpose("channel", "ping_time").values, dtype=np.float64)
    bch = np.asarray(beam["channel"].values)
    vch = np.asarray(vend["channel"].values)
    order = [int(np.flatnonzero(vch == c)[0]) for c in bch]
    pl = np.asarray(vend["pulse_length"].transpose("channel", "pulse_length_bin").values, dtype=np.float64)[order]
    tb = np.asarray(vend[param].transpose("channel", "pulse_length_bin").values, dtype=np.float64)[order]
    isnull = np.isnan(tau)
    if tau.shape[1] > 1 and bool(((tau == tau[:, :1]) | (isnull & isnull[:, :1])).all()):
        tau_u, collapse = tau[:, :1], True  # constant per channel: O(C) lookup, broadcast afterwards
    else:
        tau_u, collapse = tau, False
    d = np.abs(tau_u[:, :, None] - pl[:, None, :])
    idx = np.argmin(np.where(np.isnan(d), np.inf, d), axis=2)
    val = np.take_along_axis(tb, idx, axis=1)
    val = np.where(np.isnan(tau_u), np.nan, val)
    if collapse:
        val = np.broadcast_to(val, tau.shape)
    out = DataArray(val, dims=["channel", "ping_time"], coords={"channel": bch, "ping_time": tau_da.coords.get("ping_time", np.arange(tau.shape[1]))}, name=param)
    if tau_da.dims == ("ping_time", "channel"):
        out = out.transpose("ping_time", "channel")
    return out


def get_cal_params_AZFP(beam: Dataset, vend: Dataset, user_dict: dict) -> dict:
    out_dict = sanitize_user_cal_dict(user_dict=user_dict, channel=beam["channel"], sonar_type="AZFP")
    for p, v in out_dict.items():
        if v is None:
            if p == "equivalent_beam_angle":
                out_dict[p] = beam[p]
            elif p in ["EL", "DS", "TVR", "VTX0", "Sv_offset"]:
                out_dict[p] = vend[p]
    return out_dict


def get_cal_params_EK(waveform_mode, freq_center: DataArray, beam: Dataset, vend: Dataset, user_dict: Dict,
                      default_params: Dict = EK80_DEFAULT_PARAMS, sonar_type: str = "EK80") -> Dict:
    if not isinstance(waveform_mode, str):
        raise TypeError("waveform_mode is not type string")
    elif waveform_mode not in ["CW", "BB"]:
        raise ValueError("waveform_mode must be 'CW' or 'BB'")

    def _get_fs():
        if "receiver_sampling_frequency" in vend and not np.isclose(vend["receiver_sampling_frequency"].values, 0).all():
            return vend["receiver_sampling_frequency"]
        fs = []
        for ci in range(len(vend["channel"])):
            tcvr_type = str(vend["transceiver_type"].values[ci]).upper()
            fs.append(default_params["receiver_sampling_frequency"][tcvr_type])
        return DataArray(np.asarray(fs), dims=["channel"], coords={"channel": vend["channel"].values})

    PARAM_BEAM_NAME_MAP = {
        "angle_offset_alongship": "angle_offset_alongship",
        "angle_offset_athwartship": "angle_offset_athwartship",
        "angle_sensitivity_alongship": "angle_sensitivity_alongship",
        "angle_sensitivity_athwartship": "angle_sensitivity_athwartship",
        "beamwidth_alongship": "beamwidth_twoway_alongship",
        "beamwidth_athwartship": "beamwidth_twoway_athwartship",
        "equivalent_beam_angle": "equivalent_beam_angle",
    }
    if waveform_mode == "BB":
        PARAM_BEAM_NAME_MAP.pop("equivalent_beam_angle")

    out_dict = sanitize_user_cal_dict(user_dict=user_dict, channel=beam["channel"], sonar_type=sonar_type)
    for p, v in out_dict.items():
        if v is not None and "cal_channel_id" in v.coords:
            out_dict[p] = _get_interp_da(v, freq_center, np.nan)

    for p, v in out_dict.items():
        if v is not None:
            continue
        if p == "sa_correction":
            out_dict[p] = get_vend_cal_params_power(beam=beam, vend=vend, param=p)
        elif p == "impedance_transceiver":
            out_dict[p] = default_params[p] if p not in vend else vend["impedance_transceiver"]
        elif p == "receiver_sampling_frequency":
            out_dict[p] = _get_fs()
        elif waveform_mode == "CW":
            if p in PARAM_BEAM_NAME_MAP:
                if PARAM_BEAM_NAME_MAP[p] in beam:
                    out_dict[p] = beam[PARAM_BEAM_NAME_MAP[p]]
            elif p == "gain_correction":
                out_dict[p] = get_vend_cal_params_power(beam=beam, vend=vend, param=p)
            elif p == "impedance_transducer":
                out_dict[p] = _get_interp_da(None if p not in vend else vend[p], freq_center, default_params[p])
            else:
                raise ValueError(f"{p} not in the defined set of calibration parameters.")
        else:  # BB
            if p in PARAM_BEAM_NAME_MAP:
                if p in ["angle_sensitivity_alongship", "angle_sensitivity_athwartship"]:
                    BB_factor = freq_center / beam["frequency_nominal"]
                elif p in ["beamwidth_alongship", "beamwidth_athwartship"]:
                    BB_factor = beam["frequency_nominal"] / freq_center
                else:
                    BB_factor = 1
                if PARAM_BEAM_NAME_MAP[p] in beam:
                    out_dict[p] = _get_interp_da(None if p not in vend else vend[p], freq_center, beam[PARAM_BEAM_NAME_MAP[p]], BB_factor)
            elif p == "equivalent_beam_angle":
                ratio = beam["frequency_nominal"] / freq_center
                out_dict[p] = beam[p] + DataArray(20 * np.log10(ratio.values), ratio.dims, ratio.coords)
            elif p == "gain_correction":
                out_dict[p] = _get_interp_da(
                    None if "gain" not in vend else vend["gain"], freq_center,
                    get_vend_cal_params_power(beam=beam, vend=vend, param=p),
                )
            elif p == "impedance_transducer":
                out_dict[p] = _get_interp_da(None if p not in vend else vend[p], freq_center, default_params[p])
            else:
                raise ValueError(f"{p} not in the defined set of calibration parameters.")
    return out_dict
