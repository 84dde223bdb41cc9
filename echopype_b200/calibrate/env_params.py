"""Host-side environmental-parameter assembly.  Same names, argument meaning and errors as
echopype/calibrate/env_params.py (harmonize_env_param_time :24-71, sanitize_user_env_dict :74-157,
get_env_params_AZFP :160-221, get_env_params_EK :224-353) and echopype/utils/align.py:9-61."""

from typing import Dict, List, Optional, Union

import numpy as np

from ..dataset import DataArray, Dataset
from ..utils import uwa
from .cal_params import _chan_values, param2da

ENV_PARAMS = (
    "sound_speed", "sound_absorption", "temperature", "salinity", "pressure", "pH",
    "formula_sound_speed", "formula_absorption",
)


def _as_ns(t):
    t = np.asarray(getattr(t, "values", t))
    if t.dtype.kind == "M":
        return t.astype("datetime64[ns]").astype(np.int64)
    return t.astype(np.int64)


def _interp_extrap(x, xp, fp):
    out = np.interp(x, xp, fp)
    if xp.size >= 2:
        lo, hi = x < xp[0], x > xp[-1]
        if lo.any():
            out[lo] = fp[0] + (x[lo] - xp[0]) * (fp[1] - fp[0]) / (xp[1] - xp[0])
        if hi.any():
            out[hi] = fp[-1] + (x[hi] - xp[-1]) * (fp[-1] - fp[-2]) / (xp[-1] - xp[-2])
    return out


def align_to_ping_time(external_da: DataArray, external_time_name: str, ping_time_da, method: str = "nearest") -> DataArray:
    """utils/align.py:9-61 for a DataArray with dims (time,) or (channel, time)."""
    pt_vals = np.asarray(getattr(ping_time_da, "values", ping_time_da))
    t_ext = np.asarray(external_da.coords[external_time_name])
    ax = external_da.dims.index(external_time_name)
    v = np.asarray(external_da.values, dtype=np.float64)
    new_dims = tuple("ping_time" if d == external_time_name else d for d in external_da.dims)
    coords = {k: c for k, c in external_da.coords.items() if k != external_time_name}
    coords["ping_time"] = pt_vals
    if t_ext.shape == pt_vals.shape and np.array_equal(_as_ns(t_ext), _as_ns(pt_vals)):
        return DataArray(v, new_dims, coords, external_da.attrs)
    if t_ext.size == 1:
        shape = list(v.shape)
        shape[ax] = pt_vals.size
        return DataArray(np.broadcast_to(v, shape).astype(np.float64).copy(), new_dims, coords, external_da.attrs)
    if t_ext.size == 0:
        shape = list(v.shape)
        shape[ax] = pt_vals.size
        return DataArray(np.full(shape, np.nan), new_dims, coords, external_da.attrs)
    t0 = _as_ns(t_ext)[0]
    x = (_as_ns(pt_vals) - t0).astype(np.float64)
    xp = (_as_ns(t_ext) - t0).astype(np.float64)
    vm = np.moveaxis(v, ax, -1)
    flat = vm.reshape(-1, vm.shape[-1])
    if method == "linear":
        res = np.stack([_interp_extrap(x, xp, row) for row in flat])
    elif method == "nearest":
        idx = np.clip(np.searchsorted(xp, x), 1, xp.size - 1)
        idx = np.where(np.abs(x - xp[idx - 1]) <= np.abs(xp[idx] - x), idx - 1, idx)
        res = flat[:, idx]
    else:
        raise ValueError(f"unsupported interpolation method {method!r}")
    res = np.moveaxis(res.reshape(vm.shape[:-1] + (x.size,)), -1, ax)
    return DataArray(res, new_dims, coords, external_da.attrs)


def harmonize_env_param_time(p, ping_time=None):
    if isinstance(p, DataArray):
        if "time1" not in p.coords and "time1" not in p.dims:
            return p
        ax = p.dims.index("time1")
        t1 = np.asarray(p.coords.get("time1", np.arange(p.shape[ax])))
        v = np.asarray(p.values)
        other = {k: c for k, c in p.coords.items() if k != "time1"}
        dims_wo = tuple(d for d in p.dims if d != "time1")
        if t1.size == 1:
            return DataArray(np.take(v, 0, axis=ax), dims_wo, other, p.attrs, p.name)
        keep = ~np.isnan(np.moveaxis(np.asarray(v, dtype=np.float64), ax, 0).reshape(t1.size, -1)).any(axis=1)
        if int(keep.sum()) == 1 and v.ndim == 1:
            return DataArray(v[keep][0], (), other, p.attrs, p.name)
        if ping_time is None:
            raise ValueError(f"ping_time needs to be provided for comparison or interpolating {p.name}")
        allnan = np.isnan(np.moveaxis(np.asarray(v, dtype=np.float64), ax, 0).reshape(t1.size, -1)).all(axis=1)
        pv = DataArray(np.compress(~allnan, v, axis=ax), p.dims, {**other, "time1": t1[~allnan]}, p.attrs, p.name)
        return align_to_ping_time(pv, "time1", ping_time, method="linear")
    return p


def sanitize_user_env_dict(user_dict: Dict, channel: Union[List, DataArray]) -> Dict:
    if not isinstance(channel, (list, DataArray)):
        raise ValueError("'channel' has to be a list or an xr.DataArray")
    channel_sorted = sorted(_chan_values(channel))
    out_dict = dict.fromkeys(ENV_PARAMS)
    for p_name, p_val in (user_dict or {}).items():
        if p_name not in out_dict:
            continue
        if p_name == "sound_absorption" and not isinstance(p_val, (DataArray, list)):
            raise ValueError(
                "The 'sound_absorption' parameter has to be a list or an xr.DataArray, with 'channel' as an coordinate."
            )
        if isinstance(p_val, DataArray):
            if "channel" in p_val.coords:
                if sorted(p_val.coords["channel"].tolist()) != channel_sorted:
                    raise ValueError(f"The 'channel' coordinate of {p_name} has to match that of the data to be calibrated")
            else:
                raise ValueError(f"{p_name} has to have 'channel' as a coordinate")
            out_dict[p_name] = p_val
        elif isinstance(p_val, (int, float, str)):
            out_dict[p_name] = p_val
        elif isinstance(p_val, list):
            out_dict[p_name] = param2da(p_val, channel)
        else:
            raise ValueError(f"{p_name} has to be a scalar, list, or an xr.DataArray")
    return out_dict


def _vals(x):
    return x.values if isinstance(x, DataArray) else x


def _wrap_like(res, *sources):
    """Give a numpy result the dims of the highest-rank DataArray among its inputs."""
    das = [s for s in sources if isinstance(s, DataArray) and s.ndim == np.ndim(res)]
    if das and np.ndim(res):
        return DataArray(res, das[0].dims, das[0].coords)
    return res


def _calc_absorption(frequency, T, S, P, pH=None, c=None, formula="FG"):
    """uwa.calc_absorption with dims-aware broadcasting of frequency (channel[, ping_time]) against T/S/P."""
    f, t, s, p, ph, cc = (x if isinstance(x, DataArray) else (None if x is None else DataArray(np.asarray(x, dtype=np.float64)))
                          for x in (frequency, T, S, P, pH, c))
    kw = {}
    if ph is not None:
        kw["pH"] = ph
    if cc is not None:
        kw["sound_speed"] = cc
    # The formulae are elementwise: inputs that do not vary along ping_time (the usual case -- one transmit setting
    # and one environment record per file) are evaluated on a single ping and the result repeated, instead of on
    # the full (channel, ping_time) grid.
    nP = 0
    ins = [x for x in (f, t, s, p, ph, cc) if x is not None and "ping_time" in x.dims]
    if ins and all(x.sizes["ping_time"] == ins[0].sizes["ping_time"] for x in ins) and ins[0].sizes["ping_time"] > 1:
        def const(x):
            v = np.moveaxis(x.values, x.dims.index("ping_time"), -1)
            return bool(np.all(v == v[..., :1]))
        if all(const(x) for x in ins):
            nP, pt = ins[0].sizes["ping_time"], ins[0].coords.get("ping_time")
            f, t, s, p, ph, cc = (x.isel(ping_time=slice(0, 1)) if x is not None and "ping_time" in x.dims else x
                                  for x in (f, t, s, p, ph, cc))
    # DataArray arithmetic broadcasts by dim name; np.sqrt / np.exp / np.all inside need plain arrays
    ref = f
    for other in (t, s, p, ph, cc):
        if other is not None:
            ref = ref + other * 0
    dims, coords = ref.dims, ref.coords

    def b(x):
        return None if x is None else (x + ref * 0).transpose(*dims).values if dims else (x + ref * 0).values

    out = uwa.calc_absorption(
        frequency=b(f), temperature=b(t), salinity=b(s), pressure=b(p),
        **({"pH": b(ph)} if ph is not None else {}), **({"sound_speed": b(cc)} if cc is not None else {}),
        formula_source=formula,
    )
    if nP:
        ax = dims.index("ping_time")
        out = np.repeat(out, nP, axis=ax)
        coords = dict(coords)
        if pt is not None:
            coords["ping_time"] = pt
    return DataArray(out, dims, coords)


def get_env_params_AZFP(echodata, user_dict: Optional[dict] = None):
    beam = echodata["Sonar/Beam_group1"]
    out_dict = sanitize_user_env_dict(user_dict=user_dict, channel=beam["channel"])
    out_dict.pop("pH")
    # the reference's membership test (env_params.py:184) is on a dict built with fromkeys, i.e. always
    # true; what actually fails downstream is a None salinity/pressure - raise the intended error here
    if out_dict.get("salinity") is None or out_dict.get("pressure") is None:
        raise ReferenceError("Please supply both salinity and pressure in env_params.")
    if out_dict["temperature"] is None:
        out_dict["temperature"] = echodata["Environment"]["temperature"]
    if out_dict["formula_sound_speed"] is None:
        out_dict["formula_sound_speed"] = "AZFP"
    if out_dict["formula_absorption"] is None:
        out_dict["formula_absorption"] = "AZFP"
    for p, v in out_dict.items():
        if v is None:
            if p == "sound_speed":
                T = out_dict["temperature"]
                res = uwa.calc_sound_speed(
                    temperature=_vals(T), salinity=_vals(out_dict["salinity"]), pressure=_vals(out_dict["pressure"]),
                    formula_source=out_dict["formula_sound_speed"],
                )
                out_dict[p] = _wrap_like(res, T, out_dict["salinity"], out_dict["pressure"])
            elif p == "sound_absorption":
                out_dict[p] = _calc_absorption(
                    beam["frequency_nominal"], out_dict["temperature"], out_dict["salinity"], out_dict["pressure"],
                    formula=out_dict["formula_absorption"],
                )
    for p in out_dict.keys():
        out_dict[p] = harmonize_env_param_time(out_dict[p], ping_time=beam["ping_time"])
    return out_dict


def get_env_params_EK(sonar_type, beam: Dataset, env: Dataset, user_dict: Optional[Dict] = None, freq: DataArray = None) -> Dict:
    if sonar_type not in ["EK60", "EK80"]:
        raise ValueError("'sonar_type' has to be 'EK60' or 'EK80'")
    if sonar_type == "EK80":
        if freq is None:
            raise ValueError("'freq' is required for calibrating EK80-style data.")
    else:
        freq = beam["frequency_nominal"]
    user_dict = user_dict or {}
    out_dict = sanitize_user_env_dict(user_dict=user_dict, channel=beam["channel"])
    if out_dict["formula_absorption"] not in [None, "AM", "FG"]:
        raise ValueError("'formula_absorption' has to be None, 'FG' or 'AM' for EK echosounders.")
    if out_dict["formula_sound_speed"] not in (None, "Mackenzie"):
        raise ValueError("'formula_absorption' has to be None or 'Mackenzie' for EK echosounders.")
    tspa_all_exist = all(out_dict[p] is not None for p in ["temperature", "salinity", "pressure", "pH"])
    if not tspa_all_exist and sonar_type == "EK80":
        for p_user, p_data in zip(["temperature", "salinity", "pressure", "pH"], ["temperature", "salinity", "depth", "acidity"]):
            out_dict[p_user] = user_dict.get(p_user, env[p_data])
    if out_dict["sound_speed"] is None:
        if not tspa_all_exist:
            out_dict["sound_speed"] = env["sound_speed_indicative"]
            out_dict.pop("formula_sound_speed")
        else:
            if out_dict["formula_sound_speed"] is None:
                out_dict["formula_sound_speed"] = "Mackenzie"
            res = uwa.calc_sound_speed(
                temperature=_vals(out_dict["temperature"]), salinity=_vals(out_dict["salinity"]),
                pressure=_vals(out_dict["pressure"]), formula_source=out_dict["formula_sound_speed"],
            )
            out_dict["sound_speed"] = _wrap_like(res, out_dict["temperature"], out_dict["salinity"], out_dict["pressure"])
    else:
        out_dict.pop("formula_sound_speed")
    if out_dict["sound_absorption"] is None:
        if not tspa_all_exist and sonar_type != "EK80":
            out_dict["sound_absorption"] = env["absorption_indicative"]
            out_dict.pop("formula_absorption")
        else:
            if out_dict["formula_absorption"] is None:
                out_dict["formula_absorption"] = "FG"
            out_dict["sound_absorption"] = _calc_absorption(
                freq, out_dict["temperature"], out_dict["salinity"], out_dict["pressure"], out_dict["pH"],
                out_dict["sound_speed"], out_dict["formula_absorption"],
            )
    else:
        out_dict.pop("formula_absorption")
    if not ("formula_sound_speed" in out_dict or "formula_absorption" in out_dict):
        for p in ["temperature", "salinity", "pressure", "pH"]:
            out_dict.pop(p)
    for p in out_dict.keys():
        out_dict[p] = harmonize_env_param_time(out_dict[p], ping_time=beam["ping_time"])
    return out_dict
