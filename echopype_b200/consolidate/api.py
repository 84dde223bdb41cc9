"""consolidate.add_depth (echopype/consolidate/api.py:66-247, ek_depth_utils.py, utils/align.py:5-61) with the
O(channel x ping x range) affine transform ``depth = transducer_depth[p] + orientation * echo_range * scaling`` on the
device (epb_add_depth).  When ``echo_range`` came from this package's compute_Sv, the new ``depth`` variable carries
the exact range law plus the per-ping (offset, scale), so compute_MVBS(range_var="depth") stays on the index-space
binning path (bit-identical bin membership with the float64 reference)."""

import datetime
from numbers import Number
from typing import Optional, Union

import numpy as np

from .. import kernels
from ..dataset import DataArray, Dataset, EchoData, as_dataset
from ..device import require_cuda, to_device_f32
from ..utils.log import _init_logger
from ..utils.prov import add_processing_level

logger = _init_logger(__name__)


def _ns(t):
    return np.asarray(t).astype("datetime64[ns]").astype(np.int64)


def align_to_ping_time(values, times, ping_time, method: str = "nearest") -> np.ndarray:
    """utils/align.py:5-61 for a 1-D series: identical times -> as is; one value -> broadcast; empty -> NaN;
    otherwise interp1d(kind=method, fill_value="extrapolate") ("nearest": ties go to the earlier sample)."""
    v = np.asarray(values, dtype=np.float64)
    t, p = _ns(times), _ns(ping_time)
    if t.shape == p.shape and np.array_equal(t, p):
        return v.copy()
    if v.size == 1:
        return v.reshape(-1)[0] * np.ones(len(p), dtype=np.float64)
    if v.size == 0:
        return np.full(len(p), np.nan)
    if method == "nearest":
        mid = (t[:-1].astype(np.float64) + t[1:].astype(np.float64)) / 2
        return v[np.searchsorted(mid, p.astype(np.float64), side="left")]
    if method == "linear":
        tf, pf = t.astype(np.float64), p.astype(np.float64)
        out = np.interp(pf, tf, v)
        lo, hi = pf < tf[0], pf > tf[-1]  # linear extrapolation
        out[lo] = v[0] + (pf[lo] - tf[0]) * (v[1] - v[0]) / (tf[1] - tf[0])
        out[hi] = v[-1] + (pf[hi] - tf[-1]) * (v[-1] - v[-2]) / (tf[-1] - tf[-2])
        return out
    raise ValueError(f"unsupported interpolation method {method!r}")


def _check_and_log_nans(group, group_name, names):
    for n in names:
        if np.any(np.isnan(np.asarray(group[n].values, dtype=np.float64))):
            logger.warning(
                f"The Echodata `{group_name}` group `{n}` variable array contains "
                "NaNs. This will result in NaNs in the final `depth` array. Consider filling the "
                "NaNs and calling `.add_depth(...)` again."
            )


def ek_use_platform_vertical_offsets(platform_ds, ping_time):
    """ek_depth_utils.py:30-52."""
    _check_and_log_nans(platform_ds, "Platform", ["water_level", "vertical_offset", "transducer_offset_z"])
    td = (np.asarray(platform_ds["transducer_offset_z"].values, dtype=np.float64)
          - (np.asarray(platform_ds["water_level"].values, dtype=np.float64) + np.asarray(platform_ds["vertical_offset"].values, dtype=np.float64)))
    td = np.atleast_1d(td)
    return align_to_ping_time(td, platform_ds["time2"].values if td.size > 1 else np.asarray(ping_time)[:1], ping_time)


def ek_use_platform_angles(platform_ds, ping_time):
    """ek_depth_utils.py:55-75: element [2, 2] of the ZYX (yaw = 0, pitch, roll) rotation = cos(pitch) cos(roll)."""
    _check_and_log_nans(platform_ds, "Platform", ["pitch", "roll"])
    pitch = np.deg2rad(np.atleast_1d(np.asarray(platform_ds["pitch"].values, dtype=np.float64)))
    roll = np.deg2rad(np.atleast_1d(np.asarray(platform_ds["roll"].values, dtype=np.float64)))
    scaling = np.cos(pitch) * np.cos(roll)
    return align_to_ping_time(scaling, platform_ds["time2"].values if scaling.size > 1 else np.asarray(ping_time)[:1], ping_time)


def ek_use_beam_angles(beam_ds):
    """ek_depth_utils.py:78-120: normalised z component of the beam direction per channel (NaN for a zero vector)."""
    _check_and_log_nans(beam_ds, "Sonar/Beam_group1", ["beam_direction_x", "beam_direction_y", "beam_direction_z"])
    x, y, z = (np.asarray(beam_ds[f"beam_direction_{a}"].values, dtype=np.float64) for a in "xyz")
    norm = np.sqrt(x**2 + y**2 + z**2)
    tol = 1e-8
    if ((norm > tol) & (np.abs(norm - 1) > tol)).any():
        logger.warning("Beam direction vector was not normalized; applying normalization. By definition, it should have been normalized.")
    if (norm < tol).any():
        logger.warning("Some beam direction vectors are zero. Outputting NaN for those channels.")
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(norm < tol, np.nan, z / norm)


@add_processing_level("L2A")
def add_depth(
    ds,
    echodata: Optional[EchoData] = None,
    depth_offset: Optional[Union[Number, DataArray]] = None,
    tilt: Optional[Union[Number, DataArray]] = None,
    downward: bool = True,
    use_platform_vertical_offsets: bool = False,
    use_platform_angles: bool = False,
    use_beam_angles: bool = False,
) -> Dataset:
    """Create a ``depth`` variable from ``echo_range`` (arguments, errors and attrs as echopype.consolidate.add_depth)."""
    ds = as_dataset(ds)
    if (not echodata) and (use_platform_vertical_offsets or use_platform_angles or use_beam_angles):
        raise ValueError(
            "If any of `use_platform_vertical_offsets`, "
            + "`use_platform_angles` "
            + "or `use_beam_angles` is `True`, "
            + "then `echodata` cannot be `None`."
        )
    if use_platform_angles and use_beam_angles:
        raise NotImplementedError("Computing depth with both platform and beam angles is not implemented yet.")
    if depth_offset is not None and use_platform_vertical_offsets:
        logger.warning("When `depth_offset` is specified, platform vertical offset variables will not be used.")
    if tilt is not None and (use_beam_angles or use_platform_angles):
        logger.warning("When `tilt` is specified, beam/platform angle variables will not be used.")
    sonar_model = None
    if echodata:
        sonar_model = echodata.sonar_model
        if sonar_model not in ["EK60", "EK80"] and (use_platform_vertical_offsets or use_platform_angles or use_beam_angles):
            raise NotImplementedError(f"`use_platform/beam_...` not implemented yet for `{sonar_model}`.")

    pt = np.asarray(ds["ping_time"].values)
    P = len(pt)
    transducer_depth = 0.0
    if isinstance(depth_offset, Number):
        transducer_depth = depth_offset
    if isinstance(depth_offset, DataArray):
        if len(depth_offset.dims) != 1:
            raise ValueError("If depth_offset is passed in as an xr.DataArray, it must contain a single dimension.")
        transducer_depth = align_to_ping_time(depth_offset.values, depth_offset.coords[depth_offset.dims[0]], pt)
    elif echodata and sonar_model in ["EK60", "EK80"] and use_platform_vertical_offsets and depth_offset is None:
        transducer_depth = ek_use_platform_vertical_offsets(echodata["Platform"], pt)

    scaling = 1.0  # scalar, (P,) per ping, or (C,) per channel
    per_channel = False
    beam_group_name = "Beam_group1"
    if isinstance(tilt, Number):
        scaling = np.cos(np.deg2rad(tilt))
    if isinstance(tilt, DataArray):
        if len(tilt.dims) != 1:
            raise ValueError("If tilt is passed in as an xr.DataArray, it must contain a single dimension.")
        scaling = np.cos(np.deg2rad(align_to_ping_time(tilt.values, tilt.coords[tilt.dims[0]], pt)))
    elif echodata and sonar_model in ["EK60", "EK80"] and tilt is None:
        if use_platform_angles:
            scaling = ek_use_platform_angles(echodata["Platform"], pt)
        elif use_beam_angles:
            b1 = echodata["Sonar/Beam_group1"]
            same = np.array_equal(np.asarray(b1["channel"].values), np.asarray(ds["channel"].values))
            beam_group_name = "Beam_group1" if same else "Beam_group2"
            scaling = ek_use_beam_angles(echodata[f"Sonar/{beam_group_name}"])
            per_channel = True

    mult = 1 if downward else -1
    dev = require_cuda()
    er = ds["echo_range"]
    C, Pr, R = er.shape
    if Pr != P:
        raise ValueError("echo_range and ping_time lengths differ")
    off_p = np.broadcast_to(np.asarray(transducer_depth, dtype=np.float64), (P,)).copy()
    if per_channel:
        scale_cp = np.repeat((mult * np.asarray(scaling, dtype=np.float64))[:, None], P, axis=1)
    else:
        scale_cp = np.broadcast_to(mult * np.asarray(scaling, dtype=np.float64), (P,)).copy()
    rng_t = to_device_f32(er.data, dev)
    depth_t = kernels.add_depth(rng_t, off_p, scale_cp, C, P, R)
    da = DataArray(depth_t, er.dims, name="depth",
                   attrs={"long_name": "Depth", "standard_name": "depth", "units": "m"})
    law = getattr(er, "law", None)
    monotone = bool(np.all(np.isfinite(scale_cp)) and np.all(scale_cp > 0) and np.all(np.isfinite(off_p)))  # index-space binning needs an increasing law
    if law is not None and law.get("rows") is not None and law.get("kind") == "echo_range" and not per_channel and monotone:
        da.law = {"rows": law["rows"], "kind": "depth", "minmax": None,
                  "off": kernels.to_device_f64(off_p, dev), "scale": kernels.to_device_f64(scale_cp, dev)}
    ds = ds.copy()
    ds["depth"] = da

    # history attribute (consolidate/api.py:224-239)
    used_v = use_platform_vertical_offsets and not depth_offset
    used_a = use_platform_angles and not tilt
    used_b = use_beam_angles and not tilt
    hist = f"{datetime.datetime.now(datetime.UTC)}. `depth` calculated using:"
    hist += (
        " Sv `echo_range`"
        f"{', Echodata `Platform` Vertical Offsets' if used_v else ''}"
        f"{', Echodata `Platform` Angles' if used_a else ''}"
        f"{', Echodata `%s` Angles' % beam_group_name if used_b else ''}"
        "."
    )
    ds["depth"].attrs["history"] = hist
    return ds
