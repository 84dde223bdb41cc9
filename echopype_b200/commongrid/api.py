"""Common-grid reductions with the reference's signatures (echopype/commongrid/api.py:31-191 compute_MVBS,
:195-266 compute_MVBS_index_binning, :270-416 compute_NASC).  The (channel x ping x range) reduction that the
reference delegates to flox runs in libepb200 (epb_bin_reduce / epb_bin_reduce_law / epb_coarsen +
epb_bin_finalize); bin edges, ping -> bin assignment and the small output Dataset are assembled here."""

import logging
from typing import Literal

import numpy as np
import torch

from .. import kernels
from ..dataset import DataArray, Dataset, as_dataset
from ..device import require_cuda, to_device_f32
from ..utils.prov import add_processing_level, echopype_prov_attrs, insert_input_processing_level
from .utils import (
    POSITION_VARIABLES,
    _parse_x_bin,
    _setup_and_validate,
    assign_bins,
    binned_nanmean,
    get_distance_from_latlon,
    ping_time_bin_parsing_and_conversion,
    ping_time_edges,
    range_edges,
)

DIMS = ("channel", "ping_time", "range_sample")
_AGG_MSG = (
    "Aggregation may be negatively impacted since Flox will not aggregate any "
    "```Sv``` values that have corresponding NaN coordinate values. Consider handling "
    "these values before calling your intended commongrid function."
)


def _set_var_attrs(da, long_name, units, round_digits=None, standard_name=None):
    da.attrs.clear()
    da.attrs.update({"long_name": long_name, "units": units})
    if standard_name:
        da.attrs["standard_name"] = standard_name


def _set_MVBS_attrs(ds):
    ds["ping_time"].attrs.clear()
    ds["ping_time"].attrs.update({"long_name": "Ping time", "standard_name": "time", "axis": "T"})
    _set_var_attrs(ds["Sv"], "Mean volume backscattering strength (MVBS, mean Sv re 1 m-1)", "dB", 2)


def _first_dim(ds_Sv):
    """commongrid/utils.py:610-611: the grouping dimension is the first one (channel or frequency_nominal)."""
    d0 = ds_Sv["Sv"].dims[0]
    return d0


def _range_tensor(da, dev):
    """Range variable as a device tensor: float64 host arrays stay float64 (exact binning), everything else float32."""
    data = da.data
    if isinstance(data, torch.Tensor):
        t = data if data.dtype in (torch.float32, torch.float64) else data.float()
        return t.to(dev).contiguous()
    a = np.asarray(data)
    if a.dtype == np.float64:
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return to_device_f32(a, dev)


def _range_max_and_nan(da, rng_t):
    """nanmax of the range variable and whether it holds NaNs; cached extrema of compute_Sv are reused."""
    law = getattr(da, "law", None)
    if law is not None and law.get("minmax") is not None:
        mm = law["minmax"].tolist()
        hi = float("nan") if mm[3] == float("-inf") else mm[3]
        return hi, None
    if rng_t.dtype == torch.float32:
        _, hi, has_nan = kernels.minmax(rng_t)
        return hi, has_nan
    m = torch.isnan(rng_t)
    has_nan = bool(m.any())
    hi = float(torch.where(m, torch.full_like(rng_t, float("-inf")), rng_t).max().item())
    return (float("nan") if hi == float("-inf") else hi), has_nan


def _warn_nan_coords(x_name, x_vals, range_var, range_has_nan):
    xv = np.asarray(x_vals)
    x_nan = np.isnat(xv).any() if xv.dtype.kind == "M" else np.isnan(xv.astype(np.float64)).any()
    if x_nan:
        logging.warning(f"The ```{x_name}``` coordinate array contain NaNs. {_AGG_MSG}")
    if range_has_nan:
        logging.warning(f"The ```{range_var}``` coordinate array contain NaNs. {_AGG_MSG}")


def _reduce(ds_Sv, range_var, xbin_np, nX, r_edges_np, closed, skipna, fill_value, to_db, with_height=False):
    """Device reduction shared by MVBS and NASC.  Returns (mean[C,nX,nR] float64 numpy, heights or None)."""
    dev = require_cuda()
    sv = ds_Sv["Sv"]
    C, P, R = sv.shape
    Sv_t = to_device_f32(sv.data, dev)
    rda = ds_Sv[range_var]
    xbin = torch.from_numpy(np.ascontiguousarray(xbin_np, dtype=np.int32)).to(dev)
    edges = torch.from_numpy(np.ascontiguousarray(r_edges_np, dtype=np.float64)).to(dev)
    nR = len(r_edges_np) - 1
    acc = kernels.new_acc(C, nX, nR, dev)
    law = getattr(rda, "law", None)
    use_law = (
        law is not None and law.get("rows") is not None and law.get("kind") in ("echo_range", "depth") and skipna
        and np.isnan(fill_value) and nR <= 511 and not with_height and tuple(rda.shape) == (C, P, R)
    )
    if use_law:
        # depth = off[p] + echo_range * scale[p] (consolidate.add_depth): boundaries on the exact float64 law
        kernels.bin_reduce_law(Sv_t, law["rows"], xbin, edges, acc, C, P, R, nX, closed_right=(closed == "right"),
                               depth_off=law.get("off"), depth_scale=law.get("scale"))
    else:
        rng_t = _range_tensor(rda, dev)
        if tuple(rng_t.shape) != (C, P, R):
            raise ValueError(f"{range_var} must have the shape of Sv")
        kernels.bin_reduce(Sv_t, rng_t, xbin, edges, acc, C, P, R, nX, closed_right=(closed == "right"), with_height=with_height)
    out, h = kernels.bin_finalize(acc, skipna=skipna, fill_value=fill_value, to_db=to_db, want_height=with_height)
    return out.cpu().numpy().astype(np.float64), (h.cpu().numpy() if h is not None else None)


@add_processing_level("L3*")
def compute_MVBS(
    ds_Sv,
    range_var: Literal["echo_range", "depth"] = "echo_range",
    range_bin: str = "20m",
    ping_time_bin: str = "20s",
    method="map-reduce",
    reindex=False,
    skipna=True,
    fill_value=np.nan,
    closed: Literal["left", "right"] = "left",
    range_var_max: str = None,
    **flox_kwargs,
):
    """
    Compute Mean Volume Backscattering Strength (MVBS) based on intervals of range (``echo_range``) or depth
    (``depth``) and ``ping_time`` specified in physical units.  Arguments and output as
    echopype.commongrid.compute_MVBS; ``method`` / ``reindex`` / ``flox_kwargs`` are accepted for compatibility
    (the reduction is a single device pass).
    """
    if method != "map-reduce" and reindex is not None:
        raise ValueError(f"Passing in reindex={reindex} is only allowed when method='map_reduce'.")
    ds_Sv = as_dataset(ds_Sv)
    ds_Sv, range_bin = _setup_and_validate(ds_Sv, range_var, range_bin, closed)
    if not isinstance(ping_time_bin, str):
        raise TypeError("ping_time_bin must be a string")

    dev = require_cuda()
    rda = ds_Sv[range_var]
    range_has_nan = None
    if range_var_max is None:
        law = getattr(rda, "law", None)
        rng_t = None if (law is not None and law.get("minmax") is not None) else _range_tensor(rda, dev)
        rmax, range_has_nan = _range_max_and_nan(rda, rng_t)
    else:
        rmax = _parse_x_bin(range_var_max) + 1e-8
    r_edges = range_edges(rmax, range_bin)
    pt = np.asarray(ds_Sv["ping_time"].values)
    p_edges = ping_time_edges(pt, ping_time_bin)
    _warn_nan_coords("ping_time", pt, range_var, bool(range_has_nan))
    xbin = assign_bins(pt, p_edges, closed)
    nX = len(p_edges) - 1
    mvbs, _ = _reduce(ds_Sv, range_var, xbin, nX, r_edges, closed, skipna, fill_value, to_db=True)

    dim_0 = _first_dim(ds_Sv)
    ds_MVBS = Dataset(
        data_vars={"Sv": ((dim_0, "ping_time", range_var), mvbs)},
        coords={"ping_time": p_edges[:-1], dim_0: ds_Sv[dim_0].values, range_var: r_edges[:-1]},
    )
    if all(v in ds_Sv for v in POSITION_VARIABLES):  # commongrid/utils.py:453-501
        for var in POSITION_VARIABLES:
            ds_MVBS[var] = (("ping_time",), binned_nanmean(ds_Sv[var].values, xbin, nX), dict(ds_Sv[var].attrs))
    if range_var == "echo_range" and "water_level" in ds_Sv.data_vars:
        ds_MVBS["water_level"] = ds_Sv["water_level"]

    _set_MVBS_attrs(ds_MVBS)
    ds_MVBS[range_var].attrs.clear()
    ds_MVBS[range_var].attrs.update({"long_name": "Range distance", "units": "m"})
    resvalue, reslabel = ping_time_bin_parsing_and_conversion(ping_time_bin)
    ds_MVBS["Sv"].attrs.update(
        {
            "cell_methods": (
                f"ping_time: mean (interval: {resvalue} {reslabel} "
                "comment: ping_time is the interval start) "
                f"{range_var}: mean (interval: {range_bin} meter "
                f"comment: {range_var} is the interval start)"
            ),
            "binning_mode": "physical units",
            "range_meter_interval": str(range_bin) + "m",
            "ping_time_interval": ping_time_bin,
        }
    )
    prov_dict = echopype_prov_attrs(process_type="processing")
    prov_dict["processing_function"] = "commongrid.compute_MVBS"
    ds_MVBS = ds_MVBS.assign_attrs(prov_dict)
    ds_MVBS["frequency_nominal"] = ds_Sv["frequency_nominal"]
    if "channel" in ds_Sv:
        ds_MVBS["channel"] = ds_Sv["channel"]
    ds_MVBS = insert_input_processing_level(ds_MVBS, input_ds=ds_Sv)
    return ds_MVBS


def _coarsen_time_mean(ping_time, ping_num):
    """ping_time of DataArray.coarsen(ping_time=ping_num, boundary="pad").mean() (commongrid/api.py:219-238): the
    coordinate is coarsened with coord_func="mean", i.e. the MEAN time of each tile's pings (the NaT padding of a short
    last tile is skipped), not the tile's first ping.  int64 nanoseconds, averaged as offsets from the tile's first ping."""
    t = np.asarray(ping_time).astype("datetime64[ns]").astype(np.int64)
    n = t.shape[0]
    nP = -(-n // ping_num)
    # vectorised over the tiles (a Python loop over 10^4 tiles cost 50 ms): offsets from each tile's first ping as float64
    # (integers far below 2^53: their sum is exact in any order, so the mean equals xarray's), padding excluded
    first = t[::ping_num]
    pad = np.full(nP * ping_num, np.iinfo(np.int64).min, dtype=np.int64)
    pad[:n] = t
    tiles = pad.reshape(nP, ping_num)
    valid = np.arange(nP * ping_num).reshape(nP, ping_num) < n
    off = np.where(valid, (tiles - first[:, None]).astype(np.float64), 0.0)
    mean = off.sum(axis=1) / valid.sum(axis=1)
    out = first + mean.astype(np.int64)  # truncated to ns like int(np.mean(...))
    return out.astype("datetime64[ns]")


@add_processing_level("L3*")
def compute_MVBS_index_binning(ds_Sv, range_sample_num=100, ping_num=100):
    """
    Compute MVBS based on intervals of ``range_sample`` and ping number specified in index number
    (echopype.commongrid.compute_MVBS_index_binning).
    """
    ds_Sv = as_dataset(ds_Sv)
    dev = require_cuda()
    sv = ds_Sv["Sv"]
    if tuple(sv.dims) != DIMS:
        raise ValueError(f"Sv must have dims {DIMS}, got {tuple(sv.dims)}")
    C, P, R = sv.shape
    Sv_t = to_device_f32(sv.data, dev)
    er_t = to_device_f32(ds_Sv["echo_range"].data, dev)
    out, er = kernels.coarsen(Sv_t, er_t, C, P, R, int(ping_num), int(range_sample_num))
    out_np = out.cpu().numpy().astype(np.float64)
    nP, nRs = out_np.shape[1:]
    ds_MVBS = Dataset(
        data_vars={"Sv": (DIMS, out_np), "echo_range": (DIMS, er.cpu().numpy().astype(np.float64))},
        coords={
            "channel": ds_Sv["channel"].values,
            "ping_time": _coarsen_time_mean(np.asarray(ds_Sv["ping_time"].values), int(ping_num))[:nP],
            "range_sample": ("range_sample", np.arange(nRs), {"long_name": "Along-range sample number, base 0"}),
        },
    )
    _set_MVBS_attrs(ds_MVBS)
    with np.errstate(invalid="ignore"):
        finite = out_np[~np.isnan(out_np)]
    lo, hi = (float(finite.min()), float(finite.max())) if finite.size else (float("nan"), float("nan"))
    ds_MVBS["Sv"].attrs.update(
        {
            "cell_methods": (
                f"ping_time: mean (interval: {ping_num} pings "
                "comment: ping_time is the interval start) "
                f"range_sample: mean (interval: {range_sample_num} samples along range "
                "comment: range_sample is the interval start)"
            ),
            "comment": "MVBS binned on the basis of range_sample and ping number specified as index numbers",
            "binning_mode": "sample number",
            "range_sample_interval": f"{range_sample_num} samples along range",
            "ping_interval": f"{ping_num} pings",
            "actual_range": [round(lo, 2), round(hi, 2)],
        }
    )
    prov_dict = echopype_prov_attrs(process_type="processing")
    prov_dict["processing_function"] = "commongrid.compute_MVBS_index_binning"
    ds_MVBS = ds_MVBS.assign_attrs(prov_dict)
    ds_MVBS["frequency_nominal"] = ds_Sv["frequency_nominal"]
    ds_MVBS = insert_input_processing_level(ds_MVBS, input_ds=ds_Sv)
    return ds_MVBS


@add_processing_level("L4")
def compute_NASC(
    ds_Sv,
    range_bin: str = "10m",
    dist_bin: str = "0.5nmi",
    method: str = "map-reduce",
    skipna=True,
    closed: Literal["left", "right"] = "left",
    **flox_kwargs,
):
    """
    Compute Nautical Areal Scattering Coefficient (NASC) from an Sv dataset holding ``depth``, ``latitude`` and
    ``longitude`` (echopype.commongrid.compute_NASC).
    """
    ds_Sv = as_dataset(ds_Sv)
    range_var = "depth"
    ds_Sv, range_bin = _setup_and_validate(ds_Sv, range_var, range_bin, closed, required_data_vars=POSITION_VARIABLES)
    if not isinstance(dist_bin, str):
        raise TypeError("dist_bin must be a string")
    dist_bin = _parse_x_bin(dist_bin, "dist_bin")
    dev = require_cuda()
    dist_nmi = get_distance_from_latlon(ds_Sv["latitude"].values, ds_Sv["longitude"].values)
    rda = ds_Sv[range_var]
    rng_t = _range_tensor(rda, dev)
    rmax, range_has_nan = _range_max_and_nan(DataArray(rng_t, rda.dims), rng_t)
    r_edges = range_edges(rmax, range_bin)
    d_edges = np.arange(0, np.nanmax(dist_nmi) + dist_bin, dist_bin)
    _warn_nan_coords("distance_nmi", dist_nmi, range_var, bool(range_has_nan))
    xbin = assign_bins(dist_nmi, d_edges, closed)
    nD = len(d_edges) - 1
    ds_dev = ds_Sv.copy()
    ds_dev["depth"] = DataArray(rng_t, rda.dims)
    sv_mean, h_num = _reduce(ds_dev, range_var, xbin, nD, r_edges, closed, skipna, np.nan, to_db=False, with_height=True)
    h_den = np.bincount(xbin[xbin >= 0], minlength=nD).astype(np.float64)  # nansum of ones per distance bin
    with np.errstate(invalid="ignore", divide="ignore"):
        nasc = sv_mean * (h_num / h_den[None, :, None]) * 4 * np.pi * 1852**2  # commongrid/utils.py:202-205
    pt = np.asarray(ds_Sv["ping_time"].values).astype("datetime64[ns]")
    pt0 = pt.min()
    pt_mean = binned_nanmean((pt - pt0).astype(np.int64).astype(np.float64), xbin, nD)
    pt_out = np.where(np.isnan(pt_mean), np.datetime64("NaT"), pt0 + np.nan_to_num(pt_mean).astype("timedelta64[ns]"))

    ds_NASC = Dataset(
        data_vars={"NASC": (("channel", "distance", range_var), nasc)},
        coords={"distance": d_edges[:-1], "channel": ds_Sv["channel"].values, range_var: r_edges[:-1]},
    )
    for var in POSITION_VARIABLES:
        ds_NASC[var] = (("distance",), binned_nanmean(ds_Sv[var].values, xbin, nD), dict(ds_Sv[var].attrs))
    ds_NASC["ping_time"] = (("distance",), pt_out, dict(ds_Sv["ping_time"].attrs))
    ds_NASC["frequency_nominal"] = ds_Sv["frequency_nominal"]
    _set_var_attrs(ds_NASC["NASC"], "Nautical Areal Scattering Coefficient (NASC, m2 nmi-2)", "m2 nmi-2", 3)
    _set_var_attrs(ds_NASC["distance"], "Cumulative distance", "nmi", 3)
    _set_var_attrs(ds_NASC["depth"], "Cell depth", "m", 3, standard_name="depth")
    ds_NASC.attrs["Conventions"] = "CF-1.7,ACDD-1.3"
    ds_NASC.attrs["time_coverage_start"] = np.datetime_as_string(pt.min(), timezone="UTC")
    ds_NASC.attrs["time_coverage_end"] = np.datetime_as_string(pt.max(), timezone="UTC")
    lat, lon = np.asarray(ds_Sv["latitude"].values, float), np.asarray(ds_Sv["longitude"].values, float)
    ds_NASC.attrs["geospatial_lat_min"] = round(float(np.nanmin(lat)), 5)
    ds_NASC.attrs["geospatial_lat_max"] = round(float(np.nanmax(lat)), 5)
    ds_NASC.attrs["geospatial_lon_min"] = round(float(np.nanmin(lon)), 5)
    ds_NASC.attrs["geospatial_lon_max"] = round(float(np.nanmax(lon)), 5)
    return ds_NASC
