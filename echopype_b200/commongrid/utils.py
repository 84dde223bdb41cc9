"""Host-side helpers of the commongrid path: argument parsing / validation with the reference's messages
(echopype/commongrid/utils.py:305-377 _parse_x_bin, :380-450 _setup_and_validate, :654-698
ping_time_bin_parsing_and_conversion), bin-edge construction (commongrid/api.py:108-128, :350-362), ping ->
bin assignment, geodesic distance (utils.py:210-231) and the tiny per-ping-bin position means (:453-501).
Everything here is O(ping_time) or smaller; the O(channel x ping x range) reduction runs on the device."""

import re
from typing import Literal

import numpy as np
import pandas as pd

from ..utils.log import _init_logger

logger = _init_logger(__name__)

POSITION_VARIABLES = ["latitude", "longitude"]


def _parse_x_bin(x_bin: str, x_label="range_bin") -> float:
    """Parse a bin-size string such as '10m' or '0.5nmi' into a float (commongrid/utils.py:305-377)."""
    x_bin_map = {
        "range_bin": {"name": "Range bin", "unit": "m", "ex": "10m", "unit_label": "meters",
                      "pattern": r"([\d+]*[.,]{0,1}[\d+]*)(\s+)?(m)"},
        "dist_bin": {"name": "Distance bin", "unit": "nmi", "ex": "0.5nmi", "unit_label": "nautical miles",
                     "pattern": r"([\d+]*[.,]{0,1}[\d+]*)(\s+)?(nmi)"},
    }
    x_bin_info = x_bin_map.get(x_label, None)
    if x_bin_info is None:
        raise KeyError(f"x_label must be one of {list(x_bin_map.keys())}")
    if not isinstance(x_bin, str):
        raise TypeError("'x_bin' must be a string")
    x_bin = x_bin.strip().lower()
    match_obj = re.match(x_bin_info["pattern"], x_bin)
    if match_obj is not None:
        return float(match_obj.group(1))
    raise ValueError(f"{x_bin_info['name']} must be in {x_bin_info['unit_label']} (e.g., '{x_bin_info['ex']}').")


def _setup_and_validate(ds_Sv, range_var="echo_range", range_bin=None, closed: Literal["left", "right"] = "left",
                        required_data_vars=None):
    if range_var not in ["echo_range", "depth"]:
        raise ValueError("range_var must be one of 'echo_range' or 'depth'.")
    if required_data_vars is None:
        required_data_vars = []
    required_data_vars = set(required_data_vars + [range_var])
    if not all([var in ds_Sv.variables for var in required_data_vars]):
        raise ValueError("Input Sv dataset must contain all of " f"the following variables: {required_data_vars}")
    if not isinstance(range_bin, str):
        raise TypeError("range_bin must be a string")
    range_bin = _parse_x_bin(range_bin, "range_bin")
    if closed not in ["right", "left"]:
        raise ValueError(f"{closed} is not a valid option. Options are 'left' or 'right'.")
    if "filenames" in ds_Sv.dims:
        ds_Sv = ds_Sv.drop_dims("filenames")
    return ds_Sv, range_bin


def ping_time_bin_parsing_and_conversion(ping_time_bin: str):
    timedelta_units = {
        "d": {"nptd64": "D", "unitstr": "day"}, "h": {"nptd64": "h", "unitstr": "hour"},
        "t": {"nptd64": "m", "unitstr": "minute"}, "min": {"nptd64": "m", "unitstr": "minute"},
        "s": {"nptd64": "s", "unitstr": "second"}, "l": {"nptd64": "ms", "unitstr": "millisecond"},
        "ms": {"nptd64": "ms", "unitstr": "millisecond"}, "u": {"nptd64": "us", "unitstr": "microsecond"},
        "us": {"nptd64": "ms", "unitstr": "millisecond"}, "n": {"nptd64": "ns", "unitstr": "nanosecond"},
        "ns": {"nptd64": "ms", "unitstr": "millisecond"},
    }
    td = pd.Timedelta(ping_time_bin)
    resunit = td.resolution_string.lower()
    resvalue = int(td / np.timedelta64(1, timedelta_units[resunit]["nptd64"]))
    return resvalue, timedelta_units[resunit]["unitstr"]


def range_edges(range_var_max: float, range_bin: float) -> np.ndarray:
    """commongrid/api.py:115 / :350-352."""
    return np.arange(0, range_var_max + range_bin, range_bin)


def ping_time_edges(ping_time, ping_time_bin: str) -> np.ndarray:
    """commongrid/api.py:118-124: pandas resample grid (origin = start of day) plus one closing edge."""
    idx = pd.DatetimeIndex(np.asarray(ping_time).astype("datetime64[ns]"))
    d_index = pd.Series(np.zeros(len(idx)), index=idx).resample(ping_time_bin).first().index
    return d_index.union([d_index[-1] + pd.Timedelta(ping_time_bin)]).values.astype("datetime64[ns]")


def assign_bins(x, edges, closed="left") -> np.ndarray:
    """Interval membership of each x against ascending edges (pd.IntervalIndex.from_breaks(edges, closed)):
    returns int32 bin index, -1 for values outside every interval or NaN / NaT."""
    x = np.asarray(x)
    edges = np.asarray(edges)
    if x.dtype.kind == "M":
        bad = np.isnat(x)
        xv, ev = x.astype("datetime64[ns]").astype(np.int64), edges.astype("datetime64[ns]").astype(np.int64)
    else:
        xv, ev = x.astype(np.float64), edges.astype(np.float64)
        bad = np.isnan(xv)
    if closed == "left":
        idx = np.searchsorted(ev, xv, side="right") - 1
        ok = (xv >= ev[0]) & (xv < ev[-1])
    else:
        idx = np.searchsorted(ev, xv, side="left") - 1
        ok = (xv > ev[0]) & (xv <= ev[-1])
    return np.where(ok & ~bad, idx, -1).astype(np.int32)


def binned_nanmean(values, codes, nbins) -> np.ndarray:
    """nanmean of a per-ping series within each x bin (flox nanmean; empty bin -> NaN)."""
    v = np.asarray(values, dtype=np.float64)
    good = (codes >= 0) & ~np.isnan(v)
    s = np.bincount(codes[good], weights=v[good], minlength=nbins)
    n = np.bincount(codes[good], minlength=nbins)
    with np.errstate(invalid="ignore", divide="ignore"):
        return s / n


# ---- WGS-84 geodesic (geopy.distance.distance = Karney's geodesic; restated with Vincenty's inverse
# formula, which agrees to well below a millimetre for the ping-to-ping distances involved) -----------------
_A = 6378137.0
_F = 1 / 298.257223563
_B = (1 - _F) * _A


def geodesic_nmi(lat1, lon1, lat2, lon2) -> np.ndarray:
    """Vectorised geodesic distance in nautical miles between consecutive positions (degrees)."""
    lat1, lon1, lat2, lon2 = (np.asarray(a, dtype=np.float64) for a in (lat1, lon1, lat2, lon2))
    U1 = np.arctan((1 - _F) * np.tan(np.radians(lat1)))
    U2 = np.arctan((1 - _F) * np.tan(np.radians(lat2)))
    L = np.radians(lon2 - lon1)
    sU1, cU1, sU2, cU2 = np.sin(U1), np.cos(U1), np.sin(U2), np.cos(U2)
    lam = L.copy()
    with np.errstate(invalid="ignore", divide="ignore"):
        for _ in range(200):
            sl, cl = np.sin(lam), np.cos(lam)
            ss = np.hypot(cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl)
            cs = sU1 * sU2 + cU1 * cU2 * cl
            sig = np.arctan2(ss, cs)
            sa = np.where(ss == 0, 0.0, cU1 * cU2 * sl / np.where(ss == 0, 1.0, ss))
            c2a = 1 - sa * sa
            c2sm = np.where(c2a == 0, 0.0, cs - 2 * sU1 * sU2 / np.where(c2a == 0, 1.0, c2a))
            Cc = _F / 16 * c2a * (4 + _F * (4 - 3 * c2a))
            lam_new = L + (1 - Cc) * _F * sa * (sig + Cc * ss * (c2sm + Cc * cs * (-1 + 2 * c2sm * c2sm)))
            done = np.nanmax(np.abs(lam_new - lam), initial=0.0) < 1e-13
            lam = lam_new
            if done:
                break
        u2 = c2a * (_A * _A - _B * _B) / (_B * _B)
        A = 1 + u2 / 16384 * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
        Bc = u2 / 1024 * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
        ds = Bc * ss * (c2sm + Bc / 4 * (cs * (-1 + 2 * c2sm**2) - Bc / 6 * c2sm * (-3 + 4 * ss**2) * (-3 + 4 * c2sm**2)))
        d = _B * A * (sig - ds)
    d = np.where((lat1 == lat2) & (lon1 == lon2), 0.0, d)
    return d / 1852.0


def get_distance_from_latlon(latitude, longitude) -> np.ndarray:
    """commongrid/utils.py:210-231: dist[p] = geodesic(p, p+1) assigned to ping p, cumulative sum, then
    forward- and backward-fill of the pings without a valid pair."""
    lat = np.asarray(latitude, dtype=np.float64)
    lon = np.asarray(longitude, dtype=np.float64)
    if lat.size == 0:
        raise ValueError("All lat/lon entries are NaN!")
    lat_n = np.append(lat[1:], np.nan)
    lon_n = np.append(lon[1:], np.nan)
    ok = ~(np.isnan(lat) | np.isnan(lon) | np.isnan(lat_n) | np.isnan(lon_n))
    if not ok.any():
        raise ValueError("All lat/lon entries are NaN!")
    dist = np.full(lat.shape, np.nan)
    dist[ok] = geodesic_nmi(lat[ok], lon[ok], lat_n[ok], lon_n[ok])
    s = pd.Series(dist).cumsum()  # NaN entries are skipped by cumsum and stay NaN
    return s.ffill().bfill().values
