"""Oracle: MVBS / NASC / index-binned MVBS reductions.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  echopype/commongrid/api.py:31-191 (compute_MVBS), :195-266 (compute_MVBS_index_binning),
  :270-416 (compute_NASC);
  echopype/commongrid/utils.py:17-94 (compute_raw_MVBS), :97-207 (compute_raw_NASC),
  :210-231 (get_distance_from_latlon), :283-302 (IntervalIndex), :305-377 (_parse_x_bin),
  :504-628 (_groupby_x_along_channels), :654-698 (ping_time_bin parsing).
Third-party arithmetic absent from /root/reference:
  * flox >=0.7.2 (requirements.txt:5, unpinned) ``xarray_reduce(func="nanmean"|"mean"|"nansum",
    isbin=True, expected_groups=IntervalIndex)``: restated as np.digitize factorisation (values
    outside all intervals or with NaN coordinate get code -1 and are dropped) + bincount sums;
    bins with no members -> fill_value; ``mean`` propagates NaN values, ``nanmean`` skips them.
  * pandas resample (origin="start_day") - pandas 3.0.2 is installed and used directly.
  * geopy.distance.distance (WGS-84 geodesic, geographiclib/Karney) - absent; restated with
    Vincenty's inverse formula on the WGS-84 ellipsoid (agrees with Karney to < 0.1 mm away from
    antipodal points).
Pinned by the reference's own known-answer tests: brute-force MVBS (tests/mock_data.py:28-85),
NASC Echoview closed form (tests/commongrid/conftest.py:426-444), brute-force NASC (:467-615),
skipna NaN patterns (tests/commongrid/test_commongrid_api.py:484-556) - restated in
tests/test_oracle_golden.py.
"""

import re
import warnings

import numpy as np
import pandas as pd

from .clean import coarsen_mean, coarsen_min, lin2log, log2lin

_X_BIN = {
    "range_bin": ("Range bin", "m", "10m", "meters", r"([\d+]*[.,]{0,1}[\d+]*)(\s+)?(m)"),
    "dist_bin": ("Distance bin", "nmi", "0.5nmi", "nautical miles", r"([\d+]*[.,]{0,1}[\d+]*)(\s+)?(nmi)"),
}


def parse_x_bin(x_bin, x_label="range_bin"):
    """commongrid/utils.py:305-377."""
    info = _X_BIN.get(x_label)
    if info is None:
        raise KeyError(f"x_label must be one of {list(_X_BIN.keys())}")
    if not isinstance(x_bin, str):
        raise TypeError("'x_bin' must be a string")
    m = re.match(info[4], x_bin.strip().lower())
    if m is None:
        raise ValueError(f"{info[0]} must be in {info[3]} (e.g., '{info[2]}').")
    return float(m.group(1))


def range_edges(range_var_max, range_bin):
    """commongrid/api.py:108-115."""
    return np.arange(0, range_var_max + range_bin, range_bin)


def ping_edges(ping_time_ns, ping_time_bin):
    """commongrid/api.py:118-124.  Returns int64 ns edges (len = nbins + 1)."""
    idx = pd.DatetimeIndex(np.asarray(ping_time_ns).astype("datetime64[ns]"))
    d_index = pd.Series(np.zeros(len(idx)), index=idx).resample(ping_time_bin).first().index
    edges = d_index.union([d_index[-1] + pd.Timedelta(ping_time_bin)])
    return edges.values.astype("datetime64[ns]").astype(np.int64)


def bin_codes(x, edges, closed="left"):
    """flox factorisation of ``x`` against IntervalIndex.from_breaks(edges, closed).  -1 = dropped."""
    x = np.asarray(x)
    edges = np.asarray(edges)
    right = closed == "right"
    isnan = np.isnan(x) if x.dtype.kind == "f" else np.zeros(x.shape, dtype=bool)
    xs = np.where(isnan, edges[0], x) if x.dtype.kind == "f" else x
    idx = np.digitize(xs, edges, right=right) - 1
    within = (xs <= edges.max()) if right else (xs < edges.max())
    idx = np.where(within & ~isnan & (idx >= 0), idx, -1)
    return idx.astype(np.int64)


def groupby_mean(sv_lin, x, x_edges, rng, r_edges, skipna=True, fill_value=np.nan, closed="left"):
    """commongrid/utils.py:504-628 (the flox call at :614-627).

    sv_lin (C,X,R) linear; x (X,) ping_time ns or distance; rng (C,X,R) echo_range/depth.
    Returns mean (C, nX, nR) float64.
    """
    C, X, R = sv_lin.shape
    nX, nR = len(x_edges) - 1, len(r_edges) - 1
    xc = bin_codes(x, x_edges, closed)  # (X,)
    rc = bin_codes(rng, r_edges, closed)  # (C,X,R)
    out = np.full((C, nX, nR), np.nan)
    for c in range(C):
        code = np.where((xc[:, None] >= 0) & (rc[c] >= 0), xc[:, None] * nR + rc[c], -1).ravel()
        v = sv_lin[c].ravel()
        member = code >= 0
        nmem = np.bincount(code[member], minlength=nX * nR)
        isn = np.isnan(v)
        good = member & ~isn
        s = np.bincount(code[good], weights=v[good], minlength=nX * nR)
        n = np.bincount(code[good], minlength=nX * nR)
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = s / n  # members but all NaN -> 0/0 = NaN
        if not skipna:
            nnan = np.bincount(code[member & isn], minlength=nX * nR)
            mean = np.where(nnan > 0, np.nan, mean)
        mean = np.where(nmem == 0, fill_value, mean)
        out[c] = mean.reshape(nX, nR)
    return out


def compute_MVBS(
    Sv, rng, ping_time_ns, range_bin="20m", ping_time_bin="20s", skipna=True, fill_value=np.nan,
    closed="left", range_var_max=None,
):
    """commongrid/api.py:31-191.  Returns dict(Sv (C,nP,nR), ping_time (left edges ns), range (left edges))."""
    rb = parse_x_bin(range_bin)
    if not isinstance(ping_time_bin, str):
        raise TypeError("ping_time_bin must be a string")
    if closed not in ("right", "left"):
        raise ValueError(f"{closed} is not a valid option. Options are 'left' or 'right'.")
    if range_var_max is None:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            rmax = np.nanmax(rng)
    else:
        rmax = parse_x_bin(range_var_max) + 1e-8
    r_edges = range_edges(rmax, rb)
    p_edges = ping_edges(ping_time_ns, ping_time_bin)
    mean = groupby_mean(
        log2lin(np.asarray(Sv, dtype=np.float64)), np.asarray(ping_time_ns, dtype=np.int64), p_edges,
        np.asarray(rng, dtype=np.float64), r_edges, skipna, fill_value, closed,
    )
    return {"Sv": lin2log(mean), "ping_time": p_edges[:-1], "range": r_edges[:-1], "r_edges": r_edges, "p_edges": p_edges}


def compute_MVBS_index_binning(Sv, echo_range, range_sample_num=100, ping_num=100):
    """commongrid/api.py:195-266."""
    sv = 10 ** (np.asarray(Sv, dtype=np.float64) / 10)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = 10 * np.log10(coarsen_mean(sv, ping_num, range_sample_num))
    er = coarsen_min(np.asarray(echo_range, dtype=np.float64), ping_num, range_sample_num)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        ar = [round(float(np.nanmin(out)), 2), round(float(np.nanmax(out)), 2)]
    return {"Sv": out, "echo_range": er, "actual_range": ar}


# ------------------------------------------------------------------------------------------------
# NASC
# ------------------------------------------------------------------------------------------------
_WGS84_A = 6378137.0
_WGS84_F = 1 / 298.257223563


def geodesic_m(lat1, lon1, lat2, lon2):
    """WGS-84 geodesic distance in metres (Vincenty inverse), scalar inputs in degrees."""
    if lat1 == lat2 and lon1 == lon2:
        return 0.0
    a, f = _WGS84_A, _WGS84_F
    b = (1 - f) * a
    U1 = np.arctan((1 - f) * np.tan(np.radians(lat1)))
    U2 = np.arctan((1 - f) * np.tan(np.radians(lat2)))
    Lr = np.radians(lon2 - lon1)
    lam = Lr
    sU1, cU1, sU2, cU2 = np.sin(U1), np.cos(U1), np.sin(U2), np.cos(U2)
    for _ in range(200):
        sl, cl = np.sin(lam), np.cos(lam)
        ss = np.hypot(cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl)
        if ss == 0:
            return 0.0
        cs = sU1 * sU2 + cU1 * cU2 * cl
        sig = np.arctan2(ss, cs)
        sa = cU1 * cU2 * sl / ss
        c2a = 1 - sa * sa
        c2sm = cs - 2 * sU1 * sU2 / c2a if c2a != 0 else 0.0
        Cc = f / 16 * c2a * (4 + f * (4 - 3 * c2a))
        lam_new = Lr + (1 - Cc) * f * sa * (sig + Cc * ss * (c2sm + Cc * cs * (-1 + 2 * c2sm * c2sm)))
        if abs(lam_new - lam) < 1e-13:
            lam = lam_new
            break
        lam = lam_new
    u2 = c2a * (a * a - b * b) / (b * b)
    A = 1 + u2 / 16384 * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
    Bc = u2 / 1024 * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
    ds = Bc * ss * (c2sm + Bc / 4 * (cs * (-1 + 2 * c2sm**2) - Bc / 6 * c2sm * (-3 + 4 * ss**2) * (-3 + 4 * c2sm**2)))
    return float(b * A * (sig - ds))


def distance_from_latlon(latitude, longitude):
    """commongrid/utils.py:210-231.  dist[p] = d(p, p+1) assigned to ping p; cumsum; ffill; bfill."""
    df = pd.DataFrame({"latitude": np.asarray(latitude, float), "longitude": np.asarray(longitude, float)})
    df["latitude_prev"] = df["latitude"].shift(-1)
    df["longitude_prev"] = df["longitude"].shift(-1)
    nn = df.dropna().copy()
    if len(nn) == 0:
        raise ValueError("All lat/lon entries are NaN!")
    nn["dist"] = [
        geodesic_m(r.latitude, r.longitude, r.latitude_prev, r.longitude_prev) / 1852.0 for r in nn.itertuples()
    ]
    df = df.join(nn["dist"], how="left")
    df["dist"] = df["dist"].cumsum()
    df["dist"] = df["dist"].ffill().bfill()
    return df["dist"].values


def compute_raw_NASC(Sv, depth, dist_nmi, ping_time_ns, r_edges, d_edges, skipna=True, closed="left"):
    """commongrid/utils.py:97-207.  Sv, depth (C,X,R); dist_nmi, ping_time (X,)."""
    Sv = np.asarray(Sv, dtype=np.float64)
    depth = np.asarray(depth, dtype=np.float64)
    C, X, R = Sv.shape
    nD, nR = len(d_edges) - 1, len(r_edges) - 1
    sv_mean = groupby_mean(log2lin(Sv), dist_nmi, d_edges, depth, r_edges, skipna, np.nan, closed)
    dc = bin_codes(dist_nmi, d_edges, closed)
    ok = dc >= 0
    h_den = np.bincount(dc[ok], minlength=nD).astype(np.float64)  # nansum of ones; empty -> 0
    pt = np.asarray(ping_time_ns, dtype=np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        pt_mean = np.bincount(dc[ok], weights=(pt[ok] - pt.min()).astype(np.float64), minlength=nD) / h_den
    ping_time_mean = np.where(h_den > 0, pt_mean + pt.min(), np.nan)
    diff = depth[:, :, 1:] - depth[:, :, :-1]  # label="lower"
    lower = depth[:, :, :-1]
    rc = bin_codes(lower, r_edges, closed)
    h_num = np.zeros((C, nD, nR))
    for c in range(C):
        code = np.where((dc[:, None] >= 0) & (rc[c] >= 0), dc[:, None] * nR + rc[c], -1).ravel()
        v = diff[c].ravel()
        good = (code >= 0) & ~np.isnan(v)
        h_num[c] = np.bincount(code[good], weights=v[good], minlength=nD * nR).reshape(nD, nR)
    with np.errstate(invalid="ignore", divide="ignore"):
        h_mean = h_num / h_den[None, :, None]
        nasc = sv_mean * h_mean * 4 * np.pi * 1852**2
    return {"NASC": nasc, "ping_time": ping_time_mean, "sv_mean": sv_mean, "h_mean": h_mean}


def compute_NASC(Sv, depth, latitude, longitude, ping_time_ns, range_bin="10m", dist_bin="0.5nmi", skipna=True, closed="left"):
    """commongrid/api.py:270-416."""
    rb = parse_x_bin(range_bin)
    if not isinstance(dist_bin, str):
        raise TypeError("dist_bin must be a string")
    db = parse_x_bin(dist_bin, "dist_bin")
    dist = distance_from_latlon(latitude, longitude)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        r_edges = np.arange(0, np.nanmax(depth) + rb, rb)
        d_edges = np.arange(0, np.nanmax(dist) + db, db)
    out = compute_raw_NASC(Sv, depth, dist, ping_time_ns, r_edges, d_edges, skipna, closed)
    out.update({"distance": d_edges[:-1], "depth": r_edges[:-1], "dist_nmi": dist, "r_edges": r_edges, "d_edges": d_edges})
    return out
