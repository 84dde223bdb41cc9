"""EK80 data with several ``filter_time`` entries in ``Vendor_specific`` (echopype/calibrate/api.py:95-197).

Two host-side control paths around the same device kernels:

* ``assume_single_filter_time=True`` (api.py:103-125, calibrate_ek.py:37-52 ``_collapse_vend``): every channel uses the
  filter set whose ``filter_time`` equals the channel's first valid ping (first ping with a non-NaN
  ``transmit_duration_nominal``); the filter dimension is collapsed and the volume is calibrated in one go.
* otherwise (api.py:127-197): the volume is calibrated piece by piece - one (channel, filter interval) at a time,
  the interval running from one of the channel's filter times to the nanosecond before the next
  (calibrate_ek.py:25-34 ``_slice_beam_vend``) - and the pieces are combined like
  ``xr.merge(pieces, join="outer", compat="no_conflicts")``: sorted union of the channel / ping_time labels, NaN where no
  piece has a value, an error where two pieces disagree.
"""

import numpy as np
import torch

from ..dataset import DataArray, Dataset, EchoData

_MERGE_MSG = ("conflicting values for variable {!r} on objects to be combined. "
              "You can skip this check by specifying compat='override'.")


def _ns(t):
    return np.asarray(t).astype("datetime64[ns]").astype(np.int64)


def _with_groups(echodata: EchoData, **groups) -> EchoData:
    new = dict(echodata._groups)
    new.update({k.replace("__", "/"): v for k, v in groups.items()})
    return EchoData(echodata.sonar_model, new, source_file=echodata.source_file, converted_raw_path=echodata.converted_raw_path)


def first_valid_filter_time_per_channel(beam):
    """api.py:109-119: per channel the first ping_time whose transmit_duration_nominal is not NaN."""
    tau = np.asarray(beam["transmit_duration_nominal"].values, dtype=np.float64)
    pt = np.asarray(beam["ping_time"].values)
    out = {}
    for ci, ch in enumerate(np.asarray(beam["channel"].values)):
        ok = np.flatnonzero(~np.isnan(tau[ci]))
        if ok.size == 0:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # valid_ping_times[0] on an empty array
        out[ch] = pt[ok[0]]
    return out


def collapse_vend(vend: Dataset, first_valid: dict) -> Dataset:
    """calibrate_ek.py:37-52: Vendor_specific without the filter_time dimension, every channel taken at its own
    filter time (KeyError when that time is not one of the filter times, as ``Dataset.sel`` would raise)."""
    ft = _ns(vend["filter_time"].values)
    vch = np.asarray(vend["channel"].values)
    order = np.argsort(np.asarray([str(c) for c in first_valid]))  # the merge sorts the channel labels
    chans = [list(first_valid)[i] for i in order]
    idx = []
    for ch in chans:
        hit = np.flatnonzero(ft == int(_ns(first_valid[ch])))
        if hit.size == 0:
            raise KeyError(first_valid[ch])
        idx.append((int(np.flatnonzero(vch == ch)[0]), int(hit[0])))
    out = Dataset(attrs=dict(vend.attrs))
    for name, c in vend.coords.items():
        if name == "filter_time":
            continue
        out._set_coord(name, np.asarray(chans, dtype=object) if name == "channel" else c)
    for name in vend:
        v = vend[name]
        a = np.asarray(v.values)
        dims = tuple(v.dims)
        if "channel" not in dims:
            if "filter_time" in dims:
                raise ValueError(f"{name}: filter_time without channel cannot be collapsed")
            out[name] = (dims, a)
            continue
        ca = dims.index("channel")
        if "filter_time" in dims:
            fa = dims.index("filter_time")
            rows = []
            for ci, fi in idx:
                sl = [slice(None)] * a.ndim
                sl[ca], sl[fa] = ci, fi
                rows.append(a[tuple(sl)])
            nd = tuple(d for d in dims if d not in ("channel", "filter_time"))
            out[name] = (("channel",) + nd, np.stack(rows, axis=0))
        else:
            out[name] = (dims, np.take(a, [ci for ci, _ in idx], axis=ca))
    return out


def filter_pieces(beam, vend):
    """api.py:141-163: [(channel index, ping indices, filter_time index)] in the reference's loop order (channels in
    groupby = sorted order, filter times ascending)."""
    tau = np.asarray(beam["transmit_duration_nominal"].values, dtype=np.float64)
    pt = _ns(beam["ping_time"].values)
    ft_all = np.sort(_ns(vend["filter_time"].values))
    ft_raw = _ns(vend["filter_time"].values)
    chans = np.asarray(beam["channel"].values)
    pieces = []
    for ci in np.argsort(np.asarray([str(c) for c in chans])):
        valid = pt[~np.isnan(tau[ci])]
        fts = np.intersect1d(valid, ft_all)
        for k, start in enumerate(fts):
            end = fts[k + 1] - 1 if k + 1 < len(fts) else None
            sel = (pt >= start) if end is None else ((pt >= start) & (pt <= end))
            pieces.append((int(ci), np.flatnonzero(sel), int(np.flatnonzero(ft_raw == start)[0])))
    return pieces


def piece_echodata(echodata: EchoData, beam_group: str, ci: int, p_idx, fi: int) -> EchoData:
    """calibrate_ek.py:25-34: beam group cut to one channel and a ping interval, Vendor_specific at one filter time."""
    p_idx = np.asarray(p_idx)
    contiguous = p_idx.size > 0 and np.array_equal(p_idx, np.arange(p_idx[0], p_idx[0] + p_idx.size))
    psel = slice(int(p_idx[0]), int(p_idx[0]) + p_idx.size) if contiguous else p_idx  # an interval of a sorted axis: a view
    beam = echodata[beam_group].isel(channel=[ci]).isel(ping_time=psel)
    vend = echodata["Vendor_specific"].isel(filter_time=[fi])
    ed = _with_groups(echodata)
    ed[beam_group] = beam
    ed["Vendor_specific"] = vend
    return ed


def _isnull(a):
    if a.dtype.kind == "f":
        return np.isnan(a)
    if a.dtype.kind == "M":
        return np.isnat(a)
    if a.dtype.kind == "O":
        return np.array([x is None or (isinstance(x, float) and x != x) for x in a.ravel()], dtype=bool).reshape(a.shape)
    return np.zeros(a.shape, dtype=bool)


def merge_pieces(pieces):
    """``xr.merge(pieces, join="outer", compat="no_conflicts")`` (api.py:193-197) for the piece Datasets of compute_Sv /
    compute_TS: the (channel, ping_time, range_sample) volumes stay on the device."""
    by_name = {str(c): c for p in pieces for c in np.asarray(p["channel"].values)}
    chan = np.asarray([by_name[k] for k in sorted(by_name)], dtype=object)  # the outer join sorts the labels
    ping = np.unique(np.concatenate([np.asarray(p["ping_time"].values).astype("datetime64[ns]") for p in pieces]))
    first = pieces[0]
    coords = {}
    for name, c in first.coords.items():
        if name == "channel":
            coords[name] = chan
        elif name == "ping_time":
            coords[name] = ping
        else:
            coords[name] = c
    out = Dataset(coords=coords, attrs=dict(first.attrs))
    names = []
    for p in pieces:
        names += [n for n in p if n not in names]
    for name in names:
        src = [p for p in pieces if name in p]
        v0 = src[0][name]
        dims = tuple(v0.dims)
        shape = tuple(len(chan) if d == "channel" else len(ping) if d == "ping_time" else v0.shape[i] for i, d in enumerate(dims))
        on_dev = isinstance(v0.data, torch.Tensor)
        if on_dev:
            full = torch.full(shape, float("nan"), dtype=v0.data.dtype, device=v0.data.device)
        else:
            a0 = np.asarray(v0.values)
            if a0.dtype.kind in "iub":
                a0 = a0.astype(np.float64)  # the outer join introduces NaN
            full = np.full(shape, np.nan if a0.dtype.kind == "f" else (np.datetime64("NaT") if a0.dtype.kind == "M" else None),
                           dtype=a0.dtype if a0.dtype.kind in "fM" else object)
        for p in src:
            v = p[name]
            index = []
            for d in v.dims:
                if d == "channel":
                    index.append(np.array([int(np.flatnonzero(chan == c)[0]) for c in np.asarray(p["channel"].values)]))
                elif d == "ping_time":
                    index.append(np.searchsorted(ping, np.asarray(p["ping_time"].values).astype("datetime64[ns]")))
                else:
                    index.append(np.arange(v.shape[list(v.dims).index(d)]))
            if not index:  # scalar variable
                new = np.asarray(v.values)
                if full.shape == () and not _isnull(np.asarray(full)) and not _isnull(new) and full != new:
                    raise ValueError(_MERGE_MSG.format(name))
                full = new.astype(full.dtype) if full.dtype != object else new
                continue
            ix = np.ix_(*index)
            if on_dev:
                dev_ix = tuple(torch.as_tensor(i, device=full.device) for i in ix)
                full[dev_ix] = v.data  # pieces are disjoint (channel, ping interval) blocks
            else:
                new = np.asarray(v.values)
                old = full[ix]
                both = ~_isnull(old) & ~_isnull(new)
                if both.any() and np.any(old[both] != new[both]):
                    raise ValueError(_MERGE_MSG.format(name))
                full[ix] = np.where(_isnull(new), old, new)
        da = DataArray(full, dims, attrs=v0.attrs, name=name)
        out[name] = da
    return out
