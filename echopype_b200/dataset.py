"""Minimal labelled-array containers standing in for xarray.Dataset / DataArray / EchoData.

xarray is not installable in this environment, so the drop-in API of
``echopype_b200.calibrate / clean / commongrid`` is served through these duck-typed containers:
``ds["Sv"]``, ``.dims``, ``.sizes``, ``.shape``, ``.attrs``, ``.coords``, ``.values``, ``.data``,
``.isel``, ``.copy``, ``.assign_attrs``.  If a real xarray object is passed in, it is converted with
:func:`as_dataset`; :meth:`Dataset.to_xarray` converts back when xarray is importable.

A DataArray's ``.data`` is either a numpy array (host) or a torch CUDA tensor (device buffer).  Full-
size products stay on the device until ``.values`` is read, mirroring how the reference returns lazy
dask-backed arrays for chunked inputs.  ``.law`` optionally carries the exact range law of an
``echo_range`` / ``depth`` variable produced by this package ("index-space decisions").
"""

from __future__ import annotations

import copy as _copy

import numpy as np

try:  # torch is only used as the device-buffer type
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


class DataArray:
    __slots__ = ("data", "dims", "coords", "attrs", "name", "law")

    def __init__(self, data, dims=None, coords=None, attrs=None, name=None, law=None):
        if isinstance(data, DataArray):
            dims = data.dims if dims is None else dims
            coords = data.coords if coords is None else coords
            attrs = data.attrs if attrs is None else attrs
            name = data.name if name is None else name
            data = data.data
        if not _is_tensor(data):
            data = np.asarray(data)
        if dims is None:
            if data.ndim != 0:
                raise ValueError("dims are required for non-scalar data")
            dims = ()
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != data.ndim:
            raise ValueError(f"dims {dims} do not match data of shape {tuple(data.shape)}")
        self.data = data
        self.dims = dims
        self.coords = {}
        if coords:
            items = coords.items() if isinstance(coords, dict) else coords
            for k, v in items:
                if isinstance(v, DataArray):
                    v = v.values
                if isinstance(v, tuple) and len(v) >= 2 and isinstance(v[0], (str, tuple, list)):
                    v = v[1]
                v = np.asarray(v)
                if k in dims and v.ndim == 1 and v.shape[0] != data.shape[dims.index(k)]:
                    raise ValueError(f"coordinate {k!r} has length {v.shape[0]}, expected {data.shape[dims.index(k)]}")
                self.coords[k] = v
        self.attrs = dict(attrs) if attrs else {}
        self.name = name
        self.law = law

    # ---- basic properties --------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return len(self.dims)

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def sizes(self):
        return dict(zip(self.dims, self.shape))

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def nbytes(self):
        if _is_tensor(self.data):
            return self.data.numel() * self.data.element_size()
        return self.data.nbytes

    @property
    def on_device(self):
        return _is_tensor(self.data) and self.data.is_cuda

    @property
    def values(self):
        if _is_tensor(self.data):
            return self.data.detach().cpu().numpy()
        return self.data

    def __array__(self, dtype=None, copy=None):
        v = self.values
        return v.astype(dtype) if dtype is not None else v

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        where = "cuda" if self.on_device else "host"
        return f"<DataArray {self.name or ''} {self.sizes} {self.dtype} [{where}]>"

    def item(self):
        return self.values.item()

    def __float__(self):
        return float(self.values)

    # ---- labelled access -----------------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.coords:
                v = self.coords[key]
                return DataArray(v, dims=(key,) if v.ndim == 1 else (), coords={key: v} if v.ndim == 1 else None, name=key)
            raise KeyError(key)
        out = self.values[key]
        return out

    def isel(self, **indexers):
        idx = []
        new_dims = []
        new_coords = dict(self.coords)
        for d in self.dims:
            sel = indexers.get(d, slice(None))
            idx.append(sel)
            scalar = isinstance(sel, (int, np.integer))
            if not scalar:
                new_dims.append(d)
            if d in new_coords and np.ndim(new_coords[d]) == 1:
                new_coords[d] = new_coords[d][sel]
                if scalar:
                    new_coords.pop(d)
        data = self.data[tuple(idx)]
        return DataArray(data, dims=new_dims, coords=new_coords, attrs=self.attrs, name=self.name)

    def sel(self, **labels):
        indexers = {}
        for d, lab in labels.items():
            cv = self.coords[d]
            if isinstance(lab, (list, tuple, np.ndarray)):
                indexers[d] = [int(np.flatnonzero(cv == v)[0]) for v in lab]
            else:
                hit = np.flatnonzero(cv == lab)
                if hit.size == 0:
                    raise KeyError(lab)
                indexers[d] = int(hit[0])
        return self.isel(**indexers)

    def transpose(self, *dims):
        perm = [self.dims.index(d) for d in dims]
        data = self.data.permute(*perm) if _is_tensor(self.data) else np.transpose(self.data, perm)
        return DataArray(data, dims=dims, coords=self.coords, attrs=self.attrs, name=self.name)

    def copy(self, deep=False):
        data = self.data
        if deep:
            data = data.clone() if _is_tensor(data) else data.copy()
        return DataArray(data, self.dims, dict(self.coords), dict(self.attrs), self.name, self.law)

    def assign_attrs(self, *args, **kw):
        out = self.copy()
        for a in args:
            out.attrs.update(a)
        out.attrs.update(kw)
        return out

    def astype(self, dtype):
        out = self.copy()
        out.data = self.values.astype(dtype)
        return out

    def where(self, cond, other=np.nan):
        """Values where ``cond`` holds, ``other`` elsewhere (xarray.DataArray.where for same-shaped or dims-broadcastable
        conditions).  Device arrays stay on the device (the comparison result ``cond`` may be a device DataArray)."""
        c = cond
        cdims = tuple(c.dims) if isinstance(c, DataArray) else None
        cdata = c.data if isinstance(c, DataArray) else c
        if _is_tensor(self.data) or _is_tensor(cdata):
            x = self.data if _is_tensor(self.data) else torch.as_tensor(np.asarray(self.data))
            ct = cdata if _is_tensor(cdata) else torch.as_tensor(np.asarray(cdata))
            dev = x.device if x.is_cuda else ct.device
            x, ct = x.to(dev), ct.to(dev)
            if cdims is not None and cdims != self.dims:
                ct = ct.reshape([ct.shape[cdims.index(d)] if d in cdims else 1 for d in self.dims])
            xf = x if x.is_floating_point() else x.float()
            o = other.data if isinstance(other, DataArray) else other
            o = o.to(dev) if _is_tensor(o) else torch.as_tensor(np.asarray(o, dtype=np.float64), device=dev).to(xf.dtype)
            out = torch.where(ct.bool(), xf, o)
        else:
            x, cn = np.asarray(self.data), np.asarray(cdata)
            if cdims is not None and cdims != self.dims:
                cn = cn.reshape([cn.shape[cdims.index(d)] if d in cdims else 1 for d in self.dims])
            o = other.values if isinstance(other, DataArray) else other
            if x.dtype.kind in "iub" and np.ndim(o) == 0 and isinstance(o, float) and np.isnan(o):
                x = x.astype(np.float64)
            out = np.where(cn.astype(bool), x, o)
        return DataArray(out, self.dims, dict(self.coords), dict(self.attrs), self.name)

    def rename(self, new_name_or_name_dict=None, **names):
        if isinstance(new_name_or_name_dict, str) or (new_name_or_name_dict is None and not names):
            out = self.copy()
            out.name = new_name_or_name_dict
            return out
        m = dict(new_name_or_name_dict or {}, **names)
        return DataArray(self.data, tuple(m.get(d, d) for d in self.dims), {m.get(k, k): v for k, v in self.coords.items()},
                         dict(self.attrs), m.get(self.name, self.name), self.law)

    def isnull(self):
        v = self.values
        m = np.isnan(v) if v.dtype.kind in "fc" else (np.isnat(v) if v.dtype.kind in "mM" else np.zeros(v.shape, bool))
        return DataArray(m, self.dims, self.coords, name=self.name)

    def _reduce(self, fn_host, fn_dev):
        if _is_tensor(self.data):
            return DataArray(np.asarray(fn_dev(self.data).item()))
        return DataArray(np.asarray(fn_host(self.data)))

    def min(self, skipna=True):
        if _is_tensor(self.data):
            d = self.data
            if d.is_floating_point():
                m = torch.isnan(d)
                if bool(m.all()):
                    return DataArray(np.asarray(np.nan))
                d = torch.where(m, torch.full_like(d, float("inf")), d)
            return DataArray(np.asarray(d.min().item()))
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            return DataArray(np.asarray((np.nanmin if skipna and self.data.dtype.kind == "f" else np.min)(self.data)))

    def max(self, skipna=True):
        if _is_tensor(self.data):
            d = self.data
            if d.is_floating_point():
                m = torch.isnan(d)
                if bool(m.all()):
                    return DataArray(np.asarray(np.nan))
                d = torch.where(m, torch.full_like(d, float("-inf")), d)
            return DataArray(np.asarray(d.max().item()))
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            return DataArray(np.asarray((np.nanmax if skipna and self.data.dtype.kind == "f" else np.max)(self.data)))

    # arithmetic for the small host-side parameter arrays (dims-aware broadcasting, no alignment)
    def _binop(self, other, op):
        a_dims, a = self.dims, self.values
        if isinstance(other, DataArray):
            b_dims, b = other.dims, other.values
            coords = {**other.coords, **self.coords}
        else:
            b_dims, b = (), np.asarray(other)
            coords = dict(self.coords)
        out_dims = tuple(a_dims) + tuple(d for d in b_dims if d not in a_dims)

        def expand(x, dims):
            perm_src = [d for d in out_dims if d in dims]
            x = np.transpose(x, [dims.index(d) for d in perm_src]) if x.ndim else x
            shape = [x.shape[perm_src.index(d)] if d in dims else 1 for d in out_dims]
            return x.reshape(shape) if x.ndim else x

        with np.errstate(all="ignore"):
            res = op(expand(a, a_dims), expand(b, b_dims))
        coords = {k: v for k, v in coords.items() if k in out_dims or np.ndim(v) == 0}
        return DataArray(res, out_dims, coords)

    def __add__(self, o):
        return self._binop(o, np.add)

    def __radd__(self, o):
        return self._binop(o, lambda a, b: b + a)

    def __sub__(self, o):
        return self._binop(o, np.subtract)

    def __rsub__(self, o):
        return self._binop(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._binop(o, np.multiply)

    def __rmul__(self, o):
        return self._binop(o, lambda a, b: b * a)

    def __truediv__(self, o):
        return self._binop(o, np.divide)

    def __rtruediv__(self, o):
        return self._binop(o, lambda a, b: b / a)

    def __pow__(self, o):
        return self._binop(o, np.power)

    def __neg__(self):
        return DataArray(-self.values, self.dims, self.coords, self.attrs, self.name)

    def _cmp(self, o, op_host, op_dev):
        if _is_tensor(self.data):
            ov = o.data if isinstance(o, DataArray) else o
            return DataArray(op_dev(self.data, ov), self.dims, dict(self.coords), None, self.name)
        return self._binop(o, op_host)

    def __lt__(self, o):
        return self._cmp(o, np.less, lambda a, b: a < b)

    def __le__(self, o):
        return self._cmp(o, np.less_equal, lambda a, b: a <= b)

    def __gt__(self, o):
        return self._cmp(o, np.greater, lambda a, b: a > b)

    def __ge__(self, o):
        return self._cmp(o, np.greater_equal, lambda a, b: a >= b)

    def notnull(self):
        m = self.isnull()
        return DataArray(~m.values, m.dims, m.coords, name=self.name)


def _to_dataarray(value, name=None):
    if isinstance(value, DataArray):
        return value
    if isinstance(value, tuple):
        dims, data = value[0], value[1]
        attrs = value[2] if len(value) > 2 else None
        return DataArray(data, dims=dims, attrs=attrs, name=name)
    if hasattr(value, "dims") and hasattr(value, "values"):  # real xarray.DataArray
        coords = {k: np.asarray(v.values) for k, v in value.coords.items() if k in value.dims}
        return DataArray(np.asarray(value.values), dims=tuple(value.dims), coords=coords, attrs=dict(value.attrs), name=name)
    return DataArray(np.asarray(value), dims=(), name=name)


class Dataset:
    """Dict of named DataArrays sharing dimension coordinates (stand-in for xarray.Dataset)."""

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = {}
        self._coords = {}
        self.attrs = dict(attrs) if attrs else {}
        for k, v in (coords or {}).items():
            self._set_coord(k, v)
        for k, v in (data_vars or {}).items():
            self[k] = v

    def _set_coord(self, name, value):
        if isinstance(value, tuple) and len(value) >= 2 and isinstance(value[0], (str, tuple, list)):
            dims, data = value[0], value[1]
            attrs = value[2] if len(value) > 2 else None
            da = DataArray(np.asarray(data), dims=dims, attrs=attrs, name=name)
        elif isinstance(value, DataArray):
            da = DataArray(value.values, value.dims, attrs=value.attrs, name=name)
        else:
            arr = np.asarray(value)
            da = DataArray(arr, dims=(name,) if arr.ndim == 1 else (), name=name)
        self._coords[name] = da

    # ---- mapping interface ------------------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self[k] for k in key}, coords=self._coords, attrs=self.attrs)
        if key in self._vars:
            da = self._vars[key]
            cs = {d: self._coords[d].values for d in da.dims if d in self._coords}
            cs.update({k: v for k, v in da.coords.items() if k not in cs})
            out = DataArray(da.data, da.dims, cs, da.attrs, key, da.law)
            out.attrs = da.attrs  # share, so ds["x"].attrs.update(...) sticks like xarray
            return out
        if key in self._coords:
            c = self._coords[key]
            out = DataArray(c.data, c.dims, {key: c.values} if c.dims == (key,) else None, None, key)
            out.attrs = c.attrs
            return out
        raise KeyError(key)

    def __setitem__(self, key, value):
        da = _to_dataarray(value, key)
        for d, n in zip(da.dims, da.shape):
            if d in self._coords and self._coords[d].dims == (d,) and self._coords[d].shape[0] != n:
                raise ValueError(f"conflicting sizes for dimension {d!r}: {n} vs {self._coords[d].shape[0]}")
            have = self.sizes.get(d)
            if have is not None and have != n:
                raise ValueError(f"conflicting sizes for dimension {d!r}: {n} vs {have}")
        for d in da.dims:
            if d in da.coords and d not in self._coords:
                self._set_coord(d, da.coords[d])
        if key in self._coords and da.dims == (key,):
            self._set_coord(key, da)
            return
        if key == "Sv":
            # The index-space binning of compute_MVBS (the row table cached on echo_range / depth as ``.law``) relies on
            # Sv being NaN wherever the reference's range variable is NaN.  Arrays produced by this package from such an
            # Sv carry the marker law {"kind": "derived"}; any other array assigned as "Sv" (user data, a mask filled
            # with a finite value) ends that guarantee, so the range variables fall back to value binning.
            keeps = isinstance(da.law, dict) and da.law.get("kind") == "derived"
            old = self._vars.get("Sv")
            same = old is not None and old.data is da.data
            if not (keeps or same):
                self._drop_range_laws()
        stored = DataArray(da.data, da.dims, None, None, key, da.law)
        stored.attrs = da.attrs
        self._vars[key] = stored

    def _drop_range_laws(self):
        for name in ("echo_range", "depth"):
            v = self._vars.get(name)
            if v is not None and isinstance(v.law, dict) and v.law.get("rows") is not None:
                mm = v.law.get("minmax")
                v.law = {"kind": v.law.get("kind"), "rows": None, "minmax": mm} if mm is not None else None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        d = self.__dict__
        if name in d.get("_vars", {}) or name in d.get("_coords", {}):
            return self[name]
        raise AttributeError(f"Dataset has no variable or attribute {name!r}")

    def rename_vars(self, name_dict=None, **names):
        """xarray.Dataset.rename_vars: rename variables / coordinates, dimensions stay."""
        m = dict(name_dict or {}, **names)
        for k in m:
            if k not in self:
                raise ValueError(f"cannot rename {k!r} because it is not a variable or coordinate in this dataset")
        for k, v in m.items():
            if v in self and v not in m:
                raise ValueError(f"the new name {v!r} conflicts")
        out = Dataset(attrs=_copy.copy(self.attrs))
        for k, c in self._coords.items():
            c2 = c.copy()
            c2.name = m.get(k, k)
            out._coords[m.get(k, k)] = c2
        for k, v in self._vars.items():
            v2 = v.copy()
            v2.name = m.get(k, k)
            out._vars[m.get(k, k)] = v2
        if "Sv" in m.values():  # another variable takes the place of Sv: same rule as assignment
            sv = out._vars.get("Sv")
            if sv is not None and not (isinstance(sv.law, dict) and sv.law.get("kind") == "derived"):
                out._drop_range_laws()
        return out

    def rename(self, name_dict=None, **names):
        """xarray.Dataset.rename: variables, coordinates and dimensions."""
        m = dict(name_dict or {}, **names)
        out = self.rename_vars({k: v for k, v in m.items() if k in self})
        for store in (out._coords, out._vars):
            for k, v in list(store.items()):
                if any(d in m for d in v.dims):
                    store[k] = DataArray(v.data, tuple(m.get(d, d) for d in v.dims), None, v.attrs, v.name, v.law)
        return out

    def swap_dims(self, dims_dict=None, **dims):
        """xarray.Dataset.swap_dims: e.g. {"channel": "frequency_nominal"} makes the 1-D variable ``frequency_nominal``
        the dimension coordinate of what was the ``channel`` dimension (``channel`` stays as a non-index coordinate)."""
        m = dict(dims_dict or {}, **dims)
        out = self.copy()
        for old, new in m.items():
            if new not in out:
                raise ValueError(f"replacement dimension {new!r} is not a 1D variable along the old dimension {old!r}")
            nv = out[new]
            if nv.dims != (old,):
                raise ValueError(f"replacement dimension {new!r} is not a 1D variable along the old dimension {old!r}")
            out._vars.pop(new, None)
            out._coords[new] = DataArray(nv.values, (new,), None, nv.attrs, new)
            for store in (out._coords, out._vars):
                for k, v in list(store.items()):
                    if k != new and old in v.dims:
                        store[k] = DataArray(v.data, tuple(new if d == old else d for d in v.dims), None, v.attrs, v.name, v.law)
        return out

    def where(self, cond, other=np.nan):
        """xarray.Dataset.where over the data variables that contain every dimension of ``cond``."""
        out = self.copy()
        cdims = set(cond.dims) if isinstance(cond, DataArray) else None
        for k, v in self._vars.items():
            if cdims is None or cdims <= set(v.dims):
                out._vars[k] = DataArray(self[k].where(cond, other).data, v.dims, None, v.attrs, k)
        if "Sv" in out._vars and not (isinstance(other, float) and np.isnan(other)):
            out._drop_range_laws()
        return out

    def __contains__(self, key):
        return key in self._vars or key in self._coords

    def __iter__(self):
        return iter(self._vars)

    def __repr__(self):
        lines = [f"<Dataset sizes={self.sizes}>"]
        for k, v in self._coords.items():
            lines.append(f"  * {k} {v.dims} {v.dtype}")
        for k, v in self._vars.items():
            lines.append(f"    {k} {v.dims} {v.dtype}{' [cuda]' if v.on_device else ''}")
        return "\n".join(lines)

    def keys(self):
        return self._vars.keys()

    @property
    def data_vars(self):
        return {k: self[k] for k in self._vars}

    @property
    def variables(self):
        return {**{k: self[k] for k in self._coords}, **self.data_vars}

    @property
    def coords(self):
        return {k: self[k] for k in self._coords}

    @property
    def sizes(self):
        out = {}
        for da in list(self._vars.values()) + list(self._coords.values()):
            for d, n in zip(da.dims, da.shape):
                out.setdefault(d, n)
        return out

    dims = sizes

    def get(self, key, default=None):
        return self[key] if key in self else default

    def copy(self, deep=False):
        out = Dataset(attrs=_copy.copy(self.attrs))
        out._coords = {k: v.copy(deep) for k, v in self._coords.items()}
        out._vars = {k: v.copy(deep) for k, v in self._vars.items()}
        return out

    def assign_attrs(self, *args, **kw):
        out = self.copy()
        for a in args:
            out.attrs.update(a)
        out.attrs.update(kw)
        return out

    def assign(self, **kw):
        out = self.copy()
        for k, v in kw.items():
            out[k] = v
        return out

    def assign_coords(self, **kw):
        out = self.copy()
        for k, v in kw.items():
            out._set_coord(k, v)
        return out

    def drop_vars(self, names, errors="raise"):
        names = [names] if isinstance(names, str) else list(names)
        out = self.copy()
        for n in names:
            if n in out._vars:
                del out._vars[n]
            elif n in out._coords:
                del out._coords[n]
            elif errors == "raise":
                raise ValueError(f"cannot drop {n!r}: not found")
        return out

    def drop_dims(self, names):
        names = [names] if isinstance(names, str) else list(names)
        out = self.copy()
        out._vars = {k: v for k, v in out._vars.items() if not set(v.dims) & set(names)}
        out._coords = {k: v for k, v in out._coords.items() if not set(v.dims) & set(names)}
        return out

    def isel(self, **indexers):
        out = Dataset(attrs=dict(self.attrs))
        for k, c in self._coords.items():
            if set(c.dims) & set(indexers):
                c = c.isel(**{d: s for d, s in indexers.items() if d in c.dims})
            out._coords[k] = c
        for k, v in self._vars.items():
            if set(v.dims) & set(indexers):
                law = v.law
                v = DataArray(v.data, v.dims, None, v.attrs, k).isel(**{d: s for d, s in indexers.items() if d in v.dims})
                v.coords = {}
                _ = law  # a sliced array no longer matches its row table
            out._vars[k] = v
        return out

    def pipe(self, fn, *args, **kw):
        return fn(self, *args, **kw)

    def to_xarray(self):  # xarray is not installed in the build image: exercised with a stand-in (tests/test_host_dataset.py)
        import xarray as xr

        return xr.Dataset(
            {k: (v.dims, v.values, v.attrs) for k, v in self._vars.items()},
            coords={k: (v.dims, v.values, v.attrs) for k, v in self._coords.items()},
            attrs=self.attrs,
        )


def as_dataset(obj):
    """Accept a Dataset, a real xarray.Dataset, or a mapping name -> (dims, data[, attrs])."""
    if isinstance(obj, Dataset):
        return obj
    if hasattr(obj, "data_vars") and hasattr(obj, "coords"):
        ds = Dataset(attrs=dict(getattr(obj, "attrs", {})))
        for k, c in obj.coords.items():
            ds._set_coord(k, (tuple(c.dims), np.asarray(c.values), dict(c.attrs)))
        for k, v in obj.data_vars.items():
            ds[k] = (tuple(v.dims), np.asarray(v.values), dict(v.attrs))
        return ds
    if isinstance(obj, dict):
        return Dataset(obj)
    raise TypeError(f"cannot interpret {type(obj)} as a Dataset")


class EchoData:
    """Thin stand-in for echopype.echodata.EchoData (echodata/echodata.py:327-335): a dict of
    SONAR-netCDF4 groups ("Sonar/Beam_group1", "Environment", "Vendor_specific", "Platform", "Sonar")."""

    def __init__(self, sonar_model, groups=None, source_file=None, converted_raw_path=None):
        self.sonar_model = sonar_model
        self._groups = {k: as_dataset(v) for k, v in (groups or {}).items()}
        self.source_file = source_file
        self.converted_raw_path = converted_raw_path

    def __getitem__(self, key):
        if key not in self._groups:
            if key in ("Platform", "Top-level", "Sonar", "Provenance"):
                return Dataset()
            raise KeyError(key)
        return self._groups[key]

    def __setitem__(self, key, value):
        self._groups[key] = as_dataset(value)

    def __contains__(self, key):
        return key in self._groups

    @property
    def group_paths(self):
        return list(self._groups)
