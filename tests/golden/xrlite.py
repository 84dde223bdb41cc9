"""xrlite - a small labelled-array stand-in for the subset of xarray that echopype's
calibrate / clean array code uses.  GENERATOR-SIDE TEST INFRASTRUCTURE ONLY.

Why: xarray (and dask, flox, zarr ...) cannot be installed in the builder image, so the reference
package cannot be imported as is.  ``make_golden_calibrate.py`` registers this module as ``xarray``
and then imports the reference's OWN modules (calibrate/range.py, calibrate_ek.py, calibrate_azfp.py,
cal_params.py, env_params.py, ek80_complex.py, utils/uwa.py, utils/align.py ...) unmodified from
/root/reference, so the numbers in tests/golden/calibrate_vectors.npz are produced by the reference's
code, not by a restatement of it.  This file only supplies the container semantics those modules rely
on; they are implemented after xarray's documented behaviour:

  * arithmetic / ufuncs broadcast by dimension NAME; result dims are ordered by first appearance;
    operands are inner-joined on their index coordinates (order of the first operand);
    scalar (non-index) coordinates survive unless they conflict;
  * reductions on float / complex data skip NaN by default (skipna=None -> True);
  * ``where`` keeps values where the condition holds; ``da[bool_da]`` / ``.loc[{dim: bool_da}]``
    index the FIRST dim / the named dim by mask; integer ``isel`` keeps a scalar coordinate;
  * vectorised ``sel(dim=DataArray)`` is point-wise along dims the indexer shares with the array;
  * ``interp`` is linear 1-D (scipy ``interp1d``; datetimes as float ns offsets from their minimum);
  * ``coarsen(boundary="pad").mean()`` pads with NaN and takes nan-means; ``reindex(method="ffill")``.

Anything outside that subset raises NotImplementedError loudly instead of guessing.
"""

import numpy as np

__version__ = "xrlite-0"


def _is_scalar(x):
    return np.ndim(x) == 0 and not isinstance(x, (DataArray,))


class _Coord:
    __slots__ = ("dims", "data", "attrs")

    def __init__(self, dims, data, attrs=None):
        self.dims = tuple(dims)
        self.data = np.asarray(data)
        self.attrs = dict(attrs or {})
        assert self.data.ndim == len(self.dims), (dims, self.data.shape)


def _as_coord(name, v, default_dim=None):
    if isinstance(v, _Coord):
        return _Coord(v.dims, v.data, v.attrs)
    if isinstance(v, DataArray):
        return _Coord(v.dims, v._data, v.attrs)
    if isinstance(v, tuple) and len(v) in (2, 3) and isinstance(v[0], (str, tuple, list)):
        d = (v[0],) if isinstance(v[0], str) else tuple(v[0])
        return _Coord(d, np.asarray(v[1]), v[2] if len(v) == 3 else None)
    a = np.asarray(v)
    if a.ndim == 0:
        return _Coord((), a)
    if a.ndim == 1:
        return _Coord((default_dim or name,), a)
    raise NotImplementedError(f"coordinate {name!r} with {a.ndim} dims and no dims given")


class _CoordsView:
    """``da.coords`` / ``ds.coords``: mapping name -> DataArray."""

    def __init__(self, owner):
        self._o = owner

    def __contains__(self, k):
        return k in self._o._coords

    def __iter__(self):
        return iter(self._o._coords)

    def __len__(self):
        return len(self._o._coords)

    def keys(self):
        return self._o._coords.keys()

    def items(self):
        return [(k, self[k]) for k in self._o._coords]

    def __getitem__(self, k):
        return self._o._coord_da(k)

    def __setitem__(self, k, v):  # ds.coords["x"] = (dims, data[, attrs]): replaces the coordinate in place
        c = _as_coord(k, v)
        if hasattr(self._o, "_check_fits"):
            self._o._check_fits(k, c)
        self._o._coords[k] = c


def _index_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind in "Mm" or b.dtype.kind in "Mm":
        return bool(np.all(a.astype("datetime64[ns]") == b.astype("datetime64[ns]")))
    return bool(np.all(a == b))


def _positions(index, labels):
    """positions of ``labels`` in the 1-D ``index`` (exact match; KeyError when missing)."""
    index = np.asarray(index)
    lab = np.asarray(labels)
    if index.dtype.kind == "M":
        index = index.astype("datetime64[ns]")
        lab = lab.astype("datetime64[ns]")
    flat = lab.reshape(-1)
    out = np.empty(flat.shape, dtype=np.int64)
    for i, v in enumerate(flat):
        hit = np.nonzero(index == v)[0]
        if hit.size == 0:
            raise KeyError(f"{v!r} not found in index")
        out[i] = hit[0]
    return out.reshape(lab.shape)


def _skipna_default(a, skipna):
    if skipna is None:
        return a.dtype.kind in "fcO"
    return bool(skipna)


class DataArray:
    __array_priority__ = 60
    __hash__ = None

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            coords = coords if coords is not None else {k: v for k, v in data._coords.items()}
            dims = dims if dims is not None else data.dims
            data = data._data
        data = np.asarray(data)
        if dims is None:
            if coords is not None and hasattr(coords, "keys") and len(coords) == data.ndim:
                dims = tuple(coords.keys())
            elif data.ndim == 0:
                dims = ()
            else:
                dims = tuple(f"dim_{i}" for i in range(data.ndim))
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != data.ndim:
            raise ValueError(f"different number of dimensions on data and dims: {data.ndim} vs {len(dims)}")
        self._data = data
        self.dims = dims
        self._coords = {}
        if coords is not None and not hasattr(coords, "keys"):  # a list of coordinates, one per dimension
            coords = {d: (c._data if isinstance(c, DataArray) else c) for d, c in zip(dims, coords)}
        if coords is not None:
            for k in coords.keys():
                c = _as_coord(k, coords[k])
                for d, n in zip(c.dims, c.data.shape):
                    if d not in dims:
                        raise ValueError(f"coordinate {k} has dimension {d} not in the array dims {dims}")
                    if n != data.shape[dims.index(d)]:
                        raise ValueError(f"conflicting sizes for dimension {d!r}")
                self._coords[k] = c
        self.name = name
        self.attrs = dict(attrs) if attrs else {}

    # ---- basic properties -------------------------------------------------------------------------
    @property
    def values(self):
        return self._data

    @values.setter
    def values(self, v):
        v = np.asarray(v)
        assert v.shape == self._data.shape
        self._data = v

    data = values

    @property
    def shape(self):
        return self._data.shape

    @property
    def size(self):
        return self._data.size

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def nbytes(self):
        return self._data.nbytes

    @property
    def sizes(self):
        return dict(zip(self.dims, self._data.shape))

    @property
    def coords(self):
        return _CoordsView(self)

    @property
    def chunks(self):
        return None

    @property
    def T(self):
        return self.transpose()

    def __len__(self):
        return self._data.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __bool__(self):
        return bool(self._data)

    def __float__(self):
        return float(self._data)

    def __int__(self):
        return int(self._data)

    def __iter__(self):
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        for i in range(self.shape[0]):
            yield self.isel({self.dims[0]: i})

    def __contains__(self, v):
        return bool(np.any(self._data == v))

    def __repr__(self):
        return f"<xrlite.DataArray {self.name!r} {self.sizes} coords={list(self._coords)}>\n{self._data!r}"

    def item(self):
        return self._data.item()

    def copy(self, deep=True):
        return DataArray(self._data.copy() if deep else self._data, {k: _Coord(c.dims, c.data.copy(), c.attrs) for k, c in self._coords.items()},
                         self.dims, self.name, dict(self.attrs))

    def compute(self):
        return self

    load = compute
    persist = compute

    def _coord_da(self, k):
        c = self._coords[k]
        sub = {n: cc for n, cc in self._coords.items() if set(cc.dims) <= set(c.dims)}
        return DataArray(c.data, sub, c.dims, name=k, attrs=c.attrs)

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        co = self.__dict__.get("_coords", {})
        if k in co:
            return self._coord_da(k)
        raise AttributeError(f"xrlite.DataArray has no attribute {k!r}")

    def _new(self, data, dims=None, coords=None, name="__keep__", attrs=None):
        dims = self.dims if dims is None else tuple(dims)
        if coords is None:
            coords = {k: c for k, c in self._coords.items() if set(c.dims) <= set(dims)}
        return DataArray(data, coords, dims, self.name if name == "__keep__" else name, attrs)

    # ---- indexing -----------------------------------------------------------------------------------
    def isel(self, indexers=None, drop=False, **kw):
        ind = dict(indexers or {}, **kw)
        for d in ind:
            if d not in self.dims:
                raise ValueError(f"dimension {d!r} does not exist in {self.dims}")
        vec = {d: v for d, v in ind.items() if isinstance(v, DataArray) and v.dtype.kind != "b" and v.ndim >= 1 and v.dims != (d,)}
        if vec:
            return self._isel_vectorized(ind)
        data = self._data
        dims = list(self.dims)
        coords = {k: _Coord(c.dims, c.data, c.attrs) for k, c in self._coords.items()}
        for d, ix in ind.items():
            ax = dims.index(d)
            if isinstance(ix, DataArray):
                ix = ix._data
            scalar = (not isinstance(ix, slice)) and np.ndim(ix) == 0
            if not isinstance(ix, slice):
                ix = np.asarray(ix)
                if ix.dtype.kind == "b":
                    ix = np.nonzero(ix)[0]
            data = np.take(data, ix, axis=ax) if not isinstance(ix, slice) else data[(slice(None),) * ax + (ix,)]
            for k in list(coords):
                c = coords[k]
                if d in c.dims:
                    cax = c.dims.index(d)
                    cd = np.take(c.data, ix, axis=cax) if not isinstance(ix, slice) else c.data[(slice(None),) * cax + (ix,)]
                    cdims = tuple(x for x in c.dims if x != d) if scalar else c.dims
                    if scalar and drop and k == d:
                        del coords[k]
                    else:
                        coords[k] = _Coord(cdims, cd, c.attrs)
            if scalar:
                dims.pop(ax)
        return DataArray(data, coords, tuple(dims), self.name, self.attrs)

    def _isel_vectorized(self, ind):
        # point-wise along dims that the indexer shares with the array (xarray Variable._broadcast_indexes_vectorized)
        if len(ind) != 1:
            raise NotImplementedError("vectorised indexing along more than one dim")
        (d, ix), = ind.items()
        out_dims = []
        for x in self.dims:
            if x == d:
                for y in ix.dims:
                    if y not in out_dims:
                        out_dims.append(y)
            elif x not in out_dims:
                out_dims.append(x)
        shape = []
        for x in out_dims:
            shape.append(self.sizes[x] if x in self.dims else ix.sizes[x])
            if x in self.dims and x in ix.dims and self.sizes[x] != ix.sizes[x]:
                raise ValueError("size mismatch in vectorised indexing")

        def expand(arr, adims):
            arr = np.transpose(arr, [adims.index(x) for x in out_dims if x in adims])
            sh = [arr.shape[[y for y in out_dims if y in adims].index(x)] if x in adims else 1 for x in out_dims]
            return arr.reshape(sh)

        key = []
        for x in self.dims:
            if x == d:
                key.append(np.broadcast_to(expand(ix._data, ix.dims), shape))
            else:
                key.append(np.broadcast_to(expand(np.arange(self.sizes[x]), (x,)), shape))
        data = self._data[tuple(key)]
        coords = {k: c for k, c in self._coords.items() if d not in c.dims}
        for k, c in ix._coords.items():
            if k not in coords and k != d:
                coords[k] = c
        return DataArray(data, coords, tuple(out_dims), self.name, self.attrs)

    def _label_to_pos(self, d, lab):
        if d not in self._coords:
            raise KeyError(f"no index for dimension {d!r}")
        index = self._coords[d].data
        if isinstance(lab, slice):
            if lab.step is not None:
                raise NotImplementedError
            idx = index.astype("datetime64[ns]") if index.dtype.kind == "M" else index
            lo = 0 if lab.start is None else int(np.searchsorted(idx, np.asarray(lab.start).astype(idx.dtype), "left"))
            hi = len(idx) if lab.stop is None else int(np.searchsorted(idx, np.asarray(lab.stop).astype(idx.dtype), "right"))
            return slice(lo, hi)
        if isinstance(lab, DataArray):
            if lab.dtype.kind == "b":
                return lab._data
            pos = _positions(index, lab._data)
            if lab.ndim == 0:
                return int(pos)
            return DataArray(pos, {k: c for k, c in lab._coords.items() if k != d}, lab.dims)
        arr = np.asarray(lab)
        if arr.dtype.kind == "b" and arr.ndim == 1:
            return arr
        pos = _positions(index, arr)
        return int(pos) if arr.ndim == 0 else pos

    def sel(self, indexers=None, drop=False, method=None, **kw):
        if method is not None:
            raise NotImplementedError("sel(method=...)")
        ind = dict(indexers or {}, **kw)
        out = self
        for d, lab in ind.items():
            if d not in out.dims:
                if d in out._coords and out._coords[d].dims == ():  # selecting on a scalar coordinate
                    if not _index_equal(out._coords[d].data, np.asarray(lab)):
                        raise KeyError(lab)
                    continue
                raise KeyError(f"{d!r} is not a valid dimension or coordinate")
            pos = out._label_to_pos(d, lab)
            out = out.isel({d: pos}, drop=drop)
            if isinstance(lab, DataArray) and lab.ndim >= 1 and lab.dtype.kind != "b" and drop and d in out._coords:
                del out._coords[d]
        return out

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._coord_da(key)
        if isinstance(key, dict):
            return self.isel(key)
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) > self.ndim:
            raise IndexError("too many indices")
        ind = {}
        for d, k in zip(self.dims, key):
            if isinstance(k, DataArray) and k.dtype.kind == "b":
                if k.dims != (d,) and k.ndim == 1:
                    k = k._data
            ind[d] = k
        return self.isel(ind)

    def _set(self, ind, value):
        """positional assignment; a DataArray value is matched by dim NAME (xarray Variable.__setitem__)."""
        key = []
        kept = []
        for d in self.dims:
            k = ind.get(d, slice(None))
            if isinstance(k, DataArray):
                k = k._data
            if not isinstance(k, slice):
                k = np.asarray(k)
                if k.dtype.kind == "b":
                    k = np.nonzero(k)[0]
            if isinstance(k, slice) or k.ndim >= 1:
                kept.append(d)
            key.append(k)
        n_adv = sum(1 for k in key if not isinstance(k, slice) and np.ndim(k) >= 1)
        if n_adv > 1:
            raise NotImplementedError("assignment with several array indexers")
        if isinstance(value, DataArray):
            extra = [d for d in value.dims if d not in kept]
            if extra:
                raise ValueError(f"assigned value has dims {extra} the target lacks")
            v = value.transpose(*[d for d in kept if d in value.dims])._data
            sh = [v.shape[[x for x in kept if x in value.dims].index(d)] if d in value.dims else 1 for d in kept]
            value = v.reshape(sh)
        self._data[tuple(key)] = value

    def __setitem__(self, key, value):
        if isinstance(key, str):
            self._coords[key] = _as_coord(key, value)
            return
        if isinstance(key, dict):
            return self._set(key, value)
        if not isinstance(key, tuple):
            key = (key,)
        self._set(dict(zip(self.dims, key)), value)

    @property
    def loc(self):
        return _Loc(self)

    # ---- arithmetic -----------------------------------------------------------------------------------
    def _broadcast_data(self, dims):
        src = [d for d in dims if d in self.dims]
        a = np.transpose(self._data, [self.dims.index(d) for d in src])
        return a.reshape([a.shape[src.index(d)] if d in src else 1 for d in dims])

    @staticmethod
    def _align(arrs):
        """inner join on index coordinates (order of the first operand that carries the index)."""
        arrs = list(arrs)
        dims = []
        for a in arrs:
            for d in a.dims:
                if d not in dims:
                    dims.append(d)
        for d in dims:
            have = [a for a in arrs if d in a.dims and d in a._coords]
            if len(have) < 2:
                sizes = {a.sizes[d] for a in arrs if d in a.dims}
                if len(sizes) > 1:
                    raise ValueError(f"cannot broadcast dimension {d!r} without an index: sizes {sizes}")
                continue
            first = have[0]._coords[d].data
            if all(_index_equal(first, h._coords[d].data) for h in have[1:]):
                continue
            keep = [v for v in first if all(np.any(h._coords[d].data == v) for h in have[1:])]
            for i, a in enumerate(arrs):
                if d in a.dims and d in a._coords:
                    pos = _positions(a._coords[d].data, np.asarray(keep, dtype=first.dtype)) if keep else np.zeros(0, int)
                    if pos.size > 1 and np.any(np.diff(pos) < 0):
                        raise NotImplementedError("inner join that would reorder an operand")
                    arrs[i] = a.isel({d: pos})
                elif d in a.dims:
                    raise ValueError(f"operand without an index on {d!r} cannot be joined")
        return arrs, tuple(dims)

    @staticmethod
    def _merge_coords(arrs, dims):
        coords = {}
        conflict = set()
        for a in arrs:
            for k, c in a._coords.items():
                if not set(c.dims) <= set(dims):
                    continue
                if k in conflict:
                    continue
                if k in coords:
                    o = coords[k]
                    if o.dims != c.dims or not _index_equal(o.data, c.data):
                        if c.dims == (k,) and o.dims == (k,):
                            raise ValueError(f"index conflict on {k!r} after alignment")
                        if c.dims == (k,) or o.dims == (k,):  # an index wins over a scalar coordinate of the same name
                            coords[k] = c if c.dims == (k,) else o
                            continue
                        del coords[k]
                        conflict.add(k)
                else:
                    coords[k] = c
        return coords

    @staticmethod
    def _apply(f, *args):
        das = [a for a in args if isinstance(a, DataArray)]
        aligned, dims = DataArray._align(das)
        it = iter(aligned)
        ins = []
        ref = aligned[0]
        for a in args:
            if isinstance(a, DataArray):
                ins.append(next(it)._broadcast_data(dims))
            else:
                arr = np.asarray(a) if not isinstance(a, (str, bytes)) else a
                if not isinstance(a, (str, bytes)) and arr.ndim > 0:
                    if arr.ndim > len(dims):
                        raise ValueError("plain array with more dims than the labelled operand")
                    if any(tuple(d.dims) != tuple(dims) for d in aligned):  # positional broadcasting would be ambiguous
                        raise NotImplementedError("plain ndarray operand next to labelled operands with different dims")
                ins.append(a if isinstance(a, (str, bytes)) else arr)
        with np.errstate(all="ignore"):
            out = f(*ins)
        name = das[0].name if all(d.name == das[0].name for d in das) else None
        return DataArray(out, DataArray._merge_coords(aligned, dims), dims, name)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            raise NotImplementedError(f"ufunc method {method}")
        return DataArray._apply(lambda *a: ufunc(*a, **kwargs), *inputs)

    def _bin(self, other, f, reflexive=False):
        if isinstance(other, Dataset):
            return NotImplemented
        return DataArray._apply(f, other, self) if reflexive else DataArray._apply(f, self, other)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    def __floordiv__(self, o): return self._bin(o, np.floor_divide)
    def __mod__(self, o): return self._bin(o, np.mod)
    def __pow__(self, o): return self._bin(o, np.power)
    def __rpow__(self, o): return self._bin(o, np.power, True)
    def __lt__(self, o): return self._bin(o, np.less)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __and__(self, o): return self._bin(o, np.logical_and if self.dtype.kind == "b" else np.bitwise_and)
    def __or__(self, o): return self._bin(o, np.logical_or if self.dtype.kind == "b" else np.bitwise_or)

    def __eq__(self, o):
        return self._bin(o, lambda a, b: np.asarray(a == b))

    def __ne__(self, o):
        return self._bin(o, lambda a, b: np.asarray(a != b))

    def __neg__(self): return self._new(-self._data)
    def __pos__(self): return self._new(+self._data)
    def __abs__(self): return self._new(np.abs(self._data))
    def __invert__(self): return self._new(~self._data)

    # ---- elementwise helpers ------------------------------------------------------------------------
    def isnull(self):
        d = self._data
        if d.dtype.kind in "fc":
            m = np.isnan(d)
        elif d.dtype.kind in "Mm":
            m = np.isnat(d)
        elif d.dtype.kind == "O":
            m = np.array([x is None or (isinstance(x, float) and x != x) for x in d.reshape(-1)], dtype=bool).reshape(d.shape)
        else:
            m = np.zeros(d.shape, bool)
        return self._new(m)

    def notnull(self):
        return ~self.isnull()

    def astype(self, t, **kw):
        with np.errstate(all="ignore"):
            return self._new(self._data.astype(t), attrs=self.attrs)

    def fillna(self, v):
        return DataArray._apply(lambda a, b: np.where(_isnull_arr(a), b, a), self, v)

    def where(self, cond, other=np.nan, drop=False):
        if drop:
            raise NotImplementedError("where(drop=True)")
        if callable(cond):
            cond = cond(self)

        def f(a, c, o):
            if np.ndim(o) == 0 and isinstance(o, float) and np.isnan(o) and a.dtype.kind in "iub":
                a = a.astype(np.float64)
            return np.where(c, a, o)

        out = DataArray._apply(f, self, cond, other)
        out.name = self.name
        out.attrs = dict(self.attrs)
        return out

    def clip(self, min=None, max=None):
        return self._new(np.clip(self._data, min, max))

    def round(self, n=0):
        return self._new(np.round(self._data, n))

    # ---- shape ---------------------------------------------------------------------------------------
    def transpose(self, *dims, **kw):
        if not dims:
            dims = self.dims[::-1]
        if Ellipsis in dims:
            i = dims.index(Ellipsis)
            rest = [d for d in self.dims if d not in dims]
            dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
        if set(dims) != set(self.dims) or len(dims) != len(self.dims):
            raise ValueError(f"{dims} must be a permuted list of {self.dims}")
        return DataArray(np.transpose(self._data, [self.dims.index(d) for d in dims]), self._coords, dims, self.name, self.attrs)

    def squeeze(self, dim=None, drop=False):
        if dim is None:
            dim = [d for d, n in self.sizes.items() if n == 1]
        if isinstance(dim, str):
            dim = [dim]
        out = self
        for d in dim:
            if out.sizes[d] != 1:
                raise ValueError(f"cannot squeeze dimension {d!r} of size {out.sizes[d]}")
            out = out.isel({d: 0}, drop=drop)
        return out

    def expand_dims(self, dim=None, axis=None, **kw):
        if isinstance(dim, str):
            dim = {dim: 1}
        elif dim is not None and not isinstance(dim, dict):
            dim = {d: 1 for d in dim}
        dim = dict(dim or {}, **kw)
        if len(dim) != 1:
            raise NotImplementedError("expand_dims with several dims")
        (d, v), = dim.items()
        if d in self.dims:
            raise ValueError(f"dimension {d!r} already exists")
        ax = 0 if axis is None else (axis if not isinstance(axis, (list, tuple)) else axis[0])
        coords = dict(self._coords)
        if isinstance(v, (int, np.integer)):
            n = int(v)
            if d in coords:  # a scalar coordinate becomes the 1-element index
                c = coords[d]
                coords[d] = _Coord((d,), c.data.reshape(1), c.attrs)
        else:
            c = _as_coord(d, v)
            n = c.data.shape[0]
            coords[d] = _Coord((d,), c.data, c.attrs)
        data = np.repeat(np.expand_dims(self._data, ax), n, axis=ax)
        dims = list(self.dims)
        dims.insert(ax, d)
        return DataArray(data, coords, dims, self.name, self.attrs)

    def rename(self, new=None, **kw):
        if new is None and not kw:
            return self.copy(deep=False)
        if isinstance(new, str):
            out = self.copy(deep=False)
            out.name = new
            return out
        m = dict(new or {}, **kw)
        dims = tuple(m.get(d, d) for d in self.dims)
        coords = {m.get(k, k): _Coord(tuple(m.get(d, d) for d in c.dims), c.data, c.attrs) for k, c in self._coords.items()}
        return DataArray(self._data, coords, dims, m.get(self.name, self.name), self.attrs)

    def drop_vars(self, names, errors="raise"):
        if isinstance(names, str):
            names = [names]
        coords = dict(self._coords)
        for n in names:
            if n not in coords:
                if errors == "raise":
                    raise ValueError(f"cannot drop {n!r}: not a coordinate of this array")
                continue
            del coords[n]
        return DataArray(self._data, coords, self.dims, self.name, self.attrs)

    def reset_coords(self, names=None, drop=False):
        if not drop:
            raise NotImplementedError
        coords = {k: c for k, c in self._coords.items() if c.dims == (k,)} if names is None else {k: c for k, c in self._coords.items() if k not in ([names] if isinstance(names, str) else names)}
        return DataArray(self._data, coords, self.dims, self.name, self.attrs)

    def assign_coords(self, coords=None, **kw):
        out = self.copy(deep=False)
        for k, v in dict(coords or {}, **kw).items():
            c = _as_coord(k, v)
            for d, n in zip(c.dims, c.data.shape):
                if d not in out.dims or out.sizes[d] != n:
                    raise ValueError(f"coordinate {k!r} does not fit dimension {d!r}")
            out._coords[k] = c
        return out

    def assign_attrs(self, *a, **kw):
        out = self.copy(deep=False)
        for d in a:
            out.attrs.update(d)
        out.attrs.update(kw)
        return out

    def dropna(self, dim, how="any", thresh=None):
        if thresh is not None:
            raise NotImplementedError
        ax = tuple(i for i, d in enumerate(self.dims) if d != dim)
        null = self.isnull()._data
        bad = null.any(axis=ax) if how == "any" else null.all(axis=ax)
        return self.isel({dim: np.nonzero(~bad)[0]})

    def sortby(self, key, ascending=True):
        if not isinstance(key, DataArray) or key.ndim != 1:
            raise NotImplementedError("sortby with a non 1-D key")
        order = np.argsort(key._data, kind="stable")
        if not ascending:
            order = order[::-1]
        return self.isel({key.dims[0]: order})

    def equals(self, other):
        if not isinstance(other, DataArray) or self.dims != other.dims or self.shape != other.shape:
            return False
        a, b = self._data, other._data
        same = _index_equal(a, b) if a.dtype.kind not in "fc" else bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
        if not same:
            return False
        if set(self._coords) != set(other._coords):
            return False
        return all(self._coords[k].dims == other._coords[k].dims and _index_equal(self._coords[k].data, other._coords[k].data) for k in self._coords)

    # ---- reductions ------------------------------------------------------------------------------------
    def _reduce(self, f, nf, dim, skipna, keep_attrs=False):
        if dim is None or dim is Ellipsis:
            axes = tuple(range(self.ndim))
        else:
            dd = [dim] if isinstance(dim, str) else list(dim)
            for d in dd:
                if d not in self.dims:
                    raise ValueError(f"{d!r} not found in array dimensions {self.dims}")
            axes = tuple(self.dims.index(d) for d in dd)
        fn = nf if _skipna_default(self._data, skipna) and nf is not None else f
        import warnings

        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore", RuntimeWarning)
            out = fn(self._data, axis=axes)
        dims = tuple(d for i, d in enumerate(self.dims) if i not in axes)
        return self._new(out, dims)

    def mean(self, dim=None, skipna=None, **kw): return self._reduce(np.mean, np.nanmean, dim, skipna)
    def sum(self, dim=None, skipna=None, **kw): return self._reduce(np.sum, np.nansum, dim, skipna)
    def min(self, dim=None, skipna=None, **kw): return self._reduce(np.min, np.nanmin, dim, skipna)
    def max(self, dim=None, skipna=None, **kw): return self._reduce(np.max, np.nanmax, dim, skipna)
    def median(self, dim=None, skipna=None, **kw): return self._reduce(np.median, np.nanmedian, dim, skipna)

    def all(self, dim=None, axis=None, **kw):
        if axis is not None:
            raise NotImplementedError
        return self._reduce(np.all, None, dim, False)

    def any(self, dim=None, axis=None, **kw):
        if axis is not None:
            raise NotImplementedError
        return self._reduce(np.any, None, dim, False)

    def idxmin(self, dim, skipna=None, fill_value=np.nan):
        """coordinate label of the minimum along ``dim``; all-NaN slices give ``fill_value`` (float result)."""
        ax = self.dims.index(dim)
        a = self._data
        index = self._coords[dim].data
        if a.dtype.kind in "fc" and _skipna_default(a, skipna):
            allnan = np.isnan(a).all(axis=ax)
            pos = np.nanargmin(np.where(np.isnan(a) & np.expand_dims(allnan, ax), 0.0, a), axis=ax)
            lab = index[pos]
            if allnan.any():
                lab = np.where(allnan, fill_value, lab.astype(np.float64) if index.dtype.kind in "iu" else lab)
        else:
            lab = index[np.argmin(a, axis=ax)]
        dims = tuple(d for d in self.dims if d != dim)
        return DataArray(lab, {k: c for k, c in self._coords.items() if dim not in c.dims}, dims, dim)

    # ---- interpolation / reindexing / windows ------------------------------------------------------------
    def interp(self, coords=None, method="linear", assume_sorted=False, kwargs=None, **kw):
        from scipy.interpolate import interp1d

        ind = dict(coords or {}, **kw)
        if len(ind) != 1 or method not in ("linear", "nearest"):
            raise NotImplementedError("interp: one dim, linear / nearest only")
        (d, new), = ind.items()
        x = self._coords[d].data
        new_da = new if isinstance(new, DataArray) else None
        xn = np.asarray(new_da._data if new_da is not None else new)
        if x.dtype.kind == "M":
            x0 = x.astype("datetime64[ns]").min()
            xf = (x.astype("datetime64[ns]") - x0).astype(np.float64)
            xnf = (xn.astype("datetime64[ns]") - x0).astype(np.float64)
        else:
            xf, xnf = x.astype(np.float64), xn.astype(np.float64)
        order = np.argsort(xf, kind="stable")
        ax = self.dims.index(d)
        y = np.take(self._data.astype(np.float64), order, axis=ax)
        f = interp1d(xf[order], y, kind=method, axis=ax, bounds_error=False, **dict(kwargs or {}))  # as xarray's missing.py does
        out = f(xnf)
        if new_da is not None and new_da.ndim == 1:
            nd = new_da.dims[0]
            dims = tuple(nd if x_ == d else x_ for x_ in self.dims)
            co = {k: c for k, c in self._coords.items() if d not in c.dims}
            for k, c in new_da._coords.items():
                co.setdefault(k, c)
            co[d] = _Coord((nd,), xn)
            return DataArray(out, co, dims, self.name, self.attrs)
        if xn.ndim == 0:
            dims = tuple(x_ for x_ in self.dims if x_ != d)
            co = {k: c for k, c in self._coords.items() if d not in c.dims}
            if new_da is not None:
                for k, c in new_da._coords.items():
                    co.setdefault(k, c)
            co[d] = _Coord((), xn)
            return DataArray(out, co, dims, self.name, self.attrs)
        co = {k: c for k, c in self._coords.items() if d not in c.dims}
        co[d] = _Coord((d,), xn)
        return DataArray(out, co, self.dims, self.name, self.attrs)

    def reindex(self, indexers=None, method=None, **kw):
        ind = dict(indexers or {}, **kw)
        out = self
        for d, new in ind.items():
            newv = np.asarray(new._data if isinstance(new, DataArray) else new)
            old = out._coords[d].data
            if method == "ffill":
                if np.any(np.diff(old) <= 0):
                    raise NotImplementedError("ffill reindex on a non-monotonic index")
                pos = np.searchsorted(old, newv, side="right") - 1
                miss = pos < 0
            elif method is None:
                pos = np.array([np.nonzero(old == v)[0][0] if np.any(old == v) else -1 for v in newv])
                miss = pos < 0
            else:
                raise NotImplementedError(method)
            taken = out.isel({d: np.where(miss, 0, pos)})
            if miss.any():
                taken = taken.astype(np.float64)
                key = [slice(None)] * taken.ndim
                key[taken.dims.index(d)] = miss
                taken._data[tuple(key)] = np.nan
            taken._coords[d] = _Coord((d,), newv)
            out = taken
        return out

    def coarsen(self, dim=None, boundary="exact", side="left", coord_func="mean", **kw):
        return _Coarsen(self, dict(dim or {}, **kw), boundary, side, coord_func)

    def pipe(self, func, *args, **kw):
        return func(self, *args, **kw)

    def resample(self, indexer=None, skipna=None, **kw):
        """1-D time series only; pandas resampling (origin="start_day", closed / label "left"), like xarray's wrapper."""
        import pandas as pd

        (dim, freq), = dict(indexer or {}, **kw).items()
        if self.dims != (dim,):
            raise NotImplementedError("resample: 1-D series along the resampled dimension only")
        series = pd.Series(np.asarray(self._data), index=pd.DatetimeIndex(self._coords[dim].data))
        owner = self

        class _Resampled:
            def first(self_inner, **k):
                r = series.resample(freq).first()
                out = DataArray(r.to_numpy(), {dim: r.index.to_numpy()}, (dim,), owner.name)
                out.indexes = {dim: r.index}
                return out

        return _Resampled()

    def chunk(self, chunks=None, **kw):
        return self

    def reindex_like(self, other, method=None):
        return self.reindex({d: other._coords[d].data for d in self.dims if d in other._coords and other._coords[d].dims == (d,)}, method=method)

    def to_dataset(self, name=None):
        n = name or self.name
        if n is None:
            raise ValueError("unable to convert unnamed DataArray to a Dataset without providing an explicit name")
        ds = Dataset()
        for k, c in self._coords.items():
            ds._coords[k] = c
        ds._vars[n] = _Coord(self.dims, self._data, self.attrs)
        return ds

    def to_numpy(self):
        return self._data


def _isnull_arr(a):
    if a.dtype.kind in "fc":
        return np.isnan(a)
    return np.zeros(a.shape, bool)


class _Loc:
    def __init__(self, da):
        self._da = da

    def _pos(self, key):
        if not isinstance(key, dict):
            if not isinstance(key, tuple):
                key = (key,)
            key = dict(zip(self._da.dims, key))
        return {d: self._da._label_to_pos(d, k) for d, k in key.items()}

    def __getitem__(self, key):
        return self._da.isel(self._pos(key))

    def __setitem__(self, key, value):
        self._da._set(self._pos(key), value)


class _Coarsen:
    def __init__(self, da, windows, boundary, side, coord_func):
        if side != "left":
            raise NotImplementedError
        self.da, self.windows, self.boundary, self.coord_func = da, windows, boundary, coord_func

    def _blocks(self, data, dims):
        """reshape ``data`` so every coarsened dim d becomes (n_blocks, window); returns array, reduce axes"""
        shape, axes = [], []
        data = data.astype(np.float64) if data.dtype.kind in "iub" and self.boundary == "pad" else data
        for i, d in enumerate(dims):
            n = data.shape[i]
            if d in self.windows:
                w = int(self.windows[d])
                rem = n % w
                if rem:
                    if self.boundary == "pad":
                        pad = [(0, 0)] * data.ndim
                        pad[i] = (0, w - rem)
                        if data.dtype.kind == "M":
                            data = np.concatenate([data, np.full([w - rem if j == i else s for j, s in enumerate(data.shape)], np.datetime64("NaT"), dtype=data.dtype)], axis=i)
                        else:
                            data = np.pad(data, pad, mode="constant", constant_values=np.nan)
                    elif self.boundary == "trim":
                        data = np.take(data, np.arange(n - rem), axis=i)
                    else:
                        raise ValueError(f"Could not coarsen a dimension of size {n} with window {w}")
                shape += [data.shape[i] // w, w]
                axes.append(len(shape) - 1)
            else:
                shape.append(n)
        return data.reshape(shape), tuple(axes)

    def _reduce(self, fn_nan, fn):
        import warnings

        da = self.da
        blocks, axes = self._blocks(da._data, da.dims)
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore", RuntimeWarning)
            out = (fn_nan if blocks.dtype.kind in "fc" else fn)(blocks, axis=axes)
        coords = {}
        for k, c in da._coords.items():
            if not any(d in self.windows for d in c.dims):
                coords[k] = c
                continue
            cd = c.data
            if cd.dtype.kind == "M":  # mean of datetimes: through int64 ns, NaT padding skipped
                b, ax = self._blocks(cd.astype("datetime64[ns]"), c.dims)
                i8 = b.astype(np.int64).astype(np.float64)
                i8[np.isnat(b)] = np.nan
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", RuntimeWarning)
                    m = np.nanmean(i8, axis=ax)
                coords[k] = _Coord(c.dims, m.astype(np.int64).astype("datetime64[ns]"), c.attrs)
            else:
                b, ax = self._blocks(cd, c.dims)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", RuntimeWarning)
                    coords[k] = _Coord(c.dims, np.nanmean(b, axis=ax), c.attrs)
        return DataArray(out, coords, da.dims, da.name)

    def mean(self, skipna=None, **kw):
        return self._reduce(np.nanmean if skipna in (None, True) else np.mean, np.mean)

    def min(self, skipna=None, **kw):
        return self._reduce(np.nanmin if skipna in (None, True) else np.min, np.min)

    def max(self, skipna=None, **kw):
        return self._reduce(np.nanmax if skipna in (None, True) else np.max, np.max)

    def sum(self, skipna=None, **kw):
        return self._reduce(np.nansum if skipna in (None, True) else np.sum, np.sum)


class _VarsView:
    def __init__(self, ds, names):
        self._ds, self._names = ds, names

    def __contains__(self, k):
        return k in self._names()

    def __iter__(self):
        return iter(list(self._names()))

    def __len__(self):
        return len(list(self._names()))

    def keys(self):
        return list(self._names())

    def items(self):
        return [(k, self._ds[k]) for k in self._names()]

    def values(self):
        return [self._ds[k] for k in self._names()]

    def __getitem__(self, k):
        if k not in self:
            raise KeyError(k)
        return self._ds[k]


class Dataset:
    __hash__ = None

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = {}
        self._coords = {}
        self.attrs = dict(attrs or {})
        for k, v in (coords or {}).items():
            self._coords[k] = _as_coord(k, v)
        for k, v in (data_vars or {}).items():
            self[k] = v

    # ---- mapping surface ----------------------------------------------------------------------------------
    @property
    def sizes(self):
        s = {}
        for c in list(self._coords.values()) + list(self._vars.values()):
            for d, n in zip(c.dims, c.data.shape):
                if s.setdefault(d, n) != n:
                    raise ValueError(f"conflicting sizes for {d!r}")
        return s

    dims = sizes

    @property
    def data_vars(self):
        return _VarsView(self, lambda: self._vars.keys())

    @property
    def variables(self):
        return _VarsView(self, lambda: list(self._coords.keys()) + list(self._vars.keys()))

    @property
    def coords(self):
        return _CoordsView(self)

    def _coord_da(self, k):
        c = self._coords[k]
        sub = {n: cc for n, cc in self._coords.items() if set(cc.dims) <= set(c.dims)}
        return DataArray(c.data, sub, c.dims, name=k, attrs=c.attrs)

    def __contains__(self, k):
        return k in self._vars or k in self._coords

    def __iter__(self):
        return iter(list(self._vars))

    def __len__(self):
        return len(self._vars)

    def keys(self):
        return list(self._vars)

    def items(self):
        return [(k, self[k]) for k in self._vars]

    def __getitem__(self, k):
        if isinstance(k, (list, tuple)):
            out = Dataset(attrs=self.attrs)
            for n in k:
                out[n] = self[n]
            return out
        if k in self._vars:
            v = self._vars[k]
            co = {n: c for n, c in self._coords.items() if set(c.dims) <= set(v.dims)}
            da = DataArray(v.data, co, v.dims, k)
            da.attrs = v.attrs  # shared on purpose: ds["x"].attrs.update(...) must stick
            return da
        if k in self._coords:
            da = self._coord_da(k)
            da.attrs = self._coords[k].attrs
            return da
        raise KeyError(f"No variable named {k!r}. Variables on the dataset include {list(self._vars)}")

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        d = self.__dict__
        if k in d.get("_vars", {}) or k in d.get("_coords", {}):
            return self[k]
        raise AttributeError(f"xrlite.Dataset has no attribute {k!r}")

    def _check_fits(self, name, c):
        s = self.sizes
        for d, n in zip(c.dims, c.data.shape):
            if d in s and s[d] != n:
                raise ValueError(f"cannot add {name!r}: dimension {d!r} has size {n}, dataset has {s[d]}")

    def __setitem__(self, k, v):
        if isinstance(v, DataArray):
            for n, c in v._coords.items():
                if n in self._coords:
                    o = self._coords[n]
                    if c.dims == (n,) and not _index_equal(o.data, c.data):
                        raise NotImplementedError(f"assignment of {k!r} needs re-alignment on {n!r}")
                elif n != k:
                    self._check_fits(n, c)
                    self._coords[n] = c
            c = _Coord(v.dims, v._data, v.attrs)
        elif isinstance(v, tuple):
            c = _as_coord(k, v)
        else:
            a = np.asarray(v)
            if a.ndim != 0:
                if a.ndim == 1 and k in self.sizes:
                    c = _Coord((k,), a)
                else:
                    raise ValueError(f"cannot set variable {k!r} with {a.ndim}-dimensional data without explicit dimension names")
            else:
                c = _Coord((), a)
        self._check_fits(k, c)
        if k in self._coords:
            self._coords[k] = c
        else:
            self._vars[k] = c

    def __delitem__(self, k):
        if k in self._vars:
            del self._vars[k]
        else:
            del self._coords[k]

    def __repr__(self):
        return f"<xrlite.Dataset {self.sizes} vars={list(self._vars)} coords={list(self._coords)}>"

    def copy(self, deep=False):
        out = Dataset(attrs=dict(self.attrs))
        f = (lambda c: _Coord(c.dims, c.data.copy(), c.attrs)) if deep else (lambda c: _Coord(c.dims, c.data, c.attrs))
        out._coords = {k: f(c) for k, c in self._coords.items()}
        out._vars = {k: f(c) for k, c in self._vars.items()}
        return out

    def compute(self):
        return self

    load = compute

    # ---- indexing ----------------------------------------------------------------------------------------------
    def _map_indexed(self, how, ind, **kw):
        out = Dataset(attrs=dict(self.attrs))
        probe = {}
        for k in list(self._coords) + list(self._vars):
            da = self[k] if k in self._vars else self._coord_da(k)
            sub = {d: v for d, v in ind.items() if d in da.dims}
            probe[k] = getattr(da, how)(sub, **kw) if sub else da
        for k in self._coords:
            r = probe[k]
            out._coords[k] = _Coord(r.dims, r._data, self._coords[k].attrs)
        for k in self._vars:
            r = probe[k]
            out._vars[k] = _Coord(r.dims, r._data, self._vars[k].attrs)
            for n, c in r._coords.items():  # scalar coords created by the selection
                if n not in out._coords:
                    out._coords[n] = c
        if kw.get("drop"):
            for d, v in ind.items():
                if np.ndim(v) == 0 and not isinstance(v, slice) and d in out._coords:
                    del out._coords[d]
        return out

    def isel(self, indexers=None, drop=False, **kw):
        return self._map_indexed("isel", dict(indexers or {}, **kw), drop=drop)

    def sel(self, indexers=None, drop=False, method=None, **kw):
        ind = dict(indexers or {}, **kw)
        pos = {}
        for d, lab in ind.items():
            if d not in self._coords:
                raise KeyError(f"no index for {d!r}")
            pos[d] = self._coord_da(d)._label_to_pos(d, lab) if self._coords[d].dims == (d,) else None
            if pos[d] is None:
                raise KeyError(d)
        return self._map_indexed("isel", pos, drop=drop)

    def drop_vars(self, names, errors="raise"):
        if isinstance(names, str):
            names = [names]
        out = self.copy()
        for n in names:
            if n in out._vars:
                del out._vars[n]
            elif n in out._coords:
                del out._coords[n]
            elif errors == "raise":
                raise ValueError(f"cannot drop {n!r}: not in the dataset")
        return out

    def drop_dims(self, dims, errors="raise"):
        if isinstance(dims, str):
            dims = [dims]
        out = self.copy()
        for store in (out._vars, out._coords):
            for k in [k for k, c in store.items() if set(c.dims) & set(dims)]:
                del store[k]
        return out

    def rename(self, m=None, **kw):
        m = dict(m or {}, **kw)
        out = Dataset(attrs=dict(self.attrs))
        r = lambda c: _Coord(tuple(m.get(d, d) for d in c.dims), c.data, c.attrs)  # noqa: E731
        out._coords = {m.get(k, k): r(c) for k, c in self._coords.items()}
        out._vars = {m.get(k, k): r(c) for k, c in self._vars.items()}
        return out

    rename_vars = rename

    def assign(self, variables=None, **kw):
        out = self.copy()
        for k, v in dict(variables or {}, **kw).items():
            out[k] = v
        return out

    def assign_coords(self, coords=None, **kw):
        out = self.copy()
        for k, v in dict(coords or {}, **kw).items():
            c = _as_coord(k, v)
            out._check_fits(k, c)
            out._vars.pop(k, None)
            out._coords[k] = c
        return out

    def assign_attrs(self, *a, **kw):
        out = self.copy()
        for d in a:
            out.attrs.update(d)
        out.attrs.update(kw)
        return out

    def transpose(self, *dims):
        out = self.copy()
        for k, c in out._vars.items():
            order = [d for d in dims if d in c.dims] + [d for d in c.dims if d not in dims]
            out._vars[k] = _Coord(order, np.transpose(c.data, [c.dims.index(d) for d in order]), c.attrs)
        return out

    def merge(self, other, join="outer", compat="no_conflicts", **kw):
        return merge([self, other], join=join, compat=compat)

    def equals(self, other):
        return set(self._vars) == set(other._vars) and all(self[k].equals(other[k]) for k in self._vars)


def merge(objects, join="outer", compat="no_conflicts", **kw):
    """merge of Datasets / named DataArrays whose indexes agree, or are disjoint pieces of one outer index
    along ``channel`` (the only outer join the calibrate path performs, calibrate_ek.py:39-52)."""
    dss = [o.to_dataset() if isinstance(o, DataArray) else o for o in objects]
    idx = {}
    need_outer = set()
    for ds in dss:
        for k, c in ds._coords.items():
            if c.dims == (k,):
                if k in idx and not _index_equal(idx[k], c.data):
                    need_outer.add(k)
                idx.setdefault(k, c.data)
    if need_outer:
        if join != "outer" or need_outer != {"channel"}:
            raise NotImplementedError(f"merge needing an outer join on {need_outer}")
        return _concat_disjoint(dss, "channel")
    out = Dataset()
    for ds in dss:
        for k, c in ds._coords.items():
            if k in out._coords and not (out._coords[k].dims == c.dims and _index_equal(out._coords[k].data, c.data)):
                if c.dims == ():
                    continue
                raise ValueError(f"conflicting coordinate {k!r}")
            out._coords.setdefault(k, c)
        for k, c in ds._vars.items():
            if k in out._vars:
                o = out._vars[k]
                a, b = o.data, c.data
                ok = o.dims == c.dims and a.shape == b.shape and bool(np.all((a == b) | _isnull_arr(a) | _isnull_arr(b)))
                if not ok:
                    raise ValueError(f"conflicting values for variable {k!r}")
                out._vars[k] = _Coord(o.dims, np.where(_isnull_arr(a), b, a), o.attrs)
            else:
                out._vars[k] = c
        out.attrs.update(ds.attrs) if not out.attrs else None
    return out


def _concat_disjoint(dss, dim):
    labels = []
    for ds in dss:
        for v in ds._coords[dim].data:
            if v in labels:
                raise NotImplementedError("outer merge with overlapping labels")
            labels.append(v)
    order = sorted(range(len(labels)), key=lambda i: labels[i])  # an outer join sorts the union
    out = Dataset()
    names = []
    for ds in dss:
        for k in ds._vars:
            if k not in names:
                names.append(k)
    for k, c in dss[0]._coords.items():
        if dim not in c.dims:
            out._coords[k] = c
    out._coords[dim] = _Coord((dim,), np.array(labels, dtype=dss[0]._coords[dim].data.dtype)[order])
    for k in names:
        parts = [ds._vars[k] for ds in dss]
        if dim not in parts[0].dims:
            out._vars[k] = parts[0]
            continue
        ax = parts[0].dims.index(dim)
        out._vars[k] = _Coord(parts[0].dims, np.take(np.concatenate([p.data for p in parts], axis=ax), order, axis=ax), parts[0].attrs)
    return out


def where(cond, x, y, keep_attrs=None):
    return DataArray._apply(lambda c, a, b: np.where(c, a, b), cond, x, y)


def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=((),), vectorize=False, dask=None,
                output_dtypes=None, **kw):
    """the vectorize=True form: labelled arguments with identical loop dims, core dims moved last, python loop over the
    remaining dims, output core dims appended (ek80_complex.py:352-360, clean/api.py:253-263, 348-357)."""
    if not args or not vectorize or input_core_dims is None or len(output_core_dims) != 1 or len(input_core_dims) != len(args):
        raise NotImplementedError("apply_ufunc form not supported by xrlite")
    a = args[0]
    loop = [d for d in a.dims if d not in input_core_dims[0]]
    arrs = []
    for x, core in zip(args, input_core_dims):
        if [d for d in x.dims if d not in core] != loop and sorted(d for d in x.dims if d not in core) != sorted(loop):
            raise NotImplementedError("apply_ufunc: arguments with different loop dims")
        arrs.append(x.transpose(*loop, *core)._data)
    lshape = arrs[0].shape[: len(loop)]
    if any(r.shape[: len(loop)] != lshape for r in arrs):
        raise ValueError("apply_ufunc: loop dimensions differ in size")
    out = None
    for ix in np.ndindex(*lshape):
        r = np.asarray(func(*[r[ix] for r in arrs]))
        if out is None:
            out = np.empty(lshape + r.shape, dtype=(output_dtypes[0] if output_dtypes else r.dtype))
        out[ix] = r
    dims = tuple(loop) + tuple(output_core_dims[0])
    return DataArray(out, {k: c for k, c in a._coords.items() if set(c.dims) <= set(dims)}, dims, a.name)


def full_like(other, fill_value, dtype=None):
    return other._new(np.full(other.shape, fill_value, dtype=dtype if dtype is not None else other.dtype))


def zeros_like(other, dtype=None):
    return full_like(other, 0, dtype)


def ones_like(other, dtype=None):
    return full_like(other, 1, dtype)


def broadcast(*arrays):
    """xr.broadcast for DataArrays: every array expanded to the union of the dims (first-appearance order)."""
    aligned, dims = DataArray._align(list(arrays))
    sizes = {}
    for a in aligned:
        sizes.update(a.sizes)
    coords = DataArray._merge_coords(aligned, dims)
    out = []
    for a in aligned:
        data = np.broadcast_to(a._broadcast_data(dims), [sizes[d] for d in dims]).copy()
        out.append(DataArray(data, coords, dims, a.name, a.attrs))
    return tuple(out)


def concat(objs, dim, **kw):
    """DataArrays that each carry `dim` as a scalar coordinate (the result of isel(dim=i)) stacked along a new leading
    dimension `dim` (clean/utils.py:181, 313)."""
    objs = list(objs)
    first = objs[0]
    if not all(isinstance(o, DataArray) and o.dims == first.dims and dim not in o.dims and dim in o._coords for o in objs):
        raise NotImplementedError("xrlite.concat: only scalar-coordinate stacking of equally shaped DataArrays")
    data = np.stack([o._data for o in objs])
    coords = {k: c for k, c in first._coords.items() if k != dim and c.dims}
    coords[dim] = _Coord((dim,), np.array([np.asarray(o._coords[dim].data)[()] for o in objs]))
    return DataArray(data, coords, (dim,) + tuple(first.dims), first.name)


def set_options(**kw):
    class _Ctx:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    return _Ctx()
