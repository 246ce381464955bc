"""
TEST INFRASTRUCTURE ONLY -- a small stand-in for the part of xarray that the reference uses.

xarray (environment.yml pins 2022.12.0) is absent from the build container, so the reference's
xarray-bound functions (``integ_geopot``, ``load_delta``, ``vert_interp_delta``, ``interp_logp_4d``,
``regrid_lat_lon``, ``filter_data``, ``pgw_for_era5`` ...) cannot be run as they are.  This module
restates xarray's *published semantics* for exactly the calls those functions make -- labelled
dimensions, broadcasting by dimension name, label/position selection incl. vectorised (pointwise)
indexers, ``where``/``diff``/``expand_dims``/``transpose``/``concat``/``reindex``/``interp``
(scipy ``interp1d``, as xarray does), ``apply_ufunc(vectorize=True)``, NaN-skipping reductions, CF
time/mask decoding in ``open_dataset`` -- so that ``oracle/make_golden_glue.py`` can execute the
UNMODIFIED reference code from /root/reference and store what it returns as golden vectors
(``tests/golden/reference_glue.npz``).  It is installed as ``sys.modules['xarray']`` by that script
only; the product never imports it, and ``tests/test_xrlite.py`` checks the semantics it claims.

Deliberate limits: coordinates of operands are assumed to agree (xarray would inner-join; here a
mismatch of an index coordinate raises), NetCDF-3 files only (scipy), standard calendars only.
"""
import os
from datetime import datetime

import numpy as np
import pandas as pd
from scipy.interpolate import interp1d
from scipy.io import netcdf_file


class Variable:
    """dims + data + attrs; shared between a Dataset and the DataArrays handed out from it."""
    __slots__ = ("dims", "data", "attrs")

    def __init__(self, dims, data, attrs=None):
        self.dims = tuple(dims)
        self.data = _as_data(data)
        self.attrs = dict(attrs or {})
        assert self.data.ndim == len(self.dims), (self.dims, self.data.shape)

    def copy(self, deep=True):
        return Variable(self.dims, self.data.copy() if deep else self.data, self.attrs)


def _as_data(v):
    if isinstance(v, DataArray):
        v = v.variable.data
    if isinstance(v, datetime):
        v = np.datetime64(v, "ns")
    if isinstance(v, (pd.DatetimeIndex, pd.Index)):
        v = v.values
    a = np.asarray(v)
    if a.dtype.kind == "M" and a.dtype != np.dtype("datetime64[ns]"):
        a = a.astype("datetime64[ns]")
    if a.dtype == object and a.size and isinstance(a.flat[0], datetime):
        a = a.astype("datetime64[ns]")
    return a


def _scalar(v):
    if isinstance(v, DataArray):
        assert v.ndim == 0, "scalar label expected"
        return v.values[()]
    return v


class _Coords:
    """``da.coords``: mapping view, usable as ``coords=`` of the DataArray constructor."""

    def __init__(self, owner):
        self._o = owner

    def __contains__(self, k):
        return k in self._o._coords

    def __iter__(self):
        return iter(self._o._coords)

    def keys(self):
        return self._o._coords.keys()

    def items(self):
        return [(k, self._o[k]) for k in self._o._coords]

    def __getitem__(self, k):
        return self._o[k]

    def __len__(self):
        return len(self._o._coords)


def _merge_coords(objs, dims):
    """Union of the coordinates of ``objs`` restricted to ``dims``; index coordinates must agree."""
    out = {}
    for o in objs:
        for k, v in o._coords.items():
            if not set(v.dims) <= set(dims):
                continue
            if k in out:
                if v.dims == (k,) and out[k].dims == (k,):
                    a, b = out[k].data, v.data
                    if a.shape != b.shape or not np.array_equal(a, b):
                        raise ValueError("xrlite: index coordinate %r differs between operands" % k)
                elif out[k].dims == () and v.dims == ():
                    if not np.array_equal(out[k].data, v.data, equal_nan=False):
                        out[k] = None              # conflicting scalar coordinates are dropped
                continue
            out[k] = v
    return {k: v for k, v in out.items() if v is not None}


def _broadcast(arrs):
    """dims in order of first appearance; data of every DataArray expanded to them (views)."""
    dims = []
    for a in arrs:
        if isinstance(a, DataArray):
            for d in a.dims:
                if d not in dims:
                    dims.append(d)
    sizes = {}
    for a in arrs:
        if isinstance(a, DataArray):
            for d, n in zip(a.dims, a.shape):
                if sizes.setdefault(d, n) != n:
                    raise ValueError("xrlite: size of dimension %r differs (%d, %d)" % (d, sizes[d], n))
    out = []
    for a in arrs:
        if isinstance(a, DataArray):
            order = [a.dims.index(d) for d in dims if d in a.dims]
            x = a.variable.data.transpose(order) if order else a.variable.data
            shape = [sizes[d] if d in a.dims else 1 for d in dims]
            out.append(x.reshape(shape))
        else:
            out.append(a)
    return tuple(dims), out


class DataArray:
    __array_priority__ = 60

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None, variable=None, _coords=None):
        if variable is not None:
            self.variable, self._coords, self.name = variable, dict(_coords or {}), name
            return
        data = _as_data(data)
        cvars = {}
        if isinstance(coords, _Coords):
            src = coords._o
            if dims is None:
                dims = [k for k in src.dims]
            cvars = dict(src._coords)
        elif isinstance(coords, dict):
            if dims is None:
                dims = list(coords.keys())[:data.ndim]
            for k, v in coords.items():
                if isinstance(v, DataArray):
                    cvars[k] = v.variable
                else:
                    v = _as_data(v)
                    cvars[k] = Variable((k,) if v.ndim else (), v)
        if dims is None:
            dims = ["dim_%d" % i for i in range(data.ndim)]
        self.variable = Variable(dims, data, attrs)
        self._coords = {k: v for k, v in cvars.items() if set(v.dims) <= set(dims)}
        self.name = name

    # ---- basic properties
    dims = property(lambda s: s.variable.dims)
    shape = property(lambda s: s.variable.data.shape)
    ndim = property(lambda s: s.variable.data.ndim)
    dtype = property(lambda s: s.variable.data.dtype)
    size = property(lambda s: s.variable.data.size)
    attrs = property(lambda s: s.variable.attrs, lambda s, v: setattr(s.variable, "attrs", dict(v)))
    coords = property(lambda s: _Coords(s))

    @property
    def values(self):
        return self.variable.data

    @values.setter
    def values(self, v):
        v = _as_data(v)
        if v.shape != self.variable.data.shape:
            raise ValueError("replacement data must match the Variable's shape")
        self.variable.data = v

    data = values

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.variable.data, dtype=dtype)

    def __len__(self):
        return self.shape[0]

    def __bool__(self):
        return bool(self.variable.data)

    def __float__(self):
        return float(self.variable.data)

    def __int__(self):
        return int(self.variable.data)

    def __index__(self):
        return int(self.variable.data)

    def __repr__(self):
        return "<xrlite.DataArray %s %s>\n%r" % (self.name, dict(zip(self.dims, self.shape)), self.variable.data)

    def __getattr__(self, name):
        if name.startswith("_") or name in ("variable", "name"):
            raise AttributeError(name)
        if name in self._coords:
            return self[name]
        raise AttributeError(name)

    def _new(self, data, dims=None, coords=None, name=None):
        dims = self.dims if dims is None else tuple(dims)
        coords = self._coords if coords is None else coords
        return DataArray(variable=Variable(dims, data, self.variable.attrs),
                         _coords={k: v for k, v in coords.items() if set(v.dims) <= set(dims)},
                         name=self.name if name is None else name)

    def copy(self, deep=True):
        return DataArray(variable=self.variable.copy(deep),
                         _coords={k: v.copy(deep) for k, v in self._coords.items()}, name=self.name)

    # ---- coordinates and positional access
    def __getitem__(self, key):
        if isinstance(key, str):
            v = self._coords[key]
            return DataArray(variable=v, name=key,
                             _coords={k: c for k, c in self._coords.items() if set(c.dims) <= set(v.dims)})
        if isinstance(key, (int, np.integer)):
            return self.isel({self.dims[0]: int(key)})
        if isinstance(key, dict):
            return self.isel(key)
        raise TypeError("xrlite: unsupported key %r" % (key,))

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise TypeError("xrlite: only coordinate assignment is supported")
        data = _as_data(value)
        if isinstance(value, DataArray):
            self._coords[key] = Variable(value.dims, data, value.variable.attrs)
        else:
            self._coords[key] = Variable((key,) if data.ndim else (), data)
        assert set(self._coords[key].dims) <= set(self.dims)

    def __delitem__(self, key):
        del self._coords[key]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __reversed__(self):
        for i in range(len(self) - 1, -1, -1):
            yield self[i]

    # ---- arithmetic
    def _binary(self, other, f, reflexive=False):
        if isinstance(other, (Dataset,)):
            return NotImplemented
        if isinstance(other, DataArray):
            dims, (a, b) = _broadcast([self, other])
            coords = _merge_coords([self, other], dims)
        else:
            dims, a, b, coords = self.dims, self.variable.data, other, self._coords   # positional
        with np.errstate(all="ignore"):
            r = f(b, a) if reflexive else f(a, b)
        out = DataArray(variable=Variable(dims, r), _coords=coords)
        return out

    def _inplace(self, other, f):
        # xarray's Variable._inplace_binary_op applies operator.iadd to the arrays themselves:
        # the dtype of the left operand is KEPT (float32 += float64 stays float32)
        if isinstance(other, DataArray):
            if not set(other.dims) <= set(self.dims):
                raise ValueError("dimensions cannot change for in-place operations")
            _merge_coords([self, other], self.dims)
            dims, (a, b) = _broadcast([self, other])
            assert dims == self.dims
        else:
            b = other
        with np.errstate(all="ignore"):
            f(self.variable.data, b, out=self.variable.data, casting="same_kind")
        return self

    def _unary(self, f):
        with np.errstate(all="ignore"):
            return self._new(f(self.variable.data))

    __add__ = lambda s, o: s._binary(o, np.add)
    __radd__ = lambda s, o: s._binary(o, np.add, True)
    __sub__ = lambda s, o: s._binary(o, np.subtract)
    __rsub__ = lambda s, o: s._binary(o, np.subtract, True)
    __mul__ = lambda s, o: s._binary(o, np.multiply)
    __rmul__ = lambda s, o: s._binary(o, np.multiply, True)
    __truediv__ = lambda s, o: s._binary(o, np.true_divide)
    __rtruediv__ = lambda s, o: s._binary(o, np.true_divide, True)
    __pow__ = lambda s, o: s._binary(o, np.power)
    __and__ = lambda s, o: s._binary(o, np.logical_and)
    __or__ = lambda s, o: s._binary(o, np.logical_or)
    __lt__ = lambda s, o: s._binary(o, np.less)
    __le__ = lambda s, o: s._binary(o, np.less_equal)
    __gt__ = lambda s, o: s._binary(o, np.greater)
    __ge__ = lambda s, o: s._binary(o, np.greater_equal)
    __eq__ = lambda s, o: s._binary(o, np.equal)
    __ne__ = lambda s, o: s._binary(o, np.not_equal)
    __hash__ = None
    __neg__ = lambda s: s._unary(np.negative)
    __abs__ = lambda s: s._unary(np.abs)
    __iadd__ = lambda s, o: s._inplace(o, np.add)
    __isub__ = lambda s, o: s._inplace(o, np.subtract)
    __imul__ = lambda s, o: s._inplace(o, np.multiply)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or ufunc.signature is not None:
            return NotImplemented
        out = kwargs.pop("out", None)
        if out is not None and any(isinstance(o, DataArray) for o in out):
            raise NotImplementedError("xarray objects are not supported in `out`")
        das = [x for x in inputs if isinstance(x, DataArray)]
        if all(isinstance(x, DataArray) or np.ndim(x) == 0 for x in inputs):
            dims, arrs = _broadcast(list(inputs))
            coords = _merge_coords(das, dims)
        else:                                   # mixed with plain arrays: positional
            dims, coords = das[0].dims, das[0]._coords
            arrs = [x.variable.data if isinstance(x, DataArray) else x for x in inputs]
        with np.errstate(all="ignore"):
            r = ufunc(*arrs, **kwargs) if out is None else ufunc(*arrs, out=out, **kwargs)
        return DataArray(variable=Variable(dims, r), _coords=coords)

    # ---- reductions (skipna=True for floats, as in xarray)
    def _reduce(self, fnan, fplain, dim=None, axis=None, out=None, keepdims=False, **kw):
        x = self.variable.data
        f = fnan if x.dtype.kind in "fc" else fplain
        if dim is None and axis is None:
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    return self._new(np.asarray(f(x)), dims=())
        dims = [dim] if isinstance(dim, str) else list(dim) if dim is not None else [self.dims[a] for a in np.atleast_1d(axis)]
        ax = tuple(self.dims.index(d) for d in dims)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = f(x, axis=ax)
        return self._new(np.asarray(r), dims=[d for d in self.dims if d not in dims])

    def max(self, dim=None, **kw):
        return self._reduce(np.nanmax, np.max, dim, **kw)

    def min(self, dim=None, **kw):
        return self._reduce(np.nanmin, np.min, dim, **kw)

    def sum(self, dim=None, **kw):
        return self._reduce(np.nansum, np.sum, dim, **kw)

    def mean(self, dim=None, **kw):
        return self._reduce(np.nanmean, np.mean, dim, **kw)

    def any(self, dim=None, **kw):
        return self._reduce(np.any, np.any, dim, **kw)

    def all(self, dim=None, **kw):
        return self._reduce(np.all, np.all, dim, **kw)

    def argmin(self, dim=None, **kw):
        ax = self.dims.index(dim)
        x = self.variable.data
        r = np.nanargmin(x, axis=ax) if x.dtype.kind == "f" else np.argmin(x, axis=ax)   # all-NaN -> ValueError
        return self._new(r, dims=[d for d in self.dims if d != dim])

    # ---- shape manipulation
    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        if set(dims) != set(self.dims):
            raise ValueError("xrlite: %r is not a permutation of %r" % (dims, self.dims))
        return self._new(self.variable.data.transpose([self.dims.index(d) for d in dims]), dims=dims)

    def squeeze(self):
        keep = [i for i, n in enumerate(self.shape) if n != 1]
        coords = dict(self._coords)
        for i, n in enumerate(self.shape):
            d = self.dims[i]
            if n == 1 and d in coords and coords[d].dims == (d,):
                coords[d] = Variable((), coords[d].data[0], coords[d].attrs)
        return self._new(self.variable.data.reshape([self.shape[i] for i in keep]), dims=[self.dims[i] for i in keep],
                         coords={k: v for k, v in coords.items()})

    def rename(self, mapping):
        if isinstance(mapping, str):
            return self._new(self.variable.data, name=mapping)
        ren = lambda ds: tuple(mapping.get(d, d) for d in ds)
        coords = {mapping.get(k, k): Variable(ren(v.dims), v.data, v.attrs) for k, v in self._coords.items()}
        return DataArray(variable=Variable(ren(self.dims), self.variable.data, self.variable.attrs), _coords=coords,
                         name=self.name)

    def expand_dims(self, dim=None, axis=None):
        if isinstance(dim, str):
            dim = {dim: None}
        elif not isinstance(dim, dict):
            dim = {d: None for d in dim}
        x, dims, coords = self.variable.data, list(self.dims), dict(self._coords)
        pos = 0 if axis is None else axis
        for i, (d, vals) in enumerate(dim.items()):
            if d in dims:
                raise ValueError("dimension %r already exists" % d)
            if vals is None:
                n = 1
                if d in coords:                      # scalar coordinate -> length-1 index coordinate
                    coords[d] = Variable((d,), coords[d].data.reshape(1), coords[d].attrs)
            else:
                v = _as_data(vals)
                n = v.shape[0]
                coords[d] = Variable((d,), v, vals.variable.attrs if isinstance(vals, DataArray) else None)
            x = np.broadcast_to(np.expand_dims(x, pos + i), x.shape[:pos + i] + (n,) + x.shape[pos + i:])
            dims.insert(pos + i, d)
        return self._new(x, dims=dims, coords=coords)

    def diff(self, dim, n=1, label="upper"):
        assert n == 1
        ax = self.dims.index(dim)
        sl = slice(None, -1) if label == "lower" else slice(1, None)
        coords = {k: (Variable(v.dims, v.data[tuple(sl if d == dim else slice(None) for d in v.dims)], v.attrs)
                      if dim in v.dims else v) for k, v in self._coords.items()}
        return self._new(np.diff(self.variable.data, axis=ax), coords=coords)

    def where(self, cond, other=np.nan):
        return where(cond, self, other, _like=self)

    # ---- selection
    def _positions(self, dim, labels):
        idx = self._coords[dim].data
        lab = _as_data(labels)
        flat = np.atleast_1d(lab).ravel()
        pos = np.empty(flat.shape, dtype=np.intp)
        for i, v in enumerate(flat):
            hit = np.nonzero(idx == v)[0]
            if hit.size == 0:
                raise KeyError("%r not found in coordinate %r" % (v, dim))
            pos[i] = hit[0]
        return pos.reshape(lab.shape)

    def sel(self, indexers=None, method=None, **kw):
        assert method is None
        indexers = dict(indexers or {}, **kw)
        out = self
        for dim, lab in indexers.items():
            if isinstance(lab, DataArray) and lab.ndim > 0:
                pos = out._positions(dim, lab)
                out = out.isel({dim: DataArray(variable=Variable(lab.dims, pos), _coords=lab._coords)})
            elif isinstance(lab, (list, np.ndarray)) and np.ndim(lab) > 0:
                out = out.isel({dim: out._positions(dim, lab)})
            else:
                out = out.isel({dim: int(out._positions(dim, _scalar(lab)))})
        return out

    def isel(self, indexers=None, drop=False, **kw):
        indexers = dict(indexers or {}, **kw)
        out = self
        for dim, ind in indexers.items():
            ax = out.dims.index(dim)
            x = out.variable.data
            if isinstance(ind, DataArray) and ind.ndim > 0:
                # vectorised (pointwise) indexing: dimensions of the indexer that the array also
                # has are paired element by element, the others are new
                new_dims = [d for d in out.dims if d != dim]
                res_dims = list(ind.dims) + [d for d in new_dims if d not in ind.dims]
                idx_b = ind.variable.data
                grids = {}
                shape = [idx_b.shape[ind.dims.index(d)] if d in ind.dims else x.shape[out.dims.index(d)]
                         for d in res_dims]
                for d in new_dims:
                    s = [1] * len(res_dims)
                    s[res_dims.index(d)] = shape[res_dims.index(d)]
                    grids[d] = np.arange(shape[res_dims.index(d)]).reshape(s)
                s = [shape[res_dims.index(d)] if d in ind.dims else 1 for d in res_dims]
                idx_full = idx_b.transpose([ind.dims.index(d) for d in res_dims if d in ind.dims]).reshape(s)
                key = tuple(idx_full if d == dim else grids[d] for d in out.dims)
                r = x[key]
                coords = {}
                for k, v in out._coords.items():
                    if dim not in v.dims:
                        coords[k] = v
                    elif v.dims == (dim,):
                        coords[k] = Variable(ind.dims, v.data[idx_b], v.attrs)    # N-D, no longer an index
                for k, v in ind._coords.items():
                    coords.setdefault(k, v)
                out = out._new(r, dims=res_dims, coords=coords)
                continue
            if isinstance(ind, (int, np.integer)):
                key = int(ind)
            elif isinstance(ind, slice):
                key = ind
            else:
                key = np.asarray(list(ind) if isinstance(ind, range) else _as_data(ind))
                if key.ndim == 0:
                    key = int(key)
            full = tuple(key if i == ax else slice(None) for i in range(x.ndim))
            dims = [d for d in out.dims if not (d == dim and isinstance(key, int))]
            coords = {}
            for k, v in out._coords.items():
                if dim in v.dims:
                    kk = tuple(key if d == dim else slice(None) for d in v.dims)
                    nd = [d for d in v.dims if not (d == dim and isinstance(key, int))]
                    coords[k] = Variable(nd, v.data[kk], v.attrs)
                else:
                    coords[k] = v
            out = DataArray(variable=Variable(dims, x[full], out.variable.attrs), name=out.name,
                            _coords={k: v for k, v in coords.items() if set(v.dims) <= set(dims)})
        return out

    @property
    def loc(self):
        return _Loc(self)

    def reindex(self, indexers):
        out = self
        for dim, labels in indexers.items():
            labels = np.array([_scalar(v) for v in labels])
            out = out.isel({dim: out._positions(dim, labels)})
        return out

    def drop_sel(self, **labels):
        out = self
        for dim, lab in labels.items():
            keep = np.nonzero(out._coords[dim].data != _as_data(lab))[0]
            out = out.isel({dim: keep})
        return out

    def assign_coords(self, coords=None, **kw):
        out = DataArray(variable=self.variable, _coords=dict(self._coords), name=self.name)
        for k, v in dict(coords or {}, **kw).items():
            out[k] = v
        return out

    # ---- interpolation: scipy interp1d(kind='linear', bounds_error=False), as xarray.core.missing
    def interp(self, coords=None, method="linear", **kw):
        assert method == "linear"
        if self.dtype.kind not in "uifc":
            raise TypeError("interp only works for a numeric type array")
        out = self
        for dim, new in dict(coords or {}, **kw).items():
            ax = out.dims.index(dim)
            x = out._coords[dim].data
            new_v = _as_data(new)
            if x.dtype.kind == "M":                      # _floatize_x: ns since the smallest stamp
                xmin = x.min()
                xf = (x - xmin).astype(np.float64)
                nf = (new_v.astype("datetime64[ns]") - xmin).astype(np.float64)
            else:
                xf, nf = x.astype(np.float64), new_v.astype(np.float64)
            f = interp1d(xf, out.variable.data, kind="linear", axis=ax, bounds_error=False, fill_value=np.nan,
                         assume_sorted=False)
            r = f(np.ravel(nf))
            coords_new = {k: v for k, v in out._coords.items() if dim not in v.dims}
            if new_v.ndim == 0:
                r = np.take(r, 0, axis=ax)
                dims = [d for d in out.dims if d != dim]
                coords_new[dim] = Variable((), new_v)
            else:
                dims = list(out.dims)
                if isinstance(new, DataArray) and new.dims != (dim,):
                    dims[ax] = new.dims[0]
                    coords_new.update({k: v for k, v in new._coords.items()})
                    coords_new[dim] = Variable(new.dims, new_v)
                else:
                    coords_new[dim] = Variable((dim,), new_v,
                                               new.variable.attrs if isinstance(new, DataArray) else None)
            out = DataArray(variable=Variable(dims, r, out.variable.attrs), name=out.name, _coords=coords_new)
        return out

    # ---- conversion / output
    def to_dataset(self, name=None):
        ds = Dataset()
        for k, v in self._coords.items():
            ds._coords[k] = v
        ds._vars[name or self.name] = self.variable
        return ds

    def to_netcdf(self, path, mode="w"):
        self.to_dataset(self.name or "__xarray_dataarray_variable__").to_netcdf(path, mode)


class _Loc:
    def __init__(self, da):
        self.da = da

    def __setitem__(self, key, value):
        da = self.da
        full = [slice(None)] * da.ndim
        rest = list(da.dims)
        for dim, lab in key.items():
            full[da.dims.index(dim)] = int(da._positions(dim, _scalar(lab)))
            rest.remove(dim)
        if isinstance(value, DataArray):
            value = value.transpose(*[d for d in rest if d in value.dims]).variable.data if value.ndim else value.values
        da.variable.data[tuple(full)] = value

    def __getitem__(self, key):
        return self.da.sel(key)


def where(cond, x, y, _like=None):
    """xr.where: np.where broadcast by dimension name."""
    dims, (c, a, b) = _broadcast([cond, x, y])
    das = [o for o in (cond, x, y) if isinstance(o, DataArray)]
    if not isinstance(cond, DataArray):
        dims = das[0].dims
    return DataArray(variable=Variable(dims, np.where(c, a, b), _like.variable.attrs if _like is not None else None),
                     _coords=_merge_coords(das, dims), name=_like.name if _like is not None else None)


def zeros_like(da):
    return da._new(np.zeros_like(da.variable.data))


def full_like(da, fill_value, dtype=None):
    return da._new(np.full_like(da.variable.data, fill_value, dtype=dtype))


def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=((),), vectorize=False, **kw):
    """Only what the reference uses: core dims moved last, the rest broadcast by name, np.vectorize."""
    assert vectorize
    input_core_dims = input_core_dims or [[] for _ in args]
    das = [a for a in args if isinstance(a, DataArray)]
    loop_dims, sizes = [], {}
    for a, core in zip(args, input_core_dims):
        if isinstance(a, DataArray):
            for d, n in zip(a.dims, a.shape):
                sizes[d] = n
                if d not in core and d not in loop_dims:
                    loop_dims.append(d)
    arrs = []
    for a, core in zip(args, input_core_dims):
        if isinstance(a, DataArray):
            order = [d for d in loop_dims if d in a.dims] + list(core)
            x = a.transpose(*order).variable.data if order else a.variable.data
            shape = [sizes[d] if d in a.dims else 1 for d in loop_dims] + [sizes[d] for d in core]
            arrs.append(x.reshape(shape))
        else:
            arrs.append(a)
    if any(len(c) for c in list(input_core_dims) + list(output_core_dims)):
        sig = ",".join("(" + ",".join(c) + ")" for c in input_core_dims) + "->" + \
              ",".join("(" + ",".join(c) + ")" for c in output_core_dims)
        vf = np.vectorize(func, signature=sig)
    else:
        vf = np.vectorize(func)
    res = vf(*arrs)
    single = not isinstance(res, tuple)
    res = (res,) if single else res
    outs = []
    for r, core in zip(res, output_core_dims):
        dims = tuple(loop_dims) + tuple(core)
        outs.append(DataArray(variable=Variable(dims, np.asarray(r)), _coords=_merge_coords(das, dims)))
    return outs[0] if single else tuple(outs)


# ---------------------------------------------------------------------------------------------
class _Indexes:
    def __init__(self, ds):
        self.ds = ds

    def __getitem__(self, k):
        return pd.Index(self.ds._coords[k].data)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars, self._coords, self.attrs = {}, {}, dict(attrs or {})

    # coordinates = variables named like their only dimension, or promoted by isel/assign_coords
    def __contains__(self, k):
        return k in self._vars or k in self._coords

    def __getitem__(self, k):
        if k in self._coords:
            v = self._coords[k]
        else:
            v = self._vars[k]
        return DataArray(variable=v, name=k,
                         _coords={c: cv for c, cv in self._coords.items() if set(cv.dims) <= set(v.dims)})

    def __getattr__(self, k):
        if k.startswith("_"):
            raise AttributeError(k)
        if k in self:
            return self[k]
        raise AttributeError(k)

    def __setitem__(self, k, value):
        if isinstance(value, DataArray):
            for c, cv in value._coords.items():
                if c not in self._coords:
                    self._coords[c] = cv
                elif cv.dims == (c,) and not np.array_equal(self._coords[c].data, cv.data):
                    raise ValueError("xrlite: coordinate %r of the new variable differs" % c)
            var = Variable(value.dims, value.variable.data, value.variable.attrs)
        else:
            dims, data = value[0], value[1]
            var = Variable(dims, data)
        if k in self._coords:
            self._coords[k] = var
        else:
            self._vars[k] = var

    def __delitem__(self, k):
        if k in self._vars:
            del self._vars[k]
        else:
            del self._coords[k]

    @property
    def indexes(self):
        return _Indexes(self)

    @property
    def dims(self):
        out = {}
        for v in list(self._coords.values()) + list(self._vars.values()):
            for d, n in zip(v.dims, v.data.shape):
                out[d] = n
        return out

    @property
    def data_vars(self):
        return {k: self[k] for k in self._vars}

    def _map(self, dim, f):
        """Apply ``f(DataArray) -> DataArray`` to every variable and coordinate that has ``dim``."""
        out = Dataset(attrs=self.attrs)
        index = self._coords.get(dim)
        for src, dst in ((self._coords, out._coords), (self._vars, out._vars)):
            for k, v in src.items():
                if dim in v.dims:
                    dst[k] = f(DataArray(variable=v, _coords={dim: index} if index is not None else {}, name=k)).variable
                else:
                    dst[k] = v
        return out

    def isel(self, indexers=None, **kw):
        out = self
        for dim, ind in dict(indexers or {}, **kw).items():
            out = out._map(dim, lambda da: da.isel({dim: ind}))
        return out

    def reindex(self, indexers):
        out = self
        for dim, labels in indexers.items():
            labels = [_scalar(v) for v in labels]
            out = out._map(dim, lambda da: da.reindex({dim: labels}))
        return out

    def drop_sel(self, **labels):
        out = self
        for dim, lab in labels.items():
            out = out._map(dim, lambda da: da.drop_sel(**{dim: lab}))
        return out

    def interp(self, coords=None, method="linear", **kw):
        out = self
        for dim, new in dict(coords or {}, **kw).items():
            res = Dataset(attrs=out.attrs)
            res._coords = {k: v for k, v in out._coords.items() if dim not in v.dims}
            for k, v in out._vars.items():
                if dim in v.dims:
                    if v.data.dtype.kind in "uifc":
                        r = out[k].interp({dim: new}, method=method)
                        res._vars[k] = r.variable
                        res._coords[dim] = r._coords[dim]
                else:
                    res._vars[k] = v
            out = res
        return out

    def assign_coords(self, coords=None, **kw):
        out = Dataset(attrs=self.attrs)
        out._vars, out._coords = dict(self._vars), dict(self._coords)
        for k, v in dict(coords or {}, **kw).items():
            data = _as_data(v)
            dims = v.dims if isinstance(v, DataArray) else ((k,) if data.ndim else ())
            out._coords[k] = Variable(dims, data, v.variable.attrs if isinstance(v, DataArray) else
                                      (self._coords[k].attrs if k in self._coords else None))
        return out

    def close(self):
        pass

    def to_netcdf(self, path, mode="w"):
        with netcdf_file(path, "w", version=2) as f:
            for k, v in self.attrs.items():
                setattr(f, k, v)
            dims = self.dims
            unlimited = None
            for d, n in dims.items():
                f.createDimension(d, None if d == unlimited else n)
            for k, v in list(self._coords.items()) + list(self._vars.items()):
                data, attrs = v.data, dict(v.attrs)
                if data.dtype.kind == "M":
                    data = (data - np.datetime64("1970-01-01", "ns")).astype(np.int64) / 1e9
                    attrs.update(units="seconds since 1970-01-01 00:00:00", calendar="proleptic_gregorian")
                if data.dtype == np.int64:
                    data = data.astype(np.float64 if data.dtype.kind == "f" else np.int32)
                if data.dtype == np.bool_:
                    data = data.astype(np.int8)
                vdims = v.dims
                if unlimited in vdims and vdims[0] != unlimited:
                    raise ValueError("xrlite: record dimension must lead")
                var = f.createVariable(k, data.dtype.newbyteorder("=").char if data.dtype.kind != "f" else
                                       ("f" if data.dtype.itemsize == 4 else "d"), vdims)
                if data.ndim:
                    var[:] = data
                else:
                    var.assignValue(data[()])
                for a, av in attrs.items():
                    setattr(var, a, av)


def _decode_time(data, attrs):
    units = attrs.get("units", b"")
    units = units.decode() if isinstance(units, bytes) else units
    cal = attrs.get("calendar", b"standard")
    cal = cal.decode() if isinstance(cal, bytes) else cal
    if " since " not in units:
        return None
    if cal not in ("standard", "gregorian", "proleptic_gregorian"):
        raise NotImplementedError("xrlite: calendar %r (cftime is not available)" % cal)
    unit, ref = units.split(" since ")
    ns = {"days": 86400e9, "day": 86400e9, "hours": 3600e9, "hour": 3600e9, "minutes": 60e9, "seconds": 1e9,
          "second": 1e9}[unit.strip().lower()]
    ref = pd.Timestamp(ref.strip())
    # xarray.coding.times._decode_datetime_with_pandas: integer nanoseconds, then pandas
    flat = (np.asarray(data, dtype=np.float64) * ns).astype(np.int64)
    return (pd.to_timedelta(flat.ravel(), "ns") + ref).values.reshape(np.shape(data)).astype("datetime64[ns]")


def open_dataset(path, decode_cf=True, **kw):
    ds = Dataset()
    with netcdf_file(path, "r", mmap=False) as f:
        ds.attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in f._attributes.items()}
        for name, var in f.variables.items():
            data = np.array(var.data, copy=True)
            data = data.astype(data.dtype.newbyteorder("="))
            attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in var._attributes.items()}
            if decode_cf:
                t = _decode_time(data, attrs) if data.dtype.kind in "fiu" else None
                if t is not None:
                    data = t
                    attrs.pop("units", None), attrs.pop("calendar", None)
                elif data.dtype.kind == "f":
                    for key in ("_FillValue", "missing_value"):
                        if key in attrs:
                            data = np.where(data == np.asarray(attrs.pop(key)).astype(data.dtype), np.nan, data
                                            ).astype(data.dtype)
                if "scale_factor" in attrs or "add_offset" in attrs:
                    data = data * attrs.pop("scale_factor", 1.0) + attrs.pop("add_offset", 0.0)
            v = Variable(var.dimensions, data, attrs)
            if var.dimensions == (name,):
                ds._coords[name] = v
            else:
                ds._vars[name] = v
    return ds


def concat(objs, dim):
    """xr.concat of Datasets along ``dim``: an existing dimension of some of them (the others hold it as
    a scalar coordinate and count as length 1), or a scalar coordinate of all (new leading dimension)."""
    assert all(isinstance(o, Dataset) for o in objs)
    if not any(dim in o.dims for o in objs):
        ps = []
        for o in objs:                                  # Dataset.expand_dims(dim)
            p = Dataset(attrs=o.attrs)
            c = o._coords[dim]
            p._coords = dict(o._coords)
            p._coords[dim] = Variable((dim,), c.data.reshape(1), c.attrs)
            p._vars = {k: Variable((dim,) + v.dims, v.data[None], v.attrs) for k, v in o._vars.items()}
            ps.append(p)
        objs = ps
    out = Dataset(attrs=objs[0].attrs)
    out._coords = {k: v for k, v in objs[0]._coords.items() if k != dim and dim not in v.dims}
    for o in objs[1:]:                                  # align(join='outer', exclude=[dim]): equal here
        for k, v in o._coords.items():
            if k != dim and v.dims == (k,) and k in out._coords and not np.array_equal(out._coords[k].data, v.data):
                raise ValueError("xrlite: coordinate %r differs between the concatenated datasets" % k)
    out._coords[dim] = Variable((dim,), np.concatenate([np.atleast_1d(o._coords[dim].data) for o in objs]),
                                objs[0]._coords[dim].attrs)
    lens = [o.dims.get(dim, 1) for o in objs]
    for k in objs[0]._vars:
        vs = [o._vars[k] for o in objs]
        common = []
        for v in vs:                                    # ensure_common_dims
            for d in v.dims:
                if d not in common:
                    common.append(d)
        if dim not in common:
            common.insert(0, dim)
        sizes = {}
        for v in vs:
            sizes.update({d: n for d, n in zip(v.dims, v.data.shape) if d != dim})
        parts = []
        for v, n in zip(vs, lens):
            order = [v.dims.index(d) for d in common if d in v.dims]
            x = v.data.transpose(order).reshape([v.data.shape[v.dims.index(d)] if d in v.dims else 1 for d in common])
            parts.append(np.broadcast_to(x, [n if d == dim else sizes[d] for d in common]))
        out._vars[k] = Variable(common, np.concatenate(parts, axis=common.index(dim)), vs[0].attrs)
    return out
