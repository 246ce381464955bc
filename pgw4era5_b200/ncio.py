"""
Minimal NetCDF in/out for the drop-in CLIs.

The reference reads and writes its files with xarray (step_03_apply_to_era.py:60,378;
functions.py:203,626,675).  xarray/netCDF4 are optional here: when they are
importable they are used (any NetCDF flavour), otherwise files are handled with
``scipy.io.netcdf_file`` (NetCDF-3 classic / 64-bit offset), which is all the build
image offers.  Either way the rest of the package sees the same tiny ``Dataset``:
variables with named dimensions, numpy data and attributes, nothing decoded
except the time axis of the GCM delta files.
"""
from collections import OrderedDict
from datetime import datetime, timedelta

import numpy as np


class Variable:
    __slots__ = ("dims", "data", "attrs")

    def __init__(self, dims, data, attrs=None):
        self.dims = tuple(dims)
        self.data = np.asarray(data)
        self.attrs = dict(attrs or {})
        if self.data.ndim != len(self.dims):
            raise ValueError("variable has %d dims but data has %d" % (len(self.dims), self.data.ndim))

    @property
    def values(self):
        return self.data

    @property
    def shape(self):
        return self.data.shape

    def copy(self):
        return Variable(self.dims, self.data.copy(), self.attrs)


class Dataset:
    """Ordered name -> Variable mapping plus global attributes."""

    def __init__(self, variables=None, attrs=None):
        self.variables = OrderedDict(variables or {})
        self.attrs = dict(attrs or {})

    def __contains__(self, name):
        return name in self.variables

    def __getitem__(self, name):
        return self.variables[name]

    def __setitem__(self, name, var):
        if not isinstance(var, Variable):
            raise TypeError("Dataset values must be Variable instances")
        self.variables[name] = var

    def __delitem__(self, name):
        del self.variables[name]

    def __getattr__(self, name):
        try:
            return self.__dict__["variables"][name]
        except KeyError:
            raise AttributeError(name)

    def keys(self):
        return self.variables.keys()

    def dim_sizes(self):
        sizes = OrderedDict()
        for v in self.variables.values():
            for d, n in zip(v.dims, v.data.shape):
                if sizes.setdefault(d, n) != n:
                    raise ValueError("dimension %r has inconsistent sizes" % d)
        return sizes

    def copy(self):
        return Dataset(OrderedDict((k, v.copy()) for k, v in self.variables.items()), self.attrs)

    def close(self):
        pass

    def to_netcdf(self, path, mode="w"):
        write_dataset(self, path)


# --------------------------------------------------------------------------- reading
def _scalar_attr(v):
    if isinstance(v, bytes):
        return v.decode("utf-8", "replace")
    if isinstance(v, np.ndarray) and v.size == 1:
        return v.reshape(()).item()
    return v


def open_dataset(path, decode_cf=False):
    """Read every variable of a NetCDF file into memory (raw dtypes, like the reference's
    ``xr.open_dataset(path, decode_cf=False)``).  ``decode_cf`` is accepted for signature
    compatibility; time decoding is explicit through ``decode_time``."""
    try:
        import netCDF4  # type: ignore
    except ImportError:
        netCDF4 = None
    ds = Dataset()
    if netCDF4 is not None:
        with netCDF4.Dataset(path, "r") as nc:
            nc.set_auto_maskandscale(False)
            for name, var in nc.variables.items():
                attrs = {k: _scalar_attr(var.getncattr(k)) for k in var.ncattrs()}
                ds[name] = Variable(var.dimensions, np.array(var[...]), attrs)
            ds.attrs = {k: _scalar_attr(nc.getncattr(k)) for k in nc.ncattrs()}
        return ds
    # NetCDF-3 (classic, 64-bit offset, CDF-5): own header parser + positional reads in pieces -- scipy's
    # reader fails on records beyond 2 GiB (one global 0.25 degree ERA5 timestep is 2.3 GB)
    from .nc3raw import NotNetCDF3, RawNC3
    try:
        raw = RawNC3(path)
    except NotNetCDF3:
        raw = None
    if raw is not None:
        with open(path, "rb", buffering=0) as f:
            for name, v in raw.vars.items():
                ds[name] = Variable(v.dims, raw.read_variable(f, name), v.attrs)
        ds.attrs = dict(raw.attrs)
        return ds
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False, maskandscale=False) as nc:
        for name, var in nc.variables.items():
            attrs = {k: _scalar_attr(v) for k, v in var._attributes.items()}
            data = np.array(var.data)
            if data.dtype.byteorder == ">":
                data = data.astype(data.dtype.newbyteorder("="))
            ds[name] = Variable(var.dimensions, data, attrs)
        ds.attrs = {k: _scalar_attr(v) for k, v in nc._attributes.items()}
    return ds


def write_dataset(ds, path):
    """Write a Dataset as NetCDF-3 (64-bit offset), the first dimension named 'time' unlimited."""
    from scipy.io import netcdf_file
    sizes = ds.dim_sizes()
    with netcdf_file(path, "w", version=2) as nc:
        if "time" in sizes:                              # the record dimension has to come first
            nc.createDimension("time", None)
        for d, n in sizes.items():
            if d != "time":
                nc.createDimension(d, n)
        for name, var in ds.variables.items():
            data = var.data
            if data.dtype == np.float16:
                data = data.astype(np.float32)
            if data.dtype == np.int64:
                data = data.astype(np.int32) if np.all(np.abs(data) < 2 ** 31) else data.astype(np.float64)
            if data.dtype == np.bool_:
                data = data.astype(np.int8)
            v = nc.createVariable(name, data.dtype.newbyteorder("="), var.dims)
            if "time" in var.dims and var.dims[0] != "time":
                raise ValueError("the record dimension 'time' must come first in %r" % name)
            if data.ndim == 0:
                v.assignValue(data)
            else:
                v[:] = data
            for k, a in var.attrs.items():
                if isinstance(a, (str, bytes, int, float, np.generic, np.ndarray, list, tuple)):
                    setattr(v, k, a)
        for k, a in ds.attrs.items():
            if isinstance(a, (str, bytes, int, float, np.generic, np.ndarray, list, tuple)):
                setattr(nc, k, a)


# --------------------------------------------------------------------------- time axes
_UNIT_SECONDS = {"second": 1.0, "seconds": 1.0, "sec": 1.0, "secs": 1.0, "s": 1.0,
                 "minute": 60.0, "minutes": 60.0, "min": 60.0, "hour": 3600.0, "hours": 3600.0,
                 "hr": 3600.0, "hrs": 3600.0, "h": 3600.0, "day": 86400.0, "days": 86400.0, "d": 86400.0}


def _month_lengths(calendar, year):
    """Days per month of ``year`` in a CF calendar that is not the proleptic Gregorian one."""
    if calendar in ("noleap", "365_day"):
        feb = 28
    elif calendar in ("all_leap", "366_day"):
        feb = 29
    elif calendar == "360_day":
        return [30] * 12
    elif calendar == "julian":
        feb = 29 if year % 4 == 0 else 28
    else:
        raise ValueError("unsupported calendar %r" % calendar)
    return [31, feb, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def decode_time(var):
    """CF time axis ('<unit> since <date>') -> numpy datetime64[ns], the way the reference ends up with it after
    ``xr.open_dataset`` + ``CFTimeIndex.to_datetimeindex()`` (functions.py:203-221): in a non-standard calendar
    (noleap/365_day, all_leap/366_day, 360_day, julian) the stamp is counted in THAT calendar and its calendar
    date (year, month, day, time of day) is kept; a date the standard calendar does not have (30 February of a
    360_day file, 29 February of a non-leap year) raises ValueError, as ``to_datetimeindex`` does -- nothing is
    ever shifted.  standard / gregorian / proleptic_gregorian use the proleptic Gregorian arithmetic of
    ``datetime`` (identical for dates after 1582).  Any other calendar name raises ValueError."""
    units = str(var.attrs.get("units", ""))
    calendar = str(var.attrs.get("calendar", "standard")).lower()
    if np.issubdtype(var.data.dtype, np.datetime64):
        return var.data.astype("datetime64[ns]")
    if " since " not in units:
        raise ValueError("time variable has no CF units: %r" % units)
    unit, ref = units.split(" since ", 1)
    scale = _UNIT_SECONDS[unit.strip().lower()]
    ref = ref.strip().replace("T", " ").replace("Z", "")
    parts = ref.split()
    date = parts[0].split("-")
    hms = (parts[1].split(":") if len(parts) > 1 else []) + ["0", "0", "0"]
    oy, om, od = int(date[0]), int(date[1]), int(date[2])
    oh, omi, osec = int(hms[0]), int(hms[1]), int(float(hms[2]))
    secs = np.asarray(var.data, dtype=np.float64) * scale
    out = []
    if calendar in ("standard", "gregorian", "proleptic_gregorian"):
        origin = datetime(oy, om, od, oh, omi, osec)
        for s in secs:
            out.append(np.datetime64(origin + timedelta(seconds=float(s)), "ns"))
        return np.array(out, dtype="datetime64[ns]")
    ml0 = _month_lengths(calendar, oy)               # raises for an unknown calendar
    if not (1 <= om <= 12 and 1 <= od <= ml0[om - 1]):
        raise ValueError("reference date %s does not exist in calendar %s" % (parts[0], calendar))
    base = oh * 3600 + omi * 60 + osec
    for s in np.atleast_1d(secs):
        total = base + float(s)
        days = int(np.floor(total / 86400.0))
        rem = total - days * 86400.0
        year = oy
        doy = sum(ml0[:om - 1]) + od - 1 + days      # day of `year`, may lie outside it
        while doy < 0:
            year -= 1
            doy += sum(_month_lengths(calendar, year))
        while doy >= sum(_month_lengths(calendar, year)):
            doy -= sum(_month_lengths(calendar, year))
            year += 1
        ml = _month_lengths(calendar, year)
        m = 0
        while doy >= ml[m]:
            doy -= ml[m]
            m += 1
        try:
            stamp = datetime(year, m + 1, doy + 1)
        except ValueError:
            raise ValueError("Cannot convert date %04d-%02d-%02d of calendar %s to a date that is valid in the "
                             "standard calendar" % (year, m + 1, doy + 1, calendar))
        out.append(np.datetime64(stamp + timedelta(seconds=float(rem)), "ns"))
    return np.array(out, dtype="datetime64[ns]").reshape(np.shape(secs))


def encode_time(stamps, units="seconds since 1970-01-01 00:00:00"):
    """datetime64 -> float64 CF axis (used when writing delta files)."""
    unit, ref = units.split(" since ", 1)
    origin = np.datetime64(ref.strip().replace(" ", "T"), "ns")
    secs = (np.asarray(stamps).astype("datetime64[ns]") - origin) / np.timedelta64(1, "s")
    return Variable(("time",), secs / _UNIT_SECONDS[unit.strip().lower()],
                    {"units": units, "calendar": "standard"})
