"""
Drop-in for the reference's ``functions.py``: same function names, argument order and
error behaviour, numerics on the GPU through libpgw_b200.so (no CPU fallback).

Arrays may be numpy arrays, torch tensors, ``ncio.Variable`` objects or (when xarray is
installed) DataArrays; results come back as numpy arrays (torch CUDA tensors if the first
array argument was one).  float64 inputs are computed in float64 like the reference,
float32 inputs in float32.  Functions that the reference feeds with xarray objects
carrying coordinates (``vert_interp_delta``, ``regrid_lat_lon``, ``load_delta`` ...) take
the coordinates from ``ncio`` objects or explicit keyword arguments instead.

Each docstring cites the reference lines it replaces (menschj/PGW4ERA5).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N
from . import ncio, timeinterp
from .constants import CON_G, CON_MW_MD, CON_RD  # noqa: F401  (re-exported like the reference)
from .settings import (  # noqa: F401
    i_debug, i_use_xesmf_regridding, file_name_bases,
    TIME_ERA, LEV_ERA, HLEV_ERA, LON_ERA, LAT_ERA,
    TIME_GCM, PLEV_GCM, LON_GCM, LAT_GCM,
    LAT_GCM_OCEAN, LON_GCM_OCEAN, TIME_GCM_OCEAN,
)

_MSG_TOP = ('ERA5 top pressure is lower than climate delta top pressure. If you are certain that '
            'you do not need the data beyond to upper-most pressure level of the climate delta, '
            'you can set the flag --ignore_top_pressure_error and re-run the script.')


# --------------------------------------------------------------------------- plumbing
def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pgw4era5_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _raw(x):
    if isinstance(x, ncio.Variable):
        return x.data
    if hasattr(x, "values") and not isinstance(x, (np.ndarray, torch.Tensor)):
        return np.asarray(x.values)          # xarray DataArray
    return x


def _dev(x, dtype=None):
    x = _raw(x)
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("="))
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is None:
        dtype = torch.float64 if t.dtype == torch.float64 else torch.float32
    return t.to(_device(), dtype).contiguous()


def _work_dtype(*xs):
    for x in xs:
        x = _raw(x)
        d = x.dtype if isinstance(x, (np.ndarray, torch.Tensor)) else np.asarray(x).dtype
        if d in (np.float64, torch.float64):
            return torch.float64
    return torch.float32


def _back(t, like):
    like = _raw(like)
    if isinstance(like, torch.Tensor) and like.is_cuda:
        return t
    return t.cpu().numpy()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _sfx(dtype):
    return "f64" if dtype == torch.float64 else "f32"


def _new_err():
    return torch.zeros(1, device=_device(), dtype=torch.int32)


def _raise_interp_errors(err):
    bits = int(err.item()) & 0xFFFFFFFF
    if bits & N.ERR_SRC_NOT_ASCENDING:
        raise ValueError('Source pressure values must be ascending!')          # functions.py:501
    if bits & N.ERR_TARG_NOT_ASCENDING:
        raise ValueError('Target pressure values must be ascending!')          # functions.py:503
    if bits & N.ERR_EXTRAP_OFF:
        raise ValueError('Extrapolation deactivated but data out of bounds.')  # functions.py:565
    return bits


# --------------------------------------------------------------------------- arbitrary
def dt64_to_dt(dt64):
    """functions.py:39-51"""
    return timeinterp.to_datetime(dt64)


# --------------------------------------------------------------------------- physics
def _humidity_op(op, x, y):
    dt = _work_dtype(x, y) if y is not None else _work_dtype(x)
    xd = _dev(x, dt)
    yd = _dev(y, dt).expand_as(xd).contiguous() if y is not None else None
    out = torch.empty_like(xd)
    fn = getattr(N.lib, "pgw_humidity_op_" + _sfx(dt))
    N.check(fn(op, _p(xd), _p(yd), _p(out), xd.numel(), _stream()), "pgw_humidity_op")
    return _back(out, x)


def specific_humidity_to_vapor_pressure(hus, pa):
    """functions.py:58-64"""
    return _humidity_op(0, hus, pa)


def vapor_pressure_to_specific_humidity(vapp, pa):
    """functions.py:66-72"""
    return _humidity_op(1, vapp, pa)


def saturation_vapor_pressure_water_or_ice(pa, ta, water=True):
    """functions.py:74-89 (IFS documentation 7.93)"""
    return _humidity_op(2 if water else 3, ta, None)


def saturation_vapor_pressure_water_and_ice(pa, ta):
    """functions.py:91-105 (IFS documentation 7.92)"""
    return _humidity_op(4, ta, None)


def _hum3(name, a, pa, ta):
    dt = _work_dtype(a, pa, ta)
    ad = _dev(a, dt)
    pd_ = _dev(pa, dt).expand_as(ad).contiguous()
    td = _dev(ta, dt).expand_as(ad).contiguous()
    out = torch.empty_like(ad)
    fn = getattr(N.lib, "pgw_%s_%s" % (name, _sfx(dt)))
    N.check(fn(_p(ad), _p(pd_), _p(td), _p(out), ad.numel(), _stream()), name)
    return _back(out, a)


def specific_to_relative_humidity(hus, pa, ta):
    """functions.py:107-116"""
    return _hum3("specific_to_relative_humidity", hus, pa, ta)


def relative_to_specific_humidity(hur, pa, ta):
    """functions.py:118-125"""
    return _hum3("relative_to_specific_humidity", hur, pa, ta)


def integ_geopot(pa_hl, zgs, ta, hus, level1=None, p_ref=30000):
    """
    functions.py:128-189.  pa_hl [nt, L+1, ny, nx], zgs [nt, ny, nx], ta/hus [nt, L, ny, nx],
    p_ref scalar or [nt, ny, nx].  ``level1`` (the half-level labels) is accepted for
    signature compatibility; labels 1..L+1 are assumed.  Returns float64 [nt, ny, nx].
    float32 ``ta`` and ``hus`` with float64 pressures are not promoted: like numpy in the reference,
    Rd * Tv is then formed in float32 (functions.py:144, :151).
    """
    dt = _work_dtype(pa_hl, ta, hus)
    dts = torch.float32 if _work_dtype(ta, hus) == torch.float32 else dt
    ph, td, qd = _dev(pa_hl, dt), _dev(ta, dts), _dev(hus, dts)
    if ph.dim() != 4:
        raise ValueError("pa_hl must be (time, level1, lat, lon)")
    nt, nl1, ny, nx = ph.shape
    zd = _dev(zgs, dt).reshape(nt, ny, nx)
    pref_field = None
    pref_scalar = 0.0
    if np.ndim(_raw(p_ref)) == 0:
        pref_scalar = float(p_ref)
    else:
        pref_field = _dev(p_ref, dt).reshape(nt, ny, nx)
    out = torch.empty((nt, ny, nx), device=ph.device, dtype=torch.float64)
    err = _new_err()
    fn = getattr(N.lib, "pgw_integ_geopot_" + _sfx(dt) + ("_f32" if dts != dt else ""))
    for t in range(nt):
        N.check(fn(_p(ph[t]), _p(zd[t]), _p(td[t]), _p(qd[t]),
                   _p(pref_field[t]) if pref_field is not None else _p(None), pref_scalar,
                   _p(out[t]), nl1 - 1, ny * nx, _p(err), _stream()), "pgw_integ_geopot")
    if int(err.item()) & N.ERR_PREF_BELOW_SFC:
        raise ValueError("p_ref locally lies below the surface. Please set a lower reference "
                         "pressue (p_ref_inp) in settings.py")
    return _back(out, pa_hl)


# --------------------------------------------------------------------------- deltas
class DeltaField:
    """What ``load_delta`` returns in place of an xarray DataArray: values plus coordinates."""

    def __init__(self, values, dims, coords, name):
        self.values, self.dims, self.coords, self.name = values, tuple(dims), dict(coords), name

    @property
    def shape(self):
        return self.values.shape

    def __getitem__(self, key):
        return self.coords[key]


def _open_delta(delta_input_dir, var_name, name_base):
    ds = ncio.open_dataset(os.path.join(delta_input_dir, name_base.format(var_name)))
    if var_name not in ds:
        raise KeyError(var_name)
    var = ds[var_name]
    stamps = ncio.decode_time(ds[TIME_GCM])
    coords = {TIME_GCM: stamps}
    for d in var.dims[1:]:
        if d in ds:
            coords[d] = np.asarray(ds[d].data, dtype=np.float64)
    return var, coords


def load_delta(delta_input_dir, var_name, era5_date_time, target_date_time=None,
               name_base=file_name_bases['SCEN-HIST']):
    """
    functions.py:195-303: open ``<var>_delta.nc``, drop 29 Feb, and (if ``target_date_time`` is
    given) interpolate linearly in time to that date with the periodic year wrap.  The blend
    runs on the GPU (pgw_time_interp_f32).  Returns a ``DeltaField`` with a leading time axis.
    """
    var, coords = _open_delta(delta_input_dir, var_name, name_base)
    keep = timeinterp.drop_leap_day(coords[TIME_GCM])
    data = var.data[keep]
    coords[TIME_GCM] = coords[TIME_GCM][keep]
    if target_date_time is None:                                   # functions.py:300-301
        return DeltaField(data, var.dims, coords, var_name)
    b = timeinterp.bracket(coords[TIME_GCM], target_date_time)
    lo = _dev(data[b.ind_before], torch.float32)
    hi = _dev(data[b.ind_after], torch.float32)
    out = torch.empty_like(lo)
    N.check(N.lib.pgw_time_interp_f32(_p(lo), _p(hi), b.x_hi, b.x_new, _p(out), lo.numel(), _stream()),
            "pgw_time_interp_f32")
    coords = dict(coords)
    coords[TIME_GCM] = np.asarray(_raw(era5_date_time)).reshape(-1)[:1]      # functions.py:296
    return DeltaField(out.cpu().numpy()[None], var.dims, coords, var_name)


def replace_delta_sfc(source_P, ps_hist, delta, delta_sfc):
    """functions.py:343-366 for one column (1-D arrays) or, vectorised on the GPU, for arrays
    shaped [K, ...] with ps_hist/delta_sfc shaped [...] (the apply_ufunc of :396-402)."""
    dt = _work_dtype(source_P, delta)
    d = _dev(delta, dt)
    K = d.shape[0]
    ncol = d.numel() // K
    sp = _dev(source_P, dt)
    src_1d = int(sp.dim() == 1 and d.dim() > 1)
    ph = _dev(np.broadcast_to(np.asarray(_raw(ps_hist) if not isinstance(_raw(ps_hist), torch.Tensor)
                                         else _raw(ps_hist).cpu()), d.shape[1:]) if d.dim() > 1
              else np.asarray([float(ps_hist)]), dt).reshape(-1)
    dsf = _dev(np.broadcast_to(np.asarray(_raw(delta_sfc) if not isinstance(_raw(delta_sfc), torch.Tensor)
                                          else _raw(delta_sfc).cpu()), d.shape[1:]) if d.dim() > 1
               else np.asarray([float(delta_sfc)]), dt).reshape(-1)
    out_P = torch.empty((K, ncol), device=d.device, dtype=dt)
    out_d = torch.empty((K, ncol), device=d.device, dtype=dt)
    err = _new_err()
    fn = getattr(N.lib, "pgw_replace_delta_sfc_" + _sfx(dt))
    N.check(fn(_p(sp), _p(ph), _p(d), _p(dsf), _p(out_P), _p(out_d), K, ncol, src_1d, _p(err), _stream()),
            "pgw_replace_delta_sfc")
    if int(err.item()) & N.ERR_PS_HIST_RANGE:
        raise ValueError()                                         # functions.py:361
    return _back(out_P.reshape(d.shape), source_P), _back(out_d.reshape(d.shape), delta)


def interp_extrap_1d(src_x, src_y, targ_x, extrapolate):
    """functions.py:511-580 (x arrays hold ln p, as in the numba original)."""
    sx = _dev(src_x, torch.float64).reshape(1, -1, 1)
    sy = _dev(src_y, torch.float64).reshape(1, -1, 1)
    tx = _dev(targ_x, torch.float64).reshape(1, -1, 1)
    out = torch.zeros_like(tx)
    err = _new_err()
    N.check(N.lib.pgw_interp_logp_f64(_p(sy), _p(sx), _p(tx), _p(out), 1, sx.shape[1], tx.shape[1], 1, 0, 1,
                                      N.EXTRAP_MODES[extrapolate], _p(err), _stream()), "pgw_interp_logp_f64")
    _raise_interp_errors(err)
    return out.reshape(-1).cpu().numpy()


def interp_1d_for_timelatlon(orig_array, src_p, targ_p, interp_array, ntime, nlat, nlon, extrapolate):
    """functions.py:479-508: same in-place signature (arrays hold ln p; result written into
    ``interp_array``)."""
    dt = _work_dtype(orig_array, src_p, targ_p)
    v, sp, tp = _dev(orig_array, dt), _dev(src_p, dt), _dev(targ_p, dt)
    out = torch.zeros_like(tp)
    err = _new_err()
    fn = getattr(N.lib, "pgw_interp_logp_" + _sfx(dt))
    N.check(fn(_p(v), _p(sp), _p(tp), _p(out), ntime, sp.shape[1], tp.shape[1], nlat * nlon, 0, 1,
               N.EXTRAP_MODES[extrapolate], _p(err), _stream()), "pgw_interp_logp")
    _raise_interp_errors(err)
    if isinstance(interp_array, torch.Tensor):
        interp_array.copy_(out)
    else:
        interp_array[...] = out.cpu().numpy()


def interp_logp_4d(var, source_P, targ_P, extrapolate='off', time_key=None, lat_key=None, lon_key=None):
    """functions.py:434-477: pressures in, logarithm taken on the device.  ``source_P`` may also
    be a 1-D pressure-level table shared by all columns.  The *_key arguments are accepted for
    compatibility (arrays carry no dimension names here)."""
    if extrapolate not in ['off', 'linear', 'constant', 'nan']:
        raise ValueError('Invalid input value for "extrapolate"')
    dt = _work_dtype(var, source_P, targ_P)
    v, sp, tp = _dev(var, dt), _dev(source_P, dt), _dev(targ_P, dt)
    src_1d = int(sp.dim() == 1)
    if not src_1d:
        if v.shape[0] != sp.shape[0] or v.shape[0] != tp.shape[0]:
            raise ValueError('Time dimension of input files is inconsistent!')
        if v.shape[2] != sp.shape[2] or v.shape[2] != tp.shape[2]:
            raise ValueError('Lat dimension of input files is inconsistent!')
        if v.shape[3] != sp.shape[3] or v.shape[3] != tp.shape[3]:
            raise ValueError('Lon dimension of input files is inconsistent!')
    nt, ks, ny, nx = v.shape
    out = torch.zeros_like(tp)
    err = _new_err()
    fn = getattr(N.lib, "pgw_interp_logp_" + _sfx(dt))
    N.check(fn(_p(v), _p(sp), _p(tp), _p(out), nt, ks, tp.shape[1], ny * nx, src_1d, 0,
               N.EXTRAP_MODES[extrapolate], _p(err), _stream()), "pgw_interp_logp")
    _raise_interp_errors(err)
    return _back(out, var)


def vert_interp_delta(delta, target_P, delta_sfc=None, ps_hist=None, ignore_top_pressure_error=False,
                      plev=None):
    """
    functions.py:369-431.  ``delta`` is a ``DeltaField`` (or an array [1, K, ny, nx] with
    ``plev`` given) on pressure levels in file order; the axis is flipped to ascending pressure,
    the surface value inserted (replace_delta_sfc) and the result interpolated in ln p with
    constant extrapolation.
    """
    if isinstance(delta, DeltaField):
        plev = delta.coords[PLEV_GCM]
        delta = delta.values
    delta_in = delta
    if plev is None:
        raise ValueError("plev (pressure levels of the delta) is required")
    dt = _work_dtype(delta, target_P)
    d = _dev(delta, dt).flip(1).contiguous()                        # :383-384
    pl = torch.as_tensor(np.asarray(plev, dtype=np.float64)[::-1].copy(), device=d.device, dtype=dt)
    tp = _dev(target_P, dt)
    nt, K, ny, nx = d.shape
    if delta_sfc is not None:                                       # :395-404
        sfc = _dev(_raw(delta_sfc.values if isinstance(delta_sfc, DeltaField) else delta_sfc), dt).reshape(nt, ny * nx)
        psh = _dev(_raw(ps_hist.values if isinstance(ps_hist, DeltaField) else ps_hist), dt).reshape(nt, ny * nx)
        sp = torch.empty_like(d)
        d2 = torch.empty_like(d)
        err = _new_err()
        fn = getattr(N.lib, "pgw_replace_delta_sfc_" + _sfx(dt))
        for t in range(nt):
            N.check(fn(_p(pl), _p(psh[t]), _p(d[t]), _p(sfc[t]), _p(sp[t]), _p(d2[t]), K, ny * nx, 1, _p(err),
                       _stream()), "pgw_replace_delta_sfc")
        if int(err.item()) & N.ERR_PS_HIST_RANGE:
            raise ValueError()
        d, src = d2, sp
        min_src = float(sp.min().item())
    else:
        src = pl
        min_src = float(pl.min().item())
    if float(tp.min().item()) < min_src and not ignore_top_pressure_error:   # :417-425
        raise ValueError(_MSG_TOP)
    res = interp_logp_4d(d, src, tp, extrapolate='constant')       # :429
    return _back(res, delta_in)


def load_delta_interp(delta_input_dir, var_name, target_P, era5_date_time, target_date_time,
                      ignore_top_pressure_error=False):
    """functions.py:306-340"""
    delta = load_delta(delta_input_dir, var_name, era5_date_time, target_date_time)
    if var_name in ['ta', 'hur']:
        delta_sfc = load_delta(delta_input_dir, var_name + 's', era5_date_time, target_date_time)
        ps_hist = load_delta(delta_input_dir, 'ps', era5_date_time, target_date_time,
                             name_base=file_name_bases['HIST'])
    else:
        delta_sfc = ps_hist = None
    return vert_interp_delta(delta, target_P, delta_sfc, ps_hist, ignore_top_pressure_error)


def determine_p_ref(p_min_era, p_min_pgw, p_ref_opts, p_ref_last=None):
    """functions.py:583-598 (scalar host logic)."""
    for p in p_ref_opts:
        if (p_min_era > p) & (p_min_pgw > p):
            if p_ref_last is None:
                return p
            return min(p, p_ref_last)


def integrate_tos(tos_field, ts_field, land_frac, ice_frac):
    """functions.py:1145-1186"""
    dt = _work_dtype(tos_field, ts_field, land_frac, ice_frac)
    tos = _dev(tos_field, dt)
    ts = _dev(ts_field, dt).reshape(tos.shape)
    land = _dev(land_frac, dt).reshape(tos.shape)
    ice = _dev(ice_frac, dt).reshape(tos.shape)
    out = torch.empty_like(tos)
    fn = getattr(N.lib, "pgw_integrate_tos_" + _sfx(dt))
    N.check(fn(_p(tos), _p(ts), _p(land), _p(ice), _p(out), tos.numel(), _stream()), "pgw_integrate_tos")
    return _back(out, tos_field)


# --------------------------------------------------------------------------- step_02: smoothing
def smooth_annual_cycle(diff):
    """Array form of filter_data (functions.py:639-667): [nt, ...] -> same shape, float32."""
    d = _dev(diff, torch.float32)
    nt = d.shape[0]
    out = torch.empty_like(d)
    N.check(N.lib.pgw_smooth_harmonic_f32(_p(d), _p(out), nt, d.numel() // nt, _stream()),
            "pgw_smooth_harmonic_f32")
    return _back(out, diff)


def harmonic_ac_analysis(ts):
    """functions.py:678-740 for one series."""
    ts = np.asarray(_raw(ts))
    return smooth_annual_cycle(ts.reshape(-1, 1)).reshape(-1).astype(ts.dtype if ts.dtype.kind == 'f' else np.float64)


def filter_data(annualcycleraw, variablename_to_smooth, outputpath):
    """functions.py:606-675: smooth the annual cycle of one variable of a NetCDF file."""
    ds = ncio.open_dataset(annualcycleraw)
    var = ds[variablename_to_smooth]
    data = np.squeeze(var.data)
    dims = tuple(d for d, n in zip(var.dims, var.data.shape) if n != 1)
    print('Dimension that is assumed to be time dimension is called: ', dims[0])
    print('shape of data: ', data.shape)
    if data.ndim not in (3, 4):
        raise SystemExit('Wrong dimensions of input file should be 3 or 4-D')
    smooth = smooth_annual_cycle(data).astype(data.dtype if data.dtype.kind == 'f' else np.float32)
    print('Done with smoothing')
    out = ncio.Dataset()
    for d in dims:
        if d in ds:
            out[d] = ds[d]
    out[variablename_to_smooth] = ncio.Variable(dims, smooth, var.attrs)
    out.to_netcdf(outputpath, mode='w')


# --------------------------------------------------------------------------- step_02: regridding
def _interp_table(x, x_new):
    """scipy interp1d(kind='linear') bracket choice (searchsorted, clipped) and the weight of
    the upper node; out-of-range targets get NaN like xarray's default fill value."""
    x = np.asarray(x, dtype=np.float64)
    x_new = np.asarray(x_new, dtype=np.float64)
    idx = np.searchsorted(x, x_new).clip(1, len(x) - 1).astype(np.int64)
    lo, hi = idx - 1, idx
    w = (x_new - x[lo]) / (x[hi] - x[lo])
    w[(x_new < x[0]) | (x_new > x[-1])] = np.nan
    return lo, hi, w


def regrid_tables(lat_gcm, lon_gcm, targ_lat, targ_lon):
    """
    Host half of regrid_lat_lon's xarray-only branch (functions.py:812-892): fold the latitude
    flip, the pole rows (zonal mean of the nearest row) and the periodic longitude into index
    and weight tables.  Returns dict(j0, j1, wy, i0, i1, wx): source rows (-1 / -2 = zonal mean
    of source row 0 / last row), source columns and the weights of j1 / i1.
    """
    lat = np.asarray(lat_gcm, dtype=np.float64)
    lon = np.asarray(lon_gcm, dtype=np.float64)
    targ_lat = np.asarray(targ_lat, dtype=np.float64)
    targ_lon = np.asarray(targ_lon, dtype=np.float64)
    ny_s, nx_s = len(lat), len(lon)
    dlon_gcm = np.median(np.diff(lon))                                         # :778-779, before any flip
    dlat_gcm = np.median(np.diff(lat))
    periodic_lon = (dlon_gcm + np.max(lon) - np.min(lon)) >= 359.9              # :780-789
    rows = np.arange(ny_s)
    if lat[0] > lat[-1]:                                                       # :822-829
        lat, rows = lat[::-1], rows[::-1]
    if np.max(targ_lat) + dlat_gcm > 89.9:                                     # :833-837
        lat = np.concatenate([lat, [90.0]])
        rows = np.concatenate([rows, [-2 if rows[-1] == ny_s - 1 else -1]])
    if np.min(targ_lat) - dlat_gcm < -89.9:                                    # :838-842
        lat = np.concatenate([[-90.0], lat])
        rows = np.concatenate([[-1 if rows[0] == 0 else -2], rows])
    if (np.max(targ_lat) > np.max(lat)) or (np.min(targ_lat) < np.min(lat)):   # :845-856
        print('GCM lat: min {} max {}'.format(np.min(lat), np.max(lat)))
        print('ERA5 lat: min {} max {}'.format(np.min(targ_lat), np.max(targ_lat)))
        raise ValueError('ERA5 dataset extends further North or South than GCM dataset!. Perhaps '
                         'consider using ERA5 on a subdomain only if global coverage is not required?')
    jlo, jhi, wy = _interp_table(lat, targ_lat)                                # :859
    cols = np.arange(nx_s)
    if periodic_lon:                                                           # :866-874
        if np.max(targ_lon) > np.max(lon):
            cols = np.concatenate([cols, cols])
            lon = np.concatenate([lon, lon + 360])
        if np.min(targ_lon) < np.min(lon):
            cols = np.concatenate([cols, cols])
            lon = np.concatenate([lon - 360, lon])
    if (np.max(targ_lon) > np.max(lon)) or (np.min(targ_lon) < np.min(lon)):   # :877-888
        print('GCM lon: min {} max {}'.format(np.min(lon), np.max(lon)))
        print('ERA5 lon: min {} max {}'.format(np.min(targ_lon), np.max(targ_lon)))
        raise ValueError('ERA5 dataset extends further East or West than GCM dataset!. Perhaps '
                         'consider using ERA5 on a subdomain only if global coverage is not required?')
    ilo, ihi, wx = _interp_table(lon, targ_lon)                                # :892
    return dict(j0=rows[jlo].astype(np.int32), j1=rows[jhi].astype(np.int32), wy=wy,
                i0=cols[ilo].astype(np.int32), i1=cols[ihi].astype(np.int32), wx=wx)


_REGRID_TABLES = {}


def regrid_arrays(data, lat_gcm, lon_gcm, targ_lat, targ_lon, rows=None):
    """
    regrid_lat_lon on arrays: data [..., nlat_gcm, nlon_gcm] -> [..., len(targ_lat),
    len(targ_lon)] float32.  Tables from ``regrid_tables``; the gather (latitude pass, then
    longitude pass, float64 arithmetic) runs in pgw_regrid_bilinear_band_f32.  ``rows = (r0, r1)``:
    only that band of target rows is produced ([..., r1 - r0, len(targ_lon)]), bit-identical to the same
    rows of the whole field -- one variable split over several GPUs by target latitude
    (``parallel.regrid_banded``).
    """
    d = _dev(data, torch.float32)
    ny_s, nx_s = d.shape[-2:]
    lead = tuple(d.shape[:-2])
    nfield = int(np.prod(lead)) if lead else 1
    dev = d.device
    # the tables depend on the four coordinate axes only: built once per grid pair and device (a daily 3-D
    # variable, its HIST and SCEN-HIST files and every band of a multi-GPU split share them)
    key = tuple(np.asarray(_raw(a), dtype=np.float64).tobytes() for a in (lat_gcm, lon_gcm, targ_lat, targ_lon)) + (str(dev),)
    hit = _REGRID_TABLES.get(key)
    if hit is None:
        tb = regrid_tables(lat_gcm, lon_gcm, targ_lat, targ_lon)
        ti = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev, dtype=torch.int32)
        tf = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev, dtype=torch.float64)
        hit = (ti(tb["j0"]), ti(tb["j1"]), ti(tb["i0"]), ti(tb["i1"]), tf(tb["wy"]), tf(tb["wx"]))
        if len(_REGRID_TABLES) >= 8:
            _REGRID_TABLES.clear()
        _REGRID_TABLES[key] = hit
    j0, j1, i0, i1, wyd, wxd = hit
    ny_t, nx_t = wyd.numel(), wxd.numel()
    r0, r1 = (0, ny_t) if rows is None else (int(rows[0]), int(rows[1]))
    if not 0 <= r0 < r1 <= ny_t:
        raise ValueError("rows %r outside the %d target rows" % (rows, ny_t))
    pm = torch.empty((nfield, 2), device=dev, dtype=torch.float32)
    out = torch.empty(lead + (r1 - r0, nx_t), device=dev, dtype=torch.float32)
    st = _stream()
    N.check(N.lib.pgw_zonal_mean_f32(_p(d), _p(pm), nfield, ny_s, nx_s, st), "pgw_zonal_mean_f32")
    N.check(N.lib.pgw_regrid_bilinear_band_f32(_p(d), _p(out), _p(pm), nfield, ny_s, nx_s, ny_t, nx_t, r0, r1,
                                               _p(j0), _p(j1), _p(wyd), _p(i0), _p(i1), _p(wxd), st),
            "pgw_regrid_bilinear_band_f32")
    return _back(out, data)


def regrid_lat_lon(ds_gcm, ds_era5, var_name, method='bilinear', i_use_xesmf=0):
    """functions.py:748-898 on ``ncio.Dataset`` objects; returns a new Dataset on the ERA5 grid."""
    if i_use_xesmf:
        raise NotImplementedError("xESMF regridding is not part of the CUDA path (SURVEY.md 2, #16)")
    var = ds_gcm[var_name]
    if var.dims[-2:] != (LAT_GCM, LON_GCM):
        raise ValueError("variable %s must end in (%s, %s)" % (var_name, LAT_GCM, LON_GCM))
    targ_lat = np.asarray(ds_era5[LAT_ERA].data, dtype=np.float64)
    targ_lon = np.asarray(ds_era5[LON_ERA].data, dtype=np.float64)
    out_data = regrid_arrays(var.data, ds_gcm[LAT_GCM].data, ds_gcm[LON_GCM].data, targ_lat, targ_lon)
    out = ncio.Dataset(attrs=ds_gcm.attrs)
    for name, v in ds_gcm.variables.items():
        if name in (var_name, LAT_GCM, LON_GCM):
            continue
        if LAT_GCM in v.dims or LON_GCM in v.dims:
            continue
        out[name] = v
    out[LAT_GCM] = ncio.Variable((LAT_GCM,), targ_lat, ds_gcm[LAT_GCM].attrs)
    out[LON_GCM] = ncio.Variable((LON_GCM,), targ_lon, ds_gcm[LON_GCM].attrs)
    out[var_name] = ncio.Variable(var.dims, out_data, var.attrs)
    return out


def lonlat_to_meter(lon_deg, lat_deg, half_turn=False):
    """The coordinate mapping of nan_ignoring_interp (functions.py:946-973, :1006-1022) on the device:
    longitudes above 180 are shifted by -360, then ``lat_m = sign(lat) * geod.inv(lon, 0, lon, lat)`` and
    ``lon_m = sign(lon) * geod.inv(0, lat, lon, lat)`` on WGS84 (+ ``geod.inv(0, lat, 180, lat)``).
    Returns float64 device tensors."""
    lat = _dev(lat_deg, torch.float64).reshape(-1)
    lon = _dev(lon_deg, torch.float64).reshape(-1)
    lat_m, lon_m = torch.empty_like(lat), torch.empty_like(lat)
    ht = torch.empty_like(lat) if half_turn else None
    N.check(N.lib.pgw_geod_to_meter_f64(_p(lat), _p(lon), _p(lat_m), _p(lon_m), _p(ht), lat.numel(), _stream()),
            "pgw_geod_to_meter_f64")
    return (lat_m, lon_m, ht) if half_turn else (lat_m, lon_m)


def nan_ignoring_interp_arrays(land_fr, era5_lat, era5_lon, values, gcm_lat2d, gcm_lon2d, kernel_radius, sharpness):
    """
    nan_ignoring_interp (functions.py:900-1060) on arrays, for any number of fields (months) at once:
    ``values`` [nt, nj, ni] (or [nj, ni]) on the curvilinear ocean grid ``gcm_lat2d/gcm_lon2d`` [nj, ni]
    -> float64 [nt, len(era5_lat), len(era5_lon)].  NaN source points are ignored per field, targets with
    ``land_fr > 0.7`` and targets without a source point inside ``kernel_radius`` become NaN.
    """
    vals = _dev(values, torch.float64)
    single = vals.dim() == 2
    vals = vals.reshape(-1, vals.shape[-2] * vals.shape[-1])
    nt, npts = vals.shape
    glat = _dev(gcm_lat2d, torch.float64).reshape(-1)
    glon = _dev(gcm_lon2d, torch.float64).reshape(-1)
    if glat.numel() != npts or glon.numel() != npts:
        raise ValueError("latitude/longitude of the ocean grid must be 2-D like the data")   # see :940-943
    keep = ~torch.isnan(vals).all(dim=0)                   # points that are NaN in every field never count
    glat, glon, vals = glat[keep], glon[keep], vals[:, keep]
    if glat.numel() == 0:                                  # no source point at all: everything is the null value
        ny0, nx0 = np.size(_raw(era5_lat)), np.size(_raw(era5_lon))
        full = torch.full((nt, ny0, nx0), float("nan"), device=vals.device, dtype=torch.float64)
        return _back(full[0] if single else full, values)
    lat_m, lon_m, off = lonlat_to_meter(glon, glat, half_turn=True)
    # boundary points: the whole cloud once more to the west and to the east (:978-988)
    lat_bd = torch.cat([lat_m, lat_m, lat_m])
    lon_bd = torch.cat([lon_m - 2.0 * off, lon_m, lon_m + 2.0 * off])
    lat_bd, order = torch.sort(lat_bd)
    lon_bd = lon_bd[order].contiguous()
    n = lat_m.numel()
    era5_lat = np.asarray(_raw(era5_lat), dtype=np.float64).reshape(-1)
    era5_lon = np.asarray(_raw(era5_lon), dtype=np.float64).reshape(-1)
    ny, nx = len(era5_lat), len(era5_lon)
    e_lat_m, e_lon_m = lonlat_to_meter(np.tile(era5_lon, ny), np.repeat(era5_lat, nx))     # :1011-1022
    land = _dev(land_fr, torch.float32).reshape(-1).contiguous()
    if land.numel() != ny * nx:
        raise ValueError("land fraction must be on the ERA5 grid")
    out = torch.empty((nt, ny * nx), device=vals.device, dtype=torch.float64)
    for f0 in range(0, nt, 12):
        f1 = min(nt, f0 + 12)
        v3 = vals[f0:f1].repeat(1, 3)[:, order].contiguous()
        o = torch.empty((f1 - f0, ny * nx), device=vals.device, dtype=torch.float64)
        N.check(N.lib.pgw_gauss_interp_f64(_p(lat_bd), _p(lon_bd), _p(v3), 3 * n, f1 - f0, _p(e_lat_m), _p(e_lon_m),
                                           _p(land), _p(o), ny * nx, float(kernel_radius), float(sharpness),
                                           _stream()), "pgw_gauss_interp_f64")
        out[f0:f1] = o
    out = out.reshape(nt, ny, nx)
    return _back(out[0] if single else out, values)


def nan_ignoring_interp(da_era5_land_fr, da_delta, kernel_radius, sharpness):
    """functions.py:900-1060 with the reference's signature, for array-likes that carry coordinates the
    way xarray DataArrays do (``.values`` and ``.coords[name].values``)."""
    return nan_ignoring_interp_arrays(
        da_era5_land_fr.values, da_era5_land_fr.coords[LAT_ERA].values, da_era5_land_fr.coords[LON_ERA].values,
        da_delta.values, da_delta.coords[LAT_GCM_OCEAN].values, da_delta.coords[LON_GCM_OCEAN].values,
        kernel_radius, sharpness)


def interp_wrapper(origin_grid, target_grid, var_name, i_use_xesmf=0,
                   nan_interp_kernel_radius=300000, nan_interp_sharpness=3):
    """functions.py:1062-1141 on ``ncio.Dataset`` objects: ``tos`` and ``siconc`` (curvilinear ocean grid
    with NaNs over land) take the NaN-ignoring Gaussian-kernel scheme (:900-1060), everything else the
    bilinear regridding."""
    if var_name in ['tos', 'siconc']:
        land_fraction = np.asarray(target_grid["FR_LAND"].data)[0]                           # :1096
        var = origin_grid[var_name]
        lat_t = np.asarray(target_grid[LAT_ERA].data, dtype=np.float64)
        lon_t = np.asarray(target_grid[LON_ERA].data, dtype=np.float64)
        result = nan_ignoring_interp_arrays(land_fraction, lat_t, lon_t, var.data,
                                            origin_grid[LAT_GCM_OCEAN].data, origin_grid[LON_GCM_OCEAN].data,
                                            nan_interp_kernel_radius, nan_interp_sharpness)  # :1100-1107
        ds = ncio.Dataset(attrs=dict(description=str(var_name) + " on ERA5 grid", units="K",
                                     long_name=str(var_name)))                               # :1110-1135
        ds[TIME_GCM_OCEAN] = origin_grid[TIME_GCM_OCEAN]
        ds["lat"] = ncio.Variable(("lat",), lat_t, target_grid[LAT_ERA].attrs)
        ds["lon"] = ncio.Variable(("lon",), lon_t, target_grid[LON_ERA].attrs)
        ds[var_name] = ncio.Variable((TIME_GCM_OCEAN, "lat", "lon"), np.asarray(result, dtype=np.float64))
        return ds
    return regrid_lat_lon(origin_grid, target_grid, var_name, method='bilinear', i_use_xesmf=i_use_xesmf)
