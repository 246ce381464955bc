"""oracle/xrlite.py restates the xarray (2022.12) semantics the reference relies on; these are the documented
behaviours it has to get right for the goldens to mean anything (each case names the reference line that
depends on it).  xarray itself is not installed here, so the expectations are written out by hand."""
import numpy as np
import pytest

from oracle import xrlite as xr


def da(data, dims, **coords):
    return xr.DataArray(np.asarray(data), dims=dims, coords=coords or None)


def test_broadcasting_is_by_dimension_name_first_operand_first():
    """era_file.ak + PS * era_file.bk -> (level1, time, lat, lon), then .transpose(...) (step_03:64-66)"""
    ak = da([1., 2., 3.], ("level1",), level1=[1., 2., 3.])
    ps = da(np.arange(4.).reshape(1, 2, 2), ("time", "lat", "lon"))
    r = ak + ps * ak
    assert r.dims == ("level1", "time", "lat", "lon") and r.shape == (3, 1, 2, 2)
    np.testing.assert_array_equal(r.values[:, 0, 1, 0], [1. + 2., 2. + 4., 3. + 6.])
    t = r.transpose("time", "level1", "lat", "lon")
    assert t.dims == ("time", "level1", "lat", "lon") and t.values[0, 2, 1, 1] == 3. + 9.
    with pytest.raises(ValueError):
        ak + da([1., 2., 3.], ("level1",), level1=[1., 2., 4.])          # unequal index coordinates


def test_plain_arrays_and_python_scalars_follow_numpy_promotion():
    """0.5 * diff + ak.isel(...).values is positional (step_03:73-78); float32 * python float stays float32
    (functions.py:144: the float32 virtual temperature)"""
    a = da(np.array([1., 2., 4.], np.float32), ("level",), level=[1., 2., 3.])
    r = 0.5 * a + np.array([10., 20., 30.])
    assert r.dims == ("level",) and r.dtype == np.float64 and list(r.values) == [10.5, 21., 32.]
    assert (a * (1 + 0.61 * a)).dtype == np.float32
    assert (a * np.float64(2.0)).dtype == np.float64


def test_inplace_operators_keep_the_dtype_of_the_left_operand():
    """delta_ps += adj_ps with delta_ps = zeros_like(float32 PS) (step_03:186-194)"""
    ps = da(np.array([[1e5, 9e4]], np.float32), ("lat", "lon"))
    d = xr.zeros_like(ps)
    d += da(np.array([[0.123456789, 1.0]]), ("lat", "lon"))
    assert d.dtype == np.float32 and d.values[0, 0] == np.float32(0.123456789)
    assert (ps + d).dtype == np.float32


def test_diff_rename_and_where():
    """np.log(pa_hl).diff(dim, label='lower').rename(...) and pa_hl.where(pa_hl > 0, 1e-4) (functions.py:135-138)"""
    p = da([[0., 10., 30.]], ("time", "level1"), level1=[1., 2., 3.])
    q = p.where(p > 0, 0.0001)
    assert q.values[0, 0] == 0.0001 and q.values[0, 2] == 30.
    d = q.diff(dim="level1", label="lower").rename({"level1": "level"})
    assert d.dims == ("time", "level") and list(d["level"].values) == [1., 2.]      # labels of the LOWER index
    np.testing.assert_allclose(d.values[0], [10. - 0.0001, 20.])
    alpha = xr.where(q >= 10., 1, xr.full_like(q, np.nan))
    assert alpha.values[0, 1] == 1 and np.isnan(alpha.values[0, 0])


def test_expand_dims_loc_and_sel():
    """phi_hl = zgs.expand_dims(dim={HLEV: level1}).copy(); phi_hl.loc[{HLEV: l}] = phi_hl.sel({HLEV: l+1}) + ...
    (functions.py:141-152): the new dimension leads, the dtype is that of zgs, labels select exactly"""
    zgs = da(np.array([[[5., 6.]]], np.float32), ("time", "lat", "lon"))
    lev = da([1., 2., 3.], ("level1",), level1=[1., 2., 3.])
    phi = zgs.expand_dims(dim={"level1": lev}).copy()
    assert phi.dims == ("level1", "time", "lat", "lon") and phi.dtype == np.float32
    phi.loc[{"level1": 2.0}] = phi.sel({"level1": 3.0}) + da([[[0.5, 0.25]]], ("time", "lat", "lon"))
    np.testing.assert_array_equal(phi.values[1, 0, 0], [5.5, 6.25])
    np.testing.assert_array_equal(phi.values[0, 0, 0], [5., 6.])
    s = phi.sel({"level1": 2.0})
    assert s.dims == ("time", "lat", "lon") and float(s["level1"].values) == 2.0    # scalar coordinate kept
    with pytest.raises(KeyError):
        phi.sel({"level1": 2.5})
    t = da([[1., 2.]], ("time", "lat"), time=[np.datetime64("2000-01-16", "ns")])
    one = t.isel(time=0)
    assert one.dims == ("lat",) and one.expand_dims(dim="time", axis=0).dims == ("time", "lat")


def test_vectorised_indexers_are_pointwise():
    """ind = p_diff.argmin(dim=HLEV); p_diff[HLEV].isel({HLEV: ind}); pa_hl.sel({HLEV: hl_ref_star})
    (functions.py:160-171): dimensions shared by array and indexer are paired element by element"""
    x = da(np.arange(24.).reshape(1, 4, 2, 3), ("time", "level1", "lat", "lon"), level1=[1., 2., 3., 4.])
    pd_ = x - 9.0
    pd_ = pd_.where(pd_ >= 0, np.nan)
    ind = pd_.argmin(dim="level1")                      # NaN-skipping
    assert ind.dims == ("time", "lat", "lon")
    np.testing.assert_array_equal(ind.values[0], [[2, 2, 2], [1, 1, 1]])
    lab = pd_["level1"].isel({"level1": ind})
    assert lab.dims == ("time", "lat", "lon") and lab.values[0, 1, 2] == 2.0
    picked = x.sel({"level1": lab})
    assert picked.dims == ("time", "lat", "lon")
    np.testing.assert_array_equal(picked.values[0], [[12., 13., 14.], [9., 10., 11.]])
    assert "level1" in picked.coords and picked["level1"].dims == ("time", "lat", "lon")
    with pytest.raises(ValueError):                     # an all-NaN column: "p_ref below the surface" (:162-165)
        (x.where(x < 0, np.nan)).argmin(dim="level1")


def test_reductions_skip_nan_and_numpy_functions_dispatch():
    """np.abs(err).max().values, np.min(target_P) < np.min(source_P), .mean(dim=[time]) (step_03:134-136,:308)"""
    e = da([[1., np.nan, -7.]], ("time", "lon"))
    assert float(np.abs(e).max().values) == 7.0 and float(np.min(e)) == -7.0
    assert bool(np.min(e) < np.min(e + 1))
    m = da(np.array([[1., 3.], [np.nan, 5.]], np.float32), ("time", "lon")).mean(dim=["time"])
    assert m.dims == ("lon",) and m.dtype == np.float32 and list(m.values) == [1., 4.]
    assert bool(np.any(np.isnan(e)))


def test_interp_is_scipy_interp1d_on_float_coordinates():
    """full_delta[var].interp(time=target) (functions.py:288-292): datetime coordinates become float64
    nanoseconds since the smallest stamp, float32 data is differenced in float32 (scipy)"""
    from scipy.interpolate import interp1d
    t = np.array(["2006-07-16T12", "2006-08-16T12"], dtype="datetime64[ns]")
    y = np.array([[1.1, 2.2], [3.3, 1.1]], np.float32)
    d = da(y, ("time", "lon"), time=t)
    when = np.datetime64("2006-08-02T06", "ns")
    r = d.interp(time=when, method="linear")
    x = (t - t[0]).astype(np.float64)
    ref = interp1d(x, y, axis=0, bounds_error=False)(float((when - t[0]).astype(np.float64)))
    assert r.dims == ("lon",) and np.array_equal(r.values, ref) and r["time"].values == when
    lat = da([[0., 10., 20.]], ("time", "lat"), lat=[0., 1., 2.])
    r2 = lat.interp({"lat": da([0.5, 1.5, 3.0], ("lat",), lat=[0.5, 1.5, 3.0])})
    assert list(r2.values[0][:2]) == [5., 15.] and np.isnan(r2.values[0][2])        # outside: NaN, no error


def test_concat_along_scalar_coordinate_and_existing_dimension():
    """xr.concat([before, after], dim='time') (functions.py:287) and the pole rows / periodic copies of
    regrid_lat_lon (:833-874): variables lacking a dimension are broadcast along it"""
    ds = xr.Dataset()
    ds._coords["time"] = xr.Variable(("time",), np.array(["2000-01-16", "2000-02-16"], dtype="datetime64[ns]"))
    ds._coords["lat"] = xr.Variable(("lat",), np.array([80., 85.]))
    ds._coords["lon"] = xr.Variable(("lon",), np.array([0., 120., 240.]))
    ds._vars["ta"] = xr.Variable(("time", "lat", "lon"), np.arange(12.).reshape(2, 2, 3))
    a, b = ds.isel(time=0), ds.isel(time=1)
    both = xr.concat([a, b], dim="time")
    assert both["ta"].dims == ("time", "lat", "lon") and np.array_equal(both["ta"].values, ds["ta"].values)
    north = ds.isel({"lat": -1})
    north["lat"].values = 90
    north["ta"] = north["ta"].mean(dim=["lon"])
    ext = xr.concat([ds, north], dim="lat")
    assert ext["ta"].dims == ("time", "lat", "lon") and list(ext["lat"].values) == [80., 85., 90.]
    np.testing.assert_array_equal(ext["ta"].values[0, 2], [4., 4., 4.])             # zonal mean of the last row
    shifted = ds.assign_coords({"lon": ds["lon"] + 360})
    wide = xr.concat([ds, shifted], dim="lon")
    assert list(wide["lon"].values) == [0., 120., 240., 360., 480., 600.] and wide["ta"].shape == (2, 2, 6)
    flipped = ds.reindex({"lat": list(reversed(ds["lat"]))})
    assert list(flipped["lat"].values) == [85., 80.] and flipped["ta"].values[0, 0, 0] == 3.


def test_apply_ufunc_vectorize_moves_core_dimensions_last():
    """xr.apply_ufunc(replace_delta_sfc, source_P, ps_hist, delta, delta_sfc, input_core_dims=[[plev],[],[plev],[]],
    output_core_dims=[[plev],[plev]], vectorize=True) (functions.py:396-402)"""
    P = da(np.broadcast_to(np.array([100., 500., 900.])[None, :, None], (1, 3, 2)).copy(), ("time", "plev", "lon"))
    d = da(np.arange(6.).reshape(1, 3, 2), ("time", "plev", "lon"))
    ps = da([[600., 950.]], ("time", "lon"))
    ds_ = da([[-1., -2.]], ("time", "lon"))

    def f(p, s, v, vs):                                 # what the reference's replace_delta_sfc does
        p, v = p.copy(), v.copy()
        if s > p.max():
            p[-1], v[-1] = s, vs
        else:
            i = np.max(np.argwhere(s > p))
            v[i:], p[i] = vs, s
        return p, v
    outP, outD = xr.apply_ufunc(f, P, ps, d, ds_, input_core_dims=[["plev"], [], ["plev"], []],
                                output_core_dims=[["plev"], ["plev"]], vectorize=True)
    assert outP.dims == ("time", "lon", "plev")
    np.testing.assert_array_equal(outP.values[0, 0], [100., 600., 900.])
    np.testing.assert_array_equal(outD.values[0, 0], [0., -1., -1.])
    np.testing.assert_array_equal(outP.values[0, 1], [100., 500., 950.])
    np.testing.assert_array_equal(outD.values[0, 1], [1., 3., -2.])


def test_open_dataset_decodes_cf_time_only_when_asked(tmp_path):
    """xr.open_dataset(path) for the deltas, decode_cf=False for the ERA5 file (functions.py:203, step_03:60)"""
    from scipy.io import netcdf_file
    path = str(tmp_path / "t.nc")
    with netcdf_file(path, "w", version=2) as f:
        f.createDimension("time", 2)
        v = f.createVariable("time", "d", ("time",))
        v[:] = [15.5, 45.5]
        v.units, v.calendar = "days since 2000-01-01 00:00:00", "standard"
        w = f.createVariable("ts", "f", ("time",))
        w[:] = np.array([1.5, 2.5], np.float32)
    dec, raw = xr.open_dataset(path), xr.open_dataset(path, decode_cf=False)
    assert dec["time"].values[0] == np.datetime64("2000-01-16T12", "ns") and dec["ts"].dtype == np.float32
    assert raw["time"].values[1] == 45.5 and raw["time"].attrs["units"].startswith("days since")
    import pandas as pd
    assert isinstance(dec.indexes["time"], pd.DatetimeIndex)
