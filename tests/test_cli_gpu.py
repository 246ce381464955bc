"""The drop-in command lines end to end: NetCDF in -> CUDA pass -> NetCDF out, checked against
the oracle (step_03) and against the array operators (step_02)."""
import os
from datetime import datetime

import numpy as np
import pytest
import torch

from cases import TOL, make_case, run_oracle
from pgw4era5_b200 import ncio, settings, synthetic as S

pytestmark = pytest.mark.gpu


def _write_era(path, era, when):
    e = S.to_numpy(era)
    ds = ncio.Dataset()
    hours = (np.datetime64(when, "ns") - np.datetime64("2006-08-01T00", "ns")) / np.timedelta64(1, "h")
    ds["time"] = ncio.Variable(("time",), np.array([float(hours)]), {"units": "hours since 2006-08-01 00:00:00"})
    ds["lat"] = ncio.Variable(("lat",), e["lat"]); ds["lon"] = ncio.Variable(("lon",), e["lon"])
    ds["level"] = ncio.Variable(("level",), np.arange(1, 138, dtype=np.float64))
    ds["level1"] = ncio.Variable(("level1",), np.arange(1, 139, dtype=np.float64))
    ds["soil1"] = ncio.Variable(("soil1",), e["soil1"])
    ds["ak"] = ncio.Variable(("level1",), e["ak"]); ds["bk"] = ncio.Variable(("level1",), e["bk"])
    for name in ("PS", "FIS", "FR_LAND", "FR_SEA_ICE", "T_SKIN"):
        ds[name] = ncio.Variable(("time", "lat", "lon"), e[name])
    ds["T_SO"] = ncio.Variable(("time", "soil1", "lat", "lon"), e["T_SO"])
    for name in ("T", "QV", "U", "V"):
        ds[name] = ncio.Variable(("time", "level", "lat", "lon"), e[name])
    ds.to_netcdf(path)


def _write_deltas(ddir, deltas, lat, lon):
    d = S.to_numpy(deltas)
    for name, v in d.items():
        var, base = ("ps", settings.file_name_bases["HIST"]) if name == "ps_hist" else (name, settings.file_name_bases["SCEN-HIST"])
        ds = ncio.Dataset()
        ds["time"] = ncio.encode_time(v["time"], "days since 1850-01-01 00:00:00")
        ds["lat"] = ncio.Variable(("lat",), lat); ds["lon"] = ncio.Variable(("lon",), lon)
        if v["plev"] is not None:
            ds["plev"] = ncio.Variable(("plev",), v["plev"])
            ds[var] = ncio.Variable(("time", "plev", "lat", "lon"), v["data"])
        else:
            ds[var] = ncio.Variable(("time", "lat", "lon"), v["data"])
        ds.to_netcdf(os.path.join(ddir, base.format(var)))
        if name == "ps_hist":      # step_02 writes HIST and SCEN-HIST for every variable; -D interpolate_time
            ds.to_netcdf(os.path.join(ddir, settings.file_name_bases["SCEN-HIST"].format(var)))   # reads ps_delta


def test_step_03_cli_matches_oracle(tmp_path):
    from pgw4era5_b200 import step_03_apply_to_era as S3
    when = datetime(2006, 8, 2, 6)
    era, deltas = make_case(10, 24, 51)
    inp, out, dd = tmp_path / "in", tmp_path / "out", tmp_path / "deltas"
    for p in (inp, dd):
        p.mkdir()
    name = settings.era5_file_name_base.format(when)
    _write_era(str(inp / name), era, when)
    _write_deltas(str(dd), deltas, era["lat"], era["lon"])
    old = settings.i_debug
    settings.i_debug = 0
    try:
        n_iters = S3.main(["-i", str(inp), "-o", str(out), "-d", str(dd), "-f", "2006080206", "-l", "2006080206",
                           "-H", "6", "-t"])
    finally:
        settings.i_debug = old
    ref = run_oracle(era, deltas, when=when)
    assert n_iters == [ref["n_iter"]]
    res = ncio.open_dataset(str(out / name))
    for key in ("PS", "T", "QV", "U", "V", "T_SKIN", "T_SO", "FR_SEA_ICE"):
        g = res[key].data.astype(np.float64)
        r = np.asarray(ref[key], dtype=np.float64).reshape(g.shape)
        assert np.array_equal(np.isnan(g), np.isnan(r))
        assert np.nanmax(np.abs(g - r)) <= TOL[key], key
    assert "RELHUM" not in res and res["T"].dims == ("time", "level", "lat", "lon")
    # --debug_mode interpolate_time writes the time-interpolated deltas (step_03:387-414)
    S3.main(["-i", str(inp), "-o", str(out), "-d", str(dd), "-f", "2006080206", "-l", "2006080206", "-H", "6",
             "-D", "interpolate_time"])
    from oracle import pgw_oracle as O
    dta = ncio.open_dataset(str(out / ("delta_ta_" + name)))["ta"].data
    np.testing.assert_allclose(dta, O.load_delta(S.to_numpy(deltas)["ta"], when), rtol=0, atol=1e-6)


def test_step_02_cli(tmp_path):
    from oracle import pgw_oracle as O
    from pgw4era5_b200 import step_02_preproc_deltas as S2
    rng = np.random.default_rng(61)
    lat, lon = np.linspace(-85, 85, 18), np.arange(0.0, 360, 10.0)
    tlat, tlon = np.linspace(-90, 90, 37), np.arange(0.0, 360, 5.0)
    gdir, sdir, rdir = tmp_path / "gcm", tmp_path / "smooth", tmp_path / "regrid"
    gdir.mkdir()
    stamps = np.datetime64("2001-01-01T12", "ns") + np.arange(365) * np.timedelta64(86400 * 10 ** 9, "ns")
    data = (rng.normal(size=(365, 3, 18, 36)) + 2).astype(np.float32)
    for base in settings.file_name_bases.values():
        ds = ncio.Dataset()
        ds["time"] = ncio.encode_time(stamps, "days since 1850-01-01 00:00:00")
        ds["plev"] = ncio.Variable(("plev",), np.array([85000., 50000., 30000.]))
        ds["lat"] = ncio.Variable(("lat",), lat); ds["lon"] = ncio.Variable(("lon",), lon)
        ds["ta"] = ncio.Variable(("time", "plev", "lat", "lon"), data)
        ds.to_netcdf(str(gdir / base.format("ta")))
    era = ncio.Dataset()
    era["lat"] = ncio.Variable(("lat",), tlat); era["lon"] = ncio.Variable(("lon",), tlon)
    era.to_netcdf(str(tmp_path / "era.nc"))
    S2.main(["smoothing", "-i", str(gdir), "-o", str(sdir), "-e", str(tmp_path / "era.nc"), "-v", "ta"])
    sm = ncio.open_dataset(str(sdir / "ta_delta.nc"))["ta"].data
    ref_s = O.filter_data_fast(data)
    np.testing.assert_allclose(sm, ref_s, rtol=0, atol=2e-6)
    S2.main(["regridding", "-i", str(sdir), "-o", str(rdir), "-e", str(tmp_path / "era.nc"), "-v", "ta"])
    rg = ncio.open_dataset(str(rdir / "ta_historical.nc"))
    assert rg["ta"].data.shape == (365, 3, 37, 72)
    np.testing.assert_allclose(rg["ta"].data, O.regrid_lat_lon(sm, lat, lon, tlat, tlon), rtol=0, atol=1e-6)
    np.testing.assert_array_equal(rg["lat"].data, tlat)
    # -p 2: the (variable, period) files dealt out to two worker processes (worker i on GPU i % device_count)
    out = S2.main(["regridding", "-i", str(sdir), "-o", str(tmp_path / "regrid_p2"), "-e", str(tmp_path / "era.nc"),
                   "-v", "ta", "-p", "2"])
    assert [os.path.basename(o) for o in out] == ["ta_historical.nc", "ta_delta.nc"]
    for name in ("ta_historical.nc", "ta_delta.nc"):
        np.testing.assert_array_equal(ncio.open_dataset(str(tmp_path / "regrid_p2" / name))["ta"].data,
                                      ncio.open_dataset(str(rdir / name))["ta"].data)


def test_step_02_cli_ocean_variables(tmp_path):
    """`regridding -v tos`: the curvilinear ocean grid goes through the NaN-ignoring Gaussian-kernel scheme
    (functions.py:900-1060, :1093-1135); 12 monthly fields, NaN over GCM land and over ERA5 land."""
    from oracle import pgw_oracle as O
    from pgw4era5_b200 import step_02_preproc_deltas as S2
    from test_nanterp import _ocean_case
    glat, glon, vals3, tlat, tlon, land_fr, _ = _ocean_case(8, False)
    rng = np.random.default_rng(62)
    vals = np.concatenate([vals3] * 4, axis=0) + rng.normal(size=(12, 1, 1)).astype(np.float32)
    gdir, rdir = tmp_path / "gcm", tmp_path / "regrid"
    gdir.mkdir()
    stamps = S.monthly_stamps()
    for base in settings.file_name_bases.values():
        ds = ncio.Dataset()
        ds["time"] = ncio.encode_time(stamps, "days since 1850-01-01 00:00:00")
        ds["latitude"] = ncio.Variable(("j", "i"), glat); ds["longitude"] = ncio.Variable(("j", "i"), glon)
        ds["tos"] = ncio.Variable(("time", "j", "i"), vals)
        ds.to_netcdf(str(gdir / base.format("tos")))
    era = ncio.Dataset()
    era["time"] = ncio.Variable(("time",), np.array([0.0]))
    era["lat"] = ncio.Variable(("lat",), tlat); era["lon"] = ncio.Variable(("lon",), tlon)
    era["FR_LAND"] = ncio.Variable(("time", "lat", "lon"), land_fr[None])
    era.to_netcdf(str(tmp_path / "era.nc"))
    S2.main(["regridding", "-i", str(gdir), "-o", str(rdir), "-e", str(tmp_path / "era.nc"), "-v", "tos"])
    rg = ncio.open_dataset(str(rdir / "tos_delta.nc"))
    assert rg["tos"].dims == ("time", "lat", "lon") and rg["tos"].data.shape == (12, len(tlat), len(tlon))
    np.testing.assert_array_equal(rg["lat"].data, tlat)
    for m in (0, 1, 11):
        ref = O.nan_ignoring_interp(land_fr, tlat, tlon, vals[m], glat.copy(), glon.copy(),
                                    settings.nan_interp_kernel_radius, settings.nan_interp_sharpness)
        np.testing.assert_allclose(rg["tos"].data[m], ref, rtol=0, atol=1e-9, equal_nan=True)
        assert np.all(np.isnan(rg["tos"].data[m][land_fr > 0.7]))


def test_step_03_cli_pipelines_several_files(tmp_path):
    """Production mode over five files: reader thread -> HostPipeline (two CUDA streams) -> writer
    thread must give the same files as the one-file-at-a-time routine, in the right order."""
    from datetime import timedelta
    from pgw4era5_b200 import step_03_apply_to_era as S3
    t0 = datetime(2006, 8, 2, 0)
    inp, out, out1, dd = tmp_path / "in", tmp_path / "out", tmp_path / "out1", tmp_path / "deltas"
    for p in (inp, dd, out1):
        p.mkdir()
    era0, deltas = make_case(12, 20, 52)
    _write_deltas(str(dd), deltas, era0["lat"], era0["lon"])
    whens = [t0 + timedelta(hours=6 * i) for i in range(5)]
    for i, when in enumerate(whens):
        era = S.make_era5(12, 20, 520 + i, lat=era0["lat"], lon=era0["lon"], orog_seed=52)
        _write_era(str(inp / settings.era5_file_name_base.format(when)), era, when)
    old = settings.i_debug
    settings.i_debug = -1
    try:
        S3.IO_STATS.update(raw=0, decoded=0)
        n_pipe = S3.main(["-i", str(inp), "-o", str(out), "-d", str(dd), "-f", "2006080200", "-l", "2006080300",
                          "-H", "6", "-t"])
        assert S3.IO_STATS == dict(raw=5, decoded=0)        # NetCDF-3 float32 files: no decoding on the host
        os.environ["PGW_RAW_IO"] = "0"
        try:
            n_dec = S3.main(["-i", str(inp), "-o", str(tmp_path / "out_dec"), "-d", str(dd), "-f", "2006080200",
                             "-l", "2006080300", "-H", "6", "-t"])
        finally:
            del os.environ["PGW_RAW_IO"]
        assert S3.IO_STATS == dict(raw=5, decoded=5) and n_dec == n_pipe
        n_one = [S3.pgw_for_era5(str(inp / settings.era5_file_name_base.format(w)),
                                 str(out1 / settings.era5_file_name_base.format(w)), str(dd), w, True)
                 for w in whens]
    finally:
        settings.i_debug = old
    assert n_pipe == n_one and len(n_pipe) == 5
    for w in whens:
        name = settings.era5_file_name_base.format(w)
        a, b = ncio.open_dataset(str(out / name)), ncio.open_dataset(str(out1 / name))
        for key in ("PS", "T", "U", "V", "T_SKIN", "T_SO", "FR_SEA_ICE"):
            np.testing.assert_array_equal(a[key].data, b[key].data, err_msg="%s %s" % (name, key))
        # QV: the pipeline submits file i+1 before the iteration count of file i is known, so more
        # files take the rewrite path (k_spec > N), which recovers e from the fp32 QV: a few ulps
        np.testing.assert_allclose(a["QV"].data, b["QV"].data, rtol=0, atol=5e-9, err_msg=name)
        assert "RELHUM" not in a
        # raw and decoding pipelines: the same bytes in every variable, attributes kept
        c = ncio.open_dataset(str(tmp_path / "out_dec" / name))
        for key in a.keys():
            if key == "QV":          # k_spec of the second run starts from the first run's history (see above)
                np.testing.assert_allclose(a[key].data, c[key].data, rtol=0, atol=5e-9, err_msg=name)
            else:
                np.testing.assert_array_equal(a[key].data, c[key].data, err_msg="%s %s" % (name, key))
            assert a[key].attrs == c[key].attrs and a[key].dims == c[key].dims


def test_step_01_cfday_interp_to_plev(tmp_path):
    """model levels -> pressure levels (CFday_interp_to_plev.py:92-159)."""
    from oracle import pgw_oracle as O
    from pgw4era5_b200.step_01_extract_deltas import CFday_interp_to_plev as C1
    rng = np.random.default_rng(71)
    nt, nl, ny, nx = 3, 12, 5, 7
    b = np.linspace(1.0, 0.0, nl)                        # file order: surface first (descending pressure)
    ap = np.linspace(0.0, 3000.0, nl)
    ps = 1e5 + 2000 * rng.normal(size=(nt, ny, nx))
    ta = (250 + 10 * rng.normal(size=(nt, nl, ny, nx))).astype(np.float32)
    chunk = "20700101-20741231"
    name = "ta_CFday_MPI-ESM1-2-HR_ssp585_r1i1p1f1_gn_%s.nc" % chunk
    inp = tmp_path / "sub"; inp.mkdir()
    ds = ncio.Dataset()
    ds["time"] = ncio.Variable(("time",), np.arange(nt, dtype=np.float64), {"units": "days since 1850-01-01"})
    ds["lev"] = ncio.Variable(("lev",), np.arange(nl, dtype=np.float64))
    ds["lat"] = ncio.Variable(("lat",), np.linspace(-10, 10, ny)); ds["lon"] = ncio.Variable(("lon",), np.arange(nx) * 1.0)
    ds["ap"] = ncio.Variable(("lev",), ap); ds["b"] = ncio.Variable(("lev",), b)
    ds["ps"] = ncio.Variable(("time", "lat", "lon"), ps)
    ds["ta"] = ncio.Variable(("time", "lev", "lat", "lon"), ta)
    ds.to_netcdf(str(inp / name))
    targ = np.array([101000., 90000., 70000., 50000., 20000., 5000., 1000.])
    np.savetxt(str(tmp_path / "targ.dat"), targ)
    C1.main(["ta", "ssp585", "--inp_dir", str(inp), "--out_base_dir", str(tmp_path / "out"),
             "--target_p_file", str(tmp_path / "targ.dat"), "--times", chunk])
    res = ncio.open_dataset(str(tmp_path / "out" / "MPI-ESM1-2-HR" / name))
    np.testing.assert_array_equal(res["plev"].data, np.sort(targ)[::-1])
    src_p = (ap[::-1][None, :, None, None] + b[::-1][None, :, None, None] * ps[:, None])
    tp = np.broadcast_to(np.sort(targ)[None, :, None, None], (nt, len(targ), ny, nx))
    ref = O.interp_logp_4d(ta[:, ::-1].astype(np.float64), src_p, tp, extrapolate="constant")[:, ::-1]
    np.testing.assert_allclose(res["ta"].data, ref, rtol=0, atol=2e-5)
    assert res["ta"].dims == ("time", "plev", "lat", "lon")


def test_extpar_adapt(tmp_path):
    """postproc_cosmo/extpar_adapt.py:13-35: T_CL += annual mean of the ts delta."""
    from pgw4era5_b200.postproc_cosmo import extpar_adapt as E
    era, deltas = make_case(6, 9, 81)
    dd = tmp_path / "deltas"; dd.mkdir()
    _write_deltas(str(dd), deltas, era["lat"], era["lon"])
    t_cl = (280 + np.arange(54, dtype=np.float32)).reshape(6, 9)
    ext = ncio.Dataset()
    ext["rlat"] = ncio.Variable(("rlat",), era["lat"]); ext["rlon"] = ncio.Variable(("rlon",), era["lon"])
    ext["T_CL"] = ncio.Variable(("rlat", "rlon"), t_cl.copy())
    ext.to_netcdf(str(tmp_path / "extpar.nc"))
    E.main([str(tmp_path / "extpar.nc"), "-d", str(dd)])
    got = ncio.open_dataset(str(tmp_path / "extpar.nc"))["T_CL"].data
    ts = np.asarray(S.to_numpy(deltas)["ts"]["data"], dtype=np.float64)
    np.testing.assert_allclose(got, t_cl + ts.mean(axis=0), rtol=0, atol=1e-4)


def _several_files(tmp_path, n, seed=53):
    from datetime import timedelta
    t0 = datetime(2006, 8, 2, 0)
    inp, dd = tmp_path / "in", tmp_path / "deltas"
    for p in (inp, dd):
        p.mkdir()
    era0, deltas = make_case(12, 20, seed)
    _write_deltas(str(dd), deltas, era0["lat"], era0["lon"])
    whens = [t0 + timedelta(hours=6 * i) for i in range(n)]
    for i, when in enumerate(whens):
        era = S.make_era5(12, 20, 10 * seed + i, lat=era0["lat"], lon=era0["lon"], orog_seed=seed)
        _write_era(str(inp / settings.era5_file_name_base.format(when)), era, when)
    return inp, dd, whens


@pytest.mark.timeout(120)
def test_step_03_pipeline_surfaces_writer_failure(tmp_path):
    """Seven files into an output directory that does not exist: the writer thread fails on the first one.
    With only four output buffers in circulation (handed back by the writer alone) the main loop used to block
    for ever; now the error comes out and the reader thread ends too."""
    import threading
    from pgw4era5_b200 import step_03_apply_to_era as S3
    inp, dd, whens = _several_files(tmp_path, 7)
    steps = [dict(inp_era_file_path=str(inp / settings.era5_file_name_base.format(w)),
                  out_era_file_path=str(tmp_path / "no_such_dir" / settings.era5_file_name_base.format(w)),
                  era_step_dt=w) for w in whens]
    old = settings.i_debug
    settings.i_debug = -1
    before = threading.active_count()
    try:
        with pytest.raises((OSError, IOError)):
            S3.pgw_for_era5_files(steps, str(dd), True)
    finally:
        settings.i_debug = old
    import time
    time.sleep(0.6)
    assert threading.active_count() <= before + 8       # reader and writer gone (only pool threads may linger)


def test_step_03_reference_output_dtypes(tmp_path):
    """settings.i_reference_output_dtypes = 1: PS, T, QV, U, V are written as float64 like the reference's
    to_netcdf does (step_03_apply_to_era.py:367-381); skin, soil and sea ice keep the file's dtype; values are
    the float32 results widened."""
    from pgw4era5_b200 import step_03_apply_to_era as S3
    inp, dd, whens = _several_files(tmp_path, 2, seed=54)
    args = ["-i", str(inp), "-d", str(dd), "-f", "2006080200", "-l", "2006080206", "-H", "6", "-t"]
    old = settings.i_debug, settings.i_reference_output_dtypes
    settings.i_debug = -1
    try:
        S3.main(args + ["-o", str(tmp_path / "out32")])
        settings.i_reference_output_dtypes = 1
        S3.main(args + ["-o", str(tmp_path / "out64")])
    finally:
        settings.i_debug, settings.i_reference_output_dtypes = old
    for w in whens:
        name = settings.era5_file_name_base.format(w)
        a, b = ncio.open_dataset(str(tmp_path / "out32" / name)), ncio.open_dataset(str(tmp_path / "out64" / name))
        for key in ("PS", "T", "QV", "U", "V"):
            assert a[key].data.dtype == np.float32 and b[key].data.dtype == np.float64, key
            tol = 5e-9 if key == "QV" else 0.0          # k_spec history differs between the two runs (rewrite path)
            np.testing.assert_allclose(b[key].data, a[key].data.astype(np.float64), rtol=0, atol=tol)
        for key in ("T_SKIN", "T_SO", "FR_SEA_ICE"):
            assert b[key].data.dtype == np.float32
            np.testing.assert_array_equal(a[key].data, b[key].data)
