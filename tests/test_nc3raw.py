"""The raw NetCDF-3 access of the file pipeline (pgw4era5_b200/nc3raw.py) against scipy's reader."""
import os

import numpy as np
import pytest

from pgw4era5_b200 import ncio
from pgw4era5_b200.nc3raw import NC_FLOAT, NotNetCDF3, RawNC3


def _make(path, nt=1):
    rng = np.random.default_rng(5)
    ds = ncio.Dataset()
    ds["time"] = ncio.Variable(("time",), np.arange(nt, dtype=np.float64), {"units": "hours since 2006-08-01 00:00:00"})
    ds["lat"] = ncio.Variable(("lat",), np.linspace(30, 80, 5)); ds["lon"] = ncio.Variable(("lon",), np.arange(7.0))
    ds["level"] = ncio.Variable(("level",), np.arange(1.0, 4.0))
    ds["ak"] = ncio.Variable(("level",), rng.normal(size=3))
    ds["PS"] = ncio.Variable(("time", "lat", "lon"), rng.normal(size=(nt, 5, 7)).astype(np.float32), {"units": "Pa"})
    ds["T"] = ncio.Variable(("time", "level", "lat", "lon"), rng.normal(size=(nt, 3, 5, 7)).astype(np.float32))
    ds["FIS"] = ncio.Variable(("lat", "lon"), rng.normal(size=(5, 7)).astype(np.float32))
    ds["count"] = ncio.Variable(("time",), np.arange(nt, dtype=np.int32))
    ds.attrs["title"] = "raw access test"
    ds.to_netcdf(path)
    return ds


@pytest.mark.parametrize("nt", [1, 3])
def test_header_and_raw_reads_match_scipy(tmp_path, nt):
    path = str(tmp_path / "a.nc")
    ds = _make(path, nt)
    raw = RawNC3(path)
    assert raw.numrecs == nt and raw.dims["lat"] == 5 and raw.dims["time"] == nt
    assert raw.vars["T"].nc_type == NC_FLOAT and raw.vars["T"].is_record and not raw.vars["FIS"].is_record
    assert raw.vars["T"].shape == (nt, 3, 5, 7) and raw.vars["T"].record_shape == (3, 5, 7)
    with open(path, "rb", buffering=0) as f:
        for name in ("PS", "T", "FIS", "ak"):
            v = raw.vars[name]
            for rec in range(nt if v.is_record else 1):
                buf = np.empty(v.record_shape, dtype=v.dtype)
                raw.read_into(f, name, buf, record=rec)
                want = ds[name].data[rec] if v.is_record else ds[name].data
                np.testing.assert_array_equal(buf.astype(buf.dtype.newbyteorder("=")), want)


def test_patch_in_place(tmp_path):
    path = str(tmp_path / "b.nc")
    ds = _make(path, 2)
    raw = RawNC3(path)
    new = (np.arange(3 * 5 * 7, dtype=np.float32).reshape(3, 5, 7) + 0.5).astype(">f4")
    fd = os.open(path, os.O_RDWR)
    try:
        raw.write_from(fd, "T", new, record=1)
    finally:
        os.close(fd)
    back = ncio.open_dataset(path)
    np.testing.assert_array_equal(back["T"].data[1], new.astype(np.float32))
    np.testing.assert_array_equal(back["T"].data[0], ds["T"].data[0])          # the other record is untouched
    np.testing.assert_array_equal(back["PS"].data, ds["PS"].data)
    assert back["PS"].attrs["units"] == "Pa" and back.attrs["title"] == "raw access test"


def test_rejects_other_files(tmp_path):
    p = tmp_path / "c.nc"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(NotNetCDF3):
        RawNC3(str(p))


def test_write_raw_assembles_the_output_file(tmp_path):
    """step_03's raw writer: input file bytes + the updated fields from big-endian host buffers."""
    import torch
    from pgw4era5_b200 import settings, step_03_apply_to_era as S3
    rng = np.random.default_rng(9)
    nl, ny, nx, ns = 3, 4, 5, 2
    ds = ncio.Dataset()
    ds["time"] = ncio.Variable(("time",), np.array([6.0]), {"units": "hours since 2006-08-01 00:00:00"})
    ds["lat"] = ncio.Variable(("lat",), np.arange(ny, dtype=float)); ds["lon"] = ncio.Variable(("lon",), np.arange(nx, dtype=float))
    ds["level"] = ncio.Variable(("level",), np.arange(1.0, nl + 1)); ds["soil1"] = ncio.Variable(("soil1",), np.arange(1.0, ns + 1))
    ds["ak"] = ncio.Variable(("level",), rng.normal(size=nl))
    shapes = dict(PS=(1, ny, nx), FIS=(1, ny, nx), FR_LAND=(1, ny, nx), FR_SEA_ICE=(1, ny, nx), T_SKIN=(1, ny, nx),
                  T_SO=(1, ns, ny, nx), T=(1, nl, ny, nx), QV=(1, nl, ny, nx), U=(1, nl, ny, nx), V=(1, nl, ny, nx))
    dims = {3: ("time", "lat", "lon"), 4: ("time", "level", "lat", "lon")}
    for k, shp in shapes.items():
        d = ("time", "soil1", "lat", "lon") if k == "T_SO" else dims[len(shp)]
        ds[k] = ncio.Variable(d, rng.normal(size=shp).astype(np.float32), {"long_name": k})
    inp, out = str(tmp_path / "in.nc"), str(tmp_path / "out.nc")
    ds.to_netcdf(inp)
    names = S3._ERA_NAMES(settings.var_name_map)
    host_in = {k: torch.empty(shp, dtype=torch.float32) for k, shp in shapes.items()}
    raw = S3._raw_layout(inp, names, host_in)
    assert raw is not None
    with open(inp, "rb", buffering=0) as f:
        for k in shapes:
            raw.read_into(f, names[k], host_in[k].numpy())
    for k in shapes:          # the raw bytes are the big-endian values
        np.testing.assert_array_equal(host_in[k].numpy().view(">f4").astype(np.float32), ds[k].data)
    new = {k: rng.normal(size=shapes[k]).astype(np.float32) for k in S3._WRITTEN}
    host_out = {k: torch.from_numpy(v.astype(">f4").view(np.float32)) for k, v in new.items()}
    S3._write_raw(raw, names, inp, out, host_out)
    assert os.path.getsize(out) == os.path.getsize(inp)
    back = ncio.open_dataset(out)
    for k in shapes:
        np.testing.assert_array_equal(back[k].data, new[k] if k in new else ds[k].data)
        assert back[k].attrs["long_name"] == k
    np.testing.assert_array_equal(back["ak"].data, ds["ak"].data)
    # files the raw path must leave to the decoding path: float64 fields, RELHUM present
    ds2 = ds.copy()
    ds2["PS"] = ncio.Variable(("time", "lat", "lon"), ds["PS"].data.astype(np.float64))
    ds2.to_netcdf(str(tmp_path / "f64.nc"))
    assert S3._raw_layout(str(tmp_path / "f64.nc"), names, host_in) is None
    ds3 = ds.copy()
    ds3["RELHUM"] = ncio.Variable(("time", "level", "lat", "lon"), ds["T"].data)
    ds3.to_netcdf(str(tmp_path / "rh.nc"))
    assert S3._raw_layout(str(tmp_path / "rh.nc"), names, host_in) is None
    # two records in one file, a field of another size, a file that is not NetCDF-3, and the switch
    ds4 = ncio.Dataset()
    for k, v in ds.variables.items():
        ds4[k] = ncio.Variable(v.dims, np.concatenate([v.data, v.data]), v.attrs) if v.dims[:1] == ("time",) else v
    ds4.to_netcdf(str(tmp_path / "two.nc"))
    assert RawNC3(str(tmp_path / "two.nc")).numrecs == 2
    assert S3._raw_layout(str(tmp_path / "two.nc"), names, host_in) is None
    small = dict(host_in, T=torch.empty((1, nl, ny, nx - 1), dtype=torch.float32))
    assert S3._raw_layout(inp, names, small) is None
    (tmp_path / "h5.nc").write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    assert S3._raw_layout(str(tmp_path / "h5.nc"), names, host_in) is None
    os.environ["PGW_RAW_IO"] = "0"
    try:
        assert S3._raw_layout(inp, names, host_in) is None
    finally:
        del os.environ["PGW_RAW_IO"]
    assert S3._raw_layout(inp, names, host_in) is not None


def test_open_dataset_matches_scipy_reader(tmp_path):
    """ncio.open_dataset reads NetCDF-3 through nc3raw (no 2 GiB limit); same variables, values, dims and
    attributes as scipy's reader on a file with record and fixed variables of several types."""
    from scipy.io import netcdf_file
    path = str(tmp_path / "mix.nc")
    rng = np.random.default_rng(3)
    with netcdf_file(path, "w", version=2) as f:
        f.title, f.version_number = "mixed", np.float64(1.5)
        f.createDimension("time", None); f.createDimension("lat", 3); f.createDimension("n", 5)
        t = f.createVariable("time", "d", ("time",)); t[:] = [0.5, 1.5, 2.5]
        t.units, t.calendar = "days since 2000-01-01 00:00:00", "standard"
        a = f.createVariable("a", "f", ("time", "lat")); a[:] = rng.normal(size=(3, 3)).astype(np.float32)
        a.scale = np.array([1.0, 2.0]); a.missing_value = np.float32(-999.0)
        b = f.createVariable("b", "i", ("n",)); b[:] = np.arange(5, dtype=np.int32)
        c = f.createVariable("c", "h", ("time",)); c[:] = np.array([7, 8, 9], dtype=np.int16)
    got = ncio.open_dataset(path)
    with netcdf_file(path, "r", mmap=False) as f:
        assert set(got.keys()) == set(f.variables)
        for name, v in f.variables.items():
            want = np.array(v.data)
            np.testing.assert_array_equal(got[name].data, want.astype(want.dtype.newbyteorder("=")), err_msg=name)
            assert got[name].dims == v.dimensions and got[name].data.dtype.byteorder in "=|<"
    assert got["time"].attrs["units"] == "days since 2000-01-01 00:00:00"
    np.testing.assert_array_equal(got["a"].attrs["scale"], [1.0, 2.0])
    assert got["a"].attrs["missing_value"] == -999.0 and got.attrs["title"] == "mixed" and got.attrs["version_number"] == 1.5


def test_open_dataset_record_beyond_2gib(tmp_path):
    """One global 0.25 degree timestep is a 2.3 GB record: scipy's reader gives up there, ours must not."""
    import shutil
    if shutil.disk_usage(str(tmp_path)).free < 6e9:
        pytest.skip("not enough scratch space")
    ds = ncio.Dataset()
    ds["time"] = ncio.Variable(("time",), np.array([6.0]))
    nl, ny, nx = 137, 721, 1440
    x = np.zeros((1, nl, ny, nx), np.float32)
    for k, name in enumerate(("T", "QV", "U", "V")):
        x[0, k, 5, 7] = k + 1.0
        ds[name] = ncio.Variable(("time", "level", "lat", "lon"), x.copy())
        x[0, k, 5, 7] = 0.0
    ds["PS"] = ncio.Variable(("time", "lat", "lon"), np.full((1, ny, nx), 1e5, np.float32))
    path = str(tmp_path / "global.nc")
    ds.to_netcdf(path)
    assert os.path.getsize(path) > 2 ** 31
    raw = RawNC3(path)
    assert raw.recsize > 2 ** 31 and raw.numrecs == 1
    back = ncio.open_dataset(path)
    for k, name in enumerate(("T", "QV", "U", "V")):
        assert back[name].data.shape == (1, nl, ny, nx) and back[name].data[0, k, 5, 7] == k + 1.0
        assert float(back[name].data.sum()) == k + 1.0
    assert back["PS"].data[0, 700, 1400] == np.float32(1e5)
