#!/usr/bin/env python
"""
File -> file throughput of the step_03 drop-in: N synthetic ERA5 NetCDF files
(BASELINE configs[0] size, 201 x 281 x 137) through `python -m pgw4era5_b200.step_03_apply_to_era`,
once file by file (read, H2D, pass, D2H, write in sequence) and once with the three-stage pipeline
(reader thread / HostPipeline / writer thread).  Prints one JSON line.

    python tests/bench_files.py [--files 12] [--dir /tmp/pgw_files]
"""
import argparse
import json
import os
import shutil
import sys
import time
from datetime import datetime, timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=12)
    ap.add_argument("--dir", default="/tmp/pgw_files")
    ap.add_argument("--ny", type=int, default=201)
    ap.add_argument("--nx", type=int, default=281)
    ap.add_argument("--raw-only", action="store_true", help="skip the decoding pipeline and the file-by-file run")
    a = ap.parse_args()
    import torch
    from pgw4era5_b200 import settings, synthetic as S, step_03_apply_to_era as S3
    from test_cli_gpu import _write_deltas, _write_era
    settings.i_debug = -1
    shutil.rmtree(a.dir, ignore_errors=True)
    inp, dd = os.path.join(a.dir, "in"), os.path.join(a.dir, "deltas")
    for p in (inp, dd):
        os.makedirs(p)
    import numpy as np
    lat, lon = np.linspace(30.0, 80.0, a.ny), -20.0 + 0.25 * np.arange(a.nx)
    era0 = S.make_era5(a.ny, a.nx, 1, lat=lat, lon=lon)
    _write_deltas(dd, S.make_deltas(era0, 1), lat, lon)
    t0 = datetime(2006, 8, 1, 0)
    whens = [t0 + timedelta(hours=6 * i) for i in range(a.files)]
    for i, w in enumerate(whens):
        _write_era(os.path.join(inp, settings.era5_file_name_base.format(w)),
                   S.make_era5(a.ny, a.nx, 100 + i, lat=lat, lon=lon, orog_seed=1), w)
    fbytes = os.path.getsize(os.path.join(inp, settings.era5_file_name_base.format(whens[0])))
    last = whens[-1].strftime("%Y%m%d%H")
    argv = lambda out: ["-i", inp, "-o", os.path.join(a.dir, out), "-d", dd, "-f", "2006080100", "-l", last,
                        "-H", "6", "-t"]
    S3.main(argv("warm"))                      # loads the climatology, compiles nothing, warms the page cache
    torch.cuda.synchronize()
    S3.IO_STATS.update(raw=0, decoded=0)
    t = time.perf_counter()
    S3.main(argv("pipe"))
    t_pipe = time.perf_counter() - t
    raw_files = S3.IO_STATS["raw"]
    if a.raw_only:
        print(json.dumps({"workload": "%d ERA5 files %dx%dx137 (%.0f MB each), NetCDF-3 in -> NetCDF-3 out" %
                                      (a.files, a.ny, a.nx, fbytes / 1e6),
                          "pipelined_files_per_s": a.files / t_pipe, "raw_io_files": raw_files,
                          "pipelined_MBps_in_plus_out": 2 * fbytes * a.files / t_pipe / 1e6}))
        shutil.rmtree(a.dir, ignore_errors=True)
        return
    os.environ["PGW_RAW_IO"] = "0"             # the same pipeline with the NetCDF codec on the host (scipy)
    t = time.perf_counter()
    S3.main(argv("pipe_dec"))
    t_dec = time.perf_counter() - t
    del os.environ["PGW_RAW_IO"]
    t = time.perf_counter()
    for w in whens:
        name = settings.era5_file_name_base.format(w)
        os.makedirs(os.path.join(a.dir, "seq"), exist_ok=True)
        S3.pgw_for_era5(os.path.join(inp, name), os.path.join(a.dir, "seq", name), dd, w, True)
    t_seq = time.perf_counter() - t
    print(json.dumps({"workload": "%d ERA5 files %dx%dx137 (%.0f MB each), NetCDF-3 in -> NetCDF-3 out" %
                                  (a.files, a.ny, a.nx, fbytes / 1e6),
                      "pipelined_files_per_s": a.files / t_pipe, "raw_io_files": raw_files,
                      "pipelined_decoding_files_per_s": a.files / t_dec,
                      "file_by_file_files_per_s": a.files / t_seq,
                      "pipelined_MBps_in_plus_out": 2 * fbytes * a.files / t_pipe / 1e6,
                      "speedup": t_seq / t_pipe}))
    shutil.rmtree(a.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
