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


def breakdown(a, inp, dd, whens, fbytes):
    """The three stages of the raw file pipeline, one after the other on the first file (page cache warm), so that
    the pipelined files/s can be read against them: (1) file -> pinned host buffers (pread of the big-endian
    fields, 4 threads), (2) H2D + byte swap + fused pass + byte swap + D2H (HostPipeline, = bench.py's e2e step),
    (3) pinned buffers + untouched bytes of the input file -> output file (pwrite / copy_file_range, 4 threads)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from pgw4era5_b200 import ncio, settings, step_03_apply_to_era as S3
    from pgw4era5_b200.hostpipe import HostPipeline, IN_FIELDS
    path = os.path.join(inp, settings.era5_file_name_base.format(whens[0]))
    first = ncio.open_dataset(path, decode_cf=False)
    eng = S3.get_engine(dd, first)
    names = S3._ERA_NAMES(settings.var_name_map)
    pipe = HostPipeline(eng, a.ny, a.nx)
    h_in, h_out = pipe.alloc_host_inputs(), pipe.alloc_host_outputs()
    raw = S3._raw_layout(path, names, h_in)
    pool = ThreadPoolExecutor(8)
    out = {}

    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
        return min(ts)

    def read():
        with open(path, "rb", buffering=0) as f:
            for fut in [pool.submit(raw.read_into, f, names[k], h_in[k].numpy()) for k in IN_FIELDS]:
                fut.result()
    out["read_file_to_pinned_s"] = best(read)

    def gpu():
        pipe.run(h_in, whens[0], h_out, raw=True, ignore_top_pressure_error=True)
        pipe.drain()
        torch.cuda.synchronize()
    gpu()
    out["h2d_kernel_d2h_s"] = best(gpu)
    os.makedirs(os.path.join(a.dir, "bd"), exist_ok=True)
    dst = os.path.join(a.dir, "bd", "out.nc")
    out["write_pinned_to_file_s"] = best(lambda: S3._write_raw(raw, names, path, dst, h_out, pool))
    out["file_bytes"] = fbytes
    out["sum_of_stages_s"] = out["read_file_to_pinned_s"] + out["h2d_kernel_d2h_s"] + out["write_pinned_to_file_s"]
    out["slowest_stage_files_per_s"] = 1.0 / max(out["read_file_to_pinned_s"], out["h2d_kernel_d2h_s"],
                                                 out["write_pinned_to_file_s"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=12)
    ap.add_argument("--dir", default="/tmp/pgw_files")
    ap.add_argument("--ny", type=int, default=201)
    ap.add_argument("--nx", type=int, default=281)
    ap.add_argument("--raw-only", action="store_true", help="skip the decoding pipeline and the file-by-file run")
    ap.add_argument("--breakdown", action="store_true",
                    help="also time the three stages of the raw pipeline one by one on the first file")
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
    stages = None
    if a.breakdown:
        stages = breakdown(a, inp, dd, whens, fbytes)
    if a.raw_only:
        print(json.dumps({"workload": "%d ERA5 files %dx%dx137 (%.0f MB each), NetCDF-3 in -> NetCDF-3 out" %
                                      (a.files, a.ny, a.nx, fbytes / 1e6),
                          "pipelined_files_per_s": a.files / t_pipe, "raw_io_files": raw_files,
                          "pipelined_MBps_in_plus_out": 2 * fbytes * a.files / t_pipe / 1e6, "stages": stages}))
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
