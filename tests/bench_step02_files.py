#!/usr/bin/env python
"""
step_02 end to end, file -> file, on a slice of BASELINE configs[3]: one daily 3-D delta variable on a 1 degree GCM
grid (D days x 19 plevs x 180 x 360) through the drop-in command line, `smoothing` then `regridding` to the ERA5
0.25 degree grid (721 x 1440), NetCDF-3 in and out.  Prints one JSON line with the wall time of each stage and the
time the same fields take in the two CUDA kernels alone: the answer to whether one variable needs to be split over
several GPUs (third row) or whether dealing whole files out to GPUs (`-p N`) is enough.

    python tests/bench_step02_files.py [--days 48] [--dir /tmp/pgw_step02]

(48 days keep the regridded variable below the 4 GiB a NetCDF-3 fixed-size variable may have; all stages are linear
in the number of days.)
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=48)
    ap.add_argument("--plevs", type=int, default=19)
    ap.add_argument("--dir", default="/tmp/pgw_step02")
    a = ap.parse_args()
    import torch
    from pgw4era5_b200 import functions as F, ncio, settings, step_02_preproc_deltas as S2, synthetic as S
    shutil.rmtree(a.dir, ignore_errors=True)
    d_in, d_sm, d_out = (os.path.join(a.dir, x) for x in ("gcm", "smooth", "era"))
    for p in (d_in, d_sm, d_out):
        os.makedirs(p)
    nt, K = a.days, a.plevs
    lat_s, lon_s = np.linspace(-89.5, 89.5, 180), 0.5 + np.arange(360)
    lat_t, lon_t = np.linspace(-90.0, 90.0, 721), 0.25 * np.arange(1440)
    rng = np.random.default_rng(4)
    data = rng.normal(size=(nt, K, 180, 360)).astype(np.float32)
    times = S.monthly_stamps(ntime=nt)
    for period in ("HIST", "SCEN-HIST"):
        ds = ncio.Dataset()
        ds["time"] = ncio.encode_time(times, "days since 1850-01-01 00:00:00")
        ds["plev"] = ncio.Variable(("plev",), S.PLEV19[:K].copy())
        ds["lat"] = ncio.Variable(("lat",), lat_s)
        ds["lon"] = ncio.Variable(("lon",), lon_s)
        ds["ta"] = ncio.Variable(("time", "plev", "lat", "lon"), data)
        ds.to_netcdf(os.path.join(d_in, settings.file_name_bases[period].format("ta")))
    era = ncio.Dataset()
    era["lat"] = ncio.Variable(("lat",), lat_t)
    era["lon"] = ncio.Variable(("lon",), lon_t)
    era_path = os.path.join(a.dir, "era5_grid.nc")
    era.to_netcdf(era_path)
    in_bytes = os.path.getsize(os.path.join(d_in, settings.file_name_bases["HIST"].format("ta")))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S2.main(["smoothing", "-i", d_in, "-o", d_sm, "-v", "ta"])
    torch.cuda.synchronize()
    t_smooth = time.perf_counter() - t0
    t0 = time.perf_counter()
    S2.main(["regridding", "-i", d_sm, "-o", d_out, "-e", era_path, "-v", "ta"])
    torch.cuda.synchronize()
    t_regrid = time.perf_counter() - t0
    out_bytes = os.path.getsize(os.path.join(d_out, settings.file_name_bases["HIST"].format("ta")))
    # the kernels alone on the same fields (device resident, CUDA events)
    dev = torch.device("cuda", torch.cuda.current_device())
    d = torch.as_tensor(data, device=dev)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1), r
    ms_s, sm = timed(lambda: F.smooth_annual_cycle(d))
    ms_r, _ = timed(lambda: F.regrid_arrays(sm, lat_s, lon_s, lat_t, lon_t))
    print(json.dumps({
        "workload": "step_02 file -> file, 2 files (HIST, SCEN-HIST) of one daily 3-D variable, %d days x %d plevs x 180 x 360 "
                    "-> 721 x 1440 (BASELINE configs[3] has 365 days: every stage is linear in the days)" % (nt, K),
        "input_bytes_per_file": in_bytes, "output_bytes_per_file": out_bytes,
        "smoothing_cli_s_for_2_files": t_smooth, "regridding_cli_s_for_2_files": t_regrid,
        "cuda_kernels_ms_per_file": {"smoothing": ms_s, "regridding_incl_zonal_mean": ms_r},
        "kernel_share_of_regridding_wall": 2 * ms_r / 1000.0 / t_regrid,
        "regridding_output_gb_per_s": 2 * out_bytes / t_regrid / 1e9,
        "extrapolated_to_365_days_s_per_variable": {"smoothing": t_smooth * 365.0 / nt, "regridding": t_regrid * 365.0 / nt},
    }), flush=True)
    shutil.rmtree(a.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
