#!/usr/bin/env python
"""
step_02 on BASELINE configs[3]: annual-cycle smoothing and bilinear regridding of one daily 3-D
delta variable (365 days x 19 plevs) from a 1 degree GCM grid (180 x 360) to the ERA5 0.25 degree
grid (721 x 1440).  Device-resident in/out, CUDA events, achieved GB/s on the algorithmic bytes
(smoothing 2 x field, regridding field + 16 x field) against the measured HBM peak.

    python tests/bench_step02.py [--days 365] [--reps 5]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pgw4era5_b200 import _native as N          # noqa: E402
from pgw4era5_b200 import functions as F        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--plevs", type=int, default=19)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    nt, K, ny_s, nx_s = a.days, a.plevs, 180, 360
    lat_s = np.linspace(-89.5, 89.5, ny_s)
    lon_s = 0.5 + np.arange(nx_s)
    lat_t = np.linspace(-90.0, 90.0, 721)
    lon_t = 0.25 * np.arange(1440)
    g = torch.Generator(device=dev).manual_seed(4)
    src = torch.randn((nt, K, ny_s, nx_s), device=dev, dtype=torch.float32, generator=g)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    # ---- smoothing: [nt, npoint] series, one read and one write of the field
    npoint = K * ny_s * nx_s
    smooth = torch.empty_like(src)
    ms_s = timed(lambda: N.check(N.lib.pgw_smooth_harmonic_f32(p(src), p(smooth), nt, npoint, st), "smooth"))
    bytes_s = 2 * src.numel() * 4

    # ---- regridding: tables on the host (regrid_tables), gather kernel on the device
    tb = F.regrid_tables(lat_s, lon_s, lat_t, lon_t)
    ti = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev, dtype=torch.int32)
    tf = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev, dtype=torch.float64)
    j0, j1, i0, i1, wy, wx = ti(tb["j0"]), ti(tb["j1"]), ti(tb["i0"]), ti(tb["i1"]), tf(tb["wy"]), tf(tb["wx"])
    nfield = nt * K
    pm = torch.empty((nfield, 2), device=dev, dtype=torch.float32)
    dst = torch.empty((nt, K, 721, 1440), device=dev, dtype=torch.float32)

    def regrid():
        N.check(N.lib.pgw_zonal_mean_f32(p(smooth), p(pm), nfield, ny_s, nx_s, st), "zonal_mean")
        N.check(N.lib.pgw_regrid_bilinear_f32(p(smooth), p(dst), p(pm), nfield, ny_s, nx_s, 721, 1440,
                                              p(j0), p(j1), p(wy), p(i0), p(i1), p(wx), st), "regrid")
    ms_r = timed(regrid)
    bytes_r = (smooth.numel() + dst.numel()) * 4

    # ---- spot check against the float64 oracle on one field / a few series
    from oracle import pgw_oracle as O
    ref = O.regrid_lat_lon(smooth[3, 5].double().cpu().numpy()[None], lat_s, lon_s, lat_t, lon_t)[0]
    err_r = float(np.abs(dst[3, 5].cpu().numpy() - ref).max())
    cols = src[:, 2, 77, :4].double().cpu().numpy()
    ref_s = np.stack([O.harmonic_ac_analysis(cols[:, i]) for i in range(4)], axis=1)
    err_s = float(np.abs(smooth[:, 2, 77, :4].cpu().numpy() - ref_s).max())

    # ---- ocean variables: 12 monthly tos fields from a 1 degree curvilinear ocean grid (NaN over land)
    # to the ERA5 grid, kernel radius / sharpness of settings.py; whole call (coordinate mapping, sort, kernel)
    from pgw4era5_b200 import settings
    glat, glon = np.meshgrid(np.linspace(-78.0, 89.5, 170), 0.5 + np.arange(360), indexing="ij")
    glat = glat + 0.4 * np.sin(np.radians(glon) * 2)
    rng = np.random.default_rng(12)
    tos = (2.0 + np.cos(np.radians(glat))[None] + 0.2 * rng.normal(size=(12, 170, 360))).astype(np.float32)
    tos[:, rng.uniform(size=(170, 360)) < 0.3] = np.nan
    land_fr = (rng.uniform(size=(721, 1440)) < 0.3).astype(np.float32)
    targs = (land_fr, lat_t, lon_t, torch.as_tensor(tos, device=dev), glat, glon,
             settings.nan_interp_kernel_radius, settings.nan_interp_sharpness)
    ms_o = timed(lambda: F.nan_ignoring_interp_arrays(*targs))
    out_o = F.nan_ignoring_interp_arrays(*targs)
    sub = slice(300, 306), slice(700, 708)
    ref_o = O.nan_ignoring_interp(land_fr[sub], lat_t[sub[0]], lon_t[sub[1]], tos[4], glat, glon,
                                  settings.nan_interp_kernel_radius, settings.nan_interp_sharpness)
    err_o = float(np.nanmax(np.abs(out_o[4][sub].cpu().numpy() - ref_o)))

    print(json.dumps({
        "ocean_regridding": {"workload": "12 monthly fields, 170 x 360 curvilinear ocean grid -> 721 x 1440, radius %g m, "
                                         "sharpness %g" % (settings.nan_interp_kernel_radius, settings.nan_interp_sharpness),
                             "ms": ms_o, "max_abs_err_vs_oracle": err_o},
        "workload": "step_02, one 3-D daily delta variable (%d x %d x 180 x 360 -> 721 x 1440), BASELINE configs[3]" % (nt, K),
        "smoothing": {"ms": ms_s, "algorithmic_bytes": bytes_s, "achieved_gbs": bytes_s / ms_s / 1e6,
                      "frac_of_peak": bytes_s / ms_s / 1e6 / peak, "max_abs_err_vs_oracle": err_s},
        "regridding": {"ms": ms_r, "algorithmic_bytes": bytes_r, "achieved_gbs": bytes_r / ms_r / 1e6,
                       "frac_of_peak": bytes_r / ms_r / 1e6 / peak, "max_abs_err_vs_oracle": err_r},
        "peak_gbs": peak}))


if __name__ == "__main__":
    main()
