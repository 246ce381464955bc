"""Shared builders for the parity tests: seeded synthetic cases + oracle runs."""
from datetime import datetime

import numpy as np

from pgw4era5_b200 import synthetic as S

ERA_DATE = datetime(2006, 8, 2, 6)

# north_star tolerances (BASELINE.json): ta 1e-4 K, ps 1e-2 Pa, hus 1e-7 kg/kg
TOL = dict(T=1e-4, PS=1e-2, delta_ps=1e-2, QV=1e-7, U=1e-4, V=1e-4, T_SKIN=1e-4, T_SO=1e-4,
           FR_SEA_ICE=1e-6)


def make_case(ny, nx, seed, plev=S.PLEV19, region="EU", device="cpu"):
    if region == "EU":
        lat, lon = np.linspace(30.0, 80.0, ny), np.linspace(-20.0, 50.0, nx)
    else:
        lat, lon = np.linspace(-90.0, 90.0, ny), np.arange(nx) * (360.0 / nx)
    era = S.make_era5(ny, nx, seed, device=device, lat=lat, lon=lon)
    deltas = S.make_deltas(era, seed, plev=plev, device=device)
    return era, deltas


def run_oracle(era, deltas, when=ERA_DATE, **kw):
    from oracle import pgw_oracle as O
    kw.setdefault("ignore_top_pressure_error", True)
    return O.pgw_for_era5(S.to_numpy(era), S.to_numpy(deltas), when, **kw)


def compare(res, ref):
    """max |gpu - oracle| per output field (NaNs must coincide)."""
    out = {}
    for name in ("PS", "T", "QV", "U", "V", "T_SKIN", "T_SO", "FR_SEA_ICE"):
        g = res[name].detach().cpu().numpy().astype(np.float64).reshape(ref[name].shape)
        r = np.asarray(ref[name], dtype=np.float64)
        assert np.array_equal(np.isnan(g), np.isnan(r)), "NaN pattern differs in %s" % name
        out[name] = float(np.nanmax(np.abs(g - r))) if g.size else 0.0
    g = res["delta_ps"].detach().cpu().numpy().astype(np.float64).reshape(ref["PS"].shape)
    out["delta_ps"] = float(np.nanmax(np.abs(g - ref["deltas"]["ps"])))
    return out
