"""
Full-size parity of the fused CUDA pass against the fp64 oracle: every one of the 721 x 1440 = 1 038 240 columns
of a global timestep (BASELINE configs[1], [2], [4]).

The reference's stopping rule is field-global (step_03_apply_to_era.py:189,308), so the oracle cannot simply be
run band by band with its own stopping rule.  Instead every band runs ``n_iter_fixed = N`` iterations, N being
the count the CUDA path reported; the per-iteration maximum errors of the bands are merged (max), and the merged
vector must satisfy E_N <= thresh < E_(N-1): then N is exactly what the oracle run on the whole grid would have
returned, and the oracle's ps / hus after iteration N are its results.

Bands are worked on by forked workers (the arrays are inherited copy-on-write, nothing is pickled) that touch
numpy and the oracle's C loops only.  Used by tests/test_global_parity_gpu.py; run as a script it writes the
report that profiles/r2_parity_global.json holds:

    python tests/global_parity.py --out gpurun_out/r2_parity_global.json
"""
import json
import multiprocessing as mp
import os
import sys
import time
from datetime import datetime

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

FIELDS_3D = ("T", "QV", "U", "V")
FIELDS_2D = ("PS", "T_SKIN", "FR_SEA_ICE")
TOL = dict(T=1e-4, PS=1e-2, QV=1e-7, U=1e-4, V=1e-4, T_SKIN=1e-4, T_SO=1e-4, FR_SEA_ICE=1e-6)

_G = {}          # what the forked workers inherit


def _band_worker(rows):
    j0, j1 = rows
    from oracle import pgw_oracle as O
    era, deltas, gpu, kw = _G["era"], _G["deltas"], _G["gpu"], _G["kw"]
    e = {}
    for k, v in era.items():
        e[k] = v[..., j0:j1, :] if isinstance(v, np.ndarray) and v.ndim >= 3 else v
    d = {k: dict(v, data=v["data"][..., j0:j1, :]) for k, v in deltas.items()}
    out = O.pgw_for_era5(e, d, _G["when"], n_iter_fixed=_G["n_iter"], ignore_top_pressure_error=True, **kw)
    ref = dict(PS=out["ps_traj"][-1], QV=out["hus_traj"][-1], T=out["T"], U=out["U"], V=out["V"],
               T_SKIN=out["T_SKIN"], T_SO=out["T_SO"], FR_SEA_ICE=out["FR_SEA_ICE"])
    res = dict(errs=[float(x) for x in out["phi_max_errors"]], maxerr={}, over_tol={}, nan_mismatch=0)
    for name, r in ref.items():
        g = gpu[name][..., j0:j1, :].astype(np.float64).reshape(np.shape(r))
        r = np.asarray(r, dtype=np.float64)
        res["nan_mismatch"] += int(np.sum(np.isnan(g) != np.isnan(r)))
        with np.errstate(invalid="ignore"):
            dd = np.abs(g - r)
        res["maxerr"][name] = float(np.nanmax(dd)) if dd.size else 0.0
        res["over_tol"][name] = int(np.sum(dd > TOL[name]))
    return res


def reduce_deltas(deltas_np, when):
    """Only the two stamps that bracket ``when`` of every delta (the whole series of `ts`, whose annual mean
    enters the soil temperature, step_03_apply_to_era.py:134-136): same brackets and weights, a sixth of the
    bytes the workers have to slice."""
    from oracle import pgw_oracle as O
    out = {}
    for name, v in deltas_np.items():
        if name == "ts":
            out[name] = v
            continue
        t = np.asarray(v["time"]).astype("datetime64[ns]")
        keep, ib, ia, _, _ = O.delta_time_bracket(t, when)
        idx = sorted({keep[ib], keep[ia]})
        # a periodic wrap (ib = -1 or ia = 0 across the year boundary) keeps first and last stamp: same bracket
        out[name] = dict(v, time=t[idx], data=np.ascontiguousarray(np.asarray(v["data"])[idx]))
        k2, ib2, ia2, xh2, xn2 = O.delta_time_bracket(out[name]["time"], when)
        _, _, _, xh, xn = O.delta_time_bracket(t, when)
        assert (xh2, xn2) == (xh, xn), (name, xh2, xn2, xh, xn)
    return out


def oracle_vs_gpu(era_np, deltas_np, gpu_np, when, n_iter, thresh, band_rows=8, procs=None, **kw):
    """Run the oracle over all rows in bands of ``band_rows`` with ``n_iter`` iterations and compare with the
    CUDA results ``gpu_np`` (numpy, float32).  Returns the report dict."""
    ny = era_np["PS"].shape[-2]
    _G.update(era=era_np, deltas=reduce_deltas(deltas_np, when), gpu=gpu_np, when=when, n_iter=int(n_iter),
              kw=dict(kw, thresh_phi_ref_max_error=thresh))
    bands = [(j, min(j + band_rows, ny)) for j in range(0, ny, band_rows)]
    procs = procs or min(len(bands), os.cpu_count() or 1)
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map_async(_band_worker, bands, chunksize=1).get(timeout=1500)
    else:
        parts = [_band_worker(b) for b in bands]
    wall = time.perf_counter() - t0
    E = np.max(np.array([p["errs"] for p in parts]), axis=0)
    rep = dict(columns=int(ny * era_np["PS"].shape[-1]), bands=len(bands), procs=procs, oracle_wall_s=wall,
               n_iter_gpu=int(n_iter), oracle_max_err_per_iteration=[float(x) for x in E],
               n_iter_oracle=int(np.argmax(E <= thresh)) + 1 if np.any(E <= thresh) else 0,
               nan_mismatch=int(sum(p["nan_mismatch"] for p in parts)), maxerr={}, over_tol={})
    for name in parts[0]["maxerr"]:
        rep["maxerr"][name] = float(max(p["maxerr"][name] for p in parts))
        rep["over_tol"][name] = int(sum(p["over_tol"][name] for p in parts))
    _G.clear()
    return rep


def run_case(name, ny=721, nx=1440, seed=2, plev=None, thresh=0.15, when=datetime(2006, 8, 2, 6), band_rows=8,
             orog_seed=2, delta_seed=2, ref_dtypes=False):
    """One global timestep through the engine (TMA flavour) and the banded oracle."""
    import torch
    from pgw4era5_b200 import settings, synthetic as S
    from pgw4era5_b200.engine import DeltaSet, PGWEngine
    plev = S.PLEV19 if plev is None else plev
    era = S.make_era5(ny, nx, seed, device="cuda", orog_seed=orog_seed)
    deltas = S.make_deltas(era if seed == delta_seed else S.make_era5(ny, nx, delta_seed, device="cuda",
                                                                      orog_seed=orog_seed),
                           delta_seed, plev=plev, device="cuda")
    old = settings.thresh_phi_ref_max_error, getattr(settings, "i_reference_dtypes", 0)
    settings.thresh_phi_ref_max_error, settings.i_reference_dtypes = thresh, int(ref_dtypes)
    try:
        eng = PGWEngine(era["ak"], era["bk"], DeltaSet(deltas, device="cuda"), soil1=era["soil1"])
        p = eng.submit(era, when, ignore_top_pressure_error=True)
        import ctypes
        from pgw4era5_b200 import _native
        tma = _native.lib.pgw_timestep_uses_tma(ctypes.byref(p.args))
        res = p.result()
    finally:
        settings.thresh_phi_ref_max_error, settings.i_reference_dtypes = old
    gpu_np = {k: res[k].detach().cpu().numpy() for k in FIELDS_3D + FIELDS_2D + ("T_SO",)}
    era_np, deltas_np = S.to_numpy(era), S.to_numpy(deltas)
    del era, deltas, eng
    torch.cuda.empty_cache()
    rep = oracle_vs_gpu(era_np, deltas_np, gpu_np, when, res["n_iter"], thresh, band_rows=band_rows,
                        emulate_file_dtypes=ref_dtypes)
    rep.update(case=name, grid=[ny, nx], nplev=len(plev), thresh=thresh, when=when.isoformat(), seed=seed,
               tma_flavour=bool(tma), reference_dtypes=bool(ref_dtypes),
               phi_max_errors_gpu=[float(x) for x in res["phi_max_errors"]],
               poly_fallback_warp_iterations=int(res.get("poly_fallback", -1)),
               warps=int((ny * nx + 31) // 32))
    return rep


CASES = {
    # BASELINE configs[1]: global single timestep, plev19, default threshold
    "configs1_plev19": dict(seed=2, thresh=0.15),
    # BASELINE configs[4]: 37 pressure levels, tight threshold (worst-case iteration count)
    "configs4_plev37_tight": dict(seed=5, delta_seed=5, orog_seed=5, plev="PLEV37", thresh=1e-3),
    # BASELINE configs[2]: three dates of the 124-step month (different weather, same climatology)
    "configs2_step000": dict(seed=1000, when=datetime(2006, 8, 1, 0)),
    "configs2_step061": dict(seed=1061, when=datetime(2006, 8, 16, 6)),
    "configs2_step123": dict(seed=1123, when=datetime(2006, 8, 31, 18)),
}


def run_named(name, **over):
    from pgw4era5_b200 import synthetic as S
    kw = dict(CASES[name])
    if kw.get("plev") == "PLEV37":
        kw["plev"] = S.PLEV37
    kw.update(over)
    return run_case(name, **kw)


def check(rep):
    """The bar: identical iteration count, every field within the north_star tolerance on every column."""
    assert rep["n_iter_oracle"] == rep["n_iter_gpu"], rep
    assert rep["nan_mismatch"] == 0, rep
    for name, n_bad in rep["over_tol"].items():
        assert n_bad == 0, (name, rep["maxerr"], rep["over_tol"])
    return rep


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_parity_global.json"))
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--ny", type=int, default=721)
    ap.add_argument("--nx", type=int, default=1440)
    a = ap.parse_args()
    reports = []
    for c in a.cases.split(","):
        t0 = time.time()
        r = run_named(c, ny=a.ny, nx=a.nx)
        r["wall_s"] = time.time() - t0
        ok = True
        try:
            check(r)
        except AssertionError:
            ok = False
        r["pass"] = ok
        print(json.dumps(r), flush=True)
        reports.append(r)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(dict(host_cores=os.cpu_count(), tolerances=TOL, cases=reports), f, indent=1)
