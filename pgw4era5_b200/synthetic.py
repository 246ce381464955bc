"""
Seeded synthetic ERA5 timesteps and GCM climate deltas.

The reference ships no data (``*.nc`` is git-ignored there), and its model-level
coefficients live inside the ERA5 files, so every test and benchmark in this
repository runs on the fields generated here.  Everything is float32 storage,
generated with torch on the requested device (CPU for the parity tests so the
float64 oracle sees exactly the same numbers, CUDA for full-size benchmarks).

Conventions follow the reference's files: ERA5 fields ``(time=1, level, lat,
lon)`` with level 0 at the model top (step_03_apply_to_era.py:64-88), deltas
``(time=12, plev, lat, lon)`` with plev *descending* in pressure as CMIP6/cdo
writes them (functions.py:383-384 flips them).
"""
from datetime import datetime

import numpy as np
import torch

from .constants import CON_G, CON_RD

PLEV19 = np.array([100000, 92500, 85000, 70000, 60000, 50000, 40000, 30000, 25000,
                   20000, 15000, 10000, 7000, 5000, 3000, 2000, 1000, 500, 100],
                  dtype=np.float64)
PLEV37 = np.array([1000, 975, 950, 925, 900, 875, 850, 825, 800, 775, 750, 700, 650,
                   600, 550, 500, 450, 400, 350, 300, 250, 225, 200, 175, 150, 125,
                   100, 70, 50, 30, 20, 10, 7, 5, 3, 2, 1], dtype=np.float64) * 100.0

SOIL_DEPTHS = np.array([0.035, 0.175, 0.64, 1.945], dtype=np.float64)

GRIDS = {
    # name: (lat array, lon array)
    "EU": (np.linspace(30.0, 80.0, 201), -20.0 + 0.25 * np.arange(281)),
    "GL": (np.linspace(-90.0, 90.0, 721), 0.25 * np.arange(1440)),
}


def hybrid_coefficients(nlev=137):
    """
    Half-level hybrid coefficients ak [Pa], bk [1] (nlev+1 values, index 0 = model
    top) shaped like ECMWF L137: top half level at 0 Pa, pure pressure levels
    (bk = 0) down to half level 55 (~43 hPa), about 54 full levels below
    300 hPa for ps = 1013.25 hPa, lowest layer ~2.4 hPa thick.  Piecewise
    linear in (k, ln p); any monotone table is valid because oracle and kernels
    share it.
    """
    ps0 = 101325.0
    s = nlev / 137.0
    anchors_k = np.array([1, 20, 36, 55, 68, 76, 83, 96, 105, 114, 123, 130, 137]) * s
    anchors_p = np.array([2.0, 150.0, 1000.0, 4300.0, 10000.0, 20000.0, 30000.0,
                          50000.0, 70000.0, 85000.0, 95000.0, 99000.0, ps0])
    k = np.arange(nlev + 1, dtype=np.float64)
    p = np.exp(np.interp(k, anchors_k, np.log(anchors_p)))
    p[0] = 0.0
    k55 = int(round(55 * s))
    r = np.clip((p - p[k55]) / (ps0 - p[k55]), 0.0, 1.0)
    bk = r ** 1.5
    bk[nlev] = 1.0
    ak = p - ps0 * bk
    ak[nlev] = 0.0
    ak = np.maximum(ak, 0.0)
    for ps in (45000.0, 110000.0):
        assert np.all(np.diff(ak + ps * bk) > 0), "hybrid table not monotone"
    return ak, bk


def _smooth_noise(gen, ny, nx, coarse, device, lead=()):
    """Smooth N(0,1)-ish field: coarse noise, bicubic-free bilinear upsampling."""
    cy, cx = max(2, ny // coarse + 2), max(2, nx // coarse + 2)
    z = torch.randn(*lead, cy, cx, generator=gen, device=device, dtype=torch.float32)
    z4 = z.reshape(1, -1, cy, cx)
    out = torch.nn.functional.interpolate(z4, size=(ny, nx), mode="bilinear",
                                          align_corners=True)
    return out.reshape(*lead, ny, nx)


def _esat(ta):
    """IFS saturation vapour pressure blend (functions.py:74-105), float64 torch."""
    T0, Ti = 273.16, 250.16
    ew = 611.21 * torch.exp(17.502 * (ta - T0) / (ta - 32.19))
    ei = 611.21 * torch.exp(22.587 * (ta - T0) / (ta + 0.7))
    alpha = torch.clamp((ta - Ti) / (T0 - Ti), 0.0, 1.0) ** 2
    return alpha * ew + (1 - alpha) * ei


def make_era5(ny, nx, seed, device="cpu", nlev=137, lat=None, lon=None, orog_seed=None):
    """
    One synthetic ERA5 timestep on an (ny, nx) grid.  Returns a dict of float32
    torch tensors (plus float64 numpy ``ak``/``bk``/``soil1``/``lat``/``lon``).
    ``orog_seed`` (default: ``seed``) seeds orography, surface pressure and the
    land mask, so several timesteps over the same terrain can be generated.
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed if orog_seed is None else orog_seed))
    if lat is None:
        lat = np.linspace(-90.0, 90.0, ny) if ny > 1 else np.zeros(1)
    if lon is None:
        lon = np.arange(nx) * (360.0 / nx)
    ak, bk = hybrid_coefficients(nlev)
    akm = 0.5 * (ak[1:] + ak[:-1])
    bkm = 0.5 * (bk[1:] + bk[:-1])
    f64 = torch.float64

    # orography: positive part of smooth noise, a few high ranges
    zs = torch.clamp(_smooth_noise(gen, ny, nx, 12, dev), min=0.0) ** 2 * 900.0
    zs = torch.clamp(zs + 40.0 * torch.clamp(_smooth_noise(gen, ny, nx, 4, dev), min=0.0), 0.0, 5500.0)
    fis = (CON_G * zs).to(torch.float32)
    ps = 101325.0 * torch.exp(-CON_G * zs.to(f64) / (CON_RD * 270.0))
    ps = ps * (1.0 + 0.01 * _smooth_noise(gen, ny, nx, 10, dev).to(f64))
    PS = ps.to(torch.float32)
    if orog_seed is not None:
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed))
        PS = (ps * (1.0 + 0.004 * _smooth_noise(gen, ny, nx, 10, dev).to(f64))).to(torch.float32)

    tsfc = 288.0 - 0.0065 * zs + 8.0 * _smooth_noise(gen, ny, nx, 16, dev)
    lat_t = torch.as_tensor(lat, device=dev, dtype=torch.float32).reshape(ny, 1)
    tsfc = tsfc - 25.0 * (torch.abs(lat_t) / 90.0) ** 2

    akm_t = torch.as_tensor(akm, device=dev, dtype=f64).reshape(nlev, 1, 1)
    bkm_t = torch.as_tensor(bkm, device=dev, dtype=f64).reshape(nlev, 1, 1)
    p = akm_t + PS.to(f64).unsqueeze(0) * bkm_t                     # [L,ny,nx]
    sig = (p / PS.to(f64).unsqueeze(0)).to(torch.float32)
    T = torch.clamp(tsfc.unsqueeze(0) * sig ** (CON_RD * 0.0065 / CON_G), min=215.0)
    T = T + 0.5 * torch.randn(nlev, ny, nx, generator=gen, device=dev)
    rh = torch.clamp(70.0 * sig ** 1.5 + 10.0 * torch.randn(nlev, ny, nx, generator=gen, device=dev),
                     1.0, 100.0)
    T = T.to(torch.float32)
    e = rh.to(f64) / 100.0 * _esat(T.to(f64))
    QV = torch.clamp(0.622 * e / (p - 0.378 * e), 2.0e-6, 0.03).to(torch.float32)
    U = 10.0 * torch.randn(nlev, ny, nx, generator=gen, device=dev)
    V = 10.0 * torch.randn(nlev, ny, nx, generator=gen, device=dev)

    land = (zs > 1.0).to(torch.float32)
    fr_land = torch.clamp(land + 0.0, 0.0, 1.0)
    ice = torch.clamp((torch.abs(lat_t) - 65.0) / 15.0, 0.0, 1.0) * torch.ones(ny, nx, device=dev)
    ice = torch.clamp(ice + 0.1 * _smooth_noise(gen, ny, nx, 8, dev), 0.0, 1.0)
    ice = torch.where(fr_land > 0.5, torch.full_like(ice, float("nan")), ice)
    tskin = tsfc + 0.5 * _smooth_noise(gen, ny, nx, 6, dev)
    nsoil = len(SOIL_DEPTHS)
    tso = tskin.unsqueeze(0).repeat(nsoil, 1, 1) - 0.5 * torch.arange(
        nsoil, device=dev, dtype=torch.float32).reshape(nsoil, 1, 1)

    c = lambda a: a.to(torch.float32).contiguous()
    return dict(
        ak=ak, bk=bk, akm=akm, bkm=bkm, soil1=SOIL_DEPTHS.copy(),
        lat=np.asarray(lat, dtype=np.float64), lon=np.asarray(lon, dtype=np.float64),
        PS=c(PS)[None], FIS=c(fis)[None], T=c(T)[None], QV=c(QV)[None],
        U=c(U)[None], V=c(V)[None], FR_LAND=c(fr_land)[None],
        FR_SEA_ICE=c(ice)[None], T_SKIN=c(tskin)[None], T_SO=c(tso)[None],
        zs=c(zs),
    )


def monthly_stamps(year=2000, ntime=12):
    """12 monthly stamps (16th, 12 UTC) or ``ntime`` daily stamps (12 UTC)."""
    if ntime == 12:
        return np.array([np.datetime64(datetime(year, m, 16, 12), "ns") for m in range(1, 13)])
    base = np.datetime64(datetime(2001, 1, 1, 12), "ns")
    return base + np.arange(ntime) * np.timedelta64(24 * 3600 * 10 ** 9, "ns")


def make_deltas(era, seed, plev=PLEV19, device="cpu", ntime=12, zg_noise_m=5.0):
    """
    Climate deltas consistent with ``era``'s grid.  Returns var -> dict(time,
    plev, data) with float32 torch ``data`` [ntime,(K),ny,nx]; 'ps_hist' is the
    HIST surface-pressure climatology (functions.py:330-332).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) * 7919 + 13)
    ny, nx = era["PS"].shape[-2:]
    K = len(plev)
    times = monthly_stamps(ntime=ntime)
    plev_t = torch.as_tensor(plev, device=dev, dtype=torch.float32).reshape(1, K, 1, 1)
    season = torch.cos(2 * np.pi * (torch.arange(ntime, device=dev, dtype=torch.float32) + 0.5) / ntime
                       ).reshape(ntime, 1, 1, 1)
    lnp = torch.log(plev_t)
    prof = (4.0 * (plev_t / 1e5) ** 0.3 + 1.5 * torch.exp(-((lnp - np.log(25000.0)) / 0.5) ** 2)
            - 3.0 * (plev_t < 1e4).to(torch.float32))
    amp = 1.0 + 0.15 * _smooth_noise(gen, ny, nx, 20, dev).reshape(1, 1, ny, nx)
    dta = prof * amp + 0.5 * season + 0.3 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime, K))
    dhur = 1.0 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime, K))
    dua = 0.5 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime, K))
    dva = 0.5 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime, K))
    # index of the highest-pressure level in file order
    ibot = int(np.argmax(plev))
    dtas = dta[:, ibot] + 0.2 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime,))
    dhurs = dhur[:, ibot] + 0.3 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime,))
    PS = era["PS"].to(dev)
    ps_hist = PS * (1.0 + 0.005 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime,)))
    ps_hist = torch.clamp(ps_hist, min=float(np.min(plev)) * 1.5)
    dts = dtas + 0.2 * _smooth_noise(gen, ny, nx, 10, dev, lead=(ntime,))
    land = era["FR_LAND"].to(dev)[0] > 0.5
    dtos = torch.where(land.unsqueeze(0), torch.full_like(dtas, float("nan")), dtas - 0.5)
    ice = era["FR_SEA_ICE"].to(dev)[0]
    dsic = torch.where(torch.nan_to_num(ice, nan=0.0) > 0, torch.full_like(ice, -10.0),
                       torch.zeros_like(ice)).unsqueeze(0).repeat(ntime, 1, 1)
    # zg delta: hypsometric thickness of dta from the bottom plev up + smooth offset
    order = np.argsort(-np.asarray(plev))                 # bottom -> top
    dzg = torch.zeros(ntime, K, ny, nx, device=dev, dtype=torch.float32)
    z0 = zg_noise_m * _smooth_noise(gen, ny, nx, 14, dev, lead=(ntime,))
    acc = z0.clone()
    dzg[:, order[0]] = acc
    for a, b in zip(order[:-1], order[1:]):
        thick = (CON_RD / CON_G) * 0.5 * (dta[:, a] + dta[:, b]) * float(np.log(plev[a] / plev[b]))
        acc = acc + thick
        dzg[:, b] = acc
    d3 = lambda x: dict(time=times, plev=np.asarray(plev, dtype=np.float64),
                        data=x.to(torch.float32).contiguous())
    d2 = lambda x: dict(time=times, plev=None, data=x.to(torch.float32).contiguous())
    return dict(ta=d3(dta), hur=d3(dhur), ua=d3(dua), va=d3(dva), zg=d3(dzg),
                tas=d2(dtas), hurs=d2(dhurs), ps_hist=d2(ps_hist), ts=d2(dts),
                tos=d2(dtos), siconc=d2(dsic))


def to_numpy(tree):
    """Recursively convert torch tensors in a dict tree to numpy (for the oracle)."""
    if isinstance(tree, dict):
        return {k: to_numpy(v) for k, v in tree.items()}
    if isinstance(tree, torch.Tensor):
        return tree.detach().cpu().numpy()
    return tree
