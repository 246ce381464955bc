#!/usr/bin/env python
"""
Benchmark of the PGW4ERA5 per-timestep path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): global 0.25 degree ERA5 timesteps/s.  One "step" is one
full ERA5 timestep (721x1440 columns x 137 levels, plev19 deltas) through the
fused CUDA pass.  N > 1: every rank processes its own timesteps (the
reference's own parallelism: one file = one task), so `value` = N*K / max-over-
ranks time, scaling "weak"; the delta climatology is NCCL-broadcast once before
the timed region.  Inputs cycle through a ring of distinct device-resident
synthetic timesteps (each 2.3 GB, far larger than the 126 MB L2).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from datetime import datetime, timedelta

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "global 0.25deg ERA5 timesteps/s"
GRIDS = {"GL": (721, 1440), "EU": (201, 281)}


def algorithmic_bytes(ncol, nlev=137, nplev=19, nsoil=4):
    """every input read once, every output written once, fp32."""
    n4 = ncol * 4
    reads = 4 * nlev * n4 + 4 * 2 * nplev * n4 + 7 * 2 * n4 + n4 + 5 * n4 + nsoil * n4
    writes = 4 * nlev * n4 + 3 * n4 + nsoil * n4
    return reads + writes


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms 100` process runs across the region; samples are kept by wall-clock time."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.t0, self.t1 = gpu_index, None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def summary(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
            for line in out.splitlines():
                r = [x.strip() for x in line.split(",")]
                if len(r) < 10:
                    continue
                try:
                    ts = datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    continue
                rows.append((ts, r))
        inside = [r for ts, r in rows if self.t0 is not None and self.t0 - 0.05 <= ts <= self.t1 + 0.15]
        use = inside if inside else [r for _, r in rows]
        num = lambda x: float(x) if x.replace(".", "", 1).isdigit() else None
        sm = [num(r[2]) for r in use if num(r[2]) is not None]
        mx = [num(r[3]) for r in use if num(r[3]) is not None]
        reasons = set()
        for r in use:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 r[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(use), "samples_in_region": len(inside)}


# --------------------------------------------------------------------------- CPU arm
def _cpu_sample(args):
    """One bounded sample of the global workload on one host core: ny rows x 1440 columns."""
    seed, ny, nx = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    torch.set_num_threads(1)
    from pgw4era5_b200 import synthetic as S
    from oracle import pgw_oracle as O
    lat = np.linspace(-60.0, 60.0, ny)
    era = S.make_era5(ny, nx, seed, lat=lat, lon=np.arange(nx) * 0.25)
    deltas = S.make_deltas(era, seed)
    e, d = S.to_numpy(era), S.to_numpy(deltas)
    t0 = time.perf_counter()
    out = O.pgw_for_era5(e, d, datetime(2006, 8, 2, 6), ignore_top_pressure_error=True)
    return time.perf_counter() - t0, out["n_iter"]


def cpu_arm(n_tasks, procs, rows=8, nx=1440, pool=None):
    """The oracle port of the reference path on `procs` host cores, one band per task, the
    way parallel.IterMP spreads files over workers (parallel.py:18-32).  Returns
    (timesteps/s, wall seconds, columns processed)."""
    tasks = [(1000 + i, rows, nx) for i in range(n_tasks)]
    t0 = time.perf_counter()
    if pool is not None:
        res = pool.map(_cpu_sample, tasks, chunksize=1)
    else:
        res = [_cpu_sample(t) for t in tasks]
    wall = time.perf_counter() - t0
    # time of the reference path only (the synthetic inputs are generated outside it): the slowest
    # worker when every worker holds one band, the sum on one core
    comp = [r[0] for r in res]
    if pool is None:
        wall = sum(comp)
    elif n_tasks <= procs:
        wall = max(comp)
    cols = n_tasks * rows * nx
    ts = cols / float(GRIDS["GL"][0] * GRIDS["GL"][1])
    return ts / wall, wall, cols


def run_reference(a):
    """--impl reference: the reference path's CPU implementation (oracle port; the verbatim
    reference needs xarray which this image lacks) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import build as ob
    ob.build()
    procs = max(1, os.cpu_count() or 1)
    procs = min(procs, 64)
    steps = max(1, a.steps)
    ctx = mp.get_context("spawn")
    vals, walls = [], []
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_sample, [(1, 2, 64)] * procs, chunksize=1)     # start workers, imports
        # bounded sample: bands of `rows` x 1440 columns per worker and step, sized so that the whole
        # --steps/--warmup run stays within about three minutes (one calibration step with 2 rows)
        t_cal = time.perf_counter()
        cpu_arm(procs, procs, 2, pool=pool)
        t_cal = time.perf_counter() - t_cal
        rows = int(max(1, min(8, 2 * 180.0 / (t_cal * (steps + max(a.warmup, 0))))))
        for _ in range(max(a.warmup, 0)):
            cpu_arm(procs, procs, rows, pool=pool)
        for _ in range(steps):
            v, w, c = cpu_arm(procs, procs, rows, pool=pool)
            vals.append(v); walls.append(w)
    value = float(np.mean(vals))
    frac = procs * rows * 1440 / 1038240.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "timesteps/s", "n_gpus": a.gpus,
        "steps": steps, "warmup": a.warmup, "ms_per_step": 1000.0 * float(np.mean(walls)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "global 721x1440x137 single timestep, plev19 monthly deltas (BASELINE configs[1])"},
        "cpu_baseline": {"value": value, "unit": "timesteps/s", "cores": procs, "kind": "port",
                         "sample": "per step %d bands of %dx1440 columns (%d columns = %.4f of one global "
                                   "timestep), fp64 oracle port of the reference path (numpy + C column "
                                   "loops), one band per worker process like parallel.IterMP; the verbatim "
                                   "reference needs xarray, absent in this image"
                                   % (procs, rows, procs * rows * 1440, frac)},
        "e2e": {"value": value, "unit": "timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from pgw4era5_b200 import settings, synthetic as S
    from pgw4era5_b200.engine import DeltaSet, PGWEngine
    from pgw4era5_b200 import parallel as P

    settings.i_debug = 0
    if os.environ.get("PGW_BENCH_P_REF"):        # tuning experiments only (changes the workload; not a bench line)
        settings.p_ref_inp = float(os.environ["PGW_BENCH_P_REF"])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from pgw4era5_b200.parallel import bind_to_gpu_numa
    numa = None if os.environ.get("PGW_NO_NUMA_BIND") else bind_to_gpu_numa(local)    # before any pinned allocation
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ny, nx = GRIDS[a.grid]
    ncol = ny * nx
    plev = S.PLEV19

    # ---- inputs: climatology generated on rank 0 and broadcast (the one collective of the path)
    era0 = S.make_era5(ny, nx, 2, device=dev, orog_seed=2)
    deltas = S.make_deltas(era0, 2 if rank == 0 else 99, plev=plev, device=dev)
    ds = DeltaSet(deltas, device=dev)
    del deltas
    bcast_ms = None
    if world > 1:
        bcast_ms = P.broadcast_deltas(ds, src=0)
    eng = PGWEngine(era0["ak"], era0["bk"], ds, soil1=era0["soil1"])
    # ring members: same terrain as the climatology, different weather
    ring = [era0] + [S.make_era5(ny, nx, 100 * (rank + 1) + i, device=dev, orog_seed=2)
                     for i in range(1, a.ring)]
    outs = [eng.alloc_outputs(ny, nx, len(era0["soil1"])) for _ in range(2)]
    base = datetime(2006, 8, 1, 0)
    when = lambda i: base + timedelta(hours=6 * ((i * world + rank) % 124))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    pend = None
    for i in range(a.warmup):
        p = eng.submit(ring[i % a.ring], when(i), out=outs[i % 2], ignore_top_pressure_error=True, slot=i % 2)
        if pend is not None:
            pend.result()
        pend = p
    if pend is not None:
        pend.result()
    n_iters = []

    # ---- timed region: device-resident inputs -> device-resident outputs
    sampler = ClockSampler(local)
    eng.kernel_events = []
    launches0 = eng.stats["launches"]
    barrier()
    sampler.mark_start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pend = None
    for i in range(a.steps):
        p = eng.submit(ring[i % a.ring], when(i), out=outs[i % 2], ignore_top_pressure_error=True, slot=i % 2)
        if pend is not None:
            n_iters.append(pend.result()["n_iter"])
        pend = p
    n_iters.append(pend.result()["n_iter"])
    ev1.record()
    import ctypes
    from pgw4era5_b200 import _native
    kernel_name = ("pgw_column_tma_kernel" if _native.lib.pgw_timestep_uses_tma(ctypes.byref(pend.args)) == 1
                   else "pgw_column_kernel")
    barrier()
    sampler.mark_stop()
    ms = ev0.elapsed_time(ev1)
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in eng.kernel_events]
    eng.kernel_events = None
    launches = eng.stats["launches"] - launches0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * a.steps / (ms / 1000.0)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timing
    e2e = None
    if a.e2e_steps > 0:
        from pgw4era5_b200.hostpipe import HostPipeline
        ns = a.e2e_slots
        pipe = HostPipeline(eng, ny, nx, nslots=ns)
        hin = [pipe.pin_inputs(ring[i % a.ring]) for i in range(ns)]
        houts = [pipe.alloc_host_outputs() for _ in range(ns)]
        for i in range(ns):
            pipe.run(hin[i % ns], when(i), houts[i % ns], ignore_top_pressure_error=True)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(a.e2e_steps):
            pipe.run(hin[i % ns], when(i), houts[i % ns], ignore_top_pressure_error=True)
        pipe.drain()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        t = torch.tensor([el], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * a.e2e_steps / float(t.item()), "unit": "timesteps/s",
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "steps": a.e2e_steps, "slots": ns}

    # ---- roofline of the dominant kernel (the fused column kernel)
    peak, peak_kind = measured_peak_gbs()
    abytes = algorithmic_bytes(ncol, 137, len(plev), len(era0["soil1"]))
    k_ms = float(np.mean(kernel_ms)) if kernel_ms else None
    achieved = abytes / (k_ms / 1000.0) / 1e9 if k_ms else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "column_kernel_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes": abytes,
                "peak_kind": peak_kind}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import build as ob
        ob.build()
        _cpu_sample((1, 2, 64))
        v, wall, cols = cpu_arm(12, 1, rows=8)
        cpu = {"value": v, "unit": "timesteps/s", "cores": 1, "kind": "port",
               "sample": "12 bands of 8x1440 columns (%d columns = %.4f of one global timestep), fp64 oracle "
                         "port on 1 host core, %.1f s" % (cols, cols / 1038240.0, wall)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "timesteps/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 storage; f64 geopotential/ps iteration", "data": "synthetic",
            "config": {"workload": ("global %dx%dx137 single timestep per step, plev19 monthly deltas "
                                    "(BASELINE configs[1]); N>1: timesteps sharded by rank (configs[2])" % (ny, nx))
                       if a.grid == "GL" else
                       ("European subdomain %dx%dx137 single timestep per step, plev19 monthly deltas "
                        "(BASELINE configs[0], not the metric's configuration)" % (ny, nx)),
                       "grid": a.grid, "ring": a.ring,
                       "l2": "inputs cycle through %d distinct 2.3 GB timesteps (>> 126 MB L2)" % a.ring,
                       "parallelism": "timestep-sharded x%d" % world, "numa_node": numa,
                       "n_iter": {"min": int(min(n_iters)), "max": int(max(n_iters)), "steps": len(n_iters)},
                       "engine": dict(eng.stats), "broadcast_ms": bcast_ms},
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", default="GL", choices=sorted(GRIDS))
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--e2e-slots", type=int, default=2, help="timesteps in flight in the host-buffer pipeline")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
