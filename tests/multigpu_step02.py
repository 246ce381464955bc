"""
step_02 for one variable split over GPUs by target latitude (third row; run under torchrun, one rank
per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/multigpu_step02.py [--days 365]

Parity: the gathered field equals the single-GPU regridding bit for bit.  Timing (BASELINE configs[3], one daily
3-D variable 365 x 19 x 180 x 360 -> 721 x 1440): broadcast of the source, smoothing, the band kernels, and the
gather of the 28.8 GB result -- which is what makes the split pointless unless the bands stay where they are.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    from pgw4era5_b200 import functions as F, parallel as P
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--plevs", type=int, default=19)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    lat_s, lon_s = np.linspace(-89.5, 89.5, 180), 0.5 + np.arange(360)
    lat_t, lon_t = np.linspace(-90.0, 90.0, 721), 0.25 * np.arange(1440)
    # ---- parity on a small series
    g = torch.Generator(device=dev).manual_seed(4)
    small = torch.randn((12, 3, 180, 360), device=dev, generator=g)
    if rank != 0:
        small = torch.full_like(small, float(rank))
    band, (r0, r1), whole = P.regrid_banded(small, lat_s, lon_s, lat_t, lon_t, smooth=True)
    if rank == 0:
        one = F.regrid_arrays(F.smooth_annual_cycle(small), lat_s, lon_s, lat_t, lon_t)
        assert torch.equal(whole, one), "gathered bands differ from the single-GPU result"
        assert torch.equal(band, one[..., r0:r1, :])
    dist.barrier()
    # ---- timing at the size of BASELINE configs[3]
    src = torch.randn((a.days, a.plevs, 180, 360), device=dev, generator=g)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier()
        ev[0].record()
        dist.broadcast(src, src=0)
        ev[1].record()
        sm = F.smooth_annual_cycle(src)
        ev[2].record()
        r0, r1 = P.split_rows(721, world)[rank]
        band = F.regrid_arrays(sm, lat_s, lon_s, lat_t, lon_t, rows=(r0, r1))
        ev[3].record()
        torch.cuda.synchronize()
    t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(3)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del band
    # the gather of the result (what a single output file needs), timed separately on a tenth of the fields
    nf = max(1, a.days // 10)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    P.regrid_banded(sm[:nf], lat_s, lon_s, lat_t, lon_t, gather=True)
    e1.record()
    torch.cuda.synchronize()
    tg = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    if rank == 0:
        out_bytes = a.days * a.plevs * 721 * 1440 * 4
        print(json.dumps({"workload": "step_02, one daily 3-D variable %d x %d x 180 x 360 -> 721 x 1440 (BASELINE "
                                      "configs[3]), target-latitude bands on %d GPUs" % (a.days, a.plevs, world),
                          "n_gpus": world, "parity": "gathered bands == single-GPU result, bit for bit",
                          "broadcast_source_ms": float(t[0]), "smoothing_ms_replicated": float(t[1]),
                          "regrid_band_ms": float(t[2]), "output_bytes": out_bytes,
                          "broadcast_plus_gather_regrid_of_%d_fields_ms" % (nf * a.plevs): float(tg[0])}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
