"""world_size-2 checks of the multi-rank logic on CPU with the gloo backend:
(1) latitude-band mode: each rank iterates its own band, the per-iteration error vectors are
    MAX-all-reduced and every rank derives the same global iteration count N as the reference's
    field-global stopping rule (step_03_apply_to_era.py:189,308) -- using the oracle as the
    per-band worker since no GPU is available here;
(2) timestep sharding + broadcast of the delta climatology from rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import ERA_DATE, make_case, run_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _band(era, deltas, rows):
    sub = {k: (v[..., rows[0]:rows[1], :] if isinstance(v, torch.Tensor) and v.dim() >= 3 else v) for k, v in era.items()}
    subd = {k: dict(v, data=v["data"][..., rows[0]:rows[1], :]) for k, v in deltas.items()}
    return sub, subd


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pgw4era5_b200 import parallel as P
    ny, nx, kfix = 14, 24, 9
    era, deltas = make_case(ny, nx, 3)

    # ---- (2) broadcast: non-zero ranks start from garbage and must end with rank 0's climatology
    tensors = [v["data"].clone() for v in deltas.values()]
    if rank != 0:
        for t in tensors:
            t.fill_(float(rank))
    for t in tensors:
        dist.broadcast(t, src=0)
    for t, v in zip(tensors, deltas.values()):
        assert torch.equal(torch.nan_to_num(t), torch.nan_to_num(v["data"]))
    mine = P.timesteps_for_rank(11, rank, world)
    counts = torch.zeros(11)
    counts[mine] = 1
    dist.all_reduce(counts)
    assert torch.equal(counts, torch.ones(11))                      # every timestep exactly once

    # ---- (1) latitude bands with a global stopping rule
    rows = P.split_rows(ny, world)[rank]
    sub, subd = _band(era, deltas, rows)
    loc = run_oracle(sub, subd, n_iter_fixed=kfix)
    errs = torch.tensor(loc["phi_max_errors"], dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    n_glob = P.decide_n_iter(errs.tolist(), 0.15)
    np.save(os.path.join(out_dir, "band%d.npy" % rank), loc["ps_traj"][n_glob - 1])
    np.save(os.path.join(out_dir, "n%d.npy" % rank), np.array([n_glob, P.decide_n_iter(loc["phi_max_errors"], 0.15)]))
    np.save(os.path.join(out_dir, "errs%d.npy" % rank), errs.numpy())
    dist.destroy_process_group()


def test_latband_and_broadcast_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    era, deltas = make_case(14, 24, 3)
    ref = run_oracle(era, deltas)
    n = [np.load(str(tmp_path / ("n%d.npy" % r))) for r in range(world)]
    assert n[0][0] == n[1][0] == ref["n_iter"]                        # same global count on every rank
    np.testing.assert_allclose(np.load(str(tmp_path / "errs0.npy"))[:ref["n_iter"]], ref["phi_max_errors"],
                               rtol=1e-12)
    ps = np.concatenate([np.load(str(tmp_path / ("band%d.npy" % r))) for r in range(world)], axis=1)
    np.testing.assert_allclose(ps, ref["PS"], rtol=0, atol=1e-9)      # bands reproduce the global field
    # a purely local rule would have stopped at least one band earlier or later than the global one
    assert max(n[0][1], n[1][1]) == ref["n_iter"]


def _regrid_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pgw_oracle as O
    from pgw4era5_b200 import parallel as P
    lat_s, lon_s = np.linspace(-87.5, 87.5, 36), 2.5 + 5.0 * np.arange(72)
    lat_t, lon_t = np.linspace(-90, 90, 37), np.arange(0, 360, 2.5)
    rng = np.random.default_rng(7)
    field = rng.normal(size=(3, 2, 36, 72)).astype(np.float32)
    data = torch.from_numpy(field) if rank == 0 else torch.full(field.shape, float(rank))   # garbage off rank 0

    def oracle_band(d, la, lo, tla, tlo, rows):       # stands in for the CUDA operator on this GPU-less box
        full = O.regrid_lat_lon(d.numpy().astype(np.float64), la, lo, tla, tlo)
        return torch.from_numpy(full[..., rows[0]:rows[1], :].astype(np.float32))
    band, (r0, r1), whole = P.regrid_banded(data, lat_s, lon_s, lat_t, lon_t, src=0, device="cpu",
                                            regrid_fn=oracle_band)
    assert (r0, r1) == P.split_rows(37, world)[rank] and band.shape == (3, 2, r1 - r0, 144)
    if rank == 0:
        np.save(os.path.join(out_dir, "whole.npy"), whole.numpy())
        np.save(os.path.join(out_dir, "src.npy"), field)
    else:
        assert whole is None
    dist.destroy_process_group()


def test_regrid_banded_world3(tmp_path):
    """SURVEY 8e row 3 on CPU (gloo, world 3: uneven bands 12 + 12 + 13): broadcast of the source field,
    target-latitude bands, gather -- equals the oracle's regridding of the whole field."""
    from oracle import pgw_oracle as O
    world = 3
    mp.spawn(_regrid_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    whole = np.load(str(tmp_path / "whole.npy"))
    src = np.load(str(tmp_path / "src.npy"))
    lat_s, lon_s = np.linspace(-87.5, 87.5, 36), 2.5 + 5.0 * np.arange(72)
    lat_t, lon_t = np.linspace(-90, 90, 37), np.arange(0, 360, 2.5)
    ref = O.regrid_lat_lon(src.astype(np.float64), lat_s, lon_s, lat_t, lon_t).astype(np.float32)
    np.testing.assert_array_equal(whole, ref)
