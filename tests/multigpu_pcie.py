"""
Bare host<->device copy bandwidth with N ranks copying at the same time (run under torchrun, one rank per GPU; not
collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29513 tests/multigpu_pcie.py [--mb 2206]

This is the platform bound of bench.py's `e2e` leg (2.3 GB in + 2.3 GB out per global timestep and rank): pinned
H2D and D2H copies of that size on two streams, all ranks at once, for three ways of getting page-locked memory:
torch's pinned allocator (cudaHostAlloc), write-combined cudaHostAlloc for the H2D source, and anonymous memory
with transparent huge pages registered with cudaHostRegister.  Prints one JSON line (rank 0): GB/s per direction
and rank (min / mean over ranks) for H2D alone, D2H alone and both at once.
"""
import argparse
import ctypes
import glob
import json
import mmap
import os
import time

import torch
import torch.distributed as dist


def cudart():
    base = os.path.dirname(torch.__file__)
    cand = glob.glob(os.path.join(base, "lib", "libcudart*.so*")) + \
        glob.glob(os.path.join(os.path.dirname(base), "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    return ctypes.CDLL(cand[0] if cand else "libcudart.so")


def host_buffer(kind, nbytes, rt):
    """float32 CPU tensor of nbytes page-locked bytes, or None if this kind is not available."""
    n = nbytes // 4
    if kind == "torch_pinned":
        return torch.empty(n, dtype=torch.float32, pin_memory=True)
    if kind == "write_combined":
        ptr = ctypes.c_void_p()
        rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
        if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), 0x04 | 0x01) != 0 or not ptr.value:
            return None
        buf = (ctypes.c_float * n).from_address(ptr.value)
        t = torch.frombuffer(buf, dtype=torch.float32)
        t._keep = (buf, rt)
        return t
    if kind == "thp_registered":
        try:
            m = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            if hasattr(mmap, "MADV_HUGEPAGE"):
                m.madvise(mmap.MADV_HUGEPAGE)
            t = torch.frombuffer(m, dtype=torch.float32)
            t.zero_()                                  # fault the pages in (as huge pages where the kernel allows)
            rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
            if rt.cudaHostRegister(ctypes.c_void_p(t.data_ptr()), ctypes.c_size_t(nbytes), 0x01) != 0:
                return None
            t._keep = m
            return t
        except Exception:
            return None
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2206, help="MB per direction and copy (2206 = one global timestep)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt = cudart()
    nbytes = a.mb * 1000 * 1000 // 4096 * 4096
    d_in = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
    d_out = torch.zeros(nbytes // 4, device=dev, dtype=torch.float32)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    out = {}
    for kind in ("torch_pinned", "write_combined", "thp_registered"):
        h_in = host_buffer(kind, nbytes, rt)
        h_out = host_buffer("torch_pinned" if kind == "write_combined" else kind, nbytes, rt)   # never READ write-combined memory
        ok = torch.tensor([float(h_in is not None and h_out is not None)], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            out[kind] = None
            continue
        if kind != "thp_registered":
            h_in.zero_()

        def h2d():
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        res = {}
        for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
            for f in fns:
                f()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.reps):
                for f in fns:
                    f()
            torch.cuda.synchronize()
            gbs = nbytes * a.reps / (time.perf_counter() - t0) / 1e9
            t = torch.tensor([gbs, -gbs], device=dev, dtype=torch.float64)
            s = torch.tensor([gbs], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                dist.all_reduce(s)
            res[name] = {"min": float(t[0]), "max": float(-t[1]), "mean": float(s.item() / world)}
        out[kind] = res
        del h_in, h_out
    if rank == 0:
        print(json.dumps({"n_ranks": world, "mb_per_copy": nbytes / 1e6, "host_cores": os.cpu_count(),
                          "unit": "GB/s per direction and rank", "results": out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
