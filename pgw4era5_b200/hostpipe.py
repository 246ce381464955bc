"""
Host-buffer front end of the engine: ERA5 fields arrive in (pinned) host memory,
as they do from a NetCDF reader, and results go back to host memory.  Two slots
with one CUDA stream each keep the PCIe link busy in both directions: while slot
A's results travel device->host, slot B's inputs travel host->device and its
kernel runs.  This is the path `bench.py` reports as ``e2e``.
"""
import ctypes as C

import torch

from . import _native as N

def _pinned_write_combined(n_float):
    """n_float float32 of page-locked, WRITE-COMBINED host memory (cudaHostAllocWriteCombined) as a torch
    tensor, or None.  For H2D staging buffers only: the device reads them without snooping the CPU caches,
    the CPU only ever writes them (reads from write-combined memory are very slow)."""
    import ctypes
    import glob
    import os
    try:
        base = os.path.dirname(torch.__file__)
        cand = glob.glob(os.path.join(base, "lib", "libcudart*.so*")) + \
            glob.glob(os.path.join(os.path.dirname(base), "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        rt = ctypes.CDLL(cand[0] if cand else "libcudart.so")
        ptr = ctypes.c_void_p()
        rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
        if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(4 * n_float), 0x04 | 0x01) != 0 or not ptr.value:
            return None
        buf = (ctypes.c_float * n_float).from_address(ptr.value)
        t = torch.frombuffer(buf, dtype=torch.float32)
        t._pgw_keep = (buf, rt)                      # never freed: lives as long as the process
        return t
    except Exception:
        return None


IN_FIELDS = ("PS", "FIS", "FR_LAND", "FR_SEA_ICE", "T_SKIN", "T_SO", "T", "QV", "U", "V")
OUT_FIELDS = ("PS", "T_SKIN", "FR_SEA_ICE", "T_SO", "T", "QV", "U", "V", "delta_ps")


class HostPipeline:
    def __init__(self, engine, ny, nx, nslots=2):
        self.eng, self.ny, self.nx, self.nslots = engine, ny, nx, nslots
        L, S = engine.nlev, len(engine.soil_decay)
        self.in_levels = dict(PS=1, FIS=1, FR_LAND=1, FR_SEA_ICE=1, T_SKIN=1, T_SO=S, T=L, QV=L, U=L, V=L)
        self.out_levels = dict(PS=1, T_SKIN=1, FR_SEA_ICE=1, T_SO=S, T=L, QV=L, U=L, V=L, delta_ps=1)
        ncol = ny * nx
        self.n_in = sum(self.in_levels.values()) * ncol
        self.n_out = sum(self.out_levels.values()) * ncol
        self.h2d_bytes, self.d2h_bytes = 4 * self.n_in, 4 * self.n_out
        dev = engine.device
        self.slots = []
        for _ in range(nslots):
            din = torch.empty(self.n_in, device=dev, dtype=torch.float32)
            dout = torch.empty(self.n_out, device=dev, dtype=torch.float32)
            self.slots.append(dict(stream=torch.cuda.Stream(device=dev), din=din, dout=dout,
                                   vin=self._views(din, self.in_levels), vout=self._views(dout, self.out_levels),
                                   pending=None, host_out=None, raw=False))
        self.count = 0

    @staticmethod
    def _swap(t):
        """Byte order of a float32 device buffer, in place, on the current stream (NetCDF-3 is big-endian)."""
        N.check(N.lib.pgw_byteswap32(C.c_void_p(t.data_ptr()), t.numel(),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pgw_byteswap32")

    def _views(self, flat, levels):
        out, off = {}, 0
        ncol = self.ny * self.nx
        for name, lev in levels.items():
            out[name] = flat[off:off + lev * ncol].view(1, lev, self.ny, self.nx) if lev > 1 or name == "T_SO" \
                else flat[off:off + ncol].view(1, self.ny, self.nx)
            off += lev * ncol
        return out

    def alloc_host_inputs(self):
        import os
        flat = _pinned_write_combined(self.n_in) if os.environ.get("PGW_PINNED_WC") == "1" else None
        if flat is None:
            flat = torch.empty(self.n_in, dtype=torch.float32, pin_memory=True)
        return dict(flat=flat, **self._views(flat, self.in_levels))

    def alloc_host_outputs(self):
        flat = torch.empty(self.n_out, dtype=torch.float32, pin_memory=True)
        return dict(flat=flat, **self._views(flat, self.out_levels))

    def pin_inputs(self, era):
        """Copy a dict of ERA5 tensors (any device) into one pinned host buffer."""
        h = self.alloc_host_inputs()
        for name in IN_FIELDS:
            h[name].copy_(era[name].reshape(h[name].shape))
        return h

    def _finish(self, slot):
        p = slot["pending"]
        if p is None:
            return None
        with torch.cuda.stream(slot["stream"]):
            before = self.eng.stats["reruns"]
            res = p.result()
            if self.eng.stats["reruns"] != before:          # rerun wrote new device results
                if slot["raw"]:
                    self._swap(slot["dout"])
                slot["host_out"]["flat"].copy_(slot["dout"], non_blocking=True)
            slot["stream"].synchronize()
        slot["pending"] = None
        out = dict(slot["host_out"])
        out["n_iter"] = res["n_iter"]
        return out

    def run(self, host_in, era_step_dt, host_out, raw=False, **kw):
        """Enqueue one timestep: H2D, fused pass, D2H.  Returns the finished result of the
        timestep that previously used this slot (or None).  ``raw``: the host buffers hold the
        big-endian bytes of a NetCDF-3 file (nc3raw); they are swapped on the device after the
        H2D copy and before the D2H copy."""
        slot = self.slots[self.count % self.nslots]
        sid = self.count % self.nslots
        self.count += 1
        done = self._finish(slot)
        with torch.cuda.stream(slot["stream"]):
            slot["din"].copy_(host_in["flat"], non_blocking=True)
            if raw:
                self._swap(slot["din"])
            slot["pending"] = self.eng.submit(slot["vin"], era_step_dt, out=slot["vout"], slot=sid, **kw)
            if raw:
                self._swap(slot["dout"])
            host_out["flat"].copy_(slot["dout"], non_blocking=True)
            slot["host_out"], slot["raw"] = host_out, raw
        return done

    def drain(self):
        """Finish everything in flight; results in submission order (None for idle slots)."""
        order = [(self.count + k) % self.nslots for k in range(self.nslots)]
        return [self._finish(self.slots[i]) for i in order]
