"""Host mirror of the in-library slab decomposition (csrc/dist.cu).

``Group``: all ranks in ONE process (one host thread per rank inside the library, peer
copies) — what the reference's single-process C host gets through genr_mesh()/powspec()
with POWSPEC_B200_DEVICES; listing a device several times gives virtual ranks (tests).

``NcclRank``: one rank of a one-process-per-GPU job (bench.py under torchrun); the
library issues its own NCCL calls, the host only distributes the 128-byte unique id
(here through torch.distributed, any other channel works).

The reference has no distributed path (src/genr_mesh.c:650-747 allocates one address
space); results are checked against the single-GPU path and the oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import (POWSPEC_ERR_MESH, POWSPEC_ERR_PK, Cata, Conf, Context, PowspecB200Error, _Cats,
                  _err, _ptr_of, load_library, pk_from_result)

STAGES = ["route", "assign", "halo", "fft_zy", "transpose", "fft_x", "bin", "reduce"]


def _stage_ms(L, d):
    ms = (C.c_double * len(STAGES))()
    L.psb_dist_timings(d, ms, len(STAGES))
    return {k: ms[i] for i, k in enumerate(STAGES)}


def _traffic(L, d):
    v = (C.c_double * 3)()
    L.psb_dist_traffic(d, v, 3)
    return {"transpose_bytes_sent": v[0], "route_bytes_sent": v[1], "peer_stores": bool(v[2])}


class Group:
    """psb_group: genr_mesh() + powspec() of one mesh over several (virtual) ranks."""

    def __init__(self, devices):
        self.L = load_library()
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        self.h = self.L.psb_group_create(arr, len(self.devices))
        if not self.h:
            raise _err(self.L, "psb_group_create", POWSPEC_ERR_MESH)

    def close(self):
        if getattr(self, "h", None):
            self.L.psb_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        if self.L.psb_group_set_option(self.h, name.encode(), int(value)):
            raise _err(self.L, "psb_group_set_option")

    def _cats(self, conf: Conf, cata: Cata):
        cats, keep, spaces = _Cats(), [], set()
        for i in range(cata.num):
            ptr, n, sp, ka = _ptr_of(cata.data[i])
            keep.append(ka)
            spaces.add(sp)
            cats.data[i], cats.ndata[i] = ptr, n
            if cata.wdata is not None:
                cats.wdata[i] = float(cata.wdata[i])
            elif sp == 0:
                cats.wdata[i] = float(np.sum(np.asarray(cata.data[i])[:, 3]))
            else:
                raise PowspecB200Error("Cata.wdata is required for device-resident catalogues")
        if len(spaces) != 1:
            raise PowspecB200Error("all catalogues must live in the same memory space")
        cats.memspace = spaces.pop()
        return cats, keep

    def genr_mesh(self, conf: Conf, cata: Cata):
        cats, keep = self._cats(conf, cata)
        p = conf._c()
        if self.L.psb_group_mesh(self.h, C.byref(p), C.byref(cats)):
            raise _err(self.L, "genr_mesh", POWSPEC_ERR_MESH)
        self._conf = conf
        return self

    def powspec(self, conf: Conf, cata=None, mesh=None):
        p = conf._c()
        r = self.L.psb_group_power(self.h, C.byref(p))
        if not r:
            raise _err(self.L, "powspec", POWSPEC_ERR_PK)
        try:
            pk = pk_from_result(self.L, r, conf)
        finally:
            self.L.psb_result_free(r)
        pk.stages_ms = [_stage_ms(self.L, self.L.psb_group_rank(self.h, q)) for q in range(len(self.devices))]
        pk.traffic = _traffic(self.L, self.L.psb_group_rank(self.h, 0))
        return pk

    def run(self, conf: Conf, cata: Cata):
        self.genr_mesh(conf, cata)
        return self.powspec(conf)


class NcclRank:
    """psb_dist over NCCL: this process is rank `rank` of `nranks`."""

    def __init__(self, ctx: Context, nranks: int, rank: int, unique_id: bytes):
        self.L = ctx.L
        self.ctx = ctx
        self.nranks, self.rank = nranks, rank
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.h = self.L.psb_dist_create_nccl(ctx.h, nranks, rank, buf)
        if not self.h:
            raise _err(self.L, "psb_dist_create_nccl", POWSPEC_ERR_MESH)

    @staticmethod
    def unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        if L.psb_dist_unique_id(buf):
            raise _err(L, "psb_dist_unique_id")
        return buf.raw

    @classmethod
    def from_torch(cls, ctx: Context, group=None):
        """Rank / size from torch.distributed; the unique id travels through it."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", ctx.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0, group=group)
        return cls(ctx, world, rank, bytes(t.cpu().numpy().tobytes()))

    def close(self):
        if getattr(self, "h", None):
            self.L.psb_dist_destroy(self.h)
            self.h = None

    def set_option(self, name, value):
        if self.L.psb_dist_set_option(self.h, name.encode(), int(value)):
            raise _err(self.L, "psb_dist_set_option")

    def begin(self, conf: Conf):
        p = conf._c()
        if self.L.psb_dist_begin(self.h, C.byref(p)):
            raise _err(self.L, "psb_dist_begin", POWSPEC_ERR_MESH)
        self._conf = conf

    def add(self, cat: int, particles):
        """particles: (n, 4) float64 CUDA tensor or (device_ptr, n)."""
        ptr, n, sp, keep = _ptr_of(particles)
        if n and sp != 1:
            raise PowspecB200Error("psb_dist_add takes device-resident particles")
        if self.L.psb_dist_add(self.h, cat, ptr, n):
            raise _err(self.L, "psb_dist_add", POWSPEC_ERR_MESH)

    def add_host(self, cat: int, particles, nchunks: int = 1):
        """particles: (n, 4) float64 host array / CPU tensor (pinned or pageable)."""
        ptr, n, sp, keep = _ptr_of(particles)
        if n and sp != 0:
            raise PowspecB200Error("psb_dist_add_host takes host-resident particles")
        if self.L.psb_dist_add_host(self.h, cat, ptr, n, int(nchunks)):
            raise _err(self.L, "psb_dist_add_host", POWSPEC_ERR_MESH)

    def finish(self, wdata):
        w = (C.c_double * 2)(*(list(wdata) + [0.0])[:2])
        r = self.L.psb_dist_finish(self.h, w)
        if not r:
            raise _err(self.L, "psb_dist_finish", POWSPEC_ERR_PK)
        try:
            pk = pk_from_result(self.L, r, self._conf)
        finally:
            self.L.psb_result_free(r)
        pk.stages_ms = _stage_ms(self.L, self.h)
        pk.traffic = _traffic(self.L, self.h)
        pk.timings_ms = self.ctx.timings()
        pk.launches = int(self.L.psb_launch_count(self.ctx.h))
        return pk

    def run(self, conf: Conf, shares, wdata):
        """shares: per catalogue this rank's particles (tensor) or a list of chunks."""
        self.begin(conf)
        for i, sh in enumerate(shares):
            for chunk in (sh if isinstance(sh, (list, tuple)) and not (
                    len(sh) == 2 and isinstance(sh[0], int)) else [sh]):
                self.add(i, chunk)
        return self.finish(wdata)
