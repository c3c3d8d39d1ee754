"""Slab-decomposed power spectrum of ONE mesh over several GPUs.

The reference has no distributed path (one address space,
src/genr_mesh.c:650-747); this is new design for meshes that exceed or strain one
GPU (BASELINE configs 4 and 5).  One process per GPU; `torch.distributed` (NCCL
over NVLink/NVSwitch) carries the four exchanges, the per-rank work is the C ABI
of libpowspec_b200.so (`psb_slab_*`, include/powspec_b200.h):

    route      particles -> owner of their base x-cell        all-to-all-v
    assign     scatter into the slab buffer (owned + halos)   psb_slab_assign
    halo       1 plane down, 3 planes up, added by the owner  neighbour send/recv + psb_add
    fft (y,z)  batched 2-D r2c on the owned x-planes          psb_slab_fft_yz
    transpose  (x-slab, y, k) -> (x, y-slab, k)               psb_slab_pack + all-to-all
    fft (x)    1-D c2c on the y-slab                          psb_slab_fft_x
    bin        fused combine/window/L_l(mu)/reduce, y-slab    psb_slab_bin
    reduce     nl*nbin power sums                             allreduce(sum)
    finish     mode counts (pure geometry, every rank) + normalisation   psb_slab_finish

The result stays in the transposed (y-slab) layout: binning is layout-agnostic,
so no transpose back.

The orchestration (`density_to_kspace`, `slab_power`) is written against two
small interfaces — a communicator and a per-rank "engine" — so that the same
code runs
  * on N GPUs (`TorchComm` + `GpuSlabEngine`),
  * emulated on one GPU with N virtual ranks (`slab_power_emulated`; GPU tests), and
  * on CPU with gloo and a numpy engine (tests/test_distributed_cpu.py), which
    checks the exchange logic (split sizes, neighbours, transpose layout).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

HALO_LO, HALO_HI = 1, 3


# ---------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------
class TorchComm:
    """torch.distributed (nccl on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.has_a2a = dist.get_backend(group) != "gloo"

    def all_to_all_v(self, send, send_counts, row_elems):
        """send: 1-D tensor grouped by destination; counts in rows of row_elems."""
        import torch
        dist = self.dist
        sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
        rc = torch.empty_like(sc)
        self._a2a_equal(rc, sc)
        recv_counts = [int(x) for x in rc.tolist()]
        recv = torch.empty(sum(recv_counts) * row_elems, dtype=send.dtype, device=send.device)
        ins = [c * row_elems for c in send_counts]
        outs = [c * row_elems for c in recv_counts]
        if self.has_a2a:
            dist.all_to_all_single(recv, send, output_split_sizes=outs, input_split_sizes=ins,
                                   group=self.group)
        else:
            self._a2a_p2p(recv, outs, send, ins)
        return recv, recv_counts

    def all_to_all(self, recv, send):
        self._a2a_equal(recv, send)

    def _a2a_equal(self, recv, send):
        if self.has_a2a:
            self.dist.all_to_all_single(recv, send, group=self.group)
        else:
            n = send.numel() // self.size
            self._a2a_p2p(recv, [n] * self.size, send, [n] * self.size)

    def _a2a_p2p(self, recv, outs, send, ins):
        import torch
        dist = self.dist
        so = np.concatenate([[0], np.cumsum(ins)])
        ro = np.concatenate([[0], np.cumsum(outs)])
        ops = []
        for q in range(self.size):
            if q == self.rank:
                recv[ro[q]:ro[q + 1]] = send[so[q]:so[q + 1]]
                continue
            if ins[q]:
                ops.append(dist.P2POp(dist.isend, send[so[q]:so[q + 1]].contiguous(), q, self.group))
            if outs[q]:
                ops.append(dist.P2POp(dist.irecv, recv[ro[q]:ro[q + 1]], q, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if recv.is_cuda:
            torch.cuda.synchronize()

    def halo_exchange(self, to_prev, to_next):
        """send `to_prev` to rank-1 and `to_next` to rank+1 (periodic); returns
        (from_next, from_prev) with the shapes of what the neighbours sent."""
        import torch
        dist = self.dist
        prv, nxt = (self.rank - 1) % self.size, (self.rank + 1) % self.size
        from_next = torch.empty_like(to_prev)
        from_prev = torch.empty_like(to_next)
        ops = [dist.P2POp(dist.isend, to_prev, prv, self.group),
               dist.P2POp(dist.isend, to_next, nxt, self.group),
               dist.P2POp(dist.irecv, from_next, nxt, self.group),
               dist.P2POp(dist.irecv, from_prev, prv, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if from_next.is_cuda:
            torch.cuda.synchronize()
        return from_next, from_prev

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t


# ---------------------------------------------------------------------------
# per-rank engines
# ---------------------------------------------------------------------------
@dataclass
class SlabShape:
    ng: int
    nranks: int
    rank: int
    precision: int

    @property
    def nx(self):
        return self.ng // self.nranks

    @property
    def ngk(self):
        return self.ng // 2 + 1

    @property
    def rowlen(self):
        return 2 * self.ngk

    @property
    def planes(self):
        return self.ng if self.nranks == 1 else self.nx + HALO_LO + HALO_HI

    @property
    def plane_elems(self):
        return self.ng * self.rowlen

    @property
    def lo(self):           # index of the first owned plane in the buffer
        return 0 if self.nranks == 1 else HALO_LO


class _Slab(C.Structure):
    _fields_ = [("nranks", C.c_int), ("rank", C.c_int)]


class GpuSlabEngine:
    """The C ABI building blocks for one (real or virtual) rank on one GPU."""

    def __init__(self, ctx, conf, nranks, rank):
        import torch
        self.torch = torch
        self.ctx, self.conf = ctx, conf
        self.L = ctx.L
        self.par = conf._c()
        self.shape = SlabShape(conf.gsize, nranks, rank, conf.precision)
        self.slab = _Slab(nranks, rank)
        self.device = torch.device("cuda", ctx.device)
        self.rdtype = torch.float64 if conf.precision == 8 else torch.float32
        self._cache = {}
        L = self.L
        if not hasattr(L, "_slab_ready"):
            vp, P, S = C.c_void_p, C.POINTER(type(self.par)), C.POINTER(_Slab)
            L.psb_slab_mesh_elems.restype = C.c_size_t
            L.psb_slab_mesh_elems.argtypes = [P, S]
            L.psb_slab_partition.argtypes = [vp, P, C.c_int, vp, C.c_size_t, vp, C.POINTER(C.c_size_t)]
            L.psb_slab_assign.argtypes = [vp, P, S, vp, C.c_size_t, C.c_double, vp, vp]
            L.psb_add.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
            L.psb_slab_fft_yz.argtypes = [vp, P, S, vp]
            L.psb_slab_pack.argtypes = [vp, P, S, vp, vp]
            L.psb_slab_fft_x.argtypes = [vp, P, S, vp]
            L.psb_slab_bin.argtypes = [vp, P, S, vp, vp, vp, vp, vp]
            L.psb_slab_finish.restype = vp
            L.psb_slab_finish.argtypes = [vp, P, vp, vp, vp, C.POINTER(C.c_double)]
            L._slab_ready = True

    def _chk(self, rc, what):
        if rc:
            from .api import _err
            raise _err(self.L, what)

    def _enter(self):
        # torch (allocation memsets, copies, NCCL) works on its own streams and
        # the library on the context's stream; every psb_slab_* call drains its
        # stream before returning, so one device synchronisation on entry makes
        # the hand-over safe in both directions
        self.torch.cuda.synchronize(self.device)

    # route
    def partition(self, particles):
        t = self.torch
        self._enter()
        n = particles.shape[0]
        out = t.empty_like(particles)
        counts = (C.c_size_t * self.shape.nranks)()
        self._chk(self.L.psb_slab_partition(self.ctx.h, C.byref(self.par), self.shape.nranks,
                                            particles.data_ptr(), n, out.data_ptr(), counts),
                  "psb_slab_partition")
        return out.reshape(-1), [int(c) for c in counts]

    # Persistent buffers: the slab buffers (one set per catalogue) double as the
    # receive buffers of the transpose (their content is dead once packed), and
    # one send buffer is shared by all fields.  Nothing of this size is allocated
    # or freed inside a run, so the peak footprint is fixed:
    #   ncat * nfields * (nx + halos) planes  +  one packed field.
    def _cached(self, key, shape, dtype):
        t = self.torch
        buf = self._cache.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            self._cache[key] = buf = t.empty(shape, dtype=dtype, device=self.device)
        return buf

    def alloc_meshes(self, cat=0):
        """slab buffers [field] as (planes, ng, rowlen) tensors, zeroed"""
        s = self.shape
        nf = 2 if self.conf.intlace else 1
        out = []
        for f in range(nf):
            m = self._cached(("mesh", cat, f), (s.planes, s.ng, s.rowlen), self.rdtype)
            m.zero_()
            out.append(m)
        return out

    # scatter routed particles (flat n*4 tensor) into the slab buffers (accumulates)
    def assign_into(self, meshes, particles_flat):
        n = particles_flat.numel() // 4
        if n == 0:
            return
        self._enter()
        self._chk(self.L.psb_slab_assign(self.ctx.h, C.byref(self.par), C.byref(self.slab),
                                         particles_flat.data_ptr(), n, 1.0, meshes[0].data_ptr(),
                                         meshes[1].data_ptr() if len(meshes) == 2 else None),
                  "psb_slab_assign")

    def add_into(self, dst, src):
        self._enter()
        self._chk(self.L.psb_add(self.ctx.h, dst.data_ptr(), src.data_ptr(), src.numel(),
                                 self.conf.precision), "psb_add")

    # 2-D FFT on owned planes + pack for the transpose; returns the send buffer.
    # A single rank needs no transpose: (x-slab, y, k) is already (x, y-slab, k).
    def fft_yz_pack(self, mesh):
        s = self.shape
        self._enter()
        owned = mesh[s.lo:s.lo + s.nx]
        self._chk(self.L.psb_slab_fft_yz(self.ctx.h, C.byref(self.par), C.byref(self.slab),
                                         owned.data_ptr()), "psb_slab_fft_yz")
        if s.nranks == 1:
            return mesh.reshape(-1)
        send = self._cached("send", (s.nx * s.ng * s.ngk * 2,), self.rdtype)
        self._chk(self.L.psb_slab_pack(self.ctx.h, C.byref(self.par), C.byref(self.slab),
                                       owned.data_ptr(), send.data_ptr()), "psb_slab_pack")
        return send

    def recv_view(self, mesh):
        """The transpose is received into the slab buffer it came from."""
        s = self.shape
        return mesh.reshape(-1)[:s.nx * s.ng * s.ngk * 2]

    def fft_x(self, buf):
        self._enter()
        self._chk(self.L.psb_slab_fft_x(self.ctx.h, C.byref(self.par), C.byref(self.slab),
                                        buf.data_ptr()), "psb_slab_fft_x")

    def bin(self, fa, fb):
        """fa/fb: lists [field0, field1?] of transposed k-space buffers."""
        t = self.torch
        nl = len(self.conf.poles)
        nbin = self.nbin()
        pl = t.zeros(nl * nbin, dtype=t.float64, device=self.device)
        il = self.conf.intlace
        self._enter()
        self._chk(self.L.psb_slab_bin(self.ctx.h, C.byref(self.par), C.byref(self.slab),
                                      fa[0].data_ptr(), fa[1].data_ptr() if il else None,
                                      fb[0].data_ptr(), fb[1].data_ptr() if il else None,
                                      pl.data_ptr()), "psb_slab_bin")
        return pl

    def nbin(self):
        """powspec_init's bin count (src/multipole.c:335-351), host arithmetic."""
        c = self.conf
        bmax = max(c.bsize)
        kny = np.pi * c.gsize / bmax
        if c.logscale:
            kny = np.log10(kny)
        kmax = kny if not (c.kmax > 0 and kny > c.kmax) else c.kmax
        # C's round(): half away from zero
        x = (kmax - c.kmin) / c.kbin
        nb = int(np.floor(x + 0.5)) if x >= 0 else int(np.ceil(x - 0.5))
        if c.kmin + c.kbin * nb > kny:
            nb -= 1
        return nb

    def finish(self, pl, xpl, wdata):
        """pl: list of per-catalogue allreduced tensors (or None), xpl likewise."""
        from .api import _err, pk_from_result
        host = [None if p is None else np.ascontiguousarray(p.cpu().numpy()) for p in pl] + \
               [None if xpl is None else np.ascontiguousarray(xpl.cpu().numpy())]
        while len(host) < 3:
            host.insert(1, None)
        w = (C.c_double * 2)(*(list(wdata) + [0.0])[:2])
        ptr = [None if h is None else h.ctypes.data for h in host]
        r = self.L.psb_slab_finish(self.ctx.h, C.byref(self.par), ptr[0], ptr[1], ptr[2], w)
        if not r:
            raise _err(self.L, "psb_slab_finish")
        try:
            return pk_from_result(self.L, r, self.conf)
        finally:
            self.L.psb_result_free(r)


# ---------------------------------------------------------------------------
# the orchestration, shared by all back ends
# ---------------------------------------------------------------------------
class _Prof:
    """Optional host-side stage timers (PSB_SLAB_PROFILE=1): device-synchronised
    wall time per stage, accumulated per process; for development only."""

    def __init__(self):
        import os
        self.on = bool(os.environ.get("PSB_SLAB_PROFILE"))
        self.t = {}

    def __call__(self, name):
        prof = self

        class _S:
            def __enter__(self_):
                if prof.on:
                    import time

                    import torch
                    if torch.cuda.is_available():
                        torch.cuda.synchronize()
                    self_.t0 = time.perf_counter()

            def __exit__(self_, *a):
                if prof.on:
                    import time

                    import torch
                    if torch.cuda.is_available():
                        torch.cuda.synchronize()
                    prof.t[name] = prof.t.get(name, 0.0) + time.perf_counter() - self_.t0
        return _S()


PROF = _Prof()


def _chunks(particles):
    """A catalogue is a (n, 4) tensor or an iterable of such chunks (catalogues
    larger than one GPU's memory are generated / read chunk by chunk)."""
    if hasattr(particles, "shape"):
        yield particles
    else:
        yield from particles


def density_to_kspace(engine, comm, particles, cat=0):
    """One catalogue: this rank's particles (any distribution; a tensor or an
    iterable of chunks) -> list over fields of this rank's y-slab of delta(k),
    shape (Ng_x, ny, Ngk) complex, flattened."""
    s = engine.shape
    with PROF("zero_meshes"):
        meshes = engine.alloc_meshes(cat)
    for chunk in _chunks(particles):
        # 1. route the particles to the owner of their base x-cell
        with PROF("partition"):
            sorted_p, counts = engine.partition(chunk)
        with PROF("route_a2av"):
            if comm.size > 1:
                mine, _ = comm.all_to_all_v(sorted_p, counts, 4)
            else:
                mine = sorted_p
        del sorted_p
        # 2. scatter into the slab buffer (owned planes + halo planes)
        with PROF("assign"):
            engine.assign_into(meshes, mine)
        del mine
    out = []
    for mesh in meshes:
        # 3. halo planes go to their owners and are added there
        if comm.size > 1:
            with PROF("halo"):
                to_prev = mesh[0:HALO_LO].contiguous()
                to_next = mesh[s.lo + s.nx:s.lo + s.nx + HALO_HI].contiguous()
                from_next, from_prev = comm.halo_exchange(to_prev, to_next)
                engine.add_into(mesh[s.lo + s.nx - HALO_LO:s.lo + s.nx], from_next)
                engine.add_into(mesh[s.lo:s.lo + HALO_HI], from_prev)
        # 4./5. 2-D FFT of the owned planes, pack, transpose (received into the
        # slab buffer itself: its content is dead once packed)
        with PROF("fft_yz_pack"):
            send = engine.fft_yz_pack(mesh)
        with PROF("transpose_a2a"):
            if comm.size > 1:
                recv = engine.recv_view(mesh)
                comm.all_to_all(recv, send)
            else:
                recv = send
        # 6. 1-D FFT along x on the y-slab
        with PROF("fft_x"):
            engine.fft_x(recv)
        out.append(recv)
    return out


def slab_power(engine, comm, catalogues, wdata, isauto=None, iscross=None):
    """catalogues: list (1 or 2) of this rank's share of each catalogue, (n,4)
    device tensors in any distribution; wdata: GLOBAL sum of weights per catalogue."""
    nc = len(catalogues)
    if isauto is None:
        isauto = [True] * nc
    if iscross is None:
        iscross = nc == 2
    fk = [density_to_kspace(engine, comm, p, cat=i) for i, p in enumerate(catalogues)]
    pl = [None, None]
    with PROF("bin_allreduce"):
        for i in range(nc):
            if isauto[i]:
                pl[i] = comm.all_reduce_sum(engine.bin(fk[i], fk[i])) if comm.size > 1 else engine.bin(fk[i], fk[i])
        xpl = None
        if iscross and nc == 2:
            xpl = engine.bin(fk[0], fk[1])
            if comm.size > 1:
                xpl = comm.all_reduce_sum(xpl)
    with PROF("finish"):
        return engine.finish(pl, xpl, wdata)


# ---------------------------------------------------------------------------
# single-process emulation of N ranks (GPU tests on one device)
# ---------------------------------------------------------------------------
def slab_power_emulated(engines, catalogues_per_rank, wdata, isauto=None, iscross=None):
    """engines: one per virtual rank (same GPU); catalogues_per_rank[r][c]."""
    import torch
    G = len(engines)
    nc = len(catalogues_per_rank[0])
    s0 = engines[0].shape
    fk = [[None] * nc for _ in range(G)]
    for c in range(nc):
        parts = [engines[r].partition(catalogues_per_rank[r][c]) for r in range(G)]
        mine = []
        for r in range(G):
            chunks = []
            for q in range(G):
                sp, cnt = parts[q]
                off = sum(cnt[:r]) * 4
                chunks.append(sp[off:off + cnt[r] * 4])
            mine.append(torch.cat(chunks))
        meshes = [engines[r].alloc_meshes(c) for r in range(G)]
        for r in range(G):
            engines[r].assign_into(meshes[r], mine[r])
        nf = len(meshes[0])
        fk_c = [[None] * nf for _ in range(G)]
        for f in range(nf):
            if G > 1:
                halos = [(meshes[r][f][0:HALO_LO].clone(),
                          meshes[r][f][s0.lo + s0.nx:s0.lo + s0.nx + HALO_HI].clone()) for r in range(G)]
                for r in range(G):
                    s = engines[r].shape
                    engines[r].add_into(meshes[r][f][s.lo + s.nx - HALO_LO:s.lo + s.nx], halos[(r + 1) % G][0])
                    engines[r].add_into(meshes[r][f][s.lo:s.lo + HALO_HI], halos[(r - 1) % G][1])
            sends = [engines[r].fft_yz_pack(meshes[r][f]).clone() for r in range(G)]
            blk = sends[0].numel() // G
            for r in range(G):
                recv = torch.cat([sends[q][r * blk:(r + 1) * blk] for q in range(G)]) if G > 1 else sends[0]
                engines[r].fft_x(recv)
                fk_c[r][f] = recv
        for r in range(G):
            fk[r][c] = fk_c[r]
    if isauto is None:
        isauto = [True] * nc
    if iscross is None:
        iscross = nc == 2
    pl = [None, None]
    for i in range(nc):
        if isauto[i]:
            pl[i] = sum(engines[r].bin(fk[r][i], fk[r][i]) for r in range(G))
    xpl = None
    if iscross and nc == 2:
        xpl = sum(engines[r].bin(fk[r][0], fk[r][1]) for r in range(G))
    return engines[0].finish(pl, xpl, wdata)
