"""world_size-2 gloo test (CPU) of the slab-decomposition ORCHESTRATION in
powspec_b200/distributed.py: particle routing (all-to-all-v), halo-plane
exchange, and the FFT transpose.  The per-rank kernels are replaced by a numpy
engine with the same interface (CIC deposit with halos, numpy FFTs), so what is
checked is the host logic: split sizes, neighbours, buffer layouts.  The CUDA
engine is checked against the same driver in tests/test_gpu_slab.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from powspec_b200.distributed import HALO_HI, HALO_LO, SlabShape, TorchComm, density_to_kspace

NG, L, NPART = 16, 100.0, 4000


class NumpyEngine:
    """CPU stand-in for GpuSlabEngine (CIC only), torch CPU tensors in/out."""

    def __init__(self, nranks, rank):
        self.shape = SlabShape(NG, nranks, rank, 8)

    def partition(self, particles):
        p = particles.numpy().reshape(-1, 4)
        cx = np.minimum((p[:, 0] * NG / L).astype(int), NG - 1)
        owner = cx // self.shape.nx
        order = np.argsort(owner, kind="stable")
        counts = [int((owner == r).sum()) for r in range(self.shape.nranks)]
        return torch.from_numpy(p[order].copy().reshape(-1)), counts

    def alloc_meshes(self, cat=0):
        s = self.shape
        return [torch.zeros((s.planes, s.ng, s.rowlen), dtype=torch.float64)]

    def assign_into(self, meshes, flat):
        s = self.shape
        p = flat.numpy().reshape(-1, 4)
        mesh = meshes[0].numpy()
        xbase = 0 if s.nranks == 1 else (s.rank * s.nx - HALO_LO) % NG
        t = p[:, :3] * NG / L
        c = np.floor(t).astype(int)
        d = t - c
        for a in (0, 1):
            for b in (0, 1):
                for e in (0, 1):
                    w = p[:, 3] * np.where(a, d[:, 0], 1 - d[:, 0]) * np.where(b, d[:, 1], 1 - d[:, 1]) \
                        * np.where(e, d[:, 2], 1 - d[:, 2])
                    lp = ((c[:, 0] + a) % NG - xbase) % NG
                    assert lp.max(initial=0) < s.planes
                    np.add.at(mesh, (lp, (c[:, 1] + b) % NG, (c[:, 2] + e) % NG), w)

    def add_into(self, dst, src):
        dst += src

    def fft_yz_pack(self, mesh):
        s = self.shape
        owned = mesh[s.lo:s.lo + s.nx, :, :NG].numpy()
        f = np.fft.rfftn(owned, axes=(1, 2))                    # (nx, ng, ngk)
        ny = s.nx
        blocks = [f[:, q * ny:(q + 1) * ny, :] for q in range(s.nranks)]    # [q][xl][yl][k]
        packed = np.ascontiguousarray(np.stack(blocks))
        return torch.from_numpy(packed.view(np.float64).reshape(-1).copy())

    def recv_view(self, mesh):
        s = self.shape
        return torch.empty(s.nx * s.ng * s.ngk * 2, dtype=torch.float64)

    def fft_x(self, buf):
        s = self.shape
        a = buf.numpy().view(np.complex128).reshape(NG, s.nx, s.ngk)    # (x, y-slab, k)
        a[...] = np.fft.fft(a, axis=0)


def _full_reference(parts):
    p = np.concatenate(parts)
    eng = NumpyEngine(1, 0)
    meshes = eng.alloc_meshes()
    eng.assign_into(meshes, torch.from_numpy(p.reshape(-1)))
    return np.fft.rfftn(meshes[0].numpy()[:, :, :NG])


def _catalogue(rank):
    rng = np.random.default_rng(100 + rank)
    p = np.c_[rng.random((NPART, 3)) * L, rng.uniform(0.5, 1.5, NPART)]
    p[0, :3] = [L * (1 - 1e-12), 0.0, L / 2]       # wraps around the periodic boundary
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = TorchComm()
        eng = NumpyEngine(world, rank)
        cat = torch.from_numpy(_catalogue(rank))
        fk = density_to_kspace(eng, comm, [cat[:1500], cat[1500:1500], cat[1500:]])[0]     # chunked
        got = fk.numpy().view(np.complex128).reshape(NG, eng.shape.nx, eng.shape.ngk)
        want = _full_reference([_catalogue(r) for r in range(world)])
        want = want[:, rank * eng.shape.nx:(rank + 1) * eng.shape.nx, :]
        err = np.abs(got - want).max() / np.abs(want).max()
        q.put((rank, float(err)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_slab_orchestration_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert sorted(res) == list(range(world))
    assert max(res.values()) < 1e-12, res


def test_single_rank_is_the_identity_layout():
    class NoComm:
        size, rank = 1, 0
    eng = NumpyEngine(1, 0)
    fk = density_to_kspace(eng, NoComm(), torch.from_numpy(_catalogue(0)))[0]
    got = fk.numpy().view(np.complex128).reshape(NG, NG, NG // 2 + 1)
    want = _full_reference([_catalogue(0)])
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
