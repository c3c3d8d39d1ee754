"""Real multi-GPU run of the slab-decomposed path over NCCL (needs >= 2 GPUs;
skipped on the single-GPU box).  Each rank starts with an arbitrary share of the
catalogue; the result must match the oracle."""
import os
import socket

import numpy as np
import pytest

from tests.parity import TOL_DOUBLE, assert_spectra_close

pytestmark = pytest.mark.gpu

NG, BOX, N = 64, 400.0, 200_000
KW = dict(ng=NG, assign="TSC", interlace=True, poles=(0, 2, 4), box=BOX, kbin=0.02)


def _catalogue():
    rng = np.random.default_rng(23)
    return np.c_[rng.random((N, 3)) * BOX, rng.uniform(0.5, 1.5, N)]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import powspec_b200 as pb
    from powspec_b200.distributed import GpuSlabEngine, TorchComm, slab_power
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        d = _catalogue()
        share = np.array_split(d, world)[rank]
        ctx = pb.Context(rank)
        conf = pb.Conf(ndata=1, issim=True, bsize=(BOX,) * 3, gsize=NG, assign=2, intlace=True,
                       poles=(0, 2, 4), kbin=0.02, device=rank)
        eng = GpuSlabEngine(ctx, conf, world, rank)
        pk = slab_power(eng, TorchComm(), [torch.from_numpy(share).cuda()], [float(d[:, 3].sum())])
        if rank == 0:
            q.put(dict(nbin=pk.nbin, nl=pk.nl, k=pk.k, kedge=pk.kedge, km=pk.km, cnt=pk.cnt,
                       lcnt=pk.lcnt, pl=pk.pl, xpl=pk.xpl, shot=pk.shot, norm=pk.norm))
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_nccl_slabs_match_oracle(port_oracle):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0

    class R:
        pass
    r = R()
    r.__dict__.update(got)
    want = port_oracle.run(_catalogue(), **KW)
    assert_spectra_close(r, want, TOL_DOUBLE, f"nccl slabs x{world}")


def _worker_lib(rank, world, port, q, p2p):
    """The in-library path: the library issues the NCCL calls itself (csrc/dist.cu)."""
    import torch
    import torch.distributed as dist

    import powspec_b200 as pb
    from powspec_b200.dist import NcclRank
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        d = _catalogue()
        share = np.array_split(d, world)[rank]
        ctx = pb.Context(rank)
        conf = pb.Conf(ndata=1, issim=True, bsize=(BOX,) * 3, gsize=NG, assign=2, intlace=True,
                       poles=(0, 2, 4), kbin=0.02, device=rank)
        eng = NcclRank.from_torch(ctx)
        eng.set_option("p2p", p2p)
        halves = np.array_split(share, 2)           # two chunks per rank
        pk = eng.run(conf, [[torch.from_numpy(h).cuda() for h in halves]], [float(d[:, 3].sum())])
        if rank == 0:
            q.put(dict(nbin=pk.nbin, nl=pk.nl, k=pk.k, kedge=pk.kedge, km=pk.km, cnt=pk.cnt,
                       lcnt=pk.lcnt, pl=pk.pl, xpl=pk.xpl, shot=pk.shot, norm=pk.norm,
                       traffic=pk.traffic))
        eng.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [0, 1])
def test_library_nccl_slabs_match_oracle(port_oracle, p2p):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker_lib, args=(r, world, port, q, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0

    class R:
        pass
    r = R()
    r.__dict__.update(got)
    want = port_oracle.run(_catalogue(), **KW)
    assert_spectra_close(r, want, TOL_DOUBLE, f"library nccl slabs x{world} p2p={p2p}")
