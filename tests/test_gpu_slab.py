"""GPU tests of the slab-decomposed path: N virtual ranks emulated on one device
(the same driver logic as the NCCL run, local copies instead of collectives)
must reproduce the single-mesh result and the oracle."""
import numpy as np
import pytest

from tests.parity import TOL_DOUBLE, TOL_SINGLE, assert_spectra_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import powspec_b200
    c = powspec_b200.Context(0)
    yield c
    c.close()


def _conf(pb, ncat, ng, assign, interlace, poles, box, kbin, precision=8, los=(0, 0, 1)):
    return pb.Conf(ndata=ncat, issim=True, bsize=(box,) * 3, gsize=ng,
                   assign=pb.powspec_assign_names.index(assign), intlace=interlace, poles=poles,
                   kbin=kbin, isauto=(True, ncat == 2), iscross=ncat == 2, precision=precision, los=los)


@pytest.mark.parametrize("nranks", [1, 2, 4])
@pytest.mark.parametrize("assign,interlace", [("TSC", True), ("PCS", True), ("CIC", False), ("NGP", True)])
def test_emulated_slabs_match_oracle(nranks, assign, interlace, ctx, port_oracle):
    import torch

    import powspec_b200 as pb
    from powspec_b200.distributed import GpuSlabEngine, slab_power_emulated
    rng = np.random.default_rng(17)
    n, box, ng = 60_000, 300.0, 32
    d = np.c_[rng.random((n, 3)) * box, rng.uniform(0.5, 1.5, n)]
    d[0, :3] = [box * (1 - 1e-13), box * (1 - 1e-13), 0.0]
    kw = dict(ng=ng, assign=assign, interlace=interlace, poles=(0, 2, 4), box=box, kbin=0.02)
    want = port_oracle.run(d, **kw)
    conf = _conf(pb, 1, ng, assign, interlace, (0, 2, 4), box, 0.02)
    engines = [GpuSlabEngine(ctx, conf, nranks, r) for r in range(nranks)]
    # every virtual rank starts with an arbitrary share of the catalogue
    shares = np.array_split(d, nranks)
    cats = [[torch.from_numpy(s).cuda()] for s in shares]
    got = slab_power_emulated(engines, cats, [float(d[:, 3].sum())])
    assert_spectra_close(got, want, TOL_DOUBLE, f"slab G={nranks} {assign}")


def test_emulated_slabs_cross_and_single_precision(ctx, port_oracle):
    import torch

    import powspec_b200 as pb
    from powspec_b200.distributed import GpuSlabEngine, slab_power_emulated
    rng = np.random.default_rng(18)
    box, ng = 200.0, 24
    a = np.c_[rng.random((30_000, 3)) * box, np.ones(30_000)]
    b = np.c_[rng.random((20_000, 3)) * box, rng.uniform(0.5, 1.5, 20_000)]
    kw = dict(ng=ng, assign="TSC", interlace=True, poles=(0, 1, 2), box=box, kbin=0.03, los=(0.6, 0.0, 0.8))
    want = port_oracle.run([a, b], **kw)
    for prec, tol in ((8, TOL_DOUBLE), (4, TOL_SINGLE)):
        conf = _conf(pb, 2, ng, "TSC", True, (0, 1, 2), box, 0.03, precision=prec, los=(0.6, 0.0, 0.8))
        engines = [GpuSlabEngine(ctx, conf, 2, r) for r in range(2)]
        cats = [[torch.from_numpy(x).cuda() for x in (sa, sb)]
                for sa, sb in zip(np.array_split(a, 2), np.array_split(b, 2))]
        got = slab_power_emulated(engines, cats, [float(a[:, 3].sum()), float(b[:, 3].sum())])
        assert_spectra_close(got, want, tol, f"slab cross prec={prec}")


@pytest.mark.parametrize("seed", range(24))
def test_emulated_slabs_on_random_configurations(seed, ctx, port_oracle):
    """The seeded random box configurations of tests/fuzz_cases.py through the slab
    path, over the largest admissible rank count (GRID_SIZE divisible, at least 3
    owned planes per rank): odd mesh sizes, slabs as thin as the halo, non-cubic
    boxes, any line of sight / bins / multipoles, cross spectra."""
    import torch

    import powspec_b200 as pb
    from powspec_b200.distributed import GpuSlabEngine, slab_power_emulated
    from tests.fuzz_cases import fuzz_case
    from tests.parity import noise_floor
    cats, kw = fuzz_case(seed)
    ng = kw["ng"]
    ranks = [g for g in (8, 6, 5, 4, 3, 2) if ng % g == 0 and ng // g >= 3]
    if not ranks:
        pytest.skip(f"GRID_SIZE {ng} has no admissible slab count")
    G = ranks[0]
    want = port_oracle.run(cats if len(cats) > 1 else cats[0], **kw)
    ncat = len(cats)
    isauto = kw.get("isauto", [True] * ncat)
    conf = pb.Conf(ndata=ncat, issim=True, bsize=tuple(np.broadcast_to(np.asarray(kw["box"], float), (3,))),
                   gsize=ng, assign=pb.powspec_assign_names.index(kw["assign"]), intlace=kw["interlace"],
                   poles=kw["poles"], kbin=kw["kbin"], kmin=kw.get("kmin", 0.0), kmax=kw.get("kmax", -1.0),
                   logscale=kw.get("logscale", False), los=kw["los"],
                   isauto=tuple(isauto) + (False,) * (2 - ncat), iscross=kw.get("iscross", False))
    engines = [GpuSlabEngine(ctx, conf, G, r) for r in range(G)]
    shares = [np.array_split(c, G) for c in cats]
    per_rank = [[torch.from_numpy(np.ascontiguousarray(shares[c][r])).cuda() for c in range(ncat)]
                for r in range(G)]
    got = slab_power_emulated(engines, per_rank, [float(c[:, 3].sum()) for c in cats],
                              isauto=list(isauto), iscross=kw.get("iscross", False))
    worst = assert_spectra_close(got, want, TOL_DOUBLE, f"slab fuzz {seed} G={G}: {kw}",
                                 abs_floor=noise_floor(want, kw["poles"]))
    print(f"slab fuzz {seed} G={G} ng={ng}: worst rel err {worst:.2e}")
