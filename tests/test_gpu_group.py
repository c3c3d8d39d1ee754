"""GPU tests of the slab decomposition driven from INSIDE the library (csrc/dist.cu):
psb_group runs all ranks in one process, one host thread per rank, exchanges by peer
copies.  Several virtual ranks on ONE device exercise the routing, the halo ring, the
FFT whose y pass stores the transposed layout (send buffers or straight into the
peers' buffers), the overlapped transposes and the reduction — and must reproduce the
oracle and the single-GPU path.  The same per-rank code runs over NCCL on real
multi-GPU boxes (tests/test_gpu_multi.py, bench.py --gpus N)."""
import numpy as np
import pytest

from tests.parity import TOL_DOUBLE, TOL_SINGLE, assert_spectra_close

pytestmark = pytest.mark.gpu


def _conf(pb, ncat, ng, assign, interlace, poles, box, kbin, precision=8, los=(0, 0, 1)):
    return pb.Conf(ndata=ncat, issim=True, bsize=(box,) * 3, gsize=ng,
                   assign=pb.powspec_assign_names.index(assign), intlace=interlace, poles=poles,
                   kbin=kbin, isauto=(True, ncat == 2), iscross=ncat == 2, precision=precision, los=los)


@pytest.mark.parametrize("nranks", [1, 2, 4])
@pytest.mark.parametrize("assign,interlace", [("TSC", True), ("PCS", True), ("CIC", False), ("NGP", True)])
def test_group_matches_oracle(nranks, assign, interlace, port_oracle):
    import powspec_b200 as pb
    from powspec_b200.dist import Group
    rng = np.random.default_rng(17)
    n, box, ng = 60_000, 300.0, 32
    d = np.c_[rng.random((n, 3)) * box, rng.uniform(0.5, 1.5, n)]
    d[0, :3] = [box * (1 - 1e-13), box * (1 - 1e-13), 0.0]      # wraps over the periodic ring
    kw = dict(ng=ng, assign=assign, interlace=interlace, poles=(0, 2, 4), box=box, kbin=0.02)
    want = port_oracle.run(d, **kw)
    conf = _conf(pb, 1, ng, assign, interlace, (0, 2, 4), box, 0.02)
    g = Group([0] * nranks)
    try:
        g.set_option("group_chunk", 9000)          # several upload chunks per rank
        got = g.run(conf, pb.Cata(data=[d]))
        assert_spectra_close(got, want, TOL_DOUBLE, f"group G={nranks} {assign}")
        # a second run on the same group (buffers and peer tables reused)
        got = g.run(conf, pb.Cata(data=[d]))
        assert_spectra_close(got, want, TOL_DOUBLE, f"group G={nranks} {assign} (rerun)")
    finally:
        g.close()


def test_group_cross_spectra_and_single_precision(port_oracle):
    import powspec_b200 as pb
    from powspec_b200.dist import Group
    rng = np.random.default_rng(18)
    box, ng = 200.0, 24
    a = np.c_[rng.random((30_000, 3)) * box, np.ones(30_000)]
    b = np.c_[rng.random((20_000, 3)) * box, rng.uniform(0.5, 1.5, 20_000)]
    kw = dict(ng=ng, assign="TSC", interlace=True, poles=(0, 1, 2), box=box, kbin=0.03, los=(0.6, 0.0, 0.8))
    want = port_oracle.run([a, b], **kw)
    for prec, tol in ((8, TOL_DOUBLE), (4, TOL_SINGLE)):
        conf = _conf(pb, 2, ng, "TSC", True, (0, 1, 2), box, 0.03, precision=prec, los=(0.6, 0.0, 0.8))
        g = Group([0, 0, 0])
        try:
            got = g.run(conf, pb.Cata(data=[a, b]))
        finally:
            g.close()
        assert_spectra_close(got, want, tol, f"group cross prec={prec}")


@pytest.mark.parametrize("p2p", [1, 0])
@pytest.mark.parametrize("precision", [8, 4])
def test_group_512_hand_written_passes_match_single_gpu(p2p, precision):
    """At 512^3 the hand-written FFT passes run: the y pass writes the transposed layout
    (peer stores when p2p, send buffers + all-to-all otherwise).  4 virtual ranks against
    the one-GPU path on the same device-resident catalogue."""
    import powspec_b200 as pb
    from powspec_b200.dist import Group
    n, box, ng = 3_000_000, 1000.0, 512
    ctx = pb.Context(0)
    dev = ctx.generate_catalog(n, box, kind=1, seed=11)
    conf = pb.Conf(ndata=1, issim=True, bsize=(box,) * 3, gsize=ng, assign=2, intlace=True,
                   poles=(0, 2, 4), kbin=0.01, precision=precision)
    cata = pb.Cata(data=[dev], wdata=[float(n)])
    want = ctx.powspec(conf, cata, ctx.genr_mesh(conf, cata))
    g = Group([0, 0, 0, 0])
    try:
        g.set_option("p2p", p2p)
        got = g.run(conf, cata)
        assert got.traffic["peer_stores"] == bool(p2p)
        assert got.traffic["transpose_bytes_sent"] > 0
    finally:
        g.close()
        ctx.free_catalog(dev)
        ctx.close()
    assert_spectra_close(got, want, 1e-9 if precision == 8 else 2e-5, f"group 512 p2p={p2p}")


def test_group_rejects_particles_outside_the_box():
    """def_box's checks (src/genr_mesh.c:516-531) on the slab path: the reference fails
    for a coordinate < 0 or >= BOX_SIZE; so must every rank."""
    import powspec_b200 as pb
    from powspec_b200.dist import Group
    rng = np.random.default_rng(3)
    box = 100.0
    d = np.c_[rng.random((5000, 3)) * box, np.ones(5000)]
    conf = _conf(pb, 1, 16, "TSC", True, (0, 2), box, 0.05)
    for bad, msg in (((-0.5, 1.0, 1.0), "below 0"), ((1.0, box, 1.0), "not smaller than BOX_SIZE")):
        e = d.copy()
        e[1234, :3] = bad
        g = Group([0, 0])
        try:
            with pytest.raises(pb.PowspecB200Error, match=msg):
                g.run(conf, pb.Cata(data=[e]))
            # the group stays usable after a failed run
            ok = g.run(conf, pb.Cata(data=[d]))
            assert ok.nbin > 0
        finally:
            g.close()
