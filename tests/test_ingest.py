"""Binary catalogue ingest: psb_catalog_load must produce, from
a .npy file, the records and sums that the reference's read_ascii_data() produces
from the same numbers (io/read_ascii.c:868-902: w = wcomp * wfkp, sum wcomp,
sum w^2, sum wcomp wfkp^2 n(z)) and read_cata()'s alpha / shot / norm
(src/read_cata.c:160-183).  The records are bit-exact; the sums agree to 1e-13
(different but fixed summation order)."""
import ctypes as C
import os

import numpy as np
import pytest


def _probe(path):
    from powspec_b200.api import load_library
    L = load_library()
    n, nc, el = C.c_size_t(), C.c_int(), C.c_int()
    rc = L.psb_catalog_probe(os.fsencode(str(path)), C.byref(n), C.byref(nc), C.byref(el))
    return rc, n.value, nc.value, el.value, L.psb_last_error().decode()


def test_npy_header_probe(tmp_path):
    a = np.arange(42, dtype=np.float64).reshape(7, 6)
    np.save(tmp_path / "a.npy", a)
    assert _probe(tmp_path / "a.npy")[:4] == (0, 7, 6, 8)
    np.save(tmp_path / "b.npy", a.astype(np.float32))
    assert _probe(tmp_path / "b.npy")[:4] == (0, 7, 6, 4)
    np.save(tmp_path / "empty.npy", np.zeros((0, 4)))
    assert _probe(tmp_path / "empty.npy")[:4] == (0, 0, 4, 8)
    np.save(tmp_path / "f.npy", np.asfortranarray(a))
    rc, *_, msg = _probe(tmp_path / "f.npy")
    assert rc != 0 and "Fortran" in msg
    np.save(tmp_path / "i.npy", a.astype(np.int64))
    rc, *_, msg = _probe(tmp_path / "i.npy")
    assert rc != 0 and "dtype" in msg
    np.save(tmp_path / "v.npy", a.ravel())
    rc, *_, msg = _probe(tmp_path / "v.npy")
    assert rc != 0 and "2-D" in msg
    (tmp_path / "x.npy").write_bytes(b"not a numpy file at all")
    rc, *_, msg = _probe(tmp_path / "x.npy")
    assert rc != 0 and "not a .npy" in msg
    rc, *_, msg = _probe(tmp_path / "missing.npy")
    assert rc != 0 and "cannot open" in msg


@pytest.fixture(scope="module")
def ctx():
    import powspec_b200
    c = powspec_b200.Context(0)
    yield c
    c.close()


def _table(seed, n, dtype=np.float64):
    r = np.random.default_rng(seed)
    ra, dec = np.deg2rad(r.uniform(100, 200, n)), np.deg2rad(r.uniform(-10, 60, n))
    d = np.cbrt(r.uniform(1065.0 ** 3, 2560.0 ** 3, n))
    nz = r.uniform(1e-4, 5e-4, n)
    wc, wf = r.uniform(0.8, 1.2, n), 1 / (1 + 1e4 * nz)
    # junk column in front: columns are selected, not assumed
    return np.c_[r.random(n), d * np.cos(dec) * np.cos(ra), d * np.cos(dec) * np.sin(ra), d * np.sin(dec),
                 wc, wf, nz].astype(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_records_and_sums_match_read_ascii_semantics(dtype, ctx, tmp_path):
    n = 5_000_123            # more than one 2^22-row chunk, ragged tail
    t = _table(1, n, dtype)
    np.save(tmp_path / "cat.npy", t)
    t = t.astype(np.float64)
    cat, sums = ctx.load_catalog(tmp_path / "cat.npy", pos=(1, 2, 3), wcomp=4, wfkp=5, nz=6, issim=False)
    got = ctx.catalog_to_host(cat)
    ctx.free_catalog(cat)
    want = np.c_[t[:, 1:4], t[:, 4] * t[:, 5]]
    assert np.array_equal(got, want)
    assert sums["n"] == n
    assert abs(sums["sumw"] - t[:, 4].sum()) < 1e-13 * t[:, 4].sum()
    w2 = (want[:, 3] ** 2).sum()
    assert abs(sums["sumw2"] - w2) < 1e-13 * w2
    w2n = (t[:, 4] * t[:, 5] * t[:, 5] * t[:, 6]).sum()
    assert abs(sums["sumw2n"] - w2n) < 1e-13 * w2n
    # simulation box: w = wcomp, no FKP / n(z) (io/read_ascii.c:903)
    cat, sums = ctx.load_catalog(tmp_path / "cat.npy", pos=(1, 2, 3), wcomp=4, issim=True)
    got = ctx.catalog_to_host(cat)
    ctx.free_catalog(cat)
    assert np.array_equal(got, np.c_[t[:, 1:4], t[:, 4]])
    assert sums["sumw2"] == 0 and sums["sumw2n"] == 0
    # no weight column: w = 1
    cat, sums = ctx.load_catalog(tmp_path / "cat.npy", pos=(1, 2, 3), issim=True)
    assert sums["sumw"] == n
    ctx.free_catalog(cat)
    # same file twice: same bits
    a = ctx.load_catalog(tmp_path / "cat.npy", pos=(1, 2, 3), wcomp=4, wfkp=5, nz=6, issim=False)
    b = ctx.load_catalog(tmp_path / "cat.npy", pos=(1, 2, 3), wcomp=4, wfkp=5, nz=6, issim=False)
    assert a[1] == b[1]
    ctx.free_catalog(a[0]); ctx.free_catalog(b[0])


@pytest.mark.gpu
def test_ingest_errors(ctx, tmp_path):
    from powspec_b200.api import PowspecB200Error
    np.save(tmp_path / "c.npy", np.zeros((10, 3)))
    with pytest.raises(PowspecB200Error, match="not enough columns"):
        ctx.load_catalog(tmp_path / "c.npy", pos=(0, 1, 2), wcomp=3)
    with pytest.raises(PowspecB200Error, match="cannot open"):
        ctx.load_catalog(tmp_path / "nope.npy")
    np.save(tmp_path / "e.npy", np.zeros((0, 4)))
    cat, sums = ctx.load_catalog(tmp_path / "e.npy", pos=(0, 1, 2), wcomp=3)
    assert sums["n"] == 0 and sums["sumw"] == 0
    ctx.free_catalog(cat)


@pytest.mark.gpu
def test_survey_from_binary_catalogues(ctx, port_oracle, tmp_path):
    """read_cata for .npy catalogues -> genr_mesh -> powspec against the CPU oracle fed
    with the same arrays and the scalars of the reference's read_cata restatement."""
    import powspec_b200
    from oracle.oracle import survey_scalars
    from tests.parity import TOL_DOUBLE, assert_spectra_close
    D, R = _table(11, 120_000), _table(12, 600_000)
    np.save(tmp_path / "d.npy", D)
    np.save(tmp_path / "r.npy", R)
    sc = survey_scalars(D[:, 4], D[:, 5], D[:, 6], R[:, 4], R[:, 5], R[:, 6])
    conf = powspec_b200.Conf(ndata=1, issim=False, gsize=96, assign=3, intlace=False, poles=(0, 2, 4),
                             kbin=0.005, isauto=(True, False), iscross=False)
    cata = ctx.read_cata(conf, [tmp_path / "d.npy"], [tmp_path / "r.npy"], pos=(1, 2, 3), wcomp=4,
                         wfkp=5, nz=6)
    for key in ("wdata", "wrand", "alpha", "shot", "norm"):
        assert abs(getattr(cata, key)[0] - sc[key]) < 1e-12 * abs(sc[key]), key
    mesh = ctx.genr_mesh(conf, cata)
    got = ctx.powspec(conf, cata, mesh)
    ctx.free_catalog(cata.data[0]); ctx.free_catalog(cata.rand[0])
    want = port_oracle.run(np.c_[D[:, 1:4], D[:, 4] * D[:, 5]], rand=[np.c_[R[:, 1:4], R[:, 4] * R[:, 5]]],
                           scalars=[sc], ng=96, assign="PCS", interlace=False, poles=(0, 2, 4),
                           issim=False, kbin=0.005)
    worst = assert_spectra_close(got, want, TOL_DOUBLE, "survey from binary catalogues")
    print(f"survey from .npy catalogues: worst {worst:.2e}")
