"""Coordinate conversion (cnvt_coord, src/cnvt_coord.c:549-582).

CPU part: the restatement in oracle/pspec_port.c against the goldens written by the
UNMODIFIED reference (tests/golden/make_cnvt_golden.py), and — where the prebuilt
reference library travels with the repo — the reference itself against its goldens.
GPU part (-m gpu): the device conversion through the C ABI against the same goldens,
and a survey power spectrum from (RA, Dec, z) catalogues converted on the device
against the oracle fed with reference-converted coordinates.

Tolerance: a coordinate may differ by a few ulp of the comoving distance (CUDA's
sin/cos/pow vs glibc's); written as 1e-14 relative to the distance."""
import os

import numpy as np
import pytest

from tests.golden.cnvt_cases import CNVT_CASES, cnvt_inputs, distance_table

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cnvt_golden.npz")
TOL_COORD = 1e-14


@pytest.fixture(scope="module")
def golden():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def rel_coord_err(got, ref):
    dist = np.linalg.norm(ref[:, :3], axis=1)
    return float((np.abs(got[:, :3] - ref[:, :3]).max(axis=1) / dist).max())


@pytest.mark.parametrize("case", CNVT_CASES, ids=[c["name"] for c in CNVT_CASES])
def test_port_cnvt_matches_reference_goldens(case, golden):
    from oracle.oracle import port_cnvt
    kw = dict(case["cosmo"])
    if case.get("table"):
        kw["samples"] = distance_table(case)
    out, order = port_cnvt(cnvt_inputs(case), **kw)
    assert (order == 0) == bool(case.get("table"))
    for i, o in enumerate(out):
        ref = golden[f"{case['name']}_{i}"]
        assert np.array_equal(o[:, 3], ref[:, 3])
        assert rel_coord_err(o, ref) < 1e-15


def test_reference_cnvt_reproduces_its_goldens(golden, tmp_path):
    from oracle.oracle import have_ref_cnvt, ref_cnvt
    if not have_ref_cnvt():
        pytest.skip("oracle/_ref/libpowspec_ref_cnvt.so not built (needs /root/reference)")
    case = CNVT_CASES[0]
    out = ref_cnvt(cnvt_inputs(case), **case["cosmo"])
    assert np.array_equal(out[0], golden[f"{case['name']}_0"])


def test_port_cnvt_rejects_negative_redshift():
    from oracle.oracle import port_cnvt
    a = cnvt_inputs(CNVT_CASES[0])[0]
    a[17, 2] = -0.01
    with pytest.raises(RuntimeError):
        port_cnvt([a], **CNVT_CASES[0]["cosmo"])


def test_library_order_selection_matches_oracle_on_cpu():
    """The host-side order selection of the product library (no GPU needed) against the
    reference-pinned restatement over a grid of cosmologies, error bounds and ranges."""
    import ctypes as C

    from oracle.oracle import port_cnvt
    from powspec_b200.api import Conf, load_library
    L = load_library()
    rng = np.random.default_rng(0)
    seen = set()
    for om, ok, w in ((0.31, 0.0, -1.0), (0.25, 0.05, -1.0), (0.3, 0.0, -0.8), (0.35, -0.02, -1.1)):
        for err in (1e-4, 1e-6, 1e-8, 1e-10):
            for zlo, zhi in ((0.0, 0.2), (0.4, 1.1), (0.8, 3.0), (0.5, 0.5)):
                conf = Conf(cnvt=True, omega_m=om, omega_l=1 - om - ok, omega_k=ok, eos_w=w, ecdst=err)
                cosmo, _ = conf._cosmo()
                got = L.psb_cnvt_order(C.byref(cosmo), zlo, zhi)
                a = np.zeros((64, 4))
                a[:, 2] = np.r_[zlo, zhi, rng.uniform(zlo, zhi, 62)]
                _, want = port_cnvt([a], omega_m=om, omega_l=1 - om - ok, omega_k=ok, eos_w=w, ecdst=err)
                assert got == want, (om, ok, w, err, zlo, zhi, got, want)
                seen.add(got)
    assert len(seen) >= 4           # the grid exercises several orders
    conf = Conf(cnvt=True, ecdst=1e-8)
    cosmo, _ = conf._cosmo()
    assert L.psb_cnvt_order(C.byref(cosmo), -0.1, 1.0) == -1


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx():
    import powspec_b200
    c = powspec_b200.Context(0)
    yield c
    c.close()


def _conf(case, **extra):
    from powspec_b200.api import Conf
    kw = dict(case["cosmo"])
    if case.get("table"):
        kw["fcdst"] = distance_table(case)
    return Conf(cnvt=True, dcnvt=(True, True), rcnvt=(True, True), **kw, **extra)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CNVT_CASES, ids=[c["name"] for c in CNVT_CASES])
def test_device_cnvt_matches_reference_goldens(case, golden, ctx):
    import torch
    from oracle.oracle import port_cnvt
    arrays = cnvt_inputs(case)
    kw = dict(case["cosmo"])
    if case.get("table"):
        kw["samples"] = distance_table(case)
    _, want_order = port_cnvt(arrays, **kw)
    dev = [torch.from_numpy(a).cuda() for a in arrays]
    order = ctx.cnvt_coord(_conf(case), dev)
    assert order == want_order
    for i, t in enumerate(dev):
        got = t.cpu().numpy()
        ref = golden[f"{case['name']}_{i}"]
        assert np.array_equal(got[:, 3], ref[:, 3])
        err = rel_coord_err(got, ref)
        print(f"{case['name']}[{i}]: order {order}, max coordinate error {err:.2e} of the distance")
        assert err < TOL_COORD


@pytest.mark.gpu
def test_device_cnvt_error_conditions(ctx):
    import torch
    from powspec_b200.api import PowspecB200Error
    case = CNVT_CASES[0]
    a = cnvt_inputs(case)[0]
    a[5, 2] = -1e-3
    with pytest.raises(PowspecB200Error, match="negative redshift"):
        ctx.cnvt_coord(_conf(case), [torch.from_numpy(a).cuda()])
    # outside the tabulated range: HUGE_VAL as in the reference (src/cnvt_coord.c:107)
    tcase = next(c for c in CNVT_CASES if c.get("table"))
    b = cnvt_inputs(tcase)[0][:8].copy()
    b[3, 2] = 99.0
    t = torch.from_numpy(b).cuda()
    ctx.cnvt_coord(_conf(tcase), [t])
    assert np.isinf(t.cpu().numpy()[3, :3]).any()


@pytest.mark.gpu
def test_survey_from_sky_coordinates_on_device(ctx, tmp_path):
    """Survey P_0/P_2 from (RA, Dec, z) catalogues: conversion on the device inside
    genr_mesh (host arrays untouched) against the CPU oracle fed with coordinates
    converted by the reference-pinned restatement."""
    import powspec_b200
    from oracle import load_oracle
    from oracle.oracle import port_cnvt, survey_scalars
    from tests.parity import TOL_DOUBLE, assert_spectra_close
    from powspec_b200.api import Cata

    def sky(seed, n):
        r = np.random.default_rng(seed)
        a = np.empty((n, 4))
        a[:, 0] = r.uniform(110.0, 250.0, n)
        a[:, 1] = np.rad2deg(np.arcsin(r.uniform(np.sin(np.deg2rad(-5)), np.sin(np.deg2rad(60)), n)))
        a[:, 2] = r.uniform(0.4, 1.0, n)
        nz = r.uniform(1e-4, 5e-4, n)
        wc, wf = r.uniform(0.8, 1.2, n), 1 / (1 + 1e4 * nz)
        a[:, 3] = wc * wf
        return a, (wc, wf, nz)
    D, dc = sky(31, 150_000)
    R, rc = sky(32, 700_000)
    sc = survey_scalars(*dc, *rc)
    cosmo = dict(omega_m=0.31, omega_l=0.69, omega_k=0.0, eos_w=-1.0, ecdst=1e-8)
    (Dx, Rx), _ = port_cnvt([D, R], **cosmo)
    kw = dict(ng=128, assign="TSC", interlace=True, poles=(0, 2), issim=False, kbin=0.005)
    want = load_oracle("port").run(Dx, rand=[Rx], scalars=[sc], **kw)
    D0, R0 = D.copy(), R.copy()
    conf = powspec_b200.Conf(ndata=1, issim=False, gsize=128, assign=2, intlace=True, poles=(0, 2),
                             kbin=0.005, isauto=(True, False), iscross=False, cnvt=True,
                             dcnvt=(True, False), rcnvt=(True, False), **cosmo)
    cata = Cata(data=[D], rand=[R], wdata=[sc["wdata"]], wrand=[sc["wrand"]], alpha=[sc["alpha"]],
                shot=[sc["shot"]], norm=[sc["norm"]])
    mesh = ctx.genr_mesh(conf, cata)
    got = ctx.powspec(conf, cata, mesh)
    assert np.array_equal(D, D0) and np.array_equal(R, R0)
    worst = assert_spectra_close(got, want, TOL_DOUBLE, "survey from sky coordinates")
    print(f"survey from (RA, Dec, z) on the device: worst {worst:.2e}, cnvt {got.timings_ms.get('cnvt'):.3f} ms")
