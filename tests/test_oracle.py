"""CPU tests of the oracle itself (no GPU): the restatement (port) and, where it
has been built, the unmodified reference, against the committed golden vectors
and the known-answer tests of SURVEY.md §4."""
import ctypes as C

import numpy as np
import pytest

from tests.golden.cases import CASES, SINGLE_CASES, run_case
from tests.parity import assert_spectra_close, rel_err


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_port_matches_golden(case, port_oracle, golden_inputs, golden_outputs):
    res = run_case(port_oracle, case, golden_inputs)
    worst = assert_spectra_close(res, golden_outputs["double"][case["name"]], 1e-9, case["name"])
    assert worst < 1e-9


@pytest.mark.parametrize("name", ["goldenA_tsc_il", "sim_cross_pcs_il", "survey_cross_tsc_il",
                                  "survey_pcs_il_allpoles"])
def test_ref_matches_golden(name, ref_oracle, golden_inputs, golden_outputs):
    """The unmodified reference reproduces its own goldens.  Survey cases run on
    ONE OpenMP thread: the reference's get_coord_bound has a data race — only the
    first of the six `if`s that merge the thread-private bounds is inside the
    `omp critical` (src/genr_mesh.c:481-490) — so with several threads the box of
    a survey (and with it every P_l) changes from run to run at the 1e-3 level.
    The goldens were generated single-threaded."""
    case = next(c for c in CASES if c["name"] == name)
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(1 if "rand" in case else 4)
    try:
        res = run_case(ref_oracle, case, golden_inputs)
    finally:
        gomp.omp_set_num_threads(4)
    assert_spectra_close(res, golden_outputs["double"][name], 1e-10, name)


def test_survey_md_goldens(golden_outputs):
    """The 10-digit rows recorded in SURVEY.md §4 (Golden-A / Golden-B), produced
    there by the reference with two different DFT stubs."""
    a = golden_outputs["double"]["goldenA_tsc_il"]
    rows = [(19, -18.80953655, -350.1501814, 137.1561747),
            (128, -36.51085411, 35.77215525, 249.2895076),
            (314, 34.36493294, -39.24915112, 117.8939132),
            (584, 28.313801, 42.23127954, 9.535708838),
            (1058, -14.29849952, 73.33924604, 18.23635977)]
    for b, (n, p0, p2, p4) in enumerate(rows):
        assert a["cnt"][b] == n
        for l, p in enumerate((p0, p2, p4)):
            assert float("%.10g" % a["pl"][0][l][b]) == p
    assert a["shot"][0] == 500 and a["norm"][0] == 4
    bq = golden_outputs["double"]["goldenB_cic"]
    rows = [(-19.90381542, -346.3965631), (-33.31877869, 23.46801578),
            (40.39806627, -25.61724895), (42.03995688, 47.29885383),
            (2.616811546, 54.94575496)]
    for b, (p0, p2) in enumerate(rows):
        assert float("%.10g" % bq["pl"][0][0][b]) == p0
        assert float("%.10g" % bq["pl"][0][1][b]) == p2


def test_kat_counts(golden_outputs):
    """KAT-counts / KAT-kavg of SURVEY.md §4 (bin 0 includes the DC mode, Q3)."""
    g = golden_outputs["double"]["sim_counts_256"]
    assert g["nbin"] == 80
    assert g["cnt"][:6] == [19, 128, 314, 584, 1058, 1640]
    g = golden_outputs["double"]["goldenA_tsc_il"]
    assert g["cnt"] == [19, 128, 314, 584, 1058]
    assert np.allclose(g["km"][:3], [0.0759622644, 0.165986824, 0.2586445572], rtol=1e-9)


def test_streaming_mode_counts(port_oracle, golden_outputs):
    lib = port_oracle.lib
    lib.oracle_mode_counts.restype = C.c_int
    lib.oracle_mode_counts.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double,
                                       C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    cnt = np.zeros(512, dtype=np.uint64)
    km = np.zeros(512)
    bs = (C.c_double * 3)(1000.0, 1000.0, 1000.0)
    nb = lib.oracle_mode_counts(256, bs, 0.0, -1.0, 0.01, 0, 512, cnt.ctypes.data, km.ctypes.data)
    g = golden_outputs["double"]["sim_counts_256"]
    assert nb == g["nbin"]
    assert cnt[:nb].tolist() == g["cnt"]
    assert rel_err(km[:nb], g["km"]) < 1e-11  # summation order (thread-dependent in the reference too)


@pytest.mark.parametrize("assign", ["NGP", "CIC", "TSC", "PCS"])
@pytest.mark.parametrize("interlace", [False, True])
def test_kat_lattice(assign, interlace, port_oracle):
    """One particle on every grid point: delta(k != 0) = 0, so
    P_l = -(2l+1) * shot * lcnt_l / cnt  (src/multipole.c:1073-1083)."""
    ng, L = 12, 60.0
    g = (np.arange(ng) + 0.25) * (L / ng)
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    d = np.c_[xyz, np.ones(len(xyz))]
    r = port_oracle.run(d, ng=ng, assign=assign, interlace=interlace, poles=(0, 2, 4), box=L, kbin=0.1)
    shot = r.shot[0]
    for l, ell in enumerate((0, 2, 4)):
        want = -(2 * ell + 1) * shot * r.lcnt[l] / r.cnt
        assert np.allclose(r.pl[0][l], want, rtol=0, atol=1e-9 * shot)


def test_kat_delta(port_oracle):
    """A single particle at the origin with NGP: |delta(k)| = 1 for every mode,
    so raw - shot*lcnt/cnt cancels (SURVEY.md §4 KAT-delta)."""
    d = np.array([[0.0, 0.0, 0.0, 1.0]])
    r = port_oracle.run(d, ng=16, assign="NGP", interlace=False, poles=(0, 2, 4), box=100.0, kbin=0.1)
    assert np.max(np.abs(r.pl[0])) < 1e-9 * 100.0 ** 3


def test_single_precision_port_not_needed_marker(golden_outputs):
    # the single-precision goldens exist for the GPU float path
    assert set(golden_outputs["single"]) == set(SINGLE_CASES)


def test_ylm_restatement_matches_reference_closed_forms(ref_oracle, port_oracle):
    """The reference instantiates YlmR_l1..6 as external symbols
    (src/multipole.c:518-556, math/spherical.h:54-260)."""
    rng = np.random.default_rng(0)
    port = port_oracle.lib.port_ylm_real
    port.restype = C.c_double
    port.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4
    for l in range(1, 7):
        f = getattr(ref_oracle.lib, f"YlmR_l{l}")
        f.restype = C.c_double
        f.argtypes = [C.c_int] + [C.c_double] * 4
        for m in range(-l, l + 1):
            for _ in range(20):
                th, ph = rng.uniform(0, np.pi), rng.uniform(0, 2 * np.pi)
                a = (np.cos(th), np.sin(th), np.cos(ph), np.sin(ph))
                assert abs(f(m, *a) - port(l, m, *a)) < 5e-15


def test_cpu_fft_against_numpy(port_oracle):
    lib = port_oracle.lib
    lib.fftcpu_d_plan3d_create.restype = C.c_void_p
    lib.fftcpu_d_plan3d_create.argtypes = [C.c_int] * 3
    lib.fftcpu_d_r2c_3d.argtypes = [C.c_void_p] * 3
    lib.fftcpu_d_c2r_3d.argtypes = [C.c_void_p] * 3
    rng = np.random.default_rng(3)
    for shape in [(16, 16, 16), (15, 15, 15), (12, 10, 14), (24, 24, 24), (7, 22, 9)]:
        a = rng.standard_normal(shape)
        out = np.zeros(shape[:2] + (shape[2] // 2 + 1, 2))
        p = lib.fftcpu_d_plan3d_create(*shape)
        lib.fftcpu_d_r2c_3d(p, a.ctypes.data, out.ctypes.data)
        want = np.fft.rfftn(a)
        got = out[..., 0] + 1j * out[..., 1]
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
        back = np.zeros(shape)
        lib.fftcpu_d_c2r_3d(p, out.ctypes.data, back.ctypes.data)
        assert np.abs(back / a.size - a).max() < 1e-12


@pytest.mark.parametrize("seed", range(__import__("tests.fuzz_cases", fromlist=["NCASES"]).NCASES))
def test_port_matches_reference_on_random_configurations(seed, port_oracle, ref_oracle):
    """Seeded random simulation-box configurations (tests/fuzz_cases.py): the
    restatement against the unmodified reference — what licenses the restatement as
    the checker of the same cases on the GPU (tests/test_gpu_fuzz.py)."""
    from tests.fuzz_cases import fuzz_case
    cats, kw = fuzz_case(seed)
    data = cats if len(cats) > 1 else cats[0]
    want = ref_oracle.run(data, **kw)
    got = port_oracle.run(data, **kw)
    assert want.nbin > 0
    from tests.parity import noise_floor
    assert_spectra_close(got, want, 1e-10, f"fuzz {seed}: {kw}", abs_floor=noise_floor(want, kw["poles"]))


@pytest.mark.parametrize("seed", range(__import__("tests.fuzz_cases", fromlist=["NSURVEY"]).NSURVEY))
def test_port_matches_reference_on_random_surveys(seed, port_oracle, ref_oracle):
    """Seeded random survey configurations (data + randoms, FKP weights, automatic or
    given box, any scheme / interlacing / multipoles, auto and cross).  The reference
    runs on one thread (data race in its get_coord_bound, see above)."""
    from tests.fuzz_cases import survey_case
    data, kw = survey_case(seed)
    data = data if len(data) > 1 else data[0]
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(1)
    try:
        want = ref_oracle.run(data, **kw)
    finally:
        gomp.omp_set_num_threads(4)
    got = port_oracle.run(data, **kw)
    assert want.nbin > 0
    assert_spectra_close(got, want, 1e-9, f"survey fuzz {seed}")


@pytest.mark.parametrize("seed", range(__import__("tests.fuzz_cases", fromlist=["NMEDIUM"]).NMEDIUM))
def test_port_matches_reference_on_medium_random_configurations(seed, port_oracle, ref_oracle):
    """The random box configurations at sizes that engage the GPU's particle sort
    (tests/test_gpu_fuzz.py::test_medium_random_configuration_against_oracle)."""
    from tests.fuzz_cases import medium_case
    from tests.parity import noise_floor
    cats, kw = medium_case(seed)
    data = cats if len(cats) > 1 else cats[0]
    want = ref_oracle.run(data, **kw)
    got = port_oracle.run(data, **kw)
    assert_spectra_close(got, want, 1e-10, f"medium fuzz {seed}: {kw}", abs_floor=noise_floor(want, kw["poles"]))
