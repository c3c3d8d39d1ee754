"""GPU parity on seeded random configurations (tests/fuzz_cases.py): the CUDA
path through the C ABI against the CPU restatement, which the CPU suite checks
against the unmodified reference on exactly these cases
(tests/test_oracle.py::test_port_matches_reference_on_random_configurations).
Mode counts bit-exact, P_ell within 1e-6 (double) / 1e-4 (single)."""
import numpy as np
import pytest

from tests.fuzz_cases import NCASES, fuzz_case
from tests.parity import TOL_DOUBLE, TOL_SINGLE, assert_spectra_close, noise_floor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import powspec_b200
    c = powspec_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("seed", range(NCASES))
def test_random_configuration_against_oracle(seed, ctx, port_oracle):
    import powspec_b200
    cats, kw = fuzz_case(seed)
    data = cats if len(cats) > 1 else cats[0]
    want = port_oracle.run(data, **kw)
    got = powspec_b200.run(data, ctx=ctx, **kw)
    worst = assert_spectra_close(got, want, TOL_DOUBLE, f"fuzz {seed}: {kw}",
                                 abs_floor=noise_floor(want, kw["poles"]))
    print(f"fuzz {seed}: worst rel err {worst:.2e}")
    assert got.launches > 0
    if seed % 4 == 0:
        # every fourth case also with a single-precision mesh, against the reference's
        # own -DSINGLE_PREC build where it was prebuilt (oracle/_ref travels to the GPU
        # box), else against the double-precision restatement
        from oracle import have_ref, load_oracle
        want4 = load_oracle("ref", single=True).run(data, **kw) if have_ref(single=True) else want
        got4 = powspec_b200.run(data, ctx=ctx, precision=4, **kw)
        # float meshes: values below 1e-2 of the spectrum's maximum are within ~100 float
        # roundings (6e-8 each) of 1e-4 relative, so the error is measured against
        # max(|P|, 1e-2 max|P|) there instead of the 1e-3 of the double-precision checks
        top = max(float(np.abs(np.asarray(p)).max()) for p in (*want4.pl, want4.xpl) if p is not None)
        worst4 = assert_spectra_close(got4, want4, TOL_SINGLE, f"fuzz {seed} single: {kw}",
                                      abs_floor=max(noise_floor(want, kw["poles"]), 1e-2 * top))
        print(f"fuzz {seed} single: worst rel err {worst4:.2e}")


@pytest.mark.parametrize("seed", range(__import__("tests.fuzz_cases", fromlist=["NSURVEY"]).NSURVEY))
def test_random_survey_against_oracle(seed, ctx, port_oracle):
    import powspec_b200
    from tests.fuzz_cases import survey_case
    data, kw = survey_case(seed)
    data = data if len(data) > 1 else data[0]
    want = port_oracle.run(data, **kw)
    for direct in (1, 0):       # l > 0 binned directly per m / through the Fkl field
        ctx.set_option("survey_direct", direct)
        try:
            got = powspec_b200.run(data, ctx=ctx, **kw)
        finally:
            ctx.set_option("survey_direct", 1)
        worst = assert_spectra_close(got, want, TOL_DOUBLE, f"survey fuzz {seed} direct={direct}")
        print(f"survey fuzz {seed} direct={direct}: worst rel err {worst:.2e}")


@pytest.mark.parametrize("seed", range(__import__("tests.fuzz_cases", fromlist=["NMEDIUM"]).NMEDIUM))
def test_medium_random_configuration_against_oracle(seed, ctx, port_oracle):
    """Random box configurations with 7e4..2.5e5 particles on 33..80 cells per side:
    the counting sort, the strip order (mesh sizes that are not a multiple of the
    strip height, odd sizes) and the z-coalesced scatter are engaged.  The density
    meshes are compared cell by cell as well as the spectra."""
    import powspec_b200
    from tests.fuzz_cases import medium_case
    cats, kw = medium_case(seed)
    data = cats if len(cats) > 1 else cats[0]
    want = port_oracle.run(data, keep_mesh=True, **kw)
    got = powspec_b200.run(data, ctx=ctx, keep_mesh=True, **kw)
    # dense catalogues take the owner-computes assignment, whose fixed-point contributions
    # are rounded to ~1e-12 of the cell values (csrc/assign_tiles.cu)
    mesh_tol = 1e-12 if ctx.L.psb_assign_path(ctx.h) == 0 else 1e-10
    for i in range(len(cats)):
        scale = np.abs(want.Fr[i]).max()
        assert np.abs(got.Fr[i] - want.Fr[i]).max() < mesh_tol * scale, f"medium fuzz {seed}: mesh {i}"
        if kw["interlace"]:
            assert np.abs(got.Frl[i] - want.Frl[i]).max() < mesh_tol * scale, f"medium fuzz {seed}: shifted mesh {i}"
    worst = assert_spectra_close(got, want, TOL_DOUBLE, f"medium fuzz {seed}: {kw}",
                                 abs_floor=noise_floor(want, kw["poles"]))
    print(f"medium fuzz {seed}: worst rel err {worst:.2e}")
