"""GPU parity on seeded random configurations (tests/fuzz_cases.py): the CUDA
path through the C ABI against the CPU restatement, which the CPU suite checks
against the unmodified reference on exactly these cases
(tests/test_oracle.py::test_port_matches_reference_on_random_configurations).
Mode counts bit-exact, P_ell within 1e-6 (double) / 1e-4 (single)."""
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
        worst4 = assert_spectra_close(got4, want4, TOL_SINGLE, f"fuzz {seed} single: {kw}",
                                      abs_floor=noise_floor(want, kw["poles"]))
        print(f"fuzz {seed} single: worst rel err {worst4:.2e}")
