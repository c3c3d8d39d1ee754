"""Shared comparison helpers for the parity tests.

Tolerances (BASELINE.json north_star): mode counts bit-exact; P_ell(k) within
1e-6 relative in double, 1e-4 in single.  "Relative" is taken per spectrum
against max(|P_ell(k)|, shot-noise floor) — for Poisson-dominated catalogues
P_0 = raw - shot cancels to a small number and P_2/P_4 change sign, so a
per-point relative error is meaningless at zero crossings (hard
part 4), and odd multipoles of a periodic box are pure rounding noise of
cancelling +mu/-mu pairs (quirk Q6).  So the error of each point is measured
against max(|P_l(k)|, 1e-3 * max over all l,k of |P|) of that spectrum.
"""
import numpy as np

TOL_DOUBLE = 1e-6
TOL_SINGLE = 1e-4


def rel_err(got, want, floor=0.0):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = np.maximum(np.abs(want), floor)
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(got - want) / scale))


def assert_spectra_close(got, want, tol, what="", abs_floor=0.0):
    """got/want: objects with nbin, cnt, k, km, kedge, lcnt, pl (list), xpl, shot.
    `want` may be a golden dict.  abs_floor: an absolute scale below which
    differences do not count — for spectra that hold ONLY odd multipoles of a
    periodic box, whose own maximum is rounding noise (see noise_floor)."""
    if isinstance(want, dict):
        w = want
    else:
        w = dict(nbin=want.nbin, nl=want.nl, k=want.k, kedge=want.kedge, km=want.km,
                 cnt=want.cnt, lcnt=want.lcnt, pl=want.pl, xpl=want.xpl, shot=want.shot,
                 norm=want.norm)
    assert got.nbin == w["nbin"], what
    # mode counts: integers, bit-exact
    assert np.array_equal(np.asarray(got.cnt, dtype=np.uint64),
                          np.asarray(w["cnt"], dtype=np.uint64)), f"{what}: mode counts differ"
    assert rel_err(got.k, w["k"]) < 1e-13, what
    assert rel_err(got.kedge, w["kedge"]) < 1e-13, what
    # the reference sums |k| in thread-dependent order (src/multipole.c:119-170,243-248)
    assert rel_err(got.km, w["km"]) < 1e-10, f"{what}: kavg"
    assert rel_err(got.shot, w["shot"]) < 1e-13, what
    assert rel_err(got.norm, w["norm"]) < 1e-13, what
    worst = 0.0
    for i in range(2):
        wp = w["pl"][i]
        if wp is None:
            assert got.pl[i] is None, what
            continue
        wp = np.asarray(wp)
        assert got.pl[i] is not None, what
        floor = max(np.max(np.abs(wp)), abs_floor * 1e3)
        for l in range(wp.shape[0]):
            e = rel_err(got.pl[i][l], wp[l], floor * 1e-3)
            worst = max(worst, e)
            assert e < tol, f"{what}: pl[{i}][{l}] rel err {e:.3e} > {tol}"
    if w["xpl"] is not None:
        wx = np.asarray(w["xpl"])
        assert got.xpl is not None, what
        floor = max(np.max(np.abs(wx)), abs_floor * 1e3)
        for l in range(wx.shape[0]):
            e = rel_err(got.xpl[l], wx[l], floor * 1e-3)
            worst = max(worst, e)
            assert e < tol, f"{what}: xpl[{l}] rel err {e:.3e} > {tol}"
    else:
        assert got.xpl is None, what
    return worst


def noise_floor(want, poles):
    """abs_floor for assert_spectra_close: 0 unless every requested multipole is odd
    (simulation boxes).  Then all of P_l is the rounding residue of cancelling
    +mu / -mu pairs (~1e-16 of the shot noise, quirk Q6) and has no scale of its
    own: differences are measured against 1e-3 of the shot-noise level instead,
    the scale P_0 would have."""
    if any(p % 2 == 0 for p in poles):
        return 0.0
    shot = [s for s in np.asarray(want.shot, dtype=np.float64) if s > 0]
    return 1e-3 * float(np.sqrt(np.prod(shot)) if len(shot) == 2 else shot[0]) if shot else 0.0
