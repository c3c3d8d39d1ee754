"""Seeded random simulation-box configurations for the parity tests: every scheme,
interlacing on/off, even and odd mesh sizes, any subset of the multipoles 0..6,
arbitrary unit line of sight, cubic and non-cubic boxes, linear and logarithmic
bins with and without KMIN / KMAX, weighted and unweighted particles, auto and
cross spectra.  The same cases are run (a) on CPU: restatement against the
unmodified reference library, (b) on the GPU: CUDA path against the restatement.

Only values the reference's own configuration check accepts are drawn
(src/load_conf.c:990-1003: LINE_OF_SIGHT is a unit vector; :1352-1355: KMIN / KMAX
are handed over as log10 for logarithmic bins)."""
import numpy as np

NCASES = 24
SCHEMES = ("NGP", "CIC", "TSC", "PCS")


NMEDIUM = 8


def medium_case(seed):
    """The same draw at sizes where the particle sort and the strip-ordered,
    z-coalesced scatter are engaged (>= 65536 particles): meshes of 33..80 cells per
    side (odd sizes, sizes that are not a multiple of the strip height)."""
    cats, kw = fuzz_case(seed, ng_range=(33, 81), n_range=(70_000, 250_000), base=3000)
    kw["interlace"] = seed % 2 == 0
    return cats, kw


def fuzz_case(seed, ng_range=(9, 41), n_range=(300, 4000), base=1000):
    """-> (list of catalogues, keyword arguments of oracle.run / powspec_b200.run)"""
    r = np.random.default_rng(base + seed)
    ng = int(r.integers(*ng_range))
    cubic = r.random() < 0.6
    box = float(r.uniform(80, 400))
    bsize = (box,) * 3 if cubic else tuple(float(box * f) for f in r.uniform(0.8, 1.3, 3))
    ncat = 2 if r.random() < 0.3 else 1
    cats = []
    for _ in range(ncat):
        n = int(r.integers(*n_range))
        xyz = r.random((n, 3)) * np.asarray(bsize)
        # some particles on / next to the faces and cell boundaries (periodic wraps)
        m = min(n, 12)
        cells = r.integers(0, ng, (m, 3)) + r.choice([0.0, 0.5, 1 - 1e-12, 1e-12], (m, 3))
        xyz[:m] = np.minimum(cells / ng, 1 - 1e-15) * np.asarray(bsize)
        w = r.uniform(0.2, 2.0, n) if r.random() < 0.5 else np.ones(n)
        cats.append(np.ascontiguousarray(np.c_[xyz, w]))
    npole = int(r.integers(1, 8))
    poles = tuple(sorted(int(p) for p in r.choice(7, npole, replace=False)))
    if r.random() < 0.5:
        los = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)][int(r.integers(3))]
    else:
        v = r.normal(size=3)
        v /= np.sqrt((v * v).sum())
        los = tuple(float(x) for x in v)
    kf = 2 * np.pi / max(bsize)                     # fundamental of the longest side
    kny = np.pi * ng / max(bsize)
    kw = dict(ng=ng, assign=SCHEMES[int(r.integers(4))], interlace=bool(r.random() < 0.5),
              poles=poles, box=bsize if not cubic else box, los=los)
    if r.random() < 0.25:
        kmin = float(r.uniform(0.5, 2.0) * kf)
        kmax = float(r.uniform(0.5, 0.95) * kny)
        kw.update(logscale=True, kmin=float(np.log10(kmin)), kmax=float(np.log10(kmax)),
                  kbin=float(r.uniform(0.05, 0.15)))
    else:
        kw.update(kbin=float(r.uniform(1.0, 4.0) * kf))
        if r.random() < 0.4:
            kw.update(kmin=float(r.uniform(0.0, 2.0) * kf))
        if r.random() < 0.4:
            kw.update(kmax=float(r.uniform(0.5, 1.2) * kny))
    if ncat == 2:
        which = int(r.integers(3))
        kw.update(isauto=[[True, True], [True, False], [False, False]][which], iscross=True)
    return cats, kw


NSURVEY = 12


def survey_case(seed):
    """Survey-like configuration: data + randoms in a wedge of sky (already comoving
    Cartesian), completeness and FKP weights, automatic box with padding or a given
    one, any scheme, interlacing, any subset of the multipoles, one or two catalogues.
    -> (data list, keyword arguments incl. rand= and scalars=)"""
    from oracle.oracle import survey_scalars
    r = np.random.default_rng(5000 + seed)
    ncat = 2 if r.random() < 0.35 else 1
    ra0, dec0 = r.uniform(0, 300), r.uniform(-40, 20)
    dra, ddec = r.uniform(30, 60), r.uniform(20, 50)
    d0, d1 = r.uniform(500, 900), r.uniform(1200, 1800)

    def cat(n):
        ra = np.deg2rad(r.uniform(ra0, ra0 + dra, n))
        dec = np.deg2rad(r.uniform(dec0, dec0 + ddec, n))
        dist = r.uniform(d0, d1, n)
        nz = r.uniform(1e-4, 5e-4, n)
        wfkp = 1 / (1 + 1e4 * nz)
        wc = r.uniform(0.7, 1.3, n)
        xyz = dist[:, None] * np.c_[np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)]
        return np.ascontiguousarray(np.c_[xyz, wc * wfkp]), wc, wfkp, nz

    data, rand, scalars = [], [], []
    for _ in range(ncat):
        d, dwc, dwf, dnz = cat(int(r.integers(800, 3000)))
        q, qwc, qwf, qnz = cat(int(r.integers(5000, 12000)))
        data.append(d); rand.append(q)
        scalars.append(survey_scalars(dwc, dwf, dnz, qwc, qwf, qnz))
    npole = int(r.integers(1, 6))
    poles = tuple(sorted(int(p) for p in r.choice(7, npole, replace=False)))
    if r.random() < 0.8 and 0 not in poles:
        poles = (0,) + poles[1:] if len(poles) > 1 else (0,)
    ext = np.ptp(np.vstack(data + rand)[:, :3], axis=0)
    kw = dict(ng=int(r.integers(12, 29)), assign=SCHEMES[int(r.integers(4))],
              interlace=bool(r.random() < 0.5), poles=poles, issim=False, rand=rand, scalars=scalars)
    if r.random() < 0.3:
        kw["box"] = tuple(float(np.ceil(e * f)) for e, f in zip(ext, r.uniform(1.05, 1.4, 3)))
    else:
        kw["bpad"] = tuple(float(x) for x in r.uniform(0.0, 0.1, 3))
    side = max(kw["box"]) if "box" in kw else float(ext.max()) * 1.1
    kf = 2 * np.pi / side
    if r.random() < 0.25:
        kw.update(logscale=True, kmin=float(np.log10(r.uniform(0.8, 2.0) * kf)),
                  kbin=float(r.uniform(0.06, 0.15)))
    else:
        kw.update(kbin=float(r.uniform(1.0, 3.0) * kf))
        if r.random() < 0.4:
            kw.update(kmin=float(r.uniform(0.0, 1.5) * kf))
        if r.random() < 0.4:
            kw.update(kmax=float(r.uniform(0.5, 1.1) * np.pi * kw["ng"] / side))
    if ncat == 2:
        kw.update(isauto=[[True, True], [True, False], [False, False]][int(r.integers(3))], iscross=True)
    return data, kw
