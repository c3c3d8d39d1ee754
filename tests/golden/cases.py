"""Definition of the golden parity cases (inputs are seeded, small, and also
stored in golden_inputs.npz so the fixtures do not depend on numpy's
bit-stream).  Shared by make_golden.py (generation, needs oracle/_ref built from
/root/reference) and by the tests (checking)."""
import numpy as np

L = 100.0


def _sim_cat(seed, n, weights=False, box=L):
    r = np.random.default_rng(seed)
    xyz = r.random((n, 3)) * np.asarray(box)
    w = r.uniform(0.5, 1.5, n) if weights else np.ones(n)
    return np.c_[xyz, w]


def _survey_cat(seed, n):
    """(x,y,z,w=wcomp*wfkp), wcomp, wfkp, nz — a wedge of sky, already in
    comoving Cartesian coordinates (coordinate conversion is host work that
    happens before the boundary, src/cnvt_coord.c)."""
    r = np.random.default_rng(seed)
    ra = np.deg2rad(r.uniform(100, 160, n))
    dec = np.deg2rad(r.uniform(-10, 40, n))
    dist = r.uniform(800, 1500, n)
    x = dist * np.cos(dec) * np.cos(ra)
    y = dist * np.cos(dec) * np.sin(ra)
    z = dist * np.sin(dec)
    nz = np.full(n, 3e-4)
    wfkp = 1 / (1 + 1e4 * nz)
    wc = r.uniform(0.8, 1.2, n)
    return np.c_[x, y, z, wc * wfkp], wc, wfkp, nz


def make_inputs():
    """name -> array.  Survey entries also carry the columns needed for the
    ingest scalars."""
    inp = {}
    # the catalogue of SURVEY.md §4 Golden-A/B
    xyz = np.random.default_rng(12345).random((2000, 3)) * 100
    inp["survey_md_A"] = np.c_[xyz, np.ones(2000)]
    inp["sim_w"] = _sim_cat(7, 3000, weights=True)
    # particles on / next to the box faces exercise the periodic wraps
    inp["sim_w"][0, :3] = [0.0, 0.0, L * (1 - 1e-9)]
    inp["sim_w"][1, :3] = [L - 0.01, 0.2, 99.7]
    inp["sim_w"][2, :3] = [L / 16 * 15.5, L / 16 * 0.5, L / 16 * 7.5]
    inp["sim_b"] = _sim_cat(8, 2500)
    inp["sim_nc"] = _sim_cat(9, 3000, weights=True, box=(100.0, 120.0, 110.0))
    inp["sim_256"] = _sim_cat(10, 1000, box=1000.0)
    for tag, seed, n in (("svD1", 1, 2000), ("svR1", 2, 10000),
                         ("svD2", 3, 1500), ("svR2", 4, 9000)):
        cat, wc, wfkp, nz = _survey_cat(seed, n)
        inp[tag] = cat
        inp[tag + "_cols"] = np.c_[wc, wfkp, nz]
    return inp


# Each case: name, catalogue names, and oracle.run keyword arguments.
CASES = [
    dict(name="goldenA_tsc_il", data=["survey_md_A"],
         kw=dict(ng=16, assign="TSC", interlace=True, poles=(0, 2, 4), box=L, kbin=0.1)),
    dict(name="goldenB_cic", data=["survey_md_A"],
         kw=dict(ng=16, assign="CIC", interlace=False, poles=(0, 2), box=L, kbin=0.1)),
    dict(name="sim_ngp", data=["sim_w"],
         kw=dict(ng=16, assign="NGP", interlace=False, poles=(0, 2, 4), box=L, kbin=0.1)),
    dict(name="sim_ngp_il", data=["sim_w"],
         kw=dict(ng=16, assign="NGP", interlace=True, poles=(0, 2, 4), box=L, kbin=0.1)),
    dict(name="sim_cic_il_odd", data=["sim_w"],
         kw=dict(ng=15, assign="CIC", interlace=True, poles=(0, 1, 2, 3, 4, 5, 6),
                 box=L, kbin=0.1, los=(0.6, 0.0, 0.8))),
    dict(name="sim_tsc", data=["sim_w"],
         kw=dict(ng=16, assign="TSC", interlace=False, poles=(0, 2, 4), box=L, kbin=0.1)),
    dict(name="sim_pcs", data=["sim_w"],
         kw=dict(ng=16, assign="PCS", interlace=False, poles=(0, 2, 4), box=L, kbin=0.1)),
    dict(name="sim_pcs_il_allpoles", data=["sim_w"],
         kw=dict(ng=18, assign="PCS", interlace=True, poles=(0, 1, 2, 3, 4, 5, 6),
                 box=L, kbin=0.08, los=(0.0, 1.0, 0.0))),
    dict(name="sim_tsc_il_log_noncubic", data=["sim_nc"],
         kw=dict(ng=18, assign="TSC", interlace=True, poles=(0, 2), box=(100.0, 120.0, 110.0),
                 kbin=0.05, logscale=True, kmin=float(np.log10(0.08)),
                 kmax=float(np.log10(0.5)))),
    dict(name="sim_cross_pcs_il", data=["sim_w", "sim_b"],
         kw=dict(ng=16, assign="PCS", interlace=True, poles=(0, 2, 4), box=L, kbin=0.07,
                 kmin=0.03, kmax=0.4)),
    dict(name="sim_cross_only_cic", data=["sim_w", "sim_b"],
         kw=dict(ng=16, assign="CIC", interlace=False, poles=(0, 1, 2), box=L, kbin=0.07,
                 isauto=[True, False], iscross=True)),
    dict(name="sim_counts_256", data=["sim_256"],
         kw=dict(ng=256, assign="CIC", interlace=False, poles=(0, 2), box=1000.0, kbin=0.01)),
    dict(name="survey_tsc", data=["svD1"], rand=["svR1"],
         kw=dict(ng=24, assign="TSC", interlace=False, poles=(0, 2, 4), issim=False, kbin=0.01)),
    dict(name="survey_pcs_il_allpoles", data=["svD1"], rand=["svR1"],
         kw=dict(ng=20, assign="PCS", interlace=True, poles=(0, 1, 2, 3, 4, 5, 6),
                 issim=False, kbin=0.01)),
    dict(name="survey_cic_il_box", data=["svD1"], rand=["svR1"],
         kw=dict(ng=24, assign="CIC", interlace=True, poles=(0, 2), issim=False, kbin=0.01,
                 box=(1700.0, 1600.0, 1500.0))),
    dict(name="survey_ngp", data=["svD1"], rand=["svR1"],
         kw=dict(ng=24, assign="NGP", interlace=False, poles=(0, 2), issim=False, kbin=0.01)),
    dict(name="survey_cross_tsc_il", data=["svD1", "svD2"], rand=["svR1", "svR2"],
         kw=dict(ng=20, assign="TSC", interlace=True, poles=(0, 2, 4), issim=False, kbin=0.01)),
    dict(name="survey_log", data=["svD1"], rand=["svR1"],
         kw=dict(ng=20, assign="TSC", interlace=False, poles=(0, 2), issim=False, kbin=0.1,
                 logscale=True, kmin=-2.0)),
]

# cases also generated with the reference built with -DSINGLE_PREC
SINGLE_CASES = ["goldenA_tsc_il", "sim_pcs", "sim_cross_pcs_il", "survey_tsc"]


def survey_scalars_for(inp, dname, rname):
    from oracle.oracle import survey_scalars
    d, r = inp[dname + "_cols"], inp[rname + "_cols"]
    return survey_scalars(d[:, 0], d[:, 1], d[:, 2], r[:, 0], r[:, 1], r[:, 2])


def run_case(oracle, case, inp, **extra):
    data = [inp[n] for n in case["data"]]
    kw = dict(case["kw"])
    kw.update(extra)
    if "rand" in case:
        kw["rand"] = [inp[n] for n in case["rand"]]
        kw["scalars"] = [survey_scalars_for(inp, d, r)
                         for d, r in zip(case["data"], case["rand"])]
    return oracle.run(data if len(data) > 1 else data[0], **kw)
