"""The Python mirror of save_res (powspec_b200/save_res.py) against the files the
reference's own binary writes (oracle/_ref/POWSPEC_ref = unmodified host + hot
path, CPU only): byte for byte, headers included.  The spectra handed to the
mirror come from the same unmodified reference code run in memory
(oracle/_ref/libpowspec_ref.so), so any difference is the writer's."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "POWSPEC_ref")


def _ref_binary(conf, extra, cwd):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/POWSPEC_ref not built (needs /root/reference at build time)")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([REF_BIN, "-c", conf, *extra], cwd=cwd, env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def _structures(res, conf, datas, rands=None, scalars=None):
    """api.Conf / Cata / Mesh / PK filled from an oracle result."""
    import powspec_b200 as pb
    cata = pb.Cata(data=datas, rand=rands,
                   wdata=[s["wdata"] for s in scalars] if scalars else None,
                   wrand=[s["wrand"] for s in scalars] if scalars else None)
    mesh = pb.Mesh(ctx=None, Ng=conf.gsize, min=res.bmin, max=res.bmin + res.bsize, bsize=res.bsize,
                   issim=conf.issim, intlace=conf.intlace, assign=conf.assign, num=len(datas))
    pk = pb.PK(nl=res.nl, nbin=res.nbin, poles=list(conf.poles), k=res.k, kedge=res.kedge, km=res.km,
               cnt=res.cnt, lcnt=res.lcnt, pl=res.pl, xpl=res.xpl, shot=res.shot, norm=res.norm)
    return cata, mesh, pk


def _one_thread(fn):
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(1)
    try:
        return fn()
    finally:
        gomp.omp_set_num_threads(4)


def test_sim_auto_and_cross_files_match_the_reference_writer(tmp_path, ref_oracle):
    import powspec_b200 as pb
    rng = np.random.default_rng(31)
    cats = [np.c_[rng.random((n, 3)) * 200.0, rng.uniform(0.5, 2, n)] for n in (4000, 3000)]
    for tag, c in zip("ab", cats):
        np.savetxt(tmp_path / f"cat_{tag}.txt", c, fmt="%.17g")
    (tmp_path / "sim.conf").write_text("""
DATA_CATALOG = [cat_a.txt, cat_b.txt]
DATA_FORMATTER = ["%lf %lf %lf %lf", "%lf %lf %lf %lf"]
DATA_POSITION = [$1,$2,$3,$1,$2,$3]
DATA_WT_COMP = [$4, $4]
CUBIC_SIM = T
LINE_OF_SIGHT = [0,0,1]
BOX_SIZE = 200
GRID_SIZE = 24
PARTICLE_ASSIGN = 2
GRID_INTERLACE = T
MULTIPOLE = [0,1,2,4]
KMIN = 0
BIN_SIZE = 0.05
OVERWRITE = 1
VERBOSE = F
""")
    _ref_binary("sim.conf", ["-a", "[ref_a.txt,ref_b.txt]", "-x", "ref_x.txt"], tmp_path)
    conf = pb.Conf(ndata=2, issim=True, bsize=(200.0,) * 3, gsize=24, assign=2, intlace=True,
                   poles=(0, 1, 2, 4), kbin=0.05, isauto=(True, True), iscross=True,
                   oauto=(str(tmp_path / "our_a.txt"), str(tmp_path / "our_b.txt")),
                   ocross=str(tmp_path / "our_x.txt"))
    res = _one_thread(lambda: ref_oracle.run(cats, ng=24, assign="TSC", interlace=True, poles=(0, 1, 2, 4),
                                             box=200.0, kbin=0.05))
    cata, mesh, pk = _structures(res, conf, cats)
    pb.save_res(conf, cata, mesh, pk)
    for t in "abx":
        assert (tmp_path / f"our_{t}.txt").read_bytes() == (tmp_path / f"ref_{t}.txt").read_bytes(), t
    # without the header
    conf.oheader = False
    conf.iscross = False
    pb.save_res(conf, cata, mesh, pk)
    body = [ln for ln in (tmp_path / "ref_a.txt").read_text().splitlines(True)
            if not ln.startswith("#") or ln.startswith("# kcen")]
    assert (tmp_path / "our_a.txt").read_text() == "".join(body)


def test_survey_file_matches_the_reference_writer(tmp_path, ref_oracle):
    import powspec_b200 as pb
    from oracle.oracle import survey_scalars

    def cat(seed, n):
        r = np.random.default_rng(seed)
        ra, dec = np.deg2rad(r.uniform(100, 160, n)), np.deg2rad(r.uniform(-10, 40, n))
        d = r.uniform(800, 1500, n)
        nz = np.full(n, 3e-4)
        return np.c_[d * np.cos(dec) * np.cos(ra), d * np.cos(dec) * np.sin(ra), d * np.sin(dec),
                     r.uniform(0.8, 1.2, n), 1 / (1 + 1e4 * nz), nz]
    d, q = cat(41, 2500), cat(42, 12000)
    np.savetxt(tmp_path / "data.txt", d, fmt="%.17g")
    np.savetxt(tmp_path / "rand.txt", q, fmt="%.17g")
    (tmp_path / "survey.conf").write_text("""
DATA_CATALOG = data.txt
RAND_CATALOG = rand.txt
DATA_FORMATTER = "%lf %lf %lf %lf %lf %lf"
RAND_FORMATTER = "%lf %lf %lf %lf %lf %lf"
DATA_POSITION = [$1,$2,$3]
RAND_POSITION = [$1,$2,$3]
DATA_WT_COMP = $4
RAND_WT_COMP = $4
DATA_WT_FKP = $5
RAND_WT_FKP = $5
DATA_NZ = $6
RAND_NZ = $6
CUBIC_SIM = F
GRID_SIZE = 24
PARTICLE_ASSIGN = 1
GRID_INTERLACE = F
MULTIPOLE = [0,2]
KMIN = 0
BIN_SIZE = 0.01
OVERWRITE = 1
VERBOSE = F
""")
    _ref_binary("survey.conf", ["-a", "ref.txt"], tmp_path)
    sc = survey_scalars(d[:, 3], d[:, 4], d[:, 5], q[:, 3], q[:, 4], q[:, 5])
    datas = [np.ascontiguousarray(np.c_[d[:, :3], d[:, 3] * d[:, 4]])]
    rands = [np.ascontiguousarray(np.c_[q[:, :3], q[:, 3] * q[:, 4]])]
    res = _one_thread(lambda: ref_oracle.run(datas[0], ng=24, assign="CIC", interlace=False, poles=(0, 2),
                                             issim=False, rand=rands, scalars=[sc], kbin=0.01))
    conf = pb.Conf(ndata=1, issim=False, gsize=24, assign=1, intlace=False, poles=(0, 2), kbin=0.01,
                   isauto=(True, False), iscross=False, oauto=(str(tmp_path / "our.txt"),))
    cata, mesh, pk = _structures(res, conf, datas, rands, [sc])
    pb.save_res(conf, cata, mesh, pk)
    assert (tmp_path / "our.txt").read_bytes() == (tmp_path / "ref.txt").read_bytes()


def test_unwritable_output_is_an_error(tmp_path):
    import powspec_b200 as pb
    conf = pb.Conf(ndata=1, oauto=(str(tmp_path / "no" / "such" / "dir" / "pk.txt"),))
    cata = pb.Cata(data=[np.ones((3, 4))])
    mesh = pb.Mesh(ctx=None, Ng=4, min=np.zeros(3), max=np.ones(3), bsize=np.ones(3), issim=True,
                   intlace=False, assign=2, num=1)
    z = np.zeros(1)
    pk = pb.PK(nl=3, nbin=1, poles=[0, 2, 4], k=z, kedge=np.zeros(2), km=z, cnt=np.zeros(1, np.uint64),
               lcnt=np.zeros((3, 1)), pl=[np.zeros((3, 1)), None], xpl=None, shot=np.ones(2), norm=np.ones(2))
    with pytest.raises(pb.PowspecB200Error, match="cannot write"):
        pb.save_res(conf, cata, mesh, pk)
