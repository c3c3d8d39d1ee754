"""The seam, end to end: the reference's UNMODIFIED host program (main,
load_conf, read_cata, cnvt_coord, save_res — compiled from /root/reference by
`make -C oracle dropin`) linked against libpowspec_b200.so (POWSPEC_b200) must
write the same output files as the same host linked with the reference's own
genr_mesh.o + multipole.o (POWSPEC_ref): identical header and format, identical
mode counts, P_ell within 1e-6.  Same powspec.conf form, same ASCII catalogues."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "POWSPEC_ref")
OUR_BIN = os.path.join(ROOT, "oracle", "_ref", "POWSPEC_b200")
OUR_BIN_CNVT = os.path.join(ROOT, "oracle", "_ref", "POWSPEC_b200_cnvt")


def _need_binaries():
    if not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)):
        pytest.skip("oracle/_ref/POWSPEC_{ref,b200} not built (needs /root/reference at build time)")


def _run(binary, conf, extra, cwd, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    p = subprocess.run([binary, "-c", conf, *extra], cwd=cwd, env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def _compare(fa, fb, ncols_int=(4,)):
    la, lb = open(fa).read().splitlines(), open(fb).read().splitlines()
    assert len(la) == len(lb)
    rows_a, rows_b = [], []
    for a, b in zip(la, lb):
        if a.startswith("#"):
            assert a == b, f"header differs:\n{a}\n{b}"
        else:
            rows_a.append([float(x) for x in a.split()])
            rows_b.append([float(x) for x in b.split()])
    A, B = np.array(rows_a), np.array(rows_b)
    assert A.shape == B.shape and A.shape[0] > 0
    assert np.array_equal(A[:, 4], B[:, 4]), "nmod differs"
    assert np.allclose(A[:, :4], B[:, :4], rtol=1e-9, atol=0)
    floor = 1e-3 * np.abs(A[:, 5:]).max()
    err = np.abs(A[:, 5:] - B[:, 5:]) / np.maximum(np.abs(A[:, 5:]), floor)
    # the files carry 10 significant digits (OFMT_DBL, src/define.h:107)
    assert err.max() < 1e-6, err.max()


def test_sim_auto_and_cross(tmp_path):
    _need_binaries()
    rng = np.random.default_rng(21)
    for tag, n in (("a", 4000), ("b", 3000)):
        np.savetxt(tmp_path / f"cat_{tag}.txt", np.c_[rng.random((n, 3)) * 200.0, rng.uniform(0.5, 2, n)],
                   fmt="%.17g")
    (tmp_path / "sim.conf").write_text("""
DATA_CATALOG = [cat_a.txt, cat_b.txt]
DATA_FORMATTER = ["%lf %lf %lf %lf", "%lf %lf %lf %lf"]
DATA_POSITION = [$1,$2,$3,$1,$2,$3]
DATA_WT_COMP = [$4, $4]
CUBIC_SIM = T
LINE_OF_SIGHT = [0,0,1]
BOX_SIZE = 200
GRID_SIZE = 32
PARTICLE_ASSIGN = 3
GRID_INTERLACE = T
MULTIPOLE = [0,2,4]
KMIN = 0
BIN_SIZE = 0.05
OVERWRITE = 1
VERBOSE = F
""")
    _run(REF_BIN, "sim.conf", ["-a", "[ref_a.txt,ref_b.txt]", "-x", "ref_x.txt"], tmp_path, 4)
    os.environ["POWSPEC_B200_TIMING"] = str(tmp_path / "timing.jsonl")
    try:
        out = _run(OUR_BIN, "sim.conf", ["-a", "[our_a.txt,our_b.txt]", "-x", "our_x.txt"], tmp_path, 4)
    finally:
        del os.environ["POWSPEC_B200_TIMING"]
    assert "Generating meshes for FFT" in out and "Evaluating power spectra" in out
    import json
    rec = json.loads((tmp_path / "timing.jsonl").read_text().splitlines()[-1])
    assert rec["grid"] == 32 and rec["launches"] > 0 and rec["stages_ms"]["assign"] > 0
    for t in ("a", "b", "x"):
        _compare(tmp_path / f"ref_{t}.txt", tmp_path / f"our_{t}.txt")


@pytest.mark.parametrize("devices", ["0,0", "0-1", "0,0,0,0"])
def test_sim_over_several_devices(tmp_path, devices):
    """POWSPEC_B200_DEVICES: the reference's single-process host drives ONE mesh
    slab-decomposed over several (here also virtual: a device listed twice) ranks through
    the same genr_mesh() / powspec() seam — route, halo, distributed FFT with peer stores,
    reduction inside the library (csrc/dist.cu psb_group).  Same files as the
    all-reference binary."""
    _need_binaries()
    import torch
    ranks = []
    for part in devices.split(","):
        a, _, b = part.partition("-")
        ranks += list(range(int(a), int(b or a) + 1))
    if torch.cuda.device_count() < 1 + max(ranks):
        pytest.skip(f"needs {1 + max(ranks)} GPUs")
    rng = np.random.default_rng(33)
    for tag, n in (("a", 5000), ("b", 3500)):
        np.savetxt(tmp_path / f"cat_{tag}.txt", np.c_[rng.random((n, 3)) * 300.0, rng.uniform(0.5, 2, n)],
                   fmt="%.17g")
    (tmp_path / "sim.conf").write_text("""
DATA_CATALOG = [cat_a.txt, cat_b.txt]
DATA_FORMATTER = ["%lf %lf %lf %lf", "%lf %lf %lf %lf"]
DATA_POSITION = [$1,$2,$3,$1,$2,$3]
DATA_WT_COMP = [$4, $4]
CUBIC_SIM = T
LINE_OF_SIGHT = [0,0,1]
BOX_SIZE = 300
GRID_SIZE = 48
PARTICLE_ASSIGN = 2
GRID_INTERLACE = T
MULTIPOLE = [0,2,4]
KMIN = 0
BIN_SIZE = 0.04
OVERWRITE = 1
VERBOSE = F
""")
    _run(REF_BIN, "sim.conf", ["-a", "[ref_a.txt,ref_b.txt]", "-x", "ref_x.txt"], tmp_path, 4)
    env_add = {"POWSPEC_B200_DEVICES": devices, "POWSPEC_B200_TIMING": str(tmp_path / "timing.jsonl")}
    os.environ.update(env_add)
    try:
        out = _run(OUR_BIN, "sim.conf", ["-a", "[our_a.txt,our_b.txt]", "-x", "our_x.txt"], tmp_path, 4)
    finally:
        for k in env_add:
            del os.environ[k]
    assert "Generating meshes for FFT" in out and "Evaluating power spectra" in out
    import json
    rec = json.loads((tmp_path / "timing.jsonl").read_text().splitlines()[-1])
    assert rec["grid"] == 48 and rec["devices"] == len(ranks)
    assert rec["stages_ms"]["assign"] > 0 and rec["stages_ms"]["fft_x"] > 0
    for t in ("a", "b", "x"):
        _compare(tmp_path / f"ref_{t}.txt", tmp_path / f"our_{t}.txt")


def test_survey_with_randoms_and_fkp(tmp_path):
    _need_binaries()

    def cat(seed, n):
        r = np.random.default_rng(seed)
        ra, dec = np.deg2rad(r.uniform(100, 160, n)), np.deg2rad(r.uniform(-10, 40, n))
        d = r.uniform(800, 1500, n)
        nz = np.full(n, 3e-4)
        return np.c_[d * np.cos(dec) * np.cos(ra), d * np.cos(dec) * np.sin(ra), d * np.sin(dec),
                     r.uniform(0.8, 1.2, n), 1 / (1 + 1e4 * nz), nz]
    np.savetxt(tmp_path / "data.txt", cat(1, 3000), fmt="%.17g")
    np.savetxt(tmp_path / "rand.txt", cat(2, 15000), fmt="%.17g")
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
GRID_SIZE = 32
PARTICLE_ASSIGN = 2
GRID_INTERLACE = T
MULTIPOLE = [0,2,4]
KMIN = 0
BIN_SIZE = 0.01
OVERWRITE = 1
VERBOSE = F
""")
    # one thread for the reference: its get_coord_bound races (see tests/test_oracle.py)
    _run(REF_BIN, "survey.conf", ["-a", "ref.txt"], tmp_path, 1)
    _run(OUR_BIN, "survey.conf", ["-a", "our.txt"], tmp_path, 4)
    _compare(tmp_path / "ref.txt", tmp_path / "our.txt")


def test_survey_from_sky_coordinates(tmp_path):
    """DATA_CONVERT / RAND_CONVERT: (RA, Dec, z) catalogues.  POWSPEC_b200 keeps the
    host's cnvt_coord (CPU) in front of the device path; POWSPEC_b200_cnvt is linked
    without cnvt_coord.o, so the library's cnvt_coord is bound and the conversion
    runs on the device inside genr_mesh.  Both must reproduce the all-reference
    binary's output file."""
    _need_binaries()
    if not os.path.exists(OUR_BIN_CNVT):
        pytest.skip("oracle/_ref/POWSPEC_b200_cnvt not built")

    def cat(seed, n):
        r = np.random.default_rng(seed)
        ra, dec = r.uniform(100, 160, n), r.uniform(-10, 40, n)
        z = r.uniform(0.3, 0.7, n)
        nz = np.full(n, 3e-4)
        return np.c_[ra, dec, z, r.uniform(0.8, 1.2, n), 1 / (1 + 1e4 * nz), nz]
    np.savetxt(tmp_path / "data.txt", cat(5, 3000), fmt="%.17g")
    np.savetxt(tmp_path / "rand.txt", cat(6, 15000), fmt="%.17g")
    (tmp_path / "sky.conf").write_text("""
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
DATA_CONVERT = T
RAND_CONVERT = T
OMEGA_M = 0.31
CMVDST_ERR = 1e-8
CUBIC_SIM = F
GRID_SIZE = 32
PARTICLE_ASSIGN = 2
GRID_INTERLACE = F
MULTIPOLE = [0,2]
KMIN = 0
BIN_SIZE = 0.01
OVERWRITE = 1
VERBOSE = F
""")
    _run(REF_BIN, "sky.conf", ["-a", "ref.txt"], tmp_path, 1)
    _run(OUR_BIN, "sky.conf", ["-a", "our.txt"], tmp_path, 4)
    out = _run(OUR_BIN_CNVT, "sky.conf", ["-a", "our_dev.txt"], tmp_path, 4)
    assert "Converting coordinates" in out
    _compare(tmp_path / "ref.txt", tmp_path / "our.txt")
    _compare(tmp_path / "ref.txt", tmp_path / "our_dev.txt")

    # error contract of the seam's cnvt_coord (src/cnvt_coord.c:185-268, 549-582): a negative
    # redshift is reported by the conversion step itself — message on stderr, the host's
    # [FAIL] path, non-zero exit — not later by genr_mesh
    bad = cat(5, 3000)
    bad[17, 2] = -0.25
    np.savetxt(tmp_path / "data.txt", bad, fmt="%.17g")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for binary in (REF_BIN, OUR_BIN_CNVT):
        p = subprocess.run([binary, "-c", "sky.conf", "-a", "bad.txt"], cwd=tmp_path, env=env,
                           capture_output=True, text=True)
        assert p.returncode != 0
        assert "invalid negative redshift in the data catalog" in p.stderr
        assert "Generating meshes for FFT" not in p.stdout
