"""TEST INFRASTRUCTURE — ctypes front-end of the two oracle libraries.

* ``oracle/_ref/libpowspec_ref.so`` (+ ``_f32``): the UNMODIFIED reference
  ``genr_mesh()`` + ``powspec()`` (/root/reference/src/genr_mesh.c:874,
  src/multipole.c:1179) behind ``oracle/ref_driver.c``; built by
  ``make -C oracle ref`` where /root/reference exists, shipped prebuilt to the
  GPU box.
* ``oracle/libpowspec_port.so``: the CPU restatement ``oracle/pspec_port.c``.

Both export the ABI of ``oracle/oracle_abi.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libpowspec_ref.so")
REF_LIB_F32 = os.path.join(HERE, "_ref", "libpowspec_ref_f32.so")
PORT_LIB = os.path.join(HERE, "libpowspec_port.so")

ASSIGN = {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}


class _Params(C.Structure):
    _fields_ = [
        ("ncat", C.c_int), ("issim", C.c_int), ("intlace", C.c_int),
        ("assign", C.c_int), ("gsize", C.c_int), ("logscale", C.c_int),
        ("verbose", C.c_int), ("npole", C.c_int), ("poles", C.c_int * 8),
        ("has_bsize", C.c_int), ("isauto", C.c_int * 2), ("iscross", C.c_int),
        ("los", C.c_double * 3), ("bsize", C.c_double * 3),
        ("bpad", C.c_double * 3),
        ("kmin", C.c_double), ("kmax", C.c_double), ("kbin", C.c_double),
    ]


class _Cats(C.Structure):
    _fields_ = [
        ("data", C.c_void_p * 2), ("rand", C.c_void_p * 2),
        ("ndata", C.c_size_t * 2), ("nrand", C.c_size_t * 2),
        ("wdata", C.c_double * 2), ("wrand", C.c_double * 2),
        ("alpha", C.c_double * 2), ("shot", C.c_double * 2),
        ("norm", C.c_double * 2),
    ]


(GET_K, GET_KEDGE, GET_KM, GET_CNT, GET_LCNT, GET_PL, GET_XPL, GET_SHOT,
 GET_NORM, GET_BMIN, GET_BSIZE, GET_FR, GET_FRL) = range(13)


@dataclass
class OracleResult:
    nbin: int
    nl: int
    k: np.ndarray
    kedge: np.ndarray
    km: np.ndarray
    cnt: np.ndarray
    lcnt: np.ndarray
    pl: list            # per catalogue: (nl, nbin) or None
    xpl: np.ndarray | None
    shot: np.ndarray
    norm: np.ndarray
    bmin: np.ndarray
    bsize: np.ndarray
    t_mesh: float
    t_pk: float
    Fr: list = field(default_factory=list)
    Frl: list = field(default_factory=list)


def survey_scalars(data_wc, data_wfkp, data_nz, rand_wc, rand_wfkp, rand_nz):
    """wdata/wrand/alpha/shot/norm as the reference's ingest computes them
    (io/read_ascii.c:1031-1034, src/read_cata.c:165-183).  Host-side logic that
    stays outside the hot path; restated here so tests can feed surveys."""
    wd = float(np.sum(data_wc))
    wr = float(np.sum(rand_wc))
    sw2d = float(np.sum((data_wc * data_wfkp) ** 2))
    sw2r = float(np.sum((rand_wc * rand_wfkp) ** 2))
    sw2nd = float(np.sum(data_wc * data_wfkp ** 2 * data_nz))
    sw2nr = float(np.sum(rand_wc * rand_wfkp ** 2 * rand_nz))
    alpha = wd / wr
    shot = sw2d + alpha * alpha * sw2r
    if sw2nd == 0:
        norm = alpha * sw2nr
    elif sw2nr == 0:
        norm = sw2nd
    else:
        norm = alpha * sw2nr
    return dict(wdata=wd, wrand=wr, alpha=alpha, shot=shot, norm=norm)


class Oracle:
    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_backend.restype = C.c_char_p
        L.oracle_real_size.restype = C.c_int
        L.oracle_run.restype = C.c_void_p
        L.oracle_run.argtypes = [C.POINTER(_Params), C.POINTER(_Cats), C.c_int]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_nbin.argtypes = [C.c_void_p]
        L.oracle_nl.argtypes = [C.c_void_p]
        L.oracle_ntot.argtypes = [C.c_void_p]
        L.oracle_ntot.restype = C.c_size_t
        L.oracle_time.argtypes = [C.c_void_p, C.c_int]
        L.oracle_time.restype = C.c_double
        L.oracle_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_get.restype = C.c_long
        self.backend = L.oracle_backend().decode()
        self.real_size = L.oracle_real_size()

    def run(self, data, *, ng, assign="TSC", interlace=False, poles=(0, 2, 4),
            box=None, issim=True, rand=None, los=(0.0, 0.0, 1.0), kmin=0.0,
            kmax=-1.0, kbin=0.01, logscale=False, bpad=(0.02, 0.02, 0.02),
            isauto=None, iscross=None, scalars=None, keep_mesh=False,
            verbose=False) -> OracleResult:
        """data: (N,4) float64 array {x,y,z,w}, or a list of 1-2 such arrays.
        rand (surveys): same.  scalars (surveys): list of dicts from
        survey_scalars().  box: scalar, 3-sequence or None (auto, surveys)."""
        datas = list(data) if isinstance(data, (list, tuple)) else [data]
        rands = (list(rand) if isinstance(rand, (list, tuple)) else [rand]) \
            if rand is not None else [None] * len(datas)
        ncat = len(datas)
        par = _Params()
        par.ncat = ncat
        par.issim = int(issim)
        par.intlace = int(interlace)
        par.assign = ASSIGN[assign] if isinstance(assign, str) else int(assign)
        par.gsize = int(ng)
        par.logscale = int(logscale)
        par.verbose = int(verbose)
        poles = sorted(set(int(p) for p in poles))
        par.npole = len(poles)
        for i, p in enumerate(poles):
            par.poles[i] = p
        if box is not None:
            b = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,))
            par.has_bsize = 1
            for i in range(3):
                par.bsize[i] = float(b[i])
        for i in range(3):
            par.los[i] = float(los[i])
            par.bpad[i] = float(bpad[i])
        par.kmin, par.kmax, par.kbin = float(kmin), float(kmax), float(kbin)
        if isauto is None:
            isauto = [True] * ncat + [False] * (2 - ncat)
        if iscross is None:
            iscross = ncat == 2
        par.isauto[0], par.isauto[1] = int(isauto[0]), int(isauto[1])
        par.iscross = int(iscross)

        cats = _Cats()
        keep = []
        for i in range(ncat):
            d = np.ascontiguousarray(datas[i], dtype=np.float64)
            assert d.ndim == 2 and d.shape[1] == 4
            keep.append(d)
            cats.data[i] = d.ctypes.data
            cats.ndata[i] = d.shape[0]
            cats.wdata[i] = float(np.sum(d[:, 3])) if issim else scalars[i]["wdata"]
            if not issim:
                r = np.ascontiguousarray(rands[i], dtype=np.float64)
                keep.append(r)
                cats.rand[i] = r.ctypes.data
                cats.nrand[i] = r.shape[0]
                s = scalars[i]
                cats.wrand[i] = s["wrand"]
                cats.alpha[i] = s["alpha"]
                cats.shot[i] = s["shot"]
                cats.norm[i] = s["norm"]
        L = self.lib
        h = L.oracle_run(C.byref(par), C.byref(cats), int(keep_mesh))
        if not h:
            raise RuntimeError("oracle_run failed (see stderr)")
        try:
            nbin, nl = L.oracle_nbin(h), L.oracle_nl(h)

            def get(what, n, dtype=np.float64, idx=0):
                a = np.empty(n, dtype=dtype)
                got = L.oracle_get(h, what, idx, a.ctypes.data)
                return a if got >= 0 else None

            res = OracleResult(
                nbin=nbin, nl=nl,
                k=get(GET_K, nbin), kedge=get(GET_KEDGE, nbin + 1),
                km=get(GET_KM, nbin), cnt=get(GET_CNT, nbin, np.uint64),
                lcnt=get(GET_LCNT, nl * nbin).reshape(nl, nbin),
                pl=[None, None], xpl=None,
                shot=get(GET_SHOT, 2), norm=get(GET_NORM, 2),
                bmin=get(GET_BMIN, 3), bsize=get(GET_BSIZE, 3),
                t_mesh=L.oracle_time(h, 0), t_pk=L.oracle_time(h, 1))
            for i in range(2):
                p = get(GET_PL, nl * nbin, idx=i)
                res.pl[i] = None if p is None else p.reshape(nl, nbin)
            x = get(GET_XPL, nl * nbin)
            res.xpl = None if x is None else x.reshape(nl, nbin)
            if keep_mesh:
                ntot = L.oracle_ntot(h)
                rdt = np.float64 if self.real_size == 8 else np.float32
                for i in range(ncat):
                    f = get(GET_FR, ntot, rdt, i)
                    res.Fr.append(None if f is None else f.reshape(ng, ng, ng))
                    f = get(GET_FRL, ntot, rdt, i)
                    res.Frl.append(None if f is None else f.reshape(ng, ng, ng))
            return res
        finally:
            L.oracle_free(h)


def have_ref(single=False) -> bool:
    return os.path.exists(REF_LIB_F32 if single else REF_LIB)


def have_port() -> bool:
    return os.path.exists(PORT_LIB)


def build_port(force=False) -> str:
    if force or not have_port() or \
            os.path.getmtime(PORT_LIB) < os.path.getmtime(os.path.join(HERE, "pspec_port.c")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_LIB


def build_ref(reference="/root/reference") -> str | None:
    """Build oracle/_ref from the reference sources where they lie; a no-op
    (returning the prebuilt library, if any) when the tree is absent."""
    if os.path.isdir(os.path.join(reference, "src")):
        # `dropin` also links the reference's unchanged host against the product
        # library (oracle/_ref/POWSPEC_b200) for tests/test_dropin_binary.py
        subprocess.check_call(["make", "-C", HERE, "ref", "dropin", f"REF={reference}"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return REF_LIB if have_ref() else None


_cache: dict = {}


def load_oracle(kind="ref", single=False) -> Oracle:
    """kind: 'ref' (unmodified reference, prebuilt) or 'port' (restatement)."""
    key = (kind, single)
    if key not in _cache:
        if kind == "ref":
            path = REF_LIB_F32 if single else REF_LIB
            if not os.path.exists(path):
                raise FileNotFoundError(path + " (run `make -C oracle ref` where "
                                        "/root/reference exists)")
        elif kind == "port":
            path = build_port()
        else:
            raise ValueError(kind)
        _cache[key] = Oracle(path)
    return _cache[key]


# ---------------------------------------------------------------------------
# coordinate conversion (cnvt_coord, src/cnvt_coord.c:549-582)
# ---------------------------------------------------------------------------
REF_CNVT_LIB = os.path.join(HERE, "_ref", "libpowspec_ref_cnvt.so")


def have_ref_cnvt() -> bool:
    return os.path.exists(REF_CNVT_LIB)


def ref_cnvt(arrays, *, omega_m=0.31, omega_l=0.69, omega_k=0.0, eos_w=-1.0, ecdst=1e-8,
             fcdst=None):
    """The UNMODIFIED reference cnvt_coord() on copies of ``arrays`` ((N, 4) float64,
    {RA deg, Dec deg, z, w}); ``fcdst``: path of a (z, distance) table, or None
    for the Legendre-Gauss integration.  Up to two arrays (data catalogues)."""
    lib = C.CDLL(REF_CNVT_LIB)
    out = [np.array(a, dtype=np.float64, order="C", copy=True) for a in arrays]
    n = len(out)
    assert 1 <= n <= 2
    ptrs = (C.c_void_p * 2)(*([o.ctypes.data for o in out] + [None] * (2 - n)))
    cnts = (C.c_size_t * 2)(*([o.shape[0] for o in out] + [0] * (2 - n)))
    none = (C.c_void_p * 2)(None, None)
    zero = (C.c_size_t * 2)(0, 0)
    yes = (C.c_int * 2)(1, 1)
    no = (C.c_int * 2)(0, 0)
    lib.oracle_cnvt.restype = C.c_int
    lib.oracle_cnvt.argtypes = [C.c_double] * 5 + [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.oracle_cnvt(omega_m, omega_l, omega_k, eos_w, ecdst,
                         None if fcdst is None else os.fsencode(fcdst), n, ptrs, cnts, none, zero,
                         yes, no)
    if rc:
        raise RuntimeError(f"reference cnvt_coord failed: {rc}")
    return out


def port_cnvt(arrays, *, omega_m=0.31, omega_l=0.69, omega_k=0.0, eos_w=-1.0, ecdst=1e-8,
              samples=None):
    """CPU restatement (oracle/pspec_port.c: oracle_cnvt); ``samples`` = (z[], d[]) or
    None.  Returns (converted copies, Legendre-Gauss order)."""
    lib = C.CDLL(build_port())
    out = [np.array(a, dtype=np.float64, order="C", copy=True) for a in arrays]
    n = len(out)
    ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in out])
    cnts = (C.c_size_t * n)(*[o.shape[0] for o in out])
    order = C.c_int(0)
    sz = sd = None
    ns = 0
    if samples is not None:
        z = np.ascontiguousarray(samples[0], dtype=np.float64)
        d = np.ascontiguousarray(samples[1], dtype=np.float64)
        sz, sd, ns = z.ctypes.data, d.ctypes.data, len(z)
    lib.oracle_cnvt.restype = C.c_int
    lib.oracle_cnvt.argtypes = [C.c_double] * 5 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    rc = lib.oracle_cnvt(omega_m, omega_l, omega_k, eos_w, ecdst, sz, sd, ns, n, ptrs, cnts,
                         C.byref(order))
    if rc:
        raise RuntimeError("oracle_cnvt (port) failed")
    return out, order.value
