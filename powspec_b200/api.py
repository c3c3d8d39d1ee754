"""Python host mirror of the reference's operator interface for the hot path.

Names and argument meaning follow the reference's C entry points
(cheng-zhao/powspec):

    genr_mesh(conf, cata)        src/genr_mesh.h:86   -> Mesh
    powspec(conf, cata, mesh)    src/multipole.h:78   -> PK
    mesh_destroy(mesh)           src/genr_mesh.h:94
    powspec_destroy(pk)          src/multipole.h:86
    powspec_assign_names         src/genr_mesh.h:45

``Conf`` carries the members of CONF that the two stages read
(src/load_conf.h:40-93), ``Cata`` the members of CATA (src/read_cata.h:47-59),
``PK`` the members of PK the host reads back (src/multipole.h:38-61).  Errors:
the C layer prints the reference's P_ERR-style message and returns NULL; here
that becomes ``PowspecB200Error`` (the reference's main() maps it to
POWSPEC_ERR_MESH / POWSPEC_ERR_PK, src/powspec.c:47-60).

Everything numerical happens in ``libpowspec_b200.so``; this file only marshals
pointers.  PyTorch is optional here and only used to hand over device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
powspec_assign_names = ["NGP", "CIC", "TSC", "PCS"]

POWSPEC_ERR_CATA = -11      # src/define.h:119
POWSPEC_ERR_CNVT = -12      # src/define.h:120
POWSPEC_ERR_MESH = -13      # src/define.h:121
POWSPEC_ERR_PK = -14        # src/define.h:122

(T_H2D, T_BOUNDS, T_SORT, T_MEMSET, T_ASSIGN, T_FFT, T_GEOM, T_BIN, T_YLM, T_FFT_STRIDED,
 T_CNVT, T_TOTAL, T_COUNT) = range(13)
TIMING_NAMES = ["h2d", "bounds", "sort", "memset", "assign", "fft", "geom", "bin", "ylm",
                "fft_strided", "cnvt", "total"]

(GET_K, GET_KEDGE, GET_KM, GET_CNT, GET_LCNT, GET_PL, GET_XPL, GET_SHOT, GET_NORM,
 GET_BMIN, GET_BSIZE, GET_BMAX) = range(12)


class PowspecB200Error(RuntimeError):
    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


class _Params(C.Structure):
    _fields_ = [
        ("ncat", C.c_int), ("issim", C.c_int), ("intlace", C.c_int), ("assign", C.c_int),
        ("gsize", C.c_int), ("logscale", C.c_int), ("verbose", C.c_int), ("npole", C.c_int),
        ("poles", C.c_int * 8), ("has_bsize", C.c_int), ("isauto", C.c_int * 2),
        ("iscross", C.c_int), ("los", C.c_double * 3), ("bsize", C.c_double * 3),
        ("bpad", C.c_double * 3), ("kmin", C.c_double), ("kmax", C.c_double),
        ("kbin", C.c_double), ("precision", C.c_int), ("device", C.c_int),
    ]


class _Cosmo(C.Structure):
    _fields_ = [
        ("omega_m", C.c_double), ("omega_l", C.c_double), ("omega_k", C.c_double),
        ("eos_w", C.c_double), ("ecdst", C.c_double),
        ("sample_z", C.c_void_p), ("sample_d", C.c_void_p), ("nsample", C.c_size_t),
    ]


class _Columns(C.Structure):
    _fields_ = [("pos", C.c_int * 3), ("wcomp", C.c_int), ("wfkp", C.c_int), ("nz", C.c_int)]


class _CatalogSums(C.Structure):
    _fields_ = [("n", C.c_size_t), ("sumw", C.c_double), ("sumw2", C.c_double), ("sumw2n", C.c_double)]


class _Cats(C.Structure):
    _fields_ = [
        ("data", C.c_void_p * 2), ("rand", C.c_void_p * 2),
        ("ndata", C.c_size_t * 2), ("nrand", C.c_size_t * 2),
        ("wdata", C.c_double * 2), ("wrand", C.c_double * 2), ("alpha", C.c_double * 2),
        ("shot", C.c_double * 2), ("norm", C.c_double * 2), ("memspace", C.c_int),
        ("cnvt", C.POINTER(_Cosmo)), ("dcnvt", C.c_int * 2), ("rcnvt", C.c_int * 2),
    ]


_lib = None


def library_path() -> str:
    # POWSPEC_B200_LIBRARY: an ablation build of the same library (tools/build_variant.sh)
    return os.environ.get("POWSPEC_B200_LIBRARY") or os.path.join(HERE, "libpowspec_b200.so")


def load_library():
    """dlopen libpowspec_b200.so (built in-tree by powspec_b200/build.py).
    Fails loudly if it is missing: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise PowspecB200Error(
            f"{path} is missing: build it with `python -m powspec_b200.build` "
            "(nvcc, sm_100a). powspec_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.psb_last_error.restype = C.c_char_p
    L.psb_device_count.restype = C.c_int
    L.psb_create.restype = C.c_void_p
    L.psb_create.argtypes = [C.c_int]
    L.psb_destroy.argtypes = [C.c_void_p]
    L.psb_run.restype = C.c_void_p
    L.psb_run.argtypes = [C.c_void_p, C.POINTER(_Params), C.POINTER(_Cats)]
    L.psb_mesh.restype = C.c_int
    L.psb_mesh.argtypes = [C.c_void_p, C.POINTER(_Params), C.POINTER(_Cats)]
    L.psb_power.restype = C.c_void_p
    L.psb_power.argtypes = [C.c_void_p, C.POINTER(_Params)]
    L.psb_result_free.argtypes = [C.c_void_p]
    L.psb_result_nbin.argtypes = [C.c_void_p]
    L.psb_result_nl.argtypes = [C.c_void_p]
    L.psb_result_get.restype = C.c_long
    L.psb_result_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.psb_copy_mesh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.psb_mesh_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.psb_timings.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.psb_launch_count.restype = C.c_long
    L.psb_launch_count.argtypes = [C.c_void_p]
    L.psb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
    L.psb_assign_path.argtypes = [C.c_void_p]
    L.psb_tile_overflow.restype = C.c_long
    L.psb_tile_overflow.argtypes = [C.c_void_p]
    L.psb_stream.restype = C.c_void_p
    L.psb_stream.argtypes = [C.c_void_p]
    L.psb_generate_catalog.restype = C.c_void_p
    L.psb_generate_catalog.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_uint64]
    L.psb_device_free.argtypes = [C.c_void_p, C.c_void_p]
    L.psb_generate_into.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_uint64, C.c_uint64]
    L.psb_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.psb_catalog_probe.argtypes = [C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.psb_catalog_load.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_Columns), C.c_int,
                                   C.POINTER(C.c_void_p), C.POINTER(_CatalogSums)]
    L.psb_cnvt_coord.argtypes = [C.c_void_p, C.POINTER(_Cosmo), C.c_void_p, C.c_void_p, C.c_int,
                                 C.POINTER(C.c_int)]
    L.psb_cnvt_order.argtypes = [C.POINTER(_Cosmo), C.c_double, C.c_double]
    L.psb_fft_axis.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    vp = C.c_void_p
    L.psb_dist_unique_id.argtypes = [vp]
    L.psb_dist_create_nccl.restype = vp
    L.psb_dist_create_nccl.argtypes = [vp, C.c_int, C.c_int, vp]
    L.psb_dist_destroy.argtypes = [vp]
    L.psb_dist_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.psb_dist_begin.argtypes = [vp, C.POINTER(_Params)]
    L.psb_dist_add.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.psb_dist_add_host.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int]
    L.psb_dist_finish.restype = vp
    L.psb_dist_finish.argtypes = [vp, C.POINTER(C.c_double)]
    L.psb_dist_timings.argtypes = [vp, vp, C.c_int]
    L.psb_dist_traffic.argtypes = [vp, vp, C.c_int]
    L.psb_dist_transport.restype = C.c_char_p
    L.psb_dist_transport.argtypes = [vp]
    L.psb_dist_context.restype = vp
    L.psb_dist_context.argtypes = [vp]
    L.psb_group_create.restype = vp
    L.psb_group_create.argtypes = [C.POINTER(C.c_int), C.c_int]
    L.psb_group_destroy.argtypes = [vp]
    L.psb_group_size.argtypes = [vp]
    L.psb_group_rank.restype = vp
    L.psb_group_rank.argtypes = [vp, C.c_int]
    L.psb_group_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.psb_group_mesh.argtypes = [vp, C.POINTER(_Params), C.POINTER(_Cats)]
    L.psb_group_power.restype = vp
    L.psb_group_power.argtypes = [vp, C.POINTER(_Params)]
    L.psb_group_run.restype = vp
    L.psb_group_run.argtypes = [vp, C.POINTER(_Params), C.POINTER(_Cats)]
    _lib = L
    return L


def _err(L, what, code=None):
    msg = L.psb_last_error().decode(errors="replace").strip()
    return PowspecB200Error(f"{what}: {msg}" if msg else what, code)


# ---------------------------------------------------------------------------
# the reference's data structures, host side
# ---------------------------------------------------------------------------
@dataclass
class Conf:
    """Members of CONF read by genr_mesh() and powspec() (src/load_conf.h:40-93)."""
    ndata: int = 1                  # number of catalogues
    issim: bool = True              # CUBIC_SIM
    los: tuple = (0.0, 0.0, 1.0)    # LINE_OF_SIGHT
    bsize: tuple | None = None      # BOX_SIZE (3 values) or None
    bpad: tuple = (0.02, 0.02, 0.02)    # BOX_PAD
    gsize: int = 256                # GRID_SIZE
    assign: int = 2                 # PARTICLE_ASSIGN (index into powspec_assign_names)
    intlace: bool = False           # GRID_INTERLACE
    poles: tuple = (0, 2, 4)        # MULTIPOLE (sorted, unique)
    kmin: float = 0.0               # KMIN (log10 of it for LOG_SCALE, src/load_conf.c:1352-1355)
    kmax: float = -1.0              # KMAX; <= 0: unset (src/define.h:67-68)
    logscale: bool = False          # LOG_SCALE
    kbin: float = 0.01              # BIN_SIZE
    isauto: tuple = (True, False)
    iscross: bool = False
    verbose: bool = False
    # outputs, read by save_res() (src/save_res.c:35-229)
    oauto: tuple | None = None      # OUTPUT_AUTO: one file name per catalogue
    ocross: str | None = None       # OUTPUT_CROSS
    oheader: bool = True            # OUTPUT_HEADER (DEFAULT_HEADER, src/define.h:62)
    # coordinate conversion, read by cnvt_coord() (src/cnvt_coord.c:549-582)
    cnvt: bool = False              # any of DATA_CONVERT / RAND_CONVERT set
    dcnvt: tuple = (False, False)   # DATA_CONVERT per catalogue
    rcnvt: tuple = (False, False)   # RAND_CONVERT per catalogue
    omega_m: float = 0.31           # OMEGA_M
    omega_l: float = 0.69           # OMEGA_LAMBDA
    omega_k: float = 0.0            # 1 - OMEGA_M - OMEGA_LAMBDA
    eos_w: float = -1.0             # DE_EOS_W
    ecdst: float = 1e-8             # CMVDST_ERR
    fcdst: tuple | None = None      # Z_CMVDST_CNVT: here the file's contents, (z[], d[])
    # not in the reference's CONF: compile-time -DSINGLE_PREC there
    precision: int = 8
    device: int = 0

    def _cosmo(self):
        """(psb_cosmo, keepalive) for cnvt_coord, or (None, None)."""
        if not self.cnvt:
            return None, None
        cm = _Cosmo()
        cm.omega_m, cm.omega_l, cm.omega_k = float(self.omega_m), float(self.omega_l), float(self.omega_k)
        cm.eos_w, cm.ecdst = float(self.eos_w), float(self.ecdst)
        keep = None
        if self.fcdst is not None:
            z = np.ascontiguousarray(self.fcdst[0], dtype=np.float64)
            d = np.ascontiguousarray(self.fcdst[1], dtype=np.float64)
            cm.sample_z, cm.sample_d, cm.nsample = z.ctypes.data, d.ctypes.data, len(z)
            keep = (z, d)
        return cm, keep

    def _c(self) -> _Params:
        p = _Params()
        p.ncat = self.ndata
        p.issim, p.intlace, p.assign = int(self.issim), int(self.intlace), int(self.assign)
        p.gsize, p.logscale, p.verbose = int(self.gsize), int(self.logscale), int(self.verbose)
        poles = list(self.poles)
        p.npole = len(poles)
        for i, v in enumerate(poles[:8]):
            p.poles[i] = int(v)
        p.has_bsize = int(self.bsize is not None)
        if self.bsize is not None:
            b = np.broadcast_to(np.asarray(self.bsize, dtype=np.float64), (3,))
            for i in range(3):
                p.bsize[i] = float(b[i])
        for i in range(3):
            p.los[i] = float(self.los[i])
            p.bpad[i] = float(self.bpad[i])
        p.isauto[0], p.isauto[1] = int(self.isauto[0]), int(self.isauto[1])
        p.iscross = int(self.iscross)
        p.kmin, p.kmax, p.kbin = float(self.kmin), float(self.kmax), float(self.kbin)
        p.precision, p.device = int(self.precision), int(self.device)
        return p


@dataclass
class Cata:
    """CATA (src/read_cata.h:47-59).  ``data`` / ``rand``: per catalogue an
    (N, 4) float64 array of DATA records {x, y, z, w}: a numpy array (host), or
    a torch CUDA tensor / (device_ptr, n) pair (already resident in HBM)."""
    data: list
    rand: list | None = None
    wdata: list | None = None
    wrand: list | None = None
    alpha: list | None = None
    shot: list | None = None
    norm: list | None = None

    @property
    def num(self):
        return len(self.data)


@dataclass
class Mesh:
    """MESH metadata (src/genr_mesh.h:48-70); the fields stay on the device."""
    ctx: "Context"
    Ng: int
    min: np.ndarray
    max: np.ndarray
    bsize: np.ndarray
    issim: bool
    intlace: bool
    assign: int
    num: int

    def field(self, cat=0, shifted=False) -> np.ndarray:
        """Copy of Fr (or Frl, the half-cell shifted field) as (Ng,Ng,Ng)."""
        return self.ctx.copy_mesh(cat, int(shifted))


@dataclass
class PK:
    """PK (src/multipole.h:38-61) as the host reads it back (src/save_res.c:100-124)."""
    nl: int
    nbin: int
    poles: list
    k: np.ndarray
    kedge: np.ndarray
    km: np.ndarray
    cnt: np.ndarray
    lcnt: np.ndarray
    pl: list
    xpl: np.ndarray | None
    shot: np.ndarray
    norm: np.ndarray
    timings_ms: dict = field(default_factory=dict)
    launches: int = 0


def pk_from_result(L, r, conf) -> "PK":
    """Copy a psb_result into the host-side PK (the caller frees r)."""
    nbin, nl = L.psb_result_nbin(r), L.psb_result_nl(r)

    def get(what, n, dtype=np.float64, idx=0):
        a = np.empty(n, dtype=dtype)
        return a if L.psb_result_get(r, what, idx, a.ctypes.data) >= 0 else None

    pl = []
    for i in range(2):
        q = get(GET_PL, nl * nbin, idx=i)
        pl.append(None if q is None else q.reshape(nl, nbin))
    x = get(GET_XPL, nl * nbin)
    return PK(nl=nl, nbin=nbin, poles=list(conf.poles), k=get(GET_K, nbin),
              kedge=get(GET_KEDGE, nbin + 1), km=get(GET_KM, nbin),
              cnt=get(GET_CNT, nbin, np.uint64), lcnt=get(GET_LCNT, nl * nbin).reshape(nl, nbin),
              pl=pl, xpl=None if x is None else x.reshape(nl, nbin),
              shot=get(GET_SHOT, 2), norm=get(GET_NORM, 2))


def _ptr_of(arr):
    """(pointer, n, memspace, keepalive) of one particle array."""
    if arr is None:
        return 0, 0, 0, None
    if isinstance(arr, tuple):          # (device pointer, n)
        return int(arr[0]), int(arr[1]), 1, None
    if hasattr(arr, "data_ptr"):        # torch tensor
        t = arr
        assert t.dim() == 2 and t.shape[1] == 4 and str(t.dtype) == "torch.float64"
        t = t.contiguous()
        return t.data_ptr(), t.shape[0], (1 if t.is_cuda else 0), t
    a = np.ascontiguousarray(arr, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 4, "particles must be (N, 4) {x,y,z,w}"
    return a.ctypes.data, a.shape[0], 0, a


class Context:
    """psb_context: CUDA stream, mesh buffers, cuFFT plans (include/powspec_b200.h)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.device = device
        self.h = self.L.psb_create(device)
        if not self.h:
            raise _err(self.L, "psb_create", POWSPEC_ERR_MESH)
        self._conf = None

    def close(self):
        if getattr(self, "h", None):
            self.L.psb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        if self.L.psb_set_option(self.h, name.encode(), int(value)):
            raise _err(self.L, "psb_set_option")

    def torch_stream(self):
        """The library's compute stream as a torch stream (timing events, ordering)."""
        import torch
        return torch.cuda.ExternalStream(self.L.psb_stream(self.h), device=torch.device("cuda", self.device))

    # -- genr_mesh
    def genr_mesh(self, conf: Conf, cata: Cata) -> Mesh:
        p = conf._c()
        cats = _Cats()
        keep = []
        spaces = set()
        for i in range(cata.num):
            ptr, n, sp, ka = _ptr_of(cata.data[i])
            keep.append(ka)
            spaces.add(sp)
            cats.data[i], cats.ndata[i] = ptr, n
            if cata.wdata is not None:
                cats.wdata[i] = float(cata.wdata[i])
            elif sp == 0:
                cats.wdata[i] = float(np.sum(np.asarray(cata.data[i])[:, 3]))
            else:
                raise PowspecB200Error("Cata.wdata is required for device-resident catalogues")
            if not conf.issim:
                ptr, n, sp, ka = _ptr_of(cata.rand[i])
                keep.append(ka)
                spaces.add(sp)
                cats.rand[i], cats.nrand[i] = ptr, n
                cats.wrand[i] = float(cata.wrand[i])
                cats.alpha[i] = float(cata.alpha[i])
                cats.shot[i] = float(cata.shot[i])
                cats.norm[i] = float(cata.norm[i])
        if len(spaces) != 1:
            raise PowspecB200Error("all catalogues must live in the same memory space")
        cats.memspace = spaces.pop()
        cosmo, cosmo_keep = conf._cosmo()
        if cosmo is not None:
            cats.cnvt = C.pointer(cosmo)
            for i in range(cata.num):
                cats.dcnvt[i], cats.rcnvt[i] = int(conf.dcnvt[i]), int(conf.rcnvt[i])
        if self.L.psb_mesh(self.h, C.byref(p), C.byref(cats)):
            raise _err(self.L, "genr_mesh", POWSPEC_ERR_MESH)
        bmin, bsize, bmax = np.zeros(3), np.zeros(3), np.zeros(3)
        self.L.psb_mesh_box(self.h, bmin.ctypes.data, bsize.ctypes.data, bmax.ctypes.data)
        self._conf = conf
        return Mesh(ctx=self, Ng=conf.gsize, min=bmin, max=bmax, bsize=bsize, issim=conf.issim,
                    intlace=conf.intlace, assign=conf.assign, num=cata.num)

    def copy_mesh(self, cat: int, fld: int) -> np.ndarray:
        conf = self._conf
        ng = conf.gsize
        out = np.empty((ng, ng, ng), dtype=np.float64 if conf.precision == 8 else np.float32)
        if self.L.psb_copy_mesh(self.h, cat, fld, out.ctypes.data):
            raise _err(self.L, "psb_copy_mesh")
        return out

    # -- powspec
    def powspec(self, conf: Conf, cata: Cata | None, mesh: Mesh) -> PK:
        p = conf._c()
        r = self.L.psb_power(self.h, C.byref(p))
        if not r:
            raise _err(self.L, "powspec", POWSPEC_ERR_PK)
        try:
            pk = pk_from_result(self.L, r, conf)
            pk.timings_ms = self.timings()
            pk.launches = int(self.L.psb_launch_count(self.h))
            return pk
        finally:
            self.L.psb_result_free(r)

    def timings(self) -> dict:
        ms = (C.c_double * T_COUNT)()
        self.L.psb_timings(self.h, ms, T_COUNT)
        return {TIMING_NAMES[i]: ms[i] for i in range(T_COUNT - 1)}

    # -- synthetic catalogues on the device
    def generate_catalog(self, n: int, boxsize: float, kind: int = 0, seed: int = 1):
        ptr = self.L.psb_generate_catalog(self.h, n, float(boxsize), int(kind), int(seed))
        if not ptr:
            raise _err(self.L, "psb_generate_catalog")
        return (ptr, n)

    def generate_into(self, tensor, boxsize: float, kind: int = 0, seed: int = 1, first_index: int = 0):
        """Fill a (n, 4) float64 CUDA tensor with particles first_index.. of the catalogue."""
        if self.L.psb_generate_into(self.h, tensor.data_ptr(), tensor.shape[0], float(boxsize), int(kind),
                                    int(seed), int(first_index)):
            raise _err(self.L, "psb_generate_into")
        return tensor

    def load_catalog(self, path, pos=(0, 1, 2), wcomp=None, wfkp=None, nz=None, issim=True):
        """Binary catalogue ingest (psb_catalog_load): a .npy file (N, ncols) of float64 /
        float32 -> device records {x, y, z, w} and the sums read_ascii_data() keeps
        (io/read_ascii.c:868-902).  Returns ((device_ptr, n), sums dict); release the
        records with free_catalog()."""
        cols = _Columns()
        for i in range(3):
            cols.pos[i] = int(pos[i])
        cols.wcomp = -1 if wcomp is None else int(wcomp)
        cols.wfkp = -1 if wfkp is None else int(wfkp)
        cols.nz = -1 if nz is None else int(nz)
        ptr = C.c_void_p()
        sums = _CatalogSums()
        if self.L.psb_catalog_load(self.h, os.fsencode(path), C.byref(cols), int(issim), C.byref(ptr),
                                   C.byref(sums)):
            raise _err(self.L, "read_cata", POWSPEC_ERR_CATA)
        return (ptr.value, sums.n), dict(n=sums.n, sumw=sums.sumw, sumw2=sums.sumw2, sumw2n=sums.sumw2n)

    def read_cata(self, conf: Conf, data_files, rand_files=None, *, pos=(0, 1, 2), wcomp=None,
                  wfkp=None, nz=None) -> "Cata":
        """read_cata() (src/read_cata.c:86-189) for binary catalogues: one .npy file per
        data (and, for surveys, random) catalogue with the same column layout.  The
        records stay on the device; wdata / wrand / alpha / shot / norm are formed
        from the sums exactly as the reference does (:160-183)."""
        data, rand, wd, wr, alpha, shot, norm = [], [], [], [], [], [], []
        for i, f in enumerate(data_files):
            d, sd = self.load_catalog(f, pos, wcomp, None if conf.issim else wfkp,
                                      None if conf.issim else nz, conf.issim)
            data.append(d)
            wd.append(sd["sumw"])
            if conf.issim:
                continue
            r, sr = self.load_catalog(rand_files[i], pos, wcomp, wfkp, nz, False)
            rand.append(r)
            wr.append(sr["sumw"])
            if sd["sumw"] == 0 or sd["sumw2"] == 0:
                raise PowspecB200Error("invalid completeness or FKP weights in the data catalog",
                                       POWSPEC_ERR_CATA)
            if sr["sumw"] == 0 or sr["sumw2"] == 0:
                raise PowspecB200Error("invalid completeness or FKP weights in the random catalog",
                                       POWSPEC_ERR_CATA)
            a = sd["sumw"] / sr["sumw"]
            alpha.append(a)
            shot.append(sd["sumw2"] + a * a * sr["sumw2"])
            norm.append(sd["sumw2n"] if sr["sumw2n"] == 0 else a * sr["sumw2n"])
        if conf.issim:
            return Cata(data=data, wdata=wd)
        return Cata(data=data, rand=rand, wdata=wd, wrand=wr, alpha=alpha, shot=shot, norm=norm)

    def cnvt_coord(self, conf: Conf, tensors):
        """cnvt_coord() (src/cnvt_coord.c:549-582) in place on (N, 4) float64 CUDA
        tensors {RA deg, Dec deg, z, w}; returns the Legendre-Gauss order used
        (0 for the interpolation mode)."""
        cosmo, keep = conf._cosmo()
        if cosmo is None:
            return 0
        n = len(tensors)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        cnts = (C.c_size_t * n)(*[t.shape[0] for t in tensors])
        order = C.c_int(0)
        if self.L.psb_cnvt_coord(self.h, C.byref(cosmo), ptrs, cnts, n, C.byref(order)):
            raise _err(self.L, "cnvt_coord", POWSPEC_ERR_CNVT)
        return order.value

    @staticmethod
    def _after_torch():
        """The library works on its own (non-blocking) stream: a tensor that torch has just
        produced on ITS stream must be complete before the library touches it (found by
        running the FFT tests under compute-sanitizer, whose slowdown exposed the missing
        ordering in the test helpers)."""
        import torch
        torch.cuda.current_stream().synchronize()

    def fft_axis(self, tensor, axis: int):
        """In-place forward FFT along axis 0 or 1 of a 3-D complex CUDA tensor
        (hand-written strided pass; the other two axes are (outer, k))."""
        prec = 8 if tensor.element_size() == 16 else 4
        ng = tensor.shape[axis]
        self._after_torch()
        if self.L.psb_fft_axis(self.h, tensor.data_ptr(), prec, ng, tensor.shape[2], axis,
                               tensor.shape[1 - axis]):
            raise _err(self.L, "psb_fft_axis")
        return tensor

    def fft_rows(self, tensor, ng: int):
        """In-place r2c FFT of the rows of a real (nrows, 2 (ng/2+1)) CUDA tensor
        (hand-written z pass); returns the complex (nrows, ng/2+1) view."""
        import torch
        prec = tensor.element_size()
        self._after_torch()
        if self.L.psb_fft_axis(self.h, tensor.data_ptr(), prec, ng, ng // 2 + 1, 2, tensor.shape[0]):
            raise _err(self.L, "psb_fft_axis")
        return torch.view_as_complex(tensor.view(tensor.shape[0], ng // 2 + 1, 2))

    def fft_zy(self, tensor):
        """In-place r2c along z then c2c along y of a real (nplanes, ng, 2 (ng/2+1)) CUDA
        tensor (fused persistent kernel); returns the complex (nplanes, ng, ng/2+1) view."""
        import torch
        npl, ng, rl = tensor.shape
        self._after_torch()
        if self.L.psb_fft_axis(self.h, tensor.data_ptr(), tensor.element_size(), ng, rl // 2, 3, npl):
            raise _err(self.L, "psb_fft_axis")
        return torch.view_as_complex(tensor.view(npl, ng, rl // 2, 2))

    def free_catalog(self, cat):
        self.L.psb_device_free(self.h, cat[0])

    def catalog_to_host(self, cat) -> np.ndarray:
        out = np.empty((cat[1], 4), dtype=np.float64)
        if self.L.psb_copy_to_host(self.h, out.ctypes.data, cat[0], out.nbytes):
            raise _err(self.L, "psb_copy_to_host")
        return out


# ---------------------------------------------------------------------------
# module-level functions with the reference's names
# ---------------------------------------------------------------------------
_default_ctx: dict = {}


def _ctx_for(device: int) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def genr_mesh(conf: Conf, cata: Cata) -> Mesh:
    return _ctx_for(conf.device).genr_mesh(conf, cata)


def powspec(conf: Conf, cata: Cata | None, mesh: Mesh) -> PK:
    return mesh.ctx.powspec(conf, cata, mesh)


def mesh_destroy(mesh: Mesh | None):
    """Releases nothing on its own: buffers belong to the context and are reused."""
    return None


def powspec_destroy(pk: PK | None):
    return None


def run(data, *, ng, assign="TSC", interlace=False, poles=(0, 2, 4), box=None, issim=True,
        rand=None, los=(0.0, 0.0, 1.0), kmin=0.0, kmax=-1.0, kbin=0.01, logscale=False,
        bpad=(0.02, 0.02, 0.02), isauto=None, iscross=None, scalars=None, precision=8,
        device=0, verbose=False, ctx: Context | None = None, keep_mesh=False, wdata=None):
    """Convenience wrapper with the same keywords as oracle.Oracle.run (tests)."""
    datas = list(data) if isinstance(data, (list, tuple)) and not (
        isinstance(data, tuple) and len(data) == 2 and isinstance(data[0], int)) else [data]
    ncat = len(datas)
    if isauto is None:
        isauto = [True] * ncat + [False] * (2 - ncat)
    if iscross is None:
        iscross = ncat == 2
    conf = Conf(ndata=ncat, issim=issim, los=tuple(los),
                bsize=None if box is None else tuple(np.broadcast_to(np.asarray(box, float), (3,))),
                bpad=tuple(bpad), gsize=ng,
                assign=powspec_assign_names.index(assign) if isinstance(assign, str) else assign,
                intlace=interlace, poles=tuple(sorted(set(poles))), kmin=kmin, kmax=kmax,
                logscale=logscale, kbin=kbin, isauto=tuple(isauto), iscross=iscross,
                verbose=verbose, precision=precision, device=device)
    cata = Cata(data=datas, wdata=wdata)
    if not issim:
        rands = list(rand) if isinstance(rand, (list, tuple)) else [rand]
        cata.rand = rands
        cata.wdata = [s["wdata"] for s in scalars]
        cata.wrand = [s["wrand"] for s in scalars]
        cata.alpha = [s["alpha"] for s in scalars]
        cata.shot = [s["shot"] for s in scalars]
        cata.norm = [s["norm"] for s in scalars]
    c = ctx or _ctx_for(device)
    mesh = c.genr_mesh(conf, cata)
    fields = None
    if keep_mesh:
        fields = ([mesh.field(i) for i in range(ncat)],
                  [mesh.field(i, True) for i in range(ncat)] if interlace else [None] * ncat)
    pk = c.powspec(conf, cata, mesh)
    if keep_mesh:
        pk.Fr, pk.Frl = fields
    return pk
