"""CPU-side checks of the product library (no GPU needed): it loads, exports
every symbol the public headers declare, and fails loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", txt)
    return sorted(set(n for n in names if n.startswith(("psb_", "genr_", "mesh_", "powspec"))))


@pytest.fixture(scope="module")
def lib():
    import powspec_b200
    from powspec_b200.build import build
    build()
    return powspec_b200.load_library()


def test_exports_every_declared_symbol(lib):
    fns = _declared_functions("powspec_b200.h") + _declared_functions("powspec_refabi.h")
    assert len(fns) >= 23
    for name in fns:
        assert hasattr(lib, name), f"libpowspec_b200.so does not export {name}"
    names = (C.c_char_p * 4).in_dll(lib, "powspec_assign_names")
    assert [n.decode() for n in names] == ["NGP", "CIC", "TSC", "PCS"]


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to run instead of falling back."""
    import powspec_b200
    if lib.psb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(powspec_b200.PowspecB200Error, match="no CUDA device"):
        powspec_b200.Context(0)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under powspec_b200/ may reference it."""
    for base, _, files in os.walk(os.path.join(ROOT, "powspec_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "oracle/" not in txt and "libpowspec_port" not in txt and "libpowspec_ref" not in txt, f


def test_refabi_struct_layout_matches_reference_headers(tmp_path):
    """Compile a TU that includes BOTH the reference headers (by include path)
    and our ABI mirror, and static-assert sizes and offsets."""
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("reference tree absent (GPU box)")
    members = {
        "CONF": ["ndata", "issim", "los", "bsize", "bpad", "gsize", "assign", "intlace", "poles",
                 "npole", "kmin", "kmax", "logscale", "kbin", "isauto", "iscross", "verbose"],
        "CATA": ["num", "data", "rand", "ndata", "nrand", "wdata", "wrand", "alpha", "shot", "norm"],
        "MESH": ["num", "Ng", "Ngk", "Ntot", "Ncmplx", "min", "max", "smin", "bsize", "issim",
                 "intlace", "fft_init", "assign", "r2c", "Fr", "alias", "Fka"],
        "PK": ["issim", "log", "isauto", "iscross", "nl", "nbin", "nmu", "poles", "los", "dk",
               "kedge", "k", "km", "cnt", "lcnt", "pl", "xpl", "nomp", "pcnt", "plcnt"],
        "DATA": ["x", "w"],
    }
    src = ['#include <stddef.h>', '#include "load_conf.h"', '#include "read_cata.h"',
           '#include "genr_mesh.h"', '#include "multipole.h"',
           '#define PSB_REFABI_NO_PROTOTYPES', '#include "powspec_refabi.h"']
    for s, ms in members.items():
        src.append(f'_Static_assert(sizeof({s}) == sizeof(psb_ref_{s}), "sizeof {s}");')
        for m in ms:
            src.append(f'_Static_assert(offsetof({s}, {m}) == offsetof(psb_ref_{s}, {m}), "{s}.{m}");')
    src.append("int main(void) { return 0; }")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    import subprocess
    for extra in ([], ["-DSINGLE_PREC"]):
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-DOMP", "-fopenmp", *extra,
                               "-I", os.path.join(ROOT, "oracle", "fftw_shim"),
                               "-I", os.path.join(ref, "src"), "-I", os.path.join(ROOT, "include"),
                               "-c", str(c), "-o", str(tmp_path / "layout.o")])


def test_staging_copy_pool(lib):
    """The pageable -> pinned staging copy (csrc/hostcopy.cpp): byte-exact for aligned and
    ragged sizes / offsets, one thread and many, reused over several pieces."""
    import numpy as np
    lib.psb_test_host_copy.restype = C.c_int
    lib.psb_test_host_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, size=(24 << 20) + 4099, dtype=np.uint8)
    for nthr in (1, 3, 16):
        for off_s, off_d, n in ((0, 0, src.size), (1, 0, 5 << 20), (3, 13, (9 << 20) + 77),
                                (0, 64, 100), (5, 7, 255), (0, 0, 0), (8, 24, 4096 * 33 + 1)):
            dst = np.zeros(src.size + 128, dtype=np.uint8)
            got = lib.psb_test_host_copy(dst.ctypes.data + off_d, src.ctypes.data + off_s, n, nthr, 3)
            assert got == nthr
            assert np.array_equal(dst[off_d:off_d + n], src[off_s:off_s + n]), (nthr, off_s, off_d, n)
            assert not dst[:off_d].any() and not dst[off_d + n:].any(), (nthr, off_s, off_d, n)
