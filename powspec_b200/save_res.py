"""Host-side mirror of the reference's output writer (`save_res`,
src/save_res.c:35-229): the same text files — header lines, column indicator,
`%.10lg` numbers (OFMT_DBL, src/define.h:107) — from the Python structures of
`powspec_b200.api`.  Pure host code (no GPU); it exists so that a pipeline
driving the library from Python writes files a consumer of the reference's
outputs cannot tell apart.  The C host keeps using its own `save_res`."""
from __future__ import annotations

import numpy as np

from .api import Cata, Conf, Mesh, PK, PowspecB200Error, powspec_assign_names

POWSPEC_ERR_FILE = -3           # src/define.h
COMMENT = "#"                   # POWSPEC_SAVE_COMMENT, src/define.h:87


def _g10(x) -> str:             # OFMT_DBL = "%.10lg"
    return "%.10g" % float(x)


def _lg(x) -> str:              # "%lg"
    return "%g" % float(x)


def _count(arr) -> int:
    if arr is None:
        return 0
    if isinstance(arr, tuple):
        return int(arr[1])
    return int(arr.shape[0])


def _mesh_lines(mesh: Mesh) -> list:
    b, m = mesh.bsize, mesh.min
    return [f"{COMMENT} Box size: [{_lg(b[0])}, {_lg(b[1])}, {_lg(b[2])}]\n",
            f"{COMMENT} Box boundaries: [[{_lg(m[0])},{_lg(m[0] + b[0])}], "
            f"[{_lg(m[1])},{_lg(m[1] + b[1])}], [{_lg(m[2])},{_lg(m[2] + b[2])}]]\n",
            f"{COMMENT} Grid size: {int(mesh.Ng)} , assignment scheme: "
            f"{powspec_assign_names[int(mesh.assign)]} , grid interlacing: "
            f"{'enabled' if mesh.intlace else 'disabled'}\n"]


def _table(pk: PK, spectra) -> list:
    out = [f"{COMMENT} kcen(1) kmin(2) kmax(3) kavg(4) nmod(5)"
           + "".join(f" P_{int(p)}({l + 6})" for l, p in enumerate(pk.poles)) + "\n"]
    for i in range(pk.nbin):
        row = f"{_g10(pk.k[i])} {_g10(pk.kedge[i])} {_g10(pk.kedge[i + 1])} {_g10(pk.km[i])} {int(pk.cnt[i])}"
        out.append(row + "".join(" " + _g10(spectra[l][i]) for l in range(pk.nl)) + "\n")
    return out


def _write(path, lines):
    try:
        with open(path, "w") as f:
            f.writelines(lines)
    except OSError as ex:
        raise PowspecB200Error(f"cannot write to file: `{path}' ({ex.strerror})", POWSPEC_ERR_FILE)


def save_res(conf: Conf, cata: Cata, mesh: Mesh, pk: PK) -> None:
    """src/save_res.c:35-229.  Files: conf.oauto[n] for every catalogue with
    conf.isauto[n], conf.ocross if conf.iscross.  The catalogue sums are taken
    from `cata` (wdata / wrand; for host arrays of a simulation box the sum of
    the weights if wdata was left unset), shot noise and normalisation from `pk`
    (CATA.shot / CATA.norm as genr_mesh left them, src/genr_mesh.c:904-909)."""
    ndata = [_count(d) for d in cata.data]
    nrand = [_count(r) for r in cata.rand] if cata.rand else [0] * cata.num
    wdata = list(cata.wdata) if cata.wdata is not None else \
        [float(np.sum(np.asarray(d)[:, 3])) for d in cata.data]
    if conf.verbose:
        print("Saving outputs ...")
    for n in range(cata.num):
        if not conf.isauto[n]:
            continue
        if not conf.oauto or n >= len(conf.oauto) or not conf.oauto[n]:
            raise PowspecB200Error("OUTPUT_AUTO is not set", POWSPEC_ERR_FILE)
        lines = []
        if conf.oheader:
            lines.append(f"{COMMENT} Data catalog: {ndata[n]} objects, total weight: {_g10(wdata[n])}\n")
            if not conf.issim:
                lines.append(f"{COMMENT} Random catalog: {nrand[n]} objects, total weight: "
                             f"{_g10(cata.wrand[n])}\n")
            lines += _mesh_lines(mesh)
            shot = pk.shot[n] if conf.issim else pk.shot[n] / pk.norm[n]
            lines.append(f"{COMMENT} Shot noise: {_g10(shot)} , normalisation: {_g10(pk.norm[n])} \n")
        lines += _table(pk, pk.pl[n])
        _write(conf.oauto[n], lines)
        if conf.verbose:
            if cata.num == 2:
                print(f"  Auto power spectra for catalog {n + 1} saved to file: `{conf.oauto[n]}'")
            else:
                print(f"  Auto power spectra saved to file: `{conf.oauto[n]}'")
    if not conf.iscross:
        return
    if not conf.ocross:
        raise PowspecB200Error("OUTPUT_CROSS is not set", POWSPEC_ERR_FILE)
    lines = []
    if conf.oheader:
        lines.append(f"{COMMENT} Data catalog 1: {ndata[0]} objects, total weight: {_g10(wdata[0])}\n"
                     f"{COMMENT} Data catalog 2: {ndata[1]} objects, total weight: {_g10(wdata[1])}\n")
        if not conf.issim:
            lines.append(f"{COMMENT} Random catalog 1: {nrand[0]} objects, total weight: "
                         f"{_g10(cata.wrand[0])}\n{COMMENT} Random catalog 2: {nrand[1]} objects, "
                         f"total weight: {_g10(cata.wrand[1])}\n")
        lines += _mesh_lines(mesh)
        lines.append(f"{COMMENT} Normalisation factor 1: {_g10(pk.norm[0])} , "
                     f"Normalisation factor 2: {_g10(pk.norm[1])}\n")
    lines += _table(pk, pk.xpl)
    _write(conf.ocross, lines)
    if conf.verbose:
        print(f"  Cross power spectra saved to file: `{conf.ocross}'")


__all__ = ["save_res"]
