"""Seeded inputs of the coordinate-conversion parity cases (shared by the golden
generator and the tests)."""
import numpy as np

CNVT_CASES = [
    dict(name="flat_lcdm", seed=101, n=(4000,), zrange=(0.4, 1.1),
         cosmo=dict(omega_m=0.31, omega_l=0.69, omega_k=0.0, eos_w=-1.0, ecdst=1e-8)),
    dict(name="two_catalogues_lowz", seed=102, n=(3000, 2500), zrange=(0.0, 0.3),
         cosmo=dict(omega_m=0.27, omega_l=0.73, omega_k=0.0, eos_w=-1.0, ecdst=1e-10)),
    dict(name="curved_wcdm", seed=103, n=(3500,), zrange=(0.8, 3.5),
         cosmo=dict(omega_m=0.30, omega_l=0.65, omega_k=0.05, eos_w=-0.9, ecdst=1e-9)),
    dict(name="loose_error", seed=104, n=(2000,), zrange=(0.1, 2.0),
         cosmo=dict(omega_m=0.31, omega_l=0.69, omega_k=0.0, eos_w=-1.0, ecdst=1e-4)),
    dict(name="table_spline", seed=105, n=(4000,), zrange=(0.2, 1.4), table=True,
         cosmo=dict(omega_m=0.31, omega_l=0.69, omega_k=0.0, eos_w=-1.0, ecdst=1e-8)),
]


def cnvt_inputs(case):
    r = np.random.default_rng(case["seed"])
    out = []
    for n in case["n"]:
        a = np.empty((n, 4))
        a[:, 0] = r.uniform(0.0, 360.0, n)
        a[:, 1] = np.rad2deg(np.arcsin(r.uniform(-1.0, 1.0, n)))
        a[:, 2] = r.uniform(*case["zrange"], n)
        a[:, 3] = r.uniform(0.5, 1.5, n)
        out.append(a)
    return out


def distance_table(case, nsp=400):
    """(z, d) samples of a smooth distance-redshift relation covering the case's range
    (trapezoid rule on a fine grid: the table's provenance does not matter, both
    sides interpolate the same numbers)."""
    c = case["cosmo"]
    z = np.linspace(0.0, case["zrange"][1] * 1.05 + 0.01, nsp)
    fine = np.linspace(0.0, z[-1], 200001)
    e = np.sqrt(c["omega_m"] * (1 + fine) ** 3 + c["omega_k"] * (1 + fine) ** 2 + c["omega_l"])
    integ = np.concatenate([[0.0], np.cumsum(0.5 * (1 / e[1:] + 1 / e[:-1]) * np.diff(fine))])
    d = 2997.92458 * np.interp(z, fine, integ)
    return z, d
