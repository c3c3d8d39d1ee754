"""Generate the golden parity fixtures from the UNMODIFIED reference.

Needs oracle/_ref/libpowspec_ref{,_f32}.so, i.e. must run in the container
where /root/reference exists (`make -C oracle ref` first).  Writes
tests/golden/golden_inputs.npz and tests/golden/golden_outputs.json.

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

(one thread: makes the reference's atomics / thread-private sums run in a
fixed order, so regenerating gives byte-identical fixtures).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import load_oracle  # noqa: E402
from tests.golden.cases import CASES, SINGLE_CASES, make_inputs, run_case  # noqa: E402


def pack(res):
    out = dict(nbin=res.nbin, nl=res.nl, k=res.k.tolist(), kedge=res.kedge.tolist(),
               km=res.km.tolist(), cnt=[int(c) for c in res.cnt],
               lcnt=res.lcnt.tolist(), shot=res.shot.tolist(), norm=res.norm.tolist(),
               bmin=res.bmin.tolist(), bsize=res.bsize.tolist(),
               pl=[None if p is None else p.tolist() for p in res.pl],
               xpl=None if res.xpl is None else res.xpl.tolist())
    return out


def main():
    inp = make_inputs()
    np.savez_compressed(os.path.join(HERE, "golden_inputs.npz"), **inp)
    ref = load_oracle("ref")
    ref32 = load_oracle("ref", single=True)
    out = {"_generator": ref.backend, "_generator_f32": ref32.backend, "double": {}, "single": {}}
    for case in CASES:
        out["double"][case["name"]] = pack(run_case(ref, case, inp))
        if case["name"] in SINGLE_CASES:
            out["single"][case["name"]] = pack(run_case(ref32, case, inp))
        print("generated", case["name"])
    with open(os.path.join(HERE, "golden_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
