import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the oracle libraries use OpenMP; keep the CPU suite polite and deterministic
os.environ.setdefault("OMP_NUM_THREADS", "4")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_inputs():
    with np.load(os.path.join(GOLDEN_DIR, "golden_inputs.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_outputs():
    with open(os.path.join(GOLDEN_DIR, "golden_outputs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import load_oracle
    return load_oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import have_ref, load_oracle
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return load_oracle("ref")
