import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_inputs():
    import numpy as np
    z = np.load(os.path.join(HERE, "golden", "inputs.npz"))
    return {k: (z[k], int(z[k + "_rate"])) for k in ("tapestry16k", "tapestry22k", "negative24k")}


@pytest.fixture(scope="session")
def golden_outputs():
    import numpy as np
    z = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases
