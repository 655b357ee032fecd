import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fx = {"meta": z["meta"]}
    for pre in ("param", "in", "grad", "after"):
        fx[pre] = {k[len(pre) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre + ".")}
    for k in z.files:
        if "." not in k and k != "meta":
            fx[k] = z[k]
    return fx


@pytest.fixture
def golden():
    return load_golden
