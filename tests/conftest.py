import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_e2e():
    return np.load(os.path.join(GOLDEN, "e2e.npz"))


@pytest.fixture(scope="session")
def golden_stages():
    return np.load(os.path.join(GOLDEN, "stages.npz"))


@pytest.fixture(scope="session")
def synth_sd():
    from cuahn_vio_b200 import synthetic as S
    return S.synthetic_state_dict(0)


def unpack_masks(g, i):
    """The four {0, 1/0.95} MC-dropout masks stored bit-packed in the golden file."""
    import torch
    shapes = [(16, 5120), (16, 256), (16, 5120), (16, 256)]
    out = []
    for j, shp in enumerate(shapes):
        bits = np.unpackbits(g[f"mask{j}_{i}"])[: shp[0] * shp[1]].reshape(shp)
        out.append(torch.from_numpy(bits.astype(np.float32)) * torch.tensor(1.0 / 0.95, dtype=torch.float32))
    return out
