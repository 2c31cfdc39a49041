import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` via gpurun)")


def golden_names(prefix):
    return sorted(
        os.path.basename(p)[: -len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "__*.npz"))
    )


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible (no CPU fallback exists)")
    return torch.device("cuda:0")
