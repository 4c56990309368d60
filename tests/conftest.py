import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def nerr(a, b):
    """Normalised max error: max|a-b| / max|b|  (the '1e-3 relative fp32' metric of BASELINE.json)."""
    import torch
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    return d / s if s > 0 else d


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
