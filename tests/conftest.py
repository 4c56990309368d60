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


def oracle_threads():
    """Thread count for the CPU oracle inside GPU tests: what this process can really run on (affinity mask, cgroup quota),
    at most 32.  A pool box shows 128 CPUs to a container that may own a fraction of them; 128 OpenMP threads there made a
    full-size oracle frame 12x slower than 8 threads (bench.py::usable_cores has the same logic)."""
    n = os.cpu_count() or 8
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for quota_path, period_path in (("/sys/fs/cgroup/cpu.max", None), ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us")):
        try:
            if period_path is None:
                quota, period = open(quota_path).read().split()[:2]
                if quota == "max":
                    continue
                quota, period = int(quota), int(period)
            else:
                quota, period = int(open(quota_path).read()), int(open(period_path).read())
            if quota > 0 and period > 0:
                n = min(n, max(1, -(-quota // period)))
        except (OSError, ValueError):
            pass
    return max(1, min(n, 32))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
