import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    from tests.problems import Golden
    return Golden()


@pytest.fixture(scope="session")
def batch_golden():
    from tests.problems import BatchGolden
    return BatchGolden()


@pytest.fixture(scope="session")
def traj_batch():
    """The trajectory problems of configs[2] / configs[3], built on demand and kept for the session: traj_batch(n) -> first n."""
    from ndtpso_slam_b200 import workload
    cache = []

    def get(n):
        if len(cache) < n:
            cache.extend(workload.cfg2_batch(n - len(cache), first=len(cache)))
        return cache[:n]

    return get


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind oracle/_ref (present in the build container, and on the
    GPU box as a prebuilt .so); tests that need it are skipped where it is missing."""
    from oracle import binding
    if not os.path.exists(binding.REF_SO) and not os.path.exists("/root/reference/lib/ndtpso_slam/core.cpp"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return binding.Reference()


@pytest.fixture(scope="session")
def ctx():
    """The CUDA path, through the C ABI.  Fails loudly without a device: there is no CPU fallback."""
    from ndtpso_slam_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
