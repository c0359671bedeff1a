import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import orb_oracle
    orb_oracle.lib()  # builds oracle/liborb_oracle.so on first use
    return orb_oracle


@pytest.fixture(scope="session")
def reference():
    """The reference's own ORBextractor (oracle/_ref/liborbref.so, see oracle/orb_ref.py). Built from
    /root/reference in the development container; on the GPU box only the prebuilt binary exists."""
    from oracle import orb_ref
    if not orb_ref.available():
        orb_ref.build()
    if not orb_ref.available():
        pytest.skip("oracle/_ref/liborbref.so not built (needs /root/reference)")
    orb_ref.lib()
    return orb_ref


CONFIGS = {
    # name: (width, height, nfeatures)  -- SURVEY.md §8(a), Examples/*/*.yaml of the reference
    "tum1": (640, 480, 1000),
    "kitti": (1241, 376, 2000),
    "euroc": (752, 480, 1200),
}
