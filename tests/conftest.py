import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference checkout (build container)")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "cirkit"))


@pytest.fixture(scope="session")
def reference():
    """Make the real reference importable (build container only; never on the GPU box)."""
    if not have_reference():
        pytest.skip("reference checkout not present")
    if REFERENCE not in sys.path:
        sys.path.insert(1, REFERENCE)
    import cirkit  # noqa: F401

    return cirkit
