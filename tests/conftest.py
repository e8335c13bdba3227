import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).parent.parent.resolve()
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped, not failed, on a machine without an NVIDIA driver (plain
    `pytest` on a CPU-only host).  On a GPU box nothing is skipped here: a missing or stale
    library must fail loudly there."""
    from stencil_benchmarks_b200 import capi

    if capi.driver_present():
        return
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(pytest.mark.skip(reason="no NVIDIA driver on this machine"))


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
