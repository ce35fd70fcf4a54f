import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EXTDATA = os.path.join(GOLDEN, "extdata")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a CUDA device and the built library: skip them (loudly) elsewhere, so that a plain
    `pytest tests/` is green on a CPU-only machine; `-m gpu` on the GPU box runs them all."""
    reason = None
    try:
        from gwsem_b200 import _lib
        if _lib.lib().gwsem_device_count() <= 0:
            reason = "no CUDA device"
    except Exception as exc:   # noqa: BLE001
        reason = f"libgwsem_b200.so not usable: {exc}"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    from gwsem_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()
