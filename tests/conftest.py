import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """Measured parity errors of the GPU tests -> gpurun_out/parity.json (copied to profiles/ per round)."""
    try:
        from tests import _util
    except Exception:  # noqa: BLE001
        return
    if not _util.PARITY:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity.json")
    old = {}
    if os.path.exists(path):
        try:
            with open(path) as fh:
                old = json.load(fh)
        except Exception:  # noqa: BLE001
            old = {}
    old.update(_util.PARITY)
    with open(path, "w") as fh:
        json.dump(old, fh, indent=1, sort_keys=True)
