import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    return oracle()


@pytest.fixture(scope="session")
def golden():
    table = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    return {e["name"]: e for e in table}


@pytest.fixture(scope="session")
def ref_available():
    from oracle import ref_binary
    return ref_binary() is not None
