import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def stage():
    return np.load(GOLDEN / "stage_vectors.npz")


@pytest.fixture(scope="session")
def solves():
    both = json.loads((GOLDEN / "solves.json").read_text())
    both.update(json.loads((GOLDEN / "solves_extra.json").read_text()))        # further parameter combinations (X*)
    return both


@pytest.fixture(scope="session")
def shipped():
    return json.loads((GOLDEN / "shipped.json").read_text())


@pytest.fixture(scope="session")
def workflow_golden():
    return json.loads((GOLDEN / "workflow.json").read_text())
