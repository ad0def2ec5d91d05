import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_states():
    return load_golden("feature_states.json.gz")


@pytest.fixture(scope="session")
def golden_steps():
    return load_golden("steps.json")


@pytest.fixture(scope="session")
def golden_games():
    return load_golden("games.json")


@pytest.fixture(scope="session")
def golden_ce():
    return load_golden("ce.json")
