import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "liberate-fhe_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_contexts():
    return json.loads((GOLDEN / "context.json").read_text())


def primes_for(logN, n_small=2, n_big=2):
    """a few scale primes + 60-bit primes of the reference's cached tables for this logN"""
    ctxs = json.loads((GOLDEN / "context.json").read_text())["contexts"]
    for c in ctxs:
        if c["args"]["logN"] == logN:
            q = c["q"]
            K = c["args"]["num_special_primes"]
            small = q[:n_small]
            big = q[len(q) - K - 1:][:n_big]
            return small + big
    raise KeyError(logN)
