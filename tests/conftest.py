"""Shared fixtures.  `-m "not gpu"` = oracle vs golden vectors + host logic + ABI checks (CPU only);
`-m gpu` = parity of the CUDA path against the oracle, through the C-ABI (needs a B200)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as O

    O.build()
    return O.Oracle()


@pytest.fixture(scope="session")
def kats():
    return json.load(open(os.path.join(GOLDEN, "kats.json")))


def _num(v):
    if isinstance(v, list):
        return [_num(x) for x in v]
    return int(v)


@pytest.fixture(scope="session")
def kat(kats):
    """kat(file, test) -> (lets, asserts) with every number as int (u256 are stored as strings)."""

    def get(path, name):
        t = kats[path][name]
        lets = {k: _num(v) for k, v in t["let"].items()}
        asserts = [dict(a, rhs=_num(a["rhs"]) if not (isinstance(a["rhs"], str) and not a["rhs"].isdigit()) else a["rhs"]) for a in t["assert"]]
        return lets, asserts

    return get


def flat(v):
    """Flatten nested QM31 / point tuples to a list of ints."""
    if isinstance(v, (list, tuple)):
        out = []
        for x in v:
            out += flat(x)
        return out
    return [int(v)]
