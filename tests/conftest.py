import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


@pytest.fixture(scope="session")
def golden_indices():
    """name -> (x float32 [n,d], metric) for the committed reference indices."""
    from oracle import oracle as O
    d = os.path.join(GOLD, "indices")
    return {f: O.read_faiss_flat(os.path.join(d, f)) for f in sorted(os.listdir(d))}


@pytest.fixture(scope="session")
def golden_texts():
    import json
    g = json.load(open(os.path.join(GOLD, "phase4_records.json")))
    tg = json.load(open(os.path.join(GOLD, "tfidf_golden.json")))
    return g["chunks"], tg["queries"]
