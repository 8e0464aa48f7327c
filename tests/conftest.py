import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """One CcoContext on cuda:0 through the C ABI.  Fails loudly (no skip) if the extension or the GPU is missing."""
    import universal_recommender_b200 as ur
    c = ur.CcoContext(device=0)
    yield c
    c.close()


def prepared_from_fixture(fx):
    """events of a golden fixture -> [(event name, IndexedDataset)] through the Preparator mirror."""
    from universal_recommender_b200 import preparator
    actions = [(n, [(u, i) for (u, e, i) in fx["events"] if e == n]) for n in fx["event_names"]]
    actions = [(n, p) for n, p in actions if p]   # DataSource.scala:79-89 drops empty event RDDs
    return preparator.prepare(actions, fx.get("min_events_per_user"))
