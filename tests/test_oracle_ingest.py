"""SURVEY.md 8f-1 (the next row): the oracle's integer-id ingest against the Preparator mirror on the reference's
fixtures and on random events."""
import random

import numpy as np
import pytest

from conftest import load_golden
from universal_recommender_b200 import preparator


def tokenise(actions):
    users, items = {}, [dict() for _ in actions]
    ev = []
    for t, (_, pairs) in enumerate(actions):
        u = np.array([users.setdefault(a, len(users)) for a, _ in pairs], dtype=np.int64)
        i = np.array([items[t].setdefault(b, len(items[t])) for _, b in pairs], dtype=np.int32)
        ev.append((u, i))
    return users, items, [(u, i, len(items[t])) for t, (u, i) in enumerate(ev)]


def as_pairs(csr, user_names, item_names):
    return {(user_names[r], item_names[c]) for r in range(csr.n_rows) for c in csr.col_idx[csr.row_ptr[r]:csr.row_ptr[r + 1]]}


def check_against_preparator(orc, actions, min_ev):
    users, items, ev = tokenise(actions)
    user_map, res = orc.ingest(ev, len(users), min_ev or 0)
    prepared = preparator.prepare(actions, min_ev)
    inv_u = {new: name for name, raw in users.items() for new in [int(user_map[raw])] if new >= 0}
    assert set(inv_u.values()) == set(prepared[0][1].row_ids.inverse)
    for t, ((name, d), (csr, item_map)) in enumerate(zip(prepared, res)):
        inv_i = {int(item_map[raw]): nm for nm, raw in items[t].items() if item_map[raw] >= 0}
        assert csr.n_rows == d.n_rows and csr.n_cols == d.n_cols, name
        assert set(inv_i.values()) == set(d.column_ids.inverse)
        want = {(d.row_ids.inverse[r], d.column_ids.inverse[c]) for r in range(d.n_rows) for c in d.col_idx[d.row_ptr[r]:d.row_ptr[r + 1]]}
        assert as_pairs(csr, inv_u, inv_i) == want
        for r in range(csr.n_rows):                                   # canonical rows
            row = csr.col_idx[csr.row_ptr[r]:csr.row_ptr[r + 1]]
            assert (np.diff(row) > 0).all()


@pytest.mark.parametrize("name", ["handmade.json", "item_sets.json", "movielens_sample.json"])
def test_ingest_matches_preparator_on_fixtures(orc, name):
    fx = load_golden(name)
    actions = [(n, [(u, i) for (u, e, i) in fx["events"] if e == n]) for n in fx["event_names"]]
    actions = [(n, p) for n, p in actions if p]
    check_against_preparator(orc, actions, fx.get("min_events_per_user"))


def test_ingest_matches_preparator_on_random_events(orc):
    rng = random.Random(4)
    for _ in range(60):
        users = [f"u{i}" for i in range(rng.randrange(1, 15))]
        actions = []
        for t in range(rng.randrange(1, 4)):
            items = [f"t{t}i{i}" for i in range(rng.randrange(1, 10))]
            actions.append((f"e{t}", [(rng.choice(users), rng.choice(items)) for _ in range(rng.randrange(0 if t else 1, 50))]))
        check_against_preparator(orc, actions, rng.choice([None, 2, 4]))


def test_ingest_rejects_bad_ids(orc):
    with pytest.raises(orc.OracleError):
        orc.ingest([(np.array([0, 5]), np.array([0, 0], dtype=np.int32), 1)], 3)
    with pytest.raises(orc.OracleError):
        orc.ingest([(np.array([0, 1]), np.array([0, 2], dtype=np.int32), 2)], 3)
