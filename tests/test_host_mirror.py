"""Host-side mirror of the reference interface: Preparator semantics, IndexedDataset conversions and the
URAlgorithm.calcAll parameter plumbing (no GPU: the native call is stubbed)."""
import numpy as np
import pytest

import universal_recommender_b200 as ur
from universal_recommender_b200 import preparator, similarity_analysis, ur_algorithm
from conftest import load_golden, prepared_from_fixture


def test_preparator_min_events_counts_duplicates_and_freezes_users():
    fx = load_golden("handmade.json")
    prepared = prepared_from_fixture(fx)
    names = [n for n, _ in prepared]
    assert names == ["purchase", "view", "category-pref"]
    a = prepared[0][1]
    # u-3 has 2 purchase events (< 3) and is dropped; u-4 passes only because duplicates count (Preparator.scala:129-132)
    assert sorted(a.row_ids.inverse) == sorted(["u1", "U 2", "u-4"])
    for _, d in prepared[1:]:
        assert d.row_ids is a.row_ids and d.n_rows == a.n_rows          # one shared user dictionary
    # binary + deduplicated + sorted rows
    for _, d in prepared:
        for r in range(d.n_rows):
            cols = d.col_idx[d.row_ptr[r]:d.row_ptr[r + 1]]
            assert list(cols) == sorted(set(cols))


def test_preparator_secondary_events_of_unknown_users_are_dropped():
    actions = [("buy", [("a", "x"), ("b", "y")]), ("view", [("a", "p"), ("zzz", "q"), ("b", "p")])]
    prepared = preparator.prepare(actions)
    view = prepared[1][1]
    assert view.n_rows == 2 and view.column_ids.inverse == ["p"]           # 'q' only seen by the unknown user
    assert view.nnz == 2


def test_to_string_map_orders_by_llr_and_drops_scores():
    rows = ur.BiDictionary(["i0", "i1"])
    cols = ur.BiDictionary(["c0", "c1", "c2"])
    ids = ur.IndexedDataset(np.array([0, 3, 3]), np.array([2, 0, 1], dtype=np.int32), rows, cols,
                            values=np.array([1.0, 5.0, 5.0]))
    m = ids.to_string_map("buy")
    assert m == {"i0": {"buy": ["c0", "c1", "c2"]}, "i1": {"buy": []}}      # stable: ties keep input order (package.scala:100-108)


class _FakeCtx:
    world_size, rank = 1, 0

    def __init__(self):
        self.calls = []

    def train_csr(self, mats, params, seed, flags=0, copy_arrays=True):
        self.calls.append((params, seed, flags))
        return [(0, m[1], m[1], np.zeros(mats[0][1] + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.int32))
                for m in mats]


def _actions():
    return preparator.prepare([("buy", [("a", "x"), ("b", "y")]), ("view", [("a", "p"), ("b", "q")])])


def test_calc_all_global_params_path():
    ctx = _FakeCtx()
    out = ur_algorithm.calc_all(_actions(), ur.URAlgorithmParams(maxEventsPerEventType=77, maxCorrelatorsPerEventType=9, seed=5), ctx=ctx)
    params, seed, _ = ctx.calls[0]
    assert params == [(77, 9, None), (77, 9, None)] and seed == 5          # URAlgorithm.scala:323-329
    assert [n for n, _ in out] == ["buy", "view"]                           # names zipped back positionally (:349)
    assert out[1][1].row_ids is out[0][1].row_ids                            # rows = primary items for every indicator


def test_calc_all_per_indicator_path_and_defaults():
    ctx = _FakeCtx()
    ap = ur.URAlgorithmParams(indicators=[ur.IndicatorParams("buy"), ur.IndicatorParams("view", 20, 5, 0.5)], seed=1)
    ur_algorithm.calc_all(_actions(), ap, ctx=ctx)
    assert ctx.calls[0][0] == [(500, 50, None), (20, 5, 0.5)]               # defaults 500 / 50 / None (:336-340)


def test_calc_all_rejects_bad_recs_model_and_skips_backfill():
    with pytest.raises(ValueError):
        ur_algorithm.calc_all(_actions(), ur.URAlgorithmParams(recsModel="nope"), ctx=_FakeCtx())
    assert ur_algorithm.calc_all(_actions(), ur.URAlgorithmParams(recsModel="backfill"), ctx=_FakeCtx()) == []


def test_seed_to_int_wraps_like_scala():
    assert similarity_analysis._to_i32(0xdeadbeef) == -559038737
    assert similarity_analysis._to_i32(1 << 40 | 7) == 7


def test_engine_json_params():
    ap = ur.URAlgorithmParams.from_engine_json({"indicators": [{"name": "purchase"}, {"name": "view", "maxCorrelatorsPerItem": 50}], "seed": 3})
    assert ap.indicators[1].maxCorrelatorsPerItem == 50 and ap.indicators[0].maxItemsPerUser is None and ap.seed == 3


def test_preparator_matches_bruteforce_semantics():
    """Randomised: shared user dictionary, secondary events of unknown users dropped, duplicates collapse, minEventsPerUser
    counts duplicate primary events (Preparator.scala:44-87, 102-214)."""
    import random
    rng = random.Random(2)
    for _ in range(50):
        users = [f"u{i}" for i in range(rng.randrange(1, 12))]
        items = [f"i{i}" for i in range(rng.randrange(1, 9))]
        ev = {n: [(rng.choice(users + ["ghost"]), rng.choice(items)) for _ in range(rng.randrange(0, 40))] for n in ("buy", "view")}
        if not ev["buy"]:
            continue
        min_ev = rng.choice([None, 2, 3])
        prepared = preparator.prepare([("buy", ev["buy"]), ("view", ev["view"])], min_ev)
        counts = {}
        for u, _ in ev["buy"]:
            counts[u] = counts.get(u, 0) + 1
        keep = {u for u, c in counts.items() if min_ev is None or c >= min_ev}
        a, b = prepared[0][1], prepared[1][1]
        assert set(a.row_ids.inverse) == keep and b.row_ids is a.row_ids
        for name, d in prepared:
            want = {(u, i) for u, i in ev[name] if u in keep}
            got = {(d.row_ids.inverse[r], d.column_ids.inverse[c]) for r in range(d.n_rows)
                   for c in d.col_idx[d.row_ptr[r]:d.row_ptr[r + 1]]}
            assert got == want
            assert set(d.column_ids.inverse) == {i for _, i in want}      # item ids come from surviving events only
