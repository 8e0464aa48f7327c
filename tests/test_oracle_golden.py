"""The oracle against everything the reference's fixtures pin and against the
committed derived goldens."""
import numpy as np
import pytest

from conftest import load_golden, prepared_from_fixture


def run_oracle(orc, fx, seed=1):
    prepared = prepared_from_fixture(fx)
    mats = [orc.Csr(d.n_rows, d.n_cols, d.row_ptr, d.col_idx) for _, d in prepared]
    res = orc.train(mats, [orc.Params(*p) for p in fx["params"][:len(mats)]], seed)
    out = {}
    a_items = prepared[0][1].column_ids.inverse
    for (name, d), r in zip(prepared, res):
        cols = d.column_ids.inverse
        out[name] = {a_items[row]: [(cols[int(c)], float(l), int(k)) for c, l, k in zip(*r.row(row))] for row in range(r.n_rows)}
    return prepared, out


def test_handmade_matches_survey_appendix_b1(orc):
    fx = load_golden("handmade.json")
    prepared, got = run_oracle(orc, fx)
    b1 = fx["survey_b1"]
    a = prepared[0][1]
    assert a.n_rows == b1["n_users"] == 3                       # u-3 dropped by minEventsPerUser (duplicates count)
    assert "Surface" not in a.column_ids                          # only u-3 bought it
    col_a = np.bincount(a.col_idx, minlength=a.n_cols)
    assert {a.column_ids.inverse[i]: int(c) for i, c in enumerate(col_a)} == b1["col_a"]
    for item, want in b1["purchase"].items():
        have = got["purchase"][item]
        assert sorted(c for c, _, _ in have) == sorted(c for c, _ in want)
        for (c, l, _), (wc, wl) in zip(sorted(have), sorted(want)):
            assert c == wc and l == pytest.approx(wl, rel=1e-14)
    for name in ("view", "category-pref"):
        for item, want in b1[name].items():
            assert sorted(c for c, _, _ in got[name][item]) == sorted(want), (name, item)
    # Galaxy / Iphone 5 were bought by every user: every LLR is exactly 0 -> no correlators anywhere
    for name in got:
        assert got[name]["Galaxy"] == [] and got[name]["Iphone 5"] == []


def test_item_sets_membership_constraints(orc):
    fx = load_golden("item_sets.json")
    prepared, got = run_oracle(orc, fx)
    assert prepared[0][1].n_rows == fx["survey_b2"]["n_users"]
    purchase = got["purchase"]
    for c in fx["membership"]:
        hits = {item for item, corr in purchase.items() if any(x in {cc for cc, _, _ in corr} for x in c["query"])} - set(c["query"])
        assert hits == set(c["hits"]), f"integration-test-item-set-expected.txt:{c['line']}"
    for pair, want in fx["survey_b2"]["llr"].items():
        a, b = pair.split("|")
        assert dict((c, l) for c, l, _ in purchase[a])[b] == pytest.approx(want, rel=1e-14)
        assert dict((c, l) for c, l, _ in purchase[b])[a] == pytest.approx(want, rel=1e-14)


@pytest.mark.parametrize("name", ["handmade.json", "item_sets.json", "movielens_sample.json"])
def test_oracle_reproduces_committed_goldens(orc, name):
    fx = load_golden(name)
    _, got = run_oracle(orc, fx)
    want = fx["oracle"]["indicators"]
    assert set(got) == set(want)
    for ev in want:
        for item, rows in want[ev].items():
            have = got[ev][item]
            assert [c for c, _, _ in have] == [r[0] for r in rows], (ev, item)
            assert [k for _, _, k in have] == [r[2] for r in rows]
            assert np.allclose([l for _, l, _ in have], [r[1] for r in rows], rtol=1e-12, atol=0)


def test_rows_sorted_and_positive(orc):
    fx = load_golden("movielens_sample.json")
    _, got = run_oracle(orc, fx)
    for ev in got:
        for item, rows in got[ev].items():
            llrs = [l for _, l, _ in rows]
            assert all(l > 0 for l in llrs)
            assert llrs == sorted(llrs, reverse=True)
            assert len(rows) <= 50
