"""SURVEY.md 8f-2, the output formatter row: the Elasticsearch bulk body of the indicator model.
CPU: the restatement (oracle/format_oracle.py) against the host mirror of toStringMapRDD on the reference fixtures.
GPU: cco_format_es_bulk byte for byte against the restatement (fixtures, hostile id strings, a rank's row slice)."""
import json

import numpy as np
import pytest

from conftest import load_golden, prepared_from_fixture


def _docs(body: bytes):
    lines = body.decode("utf-8").split("\n")
    assert lines[-1] == "" and len(lines) % 2 == 1
    return [(json.loads(lines[i]), json.loads(lines[i + 1])) for i in range(0, len(lines) - 1, 2)]


@pytest.mark.parametrize("name", ["handmade.json", "item_sets.json"])
def test_oracle_bulk_body_matches_the_string_map_of_the_reference_consumer(orc, name):
    from oracle import format_oracle as fo
    fx = load_golden(name)
    prepared = prepared_from_fixture(fx)
    mats = [orc.Csr(d.n_rows, d.n_cols, d.row_ptr, d.col_idx) for _, d in prepared]
    ref = orc.train(mats, [orc.Params(*p) for p in fx["params"]], 1)
    names = [ev for ev, _ in prepared]
    a = prepared[0][1]
    body = fo.es_bulk([(r.row_ptr, r.col_idx) for r in ref], names, a.column_ids.inverse, [d.column_ids.inverse for _, d in prepared])
    docs = _docs(body)
    assert len(docs) == a.n_cols
    for r, (action, doc) in enumerate(docs):
        item = a.column_ids.inverse[r]
        assert action == {"index": {"_id": item}} and doc["id"] == item
        for ev, want in fx["oracle"]["indicators"].items():
            assert doc[ev] == [x[0] for x in want[item]]          # the ordered id lists of SURVEY Appendix B
    # same thing through the host mirror of toStringMapRDD (package.scala:82-110)
    for (ev, d), ind in zip(prepared, ref):
        sm = a.create(ind.row_ptr, ind.col_idx, a.column_ids, d.column_ids, ind.llr, ind.count).to_string_map(ev)
        for action, doc in docs:
            assert doc[ev] == sm[doc["id"]][ev]


def test_escaping_rules():
    from oracle import format_oracle as fo
    assert fo.json_escape('a"b\\c') == b'a\\"b\\\\c'
    assert fo.json_escape("tab\there\n") == b"tab\\u0009here\\u000a"
    assert fo.json_escape("Zoë ☃") == "Zoë ☃".encode("utf-8")
    for s in ['a"b\\c', "tab\there\n", "Zoë ☃", "", "\x00\x1f "]:
        assert json.loads(b'"' + fo.json_escape(s) + b'"') == s


# ---- device ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["handmade.json", "item_sets.json", "movielens_sample.json"])
def test_device_bulk_body_on_the_reference_fixtures(orc, ctx, name):
    from oracle import format_oracle as fo
    fx = load_golden(name)
    prepared = prepared_from_fixture(fx)
    mats = [(d.n_rows, d.n_cols, d.row_ptr, d.col_idx) for _, d in prepared]
    names = [ev for ev, _ in prepared]
    a = prepared[0][1]
    res, h = ctx.train_csr(mats, fx["params"], 1, keep=True)
    try:
        got = ctx.format_es_bulk(h, names, a.column_ids.inverse, [d.column_ids.inverse for _, d in prepared])
        want = fo.es_bulk([(r[3], r[4]) for r in res], names, a.column_ids.inverse, [d.column_ids.inverse for _, d in prepared])
    finally:
        ctx.free_result(h)
    assert got == want
    assert all(doc["id"] == act["index"]["_id"] for act, doc in _docs(got))


@pytest.mark.gpu
def test_device_bulk_body_with_hostile_ids_and_many_rows(orc, ctx):
    import synth
    from oracle import format_oracle as fo
    import universal_recommender_b200 as ur
    w = synth.make("small")
    rng = np.random.default_rng(3)
    alphabet = ['"', "\\", "\t", "\n", "\x01", "é", "☃", "\U0001f600", "a", "B", "7", " ", "-", "/"]

    def ids(n, salt):
        out = []
        for i in range(n):
            k = int(rng.integers(0, 9))
            out.append("".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), k)) + f"{salt}{i}")
        out[0] = ""                                   # an empty id is legal JSON
        return out
    row_ids = ids(w.n_items, "i")
    col_ids = [row_ids] + [ids(w.n_items, f"t{t}_") for t in range(1, w.n_types)]
    names = ["purchase", 'vi"ew', "category-pref"]
    res, h = ctx.train_csr(w.mats, w.params, 5, flags=ur.FLAG_RESULT_NO_COUNT | ur.FLAG_RESULT_NO_LLR, keep=True)
    try:
        got = ctx.format_es_bulk(h, names, row_ids, col_ids)
        want = fo.es_bulk([(r[3], r[4]) for r in res], names, row_ids, col_ids)
    finally:
        ctx.free_result(h)
    assert got == want
    docs = _docs(got)
    assert len(docs) == w.n_items and docs[7][1]["id"] == row_ids[7]
    assert max(len(d['vi"ew']) for _, d in docs) == 50
