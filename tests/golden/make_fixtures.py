#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the reference's own fixture files (run in the build container where
/root/reference is mounted; the GPU box has no /root/reference, so the JSON files are committed).

What is pinned and where it comes from:
  * events: parsed from /root/reference/data/*.txt with the line format of examples/import_handmade.py:30-45
    ("user,event,item"; `$set` lines are item properties, not events) and examples/import_movielens_eventserver.py
  * engine params: /root/reference/examples/handmade-engine.json, handmade-engine-item-sets.json
  * membership constraints: transcribed from data/integration-test-expected.txt and
    data/integration-test-item-set-expected.txt (line numbers in each entry)
  * llr_kats: six known-answer values of Mahout's LogLikelihoodTest
  * derived goldens (`indicators`): the oracle's output on those inputs at generation time -- they pin the oracle
    against regressions and are the vectors the CUDA path is compared with on the GPU box.
"""
import json
import os
import random
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def parse_events(path, delim=","):
    ev = []
    for line in open(path):
        d = line.rstrip("\r\n").split(delim)
        if len(d) < 3 or d[1] == "$set":
            continue
        ev.append([d[0], d[1], d[2]])
    return ev


def indicators_via_oracle(events, event_names, min_events, params):
    from oracle import oracle as orc
    from universal_recommender_b200 import preparator
    actions = [(n, [(u, i) for (u, e, i) in events if e == n]) for n in event_names]
    actions = [(n, p) for n, p in actions if p]
    prepared = preparator.prepare(actions, min_events)
    mats = [orc.Csr(d.n_rows, d.n_cols, d.row_ptr, d.col_idx) for _, d in prepared]
    res = orc.train(mats, [orc.Params(*p) for p in params[:len(mats)]], seed=1)
    a_items = prepared[0][1].column_ids.inverse
    out = {}
    for (name, d), r in zip(prepared, res):
        cols = d.column_ids.inverse
        table = {}
        for row in range(r.n_rows):
            c, l, k = r.row(row)
            table[a_items[row]] = [[cols[int(ci)], float(li), int(ki)] for ci, li, ki in zip(c, l, k)]
        out[name] = table
    return {"n_users": prepared[0][1].n_rows, "primary_items": list(a_items), "indicators": out}


def main():
    # ---- C1: handmade ------------------------------------------------------------------------------------------
    eng = json.load(open(f"{REF}/examples/handmade-engine.json"))
    ev = parse_events(f"{REF}/data/sample-handmade-data.txt")
    names = eng["datasource"]["params"]["eventNames"]
    min_ev = eng["datasource"]["params"]["minEventsPerUser"]
    ind = eng["algorithms"][0]["params"]["indicators"]
    params = [(i.get("maxItemsPerUser", 500), i.get("maxCorrelatorsPerItem", 50), i.get("minLLR")) for i in ind]
    fx = {
        "source": "data/sample-handmade-data.txt + examples/handmade-engine.json",
        "event_names": names, "min_events_per_user": min_ev, "params": params, "events": ev,
        # SURVEY.md Appendix B.1 (derived by hand from the data + constraints of integration-test-expected.txt)
        "survey_b1": {
            "n_users": 3,
            "col_a": {"Iphone 6": 1, "Iphone 5": 3, "Iphone 4": 2, "Ipad-retina": 1, "Galaxy": 3, "Nexus": 1},
            "llr_values": {"(1,0,0,2)": 3.819085009768877, "(1,1,0,1)": 1.046496287529096,
                           "(2,0,0,1)": 3.819085009768877, "(3,0,0,0)": 0.0},
            "purchase": {"Iphone 6": [["Ipad-retina", 3.819085009768877], ["Iphone 4", 1.046496287529096]],
                         "Ipad-retina": [["Iphone 6", 3.819085009768877], ["Iphone 4", 1.046496287529096]],
                         "Iphone 4": [["Ipad-retina", 1.046496287529096], ["Iphone 6", 1.046496287529096]],
                         "Nexus": [], "Iphone 5": [], "Galaxy": []},
            "view": {"Iphone 6": ["Soap"], "Ipad-retina": ["Soap"], "Iphone 4": ["Soap", "Tablets"], "Nexus": ["Tablets"],
                     "Iphone 5": [], "Galaxy": []},
            "category-pref": {"Iphone 6": ["tablets"], "Ipad-retina": ["tablets"], "Iphone 4": ["tablets"], "Nexus": ["tablets"],
                              "Iphone 5": [], "Galaxy": []},
        },
        # zero / non-zero pattern of data/integration-test-expected.txt that constrains the boundary
        "expected_file_constraints": [
            {"line": "48-50", "query": "item Galaxy", "means": "Galaxy has no correlators in any indicator (all scores 0.0)"},
            {"line": "52-54", "query": "item Surface", "means": "Surface is not a primary item (only u-3 bought it; u-3 is dropped by minEventsPerUser=3)"},
            {"line": "16", "query": "user u1", "means": "only Nexus scores > 0, via category-pref 'tablets'"},
        ],
    }
    fx["oracle"] = indicators_via_oracle(ev, names, min_ev, params)
    json.dump(fx, open(f"{HERE}/handmade.json", "w"), indent=1)

    # ---- item sets ----------------------------------------------------------------------------------------------
    eng = json.load(open(f"{REF}/examples/handmade-engine-item-sets.json"))
    ev = parse_events(f"{REF}/data/sample-handmade-item-set-data.txt")
    names = eng["datasource"]["params"]["eventNames"]
    ind = eng["algorithms"][0]["params"].get("indicators") or [{"name": n} for n in names]
    params = [(i.get("maxItemsPerUser", 500), i.get("maxCorrelatorsPerItem", 50), i.get("minLLR")) for i in ind]
    fx = {
        "source": "data/sample-handmade-item-set-data.txt + examples/handmade-engine-item-sets.json",
        "event_names": names, "min_events_per_user": eng["datasource"]["params"].get("minEventsPerUser"),
        "params": params, "events": ev,
        # data/integration-test-item-set-expected.txt: an itemSet query returns exactly the items whose `purchase`
        # indicator contains a query item, minus the query items (URAlgorithm.scala:640-646, 756-765)
        "membership": [
            {"line": 16, "query": ["iPhone 6"], "hits": ["iPhone earbuds", "iPhone 6 charging cradle", "iPhone 6 case"]},
            {"line": 20, "query": ["iPhone 7"], "hits": ["AirPods"]},
            {"line": 24, "query": ["iPhone 6p"], "hits": []},
            {"line": 28, "query": ["AirPods"], "hits": ["iPhone 7"]},
            {"line": 32, "query": ["USB type-C cable"], "hits": ["Nexus 6p case", "Nexus 6p"]},
            {"line": 36, "query": ["iPhone 6 charging cradle"], "hits": ["iPhone earbuds", "iPhone 6", "iPhone 6 case"]},
            {"line": 40, "query": ["iPhone earbuds", "iPhone 6 case"], "hits": ["iPhone 6", "iPhone 6 charging cradle"]},
        ],
        # SURVEY.md Appendix B.2
        "survey_b2": {"n_users": 8, "llr": {"AirPods|iPhone 7": 10.58501181052771, "Nexus 6p|Nexus 6p case": 10.58501181052771,
                                            "Nexus 6p|USB type-C cable": 5.1782773041320524, "iPhone 6|iPhone 6 case": 8.997362313900929,
                                            "iPhone 6|iPhone 6 charging cradle": 3.2557338578632056,
                                            "iPhone 6 charging cradle|iPhone earbuds": 6.028322580102987}},
    }
    fx["oracle"] = indicators_via_oracle(ev, names, fx["min_events_per_user"], params)
    json.dump(fx, open(f"{HERE}/item_sets.json", "w"), indent=1)

    # ---- movielens sample (no expected output in the reference; oracle-vs-GPU only) ------------------------------
    random.seed(3)  # examples/import_movielens_eventserver.py:10,16
    ev = []
    for line in open(f"{REF}/data/sample_movielens_data.txt"):
        d = line.rstrip("\r\n").split("::")
        ev.append([d[0], "rate" if random.randint(0, 1) == 1 else "buy", d[1]])
        random.randint(0, 1)  # the importer draws a second number per line for the $set category
    fx = {"source": "data/sample_movielens_data.txt labelled like examples/import_movielens_eventserver.py:21-36",
          "event_names": ["rate", "buy"], "min_events_per_user": None, "params": [(500, 50, None), (500, 50, None)], "events": ev}
    fx["oracle"] = indicators_via_oracle(ev, fx["event_names"], None, fx["params"])
    json.dump(fx, open(f"{HERE}/movielens_sample.json", "w"))

    # ---- LLR known answers ------------------------------------------------------------------------------------------
    json.dump({"source": "Mahout LogLikelihoodTest values (SURVEY.md A.3), 6 printed digits",
               "kats": [[1, 0, 0, 1, 2.772589], [10, 0, 0, 10, 27.72589], [5, 1995, 0, 100000, 39.33052],
                        [1000, 1995, 1000, 100000, 4730.737], [1000, 1000, 1000, 100000, 5734.343],
                        [1000, 1000, 1000, 99000, 5714.932]]}, open(f"{HERE}/llr_kats.json", "w"), indent=1)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
