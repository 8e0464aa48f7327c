"""Algorithm models behind DESIGN.md section 8 (round-2 plan), validated against brute force with the oracle's LLR:
the integer-domain top-k merge and the simpler level-1 integer cut are EXACT (same kept cells, same order)."""
import os
import random
import sys

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "proto"))


def test_integer_domain_selection_models_are_exact(orc):
    import cut_model
    import select_model
    rng = random.Random(3)
    stats_a, stats_b = {}, {}
    for _ in range(400):
        cells, ra, n = select_model.random_row(rng)
        k = rng.choice([1, 5, 50, 200])
        min_llr = rng.choice([None, None, 0.5])
        item, self_ = rng.randrange(0, 5000), rng.random() < 0.3
        want = select_model.brute(cells, ra, n, k, min_llr, item, self_)
        assert select_model.integer_domain(cells, ra, n, k, min_llr, item, self_, stats=stats_a) == want
        assert cut_model.level1_cut(cells, ra, n, k, min_llr, item, self_, stats=stats_b) == want
    assert stats_a["evals"] <= stats_a["cells"] and stats_b["evals"] <= stats_b["cells"]
