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


def test_entry_parallel_sampler_walk_model():
    """the chunk / row-window walk of k_sample_count attributes every stored entry to its row, decides it
    once, and its lane-range masks + window flushes give the per-row kept counts -- empty rows, rows of thousands of
    entries, blocks that do not start at entry 0 and chunk-straddling rows included."""
    import numpy as np
    import sampler_walk_model as m
    rng = np.random.default_rng(0)
    for trial in range(40):
        n_rows = int(rng.integers(1, 300))
        deg = rng.integers(0, 30, size=n_rows)
        deg[rng.random(n_rows) < 0.3] = 0
        if trial % 3 == 0:
            deg[rng.integers(0, n_rows)] = int(rng.integers(300, 3000))   # a heavy user
        if trial % 5 == 0:
            deg[: n_rows // 2] = 0                                         # a long run of empty rows
        q_lo = int(rng.integers(0, 1000))
        rp = [int(x) + q_lo for x in np.concatenate([[0], np.cumsum(deg)])]
        q_hi = rp[-1]
        decision = rng.random(q_hi - q_lo) < 0.6
        row_of = np.repeat(np.arange(n_rows), deg)

        def keep_fn(row, q):
            assert row_of[q - q_lo] == row
            return decision[q - q_lo]

        kept, flag = m.walk(rp, n_rows, q_lo, q_hi, keep_fn)
        assert kept == np.bincount(row_of[decision], minlength=n_rows).tolist()
        assert flag == decision.tolist()
