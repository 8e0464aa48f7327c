"""PopModel rank histograms (PopModel.scala:113-182).  CPU: the restatement on hand-checkable events.
GPU: cco_pop_model against the restatement on random event logs, including the empty-bucket rules."""
import numpy as np
import pytest


def test_oracle_semantics_by_hand():
    from oracle import pop_oracle as po
    # interval [0, 90): thirds [0,30) [30,60) [60,90); halves [0,45) [45,90)
    items = [0, 0, 0, 1, 1, 1, 1, 2, 2, 3]
    times = [5, 35, 65, 10, 50, 70, 80, 40, 89, 90]     # item 3's only event sits ON the end: outside [start, end)
    assert po.pop_model("popular", items, times, 0, 90) == {0: 3.0, 1: 4.0, 2: 2.0}
    assert po.pop_model("trending", items, times, 0, 90) == {0: 0.0 - 0.0 + (1 - 2), 1: 3.0 - 1.0, 2: 1.0 - 1.0}
    assert po.pop_model("hot", items, times, 0, 90) == {0: (1 - 1) - (1 - 1), 1: (2 - 1) - (1 - 1)}     # item 2 misses the first third
    assert po.pop_model("trending", [0, 1], [50, 60], 0, 90) == {}      # older half empty -> empty result
    assert po.pop_model("hot", [0, 0], [5, 70], 0, 90) == {}            # middle third empty -> empty result


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["popular", "trending", "hot"])
def test_device_pop_model_matches_the_restatement(ctx, mode):
    from oracle import pop_oracle as po
    rng = np.random.default_rng(11)
    n_items, n_events = 5000, 400_000
    items = ((rng.zipf(1.2, n_events) - 1) % n_items).astype(np.int32)
    start, end = 1_500_000_000_000, 1_500_000_000_000 + 7 * 86_400_000 + 1        # a duration that 2 and 3 do not divide
    times = rng.integers(start - 86_400_000, end + 86_400_000, n_events)            # some events outside the interval
    assert ctx.pop_model(mode, items, times, n_items, start, end) == po.pop_model(mode, items.tolist(), times.tolist(), start, end)


@pytest.mark.gpu
def test_device_pop_model_empty_bucket_rules(ctx):
    assert ctx.pop_model("trending", [0, 1], [50, 60], 4, 0, 90) == {}
    assert ctx.pop_model("hot", [0, 0], [5, 70], 4, 0, 90) == {}
    assert ctx.pop_model("popular", [], [], 4, 0, 90) == {}
    assert ctx.pop_model("popular", [2, 2, 3], [0, 89, 90], 4, 0, 90) == {2: 2.0}
