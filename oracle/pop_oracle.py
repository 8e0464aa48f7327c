"""TEST INFRASTRUCTURE: CPU restatement of PopModel's rank histograms, the checker of cco_pop_model.

Follows /root/reference/src/main/scala/PopModel.scala: calcPopular :113-122 (events per item in the interval),
calcTrending :128-148 (newer half minus older half over the items present in both -- an inner join -- and nothing at all if
the older half has no event), calcHot :153-182 (thirds; (newer - middle) - (middle - older) over the items present in all three;
nothing if the older or the middle third is empty).  Intervals are [start, end) (PEventStore.find: startTime inclusive,
untilTime exclusive); the bucket edges use Joda's integer millisecond arithmetic."""
from __future__ import annotations

from collections import Counter


def _popular(items, times, lo, hi):
    return Counter(int(j) for j, t in zip(items, times) if lo <= t < hi)


def pop_model(mode: str, items, times_ms, start_ms: int, end_ms: int) -> dict:
    dur = end_ms - start_ms
    if mode == "popular":
        return {j: float(c) for j, c in _popular(items, times_ms, start_ms, end_ms).items()}
    if mode == "trending":
        half = dur // 2
        older = _popular(items, times_ms, start_ms, start_ms + half)
        if not older:
            return {}
        newer = _popular(items, times_ms, start_ms + half, end_ms)
        return {j: float(newer[j] - older[j]) for j in newer if j in older}
    if mode == "hot":
        third = dur // 3
        older = _popular(items, times_ms, start_ms, start_ms + third)
        if not older:
            return {}
        middle = _popular(items, times_ms, start_ms + third, start_ms + 2 * third)
        if not middle:
            return {}
        newer = _popular(items, times_ms, start_ms + 2 * third, end_ms)
        new_v = {j: newer[j] - middle[j] for j in newer if j in middle}
        old_v = {j: middle[j] - older[j] for j in middle if j in older}
        return {j: float(new_v[j] - old_v[j]) for j in new_v if j in old_v}
    raise ValueError(mode)
