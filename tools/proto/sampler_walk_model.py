"""Lane-level model of k_sample_count (universal_recommender_b200/csrc/cco_sampler.cuh): a warp walks a
chunk of consecutive stored entries with a window of 32 rows held by its lanes.  The model proves the bookkeeping --
row of every entry, one decision per entry, kept counts per row through lane-range masks and window flushes -- against
a plain per-row count (tests/test_round2_models.py).  The keep decision itself is a callback."""
NO_ROW = 2 ** 63 - 1


def find_row(rp, n_rows, q):
    """largest r in [0, n_rows) with rp[r] <= q: the 32-ary probe of warp_find_row"""
    lo, hi = 0, n_rows
    while hi - lo > 1:
        step = (hi - lo + 31) // 32
        ballot = 1   # lane 0 probes rp[lo] <= q, the loop invariant
        for lane in range(32):
            p = lo + lane * step
            if p < hi and rp[p] <= q:
                ballot |= 1 << lane
        lo += (ballot.bit_length() - 1) * step
        hi = min(hi, lo + step)
    return lo


def walk(rp, n_rows, q_lo, q_hi, keep_fn, chunk=256):
    """-> (kept_per_row list, keep flag per entry); keep_fn(row, q) is called exactly once per stored entry"""
    kept = [0] * n_rows
    flag = [None] * (q_hi - q_lo)
    for c in range((q_hi - q_lo + chunk - 1) // chunk):
        q0 = q_lo + c * chunk
        q1 = min(q0 + chunk, q_hi)
        base = find_row(rp, n_rows, q0)

        def load(base):
            start, end = [NO_ROW] * 32, [NO_ROW] * 32
            for lane in range(32):
                if base + lane < n_rows:
                    start[lane], end[lane] = rp[base + lane], rp[base + lane + 1]
            return start, end, [0] * 32

        start, end, mine = load(base)
        qb, past_last_row = q0, False
        while qb < q1 and not past_last_row:
            pending = [qb + lane < q1 for lane in range(32)]
            while True:
                ballot, here = 0, [False] * 32
                for lane in range(32):
                    q, idx = qb + lane, 0
                    for step in (16, 8, 4, 2, 1):   # rows of the window ending at or before q
                        if end[idx + step - 1] <= q:
                            idx += step
                    if end[idx] <= q:
                        idx += 1
                    here[lane] = pending[lane] and idx < 32
                    if here[lane]:
                        keep = base + idx < n_rows and bool(keep_fn(base + idx, q))
                        assert flag[q - q_lo] is None
                        flag[q - q_lo] = keep
                        if keep:
                            ballot |= 1 << lane
                for lane in range(32):   # the owner lane counts its row's kept entries of this batch
                    a = 0 if start[lane] <= qb else (32 if start[lane] >= qb + 32 else start[lane] - qb)
                    b = 0 if end[lane] <= qb else (32 if end[lane] >= qb + 32 else end[lane] - qb)
                    if b > a:
                        mine[lane] += bin(ballot & ((1 << b) - 1) & ~((1 << a) - 1)).count("1")
                pending = [p and not h for p, h in zip(pending, here)]
                if not any(pending):
                    break
                for lane in range(32):   # slide the window
                    if mine[lane]:
                        kept[base + lane] += mine[lane]
                base += 32
                if base >= n_rows:
                    past_last_row = True
                    break
                start, end, mine = load(base)
            qb += 32
        if not past_last_row:
            for lane in range(32):
                if mine[lane]:
                    kept[base + lane] += mine[lane]
    return kept, flag
