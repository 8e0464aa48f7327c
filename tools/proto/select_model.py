"""Python model of the integer-domain top-k selection for one row of A'^T B' (round-2 prototype).
Validated against brute force with the oracle's LLR."""
import sys, random, heapq
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle as orc

def llr(k11, ra, cb, N):
    return orc.llr(k11, ra - k11, cb - k11, N - ra - cb + k11)

def brute(cells, ra, N, k, min_llr, item, self_):
    out = []
    for b, k11, cb in cells:
        if self_ and b == item: continue
        v = llr(k11, ra, cb, N)
        if min_llr is not None and not v >= min_llr: continue
        if v > 0: out.append((-v, b, k11))
    out.sort()
    return [(b, -nv, k11) for nv, b, k11 in out[:k]]

def integer_domain(cells, ra, N, k, min_llr, item, self_, J=3, Cmax=1024, stats=None):
    H = {}          # (j, c) -> count     (the GPU keeps J x Cmax counters)
    R = []          # cells evaluated directly
    S = []
    for b, k11, cb in cells:
        if self_ and b == item: continue
        if k11 <= J and cb < Cmax and 2 * ra * cb < k11 * N:
            H[(k11, cb)] = H.get((k11, cb), 0) + 1
            S.append((b, k11, cb))
        else:
            R.append((b, k11, cb))
    evals = 0
    E = []
    for b, k11, cb in R:
        v = llr(k11, ra, cb, N); evals += 1
        if min_llr is not None and not v >= min_llr: continue
        if v > 0: E.append((v, b, k11))
    E.sort(key=lambda t: (-t[0], t[1]))
    # level heads: ascending c among non-empty bins
    bins = {j: sorted(c for (jj, c) in H if jj == j) for j in range(1, J + 1)}
    ptr = {j: 0 for j in bins}
    head = {}
    def load(j):
        nonlocal evals
        while ptr[j] < len(bins[j]):
            c = bins[j][ptr[j]]
            v = llr(j, ra, c, N); evals += 1
            if (min_llr is not None and not v >= min_llr) or not v > 0:
                ptr[j] = len(bins[j])      # monotone: every later bin of this level fails too
                break
            head[j] = (v, c)
            return
        head.pop(j, None)
    for j in bins: load(j)
    ei = 0
    chosen_bins = []     # (j, c, take_all or None) in order
    result = []
    remaining = k
    cells_by_bin = {}
    for b, k11, cb in S: cells_by_bin.setdefault((k11, cb), []).append(b)
    while remaining > 0:
        cand = [head[j][0] for j in head]
        if ei < len(E): cand.append(E[ei][0])
        if not cand: break
        v = max(cand)
        group = []     # (col, k11) of every cell whose LLR == v
        for j in list(head):
            if head[j][0] == v:
                c = head[j][1]
                group += [(b, j) for b in cells_by_bin[(j, c)]]
                ptr[j] += 1; load(j)
        while ei < len(E) and E[ei][0] == v:
            group.append((E[ei][1], E[ei][2])); ei += 1
        group.sort()
        for b, k11 in group[:remaining]:
            result.append((b, v, k11))
        remaining -= min(remaining, len(group))
    if stats is not None:
        stats['evals'] = stats.get('evals', 0) + evals
        stats['cells'] = stats.get('cells', 0) + len(cells)
    return result

def random_row(rng):
    N = 10 ** rng.randrange(2, 8)
    ra = rng.randrange(1, max(2, min(600, N // 2)))
    n_cells = rng.randrange(0, 800)
    cells = []
    used = set()
    for _ in range(n_cells):
        b = rng.randrange(0, 5000)
        if b in used: continue
        used.add(b)
        style = rng.random()
        cb = rng.randrange(1, min(600, N - ra) + 1) if style < 0.8 else rng.randrange(1, max(2, min(N - ra, 5000)) + 1)
        k11 = 1 if rng.random() < 0.85 else rng.randrange(1, 8)
        k11 = min(k11, ra, cb)
        if N - ra - cb + k11 < 0: continue
        cells.append((b, k11, cb))
    return cells, ra, N

if __name__ == '__main__':
    rng = random.Random(7)
    stats = {}
    for t in range(3000):
        cells, ra, N = random_row(rng)
        k = rng.choice([1, 5, 50, 50, 50, 200])
        min_llr = rng.choice([None, None, None, 0.5, 5.0])
        item = rng.randrange(0, 5000); self_ = rng.random() < 0.3
        a = brute(cells, ra, N, k, min_llr, item, self_)
        b = integer_domain(cells, ra, N, k, min_llr, item, self_, stats=stats)
        assert a == b, (t, len(cells), ra, N, k, min_llr, a[:5], b[:5])
    print('3000 random rows identical; evaluated', stats['evals'], 'of', stats['cells'], 'cells =', round(stats['evals'] / stats['cells'], 3))
