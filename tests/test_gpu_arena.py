"""Result arena (cco_config_t.result_arena) and result modes: the indicator arrays land in caller-provided memory."""
import numpy as np
import pytest

import synth
import universal_recommender_b200 as ur

pytestmark = pytest.mark.gpu


def test_results_are_placed_in_the_callers_arena_and_the_arena_is_reused(orc):
    from oracle import parity as par
    arena = np.zeros(64 << 20, dtype=np.uint8)
    c = ur.CcoContext(device=0, result_arena=arena)
    try:
        w = synth.make("small")
        ref = orc.train([orc.Csr(*m) for m in w.mats], [orc.Params(*p) for p in w.params], 3)
        lo, hi = arena.ctypes.data, arena.ctypes.data + arena.nbytes
        first = None
        for _ in range(3):
            res, h = c.train_csr(w.mats, w.params, 3, keep=True)
            for r in res:
                for a in r[3:]:
                    assert lo <= a.ctypes.data < hi, "result array outside the arena"
            assert par.compare(ref, res, w.n_users)["ok"]
            addr = res[0][3].ctypes.data
            c.free_result(h)                       # every result freed -> the arena starts over
            first = first or addr
            assert addr == first
        # a result that does not fit falls back to library-owned pinned memory (still correct)
        tiny_arena = np.zeros(4096, dtype=np.uint8)
        c2 = ur.CcoContext(device=0, result_arena=tiny_arena)
        got = c2.train_csr(w.mats, w.params, 3)
        assert par.compare(ref, got, w.n_users)["ok"]
        c2.close()
    finally:
        c.close()


def test_result_modes_drop_the_arrays_the_reference_consumer_never_reads(ctx):
    w = synth.make("small")
    full = ctx.train_csr(w.mats, w.params, 5)
    no_cnt = ctx.train_csr(w.mats, w.params, 5, flags=ur.FLAG_RESULT_NO_COUNT)
    ids_only = ctx.train_csr(w.mats, w.params, 5, flags=ur.FLAG_RESULT_NO_COUNT | ur.FLAG_RESULT_NO_LLR)
    for f, n, i in zip(full, no_cnt, ids_only):
        assert np.array_equal(f[3], n[3]) and np.array_equal(f[4], n[4]) and np.array_equal(f[5], n[5]) and len(n[6]) == 0
        assert np.array_equal(f[3], i[3]) and np.array_equal(f[4], i[4]) and len(i[5]) == 0 and len(i[6]) == 0
