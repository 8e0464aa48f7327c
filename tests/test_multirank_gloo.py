"""N > 1 host path on CPU: two gloo ranks agree on the work-balanced row partition (cco_partition_rows), each
takes its slice of the indicators and the all-gather/merge reproduces the single-rank model exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import synth
        from oracle import oracle as orc
        from universal_recommender_b200 import distributed as D
        w = synth.make("tiny")
        full = orc.train([orc.Csr(*m) for m in w.mats], [orc.Params(*p) for p in w.params], 3)
        local = []
        for idx, ind in enumerate(full):
            # the same work vector on every rank -> the same partition without communication
            a_rp, a_ci = w.mats[0][2], w.mats[0][3]
            deg_b = np.diff(w.mats[idx][2])
            users = np.repeat(np.arange(w.n_users), np.diff(a_rp))
            work = np.bincount(a_ci, weights=deg_b[users], minlength=ind.n_rows).astype(np.int64)
            prefix = np.concatenate([[0], np.cumsum(work)])
            bounds = D.partition_rows(prefix, world)
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            s, e = int(ind.row_ptr[lo]), int(ind.row_ptr[hi])
            local.append((lo, hi, ind.n_cols, ind.row_ptr[lo:hi + 1] - ind.row_ptr[lo], ind.col_idx[s:e], ind.llr[s:e], ind.count[s:e]))
            box = [None] * world
            dist.all_gather_object(box, [int(b) for b in bounds])
            assert all(b == box[0] for b in box), "ranks disagree on the partition"
        merged = D.gather_indicators(dist, local)
        ok = True
        for m, ind in zip(merged, full):
            n_rows, n_cols, rp, ci, ll, cn = m
            ok &= n_rows == ind.n_rows and np.array_equal(rp, ind.row_ptr) and np.array_equal(ci, ind.col_idx)
            ok &= np.array_equal(ll, ind.llr) and np.array_equal(cn, ind.count)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_and_merge():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_partition_rows_properties():
    from universal_recommender_b200 import distributed as D
    rng = np.random.default_rng(0)
    work = (rng.zipf(1.5, 5000) % 10000).astype(np.int64)
    prefix = np.concatenate([[0], np.cumsum(work)])
    for world in (1, 2, 3, 8):
        b = D.partition_rows(prefix, world)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all()
        share = np.array([prefix[b[r + 1]] - prefix[b[r]] + (b[r + 1] - b[r]) for r in range(world)], dtype=np.float64)
        assert share.max() <= share.sum() / world + work.max() + 1          # balanced up to one row
    assert list(D.partition_rows(np.zeros(1, dtype=np.int64), 4)) == [0, 0, 0, 0, 0]


def test_merge_rejects_gaps():
    from universal_recommender_b200 import distributed as D
    z = lambda lo, hi: (lo, hi, 3, np.zeros(hi - lo + 1, dtype=np.int64), np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.int32))
    assert D.merge_row_slices([z(2, 5), z(0, 2)])[0] == 5
    with pytest.raises(ValueError):
        D.merge_row_slices([z(0, 2), z(3, 5)])
