"""Host-side glue for the multi-GPU path (one process per GPU): rendezvous of the NCCL id over
torch.distributed, the rank partition (cco_partition_rows) and the merge of per-rank indicator row slices.
torch.distributed is plumbing here (any host transport would do); the data-path collective -- the allreduce of the
column marginals -- runs inside libcco_b200.so on its own NCCL communicator."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from .similarity_analysis import CcoContext


def partition_rows(work_prefix: np.ndarray, world_size: int) -> np.ndarray:
    """Contiguous primary-item ranges of equal (products + 1 per row); bounds[r]..bounds[r+1] belongs to rank r."""
    wp = np.ascontiguousarray(work_prefix, dtype=np.int64)
    out = np.zeros(world_size + 1, dtype=np.int32)
    N.check(N.lib().cco_partition_rows(wp.ctypes.data_as(C.POINTER(C.c_int64)), len(wp) - 1, world_size,
                                       out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


def context_from_env(dist=None) -> CcoContext:
    """CcoContext for this torchrun rank (RANK / LOCAL_RANK / WORLD_SIZE); rank 0's NCCL id is broadcast with
    `dist` (an initialised torch.distributed)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    uid = None
    if world > 1:
        box = [CcoContext.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    return CcoContext(device=local, rank=rank, world_size=world, nccl_unique_id=uid)


def merge_row_slices(slices):
    """slices: list over ranks of (row_begin, row_end, n_cols, row_ptr, col_idx, llr, count) for ONE indicator ->
    the full (n_rows, n_cols, row_ptr, col_idx, llr, count).  Ranges must tile [0, n_rows) without gaps."""
    slices = sorted(slices, key=lambda s: (s[0], s[1]))
    n_cols = slices[0][2]
    expect = 0
    ptrs, cols, llrs, cnts = [np.zeros(1, dtype=np.int64)], [], [], []
    base = 0
    for rb, re_, nc, rp, ci, ll, cn in slices:
        if rb != expect or nc != n_cols or len(rp) != re_ - rb + 1:
            raise ValueError(f"row slices do not tile the item space: got [{rb},{re_}) after {expect}")
        ptrs.append(np.asarray(rp[1:], dtype=np.int64) + base)
        base += int(rp[-1])
        cols.append(ci)
        llrs.append(ll)
        cnts.append(cn)
        expect = re_
    cat = lambda xs, dt: np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dt)
    return expect, n_cols, np.concatenate(ptrs), cat(cols, np.int32), cat(llrs, np.float64), cat(cnts, np.int32)


def gather_indicators(dist, local_results):
    """all_gather every rank's result list (one tuple per indicator) and merge: every rank gets the full model."""
    world = dist.get_world_size()
    box = [None] * world
    dist.all_gather_object(box, local_results)
    n_ind = len(local_results)
    return [merge_row_slices([box[r][i] for r in range(world)]) for i in range(n_ind)]
