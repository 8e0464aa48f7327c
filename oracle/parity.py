"""TEST INFRASTRUCTURE (like everything under oracle/): the parity gate of SURVEY.md 8(d) as one function.

compare(ref, got) checks a model produced by the CUDA path (through the C ABI) against the oracle's on the same
input and returns the `parity` block bench.py prints and tests assert on:

  counts_exact : every kept cell's co-occurrence count k11 equals the oracle's, row lengths equal   (bit-exact bar)
  topk_equal   : the kept column ids of every row equal the oracle's, in order (llr desc, col asc)  (bit-exact bar)
  max_llr_rel  : max |llr_gpu - llr_ref| / |llr_ref| over kept cells                               (bar: <= 1e-6)
  cancellation_limited : kept cells with |llr| < 1e-9 * N log N, where fp64 cancellation dominates (reported)

Only tests/, __graft_entry__.smoke() and bench.py's checker legs import this module.
"""
from __future__ import annotations

import numpy as np

LLR_RTOL = 1e-6


def compare(ref, got, n_users: int | None = None) -> dict:
    """ref: list of oracle Indicator; got: list of (row_begin, row_end, n_cols, row_ptr, col_idx, llr, count) covering
    ALL rows of each indicator (merge multi-GPU slices first)."""
    out = {"indicators": len(ref), "rows": 0, "cells": 0, "counts_exact": True, "topk_equal": True, "row_lengths_equal": True,
           "max_llr_rel": 0.0, "cancellation_limited": 0, "llr_rtol": LLR_RTOL}
    if len(ref) != len(got):
        out.update(counts_exact=False, topk_equal=False, row_lengths_equal=False, error="indicator count differs")
        return out
    for r, g in zip(ref, got):
        rb, re_, nc, rp, ci, ll, cn = g
        out["rows"] += int(r.n_rows)
        out["cells"] += int(len(r.col_idx))
        same_shape = (rb, re_, nc) == (0, r.n_rows, r.n_cols) and np.array_equal(rp, r.row_ptr)
        if not same_shape:
            out.update(counts_exact=False, topk_equal=False, row_lengths_equal=False)
            continue
        if not np.array_equal(ci, r.col_idx):
            out["topk_equal"] = False
        if not np.array_equal(cn, r.count):
            out["counts_exact"] = False
        if len(ll):
            ref_l = np.asarray(r.llr)
            rel = np.abs(np.asarray(ll) - ref_l) / np.maximum(np.abs(ref_l), 1e-300)
            out["max_llr_rel"] = max(out["max_llr_rel"], float(rel.max()))
            if n_users:
                noise = 1e-9 * n_users * np.log(max(n_users, 2))
                out["cancellation_limited"] += int((np.abs(ref_l) < noise).sum())
    out["ok"] = bool(out["counts_exact"] and out["topk_equal"] and out["row_lengths_equal"] and out["max_llr_rel"] <= LLR_RTOL)
    return out


def assert_ok(p: dict, tag: str = ""):
    assert p.get("row_lengths_equal"), f"{tag}: row lengths differ from the oracle: {p}"
    assert p.get("topk_equal"), f"{tag}: kept columns differ from the oracle: {p}"
    assert p.get("counts_exact"), f"{tag}: co-occurrence counts differ from the oracle: {p}"
    assert p.get("max_llr_rel", 1.0) <= LLR_RTOL, f"{tag}: LLR beyond {LLR_RTOL} relative: {p}"
