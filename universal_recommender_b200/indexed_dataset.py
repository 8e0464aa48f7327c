"""Host-side mirror of the Mahout types that cross the hot-path boundary.

`IndexedDataset` = binary user x item matrix + the two string<->int dictionaries, i.e. what
Preparator hands to URAlgorithm as `PreparedData.actions` (Preparator.scala:91-93) and what
URModel receives back (URModel.scala:32-36).  Values are implicit 1 (Preparator.scala:201-208).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np


class BiDictionary:
    """String <-> Int dictionary (org.apache.mahout.math.indexeddataset.BiDictionary).

    Index order is first appearance; Mahout's is the arbitrary order of `distinct().collect()`
    (Preparator.scala:170,184) -- nothing downstream depends on it."""

    def __init__(self, keys: Iterable[str] = ()):
        self._fwd: dict[str, int] = {}
        self._inv: list[str] = []
        for k in keys:
            self.add(k)

    def add(self, key: str) -> int:
        i = self._fwd.get(key)
        if i is None:
            i = len(self._inv)
            self._fwd[key] = i
            self._inv.append(key)
        return i

    def get(self, key: str, default: int = -1) -> int:
        return self._fwd.get(key, default)

    def contains(self, key: str) -> bool:
        return key in self._fwd

    __contains__ = contains

    @property
    def size(self) -> int:
        return len(self._inv)

    def __len__(self) -> int:
        return len(self._inv)

    @property
    def inverse(self) -> Sequence[str]:
        return self._inv


@dataclass
class IndexedDataset:
    """matrix (CSR, binary) + rowIDs + columnIDs.  `matrix.nrow` == len(rowIDs) is enforced the way
    Preparator does with `newRowCardinality(rowIDDictionary.size)` (Preparator.scala:213)."""
    row_ptr: np.ndarray            # int64 [n_rows + 1]
    col_idx: np.ndarray            # int32 [nnz]
    row_ids: BiDictionary
    column_ids: BiDictionary
    values: np.ndarray | None = None   # fp64 LLR for indicator matrices, None for binary inputs
    counts: np.ndarray | None = None   # k11 per kept cell (indicator matrices only)
    n_rows: int = field(default=-1)
    n_cols: int = field(default=-1)

    def __post_init__(self):
        self.row_ptr = np.ascontiguousarray(self.row_ptr, dtype=np.int64)
        self.col_idx = np.ascontiguousarray(self.col_idx, dtype=np.int32)
        if self.n_rows < 0:
            self.n_rows = len(self.row_ptr) - 1
        if self.n_cols < 0:
            self.n_cols = self.column_ids.size

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    def create(self, row_ptr, col_idx, row_ids: BiDictionary, column_ids: BiDictionary, values=None, counts=None):
        """IndexedDataset.create(drm, rowIDs, columnIDs): same backend, new payload."""
        return IndexedDataset(row_ptr, col_idx, row_ids, column_ids, values, counts,
                              n_rows=len(row_ptr) - 1, n_cols=column_ids.size)

    def row(self, i: int):
        s, e = int(self.row_ptr[i]), int(self.row_ptr[i + 1])
        v = None if self.values is None else self.values[s:e]
        return self.col_idx[s:e], v

    def to_string_map(self, action_name: str) -> dict[str, dict[str, list[str]]]:
        """IndexedDatasetConversions.toStringMapRDD (package.scala:82-110): per row, non-zeros sorted by
        -LLR mapped to column id strings; LLR values are discarded.  Rows arrive pre-sorted (llr desc,
        col asc) from the device, so the stable sort is a no-op and ties keep ascending column index."""
        out: dict[str, dict[str, list[str]]] = {}
        col_inv = self.column_ids.inverse
        row_inv = self.row_ids.inverse
        for r in range(self.n_rows):
            cols, vals = self.row(r)
            if vals is not None and len(cols) > 1:
                order = np.argsort(-vals, kind="stable")
                cols = cols[order]
            item = row_inv[r] if r < len(row_inv) else "INVALID_ITEM_ID"
            out[item] = {action_name: [col_inv[c] if c < len(col_inv) else "" for c in cols]}
        return out
