"""ctypes binding of the CPU oracle (oracle/cco_oracle.c) + a dense numpy restatement.

TEST INFRASTRUCTURE ONLY (see oracle/cco_oracle.h): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs -- never by the
product package `universal_recommender_b200`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcco_oracle.so")

FLAG_ROWRATE_INTDIV = 1
FLAG_ENTROPY_VARARGS = 2


class _Csr(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int32),
                ("row_ptr", C.POINTER(C.c_int64)), ("col_idx", C.POINTER(C.c_int32))]


class _Params(C.Structure):
    _fields_ = [("max_interactions", C.c_int32), ("top_k", C.c_int32),
                ("has_min_llr", C.c_int32), ("min_llr", C.c_double)]


class _Result(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int32),
                ("row_ptr", C.POINTER(C.c_int64)), ("col_idx", C.POINTER(C.c_int32)),
                ("llr", C.POINTER(C.c_double)), ("count", C.POINTER(C.c_int32)),
                ("products", C.c_int64), ("distinct_cells", C.c_int64),
                ("nnz_a", C.c_int64), ("nnz_b", C.c_int64)]


class _Events(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("user", C.POINTER(C.c_int64)), ("item", C.POINTER(C.c_int32)), ("n_items_raw", C.c_int32)]


class _Ingested(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int32), ("row_ptr", C.POINTER(C.c_int64)),
                ("col_idx", C.POINTER(C.c_int32)), ("item_map", C.POINTER(C.c_int32))]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "cco_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_xlogx.restype = C.c_double
        L.orc_xlogx.argtypes = [C.c_int64]
        L.orc_llr.restype = C.c_double
        L.orc_llr.argtypes = [C.c_int64] * 4 + [C.c_int]
        L.orc_hash64.restype = C.c_uint64
        L.orc_hash64.argtypes = [C.c_int32, C.c_int64, C.c_int32]
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64]
        L.orc_last_error.restype = C.c_char_p
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_canonicalize.argtypes = [C.POINTER(_Csr), C.POINTER(C.POINTER(C.c_int64)),
                                       C.POINTER(C.POINTER(C.c_int32))]
        L.orc_downsample.argtypes = [C.POINTER(_Csr), C.c_int32, C.c_int32, C.c_int,
                                     C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int32)),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_cooccurrence.argtypes = [C.POINTER(_Csr), C.POINTER(_Csr), C.POINTER(C.POINTER(C.c_int64)),
                                       C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32))]
        L.orc_train.argtypes = [C.c_int, C.POINTER(_Csr), C.POINTER(_Params), C.c_int32, C.c_int, C.c_int,
                                C.POINTER(_Result)]
        L.orc_free_result.argtypes = [C.POINTER(_Result)]
        L.orc_ingest.argtypes = [C.c_int, C.POINTER(_Events), C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(_Ingested)]
        L.orc_free_ingested.argtypes = [C.POINTER(_Ingested)]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


@dataclass
class Csr:
    """Binary user x item matrix (values implicit 1)."""
    n_rows: int
    n_cols: int
    row_ptr: np.ndarray  # int64 [n_rows+1]
    col_idx: np.ndarray  # int32 [nnz]

    def __post_init__(self):
        self.row_ptr = np.ascontiguousarray(self.row_ptr, dtype=np.int64)
        self.col_idx = np.ascontiguousarray(self.col_idx, dtype=np.int32)

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    def _c(self) -> _Csr:
        return _Csr(self.n_rows, self.n_cols, self.row_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                    self.col_idx.ctypes.data_as(C.POINTER(C.c_int32)))

    def to_dense(self) -> np.ndarray:
        d = np.zeros((self.n_rows, self.n_cols), dtype=np.int64)
        for r in range(self.n_rows):
            d[r, self.col_idx[self.row_ptr[r]:self.row_ptr[r + 1]]] = 1
        return d

    @staticmethod
    def from_pairs(users, items, n_rows: int, n_cols: int) -> "Csr":
        """COO (user, item) events -> CSR keeping duplicates and arrival order inside a row."""
        users = np.asarray(users, dtype=np.int64)
        items = np.asarray(items, dtype=np.int32)
        order = np.argsort(users, kind="stable")
        counts = np.bincount(users, minlength=n_rows).astype(np.int64)
        rp = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(counts, out=rp[1:])
        return Csr(n_rows, n_cols, rp, items[order])

    @staticmethod
    def from_dense(d: np.ndarray) -> "Csr":
        d = np.asarray(d)
        rp = np.zeros(d.shape[0] + 1, dtype=np.int64)
        cols = []
        for r in range(d.shape[0]):
            nz = np.nonzero(d[r])[0]
            cols.append(nz)
            rp[r + 1] = rp[r] + len(nz)
        ci = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
        return Csr(d.shape[0], d.shape[1], rp, ci)


@dataclass
class Params:
    max_interactions: int = 500   # DefaultURAlgoParams.MaxEventsPerEventType (URAlgorithm.scala:54)
    top_k: int = 50               # DefaultURAlgoParams.MaxCorrelatorsPerEventType (URAlgorithm.scala:56)
    min_llr: float | None = None


@dataclass
class Indicator:
    """One indicator matrix: rows = primary items, columns = items of this event type."""
    n_rows: int
    n_cols: int
    row_ptr: np.ndarray
    col_idx: np.ndarray
    llr: np.ndarray
    count: np.ndarray
    products: int = 0
    distinct_cells: int = 0
    nnz_a: int = 0
    nnz_b: int = 0

    def row(self, i: int):
        s, e = int(self.row_ptr[i]), int(self.row_ptr[i + 1])
        return self.col_idx[s:e], self.llr[s:e], self.count[s:e]


def _take(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def xlogx(x: int) -> float:
    return lib().orc_xlogx(int(x))


def llr(k11: int, k12: int, k21: int, k22: int, flags: int = 0) -> float:
    return lib().orc_llr(int(k11), int(k12), int(k21), int(k22), flags)


def hash64(seed: int, u: int, j: int) -> int:
    return lib().orc_hash64(seed, u, j)


def u01(h: int) -> float:
    return lib().orc_u01(h)


def canonicalize(m: Csr) -> Csr:
    L = lib()
    rp = C.POINTER(C.c_int64)()
    ci = C.POINTER(C.c_int32)()
    cm = m._c()
    if L.orc_canonicalize(C.byref(cm), C.byref(rp), C.byref(ci)):
        raise OracleError(L.orc_last_error().decode())
    row_ptr = _take(rp, m.n_rows + 1, np.int64)
    col = _take(ci, int(row_ptr[-1]), np.int32)
    L.orc_free(rp)
    L.orc_free(ci)
    return Csr(m.n_rows, m.n_cols, row_ptr, col)


def downsample(m: Csr, max_interactions: int, seed: int, flags: int = 0):
    """-> (downsampled Csr, raw column counts, downsampled column counts). `m` must be canonical."""
    L = lib()
    rp = C.POINTER(C.c_int64)()
    ci = C.POINTER(C.c_int32)()
    raw = np.zeros(max(m.n_cols, 1), dtype=np.int32)
    new = np.zeros(max(m.n_cols, 1), dtype=np.int32)
    cm = m._c()
    if L.orc_downsample(C.byref(cm), max_interactions, seed, flags, C.byref(rp), C.byref(ci),
                        raw.ctypes.data_as(C.POINTER(C.c_int32)), new.ctypes.data_as(C.POINTER(C.c_int32))):
        raise OracleError(L.orc_last_error().decode())
    row_ptr = _take(rp, m.n_rows + 1, np.int64)
    col = _take(ci, int(row_ptr[-1]), np.int32)
    L.orc_free(rp)
    L.orc_free(ci)
    return Csr(m.n_rows, m.n_cols, row_ptr, col), raw[:m.n_cols], new[:m.n_cols]


def cooccurrence(a: Csr, b: Csr):
    """Full integer A^T B as (row_ptr, col_idx, count); inputs must be canonical."""
    L = lib()
    rp = C.POINTER(C.c_int64)()
    ci = C.POINTER(C.c_int32)()
    cn = C.POINTER(C.c_int32)()
    ca, cb = a._c(), b._c()
    if L.orc_cooccurrence(C.byref(ca), C.byref(cb), C.byref(rp), C.byref(ci), C.byref(cn)):
        raise OracleError(L.orc_last_error().decode())
    row_ptr = _take(rp, a.n_cols + 1, np.int64)
    n = int(row_ptr[-1])
    col, cnt = _take(ci, n, np.int32), _take(cn, n, np.int32)
    for p in (rp, ci, cn):
        L.orc_free(p)
    return row_ptr, col, cnt


def train(mats: list[Csr], params: list[Params], seed: int, flags: int = 0, n_threads: int = 0) -> list[Indicator]:
    """crossOccurrenceDownsampled: mats[0] is the primary; one Indicator per matrix."""
    L = lib()
    n = len(mats)
    cm = (_Csr * n)(*[m._c() for m in mats])
    cp = (_Params * n)(*[_Params(p.max_interactions, p.top_k, 0 if p.min_llr is None else 1,
                                 0.0 if p.min_llr is None else float(p.min_llr)) for p in params])
    res = (_Result * n)()
    if L.orc_train(n, cm, cp, seed, flags, n_threads, res):
        raise OracleError(L.orc_last_error().decode())
    out = []
    for r in res:
        rp = _take(r.row_ptr, r.n_rows + 1, np.int64)
        nnz = int(rp[-1])
        out.append(Indicator(int(r.n_rows), int(r.n_cols), rp, _take(r.col_idx, nnz, np.int32),
                             _take(r.llr, nnz, np.float64), _take(r.count, nnz, np.int32),
                             int(r.products), int(r.distinct_cells), int(r.nnz_a), int(r.nnz_b)))
        L.orc_free_result(C.byref(r))
    return out


def time_train(mats: list[Csr], params: list[Params], seed: int, flags: int = 0, n_threads: int = 0):
    """The CPU arm's timed call: orc_train only (no numpy copies of the result) -> (seconds, threads used, stats of the
    run: products / distinct cells / kept cells per indicator)."""
    import time
    L = lib()
    n = len(mats)
    cm = (_Csr * n)(*[m._c() for m in mats])
    cp = (_Params * n)(*[_Params(p.max_interactions, p.top_k, 0 if p.min_llr is None else 1,
                                 0.0 if p.min_llr is None else float(p.min_llr)) for p in params])
    res = (_Result * n)()
    t0 = time.perf_counter()
    rc = L.orc_train(n, cm, cp, seed, flags, n_threads, res)
    dt = time.perf_counter() - t0
    if rc:
        raise OracleError(L.orc_last_error().decode())
    stats = [(int(r.products), int(r.distinct_cells), int(r.row_ptr[r.n_rows])) for r in res]
    for r in res:
        L.orc_free_result(C.byref(r))
    return dt, (n_threads if n_threads > 0 else L.orc_max_threads()), stats


def ingest(events, n_users_raw: int, min_events_per_user: int = 0):
    """Preparator.prepare on integer-tokenised events.  events = [(users int64[], items int32[], n_items_raw)],
    type 0 = primary.  -> (user_map int32[n_users_raw], [(Csr, item_map int32[n_items_raw])])."""
    L = lib()
    n = len(events)
    keep = []
    ev = (_Events * n)()
    for t, (u, i, ni) in enumerate(events):
        u = np.ascontiguousarray(u, dtype=np.int64)
        i = np.ascontiguousarray(i, dtype=np.int32)
        keep.append((u, i))
        ev[t] = _Events(len(u), u.ctypes.data_as(C.POINTER(C.c_int64)), i.ctypes.data_as(C.POINTER(C.c_int32)), ni)
    user_map = np.zeros(max(n_users_raw, 1), dtype=np.int32)
    out = (_Ingested * n)()
    if L.orc_ingest(n, ev, n_users_raw, min_events_per_user, user_map.ctypes.data_as(C.POINTER(C.c_int32)), out):
        raise OracleError(L.orc_last_error().decode())
    res = []
    for t, o in enumerate(out):
        rp = _take(o.row_ptr, o.n_rows + 1, np.int64)
        ci = _take(o.col_idx, int(rp[-1]), np.int32)
        im = _take(o.item_map, events[t][2], np.int32)
        res.append((Csr(int(o.n_rows), int(o.n_cols), rp, ci), im))
        L.orc_free_ingested(C.byref(o))
    return user_map[:n_users_raw], res


def cooccurrences_idss(mats: list[Csr], seed: int, max_interesting: int = 50, max_interactions: int = 500,
                       flags: int = 0, n_threads: int = 0) -> list[Indicator]:
    """SimilarityAnalysis.cooccurrencesIDSs: one global (k, m) for every matrix (URAlgorithm.scala:323-329)."""
    return train(mats, [Params(max_interactions, max_interesting, None) for _ in mats], seed, flags, n_threads)


# ------------------------------------------------------------------------------------------------
# Independent dense restatement (pure Python/numpy; tiny inputs only) used to cross-check the C
# oracle itself: counts by matrix product, LLR by math.log in Mahout's evaluation order.
# ------------------------------------------------------------------------------------------------
def _xlogx_py(x: int) -> float:
    return 0.0 if x == 0 else x * math.log(x)


def llr_py(k11: int, k12: int, k21: int, k22: int) -> float:
    row = _xlogx_py(k11 + k12 + k21 + k22) - _xlogx_py(k11 + k12) - _xlogx_py(k21 + k22)
    col = _xlogx_py(k11 + k21 + k12 + k22) - _xlogx_py(k11 + k21) - _xlogx_py(k12 + k22)
    mat = _xlogx_py(k11 + k12 + k21 + k22) - _xlogx_py(k11) - _xlogx_py(k12) - _xlogx_py(k21) - _xlogx_py(k22)
    if row + col < mat:
        return 0.0
    return 2.0 * (row + col - mat)


def dense_indicator(a: np.ndarray, b: np.ndarray, k: int, self_cooc: bool, min_llr: float | None = None):
    """Dense restatement of computeSimilarities on already-downsampled 0/1 matrices.
    -> list over primary items of [(col, llr, k11)] sorted (llr desc, col asc), LLR > 0 only."""
    a = (np.asarray(a) != 0).astype(np.int64)
    b = (np.asarray(b) != 0).astype(np.int64)
    n = a.shape[0]
    c = a.T @ b
    ma, mb = a.sum(0), b.sum(0)
    rows = []
    for i in range(a.shape[1]):
        cand = []
        for j in range(b.shape[1]):
            k11 = int(c[i, j])
            if k11 == 0 or (self_cooc and i == j):
                continue
            v = llr_py(k11, int(ma[i]) - k11, int(mb[j]) - k11, n - int(ma[i]) - int(mb[j]) + k11)
            if min_llr is not None and not v >= min_llr:
                continue
            if v > 0.0:
                cand.append((j, v, k11))
        cand.sort(key=lambda t: (-t[1], t[0]))
        rows.append(cand[:k])
    return rows
