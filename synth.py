"""Synthetic Zipf-skewed user-event streams of the shapes BASELINE.json names.

item j drawn with p ~ 1/(rank+1)^s_i (s_i = 1.0), user u with p ~ 1/(rank+1)^s_u (s_u = 0.5); ranks are
mapped to ids by a fixed seeded permutation; events are split equally across event types; type t uses
seed 1234 + t; all types share the user space, each has its own item space.

The stream is counter based, so the numpy generator here and the CUDA generator of the library
(cco_synth_ingest, include/cco_b200.h) produce the SAME events bit for bit:
    h1 = mix64(mix64(seed) + (e + 1) * 0x9e3779b97f4a7c15),  h2 = mix64(h1 ^ 0x6a09e667f3bcc909)        e = 0 .. n-1
    user = user_perm[searchsorted(user_cdf, (h1 >> 11) * 2^-53, "right")],  item likewise from h2
make(name) runs on the host (numpy); make(name, ctx=<CcoContext>) generates and ingests on the B200 (seconds instead of
minutes at the 10M-user shapes) and copies the matrices back.  Either way the matrices come out exactly as Preparator
would build them: binary, deduplicated, CSR over users, minEventsPerUser applied on the primary event (duplicates count).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Workload:
    name: str
    n_users: int
    n_items: int
    n_events: int          # raw events over all types, before dedup
    n_types: int
    max_interactions: int = 500
    top_k: int = 50
    min_events_per_user: int | None = None
    mats: list | None = None      # [(n_rows, n_cols, row_ptr int64, col_idx int32)]
    events_per_type: list | None = None
    dataset: object = None        # resident dataset handle when generated on the device (ctx.free_dataset to release)

    @property
    def params(self):
        return [(self.max_interactions, self.top_k, None)] * self.n_types


CONFIGS = {
    # BASELINE.json configs[1]: MovieLens-1M-shaped
    "C2": dict(n_users=6_000, n_items=4_000, n_events=1_000_000, n_types=2),
    # configs[2]: 1M users x 100K items, 50M events, 1 primary + 3 secondary, k=50
    "C3": dict(n_users=1_000_000, n_items=100_000, n_events=50_000_000, n_types=4),
    # configs[3]: 10M x 1M, 500M events, downsampling + minEventsPerUser on
    "C4": dict(n_users=10_000_000, n_items=1_000_000, n_events=500_000_000, n_types=2, min_events_per_user=3),
    # configs[4]: 10M x 1M, 1B events, 1 + 8 types
    "C5": dict(n_users=10_000_000, n_items=1_000_000, n_events=1_000_000_000, n_types=9),
    # reduced shapes for tests / CPU-box development
    "tiny": dict(n_users=300, n_items=120, n_events=6_000, n_types=3),
    "small": dict(n_users=20_000, n_items=5_000, n_events=600_000, n_types=3),
    "C3-tenth": dict(n_users=100_000, n_items=10_000, n_events=5_000_000, n_types=4),
    # C4 at a tenth of the users and events (same 1M-column item space): the downsampling-dominated parity shape
    "C4-tenth": dict(n_users=1_000_000, n_items=1_000_000, n_events=50_000_000, n_types=2, min_events_per_user=3),
}

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_H2 = np.uint64(0x6A09E667F3BCC909)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z ^ (z >> np.uint64(30))
    z = z * np.uint64(0xBF58476D1CE4E5B9)
    z = z ^ (z >> np.uint64(27))
    z = z * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def zipf_cdf(n: int, s: float) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def user_tables(n_users: int, s_user: float = 0.5):
    """(cdf, perm) of the user space, shared by every event type"""
    return zipf_cdf(n_users, s_user), np.random.default_rng(99).permutation(n_users).astype(np.int32)


def item_tables(n_items: int, t: int, s_item: float = 1.0):
    return zipf_cdf(n_items, s_item), np.random.default_rng(1000 + t).permutation(n_items).astype(np.int32)


def type_seed(t: int) -> int:
    return 1234 + t


def events_for_type(n_users: int, n_items: int, n_events: int, t: int, tables=None):
    """-> (users int64[n_events], items int64[n_events]) raw events of event type t (the numpy twin of k_synth_events)."""
    ucdf, uperm = tables[0] if tables else user_tables(n_users)
    icdf, iperm = tables[1] if tables else item_tables(n_items, t)
    users = np.empty(n_events, dtype=np.int64)
    items = np.empty(n_events, dtype=np.int64)
    with np.errstate(over="ignore"):
        base = _mix64(np.array([type_seed(t)], dtype=np.uint64))[0]
        step = 1 << 22
        for s in range(0, n_events, step):
            e = np.arange(s + 1, min(n_events, s + step) + 1, dtype=np.uint64)
            h1 = _mix64(base + e * _GOLDEN)
            h2 = _mix64(h1 ^ _H2)
            u1 = (h1 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
            u2 = (h2 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
            users[s:s + len(e)] = uperm[np.minimum(np.searchsorted(ucdf, u1, side="right"), n_users - 1)]
            items[s:s + len(e)] = iperm[np.minimum(np.searchsorted(icdf, u2, side="right"), n_items - 1)]
    return users, items


def to_binary_csr(users: np.ndarray, items: np.ndarray, n_users: int, n_items: int):
    """dedup + sort: what IndexedDatasetSpark.apply leaves (Preparator.scala:195-208)."""
    keys = np.unique(users * np.int64(n_items) + items)
    r = keys // n_items
    c = (keys % n_items).astype(np.int32)
    rp = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n_users), out=rp[1:])
    return rp, c


def _cache_path(name: str, cfg: dict):
    """Generated matrices are cached (np.savez, /dev/shm by default) so that the processes of one GPU lease -- tests,
    the reference arm, the bench -- generate a workload once.  CCO_SYNTH_CACHE=0 disables, any other value = directory."""
    d = os.environ.get("CCO_SYNTH_CACHE", "/dev/shm/cco_synth")
    if d == "0" or cfg["n_events"] < 2_000_000:
        return None
    key = "_".join(f"{k}{cfg[k]}" for k in sorted(cfg))
    return os.path.join(d, f"v2_{name}_{key}.npz")


def _make_host(w: Workload, n_users_raw: int):
    """numpy path: Preparator semantics on the host.  The item dictionary is the RAW item space (every id of the Zipf
    support), exactly like the device path below, so both produce identical matrices."""
    per_type = w.n_events // w.n_types
    utab = user_tables(n_users_raw)
    keep_users, new_id = None, None
    for t in range(w.n_types):
        users, items = events_for_type(n_users_raw, w.n_items, per_type, t, (utab, item_tables(w.n_items, t)))
        if t == 0:
            # Preparator.scala:56-68, 129-132: users with < minEventsPerUser primary events (duplicates count; at least one
            # event) leave the user dictionary; every event type is then restricted to the passing users (:69-77)
            cnt = np.bincount(users, minlength=n_users_raw)
            keep_users = cnt >= max(w.min_events_per_user or 0, 1)
            new_id = np.cumsum(keep_users) - 1
            w.n_users = int(keep_users.sum())
        m = keep_users[users]
        users, items = new_id[users[m]], items[m]
        rp, ci = to_binary_csr(users, items, w.n_users, w.n_items)
        w.mats.append((w.n_users, w.n_items, rp, ci))
        w.events_per_type.append(per_type)


def _make_device(w: Workload, n_users_raw: int, ctx, keep_dataset: bool, pinned: bool):
    """CUDA path: events generated in HBM (k_synth_events) and ingested there (the cco_ingest pipeline); matrices copied
    back for the oracle / the end-to-end leg.  Item ids: the device ingest compacts the item dictionary to the items
    that have an event; the synthetic spaces are remapped back to the raw ids so that host and device agree."""
    per_type = w.n_events // w.n_types
    ucdf, uperm = user_tables(n_users_raw)
    types = []
    for t in range(w.n_types):
        icdf, iperm = item_tables(w.n_items, t)
        types.append((per_type, type_seed(t), icdf, iperm))
    ds = ctx.synth_dataset(types, n_users_raw, ucdf, uperm, w.min_events_per_user or 0, raw_item_space=True)
    for t in range(w.n_types):
        nr, nc, rp, ci = ctx.dataset_to_host(ds, t, pinned=pinned)
        w.n_users = nr
        w.mats.append((nr, nc, rp, ci))
        w.events_per_type.append(per_type)
    if keep_dataset:
        w.dataset = ds
    else:
        ctx.free_dataset(ds)


def make(name: str, ctx=None, keep_dataset: bool = False, pinned: bool = False, **override) -> Workload:
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    w = Workload(name=name, **cfg)
    n_users_raw = w.n_users
    w.mats, w.events_per_type = [], []
    if ctx is not None:
        _make_device(w, n_users_raw, ctx, keep_dataset, pinned)
        return w
    cp = _cache_path(name, cfg)
    if cp and os.path.exists(cp):
        try:
            z = np.load(cp)
            w.n_users = int(z["n_users"])
            w.mats = [(w.n_users, w.n_items, z[f"rp{t}"], z[f"ci{t}"]) for t in range(w.n_types)]
            w.events_per_type = [w.n_events // w.n_types] * w.n_types
            return w
        except Exception:
            w.mats, w.events_per_type = [], []
    _make_host(w, n_users_raw)
    if cp:
        try:
            os.makedirs(os.path.dirname(cp), exist_ok=True)
            tmp = cp + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, n_users=w.n_users, **{f"rp{t}": m[2] for t, m in enumerate(w.mats)},
                     **{f"ci{t}": m[3] for t, m in enumerate(w.mats)})
            os.replace(tmp, cp)
        except Exception:
            pass
    return w
