"""Synthetic Zipf-skewed user-event streams of the shapes BASELINE.json names.

item j drawn with p ~ 1/(rank+1)^s_i (s_i = 1.0), user u with p ~ 1/(rank+1)^s_u (s_u = 0.5); ranks are
mapped to ids by a fixed seeded permutation; events are split equally across event types; type t uses
seed 1234 + t; all types share the user space, each has its own item space.  Deterministic (numpy PCG64).
The matrices come out exactly as Preparator would build them: binary, deduplicated, CSR over users.
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
}


def _zipf_cdf(n: int, s: float) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def _draw(rng: np.random.Generator, cdf: np.ndarray, perm: np.ndarray, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    step = 1 << 24
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = perm[np.searchsorted(cdf, rng.random(e - s), side="right").clip(0, len(cdf) - 1)]
    return out


def events_for_type(n_users: int, n_items: int, n_events: int, t: int, s_user: float = 0.5, s_item: float = 1.0):
    """-> (users int64[n_events], items int64[n_events]) raw events of event type t."""
    rng = np.random.default_rng(1234 + t)
    user_perm = np.random.default_rng(99).permutation(n_users)     # shared by every type
    item_perm = np.random.default_rng(1000 + t).permutation(n_items)
    users = _draw(rng, _zipf_cdf(n_users, s_user), user_perm, n_events)
    items = _draw(rng, _zipf_cdf(n_items, s_item), item_perm, n_events)
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
    return os.path.join(d, f"{name}_{key}.npz")


def make(name: str, **override) -> Workload:
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    w = Workload(name=name, **cfg)
    cp = _cache_path(name, cfg)
    if cp and os.path.exists(cp):
        try:
            z = np.load(cp)
            w.n_users = int(z["n_users"])
            w.mats = [(w.n_users, w.n_items, z[f"rp{t}"], z[f"ci{t}"]) for t in range(w.n_types)]
            w.events_per_type = [w.n_events // w.n_types] * w.n_types
            return w
        except Exception:
            pass
    per_type = w.n_events // w.n_types
    w.mats, w.events_per_type = [], []
    keep_users = None
    n_users_raw = w.n_users
    for t in range(w.n_types):
        users, items = events_for_type(n_users_raw, w.n_items, per_type, t)
        if t == 0 and w.min_events_per_user:
            # Preparator.scala:56-68: users with < minEventsPerUser primary events (duplicates count) leave the
            # user dictionary; every event type is then restricted to the passing users (:69-77) and N shrinks.
            cnt = np.bincount(users, minlength=n_users_raw)
            keep_users = cnt >= w.min_events_per_user
            new_id = np.cumsum(keep_users) - 1
            w.n_users = int(keep_users.sum())
        if keep_users is not None:
            m = keep_users[users]
            users, items = new_id[users[m]], items[m]
        rp, ci = to_binary_csr(users, items, w.n_users, w.n_items)
        w.mats.append((w.n_users, w.n_items, rp, ci))
        w.events_per_type.append(per_type)
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
