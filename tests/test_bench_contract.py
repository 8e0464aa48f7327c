"""bench.py contract pieces that can run without a GPU: the reference (CPU) arm's JSON line, the synthetic generator's
determinism and the roofline arithmetic."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what torchrun sets for its children
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "2", "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, env=env, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "events/s" and d["higher_is_better"] is True
    assert d["metric"] == "CCO train events/sec to indicator model"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))           # all host threads, not OMP_NUM_THREADS=1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--gpus", "2"],
                       capture_output=True, text=True, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_synthetic_generator_is_deterministic_and_binary():
    import synth
    a, b = synth.make("tiny"), synth.make("tiny")
    for (nr, nc, rp, ci), (_, _, rp2, ci2) in zip(a.mats, b.mats):
        assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
        assert rp[0] == 0 and (np.diff(rp) >= 0).all() and len(ci) == rp[-1]
        for r in range(0, nr, 37):
            row = ci[rp[r]:rp[r + 1]]
            assert (np.diff(row) > 0).all()                        # sorted, no duplicates
    assert a.mats[0][3].tolist() != a.mats[1][3].tolist()          # event types differ (seed 1234 + t)


def test_min_events_per_user_shrinks_the_user_space_like_preparator():
    import synth
    w = synth.make("tiny", min_events_per_user=25)
    raw = synth.make("tiny")
    assert w.n_users < raw.n_users
    assert all(m[0] == w.n_users for m in w.mats)                  # one shared, compacted row space


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench

    class St:
        nnz_downsampled = [1000, 2000]
        products = [50_000, 70_000]
        distinct_cells = [40_000, 60_000]
        out_nnz = [3_000, 4_000]
    # 4 nnz(A') + 8 (I_A+1) + 8 nnz(A') + 4 P + 4 C + 4 I_A + 12 out
    assert bench.algorithmic_bytes(St, 1, 100) == 4 * 1000 + 8 * 101 + 8 * 1000 + 4 * 70_000 + 4 * 60_000 + 4 * 100 + 12 * 4_000
