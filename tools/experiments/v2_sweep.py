"""EXPERIMENT RECORD (profiles/r02_k_rows2_experiment.md): the CCO_ROWS_IMPL / CCO_V2_TUNE switches this script sets existed
only while tools/experiments/cco_rows2.cuh was compiled into the library; kept to show how the A/B numbers were taken.
Development: sweep k_rows2's development switches (CCO_V2_TUNE) -- parity on a few workloads + C3 row-kernel time."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import os, sys, time
sys.path.insert(0, os.path.dirname(%r))
import numpy as np
import synth
import universal_recommender_b200 as ur
from universal_recommender_b200 import _native as N
from oracle import oracle as orc
from oracle import parity as par
ctx = ur.CcoContext()
rng = np.random.default_rng(5)
ev = [((rng.zipf(1.4, 60000) - 1) %% 5000, ((rng.zipf(1.2, 60000) - 1) %% 800).astype(np.int32), 800) for _ in range(3)]
ds, um, im = ctx.ingest(ev, 5000, 0)
mats = [ctx.dataset_matrix(ds, t) for t in range(3)]
cases = [("ingest-like", mats, [(500, 20, None)] * 3)]
for n in os.environ.get("SWEEP_CASES", "tiny,small,C3-tenth").split(","):
    wn = synth.make(n, ctx=ctx)
    cases.append((n, wn.mats, wn.params))
for name, m, prm in cases:
    got = ctx.train_csr(m, prm, seed=9)
    ref = orc.train([orc.Csr(*x) for x in m], [orc.Params(*p) for p in prm], 9)
    p = par.compare(ref, got)
    bad_rows = 0
    if not p["ok"]:
        for r, g in zip(ref, got):
            if np.array_equal(g[3], r.row_ptr):
                rows = np.repeat(np.arange(r.n_rows), np.diff(r.row_ptr))
                bad_rows += len(np.unique(rows[g[4] != r.col_idx]))
    print(f"  {name:12s} ok={p['ok']} topk={p['topk_equal']} counts={p['counts_exact']} llr={p['max_llr_rel']:.1e} bad_rows={bad_rows}", flush=True)
if os.environ.get("SWEEP_C3", "1") != "1": sys.exit(0)
w = synth.make("C3", ctx=ctx)
d3 = ctx.upload(w.mats, ur.FLAG_ASSUME_CANONICAL)
best = 1e9
for it in range(4):
    ctx.train_dataset(d3, w.params, 42, ur.FLAG_ASSUME_CANONICAL | N.FLAG_RESULT_ON_DEVICE, copy_arrays=False)
    best = min(best, sum(ctx.last_stats.ms_indicator))
print(f"  C3 rows {best:.3f} ms / step", flush=True)
''' % here
for impl, tune in [tuple(x.split(":")) for x in os.environ.get("SWEEP_VARIANTS", "1:0,2:0,2:1,2:2,2:8,2:16,2:27").split(",")]:
    print(f"==== CCO_ROWS_IMPL={impl} CCO_V2_TUNE={tune}", flush=True)
    try:
        subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, CCO_ROWS_IMPL=impl, CCO_V2_TUNE=tune), timeout=int(os.environ.get("SWEEP_TIMEOUT", "120")))
    except subprocess.TimeoutExpired:
        print("  TIMEOUT (hang)", flush=True)
