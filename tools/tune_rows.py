"""Development: sweep the host-side launch knobs of k_rows (CCO_TUNE_* environment variables) on one workload and
check that every variant returns bit-identical indicators."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
import universal_recommender_b200 as ur
ctx = ur.CcoContext()
w = synth.make(sys.argv[1] if len(sys.argv) > 1 else "C3", ctx=ctx)
ds = ctx.upload(w.mats, ur.FLAG_ASSUME_CANONICAL)
variants = [{}, {"CCO_TUNE_GRID_MULT": "2"}, {"CCO_TUNE_GRID_MULT": "4"}, {"CCO_TUNE_CAP": "66"}, {"CCO_TUNE_CAP": "75"},
            {"CCO_TUNE_WARP_CTA": "128"}, {"CCO_TUNE_WARP_CTA": "256"}, {"CCO_TUNE_WARP_CBUF": "128"}, {"CCO_TUNE_SERIAL": "1"},
            {"CCO_TUNE_SPEC": "1"}, {"CCO_TUNE_SPEC": "2"}, {"CCO_TUNE_GRID_MULT": "4", "CCO_TUNE_CAP": "66"}]
base = None
for v in variants:
    for k in list(os.environ):
        if k.startswith("CCO_TUNE_"):
            del os.environ[k]
    os.environ.update(v)
    best = 1e9
    for it in range(4):
        res = ctx.train_dataset(ds, w.params, 42, ur.FLAG_ASSUME_CANONICAL, copy_arrays=(it == 0))
        if it == 0:
            sig = [(r[3].tobytes(), r[4].tobytes(), r[5].tobytes(), r[6].tobytes()) for r in res]
            if base is None:
                base = sig
            same = sig == base
        st = ctx.last_stats
        best = min(best, sum(st.ms_indicator))
    print(f"{str(v):60s} rows {best:7.3f} ms  prep {st.ms_prepare:.2f}  identical_output={same}", flush=True)
