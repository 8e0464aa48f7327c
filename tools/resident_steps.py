"""Development: per-step stage times of the device-resident path (cco_train_dataset) on a synthetic workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
import universal_recommender_b200 as ur
from universal_recommender_b200 import _native as N
ctx = ur.CcoContext()
w = synth.make(sys.argv[1] if len(sys.argv) > 1 else "C3", ctx=ctx)
ds = ctx.upload(w.mats, ur.FLAG_ASSUME_CANONICAL)
for flags, name in ((ur.FLAG_ASSUME_CANONICAL | N.FLAG_RESULT_ON_DEVICE, "resident/no-D2H"), (ur.FLAG_ASSUME_CANONICAL, "resident/with-D2H")):
    for it in range(6):
        ctx.timer_start(); t0 = time.perf_counter()
        ctx.train_dataset(ds, w.params, 42, flags, copy_arrays=False)
        ms = ctx.timer_stop(); wall = (time.perf_counter() - t0) * 1e3
        st = ctx.last_stats
        print(f"{name} step {it}: events {ms:.2f} ms wall {wall:.2f} ms | prep {st.ms_prepare:.2f} indicators {st.ms_cooccurrence:.2f} rows {sum(st.ms_indicator):.2f}")
