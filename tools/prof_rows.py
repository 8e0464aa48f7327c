"""Development: the shortest command that runs the resident hot path a few times (what ncu wraps).
usage: python tools/prof_rows.py [workload=C3] [trains=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
import universal_recommender_b200 as ur
from universal_recommender_b200 import _native as N
ctx = ur.CcoContext()
w = synth.make(sys.argv[1] if len(sys.argv) > 1 else "C3", ctx=ctx)
ds = ctx.upload(w.mats, ur.FLAG_ASSUME_CANONICAL)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    ctx.train_dataset(ds, w.params, 42, ur.FLAG_ASSUME_CANONICAL | N.FLAG_RESULT_ON_DEVICE, copy_arrays=False)
    st = ctx.last_stats
    print(f"train {it}: prep {st.ms_prepare:.2f} indicators {st.ms_cooccurrence:.2f} rows {[round(x, 3) for x in st.ms_indicator]} "
          f"evaluated {st.llr_evaluated} distinct {st.distinct_cells} products {st.products}", flush=True)
