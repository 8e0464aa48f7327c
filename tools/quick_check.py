"""Development smoke: GPU path vs oracle on a synthetic workload (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import synth
from oracle import oracle as orc
import universal_recommender_b200 as ur

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
do_oracle = (len(sys.argv) <= 2) or sys.argv[2] != "nooracle"
t0 = time.time(); ctx = ur.CcoContext()
w = synth.make(name, ctx=ctx); print(f"{name}: generated in {time.time()-t0:.1f}s; nnz per type {[int(m[2][-1]) for m in w.mats]}")
for it in range(int(os.environ.get("QC_ITERS", "3"))):
    t0 = time.time(); res = ctx.train_csr(w.mats, w.params, seed=42, flags=ur.FLAG_ASSUME_CANONICAL); dt = time.time() - t0
    st = ctx.last_stats
    print(f"gpu iter {it}: wall {dt*1e3:.1f} ms  total {st.ms_total:.2f} h2d {st.ms_h2d:.2f} prep {st.ms_prepare:.2f} cooc {st.ms_cooccurrence:.2f} rows {['%.3f'%x for x in st.ms_indicator]} launches {st.n_kernel_launches}")
print("llr_evaluated", st.llr_evaluated); print("products", st.products, "distinct", st.distinct_cells, "out_nnz", st.out_nnz, "nnz_ds", st.nnz_downsampled)
if do_oracle:
    mats = [orc.Csr(*m) for m in w.mats]
    t0 = time.time(); ref = orc.train(mats, [orc.Params(*p) for p in w.params], 42); print(f"oracle: {time.time()-t0:.2f}s ({orc.lib().orc_max_threads()} threads)")
    for i, (r, g) in enumerate(zip(ref, res)):
        rb, re_, nc, rp, ci, ll, cn = g
        ok_ptr = np.array_equal(rp, r.row_ptr)
        same_cols = ok_ptr and np.array_equal(ci, r.col_idx)
        same_cnt = ok_ptr and np.array_equal(cn, r.count)
        rel = np.max(np.abs(ll - r.llr) / np.maximum(np.abs(r.llr), 1e-300)) if ok_ptr and len(ll) else -1
        print(f"indicator {i}: row_ptr equal {ok_ptr} cols equal {same_cols} counts equal {same_cnt} max rel llr err {rel:.3e} | products {r.products} vs {st.products[i]} distinct {r.distinct_cells} vs {st.distinct_cells[i]} nnz_b {r.nnz_b} vs {st.nnz_downsampled[i]}")
        if ok_ptr and not same_cols:
            bad = np.nonzero(ci != r.col_idx)[0]
            print("   first col mismatches:", bad[:5], ci[bad[:5]], r.col_idx[bad[:5]], ll[bad[:5]], r.llr[bad[:5]])
