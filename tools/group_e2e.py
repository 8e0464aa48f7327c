"""End-to-end throughput of the GROUP context (cco_create_group: ONE process, one context over every GPU of the box, the
mode the JNI shim uses): cco_train from pinned host CSR to ONE merged model in host memory, wall clock.
usage: python tools/group_e2e.py [workload=C3] [steps=10] [n_gpus=all] [check=1]   -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth
import universal_recommender_b200 as ur

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "all" else torch.cuda.device_count()
check = (sys.argv[4] if len(sys.argv) > 4 else "1") == "1"
gen = ur.CcoContext(device=0)
w = synth.make(wl, ctx=gen)                      # generated + ingested on GPU 0, copied back
gen.close()
g = ur.CcoContext(devices=list(range(n)))
pinned = []
for (nr, nc, rp, ci) in w.mats:
    prp, pci = g.host_array(len(rp), np.int64), g.host_array(len(ci), np.int32)
    prp[:] = rp
    pci[:] = ci
    pinned.append((nr, nc, prp, pci))
flags = ur.FLAG_ASSUME_CANONICAL | ur.FLAG_RESULT_NO_COUNT
for _ in range(3):
    res, h = g.train_csr(pinned, w.params, 42, flags, keep=True)
    g.free_result(h)
ts = []
for _ in range(steps):
    t0 = time.perf_counter()
    res, h = g.train_csr(pinned, w.params, 42, flags, keep=True)
    ts.append(time.perf_counter() - t0)
    nnz = [int(r[3][-1]) for r in res]
    g.free_result(h)
ms = float(np.median(ts)) * 1e3
line = {"tool": "group_e2e", "workload": wl, "n_gpus": n, "mode": "single process, cco_create_group, merged result", "steps": steps,
        "ms_per_train_median": ms, "ms_min_max": [min(ts) * 1e3, max(ts) * 1e3], "events_per_s": w.n_events / (ms * 1e-3),
        "out_nnz": nnz, "h2d_bytes": int(sum(m[2].nbytes + m[3].nbytes for m in pinned))}
if check:
    from oracle import oracle as orc
    from oracle import parity as par
    sw = w if w.n_events <= 50_000_000 else synth.make(wl, n_users=w.n_users // 10 if False else synth.CONFIGS[wl]["n_users"] // 10,
                                                          n_events=synth.CONFIGS[wl]["n_events"] // 10)
    got = g.train_csr(sw.mats, sw.params, 42, ur.FLAG_ASSUME_CANONICAL)
    ref = orc.train([orc.Csr(*m) for m in sw.mats], [orc.Params(*p) for p in sw.params], 42, 0, len(os.sched_getaffinity(0)))
    p = par.compare(ref, got, sw.n_users)
    line["parity"] = {k: p[k] for k in ("ok", "counts_exact", "topk_equal", "max_llr_rel", "cells")}
    line["parity"]["sample"] = "full" if sw is w else "generator at 1/10 of users and events (host-generated)"
print(json.dumps(line), flush=True)
g.close()
