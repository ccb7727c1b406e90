"""Round-2 measurement helper: cluster mode (6 profiles) with fewer prefilter passes -- KGMA_LOAD_OK lets more profiles share one
weight table (the table holds the maximum weight over the group, so it flags more blocks; the count-table kernel checks every
profile's own bound before it builds a table).  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import kmergma_jl_b200 as K

L = K.L
ctx = K.Context(0)
W = bench.Workload("cluster", 1.0, "/tmp")
g = K.Genome.synth(W.lens, seed=W.seed, n_run_len=W.n_run, centromere_len=W.centromere, ctx=ctx)
for (r, pos, s) in W.plants:
    g.put_seq(r, pos, s)
g.make_resident(ctx)
rvs, wss, cs, thr = W.profiles(K)
out = {}
for lo in ("0.62", "0.7", "0.78", "0.85", "0.95", "1.2"):
    os.environ["KGMA_LOAD_OK"] = lo
    f = lambda: K.scan_raw(g, rvs, wss, cs, thr, 6, L.MODE_CLUSTER, 100, L.F_ALIGN | L.F_RESIDENT, -200, -1, ctx=ctx)
    f(); f()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); o = f(); ts.append((time.perf_counter() - t0) * 1e3)
    st = ctx.stats()
    out[lo] = {"ms": float(np.median(ts)), "hits": int(len(o.hits)), "passes": int(st["filter_passes"]), "filter_ms": st["filter_ms"], "exact_ms": st["exact_ms"],
               "align_ms": st["align_ms"], "blocks_flagged": int(st["blocks_flagged"]), "key": int(np.bitwise_xor.reduce(o.hits["first"] * 31 + o.hits["last"]))}
print(json.dumps(out))
