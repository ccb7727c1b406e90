import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
import kmergma_jl_b200 as K
L = K.L
ctx = K.Context(0)
lens = bench.contig_lengths(1.0)
plants = bench.plant_list(lens, n_plants=2000)
g = K.Genome.synth(lens, seed=42, n_run_len=10000, centromere_len=3000000, ctx=ctx)
for (r, pos, s) in plants: g.put_seq(r, pos, s)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
rvs, wss, cs, inv = K.cluster_ref_API(bench.TF, 6)
rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
thr = [35, 31, 38, 34, 27, 27]
def run(mode, fl):
    ts = []
    for i in range(10):
        t0 = time.perf_counter()
        if mode == "single":
            out = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | fl, -69, -1, ctx=ctx)
        else:
            out = K.scan_raw(g, rvs, wss, cs, thr, 6, L.MODE_CLUSTER, 100, L.F_ALIGN | fl, -200, -1, ctx=ctx)
        ts.append((time.perf_counter() - t0) * 1e3)
    st = ctx.stats()
    h = out.hits
    sig = int(np.sum(h["first"].astype(np.int64) * 3 + h["last"].astype(np.int64) * 7 + h["align_score"].astype(np.int64)))
    print("PAIR_MAX=%s" % os.environ.get("KGMA_ALIGN_PAIR_MAX", "0"), mode, "resident" if fl else "e2e", "median %.3f ms" % np.median(ts[3:]), "align_ms %.3f" % st["align_ms"], "hits", len(h), "sig", sig, flush=True)
run("single", 0); run("cluster", 0)
g.make_resident(ctx)
run("single", L.F_RESIDENT); run("cluster", L.F_RESIDENT)
