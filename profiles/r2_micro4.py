"""Round-2 measurement helper: where to split the pipelined streamed scan (KGMA_SPLIT = fraction of the genome in front of the
split; the split falls on the last record start below it).  e2e ms per scan for single and cluster mode.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import kmergma_jl_b200 as K

L = K.L
ctx = K.Context(0)
W = bench.Workload("single", 1.0, "/tmp")
g = K.Genome.synth(W.lens, seed=W.seed, n_run_len=W.n_run, centromere_len=W.centromere, ctx=ctx)
for (r, pos, s) in W.plants:
    g.put_seq(r, pos, s)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
rvs, wss, cs, inv = K.cluster_ref_API(bench.TF, 6)
rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
offs = np.cumsum([0] + list(W.lens))[:-1] / sum(W.lens)
out = {"record_start_fractions": [round(float(x), 3) for x in offs]}


def t(f, n=6):
    f(); f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); o = f(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), int(len(o.hits))


for split in ("default", "0.6", "0.72", "0.8", "0.86", "0.9", "0.93", "0.95", "0.97", "0.99"):
    if split != "default":
        os.environ["KGMA_SPLIT"] = split
    out["single_" + split] = t(lambda: K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN, -69, -1, ctx=ctx))
    out["cluster_" + split] = t(lambda: K.scan_raw(g, rvs, wss, cs, [35, 31, 38, 34, 27, 27], 6, L.MODE_CLUSTER, 100, L.F_ALIGN, -200, -1, ctx=ctx))
    os.environ.pop("KGMA_SPLIT", None)
print(json.dumps(out))
