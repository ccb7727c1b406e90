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
rvs, wss, cs, inv = K.cluster_ref_API(bench.TF, 6)
rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
thr = [35, 31, 38, 34, 27, 27]
for i in range(3):
    K.scan_raw(g, rvs, wss, cs, thr, 6, L.MODE_CLUSTER, 100, L.F_ALIGN, -200, -1, ctx=ctx)
os.environ["KGMA_TRACE"] = "1"
for lab in ("pipelined", "plain"):
    if lab == "plain": os.environ["KGMA_NO_PIPELINE"] = "1"
    print("----", lab, flush=True)
    t0 = time.perf_counter()
    K.scan_raw(g, rvs, wss, cs, thr, 6, L.MODE_CLUSTER, 100, L.F_ALIGN, -200, -1, ctx=ctx)
    print("wall %.2f" % ((time.perf_counter() - t0) * 1e3), flush=True)
