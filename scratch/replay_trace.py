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
g.make_resident(ctx)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
for i in range(5):
    K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx)
os.environ["KGMA_TRACE"] = "1"
for i in range(3):
    t0 = time.perf_counter()
    K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx)
    print("wall %.3f" % ((time.perf_counter() - t0) * 1e3), {k: round(v, 3) for k, v in ctx.stats().items() if k.endswith("_ms")}, flush=True)
