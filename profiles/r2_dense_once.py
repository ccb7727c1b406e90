"""one dense (KGMA_F_DENSE) scan of the cfg2 genome, resident -- the ncu target for the count-table kernel's dense pass"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import kmergma_jl_b200 as K
L = K.L
ctx = K.Context(0)
W = bench.Workload("single", float(os.environ.get("SCALE", "1.0")), "/tmp")
g = K.Genome.synth(W.lens, seed=W.seed, n_run_len=W.n_run, centromere_len=W.centromere, ctx=ctx)
for (r, pos, s) in W.plants:
    g.put_seq(r, pos, s)
g.make_resident(ctx)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
for _ in range(2):
    o = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_RESIDENT | L.F_DENSE, -69, -1, ctx=ctx)
print(len(o.hits), ctx.stats()["exact_ms"])
