import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
import kmergma_jl_b200 as K
L = K.L
ctx = K.Context(0)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lens = bench.contig_lengths(scale)
plants = bench.plant_list(lens, n_plants=max(10, int(2000 * scale)))
g = K.Genome.synth(lens, seed=42, n_run_len=10000, centromere_len=3000000, ctx=ctx)
for (r, pos, s) in plants: g.put_seq(r, pos, s)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
g.make_resident(ctx)
for fl, name in ((L.F_ALIGN | L.F_RESIDENT, "resident+align"), (L.F_RESIDENT, "resident noalign"), (L.F_ALIGN, "e2e+align")):
    for i in range(4):
        t0 = time.perf_counter()
        out = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, fl, -69, -1, ctx=ctx)
        t1 = time.perf_counter()
    st = ctx.stats()
    print(name, "python wall %.3f ms" % ((t1 - t0) * 1e3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
