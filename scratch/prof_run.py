"""short run that launches every kernel once or twice on the full-size genome (for ncu)"""
import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
import kmergma_jl_b200 as K
L = K.L
ctx = K.Context(0)
lens = bench.contig_lengths(1.0)
plants = bench.plant_list(lens)
g = K.Genome.synth(lens, seed=42, n_run_len=10000, centromere_len=3000000, ctx=ctx)
for (r, pos, s) in plants: g.put_seq(r, pos, s)
rng = np.random.default_rng(5)
query = "".join(np.asarray(list("ACGT"))[rng.integers(0, 4, size=300)])
for i in range(1000):
    r = int(rng.integers(0, len(lens))); g.put_seq(r, int(rng.integers(20000, lens[r] - 20000)), query)
g.make_resident(ctx)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
for _ in range(2):
    out = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx)
print("single", len(out.hits))
mp = C.POINTER(L.Match)(); n = C.c_int64()
for _ in range(2):
    ctx.check(ctx._lib.kgma_exact_match(ctx._h, g._h, query.encode(), len(query), 1, L.F_RESIDENT, C.byref(mp), C.byref(n)))
print("exact", n.value)
if len(sys.argv) > 1 and sys.argv[1] == "dense":
    out = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_RESIDENT | L.F_DENSE, -69, -1, ctx=ctx)
    print("dense", len(out.hits))
