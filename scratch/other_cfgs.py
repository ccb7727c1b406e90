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
# plant a 300-nt query 1000 times (cfg4)
rng = np.random.default_rng(5)
query = "".join(np.asarray(list("ACGT"))[rng.integers(0, 4, size=300)])
qpos = []
for i in range(int(1000 * scale)):
    r = int(rng.integers(0, len(lens))); p = int(rng.integers(20000, lens[r] - 20000))
    g.put_seq(r, p, query); qpos.append((r, p))
g.make_resident(ctx)
def timeit(f, n=5):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): out = f()
    return (time.perf_counter() - t0) / n * 1e3, out
total = g.total_len
# cfg2 dense
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
ms, out = timeit(lambda: K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT | L.F_DENSE, -69, -1, ctx=ctx), 2)
print("cfg2 dense: %.2f ms  %.0f Mb/s hits %d" % (ms, total / ms / 1e3, len(out.hits)), {k: round(v, 3) for k, v in ctx.stats().items() if k.endswith('_ms')})
ms, out2 = timeit(lambda: K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx))
print("cfg2 filtered: %.2f ms hits %d equal_dense %s" % (ms, len(out2.hits), np.array_equal(out.hits[['record','first','last','D']], out2.hits[['record','first','last','D']])))
# cfg3 cluster
rvs, wss, cs, inv = K.cluster_ref_API(bench.TF, 6)
rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
thr = [35, 31, 38, 34, 27, 27]
ms, out = timeit(lambda: K.scan_raw(g, rvs, wss, cs, thr, 6, L.MODE_CLUSTER, 100, L.F_ALIGN | L.F_RESIDENT, -200, -1, ctx=ctx))
st = ctx.stats()
print("cfg3 cluster: %.2f ms  %.0f Mb/s hits %d" % (ms, total / ms / 1e3, len(out.hits)), {k: round(v, 3) for k, v in st.items() if k.endswith('_ms')}, st['blocks_flagged'], st['n_runs'], st['n_align'])
# cfg4 exact match
lib = ctx._lib
import ctypes as C
def em(overlap):
    mp = C.POINTER(L.Match)(); n = C.c_int64()
    ctx.check(lib.kgma_exact_match(ctx._h, g._h, query.encode(), len(query), overlap, L.F_RESIDENT, C.byref(mp), C.byref(n)))
    k = n.value
    if k: lib.kgma_free(mp)
    return k
ms, n = timeit(lambda: em(1))
st = ctx.stats()
print("cfg4 exactMatch: %.2f ms  %.0f Mb/s matches %d planted %d kernel %.3f ms -> %.0f GB/s" % (ms, total / ms / 1e3, n, len(qpos), st['filter_ms'], total * 0.375 / st['filter_ms'] / 1e6))
# cfg5-like: k=7
RV7, ws7, cons7 = K.gen_ref_ws_cons(bench.TF, 7)
ms, out = timeit(lambda: K.scan_raw(g, [RV7], [ws7], [cons7], [30.0], 7, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx))
st = ctx.stats()
print("k=7: %.2f ms  %.0f Mb/s hits %d" % (ms, total / ms / 1e3, len(out.hits)), {k: round(v, 3) for k, v in st.items() if k.endswith('_ms')}, st['blocks_flagged'])
