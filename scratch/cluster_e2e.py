import sys, time, json, os, tempfile
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
def run(label):
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        out = K.scan_raw(g, rvs, wss, cs, thr, 6, L.MODE_CLUSTER, 100, L.F_ALIGN, -200, -1, ctx=ctx)
        ts.append((time.perf_counter() - t0) * 1e3)
    st = ctx.stats()
    print(label, "wall ms", [round(t, 2) for t in ts], "hits", len(out.hits), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
run("pipelined")
os.environ["KGMA_NO_PIPELINE"] = "1"
run("plain")
del os.environ["KGMA_NO_PIPELINE"]
for f in ("0.6", "0.8", "0.9"):
    os.environ["KGMA_SPLIT"] = f
    run("split " + f)
# t2 repeat
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
rec = int(np.argmax(lens))
seq = np.frombuffer(g.seq(rec).encode(), dtype=np.uint8)
width = 80
body = seq[:seq.size // width * width].reshape(-1, width)
with tempfile.NamedTemporaryFile(suffix=".fasta", delete=False) as fh:
    fh.write(b">contig T2 tier\n")
    fh.write(np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).tobytes())
    fh.write(seq[body.size:].tobytes() + b"\n")
    path = fh.name
for i in range(4):
    t0 = time.perf_counter()
    g2 = K.Genome.from_fasta(path)
    t1 = time.perf_counter()
    out2 = K.scan_raw(g2, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN, -69, -1, ctx=ctx)
    t2 = time.perf_counter()
    st = ctx.stats()
    print("t2 parse %.1f scan %.1f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), {k: round(v, 2) for k, v in st.items() if k.endswith("_ms")})
    del g2
