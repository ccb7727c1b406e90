import sys, time, json, os, tempfile
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
import kmergma_jl_b200 as K
L = K.L
ctx = K.Context(0)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lens = bench.contig_lengths(scale)
plants = bench.plant_list(lens, n_plants=int(2000 * scale))
g = K.Genome.synth(lens, seed=42, n_run_len=10000, centromere_len=3000000, ctx=ctx)
for (r, pos, s) in plants: g.put_seq(r, pos, s)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
t0 = time.perf_counter()
path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "t2_full.fasta")
with open(path, "wb") as fh:
    for r in range(len(lens)):
        seq = np.frombuffer(g.seq(r).encode(), dtype=np.uint8)
        body = seq[:seq.size // 80 * 80].reshape(-1, 80)
        fh.write(b">contig%d T2 tier\n" % r)
        fh.write(np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).tobytes())
        fh.write(seq[body.size:].tobytes() + b"\n")
print("wrote %.2f GB in %.1f s" % (os.path.getsize(path) / 1e9, time.perf_counter() - t0), flush=True)
ref = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN, -69, -1, ctx=ctx)
def trial(label):
    t0 = time.perf_counter()
    g2 = K.Genome.from_fasta(path)
    t1 = time.perf_counter()
    ts = []
    for i in range(3):
        ta = time.perf_counter()
        out = K.scan_raw(g2, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN, -69, -1, ctx=ctx)
        ts.append((time.perf_counter() - ta) * 1e3)
        if i == 0: st = ctx.stats()
    same = np.array_equal(out.hits[["record", "first", "last", "D"]], ref.hits[["record", "first", "last", "D"]])
    print(label, "parse %.1f ms | scans %s | equal %s | first-scan stats %s" % ((t1 - t0) * 1e3, [round(t, 1) for t in ts], same,
          {k: round(v, 2) for k, v in st.items() if k.endswith("_ms")}), flush=True)
    t0 = time.perf_counter(); del g2; print("   destroy %.1f ms" % ((time.perf_counter() - t0) * 1e3))
trial("staged")
trial("staged")
os.environ["KGMA_NO_STAGING"] = "1"
trial("pin-first")
trial("pin-first")
os.unlink(path)
