"""Round-2 helper: time the host replay (merge of run pieces, event lists, state machine) WITHOUT a GPU on the run lists dumped by
profiles/r2z_dump_runs.py (gpurun_out/r2z_runs_<config>.npz).  usage: python profiles/r2z_replay_offline.py single|cluster"""
import sys, os, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import kmergma_jl_b200 as K
TF = "/root/repo/tests/fixtures/fasta_files/Alp_V_ref.fasta"
L = K.L
lib = L.load()
cfg = sys.argv[1]
d = np.load(f"/root/repo/gpurun_out/r2z_runs_{cfg}.npz")
lens = d["lens"]
h = C.c_void_p(); lib.kgma_genome_create(C.byref(h))
mx = int(lens.max())
w2 = np.zeros((mx + 15) // 16, np.uint32); m = np.zeros((mx + 31) // 32, np.uint32)
for i, n in enumerate(lens):
    assert lib.kgma_genome_append_packed(h, f"synth{i+1}".encode(), f"synth{i+1} x".encode(), w2.ctypes.data, m.ctypes.data, int(n)) == 0
assert lib.kgma_genome_seal(h) == 0
g = K.Genome.__new__(K.Genome); g._h = h; g._lib = lib; g._resident_ctx = None
if cfg == "cluster":
    rvs, wss, cs, inv = K.cluster_ref_API(TF, 6); rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv); mode = L.MODE_CLUSTER
else:
    RV, ws, cons = K.gen_ref_ws_cons(TF, 6); rvs, wss, cs = [RV], [ws], [cons]; mode = L.MODE_SINGLE
thr = list(d["thr"])
runs = d["runs"]; fd = d["first_D"]
def f():
    return K.replay_raw(g, rvs, wss, cs, thr, 6, mode, int(d["buff"]), 0, int(d["gap_open"]), -1, runs, fd, host_only=True)
o = f()
print(cfg, len(runs) // 48, "runs ->", len(o.hits), "hits; key", int(np.bitwise_xor.reduce(o.hits["first"] * 31 + o.hits["last"] + o.hits["D"])) if len(o.hits) else 0)
ts = []
for _ in range(30):
    t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
print("median call %.3f ms, min %.3f" % (np.median(ts), min(ts)))
os.environ["KGMA_TRACE"] = "1"; os.environ["KGMA_TRACE_MERGE"] = "1"
f()
for _ in range(4): f()
