"""Round-2 measurement helper: the extension kernel alone (kgma_align_batch) over batches of 74 .. 14208 alignments of
289 x 389 cut from the cfg2 genome's hits -- does the time follow the batch size (throughput bound) or stay at the latency of
one alignment?  One JSON line; CUDA-event times from kgma_get_stats."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import kmergma_jl_b200 as K

L = K.L
ctx = K.Context(0)
W = bench.Workload("single", float(os.environ.get("SCALE", "0.25")), "/tmp")
g = K.Genome.synth(W.lens, seed=W.seed, n_run_len=W.n_run, centromere_len=W.centromere, ctx=ctx)
for (r, pos, s) in W.plants:
    g.put_seq(r, pos, s)
g.make_resident(ctx)
RV, ws, cons = K.gen_ref_ws_cons(bench.TF, 6)
o = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_RESIDENT, -69, -1, ctx=ctx)     # unextended windows: cmi-50 .. cmi+ws-1+50
h = o.hits
rec0, first0, last0 = h["record"].astype(np.int32), h["first"].astype(np.int64), h["last"].astype(np.int64)
c = cons[:ws].encode()
out = {"hits": int(len(h))}
for env in ({}, {"KGMA_ALIGN_TAIL": "all"}, {"KGMA_ALIGN_TAIL": "off"}, {"KGMA_ALIGN_KERNEL": "summary"}):
    for k_, v in env.items():
        os.environ[k_] = v
    res = {}
    for n in (74, 148, 296, 592, 1184, 1776, 2368, 3552, 7104, 14208):
        idx = np.arange(n) % len(h)
        rec, first, last = np.ascontiguousarray(rec0[idx]), np.ascontiguousarray(first0[idx]), np.ascontiguousarray(last0[idx])
        of, ol, sc = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        ms = []
        for rep in range(4):
            ctx.check(ctx._lib.kgma_align_batch(ctx._h, g._h, c, len(c), -69, -1, 0, n, rec.ctypes.data, first.ctypes.data, last.ctypes.data,
                                                of.ctypes.data, ol.ctypes.data, sc.ctypes.data))
            st = ctx.stats()
            ms.append(st["align_ms"])
        res[str(n)] = {"align_ms": float(np.median(ms[1:])), "two_sweeps": int(st["n_align_redo"]), "summary": int(st["n_align_summary"]), "head": int(st["n_align_head"])}
    out["+".join("%s=%s" % kv for kv in env.items()) or "default"] = res
    for k_ in env:
        os.environ.pop(k_, None)
print(json.dumps(out))
