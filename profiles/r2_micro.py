"""Round-2 measurement helper (run on the GPU box through gpurun): per-kernel timings on the full-size cfg2 genome for the
variants the round compares -- count-table kernel (parallel slide with 16 / 24 warps per CTA, serial slide), extension kernel
(tagged vs path-summary), prefilter -- each through the C ABI with CUDA-event times from kgma_get_stats.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
out = {}


def scan(flags, n=3, env=None):
    for k_, v in (env or {}).items():
        os.environ[k_] = v
    try:
        K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, flags | L.F_RESIDENT, -69, -1, ctx=ctx)
        acc = {}
        t0 = time.perf_counter()
        for _ in range(n):
            o = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, flags | L.F_RESIDENT, -69, -1, ctx=ctx)
            st = ctx.stats()
            for k_ in ("filter_ms", "exact_ms", "align_ms", "wall_ms", "host_replay_ms", "n_align", "n_align_redo", "n_runs", "blocks_flagged"):
                acc[k_] = acc.get(k_, 0) + st[k_] / n
        acc["call_ms"] = (time.perf_counter() - t0) / n * 1e3
        acc["hits"] = int(len(o.hits))
        return acc
    finally:
        for k_ in (env or {}):
            os.environ.pop(k_, None)


out["filtered_tagged"] = scan(L.F_ALIGN, 10)
out["filtered_summary_kernel"] = scan(L.F_ALIGN, 10, {"KGMA_ALIGN_KERNEL": "summary"})
out["filtered_serial_eval"] = scan(L.F_ALIGN, 10, {"KGMA_EVAL_KERNEL": "serial"})
out["filtered_8mer_prefilter"] = scan(L.F_ALIGN, 10, {"KGMA_PREFILTER": "8mer"})
for w in ("24", "20", "16", "12"):
    out["dense_parallel_slide_%s_warps" % w] = scan(L.F_DENSE, 2, {"KGMA_EVAL_WARPS": w})
out["dense_serial_slide"] = scan(L.F_DENSE, 1, {"KGMA_EVAL_KERNEL": "serial"})
total = g.total_len
for k_, v in out.items():
    if k_.startswith("dense"):
        v["Mb_per_s"] = total / v["call_ms"] / 1e3
print(json.dumps(out))
