"""Round-2 measurement helper, extension kernel routes: default (hinted jobs + in-kernel hand-over to the two-chain form),
KGMA_ALIGN_TAIL=all (two chains for every alignment), =off (hand-back to the path-summary kernel: the earlier two-launch flow),
KGMA_ALIGN_KERNEL=summary.  Full-size cfg2 genome, resident; CUDA-event times from kgma_get_stats.  One JSON line."""
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
KEYS = ("filter_ms", "exact_ms", "align_ms", "wall_ms", "host_replay_ms", "n_align", "n_align_redo", "n_align_summary", "n_align_head", "n_runs", "launches")


def scan(n=10, env=None):
    for k_, v in (env or {}).items():
        os.environ[k_] = v
    try:
        K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx)
        acc = {}
        t0 = time.perf_counter()
        for _ in range(n):
            o = K.scan_raw(g, [RV], [ws], [cons], [30.0], 6, L.MODE_SINGLE, 50, L.F_ALIGN | L.F_RESIDENT, -69, -1, ctx=ctx)
            st = ctx.stats()
            for k_ in KEYS:
                acc[k_] = acc.get(k_, 0) + st[k_] / n
        acc["call_ms"] = (time.perf_counter() - t0) / n * 1e3
        acc["hits"] = int(len(o.hits))
        acc["key"] = int(np.bitwise_xor.reduce(o.hits["first"] * 31 + o.hits["last"]))
        return acc
    finally:
        for k_ in (env or {}):
            os.environ.pop(k_, None)


out["default"] = scan()
for c_ in ("1", "2", "3", "4"):
    out["ctas_per_sm_%s" % c_] = scan(env={"KGMA_ALIGN_CTAS": c_})
out["tail_all"] = scan(env={"KGMA_ALIGN_TAIL": "all"})
out["tail_off"] = scan(env={"KGMA_ALIGN_TAIL": "off"})
out["summary_kernel"] = scan(env={"KGMA_ALIGN_KERNEL": "summary"})
print(json.dumps(out))
