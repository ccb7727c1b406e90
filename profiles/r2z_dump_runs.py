"""Round-2 helper: dump the run lists (kgma_run records + first-window distances) of the bench's single and cluster workloads, so
that the host replay (merge, event lists, state machine) can be profiled without a GPU.  Writes gpurun_out/r2z_runs_<config>.npz."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import kmergma_jl_b200 as K

L = K.L
ctx = K.Context(0)
for cfg in ("single", "cluster"):
    W = bench.Workload(cfg, 1.0, "/tmp")
    g = K.Genome.synth(W.lens, seed=W.seed, n_run_len=W.n_run, centromere_len=W.centromere, ctx=ctx)
    for (r, pos, s) in W.plants:
        g.put_seq(r, pos, s)
    g.make_resident(ctx)
    rvs, wss, cs, thr = W.profiles(K)
    mode = L.MODE_CLUSTER if cfg == "cluster" else L.MODE_SINGLE
    o = K.scan_raw(g, rvs, wss, cs, thr, 6, mode, W.buff, L.F_RESIDENT, W.gap_open, -1, ctx=ctx, runs_only=True)
    np.savez_compressed(f"gpurun_out/r2z_runs_{cfg}.npz", runs=o.runs, first_D=np.asarray(o.first_D, np.int64), lens=np.asarray(W.lens, np.int64),
                        buff=W.buff, gap_open=W.gap_open, thr=np.asarray(thr, np.float64))
    print(cfg, o.n_runs, "runs")
