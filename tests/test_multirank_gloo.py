"""N > 1 path on CPU: two gloo ranks, each owning half of the genome's windows, produce their run summaries,
rank 0 gathers them (the same gather_object bench.py uses), merges and replays through kgma_replay (host-only,
no device), and the result must equal the oracle's single-pass scan.  The per-shard run summaries are derived
here from the oracle's per-window distances (exact-arithmetic mode) - on the GPU box the CUDA kernels produce
them (tests/test_gpu_parity.py::test_sharded_runs_replay_equals_whole) - so this covers exactly the host side
of the multi-GPU path: shard cut, arbitrary arrival order, merge of runs split at the shard edge, replay."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, TF, GENOME

RUN_DT = np.dtype([("record", "<i4"), ("profile", "<i4"), ("t_first", "<i8"), ("t_last", "<i8"), ("t_argmin", "<i8"),
                   ("D_min", "<i8"), ("flags", "<u4"), ("reserved", "<u4")])
THR = 40.0


def runs_from_D(D, T, rec, t_lo, t_hi):
    """maximal stretches of D[t] < T for loop steps t in [t_lo, t_hi) (1-based steps; D[t-1] is step t)"""
    out = []
    t = max(t_lo, 1)
    while t < t_hi:
        if D[t - 1] < T:
            a = t
            while t < t_hi and D[t - 1] < T:
                t += 1
            seg = D[a - 1:t - 1]
            out.append((rec, 0, a, t - 1, a + int(np.argmin(seg)), int(seg.min()), 0, 0))
        else:
            t += 1
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    RV, ws, cons = K.gen_ref_ws_cons(TF, 6)
    N, k = RV.n_refs, 6
    den = 2 * k * N * N
    T = int(np.ceil(THR * den))
    g = K.Genome.from_fasta(GENOME)
    f = O.Fasta(GENOME)
    with O.exact_arithmetic(N):
        _, _, d = O.ac_gma_testing(GENOME, np.asarray(RV), cons, windowsize=ws, thr=THR, do_align=False, do_return_dists=True)
    # this rank's shard: steps [lo, hi) of every record, cut in the middle of each record (+ rank 1 goes first in the gather)
    runs, firsts, base = [], np.full(len(g), np.iinfo(np.int64).min, dtype=np.int64), 0
    for r in range(len(g)):
        steps = g.seqsize(r) - ws
        D = np.rint(d[base:base + steps] * den).astype(np.int64)
        base += steps
        cut = 6845 if r == 3 else steps // 2 + 1          # record 3: the cut falls inside the run of the hit at 6852:7140
        lo, hi = (1, cut) if rank == 0 else (cut, steps + 1)
        runs += runs_from_D(D, T, r, lo, hi)
        if rank == 0:      # window 0 of every record lives in shard 0
            c = O.kmer_count(f.seq(r)[:ws], k)
            S = np.asarray(RV.S, dtype=np.int64)
            firsts[r] = int(np.sum((N * c.astype(np.int64) - S) ** 2))
    payload = (np.array(runs, dtype=RUN_DT).tobytes(), firsts.tobytes())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank == 0:
        gathered = gathered[::-1]                                  # arrival order must not matter
        allruns = np.concatenate([np.frombuffer(p[0], dtype=np.uint8) for p in gathered])
        fd = np.max(np.stack([np.frombuffer(p[1], dtype=np.int64) for p in gathered]), axis=0)
        out = K.replay_raw(g, [RV], [ws], [cons], [THR], k, K.L.MODE_SINGLE, 50, 0, -69, -1, allruns, fd, host_only=True)
        with O.exact_arithmetic(N):
            oh, _, _ = O.ac_gma_testing(GENOME, np.asarray(RV), cons, windowsize=ws, thr=THR, do_align=False)
        got = [(int(h.record), int(h.first), int(h.last), int(h.genome_pos), float(h.dist)) for h in out.hits]
        want = [(h.record, h.first, h.last, h.genome_pos, h.dist) for h in oh]
        split = sum(1 for a in np.frombuffer(gathered[1][0], dtype=RUN_DT) for b in np.frombuffer(gathered[0][0], dtype=RUN_DT)
                    if a["record"] == b["record"] and a["t_last"] + 1 == b["t_first"])
        q.put((got == want, len(got), split))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_merge_replay():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, nhits, split = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and nhits >= 20
    assert split >= 1                                          # at least one run really was cut at a shard edge and re-joined
