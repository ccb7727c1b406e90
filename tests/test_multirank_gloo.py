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


# ------------------------------------------------------------------------------------------------------------------
# The round-2 flow: every rank ships ONE fixed-size block of (run, extension result) pairs (what kgma_result_pack lays out on
# the GPU box), the ranks exchange the blocks with a single all-gather of equal-sized tensors (NCCL in bench.py, gloo here),
# and rank 0 merges + replays on the host with kgma_replay_packed -- including the alignment step, whose results it looks up
# in the blocks.  The runs come from the oracle's per-window distances and the extension results from the oracle's aligner
# (on the GPU box kgma_scan_shard produces both: tests/test_gpu_parity.py::test_scan_shard_blocks_replayed_on_the_host_equal_whole).
PACK_MAGIC = 0x4b474d41


def pack_block(runs, ext, first_D, cap):
    """the layout of kgma_result_pack (scan.cu PackHdr): magic, version, n_runs, n_first, bytes | runs | ext | first_D"""
    runs = np.asarray(runs, dtype=RUN_DT)
    ext = np.asarray(ext, dtype=np.int64).reshape(-1, 3)
    fd = np.asarray(first_D, dtype=np.int64)
    need = 32 + runs.size * RUN_DT.itemsize + ext.size * 8 + fd.size * 8
    assert need <= cap
    buf = np.zeros(cap, dtype=np.uint8)
    buf[0:8].view(np.uint32)[:] = (PACK_MAGIC, 1)
    buf[8:32].view(np.int64)[:] = (runs.size, fd.size, need)
    o = 32
    for part in (runs.view(np.uint8).reshape(-1), ext.view(np.uint8).reshape(-1), fd.view(np.uint8).reshape(-1)):
        buf[o:o + part.size] = part
        o += part.size
    return buf


def _worker_packed(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    RV, ws, cons = K.gen_ref_ws_cons(TF, 6)
    N, k, buff = RV.n_refs, 6, 50
    den = 2 * k * N * N
    T = int(np.ceil(THR * den))
    g = K.Genome.from_fasta(GENOME)
    f = O.Fasta(GENOME)
    with O.exact_arithmetic(N):
        _, _, d = O.ac_gma_testing(GENOME, np.asarray(RV), cons, windowsize=ws, thr=THR, do_align=False, do_return_dists=True)
    runs, ext, firsts, base = [], [], np.full(len(g), np.iinfo(np.int64).min, dtype=np.int64), 0
    for r in range(len(g)):
        L = g.seqsize(r)
        steps = L - ws
        D = np.rint(d[base:base + steps] * den).astype(np.int64)
        base += steps
        cut = 6845 if r == 3 else steps // 2 + 1          # record 3: the cut falls inside the run of the hit at 6852:7140
        lo, hi = (1, cut) if rank == 0 else (cut, steps + 1)
        seq = f.seq(r)
        for run in runs_from_D(D, T, r, lo, hi):
            runs.append(run)
            cmi = k + run[4]                               # GenomeMiner.jl:85,92: CMI = i_left + 1 = k + t_argmin
            a, b = max(cmi - buff, 1), min(cmi + ws - 1 + buff, L)
            cig, score = O.pairalign_semiglobal(cons[:ws], seq[a - 1:b], -69, -1)
            l_, h_ = O.cigar_to_UnitRange(cig)
            ext.append((l_, h_, score))                    # this rank extends its own runs' candidate windows
        if rank == 0:
            c = O.kmer_count(seq[:ws], k)
            S = np.asarray(RV.S, dtype=np.int64)
            firsts[r] = int(np.sum((N * c.astype(np.int64) - S) ** 2))
    cap = 1 << 16
    mine = torch.from_numpy(pack_block(runs, ext, firsts, cap))
    allb = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(allb, mine)                            # ONE collective of equal-sized blocks
    # the same blocks through the host shared-memory exchange bench.py uses on the GPU box (kmergma.jl_b200/exchange.py):
    # several steps, so that both slots and the two-steps-behind wait are exercised
    name = "kgma_xch_t%d" % port
    x = K.HostExchange(name, 0, world, cap, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        x = K.HostExchange(name, rank, world, cap, create=False)
    shm_ok = True
    for step in range(5):
        x.begin_step()
        x.block()[:cap] = mine.numpy()
        x.publish()
        if rank == 0:
            base = x.wait_all()
            o2 = K.replay_packed(g, [RV], [ws], [cons], [THR], k, K.L.MODE_SINGLE, buff, K.L.F_ALIGN, -69, -1, base, world, x.cap, ctx=None)
            x.consumed()
            shm_ok = shm_ok and len(o2.hits) > 0
            last_shm = [(int(h.record), int(h.first), int(h.last), int(h.genome_pos), float(h.dist)) for h in o2.hits]
    dist.barrier()
    x.close()
    if rank == 0:
        blocks = np.stack([b.numpy() for b in allb[::-1]])  # arrival order must not matter
        out = K.replay_packed(g, [RV], [ws], [cons], [THR], k, K.L.MODE_SINGLE, buff, K.L.F_ALIGN, -69, -1,
                              blocks.ctypes.data, world, cap, ctx=None)
        with O.exact_arithmetic(N):
            oh, _, _ = O.ac_gma_testing(GENOME, np.asarray(RV), cons, windowsize=ws, thr=THR, buff=buff, do_align=True)
        got = [(int(h.record), int(h.first), int(h.last), int(h.genome_pos), float(h.dist)) for h in out.hits]
        want = [(h.record, h.first, h.last, h.genome_pos, h.dist) for h in oh]
        # a block that does not belong to this scan must be refused, not mis-read
        bad = blocks.copy(); bad[0, 0] ^= 0xFF
        try:
            K.replay_packed(g, [RV], [ws], [cons], [THR], k, K.L.MODE_SINGLE, buff, K.L.F_ALIGN, -69, -1, bad.ctypes.data, world, cap, ctx=None)
            refused = False
        except K.KmerGMAError:
            refused = True
        q.put((got == want and shm_ok and last_shm == want, len(got), refused, any(h.last - h.first + 1 != ws + 2 * buff for h in oh)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_packed_blocks_with_extension():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, nhits, refused, aligned = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and nhits >= 20 and refused and aligned
