#!/usr/bin/env python
"""bench.py — findGenes throughput (Mb/s = 1e6 genome bases scanned per second) on the synthetic genomes of
BASELINE.json: the headline is configs[1], ONE 3.1 Gb genome (24 hg38-like contigs, N runs, 2000 planted IGHV
homologues, single profile, k = 6, thr = 30, buffer 50, gap (-69,-1), do_align = true) at 1/2/4/8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config single|cluster|exact|k7]

One step = one complete operator call over the whole genome (prefilter + count-table kernel + run compaction +
batched extension + replay of the reference's hit state machine).

N = 1: `value` = genome resident in HBM; `e2e` = the same call from pinned pre-packed HOST planes (H2D of the 2-bit genome +
kernels + D2H of run lists / results inside the timed region).
N > 1 (torchrun, one rank per GPU): the SAME single genome, cut into N equal slices of the packed coordinate space with
window-length halos (`scaling: "strong"`).  Every rank scans its slice, merges its runs and extends, on its own GPU, the
candidate window of every run that can still become a hit (kgma_scan_shard); the ranks exchange equal-sized blocks of
(run, extension result) pairs with ONE NCCL all-gather; rank 0 merges and replays them on the host (kgma_replay_packed:
no device work after the exchange).  The N-independent-replicas number of round 1 is kept as the side key `replicas`.

  --config cluster : BASELINE configs[2], findGenes_cluster_mode's operator (5 clusters + average, buffer 100, gap (-200,-1))
  --config exact   : BASELINE configs[3], exactMatch of a 300-nt query planted 1000 times
  --config k7      : BASELINE configs[4], k = 7, 500-member family, 1e10 nt in 2000 log-uniform contigs, streamed
  --impl reference : the reference's algorithm for the same config on the box's host cores -- the C oracle port
                     (oracle/kmergma_oracle.c; Julia is not installed), one task per contig piece on every host thread,
                     the WHOLE genome per step, extension on hits included.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

FIX = os.path.join(ROOT, "tests", "fixtures")
TF = os.path.join(FIX, "fasta_files", "Alp_V_ref.fasta")

# hg38-like contig lengths in Mb: sum ~ 3.1e9
CONTIG_MB = [248.9, 242.2, 198.3, 190.2, 181.5, 170.8, 159.3, 145.1, 138.4, 133.8, 135.1, 133.3, 114.4, 107.0,
             102.0, 90.3, 83.3, 80.4, 58.6, 64.4, 46.7, 50.8, 156.0, 57.2]
SEED = 42
N_RUN = 10_000          # N run at both contig ends
CENTROMERE = 3_000_000  # one N run per contig
N_PLANTS = 2000
THR, BUFF, GAP_OPEN, GAP_EXT, KMER = 30.0, 50, -69, -1, 6
CLUSTER_THR, CLUSTER_BUFF, CLUSTER_GAP = [35, 31, 38, 34, 27, 27], 100, -200
UNIT = "Mb/s"
METRICS = {
    "single": "findGenes throughput, 3.1 Gb synthetic genome, single profile, k=6",
    "cluster": "findGenes_cluster_mode throughput, 3.1 Gb synthetic genome, 5 clusters + average profile, k=6",
    "exact": "exactMatch throughput, 300-nt query, 3.1 Gb synthetic genome",
    "k7": "findGenes throughput, 10 Gb synthetic multi-contig genome, 500-member family, k=7",
}


def contig_lengths(scale: float):
    return [max(100_000, int(round(mb * 1e6 * scale))) // 128 * 128 + 77 for mb in CONTIG_MB]


def record_offsets(lens):
    """global (padded) base offset of every record: records start at multiples of 128 (kgma_internal.h REC_ALIGN)"""
    offs, end = [], 0
    for L in lens:
        off = (end + 127) // 128 * 128
        offs.append(off)
        end = off + L
    return offs


def read_refs():
    refs, cur = [], []
    with open(TF) as fh:
        for line in fh:
            if line.startswith(">"):
                if cur:
                    refs.append("".join(cur).upper())
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        refs.append("".join(cur).upper())
    return refs


def _mutate(rng, s, rate, indel):
    s = list(s)
    for j in range(len(s)):
        if rng.random() < rate:
            s[j] = "ACGT"[int(rng.integers(0, 4))]
    if indel:
        for _ in range(int(rng.integers(1, 3))):
            p = int(rng.integers(1, len(s) - 1))
            n = int(rng.integers(1, 7))
            if rng.random() < 0.5:
                del s[p:p + n]
            else:
                s[p:p] = list("ACGT"[int(rng.integers(0, 4))] * n)
    return "".join(s)


def plant_list(lens, n_plants=N_PLANTS, seed=1234, family=None, n_run=N_RUN):
    """(record, 1-based position, residues) of the planted homologues: members of the family (default: the 84 fixture refs),
    substitution rates {0,2,5,10,15,20}%, 1-6 nt indels in 25%; every 50th copy is flush against a 2^26-aligned packed
    boundary (64 Mi bases) so that shard, chunk and segment edges are exercised."""
    rng = np.random.default_rng(seed)
    refs = family or read_refs()
    w = np.asarray(lens, dtype=np.float64)
    w /= w.sum()
    out = []
    for i in range(n_plants):
        r = int(rng.choice(len(lens), p=w))
        s = _mutate(rng, refs[int(rng.integers(0, len(refs)))], [0.0, 0.02, 0.05, 0.10, 0.15, 0.20][i % 6], i % 4 == 0)
        L = lens[r]
        nr = min(n_run, L // 4)
        lo, hi = nr + 1, L - nr - len(s)
        if hi <= lo:
            continue
        pos = int(rng.integers(lo, hi))
        if i % 50 == 0:
            b = (pos >> 26) << 26
            if b - 150 > lo:
                pos = b - 150            # straddles a 64 Mi-base boundary
        out.append((r, pos, s))
    return out


# ---- BASELINE configs[4]: k = 7, a large family, 1e10 nt in 2000 contigs ---------------------------------------------
def k7_family(tmpdir, n=500, seed=21):
    """a 500-member family re-mutated from the 84 fixture references, written as FASTA; returns (path, members)"""
    rng = np.random.default_rng(seed)
    base = read_refs()
    fam = [_mutate(rng, base[i % len(base)], float(rng.uniform(0, 0.08)), i % 9 == 0) for i in range(n)]
    path = os.path.join(tmpdir, "k7_family_%d.fasta" % os.getpid())
    with open(path, "w") as fh:
        for i, s in enumerate(fam):
            fh.write(">fam%d\n%s\n" % (i, s))
    return path, fam


def k7_contig_lengths(n_contigs=2000, total=1.0e10, seed=9):
    """log-uniform 10 kb .. 100 Mb, rescaled so that the lengths add up to `total` (never below 10 kb)"""
    rng = np.random.default_rng(seed)
    raw = np.exp(rng.uniform(np.log(1.0e4), np.log(1.0e8), size=n_contigs))
    lens = np.maximum(1.0e4, raw * (total / raw.sum()))
    lens = np.maximum(1.0e4, lens * (total / lens.sum()))
    return [int(x) // 128 * 128 + 77 for x in lens]


def k7_plant_list(lens, fam, n_plants):
    return plant_list(lens, n_plants=n_plants, seed=4321, family=fam, n_run=1000)


def k7_threshold(RV, ws, k=7, seed=42, trials=100, buffer=8.0):
    """estimate_optimal_threshold (DistanceTesting.jl:8-17) with numpy's generator instead of Julia's: the mean k-mer distance
    of `trials` random sequences of the window length to the profile, minus the reference's default buffer 8 (rounded to two
    decimals so that both bench arms and the tests use the very same number)"""
    rng = np.random.default_rng(seed)
    rv = np.asarray(RV, dtype=np.float64)
    tot = 0.0
    for _ in range(trials):
        codes = rng.integers(0, 4, size=int(ws))
        n = codes.size - k + 1
        idx = np.zeros(n, dtype=np.int64)
        for j in range(k):
            idx = idx * 4 + codes[j:j + n]
        c = np.bincount(idx, minlength=4 ** k).astype(np.float64)
        tot += (1.0 / (2 * k)) * float(np.sum((c - rv) ** 2))
    return float(np.round(tot / trials - buffer, 2))


_NVML = {}


def nvml_handle(device_index):
    """NVML is initialised once, outside every timed region (nvmlInit alone takes tens of milliseconds)"""
    if device_index not in _NVML:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            _NVML[device_index] = (pynvml, h, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), None)
        except Exception as e:       # clocks are evidence, not a dependency
            _NVML[device_index] = (None, None, None, str(e))
    return _NVML[device_index]


def clocks_sampler(device_index, stop, samples):
    """sample SM clock + throttle reasons while the timed regions run (B200_PROFILING.md clocks line)"""
    pynvml, h, mx, err = nvml_handle(device_index)
    if pynvml is None:
        samples.append(("error", err, 0))
        return
    try:
        while not stop.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            samples.append((sm, mx, int(rs)))
            time.sleep(0.0005)
    except Exception as e:
        samples.append(("error", str(e), 0))


def summarise_clocks(samples):
    good = [s for s in samples if s[0] != "error"]
    if not good:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: %s" % (samples[0][1] if samples else "no samples")]}
    names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
    bits = 0
    for s in good:
        bits |= s[2]
    return {"sm_mhz": float(np.median([s[0] for s in good])), "sm_max_mhz": float(good[0][1]),
            "reasons": [n for b, n in names.items() if bits & b and n != "gpu_idle"], "samples": len(good)}


# ------------------------------------------------------------------------------------------------
# workloads: everything both arms need to agree on (genome layout, plants, profiles, parameters)
class Workload:
    def __init__(self, config, scale, tmpdir):
        self.config = config
        self.k = 7 if config == "k7" else 6
        self.seed = SEED
        if config == "k7":
            self.lens = k7_contig_lengths(max(8, int(2000 * min(1.0, scale * 4))), 1.0e10 * scale)
            self.fam_path, fam = k7_family(tmpdir)
            self.plants = k7_plant_list(self.lens, fam, max(10, int(4000 * scale)))
            self.n_run, self.centromere = 1000, 100_000
            self.ref_path = self.fam_path
        else:
            self.lens = contig_lengths(scale)
            self.plants = plant_list(self.lens, n_plants=max(10, int(N_PLANTS * scale)))
            self.n_run, self.centromere = N_RUN, CENTROMERE
            self.ref_path = TF
        self.offs = record_offsets(self.lens)
        self.total = sum(self.lens)
        self._by_rec = None
        self.query = None
        if config == "exact":
            rng = np.random.default_rng(5)
            self.query = "".join(np.asarray(list("ACGT"))[rng.integers(0, 4, size=300)])
            for i in range(1000):
                r = int(rng.integers(0, len(self.lens)))
                self.plants.append((r, int(rng.integers(20000, self.lens[r] - 20000)), self.query))
        self.buff = CLUSTER_BUFF if config == "cluster" else BUFF
        self.gap_open = CLUSTER_GAP if config == "cluster" else GAP_OPEN

    def profiles(self, M):
        """(refVecs, windowsizes, consensus_seqs, thresholds) through module M (the product mirror or the oracle)"""
        if self.config == "cluster":
            r = M.cluster_ref_API(self.ref_path, self.k)
            rvs, wss, cs = M.eliminate_null_params(*r[:4]) if hasattr(M, "eliminate_null_params") else _drop_null(*r[:4])
            return list(rvs), list(wss), list(cs), list(CLUSTER_THR)
        RV, ws, cons = M.gen_ref_ws_cons(self.ref_path, self.k)[:3]
        thr = k7_threshold(RV, ws) if self.config == "k7" else THR
        return [RV], [ws], [cons], [thr]

    def describe(self, world):
        base = {"single": "findGenes (ac_gma_testing! operator), single profile k=6 ws=289 N=84, thr=30, buffer 50, gap (-69,-1), do_align=true (BASELINE configs[1])",
                "cluster": "findGenes_cluster_mode (Omn_KmerGMA! operator), 5 clusters + average profile, thr [35,31,38,34,27,27], buffer 100, gap (-200,-1), align_hits=true (BASELINE configs[2])",
                "exact": "exactMatch of a 300-nt query planted 1000 times, overlap=true (BASELINE configs[3])",
                "k7": "findGenes (ac_gma_testing! operator), k=7, 500-member family, estimated threshold, buffer 50, gap (-69,-1), do_align=true (BASELINE configs[4])"}[self.config]
        return "%s; ONE %.2f Gb synthetic genome (%d contigs, N runs, %d planted copies)" % (base, self.total / 1e9, len(self.lens), len(self.plants))


def _drop_null(kfvs, wss, cs, inv):
    keep = [i for i, x in enumerate(inv) if not x]
    return [kfvs[i] for i in keep], [wss[i] for i in keep], [cs[i] for i in keep]


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on host cores
def gen_piece(O, W, rec, start, n):
    """n residues of record `rec` from 0-based `start`: generator + N runs + planted copies, as kgma_genome_synth + put_seq"""
    buf = np.frombuffer(O.synth(W.seed, W.offs[rec] + start, n), dtype=np.uint8).copy()
    L = W.lens[rec]
    nr = min(W.n_run, L // 4)
    runs = [(0, nr), (L - nr, L)]
    if W.centromere > 0 and L > 4 * W.centromere:
        c0 = (L * 2 // 5) // 32 * 32
        runs.append((c0, c0 + W.centromere))
    for a, b in runs:                                    # same N runs as kgma_genome_synth
        lo, hi = max(a, start), min(b, start + n)
        if hi > lo:
            buf[lo - start:hi - start] = ord("N")
    if W._by_rec is None:
        W._by_rec = {}
        for pl in W.plants:
            W._by_rec.setdefault(pl[0], []).append(pl)
    for (r, pos, s) in W._by_rec.get(rec, ()):
        if pos - 1 + len(s) > start and pos - 1 < start + n:
            lo, hi = max(pos - 1, start), min(pos - 1 + len(s), start + n)
            buf[lo - start:hi - start] = np.frombuffer(s.encode(), dtype=np.uint8)[lo - (pos - 1):hi - (pos - 1)]
    return buf


def cpu_pieces(W, piece, overlap):
    """(record, start, length) tasks covering every record, `overlap` bases shared between consecutive pieces of a record
    (a window-length halo, so that every window start is scanned by exactly one task)"""
    out = []
    for r, L in enumerate(W.lens):
        s = 0
        while s < L:
            n = min(piece + overlap, L - s)
            out.append((r, s, n))
            if s + n >= L:
                break
            s += piece
    return out


def oracle_task(O, W, prof, f):
    """one oracle call over an in-memory piece: the reference's own loop incl. the extension of its hits"""
    rvs, wss, cs, thr = prof
    if W.config == "exact":
        res = O.exactMatch(W.query, f)
        return 0 if res == "no match" else sum(len(v) for v in res.values())
    if W.config == "cluster":
        return len(O.Omn_KmerGMA(f, [np.asarray(v) for v in rvs], wss, cs, k=W.k, thr_vec=thr, buff=W.buff, align_hits=True,
                                 gap_open_score=W.gap_open, gap_extend_score=GAP_EXT, hit_cap=1 << 14)[0])
    return len(O.ac_gma_testing(f, np.asarray(rvs[0]), cs[0], k=W.k, windowsize=wss[0], thr=thr[0], buff=W.buff, do_align=True,
                                gap_open_score=W.gap_open, gap_extend_score=GAP_EXT, hit_cap=1 << 14)[0])


def cpu_baseline(W, seconds=12.0):
    """single-thread oracle (scan + extension of its hits) on consecutive 100 Mb pieces until ~`seconds` of CPU work"""
    from oracle import oracle as O
    prof = W.profiles(O)
    spent, bases, hits = 0.0, 0, 0
    for (rec, start, n) in cpu_pieces(W, 100_000_000, 0):
        f = O.Fasta.wrap("piece", gen_piece(O, W, rec, start, n))
        t0 = time.perf_counter()
        hits += oracle_task(O, W, prof, f)
        spent += time.perf_counter() - t0
        bases += n
        if spent >= seconds:
            break
    return {"value": bases / spent / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d Mb of the same synthetic genome (N runs + planted copies included), the oracle's %s loop with extension of "
                      "its hits, %d hits, %.1f s" % (bases // 1_000_000, {"cluster": "Omn_KmerGMA!", "exact": "exactMatch"}.get(W.config, "ac_gma_testing!"), hits, spent)}


def run_reference(args):
    """--impl reference: the oracle over the WHOLE genome of the config per step, one task per contig piece (with a
    window-length overlap) over all host threads -- the reference's only parallel strategy is one task per FASTA record
    (MultiThread/GenomeMiner.jl:127-140) -- extension of the hits included."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    W = Workload(args.config, args.scale, tempfile.gettempdir())
    prof = W.profiles(O)
    T = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    halo = (300 if W.config == "exact" else max(prof[1])) - 1
    pieces = cpu_pieces(W, 24_000_000, halo)
    t_gen = time.perf_counter()
    with ThreadPoolExecutor(T) as ex:
        fastas = list(ex.map(lambda p: O.Fasta.wrap("piece r%d@%d" % (p[0], p[1]), gen_piece(O, W, *p)), pieces))
        t_gen = time.perf_counter() - t_gen
        order = sorted(range(len(fastas)), key=lambda i: -pieces[i][2])

        def step():
            return sum(ex.map(lambda i: oracle_task(O, W, prof, fastas[i]), order))

        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        nh = 0
        for _ in range(args.steps):
            nh = step()
        dt = time.perf_counter() - t0
    val = W.total * args.steps / dt / 1e6
    sample = ("the whole %.2f Gb genome per step: %d pieces of <= 24 Mb (+ %d-base halo), one oracle task per piece on %d threads, "
              "scan and extension of every hit (inputs generated in memory beforehand, %.1f s)" % (W.total / 1e9, len(pieces), halo, T, t_gen))
    print(json.dumps({
        "impl": "reference", "metric": METRICS[W.config], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": W.describe(1),
                   "note": "Julia is not installed: the reference's algorithm is timed as the C oracle port (oracle/kmergma_oracle.c), "
                           "inputs in memory; pieces are scanned as independent records (hits at piece edges may differ from the whole-contig "
                           "scan; the work is the same)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": T, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "hits_per_step": int(nh)}))


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(device_index):
    """pin this rank to the CPUs next to its GPU before any pinned host memory is allocated (first touch decides the
    NUMA node of the staging buffers; with 8 ranks streaming their slices the host links are the bottleneck)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


KEYS = ("launches", "h2d_bytes", "d2h_bytes", "filter_ms", "exact_ms", "align_ms", "total_ms", "h2d_ms", "blocks_total",
        "blocks_flagged", "exact_windows", "wall_ms", "host_setup_ms", "host_cand_ms", "host_replay_ms", "n_runs", "n_align",
        "n_align_redo", "filter_passes", "n_align_summary", "n_align_head", "rank0_replay_ms")


def run_ours(args):
    # rank 0 must print exactly ONE line on stdout, but NCCL writes its version banner there from C: park the real stdout
    # and send everything else (python prints, library chatter) to stderr until the JSON line goes out
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import ctypes as C
    import tempfile
    import torch
    import torch.distributed as dist
    import kmergma_jl_b200 as K
    L = K.L
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t_cold = time.perf_counter()
    ctx = K.Context(local)
    t_ctx = time.perf_counter() - t_cold
    W = Workload(args.config, args.scale, tempfile.gettempdir())
    rvs, wss, cs, thr = W.profiles(K)
    mode = L.MODE_CLUSTER if W.config == "cluster" else L.MODE_SINGLE
    prof_args = (rvs, wss, cs, thr, W.k, mode, W.buff)
    exact = W.config == "exact"
    prep = [None]          # the operator call marshalled once (PreparedScan): the timed loops pay the C-ABI calls only

    def make_genome(seed):
        g_ = K.Genome.synth(W.lens, seed=seed, n_run_len=W.n_run, centromere_len=W.centromere, ctx=ctx)
        for (r, pos, s) in W.plants:
            g_.put_seq(r, pos, s)
        return g_

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def exact_call(g_, resident, shard=None):
        if shard is None:
            mp = C.POINTER(L.Match)(); n = C.c_int64()
            ctx.check(ctx._lib.kgma_exact_match(ctx._h, g_._h, W.query.encode(), len(W.query), 1, L.F_RESIDENT if resident else 0, C.byref(mp), C.byref(n)))
            if n.value:
                ctx._lib.kgma_free(mp)
            return n.value
        return K.exact_match_shard(W.query, g_, shard, ctx=ctx, resident=resident)

    # ---- the exchange of a sharded step: every rank packs its block straight into a shared-memory segment of the host and
    #      publishes the step number; rank 0 replays the blocks in place (kmergma.jl_b200/exchange.py).  No collective, no
    #      device round trip on the path; NCCL only sizes the segment (an untimed step) and closes the timed regions.
    from kmergma_jl_b200 import HostExchange

    class Exchange:
        def __init__(self):
            self.cap, self.x, self.gen = 0, None, 0

        def size(self, need):
            t = torch.tensor([need], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cap = (2 * int(t.item()) + 4096 + 4095) // 4096 * 4096
            if cap > self.cap:
                if self.x:
                    self.x.close()
                self.gen += 1
                name = "kgma_xch_%s_%d" % (os.environ.get("MASTER_PORT", "0"), self.gen)
                if rank == 0:
                    self.x = HostExchange(name, 0, world, cap, create=True)
                dist.barrier()
                if rank != 0:
                    self.x = HostExchange(name, rank, world, cap, create=False)
                dist.barrier()
                self.cap = self.x.cap

    xch = Exchange() if world > 1 else None

    # N > 1, two ways to cut the one genome (north_star: "by contig/chunk with window-length halo overlap"):
    #  contigs  whole records per rank (longest-first assignment): every rank runs the ordinary kgma_scan on its records -- its own
    #           replay, exactly the reference's extensions -- and ships finished hits; used when the records balance within 10 %
    #  slices   equal slices of the packed genome with window-length halos (kgma_scan_shard + kgma_replay_packed): a genome of a
    #           few huge contigs, and exact match
    parts = None
    if world > 1 and not exact and args.shard != "slices":
        parts = K.partition_records(W.lens, world)
        if parts is None and args.shard == "contigs":
            raise SystemExit("--shard contigs: the records do not balance over %d ranks" % world)
    sub = [None, None, None]  # this rank's sub-genome, its prepared scan, rank 0's merger (contig mode)

    def contig_step(resident, count):
        fl = L.F_ALIGN | (L.F_RESIDENT if resident else 0)
        part = sub[1].scan(fl)
        st1 = ctx.stats()
        if not xch.cap:
            need = part.copy_hits(None, 0)
            part.free()
            return None, st1, need
        ptr = xch.x.begin_step()
        need = part.copy_hits(ptr, xch.cap)
        part.free()
        if need > xch.cap:
            raise RuntimeError("exchange block too small (%d > %d)" % (need, xch.cap))
        xch.x.publish()
        if rank != 0:
            return None, st1, need
        base = xch.x.wait_all()
        t0 = time.perf_counter()
        merged = sub[2].merge(base, xch.cap)
        xch.x.consumed()
        st1 = dict(st1); st1["rank0_replay_ms"] = (time.perf_counter() - t0) * 1e3
        return merged, st1, need

    def shard_step(g_, resident, count):
        """one step of the sharded single-genome run; returns (result on rank 0 | None, stats of this rank's scan)"""
        if parts is not None:
            return contig_step(resident, count)
        fl = (0 if exact else L.F_ALIGN) | (L.F_RESIDENT if resident else 0)
        if exact:
            starts = exact_call(g_, resident, (rank, world))
            st1 = ctx.stats()
            need = 16 + starts.size * 8
            if need > xch.cap:
                if count:
                    raise RuntimeError("exchange block too small (%d > %d)" % (need, xch.cap))
                return None, st1, need
            xch.x.begin_step()
            hv = xch.x.block()
            hv[:8].view(np.int64)[0] = starts.size
            hv[16:16 + starts.size * 8].view(np.int64)[:] = starts
        else:
            part = prep[0].scan_shard(fl, (rank, world))
            st1 = ctx.stats()
            if not xch.cap:
                need = part.pack(None, 0)
                part.free()
                return None, st1, need
            ptr = xch.x.begin_step()
            need = part.pack(ptr, xch.cap)
            part.free()
            if need > xch.cap:
                raise RuntimeError("exchange block too small (%d > %d)" % (need, xch.cap))
        xch.x.publish()
        if rank != 0:
            return None, st1, need
        base = xch.x.wait_all()
        if exact:
            ha = xch.x.blocks()
            allst = np.concatenate([ha[i, 16:16 + int(ha[i, :8].view(np.int64)[0]) * 8].view(np.int64) for i in range(world)])
            xch.x.consumed()
            res = K.exact_match_merge(g_, allst, len(W.query), True)
            return res, st1, need
        out = prep[0].replay_packed(fl, base, world, xch.cap)
        xch.x.consumed()
        st1 = dict(st1); st1["rank0_replay_ms"] = ctx.stats()["host_replay_ms"]
        return out, st1, need

    def whole_step(g_, resident):
        if exact:
            return exact_call(g_, resident)
        return prep[0].scan(L.F_ALIGN | (L.F_RESIDENT if resident else 0))

    def measure(g_, sharded: bool, resident: bool, regions: int):
        """W warm-up steps, then `regions` timed regions of exactly K steps each (barrier + device sync on both sides, max over
        ranks); the reported time is the median region.  Returns (seconds per K steps, last result, clocks, per-step stats)."""
        agg = {k_: 0.0 for k_ in KEYS}
        n_acc = [0]

        def acc(st):
            for k_ in KEYS:
                agg[k_] += st.get(k_, 0) if isinstance(st, dict) else st[k_]
            n_acc[0] += 1

        def step(count):
            if not sharded:
                out_ = whole_step(g_, resident)
                if count:
                    acc(ctx.stats())
                return out_
            out_, st1, _ = shard_step(g_, resident, count)
            if count:
                acc(st1)
            return out_

        if resident and not sharded:
            g_.make_resident(ctx)
        if resident and sharded and parts is not None:
            sub[0].make_resident(ctx)
        if sharded:                                    # size the exchange blocks from one untimed step (2x the largest block)
            _, _, need = shard_step(g_, resident, False)
            xch.size(need)
        t_w = time.perf_counter()
        for _ in range(args.warmup):
            step(False)
        # a step is ~1 ms: W steps are over before host and device clocks have settled (the first of five regions used to read
        # 10-35 % slower than the last).  Keep stepping, untimed, for about 80 ms more -- the same number of steps on every rank
        # (the ranks of a sharded run move in lock-step through the exchange).
        per = (time.perf_counter() - t_w) / max(1, args.warmup)
        n_extra = int(min(400, max(0, 0.08 / max(per, 1e-5))))
        if world > 1:
            t = torch.tensor([n_extra], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_extra = int(t.item())
        for _ in range(n_extra):
            step(False)
        stop, samples = threading.Event(), []
        nvml_handle(local)                               # (initialised before the timed region)
        th = threading.Thread(target=clocks_sampler, args=(local, stop, samples), daemon=True)
        th.start()
        times, out_, own = [], None, []
        for _ in range(regions):
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out_ = step(True)
            torch.cuda.synchronize()
            own.append((time.perf_counter() - t0) / args.steps * 1e3)
            barrier()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        stop.set(); th.join()
        if out_ is not None and hasattr(out_, "full"):
            out_ = out_.full()
        per = {k_: v / max(1, n_acc[0]) for k_, v in agg.items()}
        o = torch.tensor([float(np.median(own))], dtype=torch.float64, device="cuda")
        if world > 1:
            tl = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(tl, o)
            per["per_rank_ms_per_step"] = [float(x.item()) for x in tl]
        else:
            per["per_rank_ms_per_step"] = [float(o.item())]
        per["region_ms_per_step"] = [t_ / args.steps * 1e3 for t_ in times]
        return float(np.median(times)), out_, summarise_clocks(samples), per

    def allsum(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    def key_of(out_):
        if exact:
            return out_
        h = out_ if isinstance(out_, np.ndarray) else out_.hits
        return h[["record", "profile", "first", "last", "D", "genome_pos", "align_score"]]

    # ---- the genome (the same one on every rank: in the sharded run each rank only ever touches its own slice)
    t_setup = time.perf_counter()
    g = make_genome(W.seed)
    t_setup = time.perf_counter() - t_setup
    total = g.total_len
    if not exact:
        prep[0] = K.PreparedScan(g, rvs, wss, cs, thr, W.k, mode, W.buff, W.gap_open, GAP_EXT, ctx=ctx)
    if parts is not None:
        sub[0] = g.subset(parts[rank], ctx=ctx)
        sub[1] = K.PreparedScan(sub[0], rvs, wss, cs, thr, W.k, mode, W.buff, W.gap_open, GAP_EXT, ctx=ctx)
        sub[2] = K.PartitionMerger(parts, W.lens, 0 if W.config == "cluster" else int(wss[0]))
    # cold costs the warm-up hides: context creation, and the first call on a fresh context (prefilter weight-table build,
    # cudaMalloc of scratch / device planes, page-locked staging blocks)
    t0 = time.perf_counter()
    ref_out = whole_step(g, False)
    cold_first_ms = (time.perf_counter() - t0) * 1e3
    if hasattr(ref_out, "full"):
        ref_out = ref_out.full()
    REG = max(1, args.regions)

    if world == 1:
        dt_res, out_res, clocks, a_res = measure(g, False, True, REG)
        dt_e2e, out_e2e, clocks_e2e, a_e2e = measure(g, False, False, REG)
        same = (out_res == out_e2e == ref_out) if exact else (np.array_equal(key_of(out_res), key_of(out_e2e)) and np.array_equal(key_of(out_res), key_of(ref_out)))
        scaling, equal_key, equal_val = "weak", "hits_equal_resident_vs_e2e", bool(same)
        h2d_all, d2h_all = a_e2e["h2d_bytes"], a_e2e["d2h_bytes"]
        xbytes = 0
    else:
        dt_res, out_res, clocks, a_res = measure(g, True, True, REG)
        dt_e2e, out_e2e, clocks_e2e, a_e2e = measure(g, True, False, REG)
        xbytes = xch.cap
        h2d_all, d2h_all = allsum([a_e2e["h2d_bytes"], a_e2e["d2h_bytes"]])
        if rank == 0:
            if exact:
                ref_d = K.exactMatch(W.query, g, ctx=ctx)
                same = out_res == ref_d and out_e2e == ref_d
            else:
                same = np.array_equal(key_of(out_res), key_of(ref_out)) and np.array_equal(key_of(out_e2e), key_of(ref_out))
        else:
            same = True
        scaling, equal_key, equal_val = "strong", "hits_equal_unsharded", bool(same)
    launches_all, = allsum([a_res["launches"] * args.steps])

    # ---- all-GPU concurrent pinned H2D ceiling of this box: what bounds e2e (every rank copies its slice of the packed genome)
    h2d_ceiling = None
    try:
        nbytes = int(total // 4 // world) // 4096 * 4096
        src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        barrier()
        dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        h2d_ceiling = {"aggregate_gbs": world * nbytes * 5 / float(dtc.item()) / 1e9, "bytes_per_rank": nbytes,
                       "floor_ms_per_step": float(dtc.item()) / 5 * 1e3,
                       "note": "all %d ranks copying their slice of the packed genome from pinned host memory at the same time (cudaMemcpyAsync, 5 "
                               "repeats, max over ranks): the floor of e2e ms_per_step on this box" % world}
        del src, dst
    except Exception as e:
        h2d_ceiling = {"error": str(e)}

    # ---- side key for N > 1: N independent replicas (round 1's headline), one genome per GPU, nothing shared
    replicas = None
    if world > 1 and not args.no_extra:
        rdt, rout, _, ra = measure(g, False, True, 1)
        rdt2, _, _, _ = measure(g, False, False, 1)
        replicas = {"scaling": "weak", "workload": "%d independent copies of the genome, one per GPU, no communication" % world,
                    "value": world * total * args.steps / rdt / 1e6, "ms_per_step": rdt / args.steps * 1e3,
                    "e2e": {"value": world * total * args.steps / rdt2 / 1e6, "ms_per_step": rdt2 / args.steps * 1e3}, "unit": UNIT}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        filt_ms = a_res["filter_ms"]
        if exact:
            alg_bytes, npass = total / world * 0.25, 1
            kname = "kgma_exact_match_sampled"
            limiter = "HBM: one sampled 4-byte word per 32-byte sector of the 2-bit plane"
        else:
            alg_bytes, npass = a_res["blocks_total"] * 64 * 0.25, max(1, int(round(a_res["filter_passes"])))
            nine = os.environ.get("KGMA_PREFILTER") != "8mer"
            kname = ("kgma_prefilter9<%d>" if nine else "kgma_prefilter<%d>") % W.k
            limiter = ("shared-memory wavefronts, not HBM: 4.1 wavefronts per random 32-lane table gather (%d gathers per 64 bases), "
                       "DESIGN.md 5.1; %d pass(es) over the genome per step (one per group of profiles sharing a weight table)"
                       % ((64 + (10 if nine else 9) - W.k - 1) // ((10 if nine else 9) - W.k), npass))
        # filter_ms spans every prefilter pass of the step: each pass streams the slice once
        achieved = npass * alg_bytes / (filt_ms * 1e-3) / 1e9 if filt_ms > 0 else 0.0
        nh = (sum(len(v) for v in out_res.values()) if isinstance(out_res, dict) else int(out_res)) if exact else int(len(out_res if isinstance(out_res, np.ndarray) else out_res.hits))
        line = {
            "metric": METRICS[W.config], "value": total * args.steps / dt_res / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_res / args.steps * 1e3,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": W.describe(world),
                       "parallelism": ("one GPU" if world == 1 else
                                       ("the one genome partitioned by contig over %d ranks (longest-first, fullest rank %.1f %% above the even share); every rank "
                                        "runs the ordinary kgma_scan on its records and writes its finished hits (%d-byte blocks) into a shared-memory segment "
                                        "of the host (no collective, no device round trip); rank 0 renumbers records / GenomePos and concatenates"
                                        % (world, 100.0 * (max(sum(W.lens[r] for r in p) for p in parts) * world / float(sum(W.lens)) - 1.0), xbytes)) if parts is not None else
                                       ("the one genome cut into %d equal slices of the packed coordinate space with window-length halos; every rank scans "
                                        "and extends its slice (kgma_scan_shard); the ranks pack %d-byte blocks of (run, extension result) pairs into a "
                                        "shared-memory segment of the host (no collective, no device round trip); host-only merge + replay on rank 0 "
                                        "(kgma_replay_packed)" % (world, xbytes))),
                       "shard": None if world == 1 else ("contigs" if parts is not None else "slices"),
                       "l2": "input (%.0f MB packed per GPU) larger than L2; no flush needed" % (total / 4e6 / world) if total / 4 / world > 126e6 else
                             "input %.0f MB packed per GPU: smaller than the 126 MB L2 at this N; the resident timed loop re-reads it from L2/HBM as a serving loop would" % (total / 4e6 / world),
                       "timing": "host clock around blocking C-ABI calls bracketed by device sync + barrier, max over ranks, median of %d regions of "
                                 "exactly K steps (>= the CUDA-event device time reported in device_ms_per_step); after the W warm-up steps the "
                                 "loop keeps stepping untimed for about 80 ms more (clocks settle: a step is ~1 ms)" % REG},
            "e2e": {"value": total * args.steps / dt_e2e / 1e6, "unit": UNIT, "ms_per_step": dt_e2e / args.steps * 1e3,
                    "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all, "h2d_ceiling": h2d_ceiling},
            "gpu_launches": int(launches_all),
            "device_ms_per_step": {"prefilter_or_match": filt_ms, "count_table": a_res["exact_ms"], "extension": a_res["align_ms"],
                                   "scan_total": a_res["total_ms"], "e2e_h2d": a_e2e["h2d_ms"], "note": "rank 0's own slice when N > 1"},
            "host_ms_per_step": {"call_wall": a_res["wall_ms"], "setup": a_res["host_setup_ms"], "results": a_res["host_cand_ms"],
                                 "replay_or_merge": a_res["host_replay_ms"], "rank0_replay_of_all_blocks": a_res["rank0_replay_ms"], "e2e_call_wall": a_e2e["wall_ms"]},
            "per_rank_ms_per_step": {"resident": a_res["per_rank_ms_per_step"], "e2e": a_e2e["per_rank_ms_per_step"],
                                     "regions_resident": a_res["region_ms_per_step"], "regions_e2e": a_e2e["region_ms_per_step"],
                                     "note": "each rank's own timed loop before the closing barrier; ms_per_step is the max incl. the barrier"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (784.4e6 if (args.scale == 1.0 and W.config == "single" and world == 1) else None),
                         "traffic_source": "ncu --set full, profiles/r2_kernels_ncu_full_summary.csv (dram read 780.2 MB + write 4.2 MB per launch of the N=1 default config)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": alg_bytes, "limiter": limiter},
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "hits_per_step": nh, "runs_per_step": a_res["n_runs"], equal_key: equal_val,
            "extensions_per_step": a_res["n_align"], "extensions_redone_per_step": a_res["n_align_redo"], "extensions_by_summary_kernel_per_step": a_res["n_align_summary"], "extensions_third_sweep_per_step": a_res["n_align_head"],
            "prefilter_blocks_flagged_per_step": a_res["blocks_flagged"], "count_table_windows_per_step": a_res["exact_windows"],
            "cold": {"context_create_ms": t_ctx * 1e3, "first_call_ms": cold_first_ms, "steady_e2e_ms": dt_e2e / args.steps * 1e3,
                     "note": "first_call = the first operator call on a fresh context, from pinned host planes: device plane + scratch cudaMalloc, "
                             "page-locked staging blocks, prefilter weight-table build; later calls reuse all of it"},
            "setup_s": t_setup, "readme_julia_mbs": 40,
        }
        if replicas:
            line["replicas"] = replicas
        if world == 1 and not args.no_extra and W.config == "single":
            line["other_configs"] = other_configs(K, ctx, g, W, total, rvs[0], wss[0], cs[0], peak)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(W, args.cpu_seconds)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        if xch is not None and xch.x is not None:
            xch.x.close()                                     # (rank 0 unlinks the segment)
        dist.destroy_process_group()


def other_configs(K, ctx, g, W, total, RV, ws, cons, peak):
    """BASELINE configs[2] (cluster mode, 6 profiles) and configs[3] (exactMatch, 300-nt query) on the same resident genome,
    the dense count-table pass, the caller-buffer ingest tiers and the FASTA-text tier; a few iterations each, reported for
    context (their own bench lines: --config cluster|exact|k7; parity for them is in tests/)."""
    import ctypes as C
    L = K.L
    lens = W.lens

    def timeit(f, n):
        f()
        t0 = time.perf_counter()
        for _ in range(n):
            out = f()
        return (time.perf_counter() - t0) / n * 1e3, out

    res = {}
    rvs, wss, cs, inv = K.cluster_ref_API(TF, KMER)
    rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
    thr = CLUSTER_THR
    g.make_resident(ctx)
    ms, out = timeit(lambda: K.scan_raw(g, rvs, wss, cs, thr, KMER, L.MODE_CLUSTER, 100, L.F_ALIGN | L.F_RESIDENT, -200, -1, ctx=ctx), 5)
    st = ctx.stats()
    res["findGenes_cluster_mode"] = {"profiles": len(wss), "ms_per_step": ms, "value": total / ms / 1e3, "unit": UNIT, "hits": int(len(out.hits)),
                                     "device_ms": {"prefilter": st["filter_ms"], "count_table": st["exact_ms"], "extension": st["align_ms"]},
                                     "extensions": int(st["n_align"])}
    ms, out = timeit(lambda: K.scan_raw(g, rvs, wss, cs, thr, KMER, L.MODE_CLUSTER, 100, L.F_ALIGN, -200, -1, ctx=ctx), 3)
    res["findGenes_cluster_mode"]["e2e"] = {"ms_per_step": ms, "value": total / ms / 1e3, "unit": UNIT, "note": "from pinned host memory, H2D inside"}
    g.make_resident(ctx)
    ms, out = timeit(lambda: K.scan_raw(g, [RV], [ws], [cons], [THR], KMER, L.MODE_SINGLE, BUFF, L.F_ALIGN | L.F_RESIDENT | L.F_DENSE, GAP_OPEN, GAP_EXT, ctx=ctx), 2)
    st = ctx.stats()
    res["findGenes_dense_count_table"] = {"ms_per_step": ms, "value": total / ms / 1e3, "unit": UNIT, "hits": int(len(out.hits)),
                                          "device_ms": {"count_table": st["exact_ms"]},
                                          "note": "KGMA_F_DENSE: every window through the shared-memory count-table kernel, no prefilter "
                                                  "(what do_return_dists and unfilterable profiles run)"}
    # the reference's experimental strobemer search (Strobemer_findGenes): same genome, randstrobe profile of the same family,
    # always the dense count-table pass (256 codes, nothing to filter on)
    try:
        SRV, sws, scons = K.strobe_gen_ref_ws_cons(TF)
        ms, out = timeit(lambda: K.strobe_scan_raw(g, SRV, sws, scons, 20.0, 2, 3, 5, 5, BUFF, L.F_ALIGN | L.F_RESIDENT, GAP_OPEN, -5, ctx=ctx), 2)
        st = ctx.stats()
        res["Strobemer_findGenes"] = {"ms_per_step": ms, "value": total / ms / 1e3, "unit": UNIT, "hits": int(len(out.hits)),
                                      "device_ms": {"count_table": st["exact_ms"], "extension": st["align_ms"]}, "runs": int(st["n_runs"]),
                                      "note": "KGMA_MODE_STROBE, s=2 w_min=3 w_max=5 q=5, KmerDistThr 20, gap (-69,-5)"}
    except Exception as e:
        res["Strobemer_findGenes"] = {"error": str(e)}
    rng = np.random.default_rng(5)
    query = "".join(np.asarray(list("ACGT"))[rng.integers(0, 4, size=300)])
    for i in range(1000):
        r = int(rng.integers(0, len(lens)))
        g.put_seq(r, int(rng.integers(20000, lens[r] - 20000)), query)

    def em(flags):
        mp = C.POINTER(L.Match)(); n = C.c_int64()
        ctx.check(ctx._lib.kgma_exact_match(ctx._h, g._h, query.encode(), len(query), 1, flags, C.byref(mp), C.byref(n)))
        if n.value:
            ctx._lib.kgma_free(mp)
        return n.value

    ms_h, n_h = timeit(lambda: em(0), 3)
    g.make_resident(ctx)
    ms, n = timeit(lambda: em(L.F_RESIDENT), 10)
    st = ctx.stats()
    ach = total * 0.25 / (st["filter_ms"] * 1e-3) / 1e9
    res["exactMatch_300nt"] = {"ms_per_step": ms, "value": total / ms / 1e3, "unit": UNIT, "matches": int(n), "planted": 1000,
                               "e2e": {"ms_per_step": ms_h, "value": total / ms_h / 1e3, "unit": UNIT, "matches": int(n_h), "note": "from pinned host memory, H2D inside"},
                               "roofline": {"bound": "hbm", "kernel": "kgma_exact_match_sampled", "achieved": ach, "peak": peak, "unit": "GB/s",
                                            "frac": ach / peak, "algorithmic_bytes": "0.25 B/base: the 2-bit plane, one sampled word per 32 B sector; "
                                            "N is checked against the masked-run list, the ambiguity plane is not read"}}
    # ---- caller-buffer tiers: what a foreign caller (the Julia shim) pays depending on where its packed planes live
    try:
        res["caller_buffer_tiers"] = caller_buffer_tiers(K, ctx, g, W, RV, ws, cons)
    except Exception as e:
        res["caller_buffer_tiers"] = {"error": str(e)}
    # tier T2: the reference's own entry point -- findGenes(genome_path = FASTA text on disk).  The whole cfg2 genome is written
    # out as 80-column FASTA (3.1 GB), then: parallel mmap parse + 2-bit pack on the host cores, and the FIRST scan of the
    # fresh, pageable genome (staged upload through a small page-locked ring).
    path = None
    try:
        import tempfile
        width = 80
        with tempfile.NamedTemporaryFile(suffix=".fasta", delete=False, dir=os.environ.get("TMPDIR", "/tmp")) as fh:
            path = fh.name
            for r in range(len(lens)):
                seq = g.seq_array(r)
                body = seq[:seq.size // width * width].reshape(-1, width)
                fh.write(b">contig%d T2 tier\n" % (r + 1))
                fh.write(np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).tobytes())
                fh.write(seq[body.size:].tobytes() + b"\n")
        t0 = time.perf_counter()
        g2 = K.Genome.from_fasta(path)
        t1 = time.perf_counter()
        out2 = K.scan_raw(g2, [RV], [ws], [cons], [THR], KMER, L.MODE_SINGLE, BUFF, L.F_ALIGN, GAP_OPEN, GAP_EXT, ctx=ctx)
        t2 = time.perf_counter()
        K.scan_raw(g2, [RV], [ws], [cons], [THR], KMER, L.MODE_SINGLE, BUFF, L.F_ALIGN, GAP_OPEN, GAP_EXT, ctx=ctx)
        t3 = time.perf_counter()
        res["t2_from_fasta_text"] = {"bases": int(g2.total_len), "file_bytes": os.path.getsize(path), "parse_pack_ms": (t1 - t0) * 1e3,
                                     "first_scan_ms": (t2 - t1) * 1e3, "second_scan_ms": (t3 - t2) * 1e3,
                                     "value": g2.total_len / (t2 - t0) / 1e6, "unit": UNIT, "hits": int(len(out2.hits)),
                                     "host_threads": len(os.sched_getaffinity(0)),
                                     "note": "value = bases / (parse + first scan).  The genome stays in pageable memory: every scan uploads it "
                                             "through the page-locked staging ring (the first one also sets the ring up)"}
        del g2
    except Exception as e:                                        # a full /tmp must not cost the headline numbers
        res["t2_from_fasta_text"] = {"error": str(e)}
    finally:
        if path and os.path.exists(path):
            os.unlink(path)
    return res


def caller_buffer_tiers(K, ctx, g, W, RV, ws, cons):
    """e2e ms per scan for the three ways a caller's packed genome can reach the library:
    (a) library-owned page-locked planes filled in place (kgma_genome_create_pinned + kgma_genome_record_planes): zero copy;
    (b) kgma_genome_append_packed from the caller's own buffer into pageable library planes (staged upload per scan);
    the headline e2e is (a)'s layout, produced by kgma_genome_synth."""
    import ctypes as C
    L = K.L
    lib = ctx._lib
    lens = np.asarray(W.lens, dtype=np.int64)
    out = {}

    def scan(gx):
        return K.scan_raw(gx, [RV], [ws], [cons], [THR], KMER, L.MODE_SINGLE, BUFF, L.F_ALIGN, GAP_OPEN, GAP_EXT, ctx=ctx)

    def timed(gx, n=3):
        scan(gx)
        t0 = time.perf_counter()
        for _ in range(n):
            o = scan(gx)
        return (time.perf_counter() - t0) / n * 1e3, o

    # the caller's packed records: read back from the bench genome, record by record
    packed = []
    for r in range(len(lens)):
        sp, mp = C.c_void_p(), C.c_void_p()
        ctx.check(lib.kgma_genome_record_planes(g._h, r, C.byref(sp), C.byref(mp)))
        nw, nm = (int(lens[r]) + 15) // 16, (int(lens[r]) + 31) // 32
        packed.append((np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_uint32)), shape=(nw,)).copy(),
                       np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_uint32)), shape=(nm,)).copy()))
    # (a) in-place fill of library-owned pinned planes
    t0 = time.perf_counter()
    h = C.c_void_p()
    ctx.check(lib.kgma_genome_create_pinned(ctx._h, len(lens), lens.ctypes.data, C.byref(h)))
    ga = K.Genome(h, lib)
    for r, (s2, m2) in enumerate(packed):
        sp, mp = C.c_void_p(), C.c_void_p()
        ctx.check(lib.kgma_genome_record_planes(ga._h, r, C.byref(sp), C.byref(mp)))
        C.memmove(sp, s2.ctypes.data, s2.nbytes)
        C.memmove(mp, m2.ctypes.data, m2.nbytes)
        lib.kgma_genome_set_names(ga._h, r, g.identifier(r).encode(), g.description(r).encode())
    ctx.check(lib.kgma_genome_seal(ga._h))
    t_fill = (time.perf_counter() - t0) * 1e3
    ms, oa = timed(ga)
    out["pinned_in_place"] = {"ingest_ms": t_fill, "e2e_ms_per_scan": ms, "hits": int(len(oa.hits)),
                              "note": "kgma_genome_create_pinned + kgma_genome_record_planes: the caller packs straight into page-locked planes"}
    del ga
    # (b) append_packed from the caller's buffers (pageable library planes, staged upload)
    t0 = time.perf_counter()
    h = C.c_void_p()
    lib.kgma_genome_create(C.byref(h))
    gb = K.Genome(h, lib)
    for r, (s2, m2) in enumerate(packed):
        ctx.check(lib.kgma_genome_append_packed(gb._h, g.identifier(r).encode(), g.description(r).encode(), s2.ctypes.data, m2.ctypes.data, int(lens[r])))
    ctx.check(lib.kgma_genome_seal(gb._h))
    t_app = (time.perf_counter() - t0) * 1e3
    ms, ob = timed(gb)
    out["append_packed_pageable"] = {"ingest_ms": t_app, "e2e_ms_per_scan": ms, "hits": int(len(ob.hits)),
                                     "note": "kgma_genome_append_packed copies into pageable planes; every scan goes through the page-locked staging ring"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="single", choices=["single", "cluster", "exact", "k7"])
    ap.add_argument("--regions", type=int, default=5, help="timed regions of exactly --steps steps each; the median is reported")
    ap.add_argument("--scale", type=float, default=1.0, help="genome size as a fraction of the config's size (testing only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_configs / replicas blocks")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--shard", default="auto", choices=["auto", "contigs", "slices"],
                    help="N > 1: whole records per rank (when they balance within 10 %%) or equal slices of the packed genome")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
