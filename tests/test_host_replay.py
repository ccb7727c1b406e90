"""Host side of the scan on CPU: kgma_replay without a device (merge of run summaries + the sequential hit state machine)
against the oracle, on run lists that arrive the way the device produces them at its worst -- in arbitrary order, maximal
runs cut into adjacent pieces (span / chunk / shard edges), pieces reported twice (a span evaluated again).  The
per-window distances come from the oracle in exact-arithmetic mode; on the GPU box the CUDA kernels produce the run
summaries (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from conftest import TF
from test_multirank_gloo import RUN_DT, runs_from_D


_HITS_SEEN = []


def _mutate(rng, s, sub):
    v = np.frombuffer(s.encode(), np.uint8).copy()
    m = rng.random(v.size) < sub
    v[m] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=int(m.sum()))]
    return v.tobytes().decode()


@pytest.mark.parametrize("seed", range(8))
def test_host_replay_of_scrambled_run_lists(tmp_path, seed):
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    rng = np.random.default_rng(500 + seed)
    refs = O.Fasta(TF)
    k = 6
    RV, ws, cons = K.gen_ref_ws_cons(TF, k)
    N = RV.n_refs
    den = 2 * k * N * N
    thr = float(rng.choice([24.0, 30.0, 33.5, 36.5]))
    T = int(np.ceil(thr * den))
    # a few records: random sequence with planted, mutated family members (some back to back, some at the very ends)
    recs = []
    for r in range(int(rng.integers(2, 6))):
        parts = []
        for j in range(int(rng.integers(1, 9))):
            parts.append("".join(np.asarray(list("ACGT"))[rng.integers(0, 4, size=int(rng.integers(0, 3000)))]))
            parts.append(_mutate(rng, refs.seq(int(rng.integers(0, len(refs)))), 0.0 if j == 0 else float(rng.choice([0.0, 0.03, 0.1, 0.2]))))
        recs.append(("rec%d" % r, "".join(parts)))
    recs.insert(1, ("short", "ACGT" * 20))                     # shorter than the window: skipped, GenomePos not advanced
    path = tmp_path / "g.fasta"
    with open(path, "w") as fh:
        for d_, s in recs:
            fh.write(">" + d_ + "\n" + s + "\n")
    g = K.Genome.from_fasta(str(path))
    with O.exact_arithmetic(N):
        oh, _, d = O.ac_gma_testing(str(path), np.asarray(RV), cons, windowsize=ws, thr=thr, do_align=False, do_return_dists=True)
    S = np.asarray(RV.S, dtype=np.int64)
    runs, firsts, base = [], np.full(len(g), np.iinfo(np.int64).min, dtype=np.int64), 0
    for r in range(len(g)):
        L = g.seqsize(r)
        if L < ws:
            continue
        steps = L - ws
        D = np.rint(d[base:base + steps] * den).astype(np.int64)
        base += steps
        c = O.kmer_count(recs[r][1][:ws], k)
        firsts[r] = int(np.sum((N * c.astype(np.int64) - S) ** 2))
        for run in runs_from_D(D, T, r, 1, steps + 1):
            _, _, a, b, _, _, _, _ = run
            # cut the maximal run into 1..4 adjacent pieces, each with its own minimum / first argmin
            cuts = sorted(set([a, b + 1] + [int(x) for x in rng.integers(a, b + 2, size=int(rng.integers(0, 4)))]))
            pieces = []
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                if hi > lo:
                    seg = D[lo - 1:hi - 1]
                    pieces.append((r, 0, lo, hi - 1, lo + int(np.argmin(seg)), int(seg.min()), 0, 0))
            runs += pieces
            if pieces and rng.random() < 0.3:
                runs.append(pieces[int(rng.integers(0, len(pieces)))])        # the same piece reported twice
    order = rng.permutation(len(runs))
    arr = np.array([runs[i] for i in order], dtype=RUN_DT) if runs else np.zeros(0, dtype=RUN_DT)
    out = K.replay_raw(g, [RV], [ws], [cons], [thr], k, K.L.MODE_SINGLE, 50, 0, -69, -1, arr.view(np.uint8), firsts, host_only=True)
    got = [(int(h.record), int(h.first), int(h.last), int(h.genome_pos), int(h.D)) for h in out.hits]
    want = [(h.record, h.first, h.last, h.genome_pos, int(round(h.dist * den))) for h in oh]
    assert got == want
    assert len(runs) >= 1
    _HITS_SEEN.append(len(want))


def test_host_replay_cases_were_not_vacuous():
    """(runs after the parametrised cases above) most of them must have produced hits to compare"""
    if len(_HITS_SEEN) < 8:
        pytest.skip("the parametrised cases did not all run in this process")
    assert sum(1 for n in _HITS_SEEN if n >= 3) >= 5, _HITS_SEEN


@pytest.mark.parametrize("seed", range(6))
def test_host_replay_cluster_mode_of_scrambled_run_lists(tmp_path, seed):
    """the same for Omn_KmerGMA!'s state machine (OmnGenomeMiner.jl:95-157 without extension): one run list per profile, hits of
    all profiles interleaved in step order, prev_hit_range shared between the profiles of a record, GenomePos advanced by
    every record"""
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    rng = np.random.default_rng(900 + seed)
    refs = O.Fasta(TF)
    k = 6
    rvs, wss, cs, inv = K.cluster_ref_API(TF, k)
    rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
    Cn = len(wss)
    Ns = [int(v.n_refs) for v in rvs]
    dens = [2 * k * n * n for n in Ns]
    thr = [35, 31, 38, 34, 27, 27][:Cn] if seed % 2 == 0 else [float(rng.choice([26.0, 30.0, 34.5]))] * Cn
    Ts = [int(np.ceil(t * d)) for t, d in zip(thr, dens)]
    maxws = max(wss)
    recs = []
    for r in range(int(rng.integers(2, 5))):
        parts = []
        for j in range(int(rng.integers(1, 8))):
            parts.append("".join(np.asarray(list("ACGT"))[rng.integers(0, 4, size=int(rng.integers(0, 2500)))]))
            parts.append(_mutate(rng, refs.seq(int(rng.integers(0, len(refs)))), 0.0 if j == 0 else float(rng.choice([0.0, 0.03, 0.1, 0.2]))))
        recs.append(("rec%d" % r, "".join(parts) + "ACGT" * 80))
    recs.insert(1, ("short", "ACGT" * 20))
    path = tmp_path / "g.fasta"
    with open(path, "w") as fh:
        for d_, s in recs:
            fh.write(">" + d_ + "\n" + s + "\n")
    g = K.Genome.from_fasta(str(path))
    nr = len(g)
    with O.exact_arithmetic(Ns):
        oh, _, dv = O.Omn_KmerGMA(str(path), [np.asarray(v) for v in rvs], wss, cs, k=k, thr_vec=thr, buff=100, align_hits=False, do_return_dists=True)
    runs, firsts = [], np.full(Cn * nr, np.iinfo(np.int64).min, dtype=np.int64)
    for q in range(Cn):
        S = np.asarray(rvs[q].S, dtype=np.int64)
        base = 0
        for r in range(nr):
            L = g.seqsize(r)
            steps = L - maxws - k + 2                                          # view(seq, k:L-maxws+1)
            if L >= wss[q]:
                c = O.kmer_count(recs[r][1][:wss[q]], k)
                firsts[q * nr + r] = int(np.sum((Ns[q] * c.astype(np.int64) - S) ** 2))
            if steps <= 0:
                continue
            D = np.rint(dv[q][base:base + steps] * dens[q]).astype(np.int64)
            base += steps
            for run in runs_from_D(D, Ts[q], r, 1, steps + 1):
                _, _, a, b, _, _, _, _ = run
                cuts = sorted(set([a, b + 1] + [int(x) for x in rng.integers(a, b + 2, size=int(rng.integers(0, 4)))]))
                pieces = []
                for lo, hi in zip(cuts[:-1], cuts[1:]):
                    if hi > lo:
                        seg = D[lo - 1:hi - 1]
                        pieces.append((r, q, lo, hi - 1, lo + int(np.argmin(seg)), int(seg.min()), 0, 0))
                runs += pieces
                if pieces and rng.random() < 0.3:
                    runs.append(pieces[int(rng.integers(0, len(pieces)))])
        assert base == dv[q].size
    order = rng.permutation(len(runs))
    arr = np.array([runs[i] for i in order], dtype=RUN_DT) if runs else np.zeros(0, dtype=RUN_DT)
    out = K.replay_raw(g, rvs, wss, cs, thr, k, K.L.MODE_CLUSTER, 100, 0, -200, -1, arr.view(np.uint8), firsts, host_only=True)
    got = [(int(h.record), int(h.profile), int(h.first), int(h.last), int(h.genome_pos), int(h.D)) for h in out.hits]
    want = [(h.record, h.kfv, h.first, h.last, h.genome_pos, int(round(h.dist * dens[h.kfv - 1]))) for h in oh]
    assert got == want
    assert len(want) >= 2
