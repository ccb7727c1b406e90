"""CPU-side checks of the drop-in boundary: libkmergma_cuda.so loads without a GPU and exports every
function include/kmergma.h declares; host-only entry points (ingest, profile generation) agree with the oracle.
No compute entry point is called here."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT, TF, GENOME, MINI_GENOME, TEST_CONSENSUS


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "kmergma.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kgma_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import kmergma_jl_b200 as K
    lib = K.L.load()
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
        assert n in K.L.SYMBOLS, "ctypes prototype missing for " + n
    assert set(K.L.SYMBOLS) == set(names)
    assert lib.kgma_version() >= 100


def test_struct_layouts_match_header():
    import kmergma_jl_b200 as K
    L = K.L
    assert C.sizeof(L.Run) == 48 and C.sizeof(L.Hit) == 80 and C.sizeof(L.Match) == 24
    assert C.sizeof(L.ScanParams) == 64 and C.sizeof(L.Profile) == 48 and C.sizeof(L.AlignEvent) == 40
    # field by field against the header text, and the prototypes' argument counts
    cty = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "double": C.c_double}
    hs = _header_structs()
    for cls, cname in ((L.Profile, "kgma_profile"), (L.ScanParams, "kgma_scan_params"), (L.Run, "kgma_run"), (L.RunExt, "kgma_run_ext"),
                       (L.Hit, "kgma_hit"), (L.Stats, "kgma_stats"), (L.Match, "kgma_match"), (L.AlignEvent, "kgma_align_event")):
        assert [n for n, _ in cls._fields_] == [n for _, n in hs[cname]], cname
        for (n, t), (ht, _) in zip(cls._fields_, hs[cname]):
            assert (C.sizeof(t) == C.sizeof(C.c_void_p) and not issubclass(t, (C.c_int64, C.c_double))) if ht == "ptr" else t is cty[ht], (cname, n)
    arity = _header_arity()
    assert set(arity) == set(L.SYMBOLS)
    for n, (_, argt) in L.SYMBOLS.items():
        assert len(argt) == arity[n], (n, len(argt), arity[n])


def test_no_gpu_means_loud_failure():
    """there is no CPU fallback: without a device kgma_create fails with KGMA_E_CUDA"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kmergma_jl_b200 as K
    with pytest.raises(K.KmerGMAError) as e:
        K.Context(0)
    assert e.value.code == K.L.E_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kmergma.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("# oracle", ""), os.path.join(dp, f)


def test_profile_generation_matches_oracle():
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    for k in (1, 2, 6, 7):
        rv, ws, cons = K.gen_ref_ws_cons(TF, k)
        orv, ows, ocons = O.gen_ref_ws_cons(TF, k)
        assert ws == ows == 289 and cons == ocons == TEST_CONSENSUS
        assert np.array_equal(np.asarray(rv), orv)
    kf, kw, kc, kinv = K.cluster_ref_API(TF, 6)
    of = O.cluster_ref_API(TF, 6)
    assert kw == list(of[1]) == [288, 288, 288, 289, 290, 289]
    assert all(np.array_equal(np.asarray(a), b) for a, b in zip(kf, of[0]))
    assert list(kc) == list(of[2])
    # KFV -> (S, N) recovery used when a plain Float64 refVec crosses the boundary
    lib = K.L.load()
    for v in list(kf) + [K.gen_ref_ws_cons(TF, 6)[0]]:
        S = np.zeros(v.size, np.int32); n = C.c_int32()
        assert lib.kgma_profile_from_kfv(np.ascontiguousarray(v, np.float64).ctypes.data, v.size, 0, S.ctypes.data, C.byref(n)) == 0
        assert np.array_equal(S * v.n_refs, v.S.astype(np.int64) * n.value)     # same rational profile


def test_ingest_round_trip_and_packed_paths():
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    g, f = K.Genome.from_fasta(GENOME), O.Fasta(GENOME)
    assert len(g) == len(f) == 4
    for r in range(4):
        assert g.identifier(r) == f.identifier(r) and g.description(r) == f.description(r)
        assert g.seq(r) == f.seq(r)
    assert g.seq(3, 6852, 7140) == f.seq(3)[6851:7140]
    # ASCII vs pre-packed vs BioSequences 4-bit ingest give the same container
    s = "ACGTNNacgtn" * 37 + "TTGACA"
    lib = K.L.load()
    codes = np.array([{"A": 0, "C": 1, "G": 2, "T": 3, "N": 3}[c] for c in s.upper()], dtype=np.uint64)
    n = len(s)
    w2 = np.zeros((n + 15) // 16, np.uint32); m = np.zeros((n + 31) // 32, np.uint32); b4 = np.zeros((n + 15) // 16, np.uint64)
    for i, c in enumerate(s.upper()):
        w2[i >> 4] |= np.uint32(int(codes[i]) << (2 * (i & 15)))
        if c == "N":
            m[i >> 5] |= np.uint32(1 << (i & 31))
        b4[i >> 4] |= np.uint64((15 if c == "N" else 1 << int(codes[i])) << (4 * (i & 15)))
    outs = []
    for mode in ("ascii", "packed", "bio4"):
        h = C.c_void_p(); lib.kgma_genome_create(C.byref(h))
        if mode == "ascii":
            rc = lib.kgma_genome_append_ascii(h, b"x", b"x y", s.encode(), n)
        elif mode == "packed":
            rc = lib.kgma_genome_append_packed(h, b"x", b"x y", w2.ctypes.data, m.ctypes.data, n)
        else:
            rc = lib.kgma_genome_append_bio4(h, b"x", b"x y", b4.ctypes.data, n)
        assert rc == 0 and lib.kgma_genome_seal(h) == 0
        gg = K.Genome(h, lib)
        outs.append(gg.seq(0))
    assert outs[0] == outs[1] == outs[2] == s.upper()
    with pytest.raises(K.KmerGMAError):
        K.Genome.from_records([("bad", "ACGT!ACGT")])


def test_fasta_ingest_edge_cases(tmp_path):
    """parallel mmap ingest vs the oracle's line-by-line reader: CRLF, blank lines, lower case, no trailing newline,
    N / IUPAC symbols, an empty record, records larger than one 4 MB packing task, task boundaries inside a line"""
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    big = "".join(np.asarray(list("ACGTN"))[rng.integers(0, 5, size=9_500_003)])
    recs = [("r1 first record", "ACGTNNNNacgtn" * 11), ("empty", ""), ("r3\ttabbed desc", big), ("r4", "T"), ("r5 last", "GATTACA" * 3)]
    p = tmp_path / "edge.fasta"
    with open(p, "wb") as fh:
        for i, (d, s) in enumerate(recs):
            eol = b"\r\n" if i % 2 else b"\n"
            fh.write(b">" + d.encode() + eol)
            width = [7, 60, 61, 80, 5][i]
            for a in range(0, len(s), width):
                fh.write(s[a:a + width].encode() + eol)
            if i == 0:
                fh.write(eol)                                  # blank line between records
        fh.seek(-1, 1); fh.truncate()                          # no trailing newline
    g, f = K.Genome.from_fasta(str(p)), O.Fasta(str(p))
    assert len(g) == len(f) == 5
    for r in range(5):
        assert g.identifier(r) == f.identifier(r) == recs[r][0].split()[0]
        assert g.seqsize(r) == f.seqsize(r) == len(recs[r][1])
        assert g.seq(r) == f.seq(r) == recs[r][1].upper()
    q = tmp_path / "iupac.fasta"
    q.write_text(">x\nACGTRYACGT\n")
    gi = K.Genome.from_fasta(str(q))                           # IUPAC other than N: ingested (exact match may use it) ...
    assert gi.seq(0) == "ACGT??ACGT"
    bad = tmp_path / "bad.fasta"
    bad.write_text(">x\nACGT!ACGT\n")
    with pytest.raises(K.KmerGMAError):
        K.Genome.from_fasta(str(bad))


def test_host_mirror_goldens():
    """host-side mirrors of the small reference utilities, against the reference's own goldens
    (test/test_folder/test-KmerGMA.jl:2-26, :28-46 via gen_ref_ws_cons, :96-110, :116-120)"""
    import kmergma_jl_b200 as K
    ts = "ATGCATGC"                                                        # test/runtests.jl:47
    test_KFV = [0, 0, 0, 2, 1, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0]           # test/runtests.jl:50-51
    assert K.kmer_count(ts, 1).tolist() == [2, 2, 2, 2]
    assert K.kmer_count(ts[:8], 2).tolist() == test_KFV
    assert K.kmer_dist(ts * 25 + "A" + ts * 25, ts * 25 + "G" + ts * 25, 2) == 1.0
    assert K.kmer_dist(ts * 25 + "AA" + ts * 25, ts * 25 + "GT" + ts * 25, 2) == 2.0
    assert K.as_UInt(ts) == 14649 and K.as_kmer(14649, 8) == ts
    # Consensus.jl through the native profile builder: column votes, ties to the earlier symbol, shorter refs vote less
    assert K.gen_ref_ws_cons([ts, "ATGCATGG", "ATGCATGG"], 1)[2] == "ATGCATGG"
    assert K.gen_ref_ws_cons(["ACGT", "ACGTA"], 1)[1:] == (4, "ACGTA")     # Int(round(4.5)) == 4 (round half even)
    # eliminate_null_params
    kf, w, c = K.eliminate_null_params([np.array(test_KFV, float), np.array(test_KFV, float) + 0.3], [8, 9], [ts, ts + "Y"], [False, True])
    assert [v.tolist() for v in kf] == [test_KFV] and w == [8] and c == [ts]
    rvs, wss, cs, inv = K.cluster_ref_API(TF, 6, cutoffs=[7, 12, 20, 25])
    rvs, wss, cs = K.eliminate_null_params(rvs, wss, cs, inv)
    assert wss == [288, 288, 288, 289, 290, 289] and len(rvs) == len(cs) == 6
    # estimate_optimal_threshold: same algorithm, numpy's RNG instead of Julia's: the cluster goldens
    # round to the reference's values, the single-profile one lands within 1 of it
    rvs5, ws5, _, _ = K.cluster_ref_API(TF, 6, cutoffs=[7, 12, 20, 25], include_avg=False)
    assert [int(round(x)) for x in K.estimate_optimal_threshold(rvs5, ws5, buffer=8)] == [38, 33, 41, 37, 29]
    RV, _, _ = K.gen_ref_ws_cons(TF, 6)
    assert abs(K.estimate_optimal_threshold(RV, 299, buffer=12) - 27) <= 1.0


def test_strobemer_utility_goldens():
    """test/test_folder/test-StrobemerGMA.jl:1-18 (the only strobemer functions the reference tests)"""
    import kmergma_jl_b200 as K
    ts = "ATGCATGC"
    assert K.randstrobe_score("ATGC", "GTGT", 5) == 4 and K.randstrobe_score("ATGC", "GTGT", 7) == 6
    assert K.get_strobe_2_mer("ATCTCTGTTT") == "AT--CT----"
    assert K.get_strobe_2_mer(ts) == "ATGC----"
    assert K.get_strobe_2_mer("ATCTCTGTTT", withGap=False) == "ATCT"
    assert K.get_strobe_2_mer(ts, withGap=False) == "ATGC"
    counts = K.ungapped_strobe_2_mer_count(ts, s=1, w_min=2, w_max=4)
    assert round(float(np.mean(counts)), 4) == 0.3125
    assert counts[3] == 2 and counts[4] == counts[11] == counts[14] == 1


def test_strobemer_profile_and_oracle_pieces():
    """the strobemer path's building blocks: the oracle's restatement and the library's native profile generator
    (kgma_refs_strobe_profile, StrobeRefGen.jl:4-42) against the host mirror of the utilities the reference's tests pin
    (get_strobe_2_mer / ungapped_strobe_2_mer_count, test-StrobemerGMA.jl:1-18), for several (s, w_min, w_max, q)"""
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    assert np.array_equal(O.ungapped_strobe_2_mer_count("ATGCATGC", 1, 2, 4), K.ungapped_strobe_2_mer_count("ATGCATGC", s=1, w_min=2, w_max=4))
    rng = np.random.default_rng(3)
    refs = O.Fasta(TF)
    for args in ((2, 3, 5, 5), (1, 2, 4, 5), (3, 4, 5, 7), (2, 3, 5, 7), (2, 4, 6, 3)):
        for _ in range(12):
            s_ = "".join(np.asarray(list("ACGTN"))[rng.integers(0, 5, size=int(rng.integers(args[2] + args[0] - 1, 80)))])
            assert np.array_equal(O.ungapped_strobe_2_mer_count(s_, *args), K.ungapped_strobe_2_mer_count(s_, *args)), (s_, args)
        RV, ws, cons = K.strobe_gen_ref_ws_cons(TF, *args)
        orv, ows, ocons = O.strobe_gen_ref_ws_cons(TF, *args)
        assert ws == ows == 289 and cons == ocons and np.array_equal(np.asarray(RV), orv) and RV.n_refs == 84
        tot = sum(K.ungapped_strobe_2_mer_count(refs.seq(r), *args) for r in range(len(refs)))
        assert np.array_equal(np.asarray(RV.S, dtype=np.float64), tot)
    # the scan on the one-record fixture: the same three loci findGenes reports there (test-KmerGMA.jl:257-263)
    RV, ws, cons = O.strobe_gen_ref_ws_cons(TF)
    hits, loci, d = O.StrobeGMA(MINI_GENOME, RV, cons, windowsize=ws, thr=30, do_return_dists=True)
    assert [(h.first, h.last) for h in hits] == [(6852, 7140), (23907, 24201), (33845, 34133)]
    assert d.size == O.Fasta(MINI_GENOME).seqsize(0) - ws - 1                # StrobeGenomeMiner.jl:45: 1:(L-ws-1)
    with O.exact_arithmetic(84):
        he, _, de = O.StrobeGMA(MINI_GENOME, RV, cons, windowsize=ws, thr=30, do_return_dists=True)
    assert [(h.first, h.last) for h in he] == [(h.first, h.last) for h in hits] and np.max(np.abs(d - de)) < 1e-9


def test_partition_records_and_hit_merge_host_only():
    """the contig cut of a multi-GPU scan, host side only (no device): longest-first assignment of records to ranks, and
    kgma_hits_merge_partition -- blocks of [n][kgma_hit x n] with rank-local record numbers -> global record indices,
    GenomePos = summed length of the records in front (single mode: records shorter than the window do not count,
    GenomeMiner.jl:37-39,106; cluster mode: every record, OmnGenomeMiner.jl:159), ordered by record, order inside a record kept"""
    import kmergma_jl_b200 as K
    lens = [5000, 120, 9000, 3000, 289, 7000, 100, 4000]
    for world in (2, 3, 4):
        parts = K.partition_records(lens, world, tolerance=10.0)
        assert sorted(r for p in parts for r in p) == list(range(len(lens))) and all(p == sorted(p) for p in parts)
        load = [sum(lens[r] for r in p) for p in parts]
        assert max(load) - min(load) <= max(lens)                       # longest-first keeps the ranks within one record of each other
    assert K.partition_records([10, 1000], 2) is None and K.partition_records([5, 5], 3) is None
    parts = K.partition_records(lens, 3, tolerance=10.0)
    rng = np.random.default_rng(0)
    cap = 1 << 16
    blocks = np.zeros((3, cap), dtype=np.uint8)
    want = []
    for rank, p in enumerate(parts):
        n = 0
        hits = np.zeros(64, dtype=K.HIT_DT)
        for local, r in enumerate(p):
            for t in range(int(rng.integers(0, 5))):                 # several hits per record, in order
                hits[n]["record"], hits[n]["first"], hits[n]["last"], hits[n]["D"] = local, 10 * t + 1, 10 * t + 5, 1000 * r + t
                hits[n]["genome_pos"] = -12345                         # whatever the sub-genome's own count was: it is replaced
                want.append((r, 10 * t + 1, 1000 * r + t))
                n += 1
        blocks[rank, :8].view(np.int64)[0] = n
        blocks[rank, 16:16 + n * K.HIT_DT.itemsize] = hits[:n].view(np.uint8).reshape(-1)
    want.sort(key=lambda x: (x[0], x[1]))
    for min_len, counted in ((289, [l if l >= 289 else 0 for l in lens]), (0, lens)):
        out = K.merge_partition_hits(blocks, parts, lens, min_len)
        gp = np.concatenate([[0], np.cumsum(counted)])[:-1]
        assert [(int(h["record"]), int(h["first"]), int(h["D"])) for h in out] == want
        assert [int(h["genome_pos"]) for h in out] == [int(gp[r]) for r, _, _ in want]
    bad = blocks.copy(); bad[0, :8].view(np.int64)[0] = 10 ** 9       # a count that cannot fit the block is refused
    with pytest.raises(K.KmerGMAError):
        K.merge_partition_hits(bad, parts, lens, 0)


def test_masked_runs_match_numpy(tmp_path):
    """the masked-run list (what the extension / exact-match kernels consult instead of the ambiguity plane) against a
    numpy scan of the same sequences: runs at record starts/ends, one-base runs, runs spanning 32-base mask words, and a
    plane large enough for the multi-threaded slices, with one run laid across the slice edge (2^25 bases)"""
    import kmergma_jl_b200 as K
    rng = np.random.default_rng(4)
    small = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=5000)].copy()
    for a, b in ((0, 1), (5, 6), (31, 33), (64, 128), (200, 263), (1000, 1001), (4990, 5000)):
        small[a:b] = ord("N")
    big = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(1 << 25) + 300_000)].copy()
    big[(1 << 25) - 70_000:(1 << 25) + 41] = ord("N")              # crosses the slice edge whatever the record offset is
    big[:7] = ord("N"); big[-1:] = ord("N"); big[12345:12346] = ord("R")
    for s in rng.integers(100_000, big.size - 100_000, size=200):
        big[s:s + int(rng.integers(1, 200))] = ord("N")
    recs = [("a", small.tobytes().decode()), ("empty", ""), ("b", big.tobytes().decode()), ("c", "NNNN"), ("d", "ACGT")]
    g = K.Genome.from_records(recs)
    got = g.masked_runs()
    want = []
    for r, (_, s) in enumerate(recs):
        v = np.frombuffer(s.encode(), np.uint8)
        m = np.concatenate([[0], (~np.isin(v, np.frombuffer(b"ACGT", np.uint8))).astype(np.int8), [0]])
        d = np.diff(m)
        off = g.record_offset(r)
        want += [(off + int(a), off + int(b)) for a, b in zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1))]
    assert got.tolist() == [list(w) for w in want]
    # through the FASTA ingest as well
    p = tmp_path / "m.fasta"
    with open(p, "w") as fh:
        for d_, s in recs:
            fh.write(">" + d_ + "\n" + s + "\n")
    g2 = K.Genome.from_fasta(str(p))
    assert g2.masked_runs().tolist() == got.tolist()


def _random_fasta(path, seed):
    """messy FASTA text: random line widths, LF / CRLF, blank lines, lower case stretches, N runs, IUPAC codes, spaces and
    tabs inside sequence lines, records of 0 .. a few MB (several 4 MB packing tasks), with or without a final newline"""
    rng = np.random.default_rng(seed)
    recs = []
    with open(path, "wb") as fh:
        for r in range(int(rng.integers(1, 6))):
            L = int(rng.choice([0, 1, 31, 32, 33, 1000, int(rng.integers(1, 200_000)), int(rng.integers(4_000_000, 9_000_000))]))
            s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=L)].copy()
            for _ in range(int(rng.integers(0, 12))):
                if L < 2:
                    break
                a = int(rng.integers(0, L)); n = int(rng.choice([1, 2, 15, 16, 17, 31, 32, 33, 64, 1000, 70000]))
                kind = int(rng.integers(0, 3))
                if kind == 0:
                    s[a:a + n] = ord("N")
                elif kind == 1:
                    s[a:a + n] |= 0x20
                else:
                    s[a:a + min(n, 3)] = np.frombuffer(b"RYK", np.uint8)[:len(s[a:a + min(n, 3)])]
            eol = b"\r\n" if rng.random() < 0.3 else b"\n"
            width = int(rng.choice([1, 7, 31, 32, 33, 60, 64, 70, 80, 1000]))
            fh.write(b">rec%d desc %d" % (r, seed) + eol)
            text = s.tobytes()
            pieces = []
            for a in range(0, L, width):
                line = text[a:a + width]
                if rng.random() < 0.002 and len(line) > 4:
                    line = line[:2] + b" \t" + line[2:]                  # white space inside a sequence line is skipped
                pieces.append(line + eol)
                if rng.random() < 0.001:
                    pieces.append(eol)
            fh.write(b"".join(pieces))
            want = "".join(c if c in "ACGTN" else "?" for c in text.decode().upper())
            recs.append(want)
        if rng.random() < 0.5:
            fh.seek(0, 2)
            if fh.tell() > 0:
                fh.seek(-1, 2); fh.truncate()
    return recs


@pytest.mark.parametrize("seed", range(6))
def test_fasta_ingest_vectorised_and_portable_paths(tmp_path, seed):
    """the AVX2 packing pass and the portable one (KGMA_NO_AVX2, run in a child process: the choice is made once per
    process) produce the same container: sequences, lengths, masked runs"""
    import hashlib
    import subprocess
    import kmergma_jl_b200 as K
    p = str(tmp_path / "messy.fasta")
    want = _random_fasta(p, seed)
    # a final line cut by the truncation above may have lost its last byte: re-derive the expectation from the file itself
    text = open(p, "rb").read().decode()
    want = ["".join(c if c in "ACGTN" else "?" for c in "".join(body.split("\n")[1:]).replace("\r", "").replace(" ", "").replace("\t", "").upper())
            for body in text.split(">")[1:]]
    g = K.Genome.from_fasta(p)
    assert len(g) == len(want)
    for r, w in enumerate(want):
        assert g.seqsize(r) == len(w)
        assert g.seq(r) == w
    runs = g.masked_runs()
    ref = []
    for r, w in enumerate(want):
        v = np.frombuffer(w.encode(), np.uint8)
        m = np.concatenate([[0], (~np.isin(v, np.frombuffer(b"ACGT", np.uint8))).astype(np.int8), [0]])
        d = np.diff(m)
        off = g.record_offset(r)
        for a, b in zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1)):
            if ref and ref[-1][1] == off + int(a):
                ref[-1][1] = off + int(b)
            else:
                ref.append([off + int(a), off + int(b)])
    assert runs.tolist() == ref
    digest = hashlib.sha256(("|".join(g.seq(r) for r in range(len(g)))).encode() + runs.tobytes()).hexdigest()
    code = ("import sys, hashlib; sys.path.insert(0, %r); import kmergma_jl_b200 as K; g = K.Genome.from_fasta(%r); "
            "print(hashlib.sha256(('|'.join(g.seq(r) for r in range(len(g)))).encode() + g.masked_runs().tobytes()).hexdigest())") % (ROOT, p)
    out = subprocess.check_output([sys.executable, "-c", code], env=dict(os.environ, KGMA_NO_AVX2="1"), text=True).strip()
    assert out == digest


def test_fixed_operator_parameters_are_checked():
    """ac_gma_testing! / Omn_KmerGMA! take ScaleFactor, mask and Nt_bits (GenomeMiner.jl:12-14); the device path is fixed to
    1/k, 4^k - 1 and NUCLEOTIDE_BITS, so anything else is refused before any work is queued (no silent divergence)"""
    import kmergma_jl_b200 as K
    K._check_fixed_params(6, 1 / 6, 4095, {"A": 0, "C": 1, "G": 2, "T": 3, "N": 3})
    K._check_fixed_params(7, None, None, None)
    for kw in (dict(k=7, ScaleFactor=1 / 6), dict(k=5, mask=4095), dict(k=6, Nt_bits={"A": 0, "C": 1, "G": 2, "T": 2, "N": 3})):
        with pytest.raises(K.KmerGMAError) as e:
            K._check_fixed_params(kw["k"], kw.get("ScaleFactor"), kw.get("mask"), kw.get("Nt_bits"))
        assert e.value.code == K.L.E_UNSUPPORTED
    with pytest.raises(K.KmerGMAError):       # refused before a context is even asked for
        K.ac_gma_testing(genome_path=MINI_GENOME, refVec=np.zeros(4 ** 7), consensus_refseq="A" * 300, k=7, ScaleFactor=1 / 6, resultVec=[])


def test_cumulative_len_dict_golden():
    """test-KmerGMA.jl:336-343"""
    import kmergma_jl_b200 as K
    assert K.fasta_id_to_cumulative_len_dict(GENOME) == {
        "JQ684648.1 Lama glama clone V03 IgH locus genomic sequence": 0,
        "JQ684647.1 Lama glama clone F07 IgH locus genomic sequence": 121478,
        "AM773548.1 Lama pacos germline IgHV region, Vh3-S1, Vh2-S1 and vhh3-S1 genes": 444023,
        "AM773729.1 Lama pacos germline IgH locus: proximal IgHV region genes, complete IgHD region genes, complete IgHJ region genes "
        "and complete IgHC region genes": 221227}


def test_native_hit_headers_and_writer(tmp_path):
    """kgma_hit_header / kgma_result_write_fasta against the host mirror's append_hit! formatting and write_results
    (Alignment.jl:57-81, OmnGenomeMiner.jl:141-149, API.jl:234-241), incl. Julia's string(round(d, digits = 2)) on values that
    sit on rounding half-way points, print as integers, or need one decimal; the result comes from a host-only replay"""
    import kmergma_jl_b200 as K
    from test_multirank_gloo import RUN_DT
    g = K.Genome.from_fasta(GENOME)
    # golden of append_hit!: test-KmerGMA.jl:147-151
    h = np.zeros(1, dtype=K.HIT_DT)
    for d, want in ((69.1, "69.1"), (8.100000000000001, "8.1"), (24.874999, "24.87"), (9.0, "9.0"), (0.005, "0.0"), (0.015, "0.02"),
                    (2.675, "2.68"), (0.125, "0.12"), (0.375, "0.38"), (100.0, "100.0"), (12345.678, "12345.68"), (1e-9, "0.0")):
        h["record"], h["first"], h["last"], h["genome_pos"], h["dist"], h["profile"] = 1, 2, 5, 3, d, 4
        assert K.julia_float_str(K.julia_round2(d)) == want
        assert K.hit_header(g, h[0]) == "JQ684647.1 | dist = %s | MatchPos = 2:5 | GenomePos = 3 | Len = 4" % want
        assert K.hit_header(g, h[0], with_genome_pos=False) == "JQ684647.1 | dist = %s | MatchPos = 2:5 | Len = 4" % want
        assert K.hit_header(g, h[0], cluster=True) == "JQ684647.1 | Dist = %s | KFV = 4 | MatchPos = 2:5 | GenomePos = 3 | Len = 4" % want
    # a replayed result written natively == the mirror's records written by write_results
    RV, ws, cons = K.gen_ref_ws_cons(TF, 6)
    runs = np.array([(0, 0, 20300, 20340, 20330, 780000, 0, 0), (3, 0, 6790, 6830, 6801, 686000, 0, 0),
                     (3, 0, 23850, 23900, 23861, 2100000, 0, 0)], dtype=RUN_DT)
    fd = np.full(len(g), 10 ** 9, dtype=np.int64)
    out = K.replay_raw(g, [RV], [ws], [cons], [30.0], 6, K.L.MODE_SINGLE, 50, 0, -69, -1, runs.view(np.uint8), fd, host_only=True)
    assert len(out.hits) == 3
    native, mirror = tmp_path / "native.fasta", tmp_path / "mirror.fasta"
    assert K.write_hits(out, g, str(native), width=95) == 3
    recs = []
    K._emit(g, out, False, recs, None, None)
    K.write_results(recs, str(mirror), 95)
    assert native.read_text() == mirror.read_text() and native.read_text().count(">") == 3
    assert max(len(l) for l in native.read_text().splitlines() if not l.startswith(">")) == 95
    assert K.write_hits(out, g, str(native), width=60) == 3 and native.read_text().count(">") == 6      # appends, like open(path, "a")


def _header_text():
    txt = open(os.path.join(ROOT, "include", "kmergma.h")).read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def _header_structs():
    """{typedef name: [(C type, field name)]} of the anonymous POD structs in include/kmergma.h"""
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(kgma_\w+)\s*;", _header_text(), flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(?:const\s+)?(\w+)\s*(\**)\s*(.*)$", decl)
            ctype, ptr, names = m.group(1), m.group(2), m.group(3)
            for nm in names.split(","):
                nm = nm.strip()
                fields.append(("ptr" if ptr or nm.startswith("*") else ctype, nm.lstrip("* ")))
        out[name] = fields
    return out


def _header_arity():
    """{function: number of parameters} from the prototypes in include/kmergma.h"""
    out = {}
    for name, args in re.findall(r"\b(kgma_\w+)\s*\(([^()]*)\)\s*;", _header_text()):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_julia_shim_follows_the_header():
    """julia/KmerGMACuda.jl cannot run here (no Julia in the image): keep its struct mirrors and ccall signatures tied to
    include/kmergma.h textually -- same field order and widths, every ccall'ed symbol declared, same number of arguments."""
    src = open(os.path.join(ROOT, "julia", "KmerGMACuda.jl")).read()
    code = "\n".join(l.split("#", 1)[0] for l in src.splitlines())
    jl2c = {"Int32": "int32_t", "UInt32": "uint32_t", "Int64": "int64_t", "Float64": "double"}
    hs = _header_structs()
    seen = 0
    for jname, cname in re.findall(r"^struct\s+(\w+)\s*#\s*(kgma_\w+)", src, flags=re.M) + [("KgmaMatch", "kgma_match")]:
        body = re.search(r"struct\s+" + jname + r"\b(.*?)\bend\b", code, flags=re.S).group(1)
        jf = [(("ptr" if t.startswith("Ptr{") else jl2c[t]), n) for n, t in re.findall(r"(\w+)::([\w{}]+)", body)]
        assert [t for t, _ in jf] == [t for t, _ in hs[cname]], (jname, jf, hs[cname])
        assert [n for _, n in jf] == [n for _, n in hs[cname]], (jname, cname)
        seen += 1
    assert seen == 5
    arity, called = _header_arity(), set()
    for m in re.finditer(r"ccall\(\(:(kgma_\w+), LIB\),\s*[\w{}]+,\s*\(", code):
        name, i, depth = m.group(1), m.end(), 1
        start = i
        while depth:
            depth += {"(": 1, ")": -1}.get(code[i], 0)
            i += 1
        types = [t for t in re.split(r",(?![^{]*\})", code[start:i - 1]) if t.strip()]
        assert name in arity, name + " is not declared in include/kmergma.h"
        assert len(types) == arity[name], (name, types, arity[name])
        called.add(name)
    assert {"kgma_create", "kgma_scan", "kgma_exact_match", "kgma_genome_from_fasta", "kgma_result_hits"} <= called


def test_cigar_to_unitrange_text_form():
    """cigar_to_UnitRange (Alignment.jl:13-30) on CIGAR text: the goldens of test-KmerGMA.jl:129-136 (6:13, 6:15) through the
    oracle's pairalign, and the quirks -- `lower` is the first operation's count whatever it is, the last operation is never
    added, counts of I columns are"""
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    for subj, want in (("GGGGGATGCATGCAAAAA", (6, 13)), ("GGGGGATGCTTATGCAAAAA", (6, 15))):
        cig, score = O.pairalign_semiglobal("ATGCATGC", subj, -5, -1)
        assert K.cigar_to_UnitRange(cig) == want == tuple(O.cigar_to_UnitRange(cig))
        assert K.cigar_to_UnitRange(K.AlignResult(cig, score)) == want
    for cig in ("8=", "3=2I5=4D", "12D300=", "7I2=", "10D5=3I5=10D", "1D1=1D"):
        assert K.cigar_to_UnitRange(cig) == tuple(O.cigar_to_UnitRange(cig)), cig
    assert K.cigar_to_UnitRange("8=") == (1, 0) and K.cigar_to_UnitRange("3=2I5=4D") == (4, 10)


@pytest.mark.parametrize("seed", range(6))
def test_exact_match_merge_host_only(tmp_path, seed):
    """kgma_exact_match_merge needs no device: all occurrence starts of a query (found here with str.find, handed over in
    arbitrary order, as the slices of several GPUs would deliver them) -> exactMatch's result, overlapping and not
    (FindAllOverlap / FindAll, ExactMatch.jl:20-43), against the oracle.  Self-overlapping queries, copies across record
    edges (which do not match), duplicate identifiers (the later record wins, :112)."""
    import kmergma_jl_b200 as K
    from oracle import oracle as O
    rng = np.random.default_rng(40 + seed)
    q = ["ACACACA", "AAAA", "ACGTTGCAACGT", "GATTACAGATTACA", "T", "ACGTACGTAC"][seed]
    recs = []
    for r in range(5):
        s = list(np.asarray(list("ACGT"))[rng.integers(0, 4, size=int(rng.integers(30, 400)))])
        for _ in range(int(rng.integers(0, 6))):
            p = int(rng.integers(0, max(1, len(s) - 2 * len(q))))
            rep = (q * 3)[:int(rng.integers(len(q), 3 * len(q)))]               # runs of the query: self-overlaps
            s[p:p + len(rep)] = list(rep)
        recs.append(("dup" if r in (1, 3) else "r%d" % r, "".join(s)))
    recs.append(("tail", q[:len(q) // 2 + 1]))                                   # shorter than the query
    path = tmp_path / "em.fasta"
    with open(path, "w") as fh:
        for d_, s in recs:
            fh.write(">" + d_ + " x\n" + s + "\n")
    g = K.Genome.from_fasta(str(path))
    lib = K.L.load()
    starts = []
    for r, (_, s) in enumerate(recs):
        off = lib.kgma_genome_record_offset(g._h, r)
        p = s.find(q)
        while p >= 0:
            starts.append(off + p)
            p = s.find(q, p + 1)
    starts = np.asarray(starts, np.int64)[rng.permutation(len(starts))]
    f = O.Fasta(str(path))
    for overlap in (True, False):
        assert K.exact_match_merge(g, starts, len(q), overlap) == O.exactMatch(q, f, overlap=overlap), (q, overlap)
    assert K.exact_match_merge(g, np.zeros(0, np.int64), len(q), True) == "no match"
