"""Pins the CPU oracle (oracle/) against every deterministic golden vector the reference's
own test-suite holds for the hot path (test/test_folder/test-KmerGMA.jl).
Line numbers in comments are into that file."""
import numpy as np
import pytest

from conftest import TF, MINI_GENOME, GENOME, EIGHT, TEST_CONSENSUS
from oracle import oracle as O

TEST_SEQ = "ATGCATGC"                         # runtests.jl:47
TEST_KFV = [0, 0, 0, 2, 1, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0]   # runtests.jl:50-51


def test_kmer_count():                        # :2-15
    assert O.kmer_count(TEST_SEQ, 1).tolist() == [2, 2, 2, 2]
    assert O.kmer_count(TEST_SEQ, 2).tolist() == TEST_KFV
    b = np.zeros(16)
    O.kmer_count_add(TEST_SEQ, 2, b)
    assert b.tolist() == TEST_KFV


def test_kmer_dist():                         # :17-20
    s = TEST_SEQ * 25
    assert O.kmer_dist(s + "A" + s, s + "G" + s, 2) == 1.0
    assert O.kmer_dist(s + "AA" + s, s + "GT" + s, 2) == 2.0


def test_consensus_profile():                 # :28-46
    assert O.Profile(2).vecs == [[0, 0]] * 4
    a = O.Profile(8)
    O.add_consensus(a, TEST_SEQ)
    assert a.vecs == [[1, 0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 1], [0, 0, 1, 0, 0, 0, 1, 0], [0, 1, 0, 0, 0, 1, 0, 0]]
    assert a.len == 8
    O.lengthen(a, 9)
    assert a.vecs == [[1, 0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 1, 0, 0, 0, 1, 0, 0], [0, 1, 0, 0, 0, 1, 0, 0, 0]]
    assert a.len == 9
    O.add_consensus(a, TEST_SEQ[:7] + "G")
    O.add_consensus(a, TEST_SEQ[:7] + "G")
    assert a.vecs == [[3, 0, 0, 0, 3, 0, 0, 0, 0], [0, 0, 0, 3, 0, 0, 0, 1, 0], [0, 0, 3, 0, 0, 0, 3, 2, 0], [0, 3, 0, 0, 0, 3, 0, 0, 0]]
    assert O.consensus_seq(a)[:8] == TEST_SEQ[:7] + "G"


def test_gen_ref_ws_cons():                   # :49-68
    rv, ws, cons = O.gen_ref_ws_cons(TF, 1)
    assert rv.tolist() == [63.25, 73.70238095238095, 89.26190476190476, 62.38095238095238]
    assert ws == 289 and cons == TEST_CONSENSUS
    assert O.gen_ref_ws_cons(TF, 1, get_maxlen=True)[3] == 299
    assert O.gen_ref_ws_cons(TF, 2)[0].tolist() == [
        11.178571428571429, 15.964285714285714, 24.154761904761905, 11.88095238095238, 22.76190476190476,
        17.904761904761905, 8.154761904761905, 24.88095238095238, 18.607142857142858, 22.202380952380953,
        30.369047619047617, 18.07142857142857, 10.702380952380953, 17.047619047619047, 26.166666666666664,
        7.5476190476190474]
    assert O.gen_ref_ws_cons(TF, 6)[0][4:10].tolist() == [0.011904761904761904, 0.023809523809523808, 0.0, 0.0,
                                                           0.023809523809523808, 0.0]


K1 = [[62.785714285714285, 72.78571428571429, 89.78571428571429, 62.642857142857146],
      [63.13333333333333, 71.33333333333333, 90.53333333333333, 62.6],
      [63.5, 70.71428571428571, 90.78571428571429, 64.07142857142857],
      [62.54545454545455, 68.72727272727273, 91.36363636363636, 64.54545454545455],
      [63.666666666666664, 78.53333333333333, 86.9, 60.56666666666667]]


def test_cluster_ref_API():                   # :70-110
    assert O.get_cluster_index(5, [1, 2, 6, 10]) == 3
    assert O.get_cluster_index(12, [1, 2, 6, 10]) == 5
    assert O.get_cluster_index(0, [1, 2, 6, 10]) == 1
    a = O.cluster_ref_API(TF, 1, cutoffs=[7, 12, 20, 25], include_avg=False)
    assert [v.tolist() for v in a[0]] == K1
    assert a[1] == [288, 288, 289, 287, 290]
    assert len(a[2]) == 5 and a[2][0][:4] == "CAGG"
    assert a[3] == [False] * 5
    a = O.cluster_ref_API(TF, 1, cutoffs=[7, 12, 20, 25])
    assert [v.tolist() for v in a[0]] == K1 + [[63.25, 73.70238095238095, 89.26190476190476, 62.38095238095238]]
    assert a[1] == [288, 288, 289, 287, 290, 289]
    assert len(a[2]) == 6 and a[3] == [False] * 6
    rvs, wss, cons, inv = O.cluster_ref_API(TF, 6, cutoffs=[7, 12, 20, 25], eliminate_null=True)
    assert wss == [288, 288, 288, 289, 290, 289]
    assert len(rvs) == len(cons) == 6


def test_cigar_to_UnitRange():                # :129-136
    cg, _ = O.pairalign_semiglobal("ATGCATGC", "GGGGGATGCATGCAAAAA", -5, -1)
    assert cg == "5D8=5D" and O.cigar_to_UnitRange(cg) == (6, 13)
    cg, _ = O.pairalign_semiglobal("ATGCATGC", "GGGGGATGCTTATGCAAAAA", -5, -1)
    assert cg == "5D4=2D4=5D" and O.cigar_to_UnitRange(cg) == (6, 15)


def test_rss_cigar():                         # :154-156 (src/RSS.jl:11-20; N==N is '=' under isequal)
    rssv = "CACAGTG" + "N" * 12 + "ACAAAAACC"
    cg, _ = O.pairalign_semiglobal(rssv, TEST_SEQ + rssv + TEST_SEQ, -69, -1)
    assert cg == "8D28=8D"


def test_align_unitrange():                   # :138-145
    f = O.Fasta(EIGHT)
    assert O.align_unitrange(f.seq(0), (450, 900), TEST_CONSENSUS, 289, 1000) == (501, 789)


def test_append_hit_format():                 # :147-151
    h = O.Hit(0, "foo", 0, 69.1, 2, 5, 3, 0, "TGCA")
    assert h.description() == "foo | dist = 69.1 | MatchPos = 2:5 | GenomePos = 3 | Len = 4"


@pytest.fixture(scope="module")
def profile6():
    return O.gen_ref_ws_cons(TF, 6)


def test_ac_gma_no_align(profile6):           # :167-177
    rv, ws, cons = profile6
    hits, _, _ = O.ac_gma_testing(GENOME, rv, cons, windowsize=ws, thr=30, do_align=False)
    assert len(hits) == 7
    d = [h.description() for h in hits]
    assert d[1] == "JQ684648.1 | dist = 9.21 | MatchPos = 20380:20768 | GenomePos = 0 | Len = 389"
    assert d[-3] == "AM773548.1 | dist = 8.1 | MatchPos = 6807:7195 | GenomePos = 444023 | Len = 389"


def test_ac_gma_align(profile6):              # :179-193
    rv, ws, cons = profile6
    hits, loci, _ = O.ac_gma_testing(GENOME, rv, cons, windowsize=ws, thr=30, do_align=True)
    assert len(hits) == 7
    assert loci == [8543, 20425, 221912, 234018, 450875, 467930, 477868]
    d = [h.description() for h in hits]
    assert d[1] == "JQ684648.1 | dist = 9.21 | MatchPos = 20425:20713 | GenomePos = 0 | Len = 289"
    assert d[-3] == "AM773548.1 | dist = 8.1 | MatchPos = 6852:7140 | GenomePos = 444023 | Len = 289"
    assert d[5] == "AM773548.1 | dist = 24.87 | MatchPos = 23907:24201 | GenomePos = 444023 | Len = 295"


def test_ac_gma_dists(profile6):              # :195-211
    rv, ws, cons = profile6
    hits, _, dist = O.ac_gma_testing(GENOME, rv, cons, windowsize=ws, thr=10, do_align=False, do_return_dists=True)
    assert len(dist) == 484127
    assert round(float(np.mean(dist))) == 46
    assert len(hits) == 3
    assert hits[0].description() == "JQ684648.1 | dist = 9.21 | MatchPos = 20380:20768 | GenomePos = 0 | Len = 389"
    assert hits[-1].description() == "AM773548.1 | dist = 8.1 | MatchPos = 6807:7195 | GenomePos = 444023 | Len = 389"


def test_omn_buff200():                       # :214-227
    rvs, wss, cons, inv = O.cluster_ref_API(TF, 6, cutoffs=[7, 12, 20, 25], include_avg=False)
    hits, _, _ = O.Omn_KmerGMA(MINI_GENOME, rvs, wss, cons, buff=200, thr_vec=[37, 33, 38, 34, 28, 27])
    assert [h.description() for h in hits] == [
        "AM773548.1 | Dist = 20.17 | KFV = 3 | MatchPos = 6852:7139 | GenomePos = 0 | Len = 288",
        "AM773548.1 | Dist = 33.96 | KFV = 4 | MatchPos = 23907:24198 | GenomePos = 0 | Len = 292",
        "AM773548.1 | Dist = 26.17 | KFV = 3 | MatchPos = 33845:34132 | GenomePos = 0 | Len = 288"]


def test_record_KmerGMA(profile6):            # :229-250
    rv, ws, cons = profile6
    hits = O.record_KmerGMA(MINI_GENOME, 0, rv, cons, thr=30)
    assert [h.description(genome_pos=False) for h in hits] == [
        "AM773548.1 | dist = 8.1 | MatchPos = 6852:7140 | Len = 289",
        "AM773548.1 | dist = 24.87 | MatchPos = 23907:24201 | Len = 295",
        "AM773548.1 | dist = 10.99 | MatchPos = 33845:34133 | Len = 289"]


def test_findGenes_cluster_mode_golden():     # :265-271 (API.jl:161-226 with explicit thresholds, buffer 100)
    rvs, wss, cons, inv = O.cluster_ref_API(TF, 6, cutoffs=[7, 12, 20, 25], eliminate_null=True)
    hits, _, _ = O.Omn_KmerGMA(MINI_GENOME, rvs, wss, cons, buff=100, thr_vec=[35, 31, 38, 34, 27, 27])
    assert [h.description() for h in hits] == [
        "AM773548.1 | Dist = 20.17 | KFV = 3 | MatchPos = 6852:7139 | GenomePos = 0 | Len = 288",
        "AM773548.1 | Dist = 33.96 | KFV = 4 | MatchPos = 23907:24193 | GenomePos = 0 | Len = 287",
        "AM773548.1 | Dist = 26.17 | KFV = 3 | MatchPos = 33845:34132 | GenomePos = 0 | Len = 288"]


def test_exactMatch_seq():                    # :299-306
    assert O.exactMatch("GAG", "CCCCCCCGAGCTTTT") == [(8, 10)]
    assert O.exactMatch("GAG", "CGAGCCCGAGCTTTT") == [(2, 4), (8, 10)]
    assert O.exactMatch("GAG", "CGAGAGAGAAGGCCGAGCTTTT") == [(2, 4), (4, 6), (6, 8), (15, 17)]
    assert O.exactMatch("GAG", "CGAGAGAGAAGGCCGAGCTTTT", overlap=False) == [(2, 4), (6, 8), (15, 17)]
    assert O.exactMatch("GAG", "CCCCCCTTT") is None


def test_exactMatch_reader():                 # :308-334
    f = O.Fasta(TF)
    sub = f.seq(0)[41:69]
    assert O.exactMatch(sub, f) == {"AM773729|IGHV1-1*01|Vicugna": [(42, 69)]}
    assert O.exactMatch(f.seq(0), f) == {"AM773729|IGHV1-1*01|Vicugna": [(1, 296)]}
    assert O.exactMatch("AAAAAAAAA", f) == "no match"
    assert O.exactMatch("AAATT", f) == {"AM773729|IGHV1-1*01|Vicugna": [(174, 178)],
                                         "AM939700|IGHV1S5*01|Vicugna": [(174, 178)]}


def test_exact_arithmetic_mode_agrees_with_faithful_on_goldens():
    """O.exact_arithmetic runs the same state machine on the rounding-free integer distance; away from exact ties it
    must give the faithful Float64 restatement's hits (test-KmerGMA.jl:174-176,189-192,223-225 inputs)."""
    from oracle import oracle as O
    RV, ws, cons = O.gen_ref_ws_cons(TF, 6)
    for thr, align in ((30, False), (30, True), (10, False)):
        a = O.ac_gma_testing(GENOME, RV, cons, windowsize=ws, thr=thr, do_align=align)[0]
        with O.exact_arithmetic(84):
            b = O.ac_gma_testing(GENOME, RV, cons, windowsize=ws, thr=thr, do_align=align)[0]
        assert [h.description() for h in a] == [h.description() for h in b]
        assert all(abs(x.dist - y.dist) <= 1e-9 * x.dist for x, y in zip(a, b))
    cl = O.cluster_ref_API(TF, 6, include_avg=False)
    rvs, wss, cs = cl[0], cl[1], cl[2]
    a = O.Omn_KmerGMA(MINI_GENOME, rvs, wss, cs, thr_vec=[37, 33, 38, 34, 28, 27], buff=200)[0]
    with O.exact_arithmetic([14, 52, 1, 5, 12]):
        b = O.Omn_KmerGMA(MINI_GENOME, rvs, wss, cs, thr_vec=[37, 33, 38, 34, 28, 27], buff=200)[0]
    assert [h.description() for h in a] == [h.description() for h in b]
    assert a[0].description() == "AM773548.1 | Dist = 20.17 | KFV = 3 | MatchPos = 6852:7139 | GenomePos = 0 | Len = 288"


def test_oracle_records_every_cluster_mode_extension():
    """the `get_aligns` sink (OmnGenomeMiner.jl:131-133): one event per extension performed, the emitted ones being the hits"""
    rvs, wss, cs, inv = O.cluster_ref_API(TF, 6, eliminate_null=True)
    with O.align_events() as sink:
        hits = O.Omn_KmerGMA(MINI_GENOME, rvs, wss, cs, thr_vec=[35, 31, 38, 34, 27, 27], buff=100)[0]
    ev = sink.list
    assert len(hits) == 3 and len(ev) >= 3
    assert [(e[0], e[1], e[2]) for e in ev if e[5]] == [(h.record, h.kfv, h.cmi) for h in hits]
    with O.align_events() as sink:
        O.Omn_KmerGMA(MINI_GENOME, rvs, wss, cs, thr_vec=[35, 31, 38, 34, 27, 27], buff=100, align_hits=False)
    assert sink.list == []                                            # no extension, nothing pushed
