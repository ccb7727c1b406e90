"""ctypes front-end of the CPU oracle (oracle/kmergma_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of kmergma_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (kmergma.jl_b200/) never does.

Function names mirror the reference (src/*.jl) so the golden tests read like the
reference's own test-suite (test/test_folder/test-KmerGMA.jl).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "kmergma_oracle.c")


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (the recipe is oracle/Makefile)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return _SO


class _Hit(C.Structure):
    _fields_ = [("record", C.c_int32), ("kfv", C.c_int32), ("dist", C.c_double),
                ("first", C.c_int64), ("last", C.c_int64), ("genome_pos", C.c_int64),
                ("cmi", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_fasta_read.restype = C.c_void_p
        L.orc_fasta_read.argtypes = [C.c_char_p]
        L.orc_fasta_free.argtypes = [C.c_void_p]
        L.orc_fasta_wrap.restype = C.c_void_p
        L.orc_fasta_wrap.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_fasta_n.argtypes = [C.c_void_p]
        L.orc_fasta_len.restype = C.c_int64
        L.orc_fasta_len.argtypes = [C.c_void_p, C.c_int]
        for f in (L.orc_fasta_ident, L.orc_fasta_desc):
            f.restype = C.c_char_p
            f.argtypes = [C.c_void_p, C.c_int]
        L.orc_fasta_seq.restype = C.c_void_p          # (a raw address: slices are copied out with string_at)
        L.orc_fasta_seq.argtypes = [C.c_void_p, C.c_int]
        L.orc_profile_new.restype = C.c_void_p
        L.orc_profile_new.argtypes = [C.c_int64]
        L.orc_profile_del.argtypes = [C.c_void_p]
        L.orc_profile_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.orc_profile_lengthen.argtypes = [C.c_void_p, C.c_int64]
        L.orc_profile_len.restype = C.c_int64
        L.orc_profile_len.argtypes = [C.c_void_p]
        L.orc_profile_counts.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_profile_consensus.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_kmer_count.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_kmer_count_add.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_kmer_dist_kfv.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.orc_kmer_dist_seq.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_double)]
        L.orc_gen_ref_ws_cons.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_char_p, C.POINTER(C.c_int64)]
        L.orc_get_cluster_index.argtypes = [C.c_double, C.c_void_p, C.c_int]
        L.orc_cluster_ref.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_char_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_semiglobal.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.orc_cigar_to_unitrange.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_align_unitrange.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_char_p, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_ac_gma.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_double,
                                 C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                 C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_int]
        L.orc_omn_gma.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int64,
                                  C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                  C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_exact_match.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_ac_gma_seq.restype = C.c_int64
        L.orc_ac_gma_seq.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_double,
                                     C.c_int64, C.c_void_p, C.c_int64]
        L.orc_strobe_count_add.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_strobe_gen_ref_ws_cons.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64),
                                                 C.c_char_p, C.POINTER(C.c_int64)]
        L.orc_strobe_gma.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                     C.c_double, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                     C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_set_exact.restype = None
        L.orc_set_exact.argtypes = [C.c_int, C.c_void_p]
        L.orc_synth.restype = None
        L.orc_synth.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p]
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc < 0:
        names = {-1: "io", -2: "KeyError: symbol outside A,C,G,T,N", -3: "capacity", -4: "argument"}
        raise OracleError(f"{what}: {names.get(rc, rc)}")


def _b(s) -> bytes:
    return s if isinstance(s, bytes) else str(s).upper().encode()


# ---------------------------------------------------------------- FASTA
class Fasta:
    """Stand-in for open(FASTA.Reader, path): all records of a file."""

    def __init__(self, path: str):
        self.path = path
        self._keep = None
        self._h = lib().orc_fasta_read(path.encode())
        if not self._h:
            raise OracleError(f"cannot read {path}")

    @classmethod
    def wrap(cls, description: str, residues: np.ndarray) -> "Fasta":
        """one record over an in-memory uint8 array of upper-case residues (not copied; kept alive by this object)"""
        self = cls.__new__(cls)
        self.path = None
        self._keep = np.ascontiguousarray(residues, dtype=np.uint8)
        self._h = lib().orc_fasta_wrap(description.encode(), self._keep.ctypes.data, self._keep.size)
        return self

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fasta_free(self._h)
            self._h = None

    def __len__(self):
        return lib().orc_fasta_n(self._h)

    def seq(self, r: int) -> str:
        return C.string_at(lib().orc_fasta_seq(self._h, r), self.seqsize(r)).decode()

    def subseq(self, r: int, first: int, last: int) -> str:
        """view(seq, first:last), 1-based inclusive, without materialising the whole record"""
        n = max(0, last - first + 1)
        return C.string_at(lib().orc_fasta_seq(self._h, r) + first - 1, n).decode() if n else ""

    def identifier(self, r: int) -> str:
        return lib().orc_fasta_ident(self._h, r).decode()

    def description(self, r: int) -> str:
        return lib().orc_fasta_desc(self._h, r).decode()

    def seqsize(self, r: int) -> int:
        return lib().orc_fasta_len(self._h, r)


# ---------------------------------------------------------------- Kmers.jl
def kmer_count(seq, k: int) -> np.ndarray:
    """src/Kmers.jl:14-28"""
    s = _b(seq)
    bins = np.zeros(4 ** k, dtype=np.float64)
    _check(lib().orc_kmer_count(s, len(s), k, bins.ctypes.data), "kmer_count")
    return bins


def kmer_count_add(seq, k: int, bins: np.ndarray) -> None:
    """src/Kmers.jl:33-44 kmer_count! (accumulates into bins)"""
    s = _b(seq)
    assert bins.dtype == np.float64 and bins.size == 4 ** k
    _check(lib().orc_kmer_count_add(s, len(s), k, bins.ctypes.data), "kmer_count!")


def kmer_dist(seq1, seq2_or_kfv, k: int) -> float:
    """src/Kmers.jl:54-60"""
    out = C.c_double()
    s1 = _b(seq1)
    if isinstance(seq2_or_kfv, np.ndarray):
        kfv = np.ascontiguousarray(seq2_or_kfv, dtype=np.float64)
        _check(lib().orc_kmer_dist_kfv(s1, len(s1), kfv.ctypes.data, k, C.byref(out)), "kmer_dist")
    else:
        s2 = _b(seq2_or_kfv)
        _check(lib().orc_kmer_dist_seq(s1, len(s1), s2, len(s2), k, C.byref(out)), "kmer_dist")
    return out.value


# ---------------------------------------------------------------- Consensus.jl
class Profile:
    """src/Consensus.jl:6-12"""

    def __init__(self, length: int):
        self._h = lib().orc_profile_new(length)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_profile_del(self._h)
            self._h = None

    @property
    def len(self) -> int:
        return lib().orc_profile_len(self._h)

    @property
    def vecs(self) -> List[List[int]]:
        out = []
        for b in range(4):
            a = np.zeros(self.len, dtype=np.int64)
            lib().orc_profile_counts(self._h, b, a.ctypes.data)
            out.append(a.tolist())
        return out


def add_consensus(p: Profile, seq) -> None:
    s = _b(seq)
    _check(lib().orc_profile_add(p._h, s, len(s)), "add_consensus!")


def lengthen(p: Profile, new_len: int) -> None:
    lib().orc_profile_lengthen(p._h, new_len)


def consensus_seq(p: Profile) -> str:
    buf = C.create_string_buffer(p.len + 1)
    lib().orc_profile_consensus(p._h, buf)
    return buf.value.decode()


# ---------------------------------------------------------------- ReferenceGeneration.jl
def gen_ref_ws_cons(refs, k: int, get_maxlen: bool = False):
    """src/ReferenceGeneration.jl:4-41 -> (RV, windowsize, consensus[, maxlen])"""
    f = refs if isinstance(refs, Fasta) else Fasta(refs)
    rv = np.zeros(4 ** k, dtype=np.float64)
    ws = C.c_int64()
    ml = C.c_int64()
    cap = max(f.seqsize(r) for r in range(len(f))) + 1
    cons = C.create_string_buffer(cap + 1)
    _check(lib().orc_gen_ref_ws_cons(f._h, k, rv.ctypes.data, C.byref(ws), cons, C.byref(ml)), "gen_ref_ws_cons")
    if get_maxlen:
        return rv, ws.value, cons.value.decode(), ml.value
    return rv, ws.value, cons.value.decode()


def get_cluster_index(inp: float, cutoffs: Sequence[float]) -> int:
    c = np.asarray(cutoffs, dtype=np.float64)
    return lib().orc_get_cluster_index(float(inp), c.ctypes.data, c.size)


def cluster_ref_API(refs, k: int, cutoffs=(7, 12, 20, 25), include_avg: bool = True,
                    eliminate_null: bool = False, get_dists: bool = False):
    """src/ReferenceGeneration.jl:75-138 -> (KFVs, windowsizes, consensus_vec, invalid_vec[, dists]).
    eliminate_null=True additionally applies eliminate_null_params (:152-168)."""
    f = refs if isinstance(refs, Fasta) else Fasta(refs)
    cut = np.asarray(cutoffs, dtype=np.float64)
    maxc = cut.size + 2
    nb = 4 ** k
    kfvs = np.zeros((maxc, nb), dtype=np.float64)
    wss = np.zeros(maxc, dtype=np.int64)
    stride = max(f.seqsize(r) for r in range(len(f))) + 2
    cons = C.create_string_buffer(maxc * stride)
    members = np.zeros(maxc, dtype=np.int32)
    invalid = np.zeros(maxc, dtype=np.int32)
    dists = np.zeros(len(f), dtype=np.float64)
    n = lib().orc_cluster_ref(f._h, k, cut.ctypes.data, cut.size, int(include_avg), int(eliminate_null),
                              kfvs.ctypes.data, wss.ctypes.data, cons, stride,
                              members.ctypes.data, invalid.ctypes.data, dists.ctypes.data)
    _check(n, "cluster_ref_API")
    raw = cons.raw
    cv = [raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode() for i in range(n)]
    ninv = cut.size + 1 + (1 if include_avg else 0)
    res = ([kfvs[i].copy() for i in range(n)], wss[:n].tolist(), cv, [bool(x) for x in invalid[:ninv]])
    if get_dists:
        return res + (dists,)
    return res


# ---------------------------------------------------------------- Alignment.jl
def pairalign_semiglobal(a, b, gap_open: int = -69, gap_extend: int = -1, prefer_extend: bool = True):
    """pairalign(SemiGlobalAlignment(), a, b, AffineGapScoreModel(EDNAFULL, ...)) -> (cigar, score)"""
    sa, sb = _b(a), _b(b)
    cap = len(sa) + len(sb) + 2
    ops = C.create_string_buffer(cap)
    cnt = np.zeros(cap, dtype=np.int32)
    sc = C.c_int64()
    n = lib().orc_semiglobal(sa, len(sa), sb, len(sb), gap_open, gap_extend, int(prefer_extend),
                             ops, cnt.ctypes.data, cap, C.byref(sc))
    _check(n, "pairalign")
    o = ops.raw[:n].decode()
    return "".join(f"{cnt[i]}{o[i]}" for i in range(n)), sc.value


def cigar_to_UnitRange(cigar: str) -> Tuple[int, int]:
    """src/Alignment.jl:13-30 on a CIGAR string -> (first, last) of the Julia UnitRange"""
    import re
    cnt = np.asarray([int(x) for x in re.findall(r"(\d+)[=XIDM]", cigar)], dtype=np.int32)
    lo, hi = C.c_int64(), C.c_int64()
    lib().orc_cigar_to_unitrange(cnt.ctypes.data, cnt.size, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def align_unitrange(seq, rng: Tuple[int, int], consensus, windowsize: int, sequence_length: int,
                    gap_open: int = -69, gap_extend: int = -1, prefer_extend: bool = True) -> Tuple[int, int]:
    """src/Alignment.jl:33-52"""
    s, c = _b(seq), _b(consensus)
    a, b = C.c_int64(), C.c_int64()
    _check(lib().orc_align_unitrange(s, sequence_length, rng[0], rng[1], c, windowsize,
                                     gap_open, gap_extend, int(prefer_extend), C.byref(a), C.byref(b)), "align_unitrange")
    return a.value, b.value


# ---------------------------------------------------------------- record formatting
def julia_round2(x: float) -> float:
    """Julia round(x, digits=2): round-half-even of x*100, divided by 100 (Base._round_invstep)."""
    return float(np.rint(x * 100.0) / 100.0)


def julia_float_str(x: float) -> str:
    """Julia string(::Float64): shortest round-trip decimal with a fractional digit (plain range)."""
    return repr(float(x))


@dataclass
class Hit:
    record: int
    identifier: str
    kfv: int
    dist: float
    first: int
    last: int
    genome_pos: int
    cmi: int
    seq: str

    @property
    def length(self) -> int:
        return self.last - self.first + 1

    def description(self, genome_pos: bool = True) -> str:
        """src/Alignment.jl:57-81 append_hit! / src/OmnGenomeMiner.jl:141-149 header text"""
        d = julia_float_str(julia_round2(self.dist))
        if self.kfv:
            return (f"{self.identifier} | Dist = {d} | KFV = {self.kfv} | MatchPos = {self.first}:{self.last}"
                    f" | GenomePos = {self.genome_pos} | Len = {self.length}")
        gp = f" | GenomePos = {self.genome_pos}" if genome_pos else ""
        return f"{self.identifier} | dist = {d} | MatchPos = {self.first}:{self.last}{gp} | Len = {self.length}"


def _hits(f: Fasta, arr, n) -> List[Hit]:
    out = []
    for i in range(n):
        h = arr[i]
        out.append(Hit(h.record, f.identifier(h.record), h.kfv, h.dist, h.first, h.last,
                       h.genome_pos, h.cmi, f.subseq(h.record, h.first, h.last)))
    return out


class exact_arithmetic:
    """with exact_arithmetic(N) / exact_arithmetic([n_1..n_C]): run ac_gma_testing / Omn_KmerGMA with the
    rounding-free integer distance D = d*2kN^2 (see orc_set_exact in kmergma_oracle.c) instead of the reference's
    accumulated Float64.  Used to check the device at exact ties (d == thr, equal run minima), where the
    reference's own outcome depends on its rounding history."""

    def __init__(self, N):
        self.N = np.atleast_1d(np.asarray(N, dtype=np.int32))

    def __enter__(self):
        lib().orc_set_exact(int(self.N.size), self.N.ctypes.data)
        return self

    def __exit__(self, *a):
        lib().orc_set_exact(0, None)
        return False


# ---------------------------------------------------------------- GenomeMiner.jl
def ac_gma_testing(genome, refVec: np.ndarray, consensus_refseq: str, k: int = 6, windowsize: int = 289,
                   thr: float = 33.5, buff: int = 50, do_align: bool = True,
                   gap_open_score: int = -69, gap_extend_score: int = -1,
                   do_return_dists: bool = False, prefer_extend: bool = True,
                   only_record: int = -1, hit_cap: int = 1 << 20):
    """src/GenomeMiner.jl:4-109.  Returns (hits, hit_loci_vec, dist_vec|None)."""
    f = genome if isinstance(genome, Fasta) else Fasta(genome)
    rv = np.ascontiguousarray(refVec, dtype=np.float64)
    assert rv.size == 4 ** k
    arr = (_Hit * hit_cap)()
    nh, nd = C.c_int64(), C.c_int64()
    dist = None
    dcap = 0
    if do_return_dists:
        dcap = sum(max(0, f.seqsize(r) - windowsize) for r in range(len(f)))
        dist = np.zeros(max(dcap, 1), dtype=np.float64)
    cons = _b(consensus_refseq)
    rc = lib().orc_ac_gma(f._h, rv.ctypes.data, cons, len(cons), k, windowsize, float(thr), buff, int(do_align),
                          gap_open_score, gap_extend_score, int(prefer_extend),
                          arr, hit_cap, C.byref(nh),
                          dist.ctypes.data if dist is not None else None, dcap, C.byref(nd), only_record)
    _check(rc, "ac_gma_testing!")
    hits = _hits(f, arr, nh.value)
    loci = [h.first + h.genome_pos for h in hits]
    return hits, loci, (dist[:nd.value] if dist is not None else None)


def record_KmerGMA(genome, record: int, refVec, consensus_refseq, **kw):
    """src/MultiThread/GenomeMiner.jl:8-98 — ac_gma_testing! restricted to one record
    (header printed without GenomePos: Hit.description(genome_pos=False))."""
    kw.setdefault("thr", 30)
    hits, _, _ = ac_gma_testing(genome, refVec, consensus_refseq, only_record=record, **kw)
    return hits


# ---------------------------------------------------------------- StrobemerGMA/ (experimental; scan unpinned by the reference)
def ungapped_strobe_2_mer_count(seq, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5) -> np.ndarray:
    """Strobemers.jl:90-103"""
    b = _b(seq)
    bins = np.zeros(4 ** (2 * s), dtype=np.float64)
    _check(lib().orc_strobe_count_add(b, len(b), s, w_min, w_max, q, bins.ctypes.data), "ungapped_strobe_2_mer_count")
    return bins


def strobe_gen_ref_ws_cons(refs, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5):
    """StrobeRefGen.jl:4-42 -> (RV, windowsize, consensus)"""
    f = refs if isinstance(refs, Fasta) else Fasta(refs)
    rv = np.zeros(4 ** (2 * s), dtype=np.float64)
    ws = C.c_int64()
    ml = C.c_int64()
    cap = max(f.seqsize(r) for r in range(len(f))) + 1
    cons = C.create_string_buffer(cap + 1)
    _check(lib().orc_strobe_gen_ref_ws_cons(f._h, s, w_min, w_max, q, rv.ctypes.data, C.byref(ws), cons, C.byref(ml)), "gen_ref_ws_cons")
    return rv, ws.value, cons.value.decode()


def StrobeGMA(genome, refVec: np.ndarray, consensus_refseq: str, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5,
              windowsize: int = 289, thr: float = 33.5, buff: int = 50, do_align: bool = True,
              gap_open_score: int = -69, gap_extend_score: int = -5, score_threshold: int = 0,
              do_return_dists: bool = False, prefer_extend: bool = True, hit_cap: int = 1 << 20):
    """StrobeGenomeMiner.jl:5-95 (ScaleFactor = 1/(w_max+s-1) as Strobemer_findGenes passes it).
    Returns (hits, hit_loci_vec, dist_vec|None)."""
    f = genome if isinstance(genome, Fasta) else Fasta(genome)
    rv = np.ascontiguousarray(refVec, dtype=np.float64)
    assert rv.size == 4 ** (2 * s)
    arr = (_Hit * hit_cap)()
    nh, nd = C.c_int64(), C.c_int64()
    dist, dcap = None, 0
    if do_return_dists:
        dcap = sum(max(0, f.seqsize(r) - windowsize - 1) for r in range(len(f)))
        dist = np.zeros(max(dcap, 1), dtype=np.float64)
    cons = _b(consensus_refseq)
    rc = lib().orc_strobe_gma(f._h, rv.ctypes.data, cons, len(cons), s, w_min, w_max, q, windowsize, float(thr), buff, int(do_align),
                              gap_open_score, gap_extend_score, int(prefer_extend), C.c_int64(score_threshold),
                              arr, hit_cap, C.byref(nh), dist.ctypes.data if dist is not None else None, dcap, C.byref(nd))
    _check(rc, "StrobeGMA!")
    hits = _hits(f, arr, nh.value)
    loci = [h.first + h.genome_pos for h in hits]
    return hits, loci, (dist[:nd.value] if dist is not None else None)


# ---------------------------------------------------------------- OmnGenomeMiner.jl
def Omn_KmerGMA(genome, refVecs, windowsizes, consensus_seqs, k: int = 6,
                thr_vec=(35, 31, 38, 34, 27, 27), buff: int = 50, align_hits: bool = True,
                gap_open_score: int = -200, gap_extend_score: int = -1,
                do_return_dists: bool = False, prefer_extend: bool = True, hit_cap: int = 1 << 20):
    """src/OmnGenomeMiner.jl:7-162.  Returns (hits, hit_loci_vec, dist_vec_vec|None)."""
    f = genome if isinstance(genome, Fasta) else Fasta(genome)
    Cn = len(windowsizes)
    nb = 4 ** k
    rvs = np.ascontiguousarray(np.stack([np.asarray(v, dtype=np.float64) for v in refVecs]))
    assert rvs.shape == (Cn, nb)
    wss = np.asarray(windowsizes, dtype=np.int64)
    thr = np.asarray(thr_vec, dtype=np.float64)
    assert thr.size >= Cn
    stride = max(len(c) for c in consensus_seqs) + 1
    cons = C.create_string_buffer(Cn * stride)
    for i, c in enumerate(consensus_seqs):
        bb = _b(c)
        cons[i * stride:i * stride + len(bb)] = bb
    arr = (_Hit * hit_cap)()
    nh = C.c_int64()
    dist = None
    nd = np.zeros(Cn, dtype=np.int64)
    dcap = 0
    if do_return_dists:
        dcap = max(1, sum(f.seqsize(r) for r in range(len(f))))
        dist = np.zeros((Cn, dcap), dtype=np.float64)
    rc = lib().orc_omn_gma(f._h, rvs.ctypes.data, wss.ctypes.data, Cn, cons, stride, k, thr.ctypes.data, buff,
                           int(align_hits), gap_open_score, gap_extend_score, int(prefer_extend),
                           arr, hit_cap, C.byref(nh),
                           dist.ctypes.data if dist is not None else None, dcap, nd.ctypes.data)
    _check(rc, "Omn_KmerGMA!")
    hits = _hits(f, arr, nh.value)
    loci = [h.first + h.genome_pos for h in hits]
    dv = [dist[q, :nd[q]].copy() for q in range(Cn)] if dist is not None else None
    return hits, loci, dv


class align_events:
    """with align_events() as ev: Omn_KmerGMA(...); ev.list -> [(record, profile, CMI, hit_left, hit_right, emitted)] for every
    extension performed, in order (what `get_aligns` pushes, OmnGenomeMiner.jl:131-133)"""

    def __init__(self, cap: int = 1 << 16):
        self.buf = np.zeros(6 * cap, dtype=np.int64)
        self.n = C.c_int64(0)
        self.cap = cap

    def __enter__(self):
        lib().orc_set_event_sink(C.c_void_p(self.buf.ctypes.data), C.c_int64(self.cap), C.byref(self.n))
        return self

    def __exit__(self, *a):
        lib().orc_set_event_sink(None, C.c_int64(0), None)
        return False

    @property
    def list(self):
        return [tuple(int(x) for x in self.buf[6 * i:6 * i + 6]) for i in range(self.n.value)]


# ---------------------------------------------------------------- ExactMatch.jl
def exactMatch(query, subject, overlap: bool = True):
    """src/ExactMatch.jl:89-121.  subject: sequence string -> list of (first,last) or None;
    subject: Fasta -> dict identifier -> list of (first,last), or the string "no match"."""
    q = _b(query)

    def one(s: bytes):
        cap = max(1, len(s))
        out = np.zeros(cap, dtype=np.int64)
        n = C.c_int64()
        _check(lib().orc_exact_match(q, len(q), s, len(s), int(overlap), out.ctypes.data, cap, C.byref(n)), "exactMatch")
        if n.value == 0:
            return None
        return [(int(a), int(a) + len(q) - 1) for a in out[:n.value]]

    if isinstance(subject, Fasta):
        d = {}
        for r in range(len(subject)):
            rm = one(subject.seq(r).encode())
            if rm is not None:
                d[subject.identifier(r)] = rm
        return d if d else "no match"
    return one(_b(subject))


def synth(seed: int, p0: int, n: int) -> bytearray:
    """residues p0..p0+n-1 of the synthetic genome (same generator as kgma_genome_synth)"""
    buf = bytearray(n)
    cbuf = (C.c_char * n).from_buffer(buf)
    lib().orc_synth(seed, p0, n, C.addressof(cbuf))
    del cbuf
    return buf


def ac_gma_seq_count(seq: bytes, RV: np.ndarray, k: int, ws: int, thr: float, buff: int = 50) -> int:
    """bench helper: hot loop of GenomeMiner.jl:60-104 (no alignment) on an in-memory sequence."""
    rv = np.ascontiguousarray(RV, dtype=np.float64)
    cap = 1 << 16
    arr = (_Hit * cap)()
    n = lib().orc_ac_gma_seq(seq, len(seq), rv.ctypes.data, k, ws, float(thr), buff, arr, cap)
    _check(int(n), "ac_gma_seq")
    return int(n)
