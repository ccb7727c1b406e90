"""kmergma.jl_b200 — host-side mirror of KmerGMA.jl's operator interface over libkmergma_cuda.

Julia is not available in this image, so the host side above the C ABI is written in Python and
mirrors the reference's names, keyword arguments, in-place output vectors and error behaviour
(src/API.jl, src/GenomeMiner.jl, src/OmnGenomeMiner.jl, src/ExactMatch.jl,
src/ReferenceGeneration.jl); the Julia shim a maintainer would add is in julia/ and INTEGRATION.md.
All scanning, extension and matching work happens in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib as L
from .exchange import HostExchange

__all__ = [
    "KmerGMAError", "Context", "Genome", "FastaRecord", "KFV", "AlignResult", "HostExchange", "partition_records", "merge_partition_hits", "PartitionMerger",
    "gen_ref_ws_cons", "cluster_ref_API", "eliminate_null_params", "get_cluster_index",
    "estimate_optimal_threshold", "ac_gma_testing", "Omn_KmerGMA", "record_KmerGMA",
    "findGenes", "findGenes_cluster_mode", "exactMatch", "write_results", "write_hits", "hit_header",
    "fasta_id_to_cumulative_len_dict", "align_unitrange", "cigar_to_UnitRange", "firstMatch",
    "julia_round2", "julia_float_str", "kmer_count", "kmer_dist", "as_UInt", "as_kmer",
    "randstrobe_score", "get_strobe_2_mer", "ungapped_strobe_2_mer_count", "strobe_gen_ref_ws_cons", "StrobeGMA", "Strobemer_findGenes",
]


_log = logging.getLogger("KmerGMA")        # the reference's @info lines (verbose = true) go here


class KmerGMAError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[kgma {code}] {msg}")
        self.code = code


# ------------------------------------------------------------------------------------------------
class Context:
    """One kgma_ctx = one GPU. No global state lives in the library; this module keeps one default
    context per device for convenience."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.kgma_create(device, C.byref(h))
        if rc != 0:
            raise KmerGMAError(rc, (self._lib.kgma_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kgma_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            raise KmerGMAError(rc, (self._lib.kgma_last_error(self._h) or b"").decode() or what)

    def stats(self) -> dict:
        s = L.Stats()
        self._lib.kgma_get_stats(self._h, C.byref(s))
        return {f: getattr(s, f) for f, _ in L.Stats._fields_}


_default_ctx: Dict[int, Context] = {}


def default_context(device: Optional[int] = None) -> Context:
    if device is None:
        device = int(os.environ.get("KGMA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ------------------------------------------------------------------------------------------------
@dataclass
class FastaRecord:
    """Stand-in for FASTX.FASTA.Record: `description` is the whole header, `identifier` the text before
    the first whitespace, `sequence` the residues."""
    description: str
    sequence: str

    @property
    def identifier(self) -> str:
        return self.description.split(None, 1)[0] if self.description.strip() else ""

    def __eq__(self, other):
        return isinstance(other, FastaRecord) and self.description == other.description and \
            self.sequence.upper() == other.sequence.upper()


@dataclass
class AlignResult:
    """What do_return_align yields per hit (the reference pushes a BioAlignments PairwiseAlignmentResult)."""
    cigar: str
    score: int


class Genome:
    """Packed genome (2 bits/base + ambiguity mask + record table) owned by the C library."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib
        self._resident_ctx = None          # the Context whose GPU was last asked to hold this genome (make_resident)

    @classmethod
    def from_fasta(cls, path: str) -> "Genome":
        lib = L.load()
        h = C.c_void_p()
        rc = lib.kgma_genome_from_fasta(os.fsencode(path), C.byref(h))
        if rc != 0:
            raise KmerGMAError(rc, f"cannot ingest {path}" + (" (symbol outside IUPAC DNA)" if rc == L.E_SYMBOL else ""))
        return cls(h, lib)

    @classmethod
    def from_records(cls, records: Sequence[Union[FastaRecord, Tuple[str, str]]]) -> "Genome":
        lib = L.load()
        h = C.c_void_p()
        lib.kgma_genome_create(C.byref(h))
        g = cls(h, lib)
        for r in records:
            desc, seq = (r.description, r.sequence) if isinstance(r, FastaRecord) else r
            ident = desc.split(None, 1)[0] if desc.strip() else ""
            s = seq.encode() if isinstance(seq, str) else bytes(seq)
            rc = lib.kgma_genome_append_ascii(h, ident.encode(), desc.encode(), s, len(s))
            if rc != 0:
                raise KmerGMAError(rc, "invalid sequence character")
        rc = lib.kgma_genome_seal(h)
        if rc != 0:
            raise KmerGMAError(rc, "seal failed")
        return g

    @classmethod
    def synth(cls, rec_len: Sequence[int], seed: int = 42, n_run_len: int = 0, centromere_len: int = 0,
              ctx: Optional[Context] = None) -> "Genome":
        ctx = ctx or default_context()
        lens = np.asarray(rec_len, dtype=np.int64)
        h = C.c_void_p()
        ctx.check(ctx._lib.kgma_genome_synth(ctx._h, lens.size, lens.ctypes.data, seed, n_run_len, centromere_len, C.byref(h)))
        return cls(h, ctx._lib)

    def subset(self, records: Sequence[int], ctx: Optional[Context] = None) -> "Genome":
        """a genome of its own holding copies of `records` (in that order) in page-locked planes: what one rank of a
        contig-partitioned multi-GPU scan works on (north_star: "partitioned ... by contig").  Records start at multiples of
        128 bases in both genomes, so the packed words are copied as they are."""
        ctx = ctx or default_context()
        recs = np.ascontiguousarray([int(r) for r in records], dtype=np.int32)
        h = C.c_void_p()
        ctx.check(ctx._lib.kgma_genome_subset(ctx._h, self._h, recs.ctypes.data, recs.size, C.byref(h)))
        return Genome(h, ctx._lib)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.kgma_genome_destroy(self._h)
            self._h = None

    def __len__(self):
        return self._lib.kgma_genome_n_records(self._h)

    @property
    def total_len(self) -> int:
        return self._lib.kgma_genome_total_len(self._h)

    def seqsize(self, r: int) -> int:
        return self._lib.kgma_genome_record_len(self._h, r)

    def identifier(self, r: int) -> str:
        return self._lib.kgma_genome_identifier(self._h, r).decode()

    def description(self, r: int) -> str:
        return self._lib.kgma_genome_description(self._h, r).decode()

    def seq(self, r: int, first: int = 1, last: Optional[int] = None) -> str:
        """view(seq, first:last), 1-based inclusive."""
        if last is None:
            last = self.seqsize(r)
        n = max(0, last - first + 1)
        buf = C.create_string_buffer(n + 1)
        rc = self._lib.kgma_genome_get_seq(self._h, r, first, last, buf)
        if rc != 0:
            raise KmerGMAError(rc, f"range {first}:{last} outside record {r}")
        return buf.value.decode()

    def seq_array(self, r: int, first: int = 1, last: Optional[int] = None) -> np.ndarray:
        """view(seq, first:last) as a uint8 array of upper-case residues (no Python string: records can be 250 Mb)"""
        if last is None:
            last = self.seqsize(r)
        n = max(0, last - first + 1)
        buf = np.empty(n + 1, dtype=np.uint8)
        rc = self._lib.kgma_genome_get_seq(self._h, r, first, last, buf.ctypes.data_as(C.c_char_p))
        if rc != 0:
            raise KmerGMAError(rc, f"range {first}:{last} outside record {r}")
        return buf[:n]

    def record_offset(self, r: int) -> int:
        return self._lib.kgma_genome_record_offset(self._h, r)

    def masked_runs(self) -> np.ndarray:
        """maximal runs of masked (non-ACGT) residues, [start, end) rows in packed coordinates (see record_offset)"""
        p = C.POINTER(C.c_int64)()
        n = C.c_int64()
        rc = self._lib.kgma_genome_masked_runs(self._h, C.byref(p), C.byref(n))
        if rc != 0:
            raise KmerGMAError(rc, "masked_runs failed")
        out = np.ctypeslib.as_array(p, shape=(max(1, 2 * n.value),))[:2 * n.value].copy().reshape(-1, 2)
        self._lib.kgma_free(p)
        return out

    def put_seq(self, r: int, first: int, seq: str):
        s = seq.encode()
        rc = self._lib.kgma_genome_put_seq(self._h, r, first, s, len(s))
        if rc != 0:
            raise KmerGMAError(rc, "put_seq failed")

    def make_resident(self, ctx: Optional[Context] = None):
        ctx = ctx or default_context()
        ctx.check(ctx._lib.kgma_genome_make_resident(ctx._h, self._h))
        self._resident_ctx = ctx           # a hint only: the library re-uploads when its device copy is not this genome's


def _as_genome(genome) -> Genome:
    if isinstance(genome, Genome):
        return genome
    if isinstance(genome, (str, os.PathLike)):
        return Genome.from_fasta(os.fspath(genome))
    if isinstance(genome, (list, tuple)):
        return Genome.from_records(genome)
    raise TypeError("genome must be a FASTA path, a Genome or a list of FastaRecord")


# ------------------------------------------------------------------------------------------------
class KFV(np.ndarray):
    """A Float64 k-mer frequency vector that remembers the integer sums it came from
    (RV = S * (1/N), src/ReferenceGeneration.jl:35) so the device path stays exact."""

    def __new__(cls, values, S=None, n_refs=None):
        obj = np.asarray(values, dtype=np.float64).view(cls)
        obj.S = None if S is None else np.ascontiguousarray(S, dtype=np.int32)
        obj.n_refs = n_refs
        return obj

    def __array_finalize__(self, obj):
        self.S = getattr(obj, "S", None)
        self.n_refs = getattr(obj, "n_refs", None)


def _ints_of(refVec, lib) -> Tuple[np.ndarray, int]:
    """(S, N) behind a refVec: carried by KFV, otherwise recovered with kgma_profile_from_kfv."""
    S = getattr(refVec, "S", None)
    N = getattr(refVec, "n_refs", None)
    if S is not None and N and np.asarray(refVec).shape == S.shape:
        return np.ascontiguousarray(S, dtype=np.int32), int(N)
    v = np.ascontiguousarray(refVec, dtype=np.float64)
    S = np.zeros(v.size, dtype=np.int32)
    n = C.c_int32()
    rc = lib.kgma_profile_from_kfv(v.ctypes.data, v.size, 0, S.ctypes.data, C.byref(n))
    if rc != 0:
        raise KmerGMAError(rc, "refVec is not (k-mer count sums) / (family size): the exact device path needs a rational profile")
    return S, n.value


class _Refs:
    def __init__(self, src):
        self._lib = L.load()
        h = C.c_void_p()
        if isinstance(src, (str, os.PathLike)):
            rc = self._lib.kgma_refs_from_fasta(os.fsencode(src), C.byref(h))
            if rc != 0:
                raise KmerGMAError(rc, f"cannot read references from {src}")
        else:
            self._lib.kgma_refs_create(C.byref(h))
            for r in src:
                s = (r.sequence if isinstance(r, FastaRecord) else r).encode()
                rc = self._lib.kgma_refs_append_ascii(h, s, len(s))
                if rc != 0:
                    raise KmerGMAError(rc, "KeyError: reference symbol outside A,C,G,T,N")
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.kgma_refs_destroy(self._h)
            self._h = None


def gen_ref_ws_cons(reference_seqs, k: int, get_maxlen: bool = False):
    """src/ReferenceGeneration.jl:4-41 -> (RV, windowsize, consensus[, maxlen]); RV is a KFV."""
    refs = _Refs(reference_seqs)
    lib = refs._lib
    S = np.zeros(4 ** k, dtype=np.int32)
    n, ws = C.c_int32(), C.c_int64()
    maxlen = lib.kgma_refs_maxlen(refs._h)
    cons = C.create_string_buffer(maxlen + 2)
    rc = lib.kgma_refs_profile(refs._h, k, S.ctypes.data, C.byref(n), C.byref(ws), cons)
    if rc != 0:
        raise KmerGMAError(rc, "gen_ref_ws_cons failed")
    rv = KFV(S.astype(np.float64) * (1.0 / n.value), S, n.value)      # answer .* (1/len)
    if get_maxlen:
        return rv, ws.value, cons.value.decode(), maxlen
    return rv, ws.value, cons.value.decode()


def get_cluster_index(inp, cutoffs) -> int:
    """src/ReferenceGeneration.jl:50-57"""
    answer = 1
    for num in cutoffs:
        if inp <= num:
            return answer
        answer += 1
    return answer


def cluster_ref_API(reference_seqs, k: int, cutoffs=(7, 12, 20, 25), get_dists: bool = False, include_avg: bool = True):
    """src/ReferenceGeneration.jl:75-138 -> (KFVs, windowsizes, consensus_vec, invalid_vec[, dists])"""
    refs = _Refs(reference_seqs)
    lib = refs._lib
    cut = np.asarray(cutoffs, dtype=np.float64)
    maxc = cut.size + 2
    nb = 4 ** k
    S = np.zeros((maxc, nb), dtype=np.int32)
    members = np.zeros(maxc, dtype=np.int32)
    wss = np.zeros(maxc, dtype=np.int64)
    stride = lib.kgma_refs_maxlen(refs._h) + 2
    cons = C.create_string_buffer(maxc * stride)
    invalid = np.zeros(maxc, dtype=np.int32)
    dists = np.zeros(lib.kgma_refs_count(refs._h), dtype=np.float64)
    n = lib.kgma_refs_cluster(refs._h, k, cut.ctypes.data, cut.size, int(include_avg), 0, S.ctypes.data,
                              members.ctypes.data, wss.ctypes.data, cons, stride, invalid.ctypes.data, dists.ctypes.data)
    if n < 0:
        raise KmerGMAError(n, "cluster_ref_API failed")
    raw = cons.raw
    kfvs, cv = [], []
    ncl = cut.size + 1
    for i in range(n):
        m = int(members[i])
        if m == 0:
            kfvs.append(KFV(np.zeros(nb), S[i].copy(), 0))
        elif include_avg and i == n - 1:
            kfvs.append(KFV(S[i].astype(np.float64) * (1.0 / m), S[i].copy(), m))      # average_KFV (:129)
        else:
            kfvs.append(KFV(S[i].astype(np.float64) / m, S[i].copy(), m))               # KFVs[i] ./= lens[i] (:118)
        cv.append(raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode())
    inv = [bool(x) for x in invalid[:ncl + (1 if include_avg else 0)]]
    out = (kfvs, [int(x) for x in wss[:n]], cv, inv)
    return out + (dists,) if get_dists else out


def eliminate_null_params(KFVs, windowsizes, consensus_vec, invalid_vec):
    """src/ReferenceGeneration.jl:152-168"""
    keep = [i for i, inv in enumerate(invalid_vec) if not inv]
    return [KFVs[i] for i in keep], [windowsizes[i] for i in keep], [consensus_vec[i] for i in keep]


def _kmer_count_np(codes: np.ndarray, k: int) -> np.ndarray:
    n = codes.size - k + 1
    idx = np.zeros(n, dtype=np.int64)
    for j in range(k):
        idx = idx * 4 + codes[j:j + n]
    return np.bincount(idx, minlength=4 ** k).astype(np.float64)


_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("N", 3)):       # src/Consts.jl:22-28 (N -> 3)
    _CODE[ord(_c)] = _CODE[ord(_c.lower())] = _v


def _codes(seq) -> np.ndarray:
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    c = _CODE[np.frombuffer(b, dtype=np.uint8)]
    if c.size and c.max() == 255:
        raise KeyError("symbol outside A,C,G,T,N")                          # the reference's Dict lookup fails the same way
    return c


def kmer_count(seq, k: int) -> np.ndarray:
    """kmer_count (src/Kmers.jl:14-28): Float64 counts of the 4^k k-mers, first base most significant."""
    c = _codes(seq)
    return _kmer_count_np(c.astype(np.int64), k) if c.size >= k else np.zeros(4 ** k)


def kmer_dist(seq1, seq2, k: int) -> float:
    """kmer_dist (src/Kmers.jl:54-60): (1/2k) * sqeuclidean(kmer_count(seq1), kmer_count(seq2) or a KFV)."""
    a = kmer_count(seq1, k)
    b = kmer_count(seq2, k) if isinstance(seq2, (str, bytes)) else np.asarray(seq2, dtype=np.float64)
    return (1.0 / (2 * k)) * float(np.sum((a - b) ** 2))


def as_UInt(seq) -> int:
    """as_UInt (src/Kmers.jl:101): the 2-bit integer of a k-mer, first base most significant."""
    v = 0
    for c in _codes(seq):
        v = (v << 2) | int(c)
    return v


def as_kmer(value: int, length: int) -> str:
    """as_kmer (src/Kmers.jl:80): inverse of as_UInt."""
    return "".join("ACGT"[(value >> (2 * (length - 1 - i))) & 3] for i in range(length))


# ---- randstrobe utilities of the experimental strobemer path (src/StrobemerGMA/Strobemers.jl).  Host-side mirrors of the
# functions the reference's tests pin; the strobemer scan itself (StrobeGMA!) is not part of the hot path and not built.
def randstrobe_score(s1, s2, q: int) -> int:
    """randstrobe_score (Strobemers.jl:12-14)"""
    return (as_UInt(s1) + as_UInt(s2)) % q


def get_strobe_2_mer(seq, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5, withGap: bool = True) -> str:
    """get_strobe_2_mer (Strobemers.jl:45-65): first s-mer + a second s-mer from starts w_min..w_max.  The reference seeds
    its running minimum with `2 << 63`, which is 0 in Int64, so the second strobe is the LAST start whose score is 0, else
    w_min -- the goldens (test-StrobemerGMA.jl:6-7) pin exactly that."""
    seq = str(seq).upper()
    first = seq[:s]
    min_score, min_ind = 0, w_min
    for i in range(w_min, w_max + 1):                     # 1-based window starts, as in the reference
        cur = randstrobe_score(first, seq[i - 1:i - 1 + s], q)
        if cur <= min_score:
            min_score, min_ind = cur, i
    second = seq[min_ind - 1:min_ind - 1 + s]
    if not withGap:
        return first + second
    return first + "-" * (min_ind - s - 1) + second + "-" * (len(seq) - min_ind - s + 1)


def ungapped_strobe_2_mer_count(seq, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5) -> np.ndarray:
    """ungapped_strobe_2_mer_count (Strobemers.jl:90-103): 4^(2s) bins of the gap-free 2-strobemers of every k = w_max+s-1 window."""
    seq = str(seq).upper()
    k = w_max + s - 1
    bins = np.zeros(4 ** (2 * s))
    for i in range(len(seq) - k + 1):
        bins[as_UInt(get_strobe_2_mer(seq[i:i + k], s, w_min, w_max, q, withGap=False))] += 1
    return bins


def estimate_optimal_threshold(RV, average_length, seed: int = 42, num_trials: int = 100, buffer: float = 8):
    """src/DistanceTesting.jl:8-32: mean kmer_dist of `num_trials` random sequences to RV, minus `buffer`.
    The reference draws them with Julia's RNG after Random.seed!(42); that stream is not reproducible
    outside Julia, so this uses numpy's PCG64 with the same seed (values agree to within sampling noise,
    see DESIGN.md).  One generator is consumed sequentially across clusters like the reference's."""
    rng = np.random.default_rng(seed)

    def one(rv, length):
        rv = np.asarray(rv, dtype=np.float64)
        k = int(round(np.log(rv.size) / np.log(4)))
        tot = 0.0
        for _ in range(num_trials):
            c = _kmer_count_np(rng.integers(0, 4, size=length), k)
            tot += (1.0 / (2 * k)) * float(np.sum((c - rv) ** 2))
        return tot / num_trials - buffer

    if isinstance(average_length, (list, tuple, np.ndarray)):
        return [one(rv, int(l)) for rv, l in zip(RV, average_length)]
    return one(RV, int(average_length))


# ------------------------------------------------------------------------------------------------
def julia_round2(x: float) -> float:
    """round(x, digits = 2) as Base does it: round-half-even of x*100, then /100."""
    return float(np.rint(x * 100.0) / 100.0)


def julia_float_str(x: float) -> str:
    """string(::Float64): shortest round-trip decimal, always with a fractional part."""
    return repr(float(x))


def _header(ident: str, h, cluster: bool, with_genome_pos: bool = True) -> str:
    d = julia_float_str(julia_round2(float(h.dist)))
    rng = f"{h.first}:{h.last}"
    ln = h.last - h.first + 1
    if cluster:            # src/OmnGenomeMiner.jl:141-149
        return f"{ident} | Dist = {d} | KFV = {h.profile} | MatchPos = {rng} | GenomePos = {h.genome_pos} | Len = {ln}"
    gp = f" | GenomePos = {h.genome_pos}" if with_genome_pos else ""
    return f"{ident} | dist = {d} | MatchPos = {rng}{gp} | Len = {ln}"      # src/Alignment.jl:71-78


HIT_DT = np.dtype([("record", "<i4"), ("profile", "<i4"), ("cmi", "<i8"), ("first", "<i8"), ("last", "<i8"),
                   ("genome_pos", "<i8"), ("D", "<i8"), ("dist", "<f8"), ("align_score", "<i8"), ("flags", "<u4"),
                   ("cigar_off", "<u4"), ("cigar_len", "<u4"), ("reserved", "<u4")])
RUN_DT = np.dtype([("record", "<i4"), ("profile", "<i4"), ("t_first", "<i8"), ("t_last", "<i8"), ("t_argmin", "<i8"),
                   ("D_min", "<i8"), ("flags", "<u4"), ("reserved", "<u4")])
assert HIT_DT.itemsize == C.sizeof(L.Hit) and RUN_DT.itemsize == C.sizeof(L.Run)


class ScanOutput:
    """Raw result of one kgma_scan call: `hits` is a numpy record array over kgma_hit (h.record, h.first, ...; the
    kgma_hit.flags field must be read as h["flags"], `.flags` being numpy's own attribute),
    `runs` the raw bytes of the kgma_run list (view with RUN_DT), plus optional dists / cigars."""

    def __init__(self, lib, res_handle):
        n = lib.kgma_result_n_hits(res_handle)
        hp = lib.kgma_result_hits(res_handle)
        if n:
            raw = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(n * HIT_DT.itemsize,)).copy()
            self.hits = raw.view(HIT_DT).view(np.recarray)
        else:
            self.hits = np.zeros(0, dtype=HIT_DT).view(np.recarray)
        nr = lib.kgma_result_n_runs(res_handle)
        rp = lib.kgma_result_runs(res_handle)
        self.runs = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint8)), shape=(nr * C.sizeof(L.Run),)).copy() if nr else np.zeros(0, np.uint8)
        self.n_runs = nr
        self.first_D = None
        self.dists: List[np.ndarray] = []
        self._cigars = None
        self._lib, self._res = lib, res_handle

    @property
    def cigars(self) -> List[Optional[AlignResult]]:
        if self._cigars is None:
            ops, cnt = self._lib.kgma_result_cigar_ops(self._res), self._lib.kgma_result_cigar_counts(self._res)
            out = []
            for h in self.hits:
                if h.cigar_len and ops:
                    s = "".join(f"{cnt[int(h.cigar_off) + t]}{ops[int(h.cigar_off) + t].decode()}" for t in range(int(h.cigar_len)))
                    out.append(AlignResult(s, int(h.align_score)))
                else:
                    out.append(None)
            self._cigars = out
        return self._cigars

    @property
    def align_events(self) -> List[tuple]:
        """kgma_result_align_events: (record, profile, cmi, AlignResult, emitted) for every extension of a cluster-mode scan with
        KGMA_F_WANT_CIGARS, in the reference's order (OmnGenomeMiner.jl:131-133: pushed before the second overlap test)"""
        n = self._lib.kgma_result_n_align_events(self._res)
        ev = self._lib.kgma_result_align_events(self._res)
        ops, cnt = self._lib.kgma_result_cigar_ops(self._res), self._lib.kgma_result_cigar_counts(self._res)
        out = []
        for i in range(n):
            e = ev[i]
            s = "".join(f"{cnt[e.cigar_off + t]}{ops[e.cigar_off + t].decode()}" for t in range(e.cigar_len)) if ops else ""
            out.append((int(e.record), int(e.profile), int(e.cmi), AlignResult(s, int(e.align_score)), bool(e.emitted)))
        return out

    def load_dists(self, n_profiles: int):
        for q in range(n_profiles):
            n = self._lib.kgma_result_n_dists(self._res, q)
            p = self._lib.kgma_result_dists(self._res, q)
            self.dists.append(np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0))

    def load_run_ext(self) -> np.ndarray:
        """kgma_result_run_ext as an (n_runs, 3) int64 array [lo, hi, score] (kgma_scan_shard results; lo == 0: not extended)"""
        p = self._lib.kgma_result_run_ext(self._res)
        if not p or not self.n_runs:
            return np.zeros((0, 3), np.int64)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(self.n_runs, 3)).copy()

    def load_first_D(self, n_profiles: int, n_records: int):
        p = self._lib.kgma_result_first_D(self._res)
        self.first_D = np.ctypeslib.as_array(p, shape=(n_profiles * n_records,)).copy() if p else np.zeros(0, np.int64)

    def free(self):
        if self._res:
            self._lib.kgma_result_free(self._res)
            self._res = None

    def __del__(self):
        self.free()


def _make_profiles(refVecs, windowsizes, consensus_seqs, thrs, k, lib):
    keep = []
    arr = (L.Profile * len(refVecs))()
    for i, rv in enumerate(refVecs):
        if np.asarray(rv).size != 4 ** k:
            raise KmerGMAError(L.E_ARG, f"refVec length {np.asarray(rv).size} != 4^k")
        S, N = _ints_of(rv, lib)
        cons = consensus_seqs[i].upper().encode() if consensus_seqs[i] is not None else b""
        keep.append((S, cons))
        arr[i].k, arr[i].n_refs, arr[i].window = k, N, int(windowsizes[i])
        arr[i].S = S.ctypes.data_as(C.POINTER(C.c_int32))
        arr[i].consensus, arr[i].consensus_len = cons, len(cons)
        arr[i].thr = float(thrs[i])
    return arr, keep


def scan_raw(genome: Genome, refVecs, windowsizes, consensus_seqs, thrs, k: int, mode: int, buff: int,
             flags: int, gap_open: int, gap_extend: int, only_record: int = -1,
             ctx: Optional[Context] = None, runs_only: bool = False, shard: Tuple[int, int] = (0, 1)) -> ScanOutput:
    """Thin call into kgma_scan / kgma_scan_runs."""
    ctx = ctx or default_context()
    arr, keep = _make_profiles(refVecs, windowsizes, consensus_seqs, thrs, k, ctx._lib)
    P = L.ScanParams(mode, flags, buff, gap_open, gap_extend, shard[0], shard[1], only_record, 0)
    res = C.c_void_p()
    fn = ctx._lib.kgma_scan_runs if runs_only else ctx._lib.kgma_scan
    ctx.check(fn(ctx._h, genome._h, arr, len(refVecs), C.byref(P), C.byref(res)))
    out = ScanOutput(ctx._lib, res)
    if flags & L.F_WANT_DISTS:
        out.load_dists(len(refVecs))
    if runs_only:
        out.load_first_D(len(refVecs), len(genome))
    return out


class PreparedScan:
    """The arguments of one operator call, marshalled once: profiles (integer sums, consensus bytes), parameters and the
    genome handle.  A serving loop (bench.py's timed region) then pays only the C-ABI call itself per scan -- building the
    ctypes structures costs as much as a resident scan's host work otherwise.  `scan()` / `scan_shard()` / `replay_packed()`
    return the raw kgma_result handle wrapped in a LightResult (hit count and stats without copying the hit list)."""

    def __init__(self, genome: Genome, refVecs, windowsizes, consensus_seqs, thrs, k: int, mode: int, buff: int,
                 gap_open: int, gap_extend: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx._lib
        self.genome = genome
        self.n = len(refVecs)
        self.arr, self._keep = _make_profiles(refVecs, windowsizes, consensus_seqs, thrs, k, self.lib)
        self._args = (mode, buff, gap_open, gap_extend)

    def _params(self, flags: int, shard=(0, 1)):
        mode, buff, go, ge = self._args
        return L.ScanParams(mode, flags, buff, go, ge, shard[0], shard[1], -1, 0)

    def scan(self, flags: int) -> "LightResult":
        P = self._params(flags)
        res = C.c_void_p()
        self.ctx.check(self.lib.kgma_scan(self.ctx._h, self.genome._h, self.arr, self.n, C.byref(P), C.byref(res)))
        return LightResult(self.lib, res)

    def scan_shard(self, flags: int, shard: Tuple[int, int]) -> "LightResult":
        P = self._params(flags, shard)
        res = C.c_void_p()
        self.ctx.check(self.lib.kgma_scan_shard(self.ctx._h, self.genome._h, self.arr, self.n, C.byref(P), C.byref(res)))
        return LightResult(self.lib, res)

    def replay_packed(self, flags: int, blocks_ptr: int, n_blocks: int, stride: int) -> "LightResult":
        P = self._params(flags)
        res = C.c_void_p()
        rc = self.lib.kgma_replay_packed(self.ctx._h, self.genome._h, self.arr, self.n, C.byref(P), blocks_ptr, n_blocks, stride, C.byref(res))
        self.ctx.check(rc)
        return LightResult(self.lib, res)


class LightResult:
    """a kgma_result handle: counts and packing without materialising numpy copies; `.full()` converts to a ScanOutput"""

    def __init__(self, lib, res):
        self._lib, self._res = lib, res

    @property
    def n_hits(self) -> int:
        return int(self._lib.kgma_result_n_hits(self._res))

    def pack(self, buf_ptr, cap: int) -> int:
        return int(self._lib.kgma_result_pack(self._res, buf_ptr, cap))

    def copy_hits(self, buf_ptr, cap: int) -> int:
        """[int64 n][kgma_hit x n] into caller memory (a rank's block of a contig-partitioned scan); returns the bytes needed"""
        n = self.n_hits
        need = 16 + n * C.sizeof(L.Hit)
        if buf_ptr and cap >= need:
            C.memmove(buf_ptr, C.byref(C.c_int64(n)), 8)
            if n:
                C.memmove(buf_ptr + 16, self._lib.kgma_result_hits(self._res), n * C.sizeof(L.Hit))
        return need

    def full(self) -> "ScanOutput":
        out = ScanOutput(self._lib, self._res)
        self._res = None
        return out

    def free(self):
        if self._res:
            self._lib.kgma_result_free(self._res)
            self._res = None

    def __del__(self):
        self.free()


def scan_shard_raw(genome: Genome, refVecs, windowsizes, consensus_seqs, thrs, k: int, mode: int, buff: int,
                   flags: int, gap_open: int, gap_extend: int, shard: Tuple[int, int], ctx: Optional[Context] = None) -> ScanOutput:
    """kgma_scan_shard: one rank's share of a multi-GPU scan -- its shard's merged runs plus, with F_ALIGN, the extension
    result of every run's own candidate window (computed on this GPU).  Ship `pack_shard(out, ...)` blocks between ranks and
    hand all of them to `replay_packed`."""
    ctx = ctx or default_context()
    arr, keep = _make_profiles(refVecs, windowsizes, consensus_seqs, thrs, k, ctx._lib)
    P = L.ScanParams(mode, flags, buff, gap_open, gap_extend, shard[0], shard[1], -1, 0)
    res = C.c_void_p()
    ctx.check(ctx._lib.kgma_scan_shard(ctx._h, genome._h, arr, len(refVecs), C.byref(P), C.byref(res)))
    out = ScanOutput(ctx._lib, res)
    out.load_first_D(len(refVecs), len(genome))
    return out


def pack_shard(out: ScanOutput, buf_ptr: int, cap: int) -> int:
    """kgma_result_pack into caller memory (e.g. a pinned torch tensor's data_ptr); returns the bytes needed."""
    return int(out._lib.kgma_result_pack(out._res, buf_ptr, cap))


def replay_packed(genome: Genome, refVecs, windowsizes, consensus_seqs, thrs, k: int, mode: int, buff: int,
                  flags: int, gap_open: int, gap_extend: int, blocks_ptr: int, n_blocks: int, stride: int,
                  ctx: Optional[Context] = None) -> ScanOutput:
    """kgma_replay_packed: host-only merge + replay of the shards' blocks (ctx is only used for error text / stats)."""
    lib = L.load()
    arr, keep = _make_profiles(refVecs, windowsizes, consensus_seqs, thrs, k, lib)
    P = L.ScanParams(mode, flags, buff, gap_open, gap_extend, 0, 1, -1, 0)
    res = C.c_void_p()
    handle = ctx._h if ctx is not None else None
    rc = lib.kgma_replay_packed(handle, genome._h, arr, len(refVecs), C.byref(P), blocks_ptr, n_blocks, stride, C.byref(res))
    if rc != 0:
        raise KmerGMAError(rc, (lib.kgma_last_error(handle) or b"").decode() or "kgma_replay_packed failed")
    return ScanOutput(lib, res)


def exact_match_shard(query: str, genome: Genome, shard: Tuple[int, int], ctx: Optional[Context] = None, resident: bool = False) -> np.ndarray:
    """kgma_exact_match_shard: occurrence starts (0-based packed coordinates) owned by slice shard[0] of shard[1]."""
    ctx = ctx or default_context()
    q = str(query).upper().encode()
    sp = C.POINTER(C.c_int64)()
    n = C.c_int64()
    ctx.check(ctx._lib.kgma_exact_match_shard(ctx._h, genome._h, q, len(q), L.F_RESIDENT if resident else 0, shard[0], shard[1],
                                              C.byref(sp), C.byref(n)))
    out = np.ctypeslib.as_array(sp, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int64)
    if n.value:
        ctx._lib.kgma_free(sp)
    return out


def exact_match_merge(genome: Genome, starts: np.ndarray, qlen: int, overlap: bool = True):
    """kgma_exact_match_merge (host only): concatenated slice starts -> what exactMatch(query, reader) returns."""
    lib = L.load()
    st = np.ascontiguousarray(starts, dtype=np.int64)
    mp = C.POINTER(L.Match)()
    n = C.c_int64()
    rc = lib.kgma_exact_match_merge(genome._h, st.ctypes.data, st.size, qlen, int(overlap), C.byref(mp), C.byref(n))
    if rc != 0:
        raise KmerGMAError(rc, "kgma_exact_match_merge failed")
    by_rec: Dict[int, List[Tuple[int, int]]] = {}
    for i in range(n.value):
        by_rec.setdefault(mp[i].record, []).append((mp[i].first, mp[i].last))
    if n.value:
        lib.kgma_free(mp)
    identify = {genome.identifier(r): by_rec[r] for r in sorted(by_rec)}
    return identify if identify else "no match"


def replay_raw(genome: Genome, refVecs, windowsizes, consensus_seqs, thrs, k: int, mode: int, buff: int,
               flags: int, gap_open: int, gap_extend: int, runs: np.ndarray, first_D: np.ndarray,
               only_record: int = -1, ctx: Optional[Context] = None, host_only: bool = False) -> ScanOutput:
    """kgma_replay over the concatenated run summaries of all shards (runs: raw bytes of kgma_run[]).
    host_only: merge + replay without a device context (only valid without F_ALIGN; the extension needs the GPU)."""
    lib = L.load()
    handle = None
    if not host_only:
        ctx = ctx or default_context()
        handle = ctx._h
    arr, keep = _make_profiles(refVecs, windowsizes, consensus_seqs, thrs, k, lib)
    P = L.ScanParams(mode, flags, buff, gap_open, gap_extend, 0, 1, only_record, 0)
    runs = np.ascontiguousarray(runs, dtype=np.uint8)
    fd = np.ascontiguousarray(first_D, dtype=np.int64)
    res = C.c_void_p()
    rc = lib.kgma_replay(handle, genome._h, arr, len(refVecs), C.byref(P), runs.ctypes.data,
                         runs.size // C.sizeof(L.Run), fd.ctypes.data, C.byref(res))
    if rc != 0:
        raise KmerGMAError(rc, (lib.kgma_last_error(handle) or b"").decode() or "kgma_replay failed")
    return ScanOutput(lib, res)


def _emit(genome: Genome, out: ScanOutput, cluster: bool, resultVec, hit_loci_vec, align_vec, with_genome_pos=True):
    cigars = out.cigars if (align_vec is not None and not cluster) else [None] * len(out.hits)
    if cluster and align_vec is not None:          # one alignment per extension performed, emitted or not (OmnGenomeMiner.jl:133)
        align_vec.extend(e[3] for e in out.align_events)
    for h, cg in zip(out.hits, cigars):
        rec, first, last = int(h.record), int(h.first), int(h.last)
        ident = genome.identifier(rec)
        seq = genome.seq(rec, first, last) if last >= first else ""
        resultVec.append(FastaRecord(_header(ident, h, cluster, with_genome_pos), seq))
        if hit_loci_vec is not None:
            hit_loci_vec.append(first + int(h.genome_pos))
        if align_vec is not None and cg is not None:
            align_vec.append(cg)


def _check_fixed_params(k: int, ScaleFactor, mask, Nt_bits):
    """The device path computes d = D / (2 k N^2), i.e. ScaleFactor = 1/k, with the k-mer mask 4^k - 1 and the standard
    NUCLEOTIDE_BITS table -- what findGenes / findGenes_cluster_mode always pass (API.jl:86-87,204-205).  The reference
    operators' own DEFAULTS are ScaleFactor = 1/6 and mask = 4095 whatever k is (GenomeMiner.jl:12-14): a direct operator call
    with k != 6 that relies on those defaults gets the findGenes scaling here, and anything explicitly different is refused."""
    if ScaleFactor is not None and abs(float(ScaleFactor) - 1.0 / k) > 1e-15:
        raise KmerGMAError(L.E_UNSUPPORTED, f"ScaleFactor = {ScaleFactor}: the device path is fixed to 1/k = {1.0 / k} (rescale thr by ScaleFactor*k instead)")
    if mask is not None and int(mask) != 4 ** k - 1:
        raise KmerGMAError(L.E_UNSUPPORTED, f"mask = {mask}: the device path is fixed to 4^k - 1 = {4 ** k - 1}")
    if Nt_bits is not None:
        std = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 3}
        got = {str(a).upper(): int(b) for a, b in dict(Nt_bits).items()}
        if got != std:
            raise KmerGMAError(L.E_UNSUPPORTED, "Nt_bits differs from NUCLEOTIDE_BITS (A0 C1 G2 T3, N -> 3): the packed genome is fixed to it")


def ac_gma_testing(*, genome_path, refVec, consensus_refseq: str, k: int = 6, windowsize: int = 289,
                   thr: float = 33.5, buff: int = 50, mask=None, Nt_bits=None, ScaleFactor=None,
                   do_align: bool = True, result_align_vec: Optional[list] = None,
                   gap_open_score: int = -69, gap_extend_score: int = -1,
                   do_return_dists: bool = False, dist_vec: Optional[list] = None,
                   do_return_align: bool = False, get_hit_loci: bool = False,
                   hit_loci_vec: Optional[list] = None, resultVec: Optional[list] = None,
                   dense: bool = False, ctx: Optional[Context] = None):
    """ac_gma_testing! (src/GenomeMiner.jl:4-109): single-profile scan; appends to resultVec /
    hit_loci_vec / result_align_vec / dist_vec in place and returns the raw ScanOutput.
    `mask`, `Nt_bits`, `ScaleFactor` are accepted for signature parity; values other than 4^k-1 / the standard table / 1/k are
    refused (see _check_fixed_params)."""
    _check_fixed_params(k, ScaleFactor, mask, Nt_bits)
    resultVec = [] if resultVec is None else resultVec
    ctx = ctx or default_context()
    g = _as_genome(genome_path)
    flags = (L.F_ALIGN if do_align else 0) | (L.F_WANT_DISTS if do_return_dists else 0) | \
            (L.F_WANT_CIGARS if (do_align and do_return_align) else 0) | (L.F_DENSE if dense else 0) | \
            (L.F_RESIDENT if g._resident_ctx is ctx else 0)
    out = scan_raw(g, [refVec], [windowsize], [consensus_refseq], [thr], k, L.MODE_SINGLE, buff, flags,
                   gap_open_score, gap_extend_score, ctx=ctx)
    _emit(g, out, False, resultVec, hit_loci_vec if get_hit_loci else None,
          result_align_vec if do_return_align else None)
    if do_return_dists and dist_vec is not None:
        dist_vec.extend(out.dists[0].tolist()) if isinstance(dist_vec, list) else None
    return out


def record_KmerGMA(*, record: Union[FastaRecord, Tuple[str, str]], refVec, consensus_refseq: str,
                   resultVec_vec: Optional[List[list]] = None, curr_kmer_freq_vec=None,
                   k: int = 6, windowsize: int = 289, thr: float = 30, buff: int = 50,
                   do_align: bool = True, gap_open_score: int = -69, gap_extend_score: int = -1,
                   ctx: Optional[Context] = None):
    """record_KmerGMA! (src/MultiThread/GenomeMiner.jl:8-98): the single-profile scan of one record; the
    header carries no GenomePos (:88-91).  Results are appended to resultVec_vec[0]."""
    resultVec_vec = [[]] if resultVec_vec is None else resultVec_vec
    g = Genome.from_records([record])
    flags = L.F_ALIGN if do_align else 0
    out = scan_raw(g, [refVec], [windowsize], [consensus_refseq], [thr], k, L.MODE_SINGLE, buff, flags,
                   gap_open_score, gap_extend_score, only_record=0, ctx=ctx)
    _emit(g, out, False, resultVec_vec[0], None, None, with_genome_pos=False)
    return out


def Omn_KmerGMA(*, genome_path, refVecs, windowsizes, consensus_seqs, resultVec: list, k: int = 6,
                ScaleFactor=None, mask=None, thr_vec=(35, 31, 38, 34, 27, 27), buff: int = 50, Nt_bits=None,
                align_hits: bool = True, align_vec: Optional[list] = None,
                gap_open_score: int = -200, gap_extend_score: int = -1, genome_pos: int = 0,
                get_hit_loci: bool = False, hit_loci_vec: Optional[list] = None, get_aligns: bool = False,
                do_return_dists: bool = False, dist_vec_vec: Optional[List[list]] = None,
                dense: bool = False, ctx: Optional[Context] = None):
    """Omn_KmerGMA! (src/OmnGenomeMiner.jl:7-162): C profiles scanned together.
    With get_aligns the reference pushes one alignment per extension it performs, including those whose hit the second overlap
    test (:139) then rejects (:133); align_vec receives exactly that list (kgma_result_align_events)."""
    _check_fixed_params(k, ScaleFactor, mask, Nt_bits)
    ctx = ctx or default_context()
    g = _as_genome(genome_path)
    Cn = len(windowsizes)
    flags = (L.F_ALIGN if align_hits else 0) | (L.F_WANT_DISTS if do_return_dists else 0) | \
            (L.F_WANT_CIGARS if (align_hits and get_aligns) else 0) | (L.F_DENSE if dense else 0) | \
            (L.F_RESIDENT if g._resident_ctx is ctx else 0)
    out = scan_raw(g, list(refVecs)[:Cn], windowsizes, list(consensus_seqs)[:Cn], list(thr_vec)[:Cn], k,
                   L.MODE_CLUSTER, buff, flags, gap_open_score, gap_extend_score, ctx=ctx)
    if genome_pos:
        out.hits.genome_pos += genome_pos
    _emit(g, out, True, resultVec, hit_loci_vec if get_hit_loci else None, align_vec if get_aligns else None)
    if do_return_dists and dist_vec_vec is not None:
        for q in range(min(Cn, len(dist_vec_vec))):
            dist_vec_vec[q].extend(out.dists[q].tolist())
    return out


# ------------------------------------------------------------------------------------------------
def warn_helper(k: int, do_return_dists: bool):
    """src/API.jl:8-11"""
    if k < 5:
        warnings.warn(f"Such a low k value of {k} likely won't yield the most accurate results")
    if do_return_dists:
        warnings.warn("Setting do_return_dists to true may be very memory intensive")


def _jl_num(x) -> str:
    return julia_float_str(x) if isinstance(x, float) else str(x)


def findGenes(*, genome_path, ref_path, k: int = 6, KmerDistThr: Union[int, float] = 0, buffer: int = 50,
              do_align: bool = True, gap_open_score: int = -69, gap_extend_score: int = -1,
              do_return_dists: bool = False, do_return_hit_loci: bool = False, do_return_align: bool = False,
              verbose: bool = True, KmerDist_threshold_buffer: float = 8.0, ctx: Optional[Context] = None) -> list:
    """KmerGMA.findGenes (src/API.jl:60-104).  Returns [hit_vector, (hit_loci), (alignments), (dists)]."""
    if verbose:
        _log.info("pre-processing references and parameters...")
    warn_helper(k, do_return_dists)
    RV, windowsize, consensus_refseq = gen_ref_ws_cons(ref_path, k)
    if k >= windowsize:
        raise KmerGMAError(L.E_WINDOW, f"the average reference sequence length {windowsize} exceeds/is equal to the chosen kmer length {k}. please reduce k. ")
    est = estimate_optimal_threshold(RV, windowsize, buffer=KmerDist_threshold_buffer)
    if KmerDistThr == 0:
        KmerDistThr = est
    elif KmerDistThr < est:     # (sic) src/API.jl:75-77
        warnings.warn(f"The kmer distance threshold {_jl_num(KmerDistThr)} for k = {k} is likely too high, and can result in many false positives")
    hit_vector, dist_vec, hit_loci_vec, alignment_vec = [], [], [], []
    if verbose:
        _log.info("initializing iteration...")
    ac_gma_testing(genome_path=genome_path, refVec=RV, consensus_refseq=consensus_refseq, k=k,
                   windowsize=windowsize, thr=KmerDistThr, buff=buffer, do_align=do_align,
                   gap_open_score=gap_open_score, gap_extend_score=gap_extend_score,
                   do_return_dists=do_return_dists, do_return_align=do_return_align,
                   get_hit_loci=do_return_hit_loci, dist_vec=dist_vec, result_align_vec=alignment_vec,
                   hit_loci_vec=hit_loci_vec, resultVec=hit_vector, ctx=ctx)
    info_str = "genome mining completed successfully, returning vector of: vector of hits"
    output_vector: list = [hit_vector]
    if do_return_hit_loci:
        output_vector.append(hit_loci_vec); info_str += ", vector of hit locations"
    if do_return_align:
        output_vector.append(alignment_vec); info_str += ", vector of alignments"
    if do_return_dists:
        output_vector.append(dist_vec); info_str += ", vector of kmer distances along the genome"
    if verbose:
        _log.info(info_str)
    return output_vector


def findGenes_cluster_mode(*, genome_path, ref_path, cluster_cutoffs=(7, 12, 20, 25), k: int = 6,
                           KmerDistThrs=(0.0,), buffer: int = 100, do_align: bool = True,
                           gap_open_score: int = -200, gap_extend_score: int = -1,
                           do_return_dists: bool = False, do_return_hit_loci: bool = False,
                           do_return_align: bool = False, verbose: bool = True,
                           kmerDist_threshold_buffer: float = 7, ctx: Optional[Context] = None) -> list:
    """KmerGMA.findGenes_cluster_mode (src/API.jl:161-226)."""
    if verbose:
        _log.info("pre-processing references and parameters...")
    warn_helper(k, do_return_dists)
    RVs, windowsizes, consensus_refseqs, invalids = cluster_ref_API(ref_path, k, cutoffs=cluster_cutoffs)
    RVs, windowsizes, consensus_refseqs = eliminate_null_params(RVs, windowsizes, consensus_refseqs, invalids)
    if k >= min(windowsizes):
        raise KmerGMAError(L.E_WINDOW, f"some/all of the average reference sequence lengths exceeds/is equal to the chosen kmer length {k}. please reduce k. ")
    est = estimate_optimal_threshold(RVs, windowsizes, buffer=kmerDist_threshold_buffer)
    KmerDistThrs = [float(x) for x in KmerDistThrs]
    if KmerDistThrs[0] == 0:
        KmerDistThrs = est
    else:
        ind_warn, num_warn = "", ""
        for i, num in enumerate(KmerDistThrs):
            if i < len(est) and num > est[i]:
                ind_warn += f"{i + 1}, "
                num_warn += f"{_jl_num(num)}, "
        if ind_warn:
            thr_txt = "[" + ", ".join(_jl_num(x) for x in KmerDistThrs) + "]"
            warnings.warn(f"The kmer distance thresholds {thr_txt} at index/indicies {ind_warn}"[:-2] +
                          f" for k = {k} is potentially too high, and may result in more false positives.")
    hit_vector, hit_loci_vec, alignment_vec = [], [], []
    dist_vec_vec = [[] for _ in windowsizes]
    if verbose:
        _log.info("initializing iteration...")
    Omn_KmerGMA(genome_path=genome_path, refVecs=RVs, windowsizes=windowsizes, consensus_seqs=consensus_refseqs,
                resultVec=hit_vector, k=k, thr_vec=KmerDistThrs, buff=buffer, align_hits=do_align,
                gap_open_score=gap_open_score, gap_extend_score=gap_extend_score, get_aligns=do_return_align,
                get_hit_loci=do_return_hit_loci, hit_loci_vec=hit_loci_vec, align_vec=alignment_vec,
                do_return_dists=do_return_dists, dist_vec_vec=dist_vec_vec, ctx=ctx)
    info_str = "genome mining completed successfully, returning vector of: vector of hits"
    output_vector: list = [hit_vector]
    if do_return_hit_loci:
        output_vector.append(hit_loci_vec); info_str += ", vector of hit locations"
    if do_return_align:
        output_vector.append(alignment_vec); info_str += ", vector of alignments"
    if do_return_dists:
        output_vector.append(dist_vec_vec); info_str += ", vector of vectors of kmer distances along the genome"
    if verbose:
        _log.info(info_str)
        _log.info("To write the results, use `KmerGMA.write_results`")
    return output_vector


# ------------------------------------------------------------------------------------------------
# Multi-GPU by contig (north_star: "the genome is partitioned across the GPUs by contig/chunk"): records are independent in both
# state machines and GenomePos is a running sum of record lengths, so a rank that holds whole records runs the ordinary kgma_scan
# on them -- its own replay, exactly the extensions the reference performs -- and ships finished hits.  Used when the records
# balance over the ranks; a genome of a few huge contigs goes by slices instead (kgma_scan_shard / kgma_replay_packed).
def partition_records(lens: Sequence[int], world: int, tolerance: float = 1.10) -> Optional[List[List[int]]]:
    """longest-processing-time assignment of records to `world` ranks (each list ascending); None when the fullest rank would
    hold more than `tolerance` times the even share"""
    lens = [int(x) for x in lens]
    load = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for r in sorted(range(len(lens)), key=lambda i: -lens[i]):
        j = min(range(world), key=lambda i: load[i])
        parts[j].append(r)
        load[j] += lens[r]
    if any(not p for p in parts) or max(load) > tolerance * sum(lens) / world:
        return None
    return [sorted(p) for p in parts]


class PartitionMerger:
    """rank 0's side of a contig-partitioned scan: the maps are built once, merge() is one C call (kgma_hits_merge_partition)"""

    def __init__(self, parts: List[List[int]], lens: Sequence[int], min_len: int = 0):
        """min_len: single mode skips records shorter than the window WITHOUT advancing GenomePos (GenomeMiner.jl:37-39,106) --
        pass the window size; cluster mode advances for every record (OmnGenomeMiner.jl:159) -- pass 0."""
        lens = np.asarray(lens, dtype=np.int64)
        self._lib = L.load()
        self.n_records = int(lens.size)
        self.glob_cum = np.ascontiguousarray(np.concatenate([[0], np.cumsum(np.where(lens >= min_len, lens, 0))])[:-1], dtype=np.int64)
        self.rec_map = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.int32) for p in parts]), dtype=np.int32)
        self.rec_off = np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(p) for p in parts])]), dtype=np.int32)
        self.world = len(parts)

    def merge(self, blocks_ptr: int, stride: int) -> np.ndarray:
        hp = C.POINTER(L.Hit)()
        n = C.c_int64()
        rc = self._lib.kgma_hits_merge_partition(blocks_ptr, self.world, stride, self.rec_map.ctypes.data, self.rec_off.ctypes.data,
                                                 self.n_records, self.glob_cum.ctypes.data, C.byref(hp), C.byref(n))
        if rc != 0:
            raise KmerGMAError(rc, "kgma_hits_merge_partition failed")
        out = np.frombuffer(C.string_at(hp, n.value * C.sizeof(L.Hit)), dtype=HIT_DT) if n.value else np.zeros(0, dtype=HIT_DT)
        self._lib.kgma_free(hp)
        return out


def merge_partition_hits(blocks: np.ndarray, parts: List[List[int]], lens: Sequence[int], min_len: int = 0) -> np.ndarray:
    """blocks[rank] = what LightResult.copy_hits wrote on that rank (records numbered within the rank's sub-genome) -> the hits of
    the whole genome in the reference's order: global record indices, GenomePos = summed length of the records in front."""
    blocks = np.ascontiguousarray(blocks)
    return PartitionMerger(parts, lens, min_len).merge(blocks.ctypes.data, blocks.shape[1])


# ------------------------------------------------------------------------------------------------
# Strobemer path (src/StrobemerGMA/, experimental in the reference; its scan has no test or golden there -- parity is against a
# line-by-line CPU restatement under tests/, assembled from the three utilities the reference's tests do pin).
def strobe_gen_ref_ws_cons(reference_seqs, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5):
    """gen_ref_ws_cons(refs; s, w_min, w_max, q) (StrobeRefGen.jl:4-42) -> (RV, windowsize, consensus); RV is a KFV over the
    4^(2s) gap-free 2-randstrobe codes."""
    refs = _Refs(reference_seqs)
    lib = refs._lib
    S = np.zeros(4 ** (2 * s), dtype=np.int32)
    n, ws = C.c_int32(), C.c_int64()
    cons = C.create_string_buffer(lib.kgma_refs_maxlen(refs._h) + 2)
    rc = lib.kgma_refs_strobe_profile(refs._h, s, w_min, w_max, q, S.ctypes.data, C.byref(n), C.byref(ws), cons)
    if rc != 0:
        raise KmerGMAError(rc, "gen_ref_ws_cons (strobemers) failed")
    return KFV(S.astype(np.float64) * (1.0 / n.value), S, n.value), ws.value, cons.value.decode()


def strobe_scan_raw(genome: Genome, refVec, windowsize: int, consensus: str, thr: float, s: int, w_min: int, w_max: int, q: int,
                    buff: int, flags: int, gap_open: int, gap_extend: int, score_threshold: int = 0,
                    ctx: Optional[Context] = None) -> ScanOutput:
    """kgma_scan with KGMA_MODE_STROBE"""
    ctx = ctx or default_context()
    S, N = _ints_of(refVec, ctx._lib)
    if S.size != 4 ** (2 * s):
        raise KmerGMAError(L.E_ARG, "refVec must have 4^(2s) entries")
    cons = str(consensus).upper().encode()
    prof = (L.Profile * 1)()
    prof[0].k, prof[0].n_refs, prof[0].window = w_max + s - 1, N, int(windowsize)
    prof[0].S = S.ctypes.data_as(C.POINTER(C.c_int32))
    prof[0].consensus, prof[0].consensus_len, prof[0].thr = cons, len(cons), float(thr)
    P = L.ScanParams(L.MODE_STROBE, flags, buff, gap_open, gap_extend, 0, 1, -1, 0, s, w_min, w_max, q, int(score_threshold))
    res = C.c_void_p()
    ctx.check(ctx._lib.kgma_scan(ctx._h, genome._h, prof, 1, C.byref(P), C.byref(res)))
    out = ScanOutput(ctx._lib, res)
    if flags & L.F_WANT_DISTS:
        out.load_dists(1)
    return out


def StrobeGMA(*, genome_path, refVec, consensus_refseq, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5,
              windowsize: int = 289, thr: float = 33.5, ScaleFactor: Optional[float] = None, buff: int = 50, do_align: bool = True,
              gap_open_score: int = -69, gap_extend_score: int = -5, score_threshold: int = 0,
              do_return_dists: bool = False, do_return_align: bool = False, get_hit_loci: bool = False,
              dist_vec: Optional[list] = None, result_align_vec: Optional[list] = None, hit_loci_vec: Optional[list] = None,
              genome_pos: int = 0, resultVec: Optional[list] = None, ctx: Optional[Context] = None):
    """StrobeGMA! (src/StrobemerGMA/StrobeGenomeMiner.jl:5-95): keyword arguments and in-place outputs as the reference's.
    The device path is fixed to ScaleFactor = 1/(w_max+s-1), what Strobemer_findGenes passes (:139)."""
    k = w_max + s - 1
    if ScaleFactor is not None and abs(float(ScaleFactor) - 1.0 / k) > 1e-15:
        raise KmerGMAError(L.E_UNSUPPORTED, "the device path uses ScaleFactor = 1/(w_max+s-1)")
    if do_return_align:
        raise KmerGMAError(L.E_UNSUPPORTED, "do_return_align is not available in strobemer mode")
    if genome_pos != 0:
        raise KmerGMAError(L.E_UNSUPPORTED, "genome_pos must start at 0")
    g = _as_genome(genome_path)
    flags = (L.F_ALIGN if do_align else 0) | (L.F_WANT_DISTS if do_return_dists else 0)
    out = strobe_scan_raw(g, refVec, windowsize, consensus_refseq, thr, s, w_min, w_max, q, buff, flags,
                          gap_open_score, gap_extend_score, score_threshold, ctx=ctx)
    resultVec = resultVec if resultVec is not None else []
    _emit(g, out, False, resultVec, hit_loci_vec if get_hit_loci else None, None)
    if do_return_dists and dist_vec is not None:
        dist_vec.extend(out.dists[0].tolist())
    return out


def Strobemer_findGenes(*, genome_path, ref_path, s: int = 2, w_min: int = 3, w_max: int = 5, q: int = 5,
                        KmerDistThr: float = 30, buffer: int = 50, do_align: bool = True, align_score_thr: int = 0,
                        do_return_dists: bool = False, do_return_hit_loci: bool = False, do_return_align: bool = False,
                        verbose: bool = True, ctx: Optional[Context] = None):
    """Strobemer_findGenes (StrobeGenomeMiner.jl:119-158): Any[hit_vector, (hit loci), (alignments), (distances)]"""
    RV, windowsize, cons = strobe_gen_ref_ws_cons(ref_path, s, w_min, w_max, q)
    hit_vector, dist_vec, hit_loci_vec = [], [], []
    if verbose:
        _log.info("initializing iteration...")
    StrobeGMA(genome_path=genome_path, refVec=RV, consensus_refseq=cons, s=s, w_min=w_min, w_max=w_max, q=q,
              windowsize=windowsize, thr=KmerDistThr, ScaleFactor=1.0 / (w_max + s - 1), buff=buffer, score_threshold=align_score_thr,
              do_align=do_align, do_return_dists=do_return_dists, do_return_align=do_return_align, get_hit_loci=do_return_hit_loci,
              dist_vec=dist_vec, hit_loci_vec=hit_loci_vec, resultVec=hit_vector, ctx=ctx)
    output_vector = [hit_vector]
    if do_return_hit_loci:
        output_vector.append(hit_loci_vec)
    if do_return_dists:
        output_vector.append(dist_vec)
    if verbose:
        _log.info("genome mining completed successfully, returning vector of: vector of hits")
    return output_vector


def write_results(KmerGMA_result_vec: Sequence[FastaRecord], file_path: str, width: int = 95):
    """src/API.jl:234-241: append the records to a FASTA file, `width` residues per line."""
    with open(file_path, "a") as fh:
        for hit in KmerGMA_result_vec:
            fh.write(">" + hit.description + "\n")
            s = hit.sequence
            for i in range(0, len(s), width):
                fh.write(s[i:i + width] + "\n")


def write_hits(out: "ScanOutput", genome: Genome, file_path: str, cluster: bool = False, with_genome_pos: bool = True, width: int = 95) -> int:
    """write_results (src/API.jl:234-241) for the hits of a scan, done by the library (kgma_result_write_fasta): headers as
    append_hit! builds them, sequences cut from the packed genome, `width` residues per line, appended to file_path."""
    n = C.c_int64()
    rc = out._lib.kgma_result_write_fasta(out._res, genome._h, os.fsencode(file_path), int(cluster), int(with_genome_pos), width, C.byref(n))
    if rc != 0:
        raise KmerGMAError(rc, f"cannot write {file_path}")
    return n.value


def hit_header(genome: Genome, hit, cluster: bool = False, with_genome_pos: bool = True) -> str:
    """the header text of one kgma_hit as the library formats it (kgma_hit_header)"""
    lib = L.load()
    h = L.Hit.from_buffer_copy(np.asarray(hit).tobytes())
    buf = C.create_string_buffer(4096)
    n = lib.kgma_hit_header(genome._h, C.byref(h), int(cluster), int(with_genome_pos), buf, 4096)
    if n < 0:
        raise KmerGMAError(int(n), "kgma_hit_header failed")
    return buf.value.decode()


def fasta_id_to_cumulative_len_dict(fasta_file_path) -> Dict[str, int]:
    """fasta_id_to_cumulative_len_dict (src/ExactMatch.jl:146-158): description of every record -> summed length of the records
    in front of it."""
    g = _as_genome(fasta_file_path)
    return {g.description(r): int(g._lib.kgma_genome_cumulative_len(g._h, r)) for r in range(len(g))}


# ------------------------------------------------------------------------------------------------
def align_unitrange(seq, seq_UnitRange: Tuple[int, int], consensus_seq: str, windowsize: int, sequence_length: int,
                    gap_open: int = -69, gap_extend: int = -1, ctx: Optional[Context] = None) -> Tuple[int, int]:
    """align_unitrange (src/Alignment.jl:33-52) on the GPU: seq is a residue string or (Genome, record)."""
    ctx = ctx or default_context()
    if isinstance(seq, tuple):
        g, rec = seq
    else:
        g, rec = Genome.from_records([("seq", seq)]), 0
    cons = consensus_seq.upper().encode()[:windowsize]
    recs = np.asarray([rec], dtype=np.int32)
    f = np.asarray([seq_UnitRange[0]], dtype=np.int64)
    l = np.asarray([seq_UnitRange[1]], dtype=np.int64)
    of, ol, sc = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.int64)
    ctx.check(ctx._lib.kgma_align_batch(ctx._h, g._h, cons, len(cons), gap_open, gap_extend, 0, 1,
                                        recs.ctypes.data, f.ctypes.data, l.ctypes.data,
                                        of.ctypes.data, ol.ctypes.data, sc.ctypes.data))
    return int(of[0]), int(ol[0])


def cigar_to_UnitRange(aligned_obj) -> Tuple[int, int]:
    """cigar_to_UnitRange (src/Alignment.jl:13-30) on an AlignResult or a CIGAR string: (lower + 1, num_sum) with `lower` the
    count of the FIRST operation whatever it is and `num_sum` the counts of all operations but the last (the loop returns when
    it reaches the last character, before that operation is added).  The scan computes the same two numbers on the device
    without a CIGAR; this is the text form for callers holding `do_return_align` results."""
    cig = aligned_obj.cigar if isinstance(aligned_obj, AlignResult) else str(aligned_obj)
    curr, nops, num_sum, lower = 0, 0, 0, 0
    for i, ch in enumerate(cig):
        if i == len(cig) - 1:
            return lower + 1, num_sum
        if ch.isdigit():
            curr = curr * 10 + int(ch)
        else:
            nops += 1
            if nops == 1:
                lower = curr
            num_sum += curr
            curr = 0
    return None                                            # empty CIGAR: the reference's loop falls through (returns nothing)


def firstMatch(reader, query, file=None, ctx: Optional[Context] = None) -> None:
    """firstMatch (src/ExactMatch.jl:8-16): prints "first:last identifier" for every record holding the query, the FIRST occurrence
    only (findfirst == the first range FindAllOverlap reports)."""
    ctx = ctx or default_context()
    g = _as_genome(reader)
    q = (query.sequence if isinstance(query, FastaRecord) else str(query)).upper().encode()
    by_rec = _exact_match_by_record(ctx, g, q, True, g._resident_ctx is ctx)
    for r in sorted(by_rec):                                # every record with a match, in file order (duplicate identifiers too)
        a, b = by_rec[r][0]
        print(f"{a}:{b} {g.identifier(r)}", file=file)


def _exact_match_by_record(ctx: Context, g: Genome, q: bytes, overlap: bool, resident: bool) -> Dict[int, List[Tuple[int, int]]]:
    mp = C.POINTER(L.Match)()
    n = C.c_int64()
    ctx.check(ctx._lib.kgma_exact_match(ctx._h, g._h, q, len(q), int(overlap), L.F_RESIDENT if resident else 0, C.byref(mp), C.byref(n)))
    by_rec: Dict[int, List[Tuple[int, int]]] = {}
    for i in range(n.value):
        by_rec.setdefault(mp[i].record, []).append((mp[i].first, mp[i].last))
    if n.value:
        ctx._lib.kgma_free(mp)
    return by_rec


def exactMatch(query, subject_seq, overlap: bool = True, ctx: Optional[Context] = None, resident: bool = False):
    """exactMatch (src/ExactMatch.jl:89-121).
    subject = residue string / FastaRecord  -> list of (first, last) ranges, or None when there is no match;
    subject = FASTA path / Genome (a Reader) -> dict identifier -> ranges, or the string "no match"."""
    ctx = ctx or default_context()
    if isinstance(query, FastaRecord):
        query = query.sequence
    q = str(query).upper().encode()
    single = isinstance(subject_seq, FastaRecord) or (isinstance(subject_seq, str) and not os.path.exists(subject_seq))
    if single:
        s = subject_seq.sequence if isinstance(subject_seq, FastaRecord) else subject_seq
        g = Genome.from_records([("subject", s)])
    else:
        g = _as_genome(subject_seq)
    by_rec = _exact_match_by_record(ctx, g, q, overlap, resident or g._resident_ctx is ctx)
    if single:
        return by_rec.get(0) or None
    identify = {}
    for r in sorted(by_rec):
        identify[g.identifier(r)] = by_rec[r]        # duplicate identifiers overwrite (ExactMatch.jl:112)
    return identify if identify else "no match"
