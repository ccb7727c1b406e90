"""ctypes binding of libkmergma_cuda.so (include/kmergma.h).  There is no fallback: if the shared
library is missing the import fails, and if no B200 is visible kgma_create fails."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libkmergma_cuda.so")

# status codes (kmergma.h)
OK, E_CUDA, E_ARG, E_SYMBOL, E_IO, E_UNSUPPORTED, E_CAPACITY, E_STATE, E_WINDOW = 0, -1, -2, -3, -4, -5, -6, -7, -8
MODE_SINGLE, MODE_CLUSTER, MODE_STROBE = 0, 1, 2
F_ALIGN, F_DENSE, F_WANT_DISTS, F_WANT_CIGARS, F_TIE_OPEN, F_RESIDENT = 1, 2, 4, 8, 16, 32
HIT_NEAR_THR, HIT_ARGMIN_TIE, HIT_ROUND_HALF = 1, 2, 4
RUN_OPEN_LEFT, RUN_OPEN_RIGHT, RUN_MARKER = 1 << 8, 1 << 9, 1 << 10


class Profile(C.Structure):
    _fields_ = [("k", C.c_int32), ("n_refs", C.c_int32), ("window", C.c_int64),
                ("S", C.POINTER(C.c_int32)), ("consensus", C.c_char_p), ("consensus_len", C.c_int32),
                ("thr", C.c_double)]


class ScanParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("flags", C.c_uint32), ("buff", C.c_int64),
                ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("shard_index", C.c_int32), ("shard_count", C.c_int32),
                ("only_record", C.c_int32), ("reserved", C.c_int32),
                ("strobe_s", C.c_int32), ("strobe_w_min", C.c_int32), ("strobe_w_max", C.c_int32), ("strobe_q", C.c_int32),
                ("score_threshold", C.c_int64)]


class Run(C.Structure):
    _fields_ = [("record", C.c_int32), ("profile", C.c_int32), ("t_first", C.c_int64), ("t_last", C.c_int64),
                ("t_argmin", C.c_int64), ("D_min", C.c_int64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class RunExt(C.Structure):
    _fields_ = [("lo", C.c_int64), ("hi", C.c_int64), ("score", C.c_int64)]


class Hit(C.Structure):
    _fields_ = [("record", C.c_int32), ("profile", C.c_int32), ("cmi", C.c_int64),
                ("first", C.c_int64), ("last", C.c_int64), ("genome_pos", C.c_int64),
                ("D", C.c_int64), ("dist", C.c_double), ("align_score", C.c_int64),
                ("flags", C.c_uint32), ("cigar_off", C.c_uint32), ("cigar_len", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("filter_ms", C.c_double), ("exact_ms", C.c_double),
                ("align_ms", C.c_double), ("total_ms", C.c_double),
                ("bases_scanned", C.c_int64), ("blocks_total", C.c_int64), ("blocks_flagged", C.c_int64),
                ("exact_windows", C.c_int64), ("n_runs", C.c_int64), ("n_align", C.c_int64),
                ("launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("wall_ms", C.c_double), ("host_setup_ms", C.c_double), ("host_cand_ms", C.c_double),
                ("host_replay_ms", C.c_double), ("n_align_redo", C.c_int64), ("filter_passes", C.c_int64),
                ("n_align_summary", C.c_int64), ("n_align_head", C.c_int64)]


class AlignEvent(C.Structure):
    _fields_ = [("record", C.c_int32), ("profile", C.c_int32), ("cmi", C.c_int64), ("align_score", C.c_int64),
                ("cigar_off", C.c_uint32), ("cigar_len", C.c_uint32), ("emitted", C.c_uint32), ("reserved", C.c_uint32)]


class Match(C.Structure):
    _fields_ = [("record", C.c_int32), ("reserved", C.c_int32), ("first", C.c_int64), ("last", C.c_int64)]


# every symbol include/kmergma.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "kgma_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "kgma_destroy": (None, [_P]),
    "kgma_last_error": (C.c_char_p, [_P]),
    "kgma_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "kgma_version": (C.c_int, []),
    "kgma_genome_create": (C.c_int, [C.POINTER(_P)]),
    "kgma_genome_from_fasta": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "kgma_genome_append_ascii": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]),
    "kgma_genome_append_packed": (C.c_int, [_P, C.c_char_p, C.c_char_p, _P, _P, C.c_int64]),
    "kgma_genome_append_bio4": (C.c_int, [_P, C.c_char_p, C.c_char_p, _P, C.c_int64]),
    "kgma_genome_create_pinned": (C.c_int, [_P, C.c_int, _P, C.POINTER(_P)]),
    "kgma_genome_record_planes": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P)]),
    "kgma_genome_set_names": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_char_p]),
    "kgma_genome_subset": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(_P)]),
    "kgma_genome_seal": (C.c_int, [_P]),
    "kgma_genome_destroy": (None, [_P]),
    "kgma_genome_n_records": (C.c_int, [_P]),
    "kgma_genome_record_len": (C.c_int64, [_P, C.c_int]),
    "kgma_genome_total_len": (C.c_int64, [_P]),
    "kgma_genome_identifier": (C.c_char_p, [_P, C.c_int]),
    "kgma_genome_description": (C.c_char_p, [_P, C.c_int]),
    "kgma_genome_get_seq": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_char_p]),
    "kgma_genome_masked_runs": (C.c_int, [_P, C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.c_int64)]),
    "kgma_genome_record_offset": (C.c_int64, [_P, C.c_int]),
    "kgma_genome_put_seq": (C.c_int, [_P, C.c_int, C.c_int64, C.c_char_p, C.c_int64]),
    "kgma_genome_synth": (C.c_int, [_P, C.c_int, _P, C.c_uint64, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "kgma_genome_make_resident": (C.c_int, [_P, _P]),
    "kgma_genome_drop_resident": (C.c_int, [_P, _P]),
    "kgma_refs_from_fasta": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "kgma_refs_create": (C.c_int, [C.POINTER(_P)]),
    "kgma_refs_append_ascii": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "kgma_refs_destroy": (None, [_P]),
    "kgma_refs_count": (C.c_int, [_P]),
    "kgma_refs_maxlen": (C.c_int64, [_P]),
    "kgma_refs_profile": (C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_char_p]),
    "kgma_refs_strobe_profile": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_char_p]),
    "kgma_refs_cluster": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_char_p, C.c_int64, _P, _P]),
    "kgma_profile_from_kfv": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.POINTER(C.c_int32)]),
    "kgma_scan": (C.c_int, [_P, _P, C.POINTER(Profile), C.c_int, C.POINTER(ScanParams), C.POINTER(_P)]),
    "kgma_scan_runs": (C.c_int, [_P, _P, C.POINTER(Profile), C.c_int, C.POINTER(ScanParams), C.POINTER(_P)]),
    "kgma_replay": (C.c_int, [_P, _P, C.POINTER(Profile), C.c_int, C.POINTER(ScanParams), _P, C.c_int64, _P, C.POINTER(_P)]),
    "kgma_scan_shard": (C.c_int, [_P, _P, C.POINTER(Profile), C.c_int, C.POINTER(ScanParams), C.POINTER(_P)]),
    "kgma_result_run_ext": (C.POINTER(RunExt), [_P]),
    "kgma_result_pack": (C.c_int64, [_P, _P, C.c_int64]),
    "kgma_replay_packed": (C.c_int, [_P, _P, C.POINTER(Profile), C.c_int, C.POINTER(ScanParams), _P, C.c_int, C.c_int64, C.POINTER(_P)]),
    "kgma_result_n_hits": (C.c_int64, [_P]),
    "kgma_result_hits": (C.POINTER(Hit), [_P]),
    "kgma_result_n_runs": (C.c_int64, [_P]),
    "kgma_result_runs": (C.POINTER(Run), [_P]),
    "kgma_result_first_D": (C.POINTER(C.c_int64), [_P]),
    "kgma_result_n_dists": (C.c_int64, [_P, C.c_int]),
    "kgma_result_dists": (C.POINTER(C.c_double), [_P, C.c_int]),
    "kgma_result_cigar_ops": (C.POINTER(C.c_char), [_P]),
    "kgma_result_cigar_counts": (C.POINTER(C.c_int32), [_P]),
    "kgma_result_n_align_events": (C.c_int64, [_P]),
    "kgma_result_align_events": (C.POINTER(AlignEvent), [_P]),
    "kgma_result_free": (None, [_P]),
    "kgma_hits_merge_partition": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, C.c_int32, _P, C.POINTER(C.POINTER(Hit)), C.POINTER(C.c_int64)]),
    "kgma_hit_header": (C.c_int64, [_P, C.POINTER(Hit), C.c_int, C.c_int, C.c_char_p, C.c_int64]),
    "kgma_result_write_fasta": (C.c_int, [_P, _P, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "kgma_genome_cumulative_len": (C.c_int64, [_P, C.c_int]),
    "kgma_align_batch": (C.c_int, [_P, _P, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_int64,
                                   _P, _P, _P, _P, _P, _P]),
    "kgma_exact_match": (C.c_int, [_P, _P, C.c_char_p, C.c_int64, C.c_int, C.c_uint32, C.POINTER(C.POINTER(Match)), C.POINTER(C.c_int64)]),
    "kgma_exact_match_shard": (C.c_int, [_P, _P, C.c_char_p, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.c_int64)]),
    "kgma_exact_match_merge": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.POINTER(Match)), C.POINTER(C.c_int64)]),
    "kgma_free": (None, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libkmergma_cuda.so and declare every prototype; raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `make -C kmergma.jl_b200/csrc` "
                              "(or __graft_entry__.build()); there is no fallback implementation")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
