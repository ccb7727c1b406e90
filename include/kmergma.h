/*
 * kmergma.h — C ABI of libkmergma_cuda, the B200 (sm_100a) implementation of
 * KmerGMA.jl's homology-scan hot path.
 *
 * The reference (pure Julia) has no FFI; the operator boundary this library
 * replaces is the set of keyword-argument functions its public API delegates to
 * (all citations are path:line in the reference repository):
 *
 *   ac_gma_testing!   src/GenomeMiner.jl:4-109     (called at src/API.jl:83-94)   -> kgma_scan (n_profiles = 1, KGMA_MODE_SINGLE)
 *   Omn_KmerGMA!      src/OmnGenomeMiner.jl:7-162  (called at src/API.jl:201-216) -> kgma_scan (KGMA_MODE_CLUSTER)
 *   record_KmerGMA!   src/MultiThread/GenomeMiner.jl:8-98                          -> kgma_scan on a one-record genome
 *   align_unitrange   src/Alignment.jl:33-52  + cigar_to_UnitRange :13-30         -> inside kgma_scan / kgma_replay, and kgma_align_batch
 *   exactMatch        src/ExactMatch.jl:89-121 (FindAll :20-30, FindAllOverlap :33-43) -> kgma_exact_match
 *   getSeq + NUCLEOTIDE_BITS   src/Consts.jl:22-39                                  -> kgma_genome_* ingest (2 bit/base + ambiguity mask)
 *   gen_ref_ws_cons / cluster_ref_API  src/ReferenceGeneration.jl:4-41,75-168      -> kgma_refs_* (integer k-mer sums, window, consensus)
 *
 * Conventions: plain C types only; every call returns an int status (0 = ok,
 * < 0 = error, text via kgma_last_error); no exceptions cross the ABI; no global
 * state (one kgma_ctx per caller thread / per GPU); all calls are blocking.
 * Positions are 1-based inclusive exactly as the reference prints them.
 * There is NO CPU fallback: kgma_create fails when no CUDA device is usable.
 */
#ifndef KMERGMA_H
#define KMERGMA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGMA_OK             0
#define KGMA_E_CUDA        (-1)   /* CUDA runtime error / no device */
#define KGMA_E_ARG         (-2)   /* invalid argument */
#define KGMA_E_SYMBOL      (-3)   /* symbol outside A,C,G,T,N (the reference's KeyError, Consts.jl:22-28) */
#define KGMA_E_IO          (-4)
#define KGMA_E_UNSUPPORTED (-5)   /* e.g. k > 7, window too large */
#define KGMA_E_CAPACITY    (-6)
#define KGMA_E_STATE       (-7)   /* e.g. genome not sealed */
#define KGMA_E_WINDOW      (-8)   /* k >= window size (error() at API.jl:70,177) */

#define KGMA_MODE_SINGLE   0      /* ac_gma_testing! semantics */
#define KGMA_MODE_CLUSTER  1      /* Omn_KmerGMA! semantics   */
#define KGMA_MODE_STROBE   2      /* StrobeGMA! semantics (src/StrobemerGMA/StrobeGenomeMiner.jl:5-95, experimental in the reference): one profile over
                                     the 4^(2s) gap-free 2-randstrobe codes; profile.k = w_max + s - 1 bases per code; always the dense pass */

/* kgma_scan_params.flags */
#define KGMA_F_ALIGN        (1u << 0)  /* do_align / align_hits */
#define KGMA_F_DENSE        (1u << 1)  /* skip the lower-bound prefilter: count-table kernel over every window */
#define KGMA_F_WANT_DISTS   (1u << 2)  /* do_return_dists (implies DENSE) */
#define KGMA_F_WANT_CIGARS  (1u << 3)  /* do_return_align: keep CIGAR ops of aligned hits */
#define KGMA_F_TIE_OPEN     (1u << 4)  /* alignment: on exact open/extend ties open a new gap (default: extend) */
#define KGMA_F_RESIDENT     (1u << 5)  /* keep / reuse the device copy of the genome (no H2D when already resident) */

/* kgma_hit.flags — hits whose reference-side Float64 result could legitimately differ (reported separately) */
#define KGMA_HIT_NEAR_THR   (1u << 0)  /* some window of the run lies within 1e-9 rel of thr */
#define KGMA_HIT_ARGMIN_TIE (1u << 1)  /* run minimum attained at more than one window, or equal to the minimum carried from an earlier run */
#define KGMA_HIT_ROUND_HALF (1u << 2)  /* dist*100 within 1e-9 of a rounding half-way point */

/* kgma_run.flags (besides KGMA_HIT_NEAR_THR / KGMA_HIT_ARGMIN_TIE) */
#define KGMA_RUN_OPEN_LEFT  (1u << 8)  /* run starts at the first window of its device segment */
#define KGMA_RUN_OPEN_RIGHT (1u << 9)  /* run reaches the last window of its device segment (merged with its neighbour on the host) */
#define KGMA_RUN_MARKER     (1u << 10) /* not a run: a single window with d >= thr inside the 1e-9 band around thr (reported, never replayed) */

typedef struct kgma_ctx    kgma_ctx;
typedef struct kgma_genome kgma_genome;
typedef struct kgma_refs   kgma_refs;
typedef struct kgma_result kgma_result;

/* A k-mer profile in exact integer form: RV[i] = S[i] / n_refs  (ReferenceGeneration.jl:35,118).
 * S is indexed like the reference's KFV (first base most significant, A0 C1 G2 T3). */
typedef struct {
    int32_t        k;
    int32_t        n_refs;        /* N: family (cluster) size */
    int64_t        window;        /* windowsize */
    const int32_t *S;             /* [4^k] summed k-mer counts over the n_refs references */
    const char    *consensus;     /* A,C,G,T,N bytes; single mode aligns consensus[1:window] (Alignment.jl:42),
                                     cluster mode the whole string (OmnGenomeMiner.jl:131) */
    int32_t        consensus_len;
    double         thr;           /* KmerDistThr / thr_vec[i] */
} kgma_profile;

typedef struct {
    int32_t  mode;                /* KGMA_MODE_* */
    uint32_t flags;               /* KGMA_F_* */
    int64_t  buff;                /* buff / buffer */
    int32_t  gap_open;            /* gap_open_score  (negative) */
    int32_t  gap_extend;          /* gap_extend_score (negative) */
    int32_t  shard_index;         /* this context scans shard_index of shard_count equal slices of the  */
    int32_t  shard_count;         /* packed genome (window-length halo handled inside); 0/1 = everything */
    int32_t  only_record;         /* >= 0: scan just this record (record_KmerGMA!); -1 = all */
    int32_t  reserved;
    /* KGMA_MODE_STROBE only (zero otherwise): randstrobe parameters s, w_min, w_max, q (Strobemers.jl:45-65) and process_hit!'s
     * score_threshold (Alignment.jl: an extended hit whose alignment scores below it is dropped) */
    int32_t  strobe_s, strobe_w_min, strobe_w_max, strobe_q;
    int64_t  score_threshold;
} kgma_scan_params;

/* One maximal stretch of consecutive loop steps with d < thr: the sufficient
 * statistic the host replays the reference's sequential state machine from. Steps are the
 * reference's loop counter t (GenomeMiner.jl:60) / i (OmnGenomeMiner.jl:89): window start (0-based) = step. */
typedef struct {
    int32_t record;
    int32_t profile;              /* 0-based */
    int64_t t_first, t_last;
    int64_t t_argmin;             /* first step attaining D_min */
    int64_t D_min;                /* exact integer distance numerator: d = D / (2 k N^2) */
    uint32_t flags;               /* KGMA_HIT_NEAR_THR | KGMA_HIT_ARGMIN_TIE */
    uint32_t reserved;
} kgma_run;

/* Extension result of a run's own candidate window (the window whose start is t_argmin), computed by the shard that
 * owns the run: what cigar_to_UnitRange returns for it, relative to the window's first base.  lo == 0: not extended. */
typedef struct { int64_t lo, hi, score; } kgma_run_ext;

typedef struct {
    int32_t  record;              /* 0-based index of the genome record */
    int32_t  profile;             /* "KFV = n" (1-based) in cluster mode, 0 in single mode */
    int64_t  cmi;                 /* CMI used to build the hit range */
    int64_t  first, last;         /* MatchPos first:last (1-based, after alignment when KGMA_F_ALIGN) */
    int64_t  genome_pos;          /* GenomePos */
    int64_t  D;                   /* exact numerator of the reported distance */
    double   dist;                /* D / (2 k N^2): the value the reference prints rounded to 2 digits */
    int64_t  align_score;
    uint32_t flags;               /* KGMA_HIT_* */
    uint32_t cigar_off, cigar_len;/* into kgma_result_cigar_* when KGMA_F_WANT_CIGARS */
    uint32_t reserved;
} kgma_hit;

/* Device-time breakdown of the last scan on a context (CUDA events, milliseconds). */
typedef struct {
    double h2d_ms, filter_ms, exact_ms, align_ms, total_ms;
    int64_t bases_scanned;        /* window starts covered by this context's shard */
    int64_t blocks_total, blocks_flagged;   /* prefilter blocks (64 bases each) */
    int64_t exact_windows;        /* windows evaluated by the count-table kernel */
    int64_t n_runs, n_align;
    int64_t launches;             /* kernels launched by the last call */
    int64_t h2d_bytes, d2h_bytes;
    /* host wall-clock breakdown of the last call (milliseconds) */
    double wall_ms;               /* whole call */
    double host_setup_ms;         /* plan, tables, scratch, small uploads (until the genome stream starts) */
    double host_cand_ms;          /* candidate list D2H + sort + segment building */
    double host_replay_ms;        /* run merge + replay of the reference state machine (without the extension kernel) */
    int64_t n_align_redo;         /* extensions whose last-row maximum is also attained at column n (second sweep of the tagged kernel) */
    int64_t filter_passes;        /* prefilter passes over the shard (one per group of profiles sharing a weight table); filter_ms spans all */
    int64_t n_align_summary;      /* extensions redone by the path-summary kernel (only with KGMA_ALIGN_TAIL=off, a testing switch) */
    int64_t n_align_head;         /* extensions whose path starts at column 0 of the slice (third sweep of the tagged kernel) */
} kgma_stats;

/* ---- context ---------------------------------------------------------------------------- */
int  kgma_create(int device, kgma_ctx **out);
void kgma_destroy(kgma_ctx *ctx);
const char *kgma_last_error(const kgma_ctx *ctx);      /* ctx may be NULL: last error of kgma_create */
int  kgma_get_stats(const kgma_ctx *ctx, kgma_stats *out);
int  kgma_version(void);

/* ---- genome ingest: replaces FASTA.Reader + getSeq + per-base NUCLEOTIDE_BITS lookups --- */
int  kgma_genome_create(kgma_genome **out);
int  kgma_genome_from_fasta(const char *path, kgma_genome **out);
/* ASCII residues (upper/lower case). A,C,G,T -> 0..3; N -> 3 + mask bit; other IUPAC -> mask bit and the
 * genome is marked ambiguous (scans then fail with KGMA_E_SYMBOL like the reference's KeyError; exact match still works). */
int  kgma_genome_append_ascii(kgma_genome *g, const char *identifier, const char *description,
                              const char *seq, int64_t len);
/* Pre-packed input (north_star's canonical path): 2 bits/base, 16 bases per little-endian uint32 (base i of
 * the record in bits 2*(i%16)), N already folded to 3; mask: 1 bit/base, 32 per uint32 (may be NULL = no N). */
int  kgma_genome_append_packed(kgma_genome *g, const char *identifier, const char *description,
                               const uint32_t *seq2, const uint32_t *mask, int64_t len);
/* BioSequences LongSequence{DNAAlphabet{4}}.data as handed over by the Julia shim: 4 bits/base, 16 per UInt64,
 * first symbol in the least significant nibble, one-hot A=1 C=2 G=4 T=8, N=15. */
int  kgma_genome_append_bio4(kgma_genome *g, const char *identifier, const char *description,
                             const uint64_t *data, int64_t len);
/* Zero-copy ingest for a caller that packs the genome itself: the library lays the records out in page-locked planes of its
 * own (records start at multiples of 128 bases), kgma_genome_record_planes returns where record r's words go (same word
 * format as kgma_genome_append_packed; the mask words are zero-initialised), the caller fills them in place and seals.
 * Scans then stream straight from these planes (the e2e tier of bench.py) instead of staging a pageable copy. */
int  kgma_genome_create_pinned(kgma_ctx *ctx, int n_records, const int64_t *rec_len, kgma_genome **out);
int  kgma_genome_record_planes(kgma_genome *g, int record, uint32_t **seq2, uint32_t **mask);
int  kgma_genome_set_names(kgma_genome *g, int record, const char *identifier, const char *description);
/* a genome of its own (page-locked planes) holding copies of the given records of g, in that order: one device's share of a
 * contig-partitioned multi-GPU scan (kgma_hits_merge_partition puts the devices' hits together) */
int  kgma_genome_subset(kgma_ctx *ctx, const kgma_genome *g, const int32_t *records, int n_records, kgma_genome **out);
int  kgma_genome_seal(kgma_genome *g);
void kgma_genome_destroy(kgma_genome *g);
int     kgma_genome_n_records(const kgma_genome *g);
int64_t kgma_genome_record_len(const kgma_genome *g, int record);
int64_t kgma_genome_total_len(const kgma_genome *g);
const char *kgma_genome_identifier(const kgma_genome *g, int record);   /* FASTA.identifier */
const char *kgma_genome_description(const kgma_genome *g, int record);  /* FASTA.description */
/* view(seq, first:last) as upper-case ASCII (N restored from the mask); 1-based inclusive */
int  kgma_genome_get_seq(const kgma_genome *g, int record, int64_t first, int64_t last, char *out);
/* Maximal runs of masked residues (N and any other non-ACGT symbol) of a sealed genome as [start,end) pairs, 0-based in
 * the packed coordinate space (record r starts at kgma_genome_record_offset).  This is the list the extension and
 * exact-match kernels consult instead of reading the ambiguity plane.  *out is malloc'd (kgma_free), *n_runs pairs. */
int  kgma_genome_masked_runs(kgma_genome *g, int64_t **out, int64_t *n_runs);
int64_t kgma_genome_record_offset(const kgma_genome *g, int record);
/* Overwrite residues in place before sealing/uploading (used to plant homologues in synthetic genomes). */
int  kgma_genome_put_seq(kgma_genome *g, int record, int64_t first, const char *seq, int64_t len);
/* Synthetic genome: record r gets rec_len[r] bases, base(p) = splitmix64(seed ^ global p) & 3,
 * generated on the GPU of ctx; N runs: n_run_len bases at both ends of every record and one run of
 * centromere_len in the middle (0 = none). */
int  kgma_genome_synth(kgma_ctx *ctx, int n_records, const int64_t *rec_len, uint64_t seed,
                       int64_t n_run_len, int64_t centromere_len, kgma_genome **out);
/* Upload and keep the packed genome in HBM (optional; scans without KGMA_F_RESIDENT stream it every call). */
int  kgma_genome_make_resident(kgma_ctx *ctx, kgma_genome *g);
int  kgma_genome_drop_resident(kgma_ctx *ctx, kgma_genome *g);

/* ---- reference family -> profiles: gen_ref_ws_cons / cluster_ref_API / eliminate_null_params ------- */
int  kgma_refs_from_fasta(const char *path, kgma_refs **out);
int  kgma_refs_create(kgma_refs **out);
int  kgma_refs_append_ascii(kgma_refs *r, const char *seq, int64_t len);
void kgma_refs_destroy(kgma_refs *r);
int  kgma_refs_count(const kgma_refs *r);
int64_t kgma_refs_maxlen(const kgma_refs *r);
/* gen_ref_ws_cons: S[4^k] (integer sums), n_refs, window = Int(round(sum_len*(1/N))), consensus[maxlen+1] */
int  kgma_refs_profile(const kgma_refs *r, int k, int32_t *S, int32_t *n_refs, int64_t *window, char *consensus);
/* gen_ref_ws_cons of the strobemer path (src/StrobemerGMA/StrobeRefGen.jl:4-42): S[4^(2s)] summed counts of the gap-free
 * 2-randstrobe codes (Strobemers.jl:45-65,105-115) */
int  kgma_refs_strobe_profile(const kgma_refs *r, int s, int w_min, int w_max, int q, int32_t *S, int32_t *n_refs, int64_t *window, char *consensus);
/* cluster_ref_API (+ eliminate_null_params when drop_empty): returns the number of profiles written (<= n_cutoffs+2)
 * or < 0.  S: [max][4^k]; consensus: [max][cons_stride] NUL-terminated (truncated to window except the appended average). */
int  kgma_refs_cluster(const kgma_refs *r, int k, const double *cutoffs, int n_cutoffs, int include_avg,
                       int drop_empty, int32_t *S, int32_t *n_members, int64_t *windows,
                       char *consensus, int64_t cons_stride, int32_t *invalid, double *ref_dists);
/* Recover (S, N) from a Float64 KFV handed to the operator interface (refVec = S .* (1/N) or S ./ N). */
int  kgma_profile_from_kfv(const double *kfv, int64_t n_bins, int32_t max_n, int32_t *S, int32_t *n_refs);

/* ---- the scan: ac_gma_testing! / Omn_KmerGMA! ------------------------------------------------------ */
/* Whole operator on one GPU: stream genome -> prefilter -> count-table kernel -> run compaction ->
 * batched extension -> host replay -> hits. */
int  kgma_scan(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
               const kgma_scan_params *params, kgma_result **out);
/* Multi-GPU form: each context scans its shard and returns run summaries only; the caller concatenates the
 * runs of all shards (any order) and calls kgma_replay once (on any context) to obtain the hits. */
int  kgma_scan_runs(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                    const kgma_scan_params *params, kgma_result **out);
int  kgma_replay(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                 const kgma_scan_params *params, const kgma_run *runs, int64_t n_runs,
                 const int64_t *first_window_D /* [n_profiles][n_records], from kgma_result_first_D */,
                 kgma_result **out);

/* Multi-GPU, one blocking call per rank and no device work afterwards: scan this context's shard, merge its runs and
 * extend, on this GPU, the candidate window of every run that can still become a hit (its own first argmin -- a superset
 * of what the replay selects).  kgma_result_runs / kgma_result_run_ext / kgma_result_first_D describe the outcome;
 * kgma_result_pack lays them out in one block (returns the bytes needed; nothing is written when cap is smaller), the
 * ranks exchange equal-sized blocks (one all-gather), and kgma_replay_packed -- host only, ctx may be NULL -- merges the
 * shards' runs and replays the reference's state machine from them, looking extension results up instead of computing. */
int  kgma_scan_shard(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                     const kgma_scan_params *params, kgma_result **out);
const kgma_run_ext *kgma_result_run_ext(const kgma_result *r);
int64_t kgma_result_pack(const kgma_result *r, void *buf, int64_t cap);
int  kgma_replay_packed(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                        const kgma_scan_params *params, const void *blocks, int n_blocks, int64_t stride, kgma_result **out);
/* The other cut: whole records per GPU (a genome whose contigs balance over the GPUs).  Every rank runs the ordinary kgma_scan on
 * a genome of its own records and ships [int64 n][8 bytes][kgma_hit x n]; this merges the blocks on one host: rec_map[rec_off[b] +
 * local record] = global record index, genome_pos_of[global record] = GenomePos of that record in the whole genome.  *out is
 * malloc'ed (kgma_free). */
int  kgma_hits_merge_partition(const void *blocks, int n_blocks, int64_t stride, const int32_t *rec_map, const int32_t *rec_off,
                               int32_t n_records, const int64_t *genome_pos_of, kgma_hit **out, int64_t *n_out);

int64_t kgma_result_n_hits(const kgma_result *r);
const kgma_hit *kgma_result_hits(const kgma_result *r);
int64_t kgma_result_n_runs(const kgma_result *r);
const kgma_run *kgma_result_runs(const kgma_result *r);
/* exact D of window 1 of every record, [n_profiles][n_records]; INT64_MIN where the record is shorter than the window
 * or outside this shard */
const int64_t *kgma_result_first_D(const kgma_result *r);
/* do_return_dists: d for steps 1..(L-ws) of every scanned record, concatenated in record order (profile p) */
int64_t kgma_result_n_dists(const kgma_result *r, int profile);
const double *kgma_result_dists(const kgma_result *r, int profile);
/* CIGAR ops of aligned hits: op chars ('=','X','I','D') and run lengths */
const char *kgma_result_cigar_ops(const kgma_result *r);
const int32_t *kgma_result_cigar_counts(const kgma_result *r);
/* Cluster mode with KGMA_F_WANT_CIGARS: one entry per extension the reference performs, in its order.  Omn_KmerGMA! pushes
 * the alignment (`get_aligns`, OmnGenomeMiner.jl:131-133) BEFORE the second overlap test (:139), so the list also holds the
 * extensions whose hit was then rejected (emitted == 0); the emitted ones are the hits, in hit order.  Empty otherwise (in
 * single mode every extension is a hit: Alignment.jl:46). */
typedef struct {
    int32_t  record;              /* 0-based */
    int32_t  profile;             /* 1-based, like kgma_hit.profile */
    int64_t  cmi;
    int64_t  align_score;
    uint32_t cigar_off, cigar_len;/* into kgma_result_cigar_* */
    uint32_t emitted;             /* 1: passed the second overlap test and became a hit */
    uint32_t reserved;
} kgma_align_event;
int64_t kgma_result_n_align_events(const kgma_result *r);
const kgma_align_event *kgma_result_align_events(const kgma_result *r);
void kgma_result_free(kgma_result *r);

/* ---- result formatting / writing: append_hit! header text (Alignment.jl:57-81; OmnGenomeMiner.jl:141-149 when cluster != 0)
 * and write_results (API.jl:234-241: records appended to a FASTA file, `width` residues per line) -------------------------- */
/* "<id> | dist = <round2> | MatchPos = a:b | GenomePos = g | Len = n" with Julia's string(round(d, digits = 2)); returns the
 * length (the text is written, NUL-terminated, when cap is larger). */
int64_t kgma_hit_header(const kgma_genome *g, const kgma_hit *h, int cluster, int with_genome_pos, char *buf, int64_t cap);
int  kgma_result_write_fasta(const kgma_result *r, const kgma_genome *g, const char *path, int cluster, int with_genome_pos,
                             int width, int64_t *n_written);
/* fasta_id_to_cumulative_len_dict (ExactMatch.jl:146-158): summed length of the records in front of `record` */
int64_t kgma_genome_cumulative_len(const kgma_genome *g, int record);

/* ---- batched semi-global extension (Alignment.jl:33-52) ------------------------------------------- */
/* For each i: align consensus (A,C,G,T,N bytes) against record[first_i:last_i]; returns the remapped
 * 1-based range exactly as align_unitrange does, plus the score. */
int  kgma_align_batch(kgma_ctx *ctx, kgma_genome *g, const char *consensus, int32_t cons_len,
                      int32_t gap_open, int32_t gap_extend, uint32_t flags, int64_t n,
                      const int32_t *record, const int64_t *first, const int64_t *last,
                      int64_t *out_first, int64_t *out_last, int64_t *out_score);

/* ---- exactMatch (ExactMatch.jl:100-121) ----------------------------------------------------------- */
typedef struct {
    int32_t record;
    int32_t reserved;
    int64_t first, last;          /* 1-based inclusive range, ascending within a record */
} kgma_match;
int  kgma_exact_match(kgma_ctx *ctx, kgma_genome *g, const char *query, int64_t qlen, int overlap,
                      uint32_t flags /* KGMA_F_RESIDENT */, kgma_match **out, int64_t *n_out);
/* Multi-GPU form (halo qlen-1): each context searches slice shard_index of shard_count of the packed genome and
 * returns the occurrence starts it owns (0-based, packed coordinate space -- see kgma_genome_record_offset; *starts is
 * malloc'd, kgma_free); every occurrence is owned by exactly one slice.  kgma_exact_match_merge (host only) turns the
 * concatenated starts of all slices into exactMatch's result, applying FindAll's non-overlap rule across slice edges. */
int  kgma_exact_match_shard(kgma_ctx *ctx, kgma_genome *g, const char *query, int64_t qlen, uint32_t flags,
                            int shard_index, int shard_count, int64_t **starts, int64_t *n_out);
int  kgma_exact_match_merge(const kgma_genome *g, const int64_t *starts, int64_t n, int64_t qlen, int overlap,
                            kgma_match **out, int64_t *n_out);
void kgma_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* KMERGMA_H */
