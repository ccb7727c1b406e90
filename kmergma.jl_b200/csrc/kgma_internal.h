// Internal declarations shared by the translation units of libkmergma_cuda.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <deque>
#include <cstring>
#include <cstdio>
#include "../../include/kmergma.h"

namespace kgma {

// ------------------------------------------------------------------ layout constants
constexpr int     REC_ALIGN   = 128;       // every record starts at a multiple of 128 bases (32 B of packed data)
constexpr int64_t TAIL_PAD    = 16384;     // zero bases after the last record (kernels may over-read)
constexpr int     FBLOCK      = 64;        // prefilter block: 64 bases = one 128-bit load per thread
constexpr int     FGROUP      = 32 * FBLOCK; // bases one warp covers per iteration
constexpr int     WFRAC       = 14;        // prefilter fixed point: weights are scaled by 2^14
constexpr uint32_t WCLAMP     = (1u << WFRAC) + 1;
constexpr int     MAX_K       = 7;
constexpr int     MAX_PROFILES = 16;

struct Record {
    std::string ident, desc;
    int64_t len = 0;
    int64_t off = 0;     // global base offset (multiple of REC_ALIGN)
};

}  // namespace kgma

struct kgma_genome {
    std::vector<kgma::Record> recs;
    // packed planes in the global coordinate space (records concatenated with alignment padding)
    uint32_t *seq2 = nullptr;      // 16 bases / word
    uint32_t *mask = nullptr;      // 32 bases / word
    int64_t   cap_bases = 0;       // allocated bases (multiple of 128)
    int64_t   G = 0;               // used bases incl. padding (set by seal; multiple of FGROUP)
    int64_t   total_len = 0;       // sum of record lengths
    bool      sealed = false;
    bool      ambiguous = false;   // a symbol other than A,C,G,T,N was ingested
    bool      any_mask = false;
    bool      pinned = false;      // planes are page-locked (cudaHostRegister / cudaHostAlloc)
    bool      host_alloc = false;  // allocated with cudaHostAlloc (else anonymous mmap, see genome_reserve)
    size_t    seq_map = 0, mask_map = 0;   // mapped bytes of the two planes when not host_alloc
    int64_t   amb_record = -1, amb_pos = -1;
    uint64_t  uid = 0;
    std::string err;
    std::vector<int64_t> nruns;    // cache: maximal runs of masked bases, [start,end) global positions (genome_nruns)
    uint64_t  nruns_uid = 0;
};

struct kgma_refs {
    std::vector<std::string> seqs;   // upper-case
    std::string err;
};

struct kgma_result {
    std::vector<kgma_hit> hits;
    std::vector<kgma_run> runs;
    std::vector<kgma_run_ext> run_ext;             // kgma_scan_shard: one per run (lo == 0: not extended)
    std::vector<int64_t>  first_D;                 // [n_profiles][n_records]
    std::vector<std::vector<double>> dists;        // per profile
    std::vector<char>     cigar_ops;
    std::vector<int32_t>  cigar_cnt;
    std::vector<kgma_align_event> align_events;    // cluster mode + KGMA_F_WANT_CIGARS: every extension of the final pass
};

// Results are recycled through a small pool: their hit / run vectors are 100-200 KB, which malloc serves with a fresh
// mmap (and page faults on first touch) per call -- a tenth of a millisecond on a 1.7 ms scan.
kgma_result *result_acquire();
void result_release(kgma_result *r);

struct kgma_ctx {
    int device = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_align = nullptr;
    cudaStream_t s_extend = nullptr;                      // stream of align_batch_device; null = s_compute (a pipelined scan moves it to s_align)
    cudaEvent_t  ev[8] = {};
    int num_sms = 0;
    size_t smem_optin = 0;
    std::string err;
    kgma_stats stats{};
    // device copy of one genome
    uint64_t  dg_uid = 0;
    uint32_t *d_seq2 = nullptr, *d_mask = nullptr;
    int64_t   d_cap_bases = 0;
    bool      d_seq_valid = false, d_mask_valid = false;
    int64_t   d_valid_lo = 0, d_valid_hi = 0;       // base range of seq2 a RESIDENT scan may reuse without upload
    int64_t   d_have_lo = 0, d_have_hi = 0;         // base range of seq2 physically present (same genome uid) - extension reads it
    // prefilter weight tables of recent scans (rebuilt only when the profiles / thresholds change)
    struct FTab {
        uint64_t key = 0; int M = 0; bool ok = false; double load = 0;
        bool nine = false;                 // tab9 (kgma_prefilter9) or tab (kgma_prefilter)
        std::vector<uint16_t> tab;         // [65536] 8-mer -> sum of the 9-k contained k-mer weights
        std::vector<uint8_t>  tab9;        // [3 * 65536] 9-mer with a ternary last base -> ceil(sum of 10-k weights / step)
        uint32_t step = 1, thrw = 0;       // flag when the covering sum of entries > thrw
    };
    std::deque<FTab> ftabs;        // deque: references stay valid while entries are appended
    // extension scratch, two slots so that extensions can be queued while a scan still streams
    void     *a_dev[2] = {}, *a_host[2] = {}; size_t a_dev_bytes[2] = {}, a_host_bytes[2] = {};
    cudaEvent_t a_ev0[2] = {}, a_ev1[2] = {}, a_done[2] = {};
    std::vector<cudaEvent_t> chunk_ev;             // one event per streamed chunk (pipelined scan)
    // staging ring for the first upload of a genome whose planes are not page-locked (scan.cu: StagedUpload)
    void *stage = nullptr; size_t stage_bytes = 0; cudaEvent_t stage_ev[16] = { nullptr };
    // scratch
    void     *d_scratch = nullptr; size_t d_scratch_bytes = 0;
    uint64_t  tab_sig = 0;          // which prefilter tables sit where in d_scratch (0 = none: whoever else carves the arena clears it)
    void     *h_scratch = nullptr; size_t h_scratch_bytes = 0;   // pinned
};

namespace kgma {

int set_err(kgma_ctx *ctx, int code, const char *fmt, ...);
#define KGMA_CUDA(ctx, call)                                                              \
    do { cudaError_t e_ = (call);                                                         \
         if (e_ != cudaSuccess) return kgma::set_err((ctx), KGMA_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                                                      cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ---- genome.cpp
int  genome_reserve(kgma_genome *g, int64_t bases);
int  genome_pin(kgma_ctx *ctx, kgma_genome *g);
const std::vector<int64_t> &genome_nruns(kgma_genome *g);
inline int base_code(const kgma_genome *g, int64_t gp) { return (g->seq2[gp >> 4] >> (2 * (gp & 15))) & 3; }
inline int base_masked(const kgma_genome *g, int64_t gp) { return (g->mask[gp >> 5] >> (gp & 31)) & 1; }

// ---- device genome management (scan.cu)
int  dev_genome_prepare(kgma_ctx *ctx, kgma_genome *g, bool need_mask);

// ---- per-profile exact integer tables (host side)
struct ProfTab {
    int k = 0; int64_t ws = 0, nk = 0; int32_t N = 0;
    int kb = 0;                      // the code space has 4^kb codes: kb = k for k-mers, 2s for strobemers (KGMA_MODE_STROBE)
    bool strobe = false;             // StrobeGMA!: codes through ScanPlan::cmap, loop of L-ws-1 steps, CMI = i, tracked-table quirk (scan.cu)
    std::vector<int32_t> S_rev;      // S in reversed-2-bit-group index order (matches packed bit order); strobemers: natural order
    int64_t N2 = 0, twoN = 0, sumS2 = 0;
    int64_t T = 0;                   // d < thr  <=>  D < T
    int64_t Tlo = 0, Thi = 0;        // |d - thr| <= 1e-9*thr  <=>  Tlo <= D < Thi  (near-threshold band)
    double  denom = 0;               // 2 k N^2
    double  thr = 0;
    uint64_t hash = 0;               // fingerprint of (S, Thi, N, nk, k): key of the cached prefilter tables
};
int  build_proftab(kgma_ctx *ctx, const kgma_profile &p, ProfTab &t, const kgma_scan_params *P = nullptr);
inline uint32_t rev_kmer(uint32_t c, int k) { uint32_t r = 0; for (int j = 0; j < k; j++) { r = (r << 2) | (c & 3); c >>= 2; } return r; }

// ---- replay.cpp
struct AlignReq { int32_t record, profile; int64_t first, last; int32_t hint = 0; };   // 1-based range to extend; hint & 1: the window's distance is
                                                                                     // close to the threshold (edge of a run: the consensus tends to overhang the slice)
inline int32_t align_hint(int64_t D, int64_t T) { return (__int128)D * 10 > (__int128)T * 9 ? 1 : 0; }
struct AlignRes { int64_t lo, hi, score; uint32_t cig_off, cig_len; };   // cigar_to_UnitRange result (relative, 1-based)
struct Pending { size_t hit; size_t req; int64_t first; size_t run; };     // hit waiting for its extension result (run: index of the run that emitted it)
void merge_runs(std::vector<kgma_run> &runs, std::vector<kgma_run_ext> *ext = nullptr);
void apply_extensions(kgma_genome *g, std::vector<kgma_hit> &hits, const std::vector<Pending> &pend, const std::vector<AlignRes> &ares);
int  replay_single_range(kgma_ctx *ctx, kgma_genome *g, const ProfTab &t, const kgma_scan_params &P,
                         std::vector<kgma_run> &runs, const std::vector<int64_t> &first_D, int r0, int r1,
                         int64_t *genome_pos_io, std::vector<kgma_hit> &hits, std::vector<AlignReq> &reqs, std::vector<Pending> &pend);
int  replay(kgma_ctx *ctx, kgma_genome *g, const std::vector<ProfTab> &tabs, const kgma_profile *profiles,
            const kgma_scan_params &P, std::vector<kgma_run> &runs, const std::vector<int64_t> &first_D,
            kgma_result *res, std::vector<kgma_run_ext> *ext = nullptr);

// ---- align.cu
struct AlignOut { long long score; int32_t lower, num_sum, nops, cig_n; };                 // nops < 0: redo with the path-summary kernel
struct AlignJob2 { long long gpos; int32_t n, a_off, m, b_off; int32_t mode, slot; };   // b_off >= 0: subject codes were uploaded (not on the device);
                                                                        // tagged kernel: mode & 2 / & 8 = second- / third-payload sweep only (a twin), mode & 1 / & 4 = such a twin is queued;
                                                                        // slot: index of the request (results are written there)
struct AlignArgs2 {
    const uint8_t *a;            // consensus codes 0..3, 4 = N
    const uint32_t *seq;         // packed 2-bit genome on the device
    const long long *nruns; int n_nruns;   // maximal runs of masked (N) bases, [start,end) global positions, ascending
    const uint8_t *b;            // uploaded subject codes for jobs whose slice is not on the device
    const AlignJob2 *jobs; int njobs;
    int *next_job;
    AlignOut *out;
    int go, ge, tie_open, ncol_cap, need_boundary;
    int tail_mode;               // tagged kernel: 0 = second sweep where needed, 1 = twins for every alignment, 2 = never (hand back instead)
    int nslots;                  // number of requests: second-sweep results go to out[nslots + slot]
};
struct AlignTicket {
    bool active = false; int slot = 0, nj = 0; void *ho = nullptr;
    // what align_collect needs to redo the alignments the tagged kernel handed back (align.cu)
    bool tagged = false; void *st = nullptr; size_t smem = 0; int maxn = 0, nq = 0;   // nq: queue entries (requests + twins)
    void *d_jobs2 = nullptr, *d_out2 = nullptr, *d_cnt = nullptr, *h_jobs = nullptr, *h_jobs2 = nullptr, *h_out2 = nullptr;
    AlignArgs2 args{};
};
int  align_enqueue(kgma_ctx *ctx, kgma_genome *g, const std::vector<AlignReq> &reqs, const kgma_profile *profiles, int n_profiles,
                   bool single_mode_truncate, int gap_open, int gap_extend, bool tie_open, cudaStream_t st, int slot, AlignTicket *t);
int  align_collect(kgma_ctx *ctx, AlignTicket *t, std::vector<AlignRes> &out);
int  align_batch_device(kgma_ctx *ctx, kgma_genome *g, const std::vector<AlignReq> &reqs,
                        const kgma_profile *profiles, int n_profiles, bool single_mode_truncate,
                        int gap_open, int gap_extend, bool tie_open, bool want_cigars,
                        std::vector<AlignRes> &out, std::vector<char> *cig_ops, std::vector<int32_t> *cig_cnt);

// ---- scratch helpers
int  dev_scratch(kgma_ctx *ctx, size_t bytes, void **out);
int  host_scratch(kgma_ctx *ctx, size_t bytes, void **out);

}  // namespace kgma
