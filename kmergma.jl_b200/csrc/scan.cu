// Rolling k-mer window scan on sm_100a: replaces the hot loops of
//   ac_gma_testing!  src/GenomeMiner.jl:60-105   and   Omn_KmerGMA!  src/OmnGenomeMiner.jl:89-158.
//
// Exact integer form: with RV = S/N,
//     D_w = sum_i (N c_w[i] - S[i])^2 = N^2 Q_w - 2N A_w + sum S^2,   d_w = D_w / (2 k N^2),
//     Q_w = sum_i c_w[i]^2,  A_w = sum_{p in w} S[kmer_p],            d < thr  <=>  D < T.
//
// Two kernels:
//  (1) kgma_prefilter<K>: streams the 2-bit genome once (0.25 B/base, one 128-bit coalesced load per
//      thread per 64 bases) and evaluates the rigorous lower bound  Q_w >= nk  =>  D_w >= N^2 nk - 2N A_w + sum S^2.
//      A_w is bounded from above by the sum of per-64-base block sums of a fixed-point weight
//      W[kmer] >= S[kmer]*2N/R (max over profiles), looked up 9-K k-mers at a time in a
//      shared-memory table indexed by 8-mers (128 KB).  Blocks whose covering sum cannot reach the
//      threshold contain no window with d < thr; the others are appended to a candidate list with
//      warp ballot/popc compaction + one atomic per warp.
//  (2) kgma_eval: the count-table kernel.  One warp per span of consecutive windows (a flagged 64-base
//      block read straight from the device candidate list, or a slice of a record in dense mode), with a
//      4^k x u16 count table in shared memory next to the profile's S table; the distance is updated
//      incrementally as one k-mer leaves and one enters the window, re-initialised per span -- 16 steps at a
//      time with all 32 lanes busy (one MATCH.ANY resolves the in-batch dependencies between the 32 events).
//      Windows are classified 16 at a time (ballot + warp min) and run summaries (maximal stretches of
//      D < T) are appended with an atomic; optionally every D is written (do_return_dists).
#include "kgma_internal.h"
#include <algorithm>
#include <type_traits>
#include <cmath>
#include <climits>
#include <cstdlib>
#include <chrono>
#include <functional>
#include <atomic>
#include <thread>
#include <sys/mman.h>
#include "staged_upload.h"
#include <cuda/barrier>

#ifndef KGMA_NSTAMP
#define KGMA_NSTAMP 2048
#endif

namespace kgma {

// =============================================================================================
// (1) prefilter
// =============================================================================================
struct FilterArgs {
    const uint4    *seq;        // packed genome, 64 bases per uint4
    const uint16_t *tab;        // [65536] 8-mer -> clamped sum of the (9-K) k-mer weights
    int64_t  blk_begin, blk_end;   // target blocks [begin,end) (multiples of 32)
    int      M;                 // covering blocks per window (<= 32)
    uint32_t thrw;              // flag when covering sum > thrw
    uint32_t *cand;             // out: flagged block ids
    uint32_t  cand_cap;
    uint32_t *cand_count;       // out: number flagged (may exceed cap -> caller falls back to dense)
    uint32_t *bitmap;           // out: one bit per block, set for flagged blocks (lets the count-table kernel join neighbours)
};

template <int K>
__device__ __forceinline__ uint32_t block_weight(const uint16_t *tab, uint4 w, uint32_t nw)
{
    constexpr int STRIDE = 9 - K;                       // k-mers fully contained in one 8-mer
    constexpr int NLOOK = (FBLOCK + STRIDE - 1) / STRIDE;
    const uint32_t W[5] = { w.x, w.y, w.z, w.w, nw };
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < NLOOK; i++) {
        constexpr int dummy = 0; (void)dummy;
        const int b = 2 * i * STRIDE;                   // bit offset of the 8-mer
        const int wi = b >> 5, s = b & 31;
        uint32_t idx;
        if (s + 16 <= 32) idx = (W[wi] >> s) & 0xFFFFu;
        else idx = __funnelshift_r(W[wi], W[wi + 1], s) & 0xFFFFu;
        sum += tab[idx];
    }
    return sum;
}

template <int K>
__global__ void __launch_bounds__(1024, 1) kgma_prefilter(FilterArgs a)
{
    extern __shared__ __align__(16) uint16_t s_tab[];
    {   // stage the 128 KB weight table into shared memory with 128-bit copies
        const uint4 *src = reinterpret_cast<const uint4 *>(a.tab);
        uint4 *dst = reinterpret_cast<uint4 *>(s_tab);
        for (int i = threadIdx.x; i < 65536 * 2 / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t grp0 = a.blk_begin >> 5, ngrp = (a.blk_end - a.blk_begin) >> 5;
    // contiguous slice of target groups for this warp
    const int64_t gb = grp0 + (ngrp * warp) / nwarps, ge = grp0 + (ngrp * (warp + 1)) / nwarps;
    if (gb >= ge) return;
    const int64_t tb0 = gb << 5, tb1 = ge << 5;         // target block range of this warp
    const int extra = (a.M - 1 + 31) >> 5;              // groups past the end needed to close the trailing sums
    const int64_t gend = ge + extra;
    const int srcl = (lane - a.M) & 31;

    uint4 cur = __ldg(a.seq + (gb << 5) + lane);
    uint4 nxt = __ldg(a.seq + ((gb + 1) << 5) + lane);
    uint32_t run = 0, prevP = 0;
    for (int64_t g = gb; g < gend; ++g) {
        uint4 nx2 = __ldg(a.seq + ((g + 2) << 5) + lane);          // prefetch distance 2 (buffer is padded)
        uint32_t nw = __shfl_down_sync(FULL, cur.x, 1);
        uint32_t n0 = __shfl_sync(FULL, nxt.x, 0);
        if (lane == 31) nw = n0;
        uint32_t G = block_weight<K>(s_tab, cur, nw);
        // inclusive warp scan of the block sums (mod 2^32 arithmetic; differences are exact)
        uint32_t P = G;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, P, d); if (lane >= d) P += t; }
        P += run;
        run = __shfl_sync(FULL, P, 31);
        uint32_t pa = __shfl_sync(FULL, P, srcl), pb = __shfl_sync(FULL, prevP, srcl);
        uint32_t F = P - (lane >= a.M ? pa : pb);        // sum of the M blocks ending at this one
        prevP = P;
        int64_t tgt = (g << 5) + lane - (a.M - 1);       // block whose windows these M blocks cover
        bool flag = (F > a.thrw) && tgt >= tb0 && tgt < tb1;
        unsigned bal = __ballot_sync(FULL, flag);
        if (bal) {                                       // warp-aggregated atomic append
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.cand_count, (uint32_t)__popc(bal));
            base = __shfl_sync(FULL, base, 0);
            if (flag) {
                uint32_t pos = base + __popc(bal & ((1u << lane) - 1));
                if (pos < a.cand_cap) a.cand[pos] = (uint32_t)tgt;
                atomicOr(a.bitmap + (tgt >> 5), 1u << (tgt & 31));
            }
        }
        cur = nxt; nxt = nx2;
    }
}

// ---------------------------------------------------------------------------------------------
// kgma_prefilter9<K>: the same filter with 10-K k-mers per table lookup instead of 9-K (k = 6: 16 lookups per 64 bases instead
// of 22, k = 7: 22 instead of 32).  The kernel is bound by shared-memory wavefronts (3.5 per random 32-lane gather), so fewer
// gathers is the only lever; a table indexed by 9-mers needs 4^9 entries, 256 KB at one byte each -- more than an SM has.
// Here the LAST base of the 9-mer is ternary {A, C, G|T}: 3 * 4^8 one-byte entries = 192 KB.  The 9-K k-mers that lie in
// the first eight bases are exact; the one k-mer that reaches the ninth base is exact for A and C and bounded by the larger of
// its G and T weights otherwise (N is folded to T in the packed genome).  Entries are weights / step, rounded up, so the filter
// stays a rigorous upper bound; flag when the covering sum of entries exceeds 2^14 / step.
// The table is staged with bulk asynchronous copies (cp.async.bulk, the TMA engine) completing on an mbarrier; the genome
// itself is read once, 128 bits per thread, straight into registers -- nothing is reused, so there is nothing to stage.
struct Filter9Args {
    const uint4   *seq;
    const uint8_t *tab;         // [3 * 65536] entry(9-mer with ternary last base)
    int64_t  blk_begin, blk_end;
    int      M;
    uint32_t thrw;              // flag when the covering sum of entries > thrw  (= floor(2^14 / step))
    uint32_t *cand; uint32_t cand_cap; uint32_t *cand_count; uint32_t *bitmap;
};
constexpr uint32_t TAB9_BYTES = 3u * 65536u;

template <int K>
__device__ __forceinline__ uint32_t block_weight9(const uint8_t *tab, uint4 w, uint32_t nw)
{
    constexpr int STRIDE = 10 - K;                      // k-mers fully contained in one 9-mer
    constexpr int NLOOK = (FBLOCK + STRIDE - 1) / STRIDE;
    const uint32_t W[5] = { w.x, w.y, w.z, w.w, nw };
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < NLOOK; i++) {
        const int b = 2 * i * STRIDE;                   // bit offset of the 9-mer
        const int wi = b >> 5, s = b & 31;
        uint32_t x;
        if (s + 18 <= 32) x = (W[wi] >> s) & 0x3FFFFu;
        else x = __funnelshift_r(W[wi], W[wi + 1], s) & 0x3FFFFu;
        x &= ~((x >> 1) & 0x10000u);                    // ninth base G (10) or T (11) -> class 2 (10)
        sum += tab[x];
    }
    return sum;
}

template <int K>
__global__ void __launch_bounds__(1024, 1) kgma_prefilter9(Filter9Args a)
{
    extern __shared__ __align__(128) uint8_t s_tab9[];
    __shared__ cuda::barrier<cuda::thread_scope_block> bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cuda::device::experimental::fence_proxy_async_shared_cta(); }
    __syncthreads();
    {
        cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
        if (threadIdx.x == 0) {
            constexpr uint32_t CHUNK = 32768;
            for (uint32_t o = 0; o < TAB9_BYTES; o += CHUNK)
                cuda::device::memcpy_async_tx(s_tab9 + o, a.tab + o, cuda::aligned_size_t<16>(CHUNK), bar);
            tok = cuda::device::barrier_arrive_tx(bar, 1, TAB9_BYTES);
        } else tok = bar.arrive();
        bar.wait(std::move(tok));
    }

    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t grp0 = a.blk_begin >> 5, ngrp = (a.blk_end - a.blk_begin) >> 5;
    const int64_t gb = grp0 + (ngrp * warp) / nwarps, ge = grp0 + (ngrp * (warp + 1)) / nwarps;
    if (gb >= ge) return;
    const int64_t tb0 = gb << 5, tb1 = ge << 5;
    const int extra = (a.M - 1 + 31) >> 5;
    const int64_t gend = ge + extra;
    const int srcl = (lane - a.M) & 31;

    uint4 cur = __ldg(a.seq + (gb << 5) + lane);
    uint4 nxt = __ldg(a.seq + ((gb + 1) << 5) + lane);
    uint32_t run = 0, prevP = 0;
    for (int64_t g = gb; g < gend; ++g) {
        uint4 nx2 = __ldg(a.seq + ((g + 2) << 5) + lane);
        uint32_t nw = __shfl_down_sync(FULL, cur.x, 1);
        uint32_t n0 = __shfl_sync(FULL, nxt.x, 0);
        if (lane == 31) nw = n0;
        uint32_t G = block_weight9<K>(s_tab9, cur, nw);
        uint32_t P = G;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, P, d); if (lane >= d) P += t; }
        P += run;
        run = __shfl_sync(FULL, P, 31);
        uint32_t pa = __shfl_sync(FULL, P, srcl), pb = __shfl_sync(FULL, prevP, srcl);
        uint32_t F = P - (lane >= a.M ? pa : pb);
        prevP = P;
        int64_t tgt = (g << 5) + lane - (a.M - 1);
        bool flag = (F > a.thrw) && tgt >= tb0 && tgt < tb1;
        unsigned bal = __ballot_sync(FULL, flag);
        if (bal) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.cand_count, (uint32_t)__popc(bal));
            base = __shfl_sync(FULL, base, 0);
            if (flag) {
                uint32_t pos = base + __popc(bal & ((1u << lane) - 1));
                if (pos < a.cand_cap) a.cand[pos] = (uint32_t)tgt;
                atomicOr(a.bitmap + (tgt >> 5), 1u << (tgt & 31));
            }
        }
        cur = nxt; nxt = nx2;
    }
}

// =============================================================================================
// (2) count-table kernel: one warp per span of consecutive windows
// =============================================================================================
// A span is either one flagged 64-base block of the candidate list (written by the prefilter, read here
// straight from device memory: no host round trip) or, in dense mode, an implicit slice of `span` windows
// of a record.  The warp keeps a 4^k x u16 k-mer count table in shared memory next to the profile's S
// table: it is initialised from the first window of the span (all lanes, packed 16-bit atomics), then one
// k-mer leaves and one enters per step (GenomeMiner.jl:69-77) while Q = sum c^2 and A = sum S[kmer] are
// updated incrementally; D = N^2 Q - 2N A + sum S^2.  Every 64 windows the warp classifies them in
// parallel (D < T by ballot, run minima by warp reduction) and appends run summaries with an atomic.
// The table is returned to zero by clearing the entries of the last window.
struct RecDev {
    long long off;          // global base offset of the record
    long long w_begin, w_end;   // window starts (== loop steps, 0-based) owned by this shard: [w_begin, w_end)
    long long item_base;    // dense mode: index of the record's first item
    long long dist_base;    // do_return_dists: index of step 1 of this record in the D output
};

struct ProfDev {
    long long N2, twoN, sumS2, T, Tlo, Thi, R;    // R = N^2 nk + sum S^2 - Thi (<= 0: cannot be bounded)
    int nk, N;
    int uThi, u_ok;                               // 32-bit pre-test: D < Thi <=> N Q - 2 A < uThi (valid when u_ok)
};

struct EvalArgs {
    const uint32_t *seq;
    const int32_t  *S;              // [C][4^k] reversed-index profile sums
    const uint32_t *cand;           // candidate mode: flagged block ids (null = dense mode); the first n_seed entries are
    const uint32_t *cand_count; uint32_t cand_cap;   // record-start blocks of which only window 0 is evaluated
    const uint32_t *bitmap; uint32_t n_seed;
    unsigned long long *next_item;  // [C] dynamic work counters (spans differ a lot in length: static striding leaves warps idle)
    long long cand_blk_lo, cand_blk_hi;   // candidate mode: only blocks in [lo, hi) (a pipelined scan evaluates record ranges separately)
    long long n_items;              // dense mode: number of implicit items
    const RecDev *recs; int nrec;
    int C, k, span;                 // span: windows per dense item
    int nq; int qlist[MAX_PROFILES];    // the profiles this launch evaluates
    ProfDev prof[MAX_PROFILES];
    kgma_run *runs; uint32_t run_cap; uint32_t *run_count;
    long long *first_D;             // [C][nrec]
    long long *dists; long long dist_stride;
    const uint16_t *cmap; uint32_t xmask;   // strobemer mode: code of every (xmask+1)-entry base pattern (null: the k-mer itself is the code)
};

__device__ __forceinline__ uint32_t kmer_at(const uint32_t *seq, long long gp, uint32_t kmask)
{
    const uint32_t *p = seq + (gp >> 4);
    return __funnelshift_r(__ldg(p), __ldg(p + 1), (int)(gp & 15) * 2) & kmask;
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

__device__ __forceinline__ void emit_run(const EvalArgs &a, int rec, int q, long long tf, long long tl, long long ta,
                                         long long dmin, uint32_t flags)
{
    uint32_t i = atomicAdd(a.run_count, 1u);
    if (i < a.run_cap) {
        kgma_run r; r.record = rec; r.profile = q; r.t_first = tf; r.t_last = tl; r.t_argmin = ta;
        r.D_min = dmin; r.flags = flags; r.reserved = 0;
        a.runs[i] = r;
    }
}

// The round-1 form of the count-table kernel, kept as a cross-check (KGMA_EVAL_KERNEL=serial): the same spans, tables and run
// bookkeeping, but the slide is one instruction stream on lane 0 (64 steps prepared by all lanes, applied by one).
__global__ void __launch_bounds__(512, 1) kgma_eval_serial(EvalArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const int nb = 1 << (2 * a.k);
    const uint32_t kmask = (uint32_t)nb - 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int32_t *sS = reinterpret_cast<int32_t *>(smem_raw);
    const size_t per_warp = (size_t)nb * 2 + 64 * 4 + 64 * 4 + 64 * 8;
    unsigned char *wb = smem_raw + (size_t)nb * 4 + (size_t)wid * per_warp;
    uint2 *qa = reinterpret_cast<uint2 *>(wb);                                // [64] (Q, A) of the current 64 windows
    uint16_t *tab = reinterpret_cast<uint16_t *>(wb + 64 * 8);                // [4^k] counts
    uint32_t *lr = reinterpret_cast<uint32_t *>(tab + nb);                    // [64] byte offsets 2*leaving | 2*entering << 16 of 64 steps
    int32_t *dS = reinterpret_cast<int32_t *>(lr + 64);                       // [64] S[entering] - S[leaving] of those steps
    for (int i = lane; i < nb / 2; i += 32) reinterpret_cast<uint32_t *>(tab)[i] = 0;
    if (nb < 2 && lane == 0) tab[0] = 0;

    long long nitems = a.n_items;
    if (a.cand) { uint32_t c = *a.cand_count; nitems = c < a.cand_cap ? c : a.cand_cap; }

    for (int qi = 0; qi < a.nq; qi++) {
        const int q = a.qlist[qi];
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += blockDim.x) sS[i] = a.S[(size_t)q * nb + i];
        __syncthreads();
        const ProfDev P = a.prof[q];
        const int nk = P.nk;
        for (;;) {
            long long item = 0;
            if (lane == 0) item = (long long)atomicAdd(a.next_item + q, 1ull);
            item = __shfl_sync(FULL, item, 0);
            if (item >= nitems) break;
            // ---- locate the span: record r, first window w0 (== loop step), n windows
            int r; long long w0, n;
            if (a.cand) {
                const long long b = (long long)a.cand[item];
                if (b < a.cand_blk_lo || b >= a.cand_blk_hi) continue;
                const long long gp = b * FBLOCK;
                int lo = 0, hi = a.nrec;                                      // last record with off <= gp
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.recs[mid].off <= gp) lo = mid; else hi = mid; }
                r = lo;
                const RecDev R = a.recs[r];
                const long long b0 = gp - R.off;
                if (b0 < 0) continue;
                if (item < (long long)a.n_seed) { w0 = 0; n = (b0 == 0 && R.w_begin == 0 && R.w_end > 0) ? 1 : 0; }   // window 0 only
                else {
                    // consecutive flagged blocks of one record are evaluated as ONE span by the warp that owns their head;
                    // spans are also cut every 8 blocks so that long flagged stretches still spread over many warps
                    const bool prev_same_rec = b0 >= FBLOCK;
                    const bool prev_set = b > 0 && ((a.bitmap[(b - 1) >> 5] >> ((b - 1) & 31)) & 1u);
                    if (prev_set && prev_same_rec && (b & 7) != 0) continue;
                    long long e = b;
                    while (((e + 1) & 7) != 0 && ((a.bitmap[(e + 1) >> 5] >> ((e + 1) & 31)) & 1u)) e++;
                    w0 = b0 > R.w_begin ? b0 : R.w_begin;
                    const long long we = (e + 1) * FBLOCK - R.off < R.w_end ? (e + 1) * FBLOCK - R.off : R.w_end;
                    n = we - w0;
                }
            } else {
                int lo = 0, hi = a.nrec;                                      // last record with item_base <= item
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.recs[mid].item_base <= item) lo = mid; else hi = mid; }
                r = lo;
                const RecDev R = a.recs[r];
                w0 = R.w_begin + (item - R.item_base) * a.span;
                n = R.w_end - w0 < a.span ? R.w_end - w0 : a.span;
            }
            if (n <= 0) continue;
            const long long gpos = a.recs[r].off + w0;
            const long long dist_base = a.recs[r].dist_base;

            // ---- candidate spans are flagged per 64-base block with the group's maximum weights over a 384-base cover; check this
            //      profile's own bound exactly before paying for the serial slide: a window can only be below thr if
            //      2N * A_w > R (D_w >= N^2 nk - 2N A_w + sum S^2), and A_w of all n windows costs one pass of lookups
            if (a.cand && P.R > 0 && !(w0 == 0 && n == 1)) {
                uint32_t A0 = 0;
                for (int p = lane; p < nk; p += 32) A0 += (uint32_t)sS[kmer_at(a.seq, gpos + p, kmask)];
#pragma unroll
                for (int d = 16; d; d >>= 1) A0 += __shfl_xor_sync(FULL, A0, d);
                long long Aw = A0, Amax = A0;                                  // A of window 0; then 32 windows per round
                for (long long wb0 = 0; wb0 + 1 < n; wb0 += 32) {
                    const long long w = wb0 + lane;                            // step w: window w -> w+1
                    int dlt = 0;
                    if (w + 1 < n) dlt = sS[kmer_at(a.seq, gpos + w + nk, kmask)] - sS[kmer_at(a.seq, gpos + w, kmask)];
                    int pre = dlt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += t; }
                    long long mine = Aw + pre;                                 // A of window w+1
                    if (w + 1 >= n) mine = 0;
#pragma unroll
                    for (int d = 16; d; d >>= 1) { const long long o = __shfl_xor_sync(FULL, mine, d); mine = o > mine ? o : mine; }
                    Amax = mine > Amax ? mine : Amax;
                    Aw += __shfl_sync(FULL, pre, 31);
                }
                if (P.twoN * Amax <= P.R) continue;                            // no window of this span can reach thr for this profile
            }
            // ---- first window of the span: build the table, Q = sum_p c[kmer_p], A = sum_p S[kmer_p]
            for (int p = lane; p < nk; p += 32) {
                const uint32_t km = kmer_at(a.seq, gpos + p, kmask);
                atomicAdd(reinterpret_cast<uint32_t *>(tab) + (km >> 1), 1u << ((km & 1) * 16));
            }
            __syncwarp();
            uint32_t Q = 0, A = 0;                                            // < 2^32: nk <= 65535, nk * max S checked on the host
            for (int p = lane; p < nk; p += 32) {
                const uint32_t km = kmer_at(a.seq, gpos + p, kmask);
                Q += tab[km]; A += (uint32_t)sS[km];
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) { Q += __shfl_xor_sync(FULL, Q, d); A += __shfl_xor_sync(FULL, A, d); }

            bool c_in = false; long long c_tf = 0, c_ta = 0, c_dmin = 0; uint32_t c_fl = 0;   // run carried across 64-window batches
            for (long long s0 = 0; s0 < n; s0 += 64) {
                const int m = (int)(n - s0 < 64 ? n - s0 : 64);
                for (int j = lane; j < 64; j += 32) {                         // step j: window s0+j -> s0+j+1
                    // everything that does not depend on the counts is prepared by all lanes: the table byte offsets of the
                    // leaving / entering k-mer and the change of A; steps past the span's last slide are made no-ops (l == r)
                    const bool slides = j < m && s0 + j + 1 < n;
                    const uint32_t l = slides ? kmer_at(a.seq, gpos + s0 + j, kmask) : 0u;
                    const uint32_t r2 = slides ? kmer_at(a.seq, gpos + s0 + j + nk, kmask) : 0u;
                    lr[j] = (2u * l) | ((2u * r2) << 16);
                    dS[j] = sS[r2] - sS[l];
                }
                __syncwarp();
                if (lane == 0) {
                    // the serial part: a single instruction stream, so every instruction counts
                    const unsigned char *tb = reinterpret_cast<const unsigned char *>(tab);
#pragma unroll 4
                    for (int j = 0; j < 64; j++) {
                        qa[j] = make_uint2(Q, A);
                        const uint32_t cur = lr[j];
                        const uint32_t lo = cur & 0xFFFFu, ro = cur >> 16;
                        if (lo != ro) {                                       // GenomeMiner.jl:69 `if left_ind != right_ind`
                            uint16_t *pl = (uint16_t *)(tb + lo), *pr = (uint16_t *)(tb + ro);
                            const uint32_t cl = *pl, cr = *pr;
                            Q += 2u * (cr - cl) + 2u; A += (uint32_t)dS[j];
                            *pl = (uint16_t)(cl - 1); *pr = (uint16_t)(cr + 1);
                        }
                    }
                }
                __syncwarp();
                // ---- classify the m windows: lanes own windows lane and lane+32
                unsigned bm[2], nm[2]; long long Dm[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int wi = lane + 32 * h;
                    const bool valid = wi < m;
                    const long long t = w0 + s0 + wi;
                    const uint2 v = qa[wi & 63];
                    const long long D = valid ? P.N2 * (long long)v.x - P.twoN * (long long)v.y + P.sumS2 : 0;
                    Dm[h] = D;
                    if (valid && t == 0) a.first_D[(size_t)q * a.nrec + r] = D;      // first window: never compared with thr
                    const bool inloop = valid && t >= 1;
                    if (a.dists && inloop) a.dists[(size_t)q * a.dist_stride + dist_base + t - 1] = D;
                    bm[h] = __ballot_sync(FULL, inloop && D < P.T);             // GenomeMiner.jl:82 `kmerDist < thr`
                    nm[h] = __ballot_sync(FULL, inloop && D >= P.Tlo && D < P.Thi);
                }
                const unsigned long long below = (unsigned long long)bm[0] | ((unsigned long long)bm[1] << 32);
                const unsigned long long near = (unsigned long long)nm[0] | ((unsigned long long)nm[1] << 32);
                const bool more = s0 + 64 < n;                                 // another batch of this span follows
                if (c_in && !(below & 1ull)) {                                 // the carried run ended with the previous batch
                    if (lane == 0) emit_run(a, r, q, c_tf, w0 + s0 - 1, c_ta, c_dmin, c_fl);
                    c_in = false;
                }
                unsigned long long rem = below;
                while (rem) {                                                  // maximal stretches of D < T inside these 64 windows
                    const int b0 = __ffsll((long long)rem) - 1;
                    const unsigned long long sh = rem >> b0;
                    const int len = (~sh == 0ull) ? 64 : (__ffsll((long long)~sh) - 1);
                    const unsigned long long stretch = (len == 64 ? ~0ull : ((1ull << len) - 1)) << b0;
                    const int b1 = b0 + len - 1;
                    long long bestD = LLONG_MAX; int bestI = 64;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int wi = lane + 32 * h;
                        if (wi >= b0 && wi <= b1 && Dm[h] < bestD) { bestD = Dm[h]; bestI = wi; }
                    }
#pragma unroll
                    for (int d = 16; d; d >>= 1) {                             // min D, earliest window on ties
                        const long long oD = __shfl_xor_sync(FULL, bestD, d); const int oI = __shfl_xor_sync(FULL, bestI, d);
                        if (oD < bestD || (oD == bestD && oI < bestI)) { bestD = oD; bestI = oI; }
                    }
                    int ties = 0;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int wi = lane + 32 * h;
                        ties += __popc(__ballot_sync(FULL, wi >= b0 && wi <= b1 && Dm[h] == bestD));
                    }
                    uint32_t fl = (ties > 1 ? KGMA_HIT_ARGMIN_TIE : 0u) | ((near & stretch) ? KGMA_HIT_NEAR_THR : 0u);
                    long long tf = w0 + s0 + b0, ta = w0 + s0 + bestI, dmin = bestD;
                    if (c_in && b0 == 0) {                                     // continues the run carried over from the previous batch
                        fl |= (c_fl & ~KGMA_HIT_ARGMIN_TIE);               // (a tie inside a piece that is not the minimum is no tie of the run)
                        if (c_dmin < dmin) { dmin = c_dmin; ta = c_ta; fl = (fl & ~KGMA_HIT_ARGMIN_TIE) | (c_fl & KGMA_HIT_ARGMIN_TIE); }
                        else if (c_dmin == dmin) { ta = c_ta; fl |= KGMA_HIT_ARGMIN_TIE; }
                        tf = c_tf; c_in = false;
                    } else if (w0 + s0 + b0 == w0 || (w0 == 0 && s0 == 0 && b0 == 1)) fl |= KGMA_RUN_OPEN_LEFT;
                    if (b1 == m - 1 && more) { c_in = true; c_tf = tf; c_ta = ta; c_dmin = dmin; c_fl = fl; }   // may continue
                    else if (lane == 0) emit_run(a, r, q, tf, w0 + s0 + b1, ta, dmin, fl | (b1 == m - 1 ? KGMA_RUN_OPEN_RIGHT : 0u));
                    rem &= ~stretch;
                }
                unsigned long long mk = near & ~below;                          // d >= thr inside the 1e-9 band: reported, never replayed
                while (mk) {
                    const int i = __ffsll((long long)mk) - 1; mk &= mk - 1;
                    if (lane == 0) emit_run(a, r, q, w0 + s0 + i, w0 + s0 + i, w0 + s0 + i,
                                            P.N2 * (long long)qa[i].x - P.twoN * (long long)qa[i].y + P.sumS2, KGMA_RUN_MARKER | KGMA_HIT_NEAR_THR);
                }
                __syncwarp();
            }
            // ---- return the table to zero: only the k-mers of the last window are still counted
            for (int p = lane; p < nk; p += 32) tab[kmer_at(a.seq, gpos + n - 1 + p, kmask)] = 0;
            __syncwarp();
        }
    }
}

// v + (t & mask) with the mask kept as data: written as `if (j >= d) v += t` the compiler re-derives four predicates per batch
__device__ __forceinline__ int tg_and_add(int v, int t, int mask)
{
    int r; asm("{ .reg .b32 x; and.b32 x, %1, %2; add.s32 %0, %3, x; }" : "=r"(r) : "r"(t), "r"(mask), "r"(v)); return r;
}

// Run bookkeeping of one warp over consecutive groups of windows of a span: an open run is carried from group to group so
// that every maximal stretch of D < T is emitted once.
struct RunCarry { bool in; long long tf, ta, dmin; uint32_t fl; };

// Classify a group of `cnt` (<= 32) consecutive windows tb .. tb+cnt-1 of record r, lane i holding the exact D of window
// tb+i (lanes >= cnt: have = false): window 0 is only recorded (GenomeMiner.jl:57: never compared with thr), every other
// one is written out when distances are wanted and tested against T (GenomeMiner.jl:82 `kmerDist < thr`) and the 1e-9 band.
// (kept out of line and fed by value: the slide loop calls it for a handful of batches per million, and must not pay for
//  its registers or for a stack copy of the kernel parameters)
struct ClsArgs {
    kgma_run *runs; uint32_t run_cap; uint32_t *run_count;
    long long *first_D; long long *dists; long long dist_stride; int nrec;
    long long T, Tlo, Thi;
};

__device__ __forceinline__ void emit_run_c(const ClsArgs &a, int rec, int q, long long tf, long long tl, long long ta, long long dmin, uint32_t flags)
{
    uint32_t i = atomicAdd(a.run_count, 1u);
    if (i < a.run_cap) {
        kgma_run r; r.record = rec; r.profile = q; r.t_first = tf; r.t_last = tl; r.t_argmin = ta;
        r.D_min = dmin; r.flags = flags; r.reserved = 0;
        a.runs[i] = r;
    }
}

__device__ __noinline__ RunCarry classify_group(ClsArgs a, int r, int q, long long w0, long long tb, int cnt,
                                                bool have, long long D, bool more, long long dist_base, RunCarry c, int lane)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const ClsArgs &P = a;
    const long long t = tb + lane;
    if (have && t == 0) a.first_D[(size_t)q * a.nrec + r] = D;
    const bool inloop = have && t >= 1;
    if (a.dists && inloop) a.dists[(size_t)q * a.dist_stride + dist_base + t - 1] = D;
    if (!c.in && !__any_sync(FULL, inloop && D < P.Thi)) return c;        // the common case: nothing at or below the threshold band
    const unsigned below = __ballot_sync(FULL, inloop && D < P.T);
    const unsigned near = __ballot_sync(FULL, inloop && D >= P.Tlo && D < P.Thi);
    if (c.in && !(below & 1u)) {                                        // the carried run ended with the previous group
        if (lane == 0) emit_run_c(a, r, q, c.tf, tb - 1, c.ta, c.dmin, c.fl);
        c.in = false;
    }
    if (!(below | near)) return c;
    unsigned rem = below;
    while (rem) {                                                      // maximal stretches of D < T inside this group
        const int b0 = __ffs((int)rem) - 1;
        const unsigned sh = rem >> b0;
        const int len = (~sh == 0u) ? 32 - b0 : (__ffs((int)~sh) - 1);
        const unsigned stretch = (len >= 32 ? ~0u : ((1u << len) - 1)) << b0;
        const int b1 = b0 + len - 1;
        const bool mine = lane >= b0 && lane <= b1;
        long long bestD = mine ? D : LLONG_MAX; int bestI = mine ? lane : 64;
#pragma unroll
        for (int d = 16; d; d >>= 1) {                                 // min D, earliest window on ties
            const long long oD = __shfl_xor_sync(FULL, bestD, d); const int oI = __shfl_xor_sync(FULL, bestI, d);
            if (oD < bestD || (oD == bestD && oI < bestI)) { bestD = oD; bestI = oI; }
        }
        const int ties = __popc(__ballot_sync(FULL, mine && D == bestD));
        uint32_t fl = (ties > 1 ? KGMA_HIT_ARGMIN_TIE : 0u) | ((near & stretch) ? KGMA_HIT_NEAR_THR : 0u);
        long long tf = tb + b0, ta = tb + bestI, dmin = bestD;
        if (c.in && b0 == 0) {                                         // continues the run carried over from the previous group
            fl |= (c.fl & ~KGMA_HIT_ARGMIN_TIE);                       // (a tie inside a piece that is not the minimum is no tie of the run)
            if (c.dmin < dmin) { dmin = c.dmin; ta = c.ta; fl = (fl & ~KGMA_HIT_ARGMIN_TIE) | (c.fl & KGMA_HIT_ARGMIN_TIE); }
            else if (c.dmin == dmin) { ta = c.ta; fl |= KGMA_HIT_ARGMIN_TIE; }
            tf = c.tf; c.in = false;
        } else if (tf == w0 || (w0 == 0 && tf == 1)) fl |= KGMA_RUN_OPEN_LEFT;
        if (b1 == cnt - 1 && more) { c.in = true; c.tf = tf; c.ta = ta; c.dmin = dmin; c.fl = fl; }   // may continue
        else if (lane == 0) emit_run_c(a, r, q, tf, tb + b1, ta, dmin, fl | (b1 == cnt - 1 ? KGMA_RUN_OPEN_RIGHT : 0u));
        rem &= ~stretch;
    }
    unsigned mk = near & ~below;                                        // d >= thr inside the 1e-9 band: reported, never replayed
    while (mk) {
        const int i = __ffs((int)mk) - 1; mk &= mk - 1;
        const long long Di = __shfl_sync(FULL, D, i);
        if (lane == 0) emit_run_c(a, r, q, tb + i, tb + i, tb + i, Di, KGMA_RUN_MARKER | KGMA_HIT_NEAR_THR);
    }
    return c;
}

// One launch per profile: k and the profile's constants are compile-time / direct kernel parameters, so the slide loop
// keeps nothing but its own state in registers.
// STROBE (StrobeGMA!, StrobeGenomeMiner.jl:45-66): the code of a position is the strobemer code of the bases starting there
// (a table lookup: MAP[k-mer]), K is the exponent of the code space (4^(2s) codes), and the table the reference tracks is
// "the nk codes from the window start on plus ONE permanent copy of the code at position nk of the record": its entering code
// is taken at i+ws-k (:53), the last code of the window that is being left.  Here P.nk is that nk (= ws - k) and the extra code
// is added when a span's table is built.
template <int K, int MAXT, bool STROBE = false>
__global__ void __launch_bounds__(MAXT, 1) kgma_eval(EvalArgs a, ProfDev P, int q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    constexpr int nb = 1 << (2 * K);
    constexpr uint32_t kmask = (uint32_t)nb - 1;
    // 16-bit counts.  k <= 6: next to them a small array of one-byte lane stamps, indexed by the low bits of the code (duplicate
    // detection, below; 2 KB at most -- with the 8 KB of counts at k = 6 that is 10 KB per warp, 20 warps per SM, where 32-bit
    // stamped entries allowed 13); k = 7: MATCH.ANY for every batch
    constexpr bool STAMP = K <= 6;
    constexpr int ESH = 1;                                                 // log2(bytes per entry)
    constexpr int NSTAMP = STAMP ? (nb < KGMA_NSTAMP ? nb : KGMA_NSTAMP) : 0;    // stamp slots (bytes) per warp
    constexpr uint32_t SMASK = NSTAMP ? (uint32_t)NSTAMP - 1 : 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t *sS = reinterpret_cast<int32_t *>(smem_raw);
    const uint32_t xmask = STROBE ? a.xmask : kmask;                        // bits of the base pattern a code is derived from
    const size_t map_bytes = STROBE ? (((size_t)xmask + 1) * 2 + 15) & ~(size_t)15 : 0;
    const uint16_t *smap = reinterpret_cast<const uint16_t *>(smem_raw + (size_t)nb * 4);
    unsigned char *tabb = smem_raw + (size_t)nb * 4 + map_bytes + (size_t)wid * (((size_t)nb << ESH) + NSTAMP);   // [4^k] counts (+ stamps) of this warp
    unsigned char *stampb = tabb + ((size_t)nb << ESH);
    auto code_at = [&](long long gp) -> uint32_t { const uint32_t v = kmer_at(a.seq, gp, xmask); return STROBE ? (uint32_t)smap[v] : v; };
    auto cnt_ptr = [&](uint32_t km) { return reinterpret_cast<uint16_t *>(tabb + ((size_t)km << ESH)); };
    for (int i = lane; i < ((nb << ESH) + NSTAMP) / 4; i += 32) reinterpret_cast<uint32_t *>(tabb)[i] = 0;

    long long nitems = a.n_items;
    if (a.cand) { uint32_t c = *a.cand_count; nitems = c < a.cand_cap ? c : a.cand_cap; }

    {
        for (int i = threadIdx.x; i < nb; i += blockDim.x) sS[i] = a.S[(size_t)q * nb + i];
        if (STROBE) for (uint32_t i = threadIdx.x; i <= xmask; i += blockDim.x) const_cast<uint16_t *>(smap)[i] = a.cmap[i];
        __syncthreads();
        const int nk = P.nk;
        ClsArgs ca;
        ca.runs = a.runs; ca.run_cap = a.run_cap; ca.run_count = a.run_count; ca.first_D = a.first_D; ca.dists = a.dists;
        ca.dist_stride = a.dist_stride; ca.nrec = a.nrec; ca.T = P.T; ca.Tlo = P.Tlo; ca.Thi = P.Thi;
        for (;;) {
            long long item = 0;
            if (lane == 0) item = (long long)atomicAdd(a.next_item + q, 1ull);
            item = __shfl_sync(FULL, item, 0);
            if (item >= nitems) break;
            // ---- locate the span: record r, first window w0 (== loop step), n windows
            int r; long long w0, n;
            if (a.cand) {
                const long long b = (long long)a.cand[item];
                if (b < a.cand_blk_lo || b >= a.cand_blk_hi) continue;
                const long long gp = b * FBLOCK;
                int lo = 0, hi = a.nrec;                                      // last record with off <= gp
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.recs[mid].off <= gp) lo = mid; else hi = mid; }
                r = lo;
                const RecDev R = a.recs[r];
                const long long b0 = gp - R.off;
                if (b0 < 0) continue;
                if (item < (long long)a.n_seed) { w0 = 0; n = (b0 == 0 && R.w_begin == 0 && R.w_end > 0) ? 1 : 0; }   // window 0 only
                else {
                    // consecutive flagged blocks of one record are evaluated as ONE span by the warp that owns their head;
                    // spans are also cut every 8 blocks so that long flagged stretches still spread over many warps
                    const bool prev_same_rec = b0 >= FBLOCK;
                    const bool prev_set = b > 0 && ((a.bitmap[(b - 1) >> 5] >> ((b - 1) & 31)) & 1u);
                    if (prev_set && prev_same_rec && (b & 7) != 0) continue;
                    long long e = b;
                    while (((e + 1) & 7) != 0 && ((a.bitmap[(e + 1) >> 5] >> ((e + 1) & 31)) & 1u)) e++;
                    w0 = b0 > R.w_begin ? b0 : R.w_begin;
                    const long long we = (e + 1) * FBLOCK - R.off < R.w_end ? (e + 1) * FBLOCK - R.off : R.w_end;
                    n = we - w0;
                }
            } else {
                int lo = 0, hi = a.nrec;                                      // last record with item_base <= item
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.recs[mid].item_base <= item) lo = mid; else hi = mid; }
                r = lo;
                const RecDev R = a.recs[r];
                w0 = R.w_begin + (item - R.item_base) * a.span;
                n = R.w_end - w0 < a.span ? R.w_end - w0 : a.span;
            }
            if (n <= 0) continue;
            const long long gpos = a.recs[r].off + w0;
            const long long dist_base = a.recs[r].dist_base;

            // ---- candidate spans are flagged per 64-base block with the group's maximum weights over a 384-base cover; check this
            //      profile's own bound exactly before paying for the table: a window can only be below thr if
            //      2N * A_w > R (D_w >= N^2 nk - 2N A_w + sum S^2), and A_w of all n windows costs one pass of lookups
            if (a.cand && P.R > 0 && !(w0 == 0 && n == 1)) {
                uint32_t A0 = 0;
                for (int p = lane; p < nk; p += 32) A0 += (uint32_t)sS[kmer_at(a.seq, gpos + p, kmask)];
#pragma unroll
                for (int d = 16; d; d >>= 1) A0 += __shfl_xor_sync(FULL, A0, d);
                long long Aw = A0, Amax = A0;                                  // A of window 0; then 32 windows per round
                for (long long wb0 = 0; wb0 + 1 < n; wb0 += 32) {
                    const long long w = wb0 + lane;                            // step w: window w -> w+1
                    int dlt = 0;
                    if (w + 1 < n) dlt = sS[kmer_at(a.seq, gpos + w + nk, kmask)] - sS[kmer_at(a.seq, gpos + w, kmask)];
                    int pre = dlt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += t; }
                    long long mine = Aw + pre;                                 // A of window w+1
                    if (w + 1 >= n) mine = 0;
#pragma unroll
                    for (int d = 16; d; d >>= 1) { const long long o = __shfl_xor_sync(FULL, mine, d); mine = o > mine ? o : mine; }
                    Amax = mine > Amax ? mine : Amax;
                    Aw += __shfl_sync(FULL, pre, 31);
                }
                if (P.twoN * Amax <= P.R) continue;                            // no window of this span can reach thr for this profile
            }
            // ---- first window of the span: build the table, Q = sum_p c[kmer_p], A = sum_p S[kmer_p]
            // (strobemer mode: the permanent extra code counts as position nk of the window)
            const uint32_t zx = STROBE ? code_at(a.recs[r].off + nk) : 0u;
            const int nel = STROBE ? nk + 1 : nk;
            for (int p = lane; p < nel; p += 32) {
                const uint32_t km = (STROBE && p == nk) ? zx : code_at(gpos + p);
                atomicAdd(reinterpret_cast<uint32_t *>(tabb) + (km >> 1), 1u << ((km & 1) * 16));
            }
            __syncwarp();
            uint32_t Qb = 0, Ab = 0;                                          // < 2^32: nk <= 65535, nk * max S checked on the host
            for (int p = lane; p < nel; p += 32) {
                const uint32_t km = (STROBE && p == nk) ? zx : code_at(gpos + p);
                Qb += *cnt_ptr(km); Ab += (uint32_t)sS[km];
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) { Qb += __shfl_xor_sync(FULL, Qb, d); Ab += __shfl_xor_sync(FULL, Ab, d); }

            RunCarry carry; carry.in = false; carry.tf = carry.ta = carry.dmin = 0; carry.fl = 0;
            // window w0 itself
            carry = classify_group(ca, r, q, w0, w0, 1, lane == 0, P.N2 * (long long)Qb - P.twoN * (long long)Ab + P.sumS2, n > 1, dist_base, carry, lane);

            // ---- the slide, 16 steps at a time with every lane busy (GenomeMiner.jl:60-87).  Lanes 0-15 hold the k-mer LEAVING the
            //      window at steps s..s+15, lanes 16-31 the one ENTERING.  The count a step sees is the table's count at the start
            //      of the batch plus the enters minus the leaves of the same k-mer at earlier steps of this batch, which one
            //      MATCH.ANY over the 32 events yields for every lane at once; Q and A then follow from a 16-wide prefix sum of
            //      2(c_r - c_l) + 2 and S[r] - S[l] (lower half-warp scans Q, upper half-warp scans A, same instructions).  The
            //      table receives one store per distinct k-mer whose count changed over the batch.
            //      Windows are pre-tested in 32 bits: D < X  <=>  N (N Q - 2 A) < X - sum S^2  <=>  N Q - 2 A < ceil((X - sum S^2) / N),
            //      exact; the 64-bit D and the run bookkeeping only happen for a batch that has a window under the band's upper edge.
            const int j = lane & 15; const bool ent = lane >= 16;
            const long long pos0 = gpos + j + (ent ? nk : 0);                  // position of my event's k-mer at s = 0
            const uint32_t *wp = a.seq + (pos0 >> 4); const int sh = (int)(pos0 & 15) * 2;   // 16 steps = one packed word: the shift never changes
            uint32_t wlo = __ldg(wp), whi = __ldg(wp + 1);
            const unsigned ltE = (((1u << j) - 1u) << 16), ltL = (1u << j) - 1u;   // enter / leave events of earlier steps
            const int scan1 = j >= 1 ? -1 : 0, scan2 = j >= 2 ? -1 : 0, scan4 = j >= 4 ? -1 : 0, scan8 = j >= 8 ? -1 : 0;
            const unsigned lt_lane = (1u << lane) - 1u;
            const int Nn = P.N, uHi = P.uThi;
            const int ni = (int)n;                                             // spans are at most 65536 windows
            int s = 0, widx = 2;
            auto batch = [&](auto full_c, auto all_c) {
                constexpr bool FB = decltype(full_c)::value;                   // all 16 steps of the batch are real slides
                constexpr bool ALL = decltype(all_c)::value;                   // every window goes through the 64-bit path (do_return_dists)
                uint32_t x = __funnelshift_r(wlo, whi, sh) & xmask;
                if (STROBE) x = smap[x];
                wlo = whi; whi = __ldg(wp + widx); widx++;
                const int left = FB ? 16 : ni - 1 - s;
                const bool valid = FB || j < left;
                // Which lanes hold the same k-mer?  Nearly always none do (32 events among 4^k values), and MATCH.ANY costs about
                // two cycles per distinct value.  So every lane writes its lane number into the stamp slot of its code and reads the
                // slot back: a lane that finds another lane's stamp shares its slot -- its k-mer, or one with the same low bits --
                // with that lane, and only a batch in which some lane does pays for the MATCH (which then finds the real duplicates).
                unsigned m = 1u << lane;
                uint32_t c0;
                if (STAMP) {
                    if (valid) stampb[x & SMASK] = (unsigned char)lane;
                    __syncwarp();
                    c0 = *cnt_ptr(x);
                    const uint32_t stp = stampb[x & SMASK];
                    if (__any_sync(FULL, valid && stp != (uint32_t)lane))
                        m = __match_any_sync(FULL, valid ? x : (0x80000000u | (unsigned)lane));
                } else {
                    m = __match_any_sync(FULL, valid ? x : (0x80000000u | (unsigned)lane));
                    c0 = *cnt_ptr(x);
                }
                if (!FB) m &= ((1u << left) - 1u) * 0x10001u;
                const int cnt = (int)c0 + __popc(m & ltE) - __popc(m & ltL);
                const int Sx = sS[x];
                const int o = __shfl_xor_sync(FULL, ent ? cnt : Sx, 16);       // lower lanes receive c_r, upper lanes S[l]
                int v = ent ? (Sx - o) : (2 * (o - cnt) + 2);                  // S[r] - S[l]  /  2(c_r - c_l) + 2
                if (((m >> (lane ^ 16)) & 1u) || !valid) v = 0;                // GenomeMiner.jl:69 `if left_ind != right_ind`
                v = tg_and_add(v, __shfl_up_sync(FULL, v, 1, 16), scan1);
                v = tg_and_add(v, __shfl_up_sync(FULL, v, 2, 16), scan2);
                v = tg_and_add(v, __shfl_up_sync(FULL, v, 4, 16), scan4);
                v = tg_and_add(v, __shfl_up_sync(FULL, v, 8, 16), scan8);
                const int ov = __shfl_xor_sync(FULL, v, 16);
                const uint32_t Qw = Qb + (uint32_t)(ent ? ov : v), Aw = Ab + (uint32_t)(ent ? v : ov);   // window s+j+1
                const int u = Nn * (int)Qw - 2 * (int)Aw;
                if (ALL || carry.in || __any_sync(FULL, valid && u < uHi)) {
                    const long long D = P.N2 * (long long)Qw - P.twoN * (long long)Aw + P.sumS2;
                    carry = classify_group(ca, r, q, w0, w0 + s + 1, left, !ent && valid, D, s + 1 + left < ni, dist_base, carry, lane);
                }
                Qb += (uint32_t)__shfl_sync(FULL, v, 15); Ab += (uint32_t)__shfl_sync(FULL, v, 31);
                // one store per distinct k-mer (by the lowest lane holding it): count at the start of the batch + enters - leaves
                if (valid && !(m & lt_lane)) *cnt_ptr(x) = (uint16_t)((int)c0 + __popc(m & 0xFFFF0000u) - __popc(m & 0x0000FFFFu));
                __syncwarp();
            };
            if (a.dists != nullptr || !P.u_ok) {
                for (; s + 16 < ni; s += 16) batch(std::true_type{}, std::true_type{});
                if (s + 1 < ni) batch(std::false_type{}, std::true_type{});
            } else {
                for (; s + 16 < ni; s += 16) batch(std::true_type{}, std::false_type{});
                if (s + 1 < ni) batch(std::false_type{}, std::false_type{});
            }
            // ---- return the table to zero: only the k-mers of the last window are still counted
            for (int p = lane; p < nk; p += 32) *cnt_ptr(code_at(gpos + n - 1 + p)) = 0;
            if (STROBE && lane == 0) *cnt_ptr(zx) = 0;
            __syncwarp();
        }
    }
}

// =============================================================================================
// synthetic genome generator (bench / tests): base(p) = splitmix64(seed ^ p) & 3
// =============================================================================================
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void kgma_synth(uint32_t *seq2, int64_t nwords, uint64_t seed)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) w |= (uint32_t)(splitmix64(seed ^ (uint64_t)(i * 16 + j)) & 3) << (2 * j);
        seq2[i] = w;
    }
}

// =============================================================================================
// host side
// =============================================================================================
int dev_scratch(kgma_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->d_scratch_bytes) {
        if (ctx->d_scratch) cudaFree(ctx->d_scratch);
        ctx->d_scratch = nullptr; ctx->d_scratch_bytes = 0; ctx->tab_sig = 0;
        size_t nb = std::max(bytes, (size_t)1 << 20);
        KGMA_CUDA(ctx, cudaMalloc(&ctx->d_scratch, nb));
        ctx->d_scratch_bytes = nb;
    }
    *out = ctx->d_scratch;
    return KGMA_OK;
}

int host_scratch(kgma_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->h_scratch_bytes) {
        if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
        ctx->h_scratch = nullptr; ctx->h_scratch_bytes = 0;
        size_t nb = std::max(bytes, (size_t)1 << 20);
        KGMA_CUDA(ctx, cudaHostAlloc(&ctx->h_scratch, nb, cudaHostAllocDefault));
        ctx->h_scratch_bytes = nb;
    }
    *out = ctx->h_scratch;
    return KGMA_OK;
}

// exact ceil(x * m) for a double x >= 0 and integer m > 0, via 128-bit arithmetic on the binary expansion of x
static bool ceil_mul_exact(double x, long long m, long long *out)
{
    if (!(x >= 0) || !std::isfinite(x)) return false;
    if (x == 0) { *out = 0; return true; }
    int e; double fr = std::frexp(x, &e);                // x = fr * 2^e, fr in [0.5,1)
    unsigned long long mant = (unsigned long long)std::ldexp(fr, 53);   // exact 53-bit integer
    int sh = e - 53;                                     // x = mant * 2^sh
    unsigned __int128 prod = (unsigned __int128)mant * (unsigned __int128)m;
    unsigned __int128 res;
    if (sh >= 0) { if (sh > 20) return false; res = prod << sh; }
    else {
        int s = -sh;
        if (s >= 127) { res = 1; }
        else { unsigned __int128 q = prod >> s, rem = prod & (((unsigned __int128)1 << s) - 1); res = q + (rem ? 1 : 0); }
    }
    if (res > (unsigned __int128)LLONG_MAX / 4) return false;
    *out = (long long)res;
    return true;
}

int build_proftab(kgma_ctx *ctx, const kgma_profile &p, ProfTab &t, const kgma_scan_params *P)
{
    if (p.k < 1 || p.k > MAX_K) return set_err(ctx, KGMA_E_UNSUPPORTED, "k = %d is outside the supported range 1..%d", p.k, MAX_K);
    const bool strobe = P && P->mode == KGMA_MODE_STROBE;
    if (strobe) {
        const int sl = P->strobe_s, a = P->strobe_w_min, b = P->strobe_w_max;
        if (sl < 1 || sl > 3 || a < 1 || b < a || P->strobe_q < 1 || b + sl - 1 != p.k)
            return set_err(ctx, KGMA_E_ARG, "strobemer parameters s=%d w_min=%d w_max=%d q=%d do not fit profile.k=%d (k = w_max + s - 1, s <= 3)", sl, a, b, P->strobe_q, p.k);
    }
    t.strobe = strobe; t.kb = strobe ? 2 * P->strobe_s : p.k;
    if (!p.S || p.n_refs <= 0) return set_err(ctx, KGMA_E_ARG, "profile needs integer sums S and n_refs > 0");
    if (p.k >= p.window) return set_err(ctx, KGMA_E_WINDOW, "the average reference sequence length %lld exceeds/is equal to the chosen kmer length %d. please reduce k. ", (long long)p.window, p.k);
    if (p.window - p.k + 1 > 60000) return set_err(ctx, KGMA_E_UNSUPPORTED, "window %lld too large for 16-bit count tables", (long long)p.window);
    if (!(p.thr >= 0) || !std::isfinite(p.thr)) return set_err(ctx, KGMA_E_ARG, "threshold must be finite and >= 0");
    t.k = p.k; t.ws = p.window; t.nk = p.window - p.k + 1; t.N = p.n_refs; t.thr = p.thr;
    size_t nb = (size_t)1 << (2 * t.kb);
    t.S_rev.assign(nb, 0);
    __int128 s2 = 0;
    for (size_t c = 0; c < nb; c++) {
        if (p.S[c] < 0) return set_err(ctx, KGMA_E_ARG, "negative k-mer sum in profile");
        t.S_rev[strobe ? c : rev_kmer((uint32_t)c, p.k)] = p.S[c];
        s2 += (__int128)p.S[c] * p.S[c];
    }
    t.N2 = (int64_t)p.n_refs * p.n_refs; t.twoN = 2LL * p.n_refs;
    {   // the count-table kernel keeps Q = sum c^2 <= nk^2 and A = sum S[kmer] <= nk * max S in 32 bits
        int64_t maxS = 0; for (size_t c = 0; c < nb; c++) maxS = std::max<int64_t>(maxS, p.S[c]);
        if ((__int128)t.nk * maxS >= ((__int128)1 << 32) || t.nk > 65535)
            return set_err(ctx, KGMA_E_UNSUPPORTED, "profile too large for the 32-bit window sums");
    }
    // magnitude check: D <= N^2 nk^2 + sumS2 + 2N*nk*maxS must stay far below 2^62
    __int128 bound = (__int128)t.N2 * t.nk * t.nk * 2 + s2 * 2;
    if (bound > ((__int128)1 << 61)) return set_err(ctx, KGMA_E_UNSUPPORTED, "profile too large for 64-bit exact distances");
    t.sumS2 = (int64_t)s2;
    long long scale = 2LL * p.k * t.N2;                  // d = D / (2 k N^2)
    t.denom = (double)scale;
    if (!ceil_mul_exact(p.thr, scale, (long long *)&t.T)) return set_err(ctx, KGMA_E_UNSUPPORTED, "threshold out of range");
    // 1e-9 relative band around thr (north_star: hits that close to the threshold are reported separately)
    double lo = p.thr * (1.0 - 1e-9), hi = p.thr * (1.0 + 1e-9);
    long long a = 0, b = 0;
    ceil_mul_exact(lo, scale, &a); ceil_mul_exact(hi, scale, &b);
    t.Tlo = a; t.Thi = std::max<long long>(b, t.T);
    {   // fingerprint of everything the prefilter table depends on (64-bit words of S, then the scalars)
        uint64_t h = 1469598103934665603ull;
        auto mixw = [&](uint64_t w) { h ^= w; h *= 1099511628211ull; h ^= h >> 29; };
        for (size_t c = 0; c + 1 < nb; c += 2) mixw(((uint64_t)(uint32_t)t.S_rev[c] << 32) | (uint32_t)t.S_rev[c + 1]);
        if (nb & 1) mixw((uint32_t)t.S_rev[nb - 1]);
        mixw((uint64_t)t.Thi); mixw((uint64_t)t.N); mixw((uint64_t)t.nk); mixw((uint64_t)t.k); mixw((uint64_t)t.kb);
        t.hash = h;
    }
    return KGMA_OK;
}

int dev_genome_prepare(kgma_ctx *ctx, kgma_genome *g, bool need_mask)
{
    int64_t need = (g->G + TAIL_PAD + 4095) / 4096 * 4096;   // same rounding as the host planes (genome_reserve / kgma_genome_synth)
    if (ctx->d_cap_bases < need) {
        // grow only: cudaFree / cudaMalloc of a genome-sized buffer cost 0.1-0.4 s, far more than a scan, so a context keeps
        // the planes of the largest genome it has seen and merely invalidates their contents when another genome arrives
        if (ctx->d_seq2) cudaFree(ctx->d_seq2);
        if (ctx->d_mask) cudaFree(ctx->d_mask);
        ctx->d_seq2 = ctx->d_mask = nullptr; ctx->d_cap_bases = 0; ctx->dg_uid = 0;
        KGMA_CUDA(ctx, cudaMalloc(&ctx->d_seq2, (size_t)need / 4));
        ctx->d_cap_bases = need;
    }
    if (ctx->dg_uid != g->uid) {
        ctx->dg_uid = g->uid;
        ctx->d_seq_valid = ctx->d_mask_valid = false; ctx->d_valid_lo = ctx->d_valid_hi = 0;
        ctx->d_have_lo = ctx->d_have_hi = 0;
    }
    if (need_mask && !ctx->d_mask) { KGMA_CUDA(ctx, cudaMalloc(&ctx->d_mask, (size_t)ctx->d_cap_bases / 8)); ctx->d_mask_valid = false; }
    return KGMA_OK;
}

// ---------------------------------------------------------------------------------------------
struct ScanPlan {
    int C = 0, k = 0; int64_t maxws = 0, maxnk = 0;
    int kb = 0;                    // code space 4^kb (= k except in strobemer mode)
    bool cluster = false, strobe = false;
    std::vector<uint16_t> cmap;    // strobemer mode: packed base pattern of k bases (first base in the low bits) -> code
    std::vector<ProfTab> tabs;
    // per record: number of loop steps (0 = record not scanned) — GenomeMiner.jl:60 / OmnGenomeMiner.jl:89
    std::vector<int64_t> steps;
};

static int make_plan(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int C, const kgma_scan_params &P, ScanPlan &pl)
{
    if (!profiles || C < 1 || C > MAX_PROFILES) return set_err(ctx, KGMA_E_ARG, "n_profiles must be 1..%d", MAX_PROFILES);
    if ((P.mode == KGMA_MODE_SINGLE || P.mode == KGMA_MODE_STROBE) && C != 1) return set_err(ctx, KGMA_E_ARG, "single / strobemer mode takes exactly one profile");
    if (P.mode != KGMA_MODE_SINGLE && P.mode != KGMA_MODE_CLUSTER && P.mode != KGMA_MODE_STROBE) return set_err(ctx, KGMA_E_ARG, "unknown mode %d", P.mode);
    pl.C = C; pl.k = profiles[0].k; pl.cluster = (P.mode == KGMA_MODE_CLUSTER); pl.strobe = (P.mode == KGMA_MODE_STROBE);
    if (pl.strobe && (P.only_record >= 0 || (P.flags & KGMA_F_TIE_OPEN))) return set_err(ctx, KGMA_E_UNSUPPORTED, "strobemer mode scans whole genomes with the default tie rule");
    pl.tabs.resize(C);
    for (int q = 0; q < C; q++) {
        if (profiles[q].k != pl.k) return set_err(ctx, KGMA_E_ARG, "all profiles must share k");
        int rc = build_proftab(ctx, profiles[q], pl.tabs[q], &P);
        if (rc) return rc;
        pl.maxws = std::max(pl.maxws, pl.tabs[q].ws);
    }
    pl.kb = pl.tabs[0].kb;
    pl.maxnk = pl.maxws - pl.k + 1;
    if (pl.strobe) {
        // get_strobe_2_mer (Strobemers.jl:45-65) for every pattern of k bases: the first strobe is bases 1..s; the running
        // minimum of randstrobe_score = (as_UInt(first) + as_UInt(candidate)) % q starts at `2 << 63` == 0 and is updated on
        // `<=`, so the second strobe starts at the LAST i in w_min..w_max whose score is 0, else at w_min; the code is
        // as_UInt(first * second), first base most significant.
        const int sl = P.strobe_s, a = P.strobe_w_min, b = P.strobe_w_max, qq = P.strobe_q, k = pl.k;
        pl.cmap.assign((size_t)1 << (2 * k), 0);
        for (uint32_t pat = 0; pat < pl.cmap.size(); pat++) {
            auto smer = [&](int i1) { uint32_t v = 0; for (int j = 0; j < sl; j++) v = (v << 2) | ((pat >> (2 * (i1 - 1 + j))) & 3u); return v; };
            const uint32_t f = smer(1);
            int min_ind = a;
            for (int i = a; i <= b; i++) if ((f + smer(i)) % (uint32_t)qq == 0) min_ind = i;
            pl.cmap[pat] = (uint16_t)((f << (2 * sl)) | smer(min_ind));
        }
    }
    if (pl.cluster && pl.k < 2) return set_err(ctx, KGMA_E_UNSUPPORTED, "cluster mode needs k >= 2 (the reference indexes past the record end for k = 1)");
    int nr = (int)g->recs.size();
    pl.steps.assign(nr, 0);
    for (int r = 0; r < nr; r++) {
        if (P.only_record >= 0 && r != P.only_record) continue;
        int64_t L = g->recs[r].len;
        int64_t st = pl.cluster ? (L - pl.maxws - pl.k + 2)       // view(seq, k:L-maxws+1)  OmnGenomeMiner.jl:89
                   : pl.strobe  ? (L - pl.maxws - 1)                // 1:(L-ws-1)              StrobeGenomeMiner.jl:45
                                : (L - pl.maxws);                   // zip(k:L-ws+k-1, ws+1:L) GenomeMiner.jl:60
        pl.steps[r] = std::max<int64_t>(0, st);
    }
    return KGMA_OK;
}

// Fixed-point prefilter table for a group of profiles (weights = max over the group).  Returns false when some profile
// cannot be filtered at all (R <= 0: even a window sharing no k-mer with the profile could be below thr).
// *load = expected covering sum of a uniformly random sequence as a fraction of the flag threshold: the closer to 1,
// the more blocks survive the filter.
static bool build_filter_table(const ScanPlan &pl, const std::vector<int> &group, kgma_ctx::FTab &ft, bool nine)
{
    const int k = pl.k; const size_t nb = (size_t)1 << (2 * k);
    if (k > 8) return false;
    int &M = ft.M;
    M = (int)((pl.maxnk - 1 + FBLOCK - 1) / FBLOCK) + 1;
    if (M > 32 || M < 1) return false;
    std::vector<uint32_t> W(nb, 0);
    for (int q : group) {
        const ProfTab &t = pl.tabs[(size_t)q];
        // candidate  <=>  2N * A > R,  R = N^2 nk + sumS2 - Thi.  Thi (>= T) is the upper edge of the 1e-9 band around
        // thr, so that every window the reference's Float64 accumulator could still see below thr is evaluated and reported.
        __int128 R = (__int128)t.N2 * t.nk + t.sumS2 - t.Thi;
        if (R <= 0) return false;
        for (size_t i = 0; i < nb; i++) {
            if (!t.S_rev[i]) continue;
            __int128 num = ((__int128)t.S_rev[i] * t.twoN) << WFRAC;
            __int128 w = (num + R - 1) / R;                     // ceil
            uint32_t wc = (w > (__int128)WCLAMP) ? WCLAMP : (uint32_t)w;
            W[i] = std::max(W[i], wc);
        }
    }
    const uint32_t kmask = (uint32_t)nb - 1;
    ft.nine = nine;
    if (!nine) {
        const int nper = 9 - k;
        ft.tab.resize(65536);
        double sum = 0;
        for (uint32_t x = 0; x < 65536; x++) {
            uint32_t v = 0;
            for (int j = 0; j < nper; j++) v += W[(x >> (2 * j)) & kmask];
            ft.tab[x] = (uint16_t)std::min<uint32_t>(v, 65535u);
            sum += ft.tab[x];
        }
        const int nlook = (FBLOCK + nper - 1) / nper;
        ft.load = sum / 65536.0 * nlook * M / (double)(1u << WFRAC);
        ft.step = 1; ft.thrw = 1u << WFRAC;
        return true;
    }
    // 9-mer table with a ternary last base (kgma_prefilter9): the k-mers at offsets 0 .. 8-k lie inside the first eight bases,
    // the one at offset 9-k ends on the ninth base; class 2 of the ninth base stands for G or T
    if (k > 9) return false;
    const int nin = 9 - k;                                       // exact k-mers per entry (the tail k-mer comes on top)
    std::vector<uint32_t> raw((size_t)TAB9_BYTES);
    uint32_t maxv = 0;
    for (uint32_t c = 0; c < 3; c++)
        for (uint32_t x8 = 0; x8 < 65536; x8++) {
            uint32_t v = 0;
            for (int j = 0; j < nin; j++) v += W[(x8 >> (2 * j)) & kmask];
            auto tail = [&](uint32_t b9) { return W[(((x8 | (b9 << 16)) >> (2 * nin)) & kmask)]; };
            v += c < 2 ? tail(c) : std::max(tail(2), tail(3));
            raw[(size_t)c * 65536 + x8] = v;
            maxv = std::max(maxv, v);
        }
    const uint32_t step = std::max<uint32_t>(1, (maxv + 254) / 255);
    ft.tab9.resize(TAB9_BYTES);
    double sum = 0, wsum = 0;
    for (size_t i = 0; i < raw.size(); i++) {
        const uint32_t e = (raw[i] + step - 1) / step;           // rounded up: still an upper bound
        ft.tab9[i] = (uint8_t)e;
        const double pw = i < 131072 ? 1.0 : 2.0;                // a random ninth base falls into class 2 twice as often
        sum += pw * e; wsum += pw;
    }
    const int nlook = (FBLOCK + (10 - k) - 1) / (10 - k);
    ft.step = step; ft.thrw = (1u << WFRAC) / step;              // sum * step > 2^14  <=>  sum > floor(2^14 / step)
    ft.load = sum / wsum * step * nlook * M / (double)(1u << WFRAC);
    return true;
}

template <int K> static void launch_filter(const FilterArgs &fa, int grid, cudaStream_t st)
{
    cudaFuncSetAttribute(kgma_prefilter<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 * 2);
    kgma_prefilter<K><<<grid, 1024, 65536 * 2, st>>>(fa);
}

static void launch_filter_k(int k, const FilterArgs &fa, int grid, cudaStream_t st)
{
    switch (k) {
    case 1: launch_filter<1>(fa, grid, st); break; case 2: launch_filter<2>(fa, grid, st); break;
    case 3: launch_filter<3>(fa, grid, st); break; case 4: launch_filter<4>(fa, grid, st); break;
    case 5: launch_filter<5>(fa, grid, st); break; case 6: launch_filter<6>(fa, grid, st); break;
    case 7: launch_filter<7>(fa, grid, st); break; default: launch_filter<8>(fa, grid, st); break;
    }
}

template <int K> static void launch_eval_strobe(const EvalArgs &ea, const ProfDev &P, int q, int grid, int threads, size_t smem, cudaStream_t st)
{
    if (threads > 512) {
        cudaFuncSetAttribute(kgma_eval<K, 640, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kgma_eval<K, 640, true><<<grid, threads, smem, st>>>(ea, P, q);
    } else {
        cudaFuncSetAttribute(kgma_eval<K, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kgma_eval<K, 512, true><<<grid, threads, smem, st>>>(ea, P, q);
    }
}

template <int K> static void launch_eval_t(const EvalArgs &ea, const ProfDev &P, int q, int grid, int threads, size_t smem, cudaStream_t st)
{
    if (threads > 512) {       // more than 16 warps per CTA: the 640-thread build, 96 registers (672 / 736 threads make ptxas drop to 80 and spill)
        cudaFuncSetAttribute(kgma_eval<K, 640>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kgma_eval<K, 640><<<grid, threads, smem, st>>>(ea, P, q);
    } else {
        cudaFuncSetAttribute(kgma_eval<K, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kgma_eval<K, 512><<<grid, threads, smem, st>>>(ea, P, q);
    }
}

static void launch_eval_k(int k, const EvalArgs &ea, const ProfDev &P, int q, int grid, int threads, size_t smem, cudaStream_t st)
{
    if (ea.cmap) {                 // strobemer codes: 4^(2s) bins, s = 1..3
        if (k == 2) launch_eval_strobe<2>(ea, P, q, grid, threads, smem, st);
        else if (k == 4) launch_eval_strobe<4>(ea, P, q, grid, threads, smem, st);
        else launch_eval_strobe<6>(ea, P, q, grid, threads, smem, st);
        return;
    }
    switch (k) {
    case 1: launch_eval_t<1>(ea, P, q, grid, threads, smem, st); break; case 2: launch_eval_t<2>(ea, P, q, grid, threads, smem, st); break;
    case 3: launch_eval_t<3>(ea, P, q, grid, threads, smem, st); break; case 4: launch_eval_t<4>(ea, P, q, grid, threads, smem, st); break;
    case 5: launch_eval_t<5>(ea, P, q, grid, threads, smem, st); break; case 6: launch_eval_t<6>(ea, P, q, grid, threads, smem, st); break;
    default: launch_eval_t<7>(ea, P, q, grid, threads, smem, st); break;
    }
}

template <int K> static void launch_filter9(const Filter9Args &fa, int grid, cudaStream_t st)
{
    cudaFuncSetAttribute(kgma_prefilter9<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB9_BYTES);
    kgma_prefilter9<K><<<grid, 1024, TAB9_BYTES, st>>>(fa);
}

static void launch_filter9_k(int k, const Filter9Args &fa, int grid, cudaStream_t st)
{
    switch (k) {
    case 1: launch_filter9<1>(fa, grid, st); break; case 2: launch_filter9<2>(fa, grid, st); break;
    case 3: launch_filter9<3>(fa, grid, st); break; case 4: launch_filter9<4>(fa, grid, st); break;
    case 5: launch_filter9<5>(fa, grid, st); break; case 6: launch_filter9<6>(fa, grid, st); break;
    default: launch_filter9<7>(fa, grid, st); break;
    }
}

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// warps per CTA and dynamic shared memory of kgma_eval for this k
static int eval_shape(const kgma_ctx *ctx, int k, bool serial, int *warps_out, size_t *smem_out, size_t map_bytes = 0)
{
    const size_t nb = (size_t)1 << (2 * k);
    map_bytes = (map_bytes + 15) & ~(size_t)15;                    // strobemer mode: the code map sits between S and the count tables
    // one 4^k x u16 count table per warp, next to the profile's S table (the serial kernel adds its 64-step staging arrays)
    const size_t per_warp = serial ? nb * 2 + 64 * 4 + 64 * 4 + 64 * 8 : nb * 2 + (k <= 6 ? std::min<size_t>(nb, KGMA_NSTAMP) : 0);
    if (ctx->smem_optin < nb * 4 + map_bytes + per_warp) return KGMA_E_UNSUPPORTED;
    int wmax = serial ? 16 : 20;
    if (const char *e = getenv("KGMA_EVAL_WARPS")) wmax = std::max(1, std::min(20, atoi(e)));
    int w = (int)std::min<size_t>((ctx->smem_optin - nb * 4 - map_bytes) / per_warp, (size_t)wmax);
    *warps_out = w; *smem_out = nb * 4 + map_bytes + (size_t)w * per_warp;
    return KGMA_OK;
}

constexpr long long NO_D = (long long)0x8080808080808080ull;   // cudaMemset(0x80) pattern = "first window not evaluated"

}  // namespace kgma

using namespace kgma;

// Pipelined streaming scan: records [0, rec_split) are evaluated as soon as their last block has been
// filtered; on_first_part is called with their runs while the rest of the genome is still being copied and filtered.
struct PhaseHook {
    int rec_split = 0;
    bool resident = false;         // the split was chosen for a genome that is already on the device (KGMA_RESIDENT_SPLIT)
    std::function<int(std::vector<kgma_run> &, const std::vector<int64_t> &)> on_first_part;
    size_t n_runs_first = 0;       // out: how many entries of res->runs belong to the first part
    bool used = false;             // out: the scan did run in two parts
};

// One prefilter pass = one group of profiles sharing a weight table, a candidate list and a block bitmap.
struct FilterGroup {
    std::vector<int> q;            // profile indices
    bool dense = false;            // no prefilter: the count-table kernel sees every window
    const kgma_ctx::FTab *ft = nullptr;
    size_t o_tab = 0, o_cand = 0, o_bits = 0;
};

static const kgma_ctx::FTab *get_ftab(kgma_ctx *ctx, const ScanPlan &pl, const std::vector<int> &group)
{
    uint64_t key = 1469598103934665603ull;                         // FNV-1a over the members' fingerprints (ProfTab::hash)
    auto mix = [&](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; for (size_t i = 0; i < n; i++) { key ^= b[i]; key *= 1099511628211ull; } };
    for (int q : group) mix(&pl.tabs[(size_t)q].hash, 8);
    mix(&pl.maxnk, 8);
    // the 9-mer table (kgma_prefilter9) is the default; KGMA_PREFILTER=8mer selects the round-1 kernel and its table
    const char *pf_env = getenv("KGMA_PREFILTER");
    const bool nine = !(pf_env && !strcmp(pf_env, "8mer")) && pl.k <= 7;
    const unsigned char nine_b = nine ? 1 : 0;
    mix(&nine_b, 1);
    for (const auto &f : ctx->ftabs) if (f.key == key) return &f;
    ctx->ftabs.emplace_back();
    kgma_ctx::FTab &f = ctx->ftabs.back();
    f.key = key; f.ok = build_filter_table(pl, group, f, nine);
    return &f;
}

// The scan proper (one context / one GPU / one shard).  Fills res->runs, res->first_D, res->dists.
// Device work is queued back to back (uploads, prefilter launches chasing the genome chunks, count-table
// kernel reading the candidate lists from device memory, result copies) with ONE host synchronisation at the end.
// Returns KGMA_E_CAPACITY with *need_runs set when the run list was too small (the caller retries once).
static int scan_runs_impl(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int C,
                          const kgma_scan_params &P, ScanPlan &pl, kgma_result *res, uint32_t run_cap, uint64_t *need_runs,
                          PhaseHook *hook = nullptr)
{
    const double t_wall0 = now_ms();
    *need_runs = 0;
    if (!g->sealed) return set_err(ctx, KGMA_E_STATE, "genome is not sealed");
    if (g->ambiguous)
        return set_err(ctx, KGMA_E_SYMBOL, "KeyError: record %lld position %lld holds a symbol outside A,C,G,T,N",
                       (long long)g->amb_record, (long long)g->amb_pos);
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = make_plan(ctx, g, profiles, C, P, pl);
    if (rc) return rc;
    kgma_stats &st = ctx->stats; st = kgma_stats{};
    const int nr = (int)g->recs.size();
    const bool want_dists = (P.flags & KGMA_F_WANT_DISTS) != 0;
    const bool force_dense = (P.flags & KGMA_F_DENSE) != 0 || want_dists || pl.strobe;   // (strobemer codes: no lower-bound table; 4^(2s) bins are all populated)

    // ---- shard range in 64-base blocks (whole warp groups)
    const int64_t nblk_total = g->G / FBLOCK;
    const int64_t ngrp_total = nblk_total / 32;
    int sc = std::max(1, P.shard_count), si = std::min(std::max(0, P.shard_index), sc - 1);
    const int64_t blk_lo = (ngrp_total * si / sc) * 32, blk_hi = (ngrp_total * (si + 1) / sc) * 32;
    const int64_t pos_lo = blk_lo * FBLOCK, pos_hi = blk_hi * FBLOCK;

    // ---- prefilter groups.  One table for all profiles when a random sequence stays well below the flag threshold
    //      under the combined (max) weights; otherwise one pass per profile; profiles that cannot be filtered go dense.
    std::vector<FilterGroup> groups;
    if (ctx->ftabs.size() > 96) ctx->ftabs.clear();                // bounded cache; nothing points into it between calls
    {
        std::vector<int> all((size_t)C); for (int q = 0; q < C; q++) all[(size_t)q] = q;
        double LOAD_OK = 0.62; const double LOAD_MAX = 0.85;
        if (const char *e = getenv("KGMA_LOAD_OK")) LOAD_OK = atof(e);     // (experiments: how full a shared table may be)
        FilterGroup dg; dg.dense = true;
        if (force_dense) dg.q = all;
        else {
            const kgma_ctx::FTab *f = get_ftab(ctx, pl, all);
            if (f->ok && (f->load <= LOAD_OK || (C == 1 && f->load <= LOAD_MAX))) { FilterGroup fg; fg.q = all; fg.ft = f; groups.push_back(fg); }
            else if (C == 1) dg.q = all;
            else for (int q = 0; q < C; q++) {
                // greedy packing: join the first group whose combined (max-weight) table still keeps random sequence well
                // below the flag threshold, else open a new group; tables are cached, so this costs nothing on repeat scans
                bool placed = false;
                for (FilterGroup &fg : groups) {
                    std::vector<int> trial = fg.q; trial.push_back(q);
                    const kgma_ctx::FTab *ft = get_ftab(ctx, pl, trial);           // (references into the deque stay valid)
                    if (ft->ok && ft->load <= LOAD_OK) { fg.q = trial; fg.ft = ft; placed = true; break; }
                }
                if (placed) continue;
                const kgma_ctx::FTab *fq = get_ftab(ctx, pl, std::vector<int>{ q });
                if (fq->ok && fq->load <= LOAD_MAX) { FilterGroup fg; fg.q = { q }; fg.ft = fq; groups.push_back(fg); }
                else dg.q.push_back(q);
            }
        }
        if (!dg.q.empty()) groups.push_back(dg);
    }
    int M = 1;
    for (const FilterGroup &fg : groups) if (!fg.dense) M = std::max(M, fg.ft->M);
    bool any_filter = false, any_dense = false;
    for (const FilterGroup &fg : groups) { any_filter |= !fg.dense; any_dense |= fg.dense; }

    rc = dev_genome_prepare(ctx, g, false);
    if (rc) return rc;

    int ewarps = 0; size_t esmem = 0;
    const char *ek_env = getenv("KGMA_EVAL_KERNEL");
    const bool eval_serial = ek_env && !strcmp(ek_env, "serial") && !pl.strobe;
    rc = eval_shape(ctx, pl.kb, eval_serial, &ewarps, &esmem, pl.strobe ? pl.cmap.size() * 2 : 0);
    if (rc) return set_err(ctx, rc, "k = %d does not fit the shared-memory count tables", pl.k);
    const size_t nb = (size_t)1 << (2 * pl.kb);
    const int egrid = ctx->num_sms;
    const int64_t total_warps = (int64_t)egrid * ewarps;

    // ---- per-record window ranges owned by this shard (+ dense items, dist offsets)
    std::vector<RecDev> recs((size_t)std::max(nr, 1));
    std::vector<uint32_t> seeds;                                   // blocks holding window 0 of a record: always evaluated
    int64_t span_windows = 0, ndist = 0;
    for (int r = 0; r < nr; r++) {
        RecDev &R = recs[(size_t)r];
        R.off = g->recs[r].off; R.w_begin = R.w_end = 0; R.item_base = 0; R.dist_base = ndist;
        if (want_dists) ndist += pl.steps[r];
        if (pl.steps[r] <= 0) continue;
        const int64_t lo = std::max<int64_t>(R.off, pos_lo), hi = std::min<int64_t>(R.off + pl.steps[r] + 1, pos_hi);   // windows 0..steps
        if (lo >= hi) continue;
        R.w_begin = lo - R.off; R.w_end = hi - R.off;
        span_windows += hi - lo;
        if (R.w_begin == 0) seeds.push_back((uint32_t)(R.off / FBLOCK));
    }
    st.bases_scanned = span_windows;
    int64_t n_items = 0; int span = 64;
    {   // dense items (also the fallback when a candidate list overflows)
        int64_t sp = span_windows / std::max<int64_t>(1, total_warps * 4) + 1;
        sp = std::min<int64_t>(std::max<int64_t>((sp + 63) / 64 * 64, 64), 1 << 16);
        span = (int)sp;
        for (int r = 0; r < nr; r++) { recs[(size_t)r].item_base = n_items; n_items += (recs[(size_t)r].w_end - recs[(size_t)r].w_begin + sp - 1) / sp; }
    }

    // ---- device scratch layout + one pinned staging block for all small uploads
    const uint32_t cand_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>((blk_hi - blk_lo) / 16, 1 << 16) + (int64_t)seeds.size(), 1 << 26);
    // runs copied back together with the counters (more only if needed); several profiles report several times the runs
    const uint32_t run_head = C == 1 ? 4096 : 32768;
    size_t o = 0;
    auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    const size_t o_S = carve((size_t)C * nb * 4), o_recs = carve(recs.size() * sizeof(RecDev)), o_cmap = carve(pl.cmap.size() * 2);
    const size_t o_seed = carve((seeds.size() + 1) * 4);
    const size_t up_small = o;                                     // ... up to here on every call; the weight tables only when they changed
    for (FilterGroup &fg : groups) if (!fg.dense) fg.o_tab = carve(fg.ft->nine ? (size_t)TAB9_BYTES : (size_t)65536 * 2);
    const size_t up_bytes = o;                                     // everything above is uploaded from the staging block
    const size_t o_cnt = carve(512), o_first = carve((size_t)C * std::max(nr, 1) * 8), o_runs = carve((size_t)run_cap * sizeof(kgma_run));
    const size_t bitmap_bytes = ((size_t)nblk_total / 32 + 4) * 4;
    for (FilterGroup &fg : groups) if (!fg.dense) { fg.o_cand = carve((size_t)cand_cap * 4); fg.o_bits = carve(bitmap_bytes); }
    const size_t o_dists = carve(want_dists ? (size_t)C * (size_t)std::max<int64_t>(ndist, 1) * 8 : 0);
    void *dsv = nullptr;
    rc = dev_scratch(ctx, o, &dsv);
    if (rc) return rc;
    unsigned char *ds = (unsigned char *)dsv;
    const size_t back_bytes = 256 + (size_t)C * std::max(nr, 1) * 8 + (size_t)run_head * sizeof(kgma_run);
    void *hsv = nullptr;
    rc = host_scratch(ctx, up_bytes + 2 * back_bytes, &hsv);
    if (rc) return rc;
    unsigned char *hs = (unsigned char *)hsv, *hback = hs + up_bytes, *hbackA = hback + back_bytes;
    for (int q = 0; q < C; q++) memcpy(hs + o_S + (size_t)q * nb * 4, pl.tabs[q].S_rev.data(), nb * 4);
    if (pl.strobe) memcpy(hs + o_cmap, pl.cmap.data(), pl.cmap.size() * 2);
    // The prefilter tables (192 KB each) stay in the device arena between scans: a serving loop with the same profiles uploads
    // them once.  The signature covers which table sits at which offset of which allocation.
    uint64_t tsig = 1469598103934665603ull;
    {
        auto mix = [&](uint64_t w) { tsig ^= w; tsig *= 1099511628211ull; tsig ^= tsig >> 31; };
        mix((uint64_t)(uintptr_t)ds);
        for (const FilterGroup &fg : groups) if (!fg.dense) { mix(fg.ft->key); mix((uint64_t)fg.o_tab); mix(fg.ft->nine ? 9 : 8); }
        if (tsig == 0) tsig = 1;
    }
    const bool tabs_resident = up_bytes > up_small && ctx->tab_sig == tsig && !getenv("KGMA_NO_TABLE_CACHE");
    if (!tabs_resident)
        for (const FilterGroup &fg : groups) if (!fg.dense) { if (fg.ft->nine) memcpy(hs + fg.o_tab, fg.ft->tab9.data(), TAB9_BYTES); else memcpy(hs + fg.o_tab, fg.ft->tab.data(), 65536 * 2); }
    memcpy(hs + o_recs, recs.data(), recs.size() * sizeof(RecDev));
    if (!seeds.empty()) memcpy(hs + o_seed, seeds.data(), seeds.size() * 4);
    { const uint32_t ns = (uint32_t)seeds.size(); memcpy(hs + o_seed + seeds.size() * 4, &ns, 4); }

    cudaStream_t sc_ = ctx->s_compute, sp = ctx->s_copy;
    cudaEvent_t e_start = ctx->ev[0], e_h2d = ctx->ev[1], e_filt = ctx->ev[2], e_exact = ctx->ev[3], e_fstart = ctx->ev[4];
    // counter block (512 B): [0] run_count, [1 + gi] candidate count of group gi, bytes 128.. next_item[q],
    // bytes 256.. next_item[q] of the first part of a pipelined scan
    uint32_t *d_counters = (uint32_t *)(ds + o_cnt);
    KGMA_CUDA(ctx, cudaEventRecord(e_start, sc_));
    KGMA_CUDA(ctx, cudaMemcpyAsync(ds, hs, tabs_resident ? up_small : up_bytes, cudaMemcpyHostToDevice, sc_));
    ctx->tab_sig = up_bytes > up_small ? tsig : 0;
    KGMA_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 512, sc_));
    KGMA_CUDA(ctx, cudaMemsetAsync(ds + o_first, 0x80, (size_t)C * std::max(nr, 1) * 8, sc_));
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const FilterGroup &fg = groups[gi];
        if (fg.dense) continue;
        KGMA_CUDA(ctx, cudaMemsetAsync(ds + fg.o_bits + (size_t)(blk_lo / 32) * 4, 0, (size_t)((blk_hi - blk_lo) / 32 + 2) * 4, sc_));
        if (!seeds.empty()) {                                      // pre-seed the candidate list
            KGMA_CUDA(ctx, cudaMemcpyAsync(ds + fg.o_cand, ds + o_seed, seeds.size() * 4, cudaMemcpyDeviceToDevice, sc_));
            KGMA_CUDA(ctx, cudaMemcpyAsync(d_counters + 1 + gi, ds + o_seed + seeds.size() * 4, 4, cudaMemcpyDeviceToDevice, sc_));
        }
    }
    st.h2d_bytes += tabs_resident ? up_small : up_bytes;
    st.host_setup_ms = now_ms() - t_wall0;

    // ---- upload range (bases): shard + halo, unless resident
    const int64_t halo = (int64_t)((M + 1) * FBLOCK + pl.maxws + 64);
    int64_t up_lo = pos_lo, up_hi = std::min(g->G + TAIL_PAD, pos_hi + halo + 3 * FGROUP);
    if (sc > 1 && (P.flags & KGMA_F_ALIGN)) {
        // kgma_scan_shard extends its own candidates: their windows reach buff bases to the left of the first owned window
        // and window + buff to the right of the last one
        const int64_t reach = (std::max<int64_t>(P.buff, 0) + pl.maxws + FGROUP - 1) / FGROUP * FGROUP;
        up_lo = std::max<int64_t>(0, pos_lo - reach);
        up_hi = std::min(g->G + TAIL_PAD, up_hi + reach);
    }
    if (si == sc - 1) up_hi = g->G + TAIL_PAD;
    up_hi = (up_hi + 127) / 128 * 128; up_hi = std::min(up_hi, g->G + TAIL_PAD);
    const bool resident_ok = (P.flags & KGMA_F_RESIDENT) && ctx->d_seq_valid && ctx->d_valid_lo <= up_lo && ctx->d_valid_hi >= up_hi;

    // pipelined scan: the records before hook->rec_split form the first part.  Candidates are divided exactly at the block
    // where record rec_split starts (records start on block boundaries); the prefilter launch in front of the first part's
    // evaluation ends on the next multiple of 32 blocks (its launches work in whole warp groups), which only means a few
    // blocks of the second part are filtered early.
    // page-locked source, or (first upload of a pageable genome) the staging ring; a resident genome is not touched at all
    // (page-locking 772 MB was measured anywhere between 14 and 380 ms on the same box, depending on how many huge pages the
    //  kernel could hand out; a staged upload costs ~2.5 ms more per scan than a page-locked one, so pageable genomes are
    //  always staged and only an explicit kgma_genome_make_resident / exact match page-locks)
    const bool staged = !resident_ok && !g->pinned && !getenv("KGMA_NO_STAGING");
    if (!resident_ok) {
        rc = staged ? StagedUpload::prepare_ring(ctx) : genome_pin(ctx, g);
        if (rc) return rc;
    }
    int64_t split_blk = -1, split_blk_f = -1;
    bool pipelined = hook && (resident_ok ? hook->resident : !staged) && any_filter && !any_dense && sc == 1 && nr > 1 &&
                     hook->rec_split > 0 && hook->rec_split < nr;
    if (pipelined) {
        split_blk = g->recs[(size_t)hook->rec_split].off / FBLOCK;
        split_blk_f = (split_blk + 31) / 32 * 32;
        if (split_blk_f >= blk_hi || split_blk <= blk_lo) pipelined = false;
    }
    bool partA_enqueued = false;
    cudaEvent_t e_partA = ctx->ev[7];
    std::function<int()> enqueue_part_a;                     // defined below, once the eval arguments exist
    // ---- count-table kernel over the candidate lists (read on the device) or over everything
    EvalArgs ea{};
    ea.seq = ctx->d_seq2; ea.S = (const int32_t *)(ds + o_S);
    ea.cmap = pl.strobe ? (const uint16_t *)(ds + o_cmap) : nullptr; ea.xmask = pl.strobe ? (uint32_t)pl.cmap.size() - 1 : 0;
    ea.cand_cap = cand_cap; ea.n_seed = (uint32_t)seeds.size();
    ea.next_item = (unsigned long long *)(d_counters + 32);        // bytes 128.. of the zeroed counter block
    ea.recs = (const RecDev *)(ds + o_recs); ea.nrec = nr;
    ea.cand_blk_lo = 0; ea.cand_blk_hi = LLONG_MAX;
    ea.C = C; ea.k = pl.kb; ea.span = span;
    for (int q = 0; q < C; q++) {
        const ProfTab &t = pl.tabs[q];
        ea.prof[q].N2 = t.N2; ea.prof[q].twoN = t.twoN; ea.prof[q].sumS2 = t.sumS2;
        ea.prof[q].T = t.T; ea.prof[q].Tlo = t.Tlo; ea.prof[q].Thi = t.Thi; ea.prof[q].nk = (int)t.nk - (t.strobe ? 1 : 0); ea.prof[q].N = t.N;   // (strobemers: nk - 1 sliding codes + the permanent one)
        {   // D < Thi  <=>  N (N Q - 2 A) < Thi - sum S^2  <=>  N Q - 2 A < ceil((Thi - sum S^2) / N); everything must fit 31 bits
            int64_t maxS = 0; for (int32_t v : t.S_rev) maxS = std::max<int64_t>(maxS, v);
            const __int128 umax = (__int128)t.N * t.nk * t.nk + 2 * (__int128)t.nk * maxS;
            const __int128 num = (__int128)t.Thi - t.sumS2;
            __int128 uq = num >= 0 ? (num + t.N - 1) / t.N : -((-num) / t.N);          // ceil for either sign
            ea.prof[q].u_ok = umax < ((__int128)1 << 30) && uq < ((__int128)1 << 30) && uq > -((__int128)1 << 30);
            ea.prof[q].uThi = ea.prof[q].u_ok ? (int)uq : 0;
        }
        {
            const __int128 R = (__int128)t.N2 * t.nk + t.sumS2 - t.Thi;
            ea.prof[q].R = R > 0 && R < ((__int128)1 << 62) ? (long long)R : 0;
        }
    }
    ea.runs = (kgma_run *)(ds + o_runs); ea.run_cap = run_cap; ea.run_count = d_counters;
    ea.first_D = (long long *)(ds + o_first);
    ea.dists = want_dists ? (long long *)(ds + o_dists) : nullptr; ea.dist_stride = std::max<int64_t>(ndist, 1);
    if (eval_serial) KGMA_CUDA(ctx, cudaFuncSetAttribute(kgma_eval_serial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    auto launch_eval = [&](const std::vector<int> &qs, bool dense_items, size_t gi, const FilterGroup *fg) -> int {
        ea.nq = (int)qs.size();
        for (size_t i = 0; i < qs.size(); i++) ea.qlist[i] = qs[i];
        if (dense_items) { ea.cand = nullptr; ea.cand_count = nullptr; ea.bitmap = nullptr; ea.n_items = n_items; }
        else { ea.cand = (const uint32_t *)(ds + fg->o_cand); ea.cand_count = d_counters + 1 + gi; ea.bitmap = (const uint32_t *)(ds + fg->o_bits); ea.n_items = 0; }
        if (eval_serial) { kgma_eval_serial<<<egrid, ewarps * 32, esmem, sc_>>>(ea); st.launches++; }
        else for (int q : qs) { launch_eval_k(pl.kb, ea, ea.prof[q], q, egrid, ewarps * 32, esmem, sc_); st.launches++; }
        KGMA_CUDA(ctx, cudaGetLastError());
        return KGMA_OK;
    };
    uint32_t cnts[64] = { 0 };
    enqueue_part_a = [&]() -> int {  // first part of a pipelined scan: evaluate its candidates, queue the copy of its results
        ea.cand_blk_lo = 0; ea.cand_blk_hi = split_blk;
        ea.next_item = (unsigned long long *)(d_counters + 64);   // bytes 256..
        int rc2 = KGMA_OK;
        for (size_t gi = 0; gi < groups.size() && !rc2; gi++) rc2 = launch_eval(groups[gi].q, false, gi, &groups[gi]);
        ea.next_item = (unsigned long long *)(d_counters + 32);
        ea.cand_blk_lo = split_blk; ea.cand_blk_hi = LLONG_MAX;    // what the final evaluation still has to do
        if (rc2) return rc2;
        KGMA_CUDA(ctx, cudaMemcpyAsync(hbackA, d_counters, 256, cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaMemcpyAsync(hbackA + 256, ds + o_first, (size_t)C * nr * 8, cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaMemcpyAsync(hbackA + 256 + (size_t)C * nr * 8, ds + o_runs, (size_t)run_head * sizeof(kgma_run), cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaEventRecord(e_partA, sc_));
        st.d2h_bytes += back_bytes;
        return KGMA_OK;
    };
    auto fetch = [&]() -> int {     // counters + first-window distances + the head of the run list, one synchronisation
        KGMA_CUDA(ctx, cudaMemcpyAsync(hback, d_counters, 256, cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaMemcpyAsync(hback + 256, ds + o_first, (size_t)C * nr * 8, cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaMemcpyAsync(hback + 256 + (size_t)C * nr * 8, ds + o_runs, (size_t)run_head * sizeof(kgma_run), cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaStreamSynchronize(sc_));
        memcpy(cnts, hback, 256);
        st.d2h_bytes += back_bytes;
        return KGMA_OK;
    };
    // ---- stream the packed genome: chunked cudaMemcpyAsync on the copy stream, prefilter on the
    //      compute stream chasing it (double buffering falls out of the two streams + per-chunk events)
    const int64_t CH = (int64_t)128 << 20;                  // bases per chunk (32 MB of packed data)
    const int fgrid = ctx->num_sms;
    int64_t done_blk = blk_lo;                               // target blocks already filtered
    const int64_t need_after = (int64_t)(M + 1) * FBLOCK + 3 * FGROUP;     // bases that must be present past a target block
    bool first_filter = true;
    auto launch_group_filter = [&](const FilterGroup &fg, size_t gi, int64_t b0, int64_t b1) {
        if (fg.ft->nine) {
            Filter9Args fa{};
            fa.seq = (const uint4 *)ctx->d_seq2; fa.tab = (const uint8_t *)(ds + fg.o_tab); fa.M = fg.ft->M; fa.thrw = fg.ft->thrw;
            fa.cand = (uint32_t *)(ds + fg.o_cand); fa.cand_cap = cand_cap; fa.cand_count = d_counters + 1 + gi;
            fa.bitmap = (uint32_t *)(ds + fg.o_bits);
            fa.blk_begin = b0; fa.blk_end = b1;
            launch_filter9_k(pl.k, fa, fgrid, sc_);
        } else {
            FilterArgs fa{};
            fa.seq = (const uint4 *)ctx->d_seq2; fa.tab = (const uint16_t *)(ds + fg.o_tab); fa.M = fg.ft->M; fa.thrw = fg.ft->thrw;
            fa.cand = (uint32_t *)(ds + fg.o_cand); fa.cand_cap = cand_cap; fa.cand_count = d_counters + 1 + gi;
            fa.bitmap = (uint32_t *)(ds + fg.o_bits);
            fa.blk_begin = b0; fa.blk_end = b1;
            launch_filter_k(pl.k, fa, fgrid, sc_);
        }
    };
    auto run_filter_to = [&](int64_t avail_hi, bool last) -> int {
        if (!any_filter) return KGMA_OK;
        int64_t lim = last ? blk_hi : std::min(blk_hi, ((avail_hi - need_after) / FBLOCK) / 32 * 32);
        if (pipelined && !partA_enqueued && lim >= split_blk_f) {
            // stop this launch on the split so that the first part can be evaluated right behind it, then carry on
            const int64_t rest = lim;
            lim = split_blk_f;
            if (lim > done_blk) {
                if (first_filter) { KGMA_CUDA(ctx, cudaEventRecord(e_fstart, sc_)); first_filter = false; }
                for (size_t gi = 0; gi < groups.size(); gi++) {
                    const FilterGroup &fg = groups[gi];
                    launch_group_filter(fg, gi, done_blk, lim);
                    KGMA_CUDA(ctx, cudaGetLastError());
                    st.launches++;
                }
                done_blk = lim;
            }
            int rc2 = enqueue_part_a();
            if (rc2) return rc2;
            partA_enqueued = true;
            lim = rest;
        }
        if (lim <= done_blk) return KGMA_OK;
        if (first_filter) { KGMA_CUDA(ctx, cudaEventRecord(e_fstart, sc_)); first_filter = false; }
        for (size_t gi = 0; gi < groups.size(); gi++) {
            const FilterGroup &fg = groups[gi];
            if (fg.dense) continue;
            launch_group_filter(fg, gi, done_blk, lim);
            KGMA_CUDA(ctx, cudaGetLastError());
            st.launches++;
        }
        done_blk = lim;
        return KGMA_OK;
    };
    if (!resident_ok) {
        cudaEvent_t e_c[2] = { ctx->ev[5], ctx->ev[6] };
        KGMA_CUDA(ctx, cudaEventRecord(ctx->ev[6 + 0], sc_));      // (ev[6] doubles as the "setup queued" marker before the loop)
        KGMA_CUDA(ctx, cudaStreamWaitEvent(sp, ctx->ev[6], 0));
        const size_t nchunks = (size_t)((up_hi - up_lo + CH - 1) / CH);
        if (pipelined) while (ctx->chunk_ev.size() < nchunks) { cudaEvent_t e; KGMA_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
        ctx->d_have_lo = up_lo; ctx->d_have_hi = up_hi;            // what the queued copies will have delivered (extension of the first part reads it)
        int ci = 0;
        StagedUpload stager;                                       // (its destructor joins the copy threads on every exit path)
        if (staged) stager.start(ctx, (const char *)g->seq2 + up_lo / 4, (size_t)(up_hi - up_lo) / 4);
        for (int64_t a = up_lo; a < up_hi; a += CH, ci++) {
            int64_t b = std::min(up_hi, a + CH);
            if (staged) {
                const size_t j0 = (size_t)((a - up_lo) / 4) / StagedUpload::SB;
                const size_t j1 = b >= up_hi ? stager.nsub : (size_t)((b - up_lo) / 4) / StagedUpload::SB;
                rc = stager.issue(j0, j1, (char *)ctx->d_seq2 + up_lo / 4, sp);
                if (rc) return rc;
            } else
            KGMA_CUDA(ctx, cudaMemcpyAsync((char *)ctx->d_seq2 + a / 4, (char *)g->seq2 + a / 4, (size_t)(b - a) / 4,
                                           cudaMemcpyHostToDevice, sp));
            cudaEvent_t ec = pipelined ? ctx->chunk_ev[(size_t)ci] : e_c[ci & 1];
            KGMA_CUDA(ctx, cudaEventRecord(ec, sp));
            KGMA_CUDA(ctx, cudaStreamWaitEvent(sc_, ec, 0));
            st.h2d_bytes += (b - a) / 4;
            rc = run_filter_to(b, b >= up_hi);
            if (rc) return rc;
            // two events are recycled: wait for the older copy before its event is recorded again.  A pipelined scan has one
            // event per chunk and queues everything without blocking, so the host is free to work on the first part's results.
            if (!pipelined && ci >= 1) KGMA_CUDA(ctx, cudaEventSynchronize(e_c[(ci - 1) & 1]));
        }
        KGMA_CUDA(ctx, cudaEventRecord(e_h2d, sc_));
        ctx->d_seq_valid = true; ctx->d_valid_lo = up_lo; ctx->d_valid_hi = up_hi;
        ctx->d_have_lo = up_lo; ctx->d_have_hi = up_hi;
    } else {
        KGMA_CUDA(ctx, cudaEventRecord(e_h2d, sc_));
        rc = run_filter_to(up_hi, true);
        if (rc) return rc;
    }
    if (first_filter) KGMA_CUDA(ctx, cudaEventRecord(e_fstart, sc_));
    KGMA_CUDA(ctx, cudaEventRecord(e_filt, sc_));

    if (nr > 0) {
        for (size_t gi = 0; gi < groups.size(); gi++) { rc = launch_eval(groups[gi].q, groups[gi].dense, gi, &groups[gi]); if (rc) return rc; }
        KGMA_CUDA(ctx, cudaEventRecord(e_exact, sc_));
        if (pipelined && partA_enqueued) {
            // everything is queued; while the rest of the genome streams, hand the first part's runs to the caller
            KGMA_CUDA(ctx, cudaEventSynchronize(e_partA));
            uint32_t ca[64]; memcpy(ca, hbackA, 256);
            bool fits = ca[0] <= run_head;
            for (size_t gi = 0; gi < groups.size(); gi++) fits = fits && ca[1 + gi] <= cand_cap;
            if (fits) {
                std::vector<kgma_run> ra((size_t)ca[0]);
                if (ca[0]) memcpy(ra.data(), hbackA + 256 + (size_t)C * nr * 8, (size_t)ca[0] * sizeof(kgma_run));
                std::vector<int64_t> fd((size_t)C * nr, INT64_MIN);
                const long long *fdp = (const long long *)(hbackA + 256);
                for (size_t i = 0; i < (size_t)C * nr; i++) if (fdp[i] != NO_D) fd[i] = fdp[i];
                hook->n_runs_first = ca[0]; hook->used = true;
                const double tA0 = now_ms();
                rc = hook->on_first_part(ra, fd);
                if (getenv("KGMA_TRACE")) fprintf(stderr, "[kgma scan] first part: %u runs, ready at %.2f ms, handled in %.2f ms\n", ca[0], tA0 - t_wall0, now_ms() - tA0);
                if (rc) { cudaStreamSynchronize(sc_); return rc; }
            }
        }
        rc = fetch(); if (rc) return rc;
        if (getenv("KGMA_TRACE")) fprintf(stderr, "[kgma scan] device done at %.2f ms\n", now_ms() - t_wall0);
        std::vector<int> redo;                                     // groups whose candidate list overflowed: evaluate every window
        for (size_t gi = 0; gi < groups.size(); gi++)
            if (!groups[gi].dense && cnts[1 + gi] > cand_cap) { redo.insert(redo.end(), groups[gi].q.begin(), groups[gi].q.end()); groups[gi].dense = true; any_dense = true; }
        if (!redo.empty() && cnts[0] <= run_cap) {
            // partial runs of the overflowed groups are duplicates of what the dense pass will report: drop them on the host below
            for (int q : redo) KGMA_CUDA(ctx, cudaMemsetAsync((unsigned long long *)(d_counters + 32) + q, 0, 8, sc_));
            rc = launch_eval(redo, true, 0, nullptr); if (rc) return rc;
            KGMA_CUDA(ctx, cudaEventRecord(e_exact, sc_));
            rc = fetch(); if (rc) return rc;
        }
    }
    const double t_cand0 = now_ms();
    st.blocks_total = any_filter ? blk_hi - blk_lo : 0;
    for (const FilterGroup &fg : groups) if (fg.ft) st.filter_passes++;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        if (groups[gi].ft) { st.blocks_flagged += cnts[1 + gi]; }
        st.exact_windows += groups[gi].dense ? span_windows * (int64_t)groups[gi].q.size() : (int64_t)std::min(cnts[1 + gi], cand_cap) * FBLOCK * (int64_t)groups[gi].q.size();
    }
    if (cnts[0] > run_cap) { *need_runs = cnts[0]; return set_err(ctx, KGMA_E_CAPACITY, "run list overflow (%u runs)", cnts[0]); }
    std::vector<kgma_run> &runs = res->runs;
    runs.resize(cnts[0]);
    if (cnts[0]) memcpy(runs.data(), hback + 256 + (size_t)C * nr * 8, (size_t)std::min(cnts[0], run_head) * sizeof(kgma_run));
    if (cnts[0] > run_head) {
        KGMA_CUDA(ctx, cudaMemcpy(runs.data() + run_head, ds + o_runs + (size_t)run_head * sizeof(kgma_run),
                                  (size_t)(cnts[0] - run_head) * sizeof(kgma_run), cudaMemcpyDeviceToHost));
        st.d2h_bytes += (size_t)(cnts[0] - run_head) * sizeof(kgma_run);
    }
    res->first_D.assign((size_t)C * nr, INT64_MIN);
    {
        const long long *fd = (const long long *)(hback + 256);
        for (size_t i = 0; i < (size_t)C * nr; i++) if (fd[i] != NO_D) res->first_D[i] = fd[i];
    }
    res->dists.assign(C, {});
    if (want_dists && ndist > 0) {
        std::vector<long long> Dh((size_t)ndist);
        for (int q = 0; q < C; q++) {
            KGMA_CUDA(ctx, cudaMemcpy(Dh.data(), ds + o_dists + (size_t)q * ndist * 8, (size_t)ndist * 8, cudaMemcpyDeviceToHost));
            st.d2h_bytes += (size_t)ndist * 8;
            res->dists[q].resize((size_t)ndist);
            const double den = pl.tabs[q].denom;
            for (int64_t i = 0; i < ndist; i++) res->dists[q][(size_t)i] = (double)Dh[(size_t)i] / den;
        }
    }
    st.n_runs = cnts[0];
    float ms = 0;
    cudaEventElapsedTime(&ms, e_start, e_h2d); st.h2d_ms = ms;
    cudaEventElapsedTime(&ms, e_fstart, e_filt); st.filter_ms = ms;
    if (nr) { cudaEventElapsedTime(&ms, e_filt, e_exact); st.exact_ms = ms; cudaEventElapsedTime(&ms, e_start, e_exact); st.total_ms = ms; }
    if (!(P.flags & KGMA_F_RESIDENT)) { ctx->d_seq_valid = false; }
    st.host_cand_ms = now_ms() - t_cand0;
    st.wall_ms = now_ms() - t_wall0;
    return KGMA_OK;
}

// run list capacity: 1 Mi runs to start with, grown to what the device counted when that was not enough
static int scan_runs_retry(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int C,
                           const kgma_scan_params &P, ScanPlan &pl, kgma_result *res,
                           PhaseHook *hook = nullptr, const std::function<void()> &reset = nullptr)
{
    uint64_t need = 0;
    int rc = scan_runs_impl(ctx, g, profiles, C, P, pl, res, 1u << 20, &need, hook);
    // a retry can overflow again: when a candidate list overflowed too, the dense re-evaluation reports every partial run a
    // second time (up to twice the count the first attempt saw), so the capacity doubles per attempt, a few times at most
    for (int attempt = 0; attempt < 4 && rc == KGMA_E_CAPACITY && need > 0 && need < (1ull << 28); attempt++) {
        pl = ScanPlan();
        if (hook) { hook->used = false; hook->n_runs_first = 0; }
        if (reset) reset();
        rc = scan_runs_impl(ctx, g, profiles, C, P, pl, res, (uint32_t)std::min<uint64_t>(2 * need + 1024, 1ull << 28), &need, nullptr);
    }
    return rc;
}

extern "C" {

int kgma_scan_runs(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                   const kgma_scan_params *params, kgma_result **out)
{
    if (!ctx || !g || !params || !out) return KGMA_E_ARG;
    kgma_result *res = result_acquire();
    ScanPlan pl;
    int rc = scan_runs_retry(ctx, g, profiles, n_profiles, *params, pl, res);
    if (rc) { result_release(res); return rc; }
    *out = res;
    return KGMA_OK;
}

// ---- multi-GPU, one call per rank ------------------------------------------------------------------------------
// The shard's runs are merged, and the candidate window of every run that can still end up as a hit (its own first
// argmin: a superset of what the replay can select) is extended right here, on the GPU that holds
// the bases.  What leaves the rank is (run, extension result) pairs -- a few KB; the replay of all shards' pairs is
// pure host logic (kgma_replay_packed) and needs no device at all, so no rank waits for another rank's kernels.
int kgma_scan_shard(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                    const kgma_scan_params *params, kgma_result **out)
{
    if (!ctx || !g || !params || !out) return KGMA_E_ARG;
    if (params->flags & (KGMA_F_WANT_CIGARS | KGMA_F_WANT_DISTS)) return set_err(ctx, KGMA_E_UNSUPPORTED, "kgma_scan_shard returns neither CIGARs nor distances");
    if (params->mode == KGMA_MODE_STROBE) return set_err(ctx, KGMA_E_UNSUPPORTED, "strobemer mode is not sharded (use one genome of whole records per device)");
    kgma_result *res = result_acquire();
    ScanPlan pl;
    int rc = scan_runs_retry(ctx, g, profiles, n_profiles, *params, pl, res);
    if (rc) { result_release(res); return rc; }
    const double t0 = now_ms();
    const kgma_scan_params &P = *params;
    merge_runs(res->runs);
    res->run_ext.assign(res->runs.size(), kgma_run_ext{ 0, 0, 0 });
    if (P.flags & KGMA_F_ALIGN) {
        std::vector<AlignReq> reqs; std::vector<size_t> of_run;
        for (size_t i = 0; i < res->runs.size(); i++) {
            const kgma_run &ru = res->runs[i];
            if (ru.flags & KGMA_RUN_MARKER) continue;
            if (ru.t_last >= pl.steps[(size_t)ru.record]) continue;          // reaches the record's last step: never emitted
            const int64_t L = g->recs[(size_t)ru.record].len, wsq = pl.tabs[(size_t)ru.profile].ws;
            const int64_t CMI = pl.cluster ? ru.t_argmin : pl.strobe ? ru.t_argmin + 1    // OmnGenomeMiner.jl:117 / StrobeGenomeMiner.jl:73,79
                                           : (int64_t)pl.k + ru.t_argmin;                  // GenomeMiner.jl:85,92
            reqs.push_back({ ru.record, ru.profile, std::max<int64_t>(CMI - P.buff, 1), std::min<int64_t>(CMI + wsq - 1 + P.buff, L), align_hint(ru.D_min, pl.tabs[(size_t)ru.profile].T) });
            of_run.push_back(i);
        }
        std::vector<AlignRes> ares;
        rc = align_batch_device(ctx, g, reqs, profiles, n_profiles, !pl.cluster, P.gap_open, P.gap_extend,
                                (P.flags & KGMA_F_TIE_OPEN) != 0, false, ares, nullptr, nullptr);
        if (rc) { result_release(res); return rc; }
        for (size_t j = 0; j < of_run.size(); j++) res->run_ext[of_run[j]] = kgma_run_ext{ ares[j].lo, ares[j].hi, ares[j].score };
        ctx->stats.n_align = (int64_t)reqs.size();
    }
    ctx->stats.n_runs = (int64_t)res->runs.size();
    const double dt = now_ms() - t0;
    ctx->stats.host_replay_ms = dt; ctx->stats.wall_ms += dt;
    *out = res;
    return KGMA_OK;
}

const kgma_run_ext *kgma_result_run_ext(const kgma_result *r) { return r && !r->run_ext.empty() ? r->run_ext.data() : nullptr; }

// Fixed-layout transport block of one shard's result: what the ranks exchange (one all-gather of equal-sized blocks).
struct PackHdr { uint32_t magic, version; int64_t n_runs, n_first, bytes; };
static const uint32_t PACK_MAGIC = 0x4b474d41u;   // "KGMA"

int64_t kgma_result_pack(const kgma_result *r, void *buf, int64_t cap)
{
    if (!r) return KGMA_E_ARG;
    const size_t n = r->runs.size(), nf = r->first_D.size();
    const bool with_ext = r->run_ext.size() == n;
    const int64_t need = (int64_t)(sizeof(PackHdr) + n * (sizeof(kgma_run) + sizeof(kgma_run_ext)) + nf * 8);
    if (!buf || cap < need) return need;
    unsigned char *p = (unsigned char *)buf;
    PackHdr h{ PACK_MAGIC, 1, (int64_t)n, (int64_t)nf, need };
    memcpy(p, &h, sizeof h); p += sizeof h;
    if (n) memcpy(p, r->runs.data(), n * sizeof(kgma_run));
    p += n * sizeof(kgma_run);
    if (n && with_ext) memcpy(p, r->run_ext.data(), n * sizeof(kgma_run_ext)); else if (n) memset(p, 0, n * sizeof(kgma_run_ext));
    p += n * sizeof(kgma_run_ext);
    if (nf) memcpy(p, r->first_D.data(), nf * 8);
    return need;
}

int kgma_replay_packed(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                       const kgma_scan_params *params, const void *blocks, int n_blocks, int64_t stride, kgma_result **out)
{
    // ctx may be NULL: with extension results in the blocks the replay is host-only
    if (!g || !params || !out || !blocks || n_blocks < 1 || stride < (int64_t)sizeof(PackHdr)) return KGMA_E_ARG;
    ScanPlan pl;
    kgma_scan_params P = *params; P.shard_index = 0; P.shard_count = 1;
    int rc = make_plan(ctx, g, profiles, n_profiles, P, pl);
    if (rc) return rc;
    const size_t nfd = (size_t)n_profiles * g->recs.size();
    kgma_result *res = result_acquire();
    res->runs.clear(); res->run_ext.clear();
    std::vector<int64_t> fd(nfd, INT64_MIN);
    for (int b = 0; b < n_blocks; b++) {
        const unsigned char *p = (const unsigned char *)blocks + (size_t)b * (size_t)stride;
        PackHdr h; memcpy(&h, p, sizeof h);
        if (h.magic != PACK_MAGIC || h.version != 1 || h.n_runs < 0 || h.bytes > stride || (size_t)h.n_first != nfd) {
            result_release(res);
            return set_err(ctx, KGMA_E_ARG, "shard block %d is not a kgma_result_pack block of this scan (or did not fit its %lld bytes)", b, (long long)stride);
        }
        const kgma_run *ru = (const kgma_run *)(p + sizeof h);
        const kgma_run_ext *ex = (const kgma_run_ext *)(ru + h.n_runs);
        const int64_t *f = (const int64_t *)(ex + h.n_runs);
        res->runs.insert(res->runs.end(), ru, ru + h.n_runs);
        res->run_ext.insert(res->run_ext.end(), ex, ex + h.n_runs);
        for (size_t i = 0; i < nfd; i++) fd[i] = std::max(fd[i], f[i]);     // INT64_MIN where a shard does not own window 0
    }
    res->first_D = fd;
    const double t0 = now_ms();
    rc = replay(ctx, g, pl.tabs, profiles, P, res->runs, fd, res, (P.flags & KGMA_F_ALIGN) ? &res->run_ext : nullptr);
    if (ctx) { ctx->stats.host_replay_ms = now_ms() - t0; }
    if (rc) { result_release(res); return rc; }
    *out = res;
    return KGMA_OK;
}

int kgma_replay(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                const kgma_scan_params *params, const kgma_run *runs, int64_t n_runs,
                const int64_t *first_window_D, kgma_result **out)
{
    // ctx may be NULL for a host-only replay (no KGMA_F_ALIGN): merging and replaying run lists needs no device
    if (!g || !params || !out || (n_runs && !runs) || !first_window_D) return KGMA_E_ARG;
    if (!ctx && (params->flags & KGMA_F_ALIGN)) return KGMA_E_ARG;
    ScanPlan pl;
    int rc = make_plan(ctx, g, profiles, n_profiles, *params, pl);
    if (rc) return rc;
    kgma_result *res = result_acquire();
    res->runs.assign(runs, runs + n_runs);
    std::vector<int64_t> fd(first_window_D, first_window_D + (size_t)n_profiles * g->recs.size());
    res->first_D = fd;
    rc = replay(ctx, g, pl.tabs, profiles, *params, res->runs, fd, res);
    if (rc) { result_release(res); return rc; }
    *out = res;
    return KGMA_OK;
}

int kgma_scan(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
              const kgma_scan_params *params, kgma_result **out)
{
    if (!ctx || !g || !params || !out) return KGMA_E_ARG;
    kgma_scan_params P = *params; P.shard_index = 0; P.shard_count = 1;
    kgma_result *res = result_acquire();
    ScanPlan pl;

    // Pipelined form of the streamed scan with extension: the records in front of a split point are replayed and extended
    // on a second stream while the tail of the genome is still being copied, so that only the last records' share of
    // replay + extension is left once the copy ends.  Records are independent in both state machines (GenomePos is a plain
    // running sum of record lengths), so the two parts' hit lists simply concatenate.
    PhaseHook hook;
    const int nr = (int)g->recs.size();
    const bool cluster = P.mode == KGMA_MODE_CLUSTER;
    const bool can_pipeline = g->sealed && P.mode != KGMA_MODE_STROBE && (cluster || n_profiles == 1) && (P.flags & KGMA_F_ALIGN) && P.only_record < 0 &&
                              !(P.flags & (KGMA_F_DENSE | KGMA_F_WANT_DISTS | KGMA_F_WANT_CIGARS)) && nr >= 2 && !getenv("KGMA_NO_PIPELINE");
    // A genome that is already on the device has no copy to hide work under, but the same two-part form lets the host replay
    // the first part's runs (and its extension batch run) next to the prefilter of the second: KGMA_RESIDENT_SPLIT = fraction
    // of the genome in the first part (0 / unset: one part).
    double res_split = 0;
    if (const char *e = getenv("KGMA_RESIDENT_SPLIT")) res_split = atof(e);
    const bool resident_now = (P.flags & KGMA_F_RESIDENT) && ctx->d_seq_valid && ctx->dg_uid == g->uid && ctx->d_valid_lo <= 0 &&
                              ctx->d_valid_hi >= g->G;
    if (can_pipeline && resident_now) {
        if (res_split > 0 && res_split < 1) {
            double best = 2;
            for (int r = 1; r < nr; r++) {
                const double frac = (double)g->recs[(size_t)r].off / (double)std::max<int64_t>(1, g->G);
                if (fabs(frac - res_split) < best) { best = fabs(frac - res_split); hook.rec_split = r; }
            }
            hook.resident = true;
        }
    } else if (can_pipeline) {
        // single mode queues one extension batch per part, so the first part should be as large as possible: split in front
        // of the last record that starts before 93% of the genome.  Cluster mode extends in rounds that block the host, a few
        // milliseconds for a whole genome: split earlier so that the first part's rounds fit under the rest of the copy.
        double hi = cluster ? 0.80 : 0.93;                        // (profiles/r2_micro4.py: cluster 16.06 ms at 0.72, 15.84 at 0.80, 16.37 at 0.90)
        if (const char *e = getenv("KGMA_SPLIT")) hi = atof(e);
        for (int r = 1; r < nr; r++) {
            const double frac = (double)g->recs[(size_t)r].off / (double)std::max<int64_t>(1, g->G);
            if (frac >= 0.5 && (hook.rec_split == 0 || frac <= hi)) hook.rec_split = r;
            if (frac > hi) break;
        }
    }
    std::vector<kgma_hit> hitsA; std::vector<AlignReq> reqsA, reqsB; std::vector<Pending> pendA, pendB;
    std::vector<kgma_run> runsA; AlignTicket tA, tB; int64_t genome_pos = 0, n_alignA = 0;
    auto reset = [&]() { hitsA.clear(); reqsA.clear(); pendA.clear(); runsA.clear(); genome_pos = 0; n_alignA = 0; if (tA.active) { std::vector<AlignRes> d; align_collect(ctx, &tA, d); } };
    hook.on_first_part = [&](std::vector<kgma_run> &ra, const std::vector<int64_t> &fd) -> int {
        if (cluster) {
            kgma_result part;
            ctx->s_extend = ctx->s_align;                 // the compute stream is busy with the tail of the genome
            int rc2 = replay(ctx, g, pl.tabs, profiles, P, ra, fd, &part);
            ctx->s_extend = nullptr;
            if (rc2) return rc2;
            hitsA.swap(part.hits); runsA.swap(ra); n_alignA = ctx->stats.n_align;
            return KGMA_OK;
        }
        merge_runs(ra);
        int rc2 = replay_single_range(ctx, g, pl.tabs[0], P, ra, fd, 0, hook.rec_split, &genome_pos, hitsA, reqsA, pendA);
        if (rc2) return rc2;
        runsA.swap(ra);
        return align_enqueue(ctx, g, reqsA, profiles, 1, true, P.gap_open, P.gap_extend, (P.flags & KGMA_F_TIE_OPEN) != 0, ctx->s_align, 0, &tA);
    };
    int rc = scan_runs_retry(ctx, g, profiles, n_profiles, P, pl, res, hook.rec_split > 0 ? &hook : nullptr, reset);
    if (rc == KGMA_OK) {
        const double t0 = now_ms();
        const double align_before = ctx->stats.align_ms;
        if (hook.used) {
            // second part: the runs reported after the first part's snapshot, restricted to its records (a dense re-evaluation
            // after a candidate overflow reports the first part's records again)
            std::vector<kgma_run> rb;
            for (size_t i = hook.n_runs_first; i < res->runs.size(); i++) if (res->runs[i].record >= hook.rec_split) rb.push_back(res->runs[i]);
            if (cluster) {
                kgma_result part;
                rc = replay(ctx, g, pl.tabs, profiles, P, rb, res->first_D, &part);
                if (rc == KGMA_OK) {
                    res->hits = hitsA; res->hits.insert(res->hits.end(), part.hits.begin(), part.hits.end());
                    res->runs = runsA; res->runs.insert(res->runs.end(), rb.begin(), rb.end());
                    ctx->stats.n_align += n_alignA;
                    ctx->stats.n_runs = (int64_t)res->runs.size();
                }
            } else {
                merge_runs(rb);
                std::vector<kgma_hit> hitsB;
                rc = replay_single_range(ctx, g, pl.tabs[0], P, rb, res->first_D, hook.rec_split, nr, &genome_pos, hitsB, reqsB, pendB);
                if (rc == KGMA_OK) rc = align_enqueue(ctx, g, reqsB, profiles, 1, true, P.gap_open, P.gap_extend, (P.flags & KGMA_F_TIE_OPEN) != 0, ctx->s_align, 1, &tB);
                std::vector<AlignRes> aA, aB;
                if (rc == KGMA_OK) rc = align_collect(ctx, &tA, aA);
                if (rc == KGMA_OK) rc = align_collect(ctx, &tB, aB);
                if (rc == KGMA_OK) {
                    apply_extensions(g, hitsA, pendA, aA);
                    apply_extensions(g, hitsB, pendB, aB);
                    res->hits = hitsA; res->hits.insert(res->hits.end(), hitsB.begin(), hitsB.end());
                    res->runs = runsA; res->runs.insert(res->runs.end(), rb.begin(), rb.end());
                    ctx->stats.n_align = (int64_t)(reqsA.size() + reqsB.size());
                    ctx->stats.n_runs = (int64_t)res->runs.size();
                } else reset();
            }
        } else rc = replay(ctx, g, pl.tabs, profiles, P, res->runs, res->first_D, res);
        const double dt = now_ms() - t0;
        if (getenv("KGMA_TRACE")) fprintf(stderr, "[kgma scan] replay after the scan: %.2f ms (%zu hits)\n", dt, res->hits.size());
        ctx->stats.host_replay_ms = dt - (ctx->stats.align_ms - align_before);
        ctx->stats.wall_ms += dt;
    } else reset();
    if (rc) { result_release(res); return rc; }
    *out = res;
    return KGMA_OK;
}

int kgma_genome_make_resident(kgma_ctx *ctx, kgma_genome *g)
{
    if (!ctx || !g) return KGMA_E_ARG;
    if (!g->sealed) return set_err(ctx, KGMA_E_STATE, "genome is not sealed");
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = dev_genome_prepare(ctx, g, false);
    if (rc) return rc;
    size_t bases = (size_t)(g->G + TAIL_PAD);
    if (!g->pinned && !getenv("KGMA_NO_STAGING")) {
        // pageable planes (FASTA / append ingest): through the staging ring, no page-locking (see StagedUpload)
        rc = StagedUpload::prepare_ring(ctx);
        if (rc) return rc;
        StagedUpload stager;
        stager.start(ctx, (const char *)g->seq2, bases / 4);
        rc = stager.issue(0, stager.nsub, (char *)ctx->d_seq2, ctx->s_copy);
        if (rc) return rc;
        KGMA_CUDA(ctx, cudaStreamSynchronize(ctx->s_copy));
    } else {
        rc = genome_pin(ctx, g);
        if (rc) return rc;
        KGMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_seq2, g->seq2, bases / 4, cudaMemcpyHostToDevice, ctx->s_compute));
        KGMA_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    }
    ctx->d_seq_valid = true; ctx->d_mask_valid = false; ctx->d_valid_lo = 0; ctx->d_valid_hi = g->G + TAIL_PAD;
    ctx->d_have_lo = 0; ctx->d_have_hi = g->G + TAIL_PAD;
    return KGMA_OK;
}

int kgma_genome_drop_resident(kgma_ctx *ctx, kgma_genome *g)
{
    if (!ctx || !g) return KGMA_E_ARG;
    if (ctx->dg_uid == g->uid) { ctx->d_seq_valid = ctx->d_mask_valid = false; }
    return KGMA_OK;
}

int kgma_genome_synth(kgma_ctx *ctx, int n_records, const int64_t *rec_len, uint64_t seed,
                      int64_t n_run_len, int64_t centromere_len, kgma_genome **out)
{
    if (!ctx || !out || n_records < 1 || !rec_len) return KGMA_E_ARG;
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    kgma_genome *g = nullptr; kgma_genome_create(&g);
    int64_t end = 0;
    for (int r = 0; r < n_records; r++) {
        kgma::Record R; R.ident = "synth" + std::to_string(r + 1); R.desc = R.ident + " synthetic contig"; R.len = rec_len[r];
        R.off = (end + REC_ALIGN - 1) / REC_ALIGN * REC_ALIGN; end = R.off + R.len;
        g->recs.push_back(R); g->total_len += R.len;
    }
    int64_t G = (end + FGROUP - 1) / FGROUP * FGROUP + FGROUP;
    int64_t cap = (G + TAIL_PAD + 4095) / 4096 * 4096;
    // page-locked planes from the start: this is the "pinned pre-packed host buffer" the e2e tier copies from
    if (cudaHostAlloc((void **)&g->seq2, (size_t)cap / 4, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&g->mask, (size_t)cap / 8, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError(); delete g;
        return set_err(ctx, KGMA_E_CUDA, "cudaHostAlloc of %lld bases failed", (long long)cap);
    }
    g->host_alloc = true; g->pinned = true; g->cap_bases = cap; g->G = G;
    memset(g->mask, 0, (size_t)cap / 8);
    int rc = dev_genome_prepare(ctx, g, false);
    if (rc) { kgma_genome_destroy(g); return rc; }
    int64_t nwords = cap / 16;
    kgma_synth<<<ctx->num_sms * 8, 256, 0, ctx->s_compute>>>(ctx->d_seq2, nwords, seed);
    KGMA_CUDA(ctx, cudaGetLastError());
    KGMA_CUDA(ctx, cudaMemcpyAsync(g->seq2, ctx->d_seq2, (size_t)cap / 4, cudaMemcpyDeviceToHost, ctx->s_compute));
    KGMA_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    // zero the padding between records and the tail, then lay down the N runs (N folds to T=3 + mask)
    auto fill = [&](int64_t lo, int64_t hi, int code, bool masked) {
        for (int64_t p = lo; p < hi; p++) {
            uint32_t &w = g->seq2[p >> 4]; int sh = 2 * (int)(p & 15);
            if ((p & 15) == 0 && p + 16 <= hi) { w = code ? 0xFFFFFFFFu : 0u; if (masked) { if ((p & 31) == 0) g->mask[p >> 5] = 0; g->mask[p >> 5] |= 0xFFFFu << (p & 31); } p += 15; continue; }
            w = (w & ~(3u << sh)) | ((uint32_t)code << sh);
            if (masked) g->mask[p >> 5] |= 1u << (p & 31);
        }
    };
    int64_t prev_end = 0;
    for (int r = 0; r < n_records; r++) {
        const auto &R = g->recs[r];
        fill(prev_end, R.off, 0, false);
        if (n_run_len > 0) {
            int64_t n = std::min(n_run_len, R.len / 4);
            fill(R.off, R.off + n, 3, true); fill(R.off + R.len - n, R.off + R.len, 3, true);
            if (n) g->any_mask = true;
        }
        if (centromere_len > 0 && R.len > 4 * centromere_len) {
            int64_t c0 = R.off + (R.len * 2 / 5) / 32 * 32;
            fill(c0, c0 + centromere_len, 3, true); g->any_mask = true;
        }
        prev_end = R.off + R.len;
    }
    fill(prev_end, cap, 0, false);
    g->sealed = true;
    ctx->d_seq_valid = ctx->d_mask_valid = false;           // device copy predates the host-side edits
    ctx->d_have_lo = ctx->d_have_hi = 0;
    *out = g;
    return KGMA_OK;
}

}  // extern "C"
