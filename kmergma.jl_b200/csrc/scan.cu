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
//  (2) kgma_exact: the count-table kernel.  One chain per thread, each with its own 4^k x u16
//      count table in shared memory next to the profile's S table; the distance is updated
//      incrementally as one k-mer leaves and one enters the window, re-initialised per segment.
//      It emits run summaries (maximal stretches of D < T) with an atomic append, and optionally
//      every D (do_return_dists).  It runs on the candidate segments, or on everything in dense mode.
#include "kgma_internal.h"
#include <algorithm>
#include <cmath>
#include <climits>

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
            }
        }
        cur = nxt; nxt = nx2;
    }
}

// =============================================================================================
// (2) count-table kernel
// =============================================================================================
struct Segment {
    int64_t gpos;      // global base position of the first window of the segment
    int64_t dist_off;  // index into the D output for window w0 (step w0), or -1
    int32_t rec;
    int32_t w0;        // first window (0-based start in the record) == loop step
    int32_t n;         // number of windows
    int32_t last_step; // last loop step of the record (a run reaching it is unterminated)
};

struct ExactArgs {
    const uint32_t *seq;
    const int32_t  *S;          // [4^k] reversed-index profile sums
    const Segment  *segs;
    int       nseg;
    int      *next_seg;         // dynamic work counter
    int       k, nk, nch;       // chains (threads) per CTA that own a table
    int       profile;
    long long N2, twoN, sumS2, T, Tlo, Thi;
    kgma_run *runs; uint32_t run_cap; uint32_t *run_count;
    long long *first_D;         // [n_records] for this profile
    long long *dists;           // optional dense output of D per step
};


struct BitReader {
    const uint32_t *p; uint32_t lo, hi; int sh;
    __device__ __forceinline__ void init(const uint32_t *seq, int64_t pos)
    { p = seq + (pos >> 4); lo = p[0]; hi = p[1]; sh = (int)(pos & 15) * 2; }
    __device__ __forceinline__ uint32_t next(uint32_t kmask)
    {
        uint32_t v = __funnelshift_r(lo, hi, sh) & kmask;
        sh += 2;
        if (sh == 32) { sh = 0; ++p; lo = hi; hi = p[1]; }
        return v;
    }
};

__device__ __forceinline__ void emit_run(const ExactArgs &a, int rec, long long tf, long long tl, long long ta,
                                         long long dmin, uint32_t flags)
{
    uint32_t i = atomicAdd(a.run_count, 1u);
    if (i < a.run_cap) {
        kgma_run r; r.record = rec; r.profile = a.profile; r.t_first = tf; r.t_last = tl; r.t_argmin = ta;
        r.D_min = dmin; r.flags = flags; r.reserved = 0;
        a.runs[i] = r;
    }
}

__global__ void __launch_bounds__(256, 1) kgma_exact(ExactArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = 1 << (2 * a.k);
    int32_t *sS = reinterpret_cast<int32_t *>(smem_raw);
    uint16_t *cnt_all = reinterpret_cast<uint16_t *>(smem_raw + (size_t)nb * 4);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sS[i] = a.S[i];
    {   // zero every chain's table once; chains return their table to zero after each segment
        uint32_t *z = reinterpret_cast<uint32_t *>(cnt_all);
        int words = a.nch * nb / 2;
        for (int i = threadIdx.x; i < words; i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if ((int)threadIdx.x >= a.nch) return;
    uint16_t *cnt = cnt_all + (size_t)threadIdx.x * nb;
    const uint32_t kmask = (uint32_t)nb - 1;
    const int nk = a.nk;

    for (;;) {
        int si = atomicAdd(a.next_seg, 1);
        if (si >= a.nseg) break;
        const Segment sg = a.segs[si];
        BitReader R, L;
        R.init(a.seq, sg.gpos);
        L.init(a.seq, sg.gpos);
        long long Q = 0, A = 0;
        bool inrun = false; long long tf = 0, ta = 0, dmin = 0; uint32_t rflags = 0;
        // unified time loop: step u enters k-mer u (u < n+nk-1) and removes k-mer u-nk (u >= nk);
        // window (u-nk+1) is complete after the update when nk-1 <= u < n+nk-1.  The tail
        // (u >= n+nk-1) only removes, which returns the table to all-zero for the next segment.
        const int total = sg.n + 2 * nk - 1;
        for (int u = 0; u < total; ++u) {
            const bool hasR = u < sg.n + nk - 1, hasL = u >= nk;
            uint32_t r = 0, l = 0;
            if (hasR) r = R.next(kmask);
            if (hasL) l = L.next(kmask);
            if (!(hasR && hasL && l == r)) {            // GenomeMiner.jl:69 `if left_ind != right_ind`
                int cl = 0, cr = 0;
                if (hasL) cl = cnt[l];
                if (hasR) cr = cnt[r];
                if (hasL) { Q -= 2 * cl - 1; A -= sS[l]; cnt[l] = (uint16_t)(cl - 1); }
                if (hasR) { Q += 2 * cr + 1; A += sS[r]; cnt[r] = (uint16_t)(cr + 1); }
            }
            const int wi = u - (nk - 1);                // window index within the segment
            if (wi >= 0 && wi < sg.n) {
                const long long D = a.N2 * Q - a.twoN * A + a.sumS2;
                const long long t = (long long)sg.w0 + wi;      // loop step == 0-based window start
                if (t == 0) { a.first_D[sg.rec] = D; }
                else {
                    if (a.dists && sg.dist_off >= 0) a.dists[sg.dist_off + wi - (sg.w0 == 0 ? 1 : 0)] = D;
                    const bool near = (D >= a.Tlo) && (D < a.Thi);
                    if (D < a.T) {                       // GenomeMiner.jl:82 `kmerDist < thr`
                        if (!inrun) { inrun = true; tf = t; ta = t; dmin = D; rflags = (wi == 0 || (wi == 1 && sg.w0 == 0)) ? KGMA_RUN_OPEN_LEFT : 0; }
                        else if (D < dmin) { dmin = D; ta = t; rflags &= ~KGMA_HIT_ARGMIN_TIE; }
                        else if (D == dmin) rflags |= KGMA_HIT_ARGMIN_TIE;
                        if (near) rflags |= KGMA_HIT_NEAR_THR;
                    } else {
                        if (inrun) { emit_run(a, sg.rec, tf, t - 1, ta, dmin, rflags); inrun = false; }
                        if (near) emit_run(a, sg.rec, t, t, t, D, KGMA_RUN_MARKER | KGMA_HIT_NEAR_THR);
                    }
                    if (wi == sg.n - 1 && inrun) {       // segment ends inside a run: host merges with the neighbour
                        emit_run(a, sg.rec, tf, t, ta, dmin, rflags | KGMA_RUN_OPEN_RIGHT);
                        inrun = false;
                    }
                }
            }
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
        ctx->d_scratch = nullptr; ctx->d_scratch_bytes = 0;
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

int build_proftab(kgma_ctx *ctx, const kgma_profile &p, ProfTab &t)
{
    if (p.k < 1 || p.k > MAX_K) return set_err(ctx, KGMA_E_UNSUPPORTED, "k = %d is outside the supported range 1..%d", p.k, MAX_K);
    if (!p.S || p.n_refs <= 0) return set_err(ctx, KGMA_E_ARG, "profile needs integer sums S and n_refs > 0");
    if (p.k >= p.window) return set_err(ctx, KGMA_E_WINDOW, "the average reference sequence length %lld exceeds/is equal to the chosen kmer length %d. please reduce k. ", (long long)p.window, p.k);
    if (p.window - p.k + 1 > 60000) return set_err(ctx, KGMA_E_UNSUPPORTED, "window %lld too large for 16-bit count tables", (long long)p.window);
    if (!(p.thr >= 0) || !std::isfinite(p.thr)) return set_err(ctx, KGMA_E_ARG, "threshold must be finite and >= 0");
    t.k = p.k; t.ws = p.window; t.nk = p.window - p.k + 1; t.N = p.n_refs; t.thr = p.thr;
    size_t nb = (size_t)1 << (2 * p.k);
    t.S_rev.assign(nb, 0);
    __int128 s2 = 0;
    for (size_t c = 0; c < nb; c++) {
        if (p.S[c] < 0) return set_err(ctx, KGMA_E_ARG, "negative k-mer sum in profile");
        t.S_rev[rev_kmer((uint32_t)c, p.k)] = p.S[c];
        s2 += (__int128)p.S[c] * p.S[c];
    }
    t.N2 = (int64_t)p.n_refs * p.n_refs; t.twoN = 2LL * p.n_refs;
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
    return KGMA_OK;
}

int dev_genome_prepare(kgma_ctx *ctx, kgma_genome *g, bool need_mask)
{
    int64_t need = (g->G + TAIL_PAD + 4095) / 4096 * 4096;   // same rounding as the host planes (genome_reserve / kgma_genome_synth)
    if (ctx->dg_uid != g->uid || ctx->d_cap_bases < need) {
        if (ctx->d_seq2) cudaFree(ctx->d_seq2);
        if (ctx->d_mask) cudaFree(ctx->d_mask);
        ctx->d_seq2 = ctx->d_mask = nullptr; ctx->d_cap_bases = 0; ctx->dg_uid = 0;
        KGMA_CUDA(ctx, cudaMalloc(&ctx->d_seq2, (size_t)need / 4));
        KGMA_CUDA(ctx, cudaMalloc(&ctx->d_mask, (size_t)need / 8));
        ctx->d_cap_bases = need; ctx->dg_uid = g->uid;
        ctx->d_seq_valid = ctx->d_mask_valid = false; ctx->d_valid_lo = ctx->d_valid_hi = 0;
    }
    (void)need_mask;
    return KGMA_OK;
}

// ---------------------------------------------------------------------------------------------
struct ScanPlan {
    int C = 0, k = 0; int64_t maxws = 0, maxnk = 0;
    bool cluster = false;
    std::vector<ProfTab> tabs;
    // per record: number of loop steps (0 = record not scanned) — GenomeMiner.jl:60 / OmnGenomeMiner.jl:89
    std::vector<int64_t> steps;
};

static int make_plan(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int C, const kgma_scan_params &P, ScanPlan &pl)
{
    if (!profiles || C < 1 || C > MAX_PROFILES) return set_err(ctx, KGMA_E_ARG, "n_profiles must be 1..%d", MAX_PROFILES);
    if (P.mode == KGMA_MODE_SINGLE && C != 1) return set_err(ctx, KGMA_E_ARG, "single mode takes exactly one profile");
    pl.C = C; pl.k = profiles[0].k; pl.cluster = (P.mode == KGMA_MODE_CLUSTER);
    pl.tabs.resize(C);
    for (int q = 0; q < C; q++) {
        if (profiles[q].k != pl.k) return set_err(ctx, KGMA_E_ARG, "all profiles must share k");
        int rc = build_proftab(ctx, profiles[q], pl.tabs[q]);
        if (rc) return rc;
        pl.maxws = std::max(pl.maxws, pl.tabs[q].ws);
    }
    pl.maxnk = pl.maxws - pl.k + 1;
    if (pl.cluster && pl.k < 2) return set_err(ctx, KGMA_E_UNSUPPORTED, "cluster mode needs k >= 2 (the reference indexes past the record end for k = 1)");
    int nr = (int)g->recs.size();
    pl.steps.assign(nr, 0);
    for (int r = 0; r < nr; r++) {
        if (P.only_record >= 0 && r != P.only_record) continue;
        int64_t L = g->recs[r].len;
        int64_t st = pl.cluster ? (L - pl.maxws - pl.k + 2)       // view(seq, k:L-maxws+1)  OmnGenomeMiner.jl:89
                                : (L - pl.maxws);                   // zip(k:L-ws+k-1, ws+1:L) GenomeMiner.jl:60
        pl.steps[r] = std::max<int64_t>(0, st);
    }
    return KGMA_OK;
}

// fixed-point prefilter table; returns false when some profile cannot be filtered (R <= 0) -> dense mode
static bool build_filter_table(const ScanPlan &pl, std::vector<uint16_t> &tab8, int &M)
{
    const int k = pl.k; const size_t nb = (size_t)1 << (2 * k);
    if (k > 8) return false;
    M = (int)((pl.maxnk - 1 + FBLOCK - 1) / FBLOCK) + 1;
    if (M > 32 || M < 1) return false;
    std::vector<uint32_t> W(nb, 0);
    for (const ProfTab &t : pl.tabs) {
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
    const int nper = 9 - k; const uint32_t kmask = (uint32_t)nb - 1;
    tab8.resize(65536);
    for (uint32_t x = 0; x < 65536; x++) {
        uint32_t s = 0;
        for (int j = 0; j < nper; j++) s += W[(x >> (2 * j)) & kmask];
        tab8[x] = (uint16_t)std::min<uint32_t>(s, 65535u);
    }
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

struct Interval { int64_t lo, hi; };   // [lo,hi) global base positions of window starts

// Turn flagged 64-base blocks (or everything, in dense mode) into per-record window segments.
static void build_segments(const kgma_genome *g, const ScanPlan &pl, const std::vector<uint32_t> *cand /*sorted, unique*/,
                           int64_t pos_lo, int64_t pos_hi, int64_t seg_max, bool want_dists,
                           const std::vector<int64_t> &dist_base, std::vector<Segment> &segs)
{
    int nr = (int)g->recs.size();
    size_t ci = 0;
    for (int r = 0; r < nr; r++) {
        if (pl.steps[r] <= 0) continue;
        const int64_t off = g->recs[r].off;
        const int64_t vlo = std::max(off, pos_lo), vhi = std::min(off + pl.steps[r] + 1, pos_hi);   // windows 0..steps
        if (vlo >= vhi) { continue; }
        std::vector<Interval> iv;
        if (!cand) iv.push_back({ vlo, vhi });
        else {
            // the first window of every record is always evaluated (GenomeMiner.jl:42-47 initialises currminim from it)
            if (off >= pos_lo && off < pos_hi) iv.push_back({ off, std::min(off + 1, vhi) });
            while (ci < cand->size() && ((int64_t)(*cand)[ci] + 1) * FBLOCK <= vlo) ci++;
            size_t cj = ci;
            while (cj < cand->size() && (int64_t)(*cand)[cj] * FBLOCK < vhi) {
                int64_t lo = std::max<int64_t>((int64_t)(*cand)[cj] * FBLOCK, vlo);
                int64_t hi = std::min<int64_t>(((int64_t)(*cand)[cj] + 1) * FBLOCK, vhi);
                if (!iv.empty() && iv.back().hi >= lo) iv.back().hi = std::max(iv.back().hi, hi);
                else iv.push_back({ lo, hi });
                cj++;
            }
            // a block may straddle into the next record: do not consume the last one
            ci = (cj > ci) ? cj - 1 : cj;
        }
        for (const Interval &I : iv) {
            int64_t len = I.hi - I.lo;
            int64_t pieces = (len + seg_max - 1) / seg_max;
            for (int64_t p = 0; p < pieces; p++) {
                int64_t a = I.lo + len * p / pieces, b = I.lo + len * (p + 1) / pieces;
                Segment s;
                s.gpos = a; s.rec = r; s.w0 = (int32_t)(a - off); s.n = (int32_t)(b - a);
                s.last_step = (int32_t)pl.steps[r];
                s.dist_off = want_dists ? dist_base[r] + std::max<int64_t>(0, (a - off) - 1) : -1;
                segs.push_back(s);
            }
        }
    }
}

static int exact_chains(const kgma_ctx *ctx, int k, int *nch_out, size_t *smem_out)
{
    size_t nb = (size_t)1 << (2 * k);
    size_t avail = ctx->smem_optin;
    size_t sbytes = nb * 4;
    if (avail < sbytes + nb * 2) return KGMA_E_UNSUPPORTED;
    int nch = (int)std::min<size_t>((avail - sbytes) / (nb * 2), 256);
    *nch_out = nch; *smem_out = sbytes + (size_t)nch * nb * 2;
    return KGMA_OK;
}

}  // namespace kgma

using namespace kgma;

// The scan proper (one context / one GPU / one shard).  Fills res->runs, res->first_D, res->dists.
static int scan_runs_impl(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int C,
                          const kgma_scan_params &P, ScanPlan &pl, kgma_result *res)
{
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
    bool dense = (P.flags & KGMA_F_DENSE) != 0 || want_dists;

    // ---- shard range in 64-base blocks (whole warp groups)
    const int64_t nblk_total = g->G / FBLOCK;
    const int64_t ngrp_total = nblk_total / 32;
    int sc = std::max(1, P.shard_count), si = std::min(std::max(0, P.shard_index), sc - 1);
    const int64_t blk_lo = (ngrp_total * si / sc) * 32, blk_hi = (ngrp_total * (si + 1) / sc) * 32;
    const int64_t pos_lo = blk_lo * FBLOCK, pos_hi = blk_hi * FBLOCK;

    std::vector<uint16_t> tab8; int M = 0;
    if (!dense && !build_filter_table(pl, tab8, M)) dense = true;

    rc = dev_genome_prepare(ctx, g, false);
    if (rc) return rc;
    rc = genome_pin(ctx, g);
    if (rc) return rc;

    cudaStream_t sc_ = ctx->s_compute, sp = ctx->s_copy;
    cudaEvent_t e_start = ctx->ev[0], e_h2d = ctx->ev[1], e_filt = ctx->ev[2], e_exact = ctx->ev[3], e_fstart = ctx->ev[4];
    KGMA_CUDA(ctx, cudaEventRecord(e_start, sc_));

    // ---- upload range (bases): shard + halo, unless resident
    const int64_t halo = (int64_t)(dense ? pl.maxws + 64 : (M + 1) * FBLOCK + pl.maxws + 64);
    int64_t up_lo = pos_lo, up_hi = std::min(g->G + TAIL_PAD, pos_hi + halo + 3 * FGROUP);
    if (si == sc - 1) up_hi = g->G + TAIL_PAD;
    up_hi = (up_hi + 127) / 128 * 128; up_hi = std::min(up_hi, g->G + TAIL_PAD);
    const bool resident_ok = (P.flags & KGMA_F_RESIDENT) && ctx->d_seq_valid && ctx->d_valid_lo <= up_lo && ctx->d_valid_hi >= up_hi;

    // ---- device scratch layout
    const uint32_t cand_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>((blk_hi - blk_lo) / 16, 1 << 16), 1 << 24);
    const size_t nb = (size_t)1 << (2 * pl.k);
    size_t o = 0;
    auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    size_t o_tab = carve(65536 * 2), o_cnt = carve(256), o_cand = carve((size_t)cand_cap * 4), o_S = carve((size_t)C * nb * 4);
    size_t o_first = carve((size_t)C * nr * 8);
    const uint32_t run_cap = 1u << 20;
    size_t o_runs = carve((size_t)run_cap * sizeof(kgma_run));
    void *dsv = nullptr;
    // segments + dists are sized later; reserve generously for segments now
    const size_t seg_cap_bytes = (size_t)64 << 20;
    size_t o_segs = carve(seg_cap_bytes);
    rc = dev_scratch(ctx, o, &dsv);
    if (rc) return rc;
    unsigned char *ds = (unsigned char *)dsv;
    uint32_t *d_counters = (uint32_t *)(ds + o_cnt);      // [0]=cand_count [1]=run_count [2]=next_seg
    KGMA_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 256, sc_));
    {   // profile tables + first_D init
        std::vector<int32_t> Sall((size_t)C * nb);
        for (int q = 0; q < C; q++) memcpy(&Sall[(size_t)q * nb], pl.tabs[q].S_rev.data(), nb * 4);
        KGMA_CUDA(ctx, cudaMemcpyAsync(ds + o_S, Sall.data(), Sall.size() * 4, cudaMemcpyHostToDevice, sc_));
        std::vector<int64_t> fd((size_t)C * nr, INT64_MIN);
        KGMA_CUDA(ctx, cudaMemcpyAsync(ds + o_first, fd.data(), fd.size() * 8, cudaMemcpyHostToDevice, sc_));
        if (!dense) KGMA_CUDA(ctx, cudaMemcpyAsync(ds + o_tab, tab8.data(), 65536 * 2, cudaMemcpyHostToDevice, sc_));
        KGMA_CUDA(ctx, cudaStreamSynchronize(sc_));         // staging vectors go out of scope
        st.h2d_bytes += Sall.size() * 4 + fd.size() * 8 + (dense ? 0 : 65536 * 2);
    }

    // ---- stream the packed genome: chunked cudaMemcpyAsync on the copy stream, prefilter on the
    //      compute stream chasing it (double buffering falls out of the two streams + per-chunk events)
    const int64_t CH = (int64_t)128 << 20;                  // bases per chunk (32 MB of packed data)
    FilterArgs fa{};
    fa.seq = (const uint4 *)ctx->d_seq2; fa.tab = (const uint16_t *)(ds + o_tab); fa.M = M; fa.thrw = 1u << WFRAC;
    fa.cand = (uint32_t *)(ds + o_cand); fa.cand_cap = cand_cap; fa.cand_count = d_counters;
    const int fgrid = ctx->num_sms;
    int64_t done_blk = blk_lo;                               // target blocks already filtered
    const int64_t need_after = (int64_t)(M + 1) * FBLOCK + 3 * FGROUP;     // bases that must be present past a target block
    bool first_filter = true;
    auto run_filter_to = [&](int64_t avail_hi, bool last) -> int {
        if (dense) return KGMA_OK;
        int64_t lim = last ? blk_hi : std::min(blk_hi, ((avail_hi - need_after) / FBLOCK) / 32 * 32);
        if (lim <= done_blk) return KGMA_OK;
        if (first_filter) { KGMA_CUDA(ctx, cudaEventRecord(e_fstart, sc_)); first_filter = false; }
        fa.blk_begin = done_blk; fa.blk_end = lim;
        launch_filter_k(pl.k, fa, fgrid, sc_);
        KGMA_CUDA(ctx, cudaGetLastError());
        st.launches++;
        done_blk = lim;
        return KGMA_OK;
    };
    if (!resident_ok) {
        cudaEvent_t e_c[2] = { ctx->ev[5], ctx->ev[6] };
        KGMA_CUDA(ctx, cudaEventRecord(ctx->ev[7], sc_));
        KGMA_CUDA(ctx, cudaStreamWaitEvent(sp, ctx->ev[7], 0));
        int ci = 0;
        for (int64_t a = up_lo; a < up_hi; a += CH, ci++) {
            int64_t b = std::min(up_hi, a + CH);
            KGMA_CUDA(ctx, cudaMemcpyAsync((char *)ctx->d_seq2 + a / 4, (char *)g->seq2 + a / 4, (size_t)(b - a) / 4,
                                           cudaMemcpyHostToDevice, sp));
            KGMA_CUDA(ctx, cudaEventRecord(e_c[ci & 1], sp));
            KGMA_CUDA(ctx, cudaStreamWaitEvent(sc_, e_c[ci & 1], 0));
            st.h2d_bytes += (b - a) / 4;
            rc = run_filter_to(b, b >= up_hi);
            if (rc) return rc;
            if (ci >= 1) KGMA_CUDA(ctx, cudaEventSynchronize(e_c[(ci - 1) & 1]));   // bound the number of in-flight events reused
        }
        KGMA_CUDA(ctx, cudaEventRecord(e_h2d, sc_));
        ctx->d_seq_valid = true; ctx->d_valid_lo = up_lo; ctx->d_valid_hi = up_hi;
    } else {
        KGMA_CUDA(ctx, cudaEventRecord(e_h2d, sc_));
        rc = run_filter_to(up_hi, true);
        if (rc) return rc;
    }
    if (first_filter) KGMA_CUDA(ctx, cudaEventRecord(e_fstart, sc_));
    KGMA_CUDA(ctx, cudaEventRecord(e_filt, sc_));

    // ---- candidates -> segments
    std::vector<uint32_t> cand;
    if (!dense) {
        uint32_t ncand = 0;
        KGMA_CUDA(ctx, cudaMemcpyAsync(&ncand, d_counters, 4, cudaMemcpyDeviceToHost, sc_));
        KGMA_CUDA(ctx, cudaStreamSynchronize(sc_));
        st.d2h_bytes += 4;
        st.blocks_total = blk_hi - blk_lo; st.blocks_flagged = ncand;
        if (ncand > cand_cap) dense = true;                  // too many survivors: evaluate everything
        else {
            cand.resize(ncand);
            if (ncand) KGMA_CUDA(ctx, cudaMemcpy(cand.data(), ds + o_cand, (size_t)ncand * 4, cudaMemcpyDeviceToHost));
            st.d2h_bytes += (size_t)ncand * 4;
            std::sort(cand.begin(), cand.end());
        }
    }
    std::vector<int64_t> dist_base(nr, 0); int64_t ndist = 0;
    if (want_dists) for (int r = 0; r < nr; r++) { dist_base[r] = ndist; ndist += pl.steps[r]; }

    int nch = 0; size_t esmem = 0;
    rc = exact_chains(ctx, pl.k, &nch, &esmem);
    if (rc) return set_err(ctx, rc, "k = %d does not fit the shared-memory count tables", pl.k);
    const int eblock = std::min(256, (nch + 31) / 32 * 32);
    const int64_t total_chains = (int64_t)nch * ctx->num_sms;
    int64_t span = 0; for (int r = 0; r < nr; r++) span += pl.steps[r] ? pl.steps[r] + 1 : 0;
    int64_t seg_max = dense ? std::min<int64_t>(std::max<int64_t>(span / (total_chains * 8) + 1, 2048), 1 << 20) : 8192;
    std::vector<Segment> segs;
    build_segments(g, pl, dense ? nullptr : &cand, pos_lo, pos_hi, seg_max, want_dists, dist_base, segs);
    if (!dense && segs.size() * sizeof(Segment) > seg_cap_bytes) {          // survivors too fragmented: evaluate everything
        dense = true; segs.clear();
        seg_max = std::min<int64_t>(std::max<int64_t>(span / (total_chains * 8) + 1, 2048), 1 << 20);
        build_segments(g, pl, nullptr, pos_lo, pos_hi, seg_max, want_dists, dist_base, segs);
    }
    if (segs.size() * sizeof(Segment) > seg_cap_bytes) return set_err(ctx, KGMA_E_CAPACITY, "too many segments (%zu)", segs.size());
    int64_t exact_windows = 0; for (auto &s : segs) exact_windows += s.n;
    st.exact_windows = exact_windows * C;
    for (int r = 0; r < nr; r++) if (pl.steps[r]) {
        int64_t lo = std::max(g->recs[r].off, pos_lo), hi = std::min(g->recs[r].off + pl.steps[r] + 1, pos_hi);
        if (hi > lo) st.bases_scanned += hi - lo;
    }

    long long *d_dists = nullptr;
    if (want_dists && ndist > 0) KGMA_CUDA(ctx, cudaMalloc(&d_dists, (size_t)ndist * 8));
    res->dists.assign(C, {});
    res->first_D.assign((size_t)C * nr, INT64_MIN);
    std::vector<kgma_run> &runs = res->runs; runs.clear();
    if (!segs.empty()) {
        KGMA_CUDA(ctx, cudaMemcpyAsync(ds + o_segs, segs.data(), segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, sc_));
        st.h2d_bytes += segs.size() * sizeof(Segment);
        KGMA_CUDA(ctx, cudaFuncSetAttribute(kgma_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
        const int egrid = (int)std::min<int64_t>(ctx->num_sms, ((int64_t)segs.size() + nch - 1) / nch);
        for (int q = 0; q < C; q++) {
            const ProfTab &t = pl.tabs[q];
            ExactArgs ea{};
            ea.seq = ctx->d_seq2; ea.S = (const int32_t *)(ds + o_S) + (size_t)q * nb;
            ea.segs = (const Segment *)(ds + o_segs); ea.nseg = (int)segs.size();
            ea.next_seg = (int *)(d_counters + 2 + q);
            ea.k = pl.k; ea.nk = (int)t.nk; ea.nch = nch; ea.profile = q;
            ea.N2 = t.N2; ea.twoN = t.twoN; ea.sumS2 = t.sumS2; ea.T = t.T; ea.Tlo = t.Tlo; ea.Thi = t.Thi;
            ea.runs = (kgma_run *)(ds + o_runs); ea.run_cap = run_cap; ea.run_count = d_counters + 1;
            ea.first_D = (long long *)(ds + o_first) + (size_t)q * nr;
            ea.dists = d_dists;
            kgma_exact<<<egrid, eblock, esmem, sc_>>>(ea);
            KGMA_CUDA(ctx, cudaGetLastError());
            st.launches++;
            if (want_dists && ndist > 0) {                   // one profile's D at a time through the same buffer
                std::vector<long long> Dh((size_t)ndist);
                KGMA_CUDA(ctx, cudaMemcpyAsync(Dh.data(), d_dists, (size_t)ndist * 8, cudaMemcpyDeviceToHost, sc_));
                KGMA_CUDA(ctx, cudaStreamSynchronize(sc_));
                st.d2h_bytes += (size_t)ndist * 8;
                res->dists[q].resize((size_t)ndist);
                const double den = t.denom;
                for (int64_t i = 0; i < ndist; i++) res->dists[q][(size_t)i] = (double)Dh[(size_t)i] / den;
            }
        }
    }
    KGMA_CUDA(ctx, cudaEventRecord(e_exact, sc_));
    uint32_t cnts[3] = { 0, 0, 0 };
    KGMA_CUDA(ctx, cudaMemcpyAsync(cnts, d_counters, 8, cudaMemcpyDeviceToHost, sc_));
    KGMA_CUDA(ctx, cudaMemcpyAsync(res->first_D.data(), ds + o_first, (size_t)C * nr * 8, cudaMemcpyDeviceToHost, sc_));
    KGMA_CUDA(ctx, cudaStreamSynchronize(sc_));
    if (d_dists) cudaFree(d_dists);
    if (cnts[1] > run_cap) return set_err(ctx, KGMA_E_CAPACITY, "run list overflow (%u runs)", cnts[1]);
    runs.resize(cnts[1]);
    if (cnts[1]) KGMA_CUDA(ctx, cudaMemcpy(runs.data(), ds + o_runs, (size_t)cnts[1] * sizeof(kgma_run), cudaMemcpyDeviceToHost));
    st.d2h_bytes += 8 + (size_t)C * nr * 8 + (size_t)cnts[1] * sizeof(kgma_run);
    st.n_runs = cnts[1];
    float ms = 0;
    cudaEventElapsedTime(&ms, e_start, e_h2d); st.h2d_ms = ms;
    cudaEventElapsedTime(&ms, e_fstart, e_filt); st.filter_ms = ms;
    cudaEventElapsedTime(&ms, e_filt, e_exact); st.exact_ms = ms;
    cudaEventElapsedTime(&ms, e_start, e_exact); st.total_ms = ms;
    if (!(P.flags & KGMA_F_RESIDENT)) { ctx->d_seq_valid = false; }
    return KGMA_OK;
}

extern "C" {

int kgma_scan_runs(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                   const kgma_scan_params *params, kgma_result **out)
{
    if (!ctx || !g || !params || !out) return KGMA_E_ARG;
    kgma_result *res = new kgma_result();
    ScanPlan pl;
    int rc = scan_runs_impl(ctx, g, profiles, n_profiles, *params, pl, res);
    if (rc) { delete res; return rc; }
    *out = res;
    return KGMA_OK;
}

int kgma_replay(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
                const kgma_scan_params *params, const kgma_run *runs, int64_t n_runs,
                const int64_t *first_window_D, kgma_result **out)
{
    if (!ctx || !g || !params || !out || (n_runs && !runs) || !first_window_D) return KGMA_E_ARG;
    ScanPlan pl;
    int rc = make_plan(ctx, g, profiles, n_profiles, *params, pl);
    if (rc) return rc;
    kgma_result *res = new kgma_result();
    res->runs.assign(runs, runs + n_runs);
    std::vector<int64_t> fd(first_window_D, first_window_D + (size_t)n_profiles * g->recs.size());
    res->first_D = fd;
    rc = replay(ctx, g, pl.tabs, profiles, *params, res->runs, fd, res);
    if (rc) { delete res; return rc; }
    *out = res;
    return KGMA_OK;
}

int kgma_scan(kgma_ctx *ctx, kgma_genome *g, const kgma_profile *profiles, int n_profiles,
              const kgma_scan_params *params, kgma_result **out)
{
    if (!ctx || !g || !params || !out) return KGMA_E_ARG;
    kgma_scan_params P = *params; P.shard_index = 0; P.shard_count = 1;
    kgma_result *res = new kgma_result();
    ScanPlan pl;
    int rc = scan_runs_impl(ctx, g, profiles, n_profiles, P, pl, res);
    if (rc == KGMA_OK) rc = replay(ctx, g, pl.tabs, profiles, P, res->runs, res->first_D, res);
    if (rc) { delete res; return rc; }
    *out = res;
    return KGMA_OK;
}

int kgma_genome_make_resident(kgma_ctx *ctx, kgma_genome *g)
{
    if (!ctx || !g) return KGMA_E_ARG;
    if (!g->sealed) return set_err(ctx, KGMA_E_STATE, "genome is not sealed");
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = dev_genome_prepare(ctx, g, true);
    if (rc) return rc;
    rc = genome_pin(ctx, g);
    if (rc) return rc;
    size_t bases = (size_t)(g->G + TAIL_PAD);
    KGMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_seq2, g->seq2, bases / 4, cudaMemcpyHostToDevice, ctx->s_compute));
    KGMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_mask, g->mask, bases / 8, cudaMemcpyHostToDevice, ctx->s_compute));
    KGMA_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    ctx->d_seq_valid = ctx->d_mask_valid = true; ctx->d_valid_lo = 0; ctx->d_valid_hi = g->G + TAIL_PAD;
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
    int rc = dev_genome_prepare(ctx, g, true);
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
    *out = g;
    return KGMA_OK;
}

}  // extern "C"
