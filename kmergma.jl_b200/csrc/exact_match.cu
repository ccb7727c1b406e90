// exactMatch on the packed genome: every start position whose qlen symbols equal the query.
// Replaces exactMatch / FindAllOverlap / FindAll (src/ExactMatch.jl:89-121,33-43,20-30), i.e. repeated
// BioSequences.findfirst(ExactSearchQuery(query), view(seq, start:end)) calls, with one streaming pass:
//   * queries of >= 31 nt: kgma_exact_match_sampled - one hash probe per sampled aligned 16-base word (every
//     128 bases for queries >= 143 nt), survivors verified over the whole query on the 2-bit plane and against the
//     list of masked (N) runs; HBM-bound (one 4-byte load per 32-byte sector);
//   * shorter queries: kgma_exact_match_kernel - every offset, a zero-byte SIMD test of the first 4 bases (4
//     offsets per word operation), survivors verified on both planes (N equals only N), ballot/popc compaction.
// Overlap / non-overlap selection is a host pass over the sorted starts (FindAll resumes at match_end+1,
// FindAllOverlap at match_start+1).  The 2-bit plane is streamed in chunks with the search chasing the copy.
#include "kgma_internal.h"
#include "staged_upload.h"
#include <algorithm>

namespace kgma {

struct MatchArgs {
    const uint4    *seq4;       // packed genome, 64 bases per uint4
    const uint32_t *seq;        // same, as words
    const uint32_t *mask;       // 32 bases per word
    const uint32_t *q2;         // query packed 2-bit (padded with one zero word)
    const uint32_t *qm;         // query mask plane (padded)
    int       qlen;
    uint32_t  q0, m0;           // first min(qlen,16) bases and their bit mask
    int64_t   blk_begin, blk_end;   // 64-base blocks (multiples of 32)
    unsigned long long *out;    // match start positions (global base index)
    uint32_t  out_cap;
    uint32_t *out_count;
};

__device__ bool verify_match(const MatchArgs &a, int64_t p)
{
    // 2-bit plane
    const int nw = (a.qlen + 15) >> 4;
    const uint32_t *w = a.seq + (p >> 4); const int sh = (int)(p & 15) * 2;
    for (int i = 0; i < nw; i++) {
        uint32_t x = __funnelshift_r(w[i], w[i + 1], sh);
        uint32_t m = (i == nw - 1 && (a.qlen & 15)) ? ((1u << (2 * (a.qlen & 15))) - 1) : 0xFFFFFFFFu;
        if ((x ^ a.q2[i]) & m) return false;
    }
    // ambiguity plane: N (or any masked symbol) only equals an N at the same query position
    const int nm = (a.qlen + 31) >> 5;
    const uint32_t *mw = a.mask + (p >> 5); const int msh = (int)(p & 31);
    for (int i = 0; i < nm; i++) {
        uint32_t x = __funnelshift_r(mw[i], mw[i + 1], msh);
        uint32_t m = (i == nm - 1 && (a.qlen & 31)) ? ((1u << (a.qlen & 31)) - 1) : 0xFFFFFFFFu;
        if ((x ^ a.qm[i]) & m) return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) kgma_exact_match_kernel(MatchArgs a)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t grp0 = a.blk_begin >> 5, grp1 = a.blk_end >> 5;
    for (int64_t g = grp0 + warp; g < grp1; g += nwarps) {
        const int64_t blk = (g << 5) + lane;
        const uint4 v = __ldg(a.seq4 + blk);
        uint32_t nw = __shfl_down_sync(FULL, v.x, 1);
        if (lane == 31) nw = __ldg(a.seq + (blk + 1) * 4);
        const uint32_t W[5] = { v.x, v.y, v.z, v.w, nw };
        unsigned long long ok = 0;                       // bit o set: the query occurs at offset o of this block
        if (a.qlen >= 4) {
            // stage 1, SIMD within a register: a word shifted by r bases holds the 4-mers at offsets r, r+4, r+8, r+12 in its
            // four bytes; XOR with the query's first 4-mer replicated and the zero-byte test flag the candidates, 4 offsets
            // per ~5 integer instructions (borrows may flag the byte above a true zero: harmless, stage 2 re-checks).
            const uint32_t Qb = (a.q0 & 0xFFu) * 0x01010101u;
            uint32_t T[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t acc = 0;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t y = r ? __funnelshift_r(W[i], W[i + 1], 2 * r) : W[i];
                    const uint32_t z = y ^ Qb;
                    const uint32_t t = (z - 0x01010101u) & ~z & 0x80808080u;
                    acc |= t >> (7 - r);                 // bit 8b+r  <=>  offset 16i + 4b + r
                }
                T[i] = acc;
            }
            // stage 2 on the survivors (1 in 64 offsets for random DNA): first 16 bases, then the whole query on both planes
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t t = T[i];
                while (t) {
                    const int bit = __ffs((int)t) - 1; t &= t - 1;
                    const int off = (bit >> 3) * 4 + (bit & 7), o = 16 * i + off;      // off < 16: the offset stays in word i
                    const uint32_t x = __funnelshift_r(W[i], W[i + 1], 2 * off);
                    if (((x ^ a.q0) & a.m0) == 0 && verify_match(a, blk * 64 + o)) ok |= 1ull << o;
                }
            }
        } else {
            unsigned long long hits = 0;                 // bit o set: the (short) query's bases match at offset o
#pragma unroll
            for (int o = 0; o < 64; o++) {
                const int wi = o >> 4, s = (o & 15) * 2;
                uint32_t x = s ? __funnelshift_r(W[wi], W[wi + 1], s) : W[wi];
                if (((x ^ a.q0) & a.m0) == 0) hits |= 1ull << o;
            }
            while (hits) {
                int o = __ffsll((long long)hits) - 1; hits &= hits - 1;
                if (verify_match(a, blk * 64 + o)) ok |= 1ull << o;
            }
        }
        int n = __popcll(ok);
        // warp-level compaction: exclusive prefix of per-lane counts, one atomic per warp
        int pre = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += t; }
        int tot = __shfl_sync(FULL, pre, 31);
        if (tot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.out_count, (uint32_t)tot);
            base = __shfl_sync(FULL, base, 0);
            uint32_t pos = base + (uint32_t)(pre - n);
            while (ok) {
                int o = __ffsll((long long)ok) - 1; ok &= ok - 1;
                if (pos < a.out_cap) a.out[pos] = (unsigned long long)(blk * 64 + o);
                pos++;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sampled search for queries of >= 31 bases (the Horspool idea of BioSequences' own search, turned into a filter that
// fits a streaming GPU pass).  Take every SW-th aligned 16-base word of the genome (SW = 1, 2, 4 or 8 with
// qlen >= 16*SW + 15): an occurrence starting at o fully contains the first sampled word at or after o, at query
// offset j = p - o < 16*SW, so that genome word must be one of the 16*SW query 16-mers at offsets 0 .. 16*SW-1.
// One hash probe per sampled word (a 512-slot table in shared memory) replaces 16*SW shifted compares; the rare
// survivors are verified over the whole query on the 2-bit plane and against the list of masked (N) runs.
// Each occurrence is found exactly once (through its first sampled word).
struct SampledArgs {
    const uint32_t *seq;
    const uint32_t *tab_key; const uint16_t *tab_j;   // [512] open-addressing table: 16-mer -> query offset (0xFFFF = empty)
    const uint32_t *q2, *qm;                          // packed query planes (padded with one zero word)
    const long long *nruns; int n_nruns;              // maximal masked runs of the genome, [start,end) ascending
    int qlen, sw, q_has_n;
    long long t_begin, t_end;                         // sampled-word indices to test
    long long g_end;                                  // end of the packed genome (bases): no occurrence may reach past it
    unsigned long long *out; uint32_t out_cap; uint32_t *out_count;
};

__device__ bool verify_sampled(const SampledArgs &a, long long p)
{
    const int nw = (a.qlen + 15) >> 4;
    const uint32_t *w = a.seq + (p >> 4); const int sh = (int)(p & 15) * 2;
    for (int i = 0; i < nw; i++) {
        uint32_t x = __funnelshift_r(__ldg(w + i), __ldg(w + i + 1), sh);
        uint32_t m = (i == nw - 1 && (a.qlen & 15)) ? ((1u << (2 * (a.qlen & 15))) - 1) : 0xFFFFFFFFu;
        if ((x ^ a.q2[i]) & m) return false;
    }
    // ambiguity: a masked genome base only equals a masked (N) query base and vice versa
    int lo = -1, hi = a.n_nruns;                       // last run with start < p + qlen
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (a.nruns[2 * mid] < p + a.qlen) lo = mid; else hi = mid; }
    const bool genome_masked_here = lo >= 0 && a.nruns[2 * lo + 1] > p;     // some run intersects [p, p+qlen)
    if (!genome_masked_here) return !a.q_has_n;
    for (int i = 0; i < a.qlen; i++) {                 // rare: a run touches the window, compare base by base
        const long long gp = p + i;
        int l2 = -1, h2 = a.n_nruns;
        while (h2 - l2 > 1) { int mid = (l2 + h2) >> 1; if (a.nruns[2 * mid] <= gp) l2 = mid; else h2 = mid; }
        const bool gm = l2 >= 0 && gp < a.nruns[2 * l2 + 1];
        const bool qmk = (a.qm[i >> 5] >> (i & 31)) & 1u;
        if (gm != qmk) return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) kgma_exact_match_sampled(SampledArgs a)
{
    __shared__ uint32_t s_key[512];
    __shared__ uint16_t s_j[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { s_key[i] = a.tab_key[i]; s_j[i] = a.tab_j[i]; }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = a.t_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.t_end; t += stride) {
        const uint32_t w = __ldg(a.seq + t * a.sw);
        uint32_t h = (w * 2654435761u) >> 23;          // 9-bit multiplicative hash
        while (s_j[h] != 0xFFFFu) {
            if (s_key[h] == w) {
                const long long start = t * a.sw * 16 - (long long)s_j[h];
                if (start >= 0 && start + a.qlen <= a.g_end && verify_sampled(a, start)) {
                    const uint32_t pos = atomicAdd(a.out_count, 1u);
                    if (pos < a.out_cap) a.out[pos] = (unsigned long long)start;
                }
            }
            h = (h + 1) & 511u;
        }
    }
}

}  // namespace kgma

using namespace kgma;

// Search the slice shard_index of shard_count of the packed genome: every occurrence is reported by exactly one slice (the
// one that owns its first sampled word / its start block), so the slices' lists simply concatenate.  The slice needs
// qlen - 1 bases past its end (an occurrence starting on its last base) and up to 127 before its start.
static int exact_match_impl(kgma_ctx *ctx, kgma_genome *g, const char *query, int64_t qlen, uint32_t flags,
                            int shard_index, int shard_count, std::vector<unsigned long long> &pos)
{
    if (!g->sealed) return set_err(ctx, KGMA_E_STATE, "genome is not sealed");
    if (qlen < 1 || qlen > 1 << 20) return set_err(ctx, KGMA_E_ARG, "query length %lld out of range", (long long)qlen);
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    // pack the query like the genome
    const int nw = (int)((qlen + 15) / 16), nm = (int)((qlen + 31) / 32);
    std::vector<uint32_t> q2((size_t)nw + 1, 0), qm((size_t)nm + 1, 0);
    for (int64_t i = 0; i < qlen; i++) {
        int code; bool masked = false;
        switch (query[i]) {
        case 'A': case 'a': code = 0; break; case 'C': case 'c': code = 1; break;
        case 'G': case 'g': code = 2; break; case 'T': case 't': code = 3; break;
        case 'N': case 'n': code = 3; masked = true; break;
        default: return set_err(ctx, KGMA_E_SYMBOL, "query symbol '%c' at %lld is outside A,C,G,T,N", query[i], (long long)(i + 1));
        }
        q2[(size_t)(i >> 4)] |= (uint32_t)code << (2 * (i & 15));
        if (masked) qm[(size_t)(i >> 5)] |= 1u << (i & 31);
    }
    int rc = dev_genome_prepare(ctx, g, qlen < 31);     // the ambiguity plane is only needed on the device for short queries
    if (rc) return rc;
    cudaStream_t st = ctx->s_compute, sp = ctx->s_copy;
    kgma_stats &S = ctx->stats; S = kgma_stats{};
    cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[1], e2 = ctx->ev[2], ek = ctx->ev[3];
    const size_t bases = (size_t)(g->G + TAIL_PAD);
    const int sc = std::max(1, shard_count), si = std::min(std::max(0, shard_index), sc - 1);
    const int64_t ngrp_total = g->G / FGROUP;
    const int64_t range_lo = (ngrp_total * si / sc) * FGROUP, range_hi = si == sc - 1 ? g->G : (ngrp_total * (si + 1) / sc) * FGROUP;
    const int64_t up_lo = sc == 1 ? 0 : std::max<int64_t>(0, range_lo - FGROUP);
    const int64_t up_hi = sc == 1 || si == sc - 1 ? (int64_t)bases : std::min<int64_t>((int64_t)bases, (range_hi + qlen + 6 * FGROUP + 127) / 128 * 128);
    const bool have = (flags & KGMA_F_RESIDENT) && ctx->d_seq_valid && ctx->d_valid_lo <= up_lo && ctx->d_valid_hi >= up_hi;
    // a resident genome is not read from the host at all; a pageable one goes through the staging ring (scan.cu)
    const bool staged = !have && !g->pinned && !getenv("KGMA_NO_STAGING");
    if (!have) { rc = staged ? StagedUpload::prepare_ring(ctx) : genome_pin(ctx, g); if (rc) return rc; }
    // sampling stride in 16-base words: the largest of 8,4,2,1 with qlen >= 16*sw + 15 (0: short query, dense compare)
    int sw = 0;
    for (int c = 8; c >= 1; c >>= 1) if (qlen >= 16 * c + 15) { sw = c; break; }
    const std::vector<int64_t> &nruns = genome_nruns(g);
    const bool need_mask_plane = sw == 0;               // only the dense kernel reads the ambiguity plane
    bool q_has_n = false; for (uint32_t w : qm) q_has_n |= w != 0;
    // hash table of the query 16-mers at offsets 0 .. 16*sw-1
    std::vector<uint32_t> tkey(512, 0); std::vector<uint16_t> tj(512, 0xFFFFu);
    for (int j = 0; j < 16 * sw; j++) {
        const uint32_t w = (uint32_t)((((uint64_t)q2[(size_t)(j >> 4) + 1] << 32 | q2[(size_t)(j >> 4)]) >> (2 * (j & 15))) & 0xFFFFFFFFu);
        uint32_t h = (w * 2654435761u) >> 23;
        while (tj[h] != 0xFFFFu) h = (h + 1) & 511u;
        tkey[h] = w; tj[h] = (uint16_t)j;
    }
    const uint32_t cap = 1u << 24;
    size_t o = 0;
    auto carve = [&](size_t b) { size_t r = o; o += (b + 255) / 256 * 256; return r; };
    size_t o_q2 = carve(q2.size() * 4), o_qm = carve(qm.size() * 4), o_tk = carve(512 * 4), o_tj = carve(512 * 2), o_nr = carve(nruns.size() * 8 + 8);
    const size_t up = o;
    size_t o_c = carve(256), o_out = carve((size_t)cap * 8);
    void *dv = nullptr, *hv = nullptr;
    rc = dev_scratch(ctx, o, &dv);
    ctx->tab_sig = 0;                                  // (the arena is carved anew: the scan's resident tables are gone)
    if (rc) return rc;
    rc = host_scratch(ctx, up, &hv);
    if (rc) return rc;
    unsigned char *d = (unsigned char *)dv, *h = (unsigned char *)hv;
    memcpy(h + o_q2, q2.data(), q2.size() * 4); memcpy(h + o_qm, qm.data(), qm.size() * 4);
    memcpy(h + o_tk, tkey.data(), 512 * 4); memcpy(h + o_tj, tj.data(), 512 * 2);
    if (!nruns.empty()) memcpy(h + o_nr, nruns.data(), nruns.size() * 8);
    KGMA_CUDA(ctx, cudaEventRecord(e0, st));
    KGMA_CUDA(ctx, cudaMemcpyAsync(d, h, up, cudaMemcpyHostToDevice, st));
    KGMA_CUDA(ctx, cudaMemsetAsync(d + o_c, 0, 256, st));
    S.h2d_bytes += up;

    MatchArgs a{};
    a.seq4 = (const uint4 *)ctx->d_seq2; a.seq = ctx->d_seq2; a.mask = ctx->d_mask;
    a.q2 = (const uint32_t *)(d + o_q2); a.qm = (const uint32_t *)(d + o_qm); a.qlen = (int)qlen;
    const int f = (int)std::min<int64_t>(qlen, 16);
    a.q0 = q2[0]; a.m0 = f == 16 ? 0xFFFFFFFFu : ((1u << (2 * f)) - 1);
    a.out = (unsigned long long *)(d + o_out); a.out_cap = cap; a.out_count = (uint32_t *)(d + o_c);
    SampledArgs sa{};
    sa.seq = ctx->d_seq2; sa.tab_key = (const uint32_t *)(d + o_tk); sa.tab_j = (const uint16_t *)(d + o_tj);
    sa.q2 = a.q2; sa.qm = a.qm; sa.nruns = (const long long *)(d + o_nr); sa.n_nruns = (int)(nruns.size() / 2);
    sa.qlen = (int)qlen; sa.sw = sw; sa.q_has_n = q_has_n ? 1 : 0; sa.g_end = g->G;
    sa.out = a.out; sa.out_cap = cap; sa.out_count = a.out_count;

    // search [done, upto) in bases (multiples of 2048), given that the genome is on the device up to avail_hi
    const int64_t G = range_hi;
    int64_t done = range_lo; bool first_kernel = true;
    auto search_to = [&](int64_t avail_hi, bool last) -> int {
        int64_t upto = last ? G : std::min<int64_t>(G, (avail_hi - qlen - 4 * FGROUP) / FGROUP * FGROUP);
        if (upto <= done) return KGMA_OK;
        if (first_kernel) { KGMA_CUDA(ctx, cudaEventRecord(ek, st)); first_kernel = false; }
        if (sw > 0) {
            sa.t_begin = done / (16 * sw); sa.t_end = upto / (16 * sw);
            const long long nt = sa.t_end - sa.t_begin;
            const int grid = (int)std::min<long long>((nt + 255) / 256, (long long)ctx->num_sms * 16);
            kgma_exact_match_sampled<<<std::max(grid, 1), 256, 0, st>>>(sa);
        } else {
            a.blk_begin = done / FBLOCK; a.blk_end = upto / FBLOCK;
            kgma_exact_match_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(a);
        }
        KGMA_CUDA(ctx, cudaGetLastError());
        S.launches++;
        done = upto;
        return KGMA_OK;
    };
    if (!have) {
        // stream the 2-bit plane in 32 MB chunks on the copy stream, the search chasing it (the ambiguity plane is only
        // uploaded for short queries; long ones check N against the list of masked runs)
        const bool whole = up_lo == 0 && up_hi == (int64_t)bases;
        if (need_mask_plane) {
            KGMA_CUDA(ctx, cudaMemcpyAsync((char *)ctx->d_mask + up_lo / 8, (char *)g->mask + up_lo / 8, (size_t)(up_hi - up_lo) / 8, cudaMemcpyHostToDevice, st));
            S.h2d_bytes += (up_hi - up_lo) / 8; ctx->d_mask_valid = whole;
        }
        cudaEvent_t e_c[2] = { ctx->ev[5], ctx->ev[6] };
        KGMA_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
        KGMA_CUDA(ctx, cudaStreamWaitEvent(sp, ctx->ev[7], 0));
        const int64_t CH = (int64_t)128 << 20; int ci = 0;
        StagedUpload stager;                                       // (its destructor joins the copy threads on every exit path)
        if (staged) stager.start(ctx, (const char *)g->seq2 + up_lo / 4, (size_t)(up_hi - up_lo) / 4);
        for (int64_t lo = up_lo; lo < up_hi; lo += CH, ci++) {
            const int64_t hi = std::min<int64_t>(up_hi, lo + CH);
            if (staged) {
                const size_t j0 = (size_t)((lo - up_lo) / 4) / StagedUpload::SB;
                const size_t j1 = hi >= up_hi ? stager.nsub : (size_t)((hi - up_lo) / 4) / StagedUpload::SB;
                rc = stager.issue(j0, j1, (char *)ctx->d_seq2 + up_lo / 4, sp);
                if (rc) return rc;
            } else
            KGMA_CUDA(ctx, cudaMemcpyAsync((char *)ctx->d_seq2 + lo / 4, (char *)g->seq2 + lo / 4, (size_t)(hi - lo) / 4, cudaMemcpyHostToDevice, sp));
            KGMA_CUDA(ctx, cudaEventRecord(e_c[ci & 1], sp));
            KGMA_CUDA(ctx, cudaStreamWaitEvent(st, e_c[ci & 1], 0));
            S.h2d_bytes += (hi - lo) / 4;
            rc = search_to(hi, hi >= up_hi);
            if (rc) return rc;
            if (ci >= 1) KGMA_CUDA(ctx, cudaEventSynchronize(e_c[(ci - 1) & 1]));
        }
        ctx->d_seq_valid = true; ctx->d_valid_lo = up_lo; ctx->d_valid_hi = up_hi;
        ctx->d_have_lo = up_lo; ctx->d_have_hi = up_hi;
        KGMA_CUDA(ctx, cudaEventRecord(e1, st));
    } else {
        if (need_mask_plane && !ctx->d_mask_valid) {
            KGMA_CUDA(ctx, cudaMemcpyAsync((char *)ctx->d_mask + up_lo / 8, (char *)g->mask + up_lo / 8, (size_t)(up_hi - up_lo) / 8, cudaMemcpyHostToDevice, st));
            S.h2d_bytes += (up_hi - up_lo) / 8; ctx->d_mask_valid = up_lo == 0 && up_hi == (int64_t)bases;
        }
        KGMA_CUDA(ctx, cudaEventRecord(e1, st));
        rc = search_to(up_hi, true);
        if (rc) return rc;
    }
    if (first_kernel) KGMA_CUDA(ctx, cudaEventRecord(ek, st));
    KGMA_CUDA(ctx, cudaEventRecord(e2, st));
    uint32_t cnt = 0;
    KGMA_CUDA(ctx, cudaMemcpyAsync(&cnt, d + o_c, 4, cudaMemcpyDeviceToHost, st));
    KGMA_CUDA(ctx, cudaStreamSynchronize(st));
    if (cnt > cap) return set_err(ctx, KGMA_E_CAPACITY, "more than %u exact matches", cap);
    pos.resize(cnt);
    if (cnt) KGMA_CUDA(ctx, cudaMemcpy(pos.data(), d + o_out, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    S.d2h_bytes += 4 + (size_t)cnt * 8;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); S.h2d_ms = ms;
    cudaEventElapsedTime(&ms, ek, e2); S.filter_ms = ms;
    cudaEventElapsedTime(&ms, e0, e2); S.total_ms = ms;
    S.bases_scanned = sc == 1 ? g->total_len : range_hi - range_lo;
    if (!(flags & KGMA_F_RESIDENT)) ctx->d_seq_valid = ctx->d_mask_valid = false;
    std::sort(pos.begin(), pos.end());
    return KGMA_OK;
}

// map sorted global starts to records; keep matches wholly inside a record; apply FindAll's non-overlap rule per record
// (ExactMatch.jl:20-30: the search resumes behind the previous match; FindAllOverlap :33-43 keeps every start)
static int matches_from_starts(kgma_ctx *ctx, const kgma_genome *g, const std::vector<unsigned long long> &pos, int64_t qlen, int overlap,
                               kgma_match **out, int64_t *n_out)
{
    std::vector<kgma_match> res;
    size_t pi = 0;
    for (int r = 0; r < (int)g->recs.size(); r++) {
        const kgma::Record &R = g->recs[r];
        while (pi < pos.size() && (int64_t)pos[pi] < R.off) pi++;
        int64_t next_ok = 0;                            // 1-based start a non-overlapping match may begin at
        while (pi < pos.size() && (int64_t)pos[pi] < R.off + R.len) {
            int64_t s1 = (int64_t)pos[pi] - R.off + 1;   // 1-based start
            pi++;
            if (s1 + qlen - 1 > R.len) continue;
            if (!overlap && s1 < next_ok) continue;
            res.push_back({ r, 0, s1, s1 + qlen - 1 });
            next_ok = s1 + qlen;
        }
    }
    *n_out = (int64_t)res.size();
    if (!res.empty()) {
        *out = (kgma_match *)malloc(res.size() * sizeof(kgma_match));
        if (!*out) return set_err(ctx, KGMA_E_CAPACITY, "out of memory");
        memcpy(*out, res.data(), res.size() * sizeof(kgma_match));
    }
    return KGMA_OK;
}

extern "C" int kgma_exact_match(kgma_ctx *ctx, kgma_genome *g, const char *query, int64_t qlen, int overlap,
                                uint32_t flags, kgma_match **out, int64_t *n_out)
{
    if (!ctx || !g || !query || !out || !n_out) return KGMA_E_ARG;
    *out = nullptr; *n_out = 0;
    std::vector<unsigned long long> pos;
    int rc = exact_match_impl(ctx, g, query, qlen, flags, 0, 1, pos);
    if (rc) return rc;
    return matches_from_starts(ctx, g, pos, qlen, overlap, out, n_out);
}

// Multi-GPU form: each context searches its slice and returns the occurrence starts it owns (0-based positions in the packed
// coordinate space, ascending; *starts is malloc'd, kgma_free); kgma_exact_match_merge, host only, turns the concatenation of
// all slices' starts (any order) into what exactMatch returns, applying the overlap rule across slice edges.
extern "C" int kgma_exact_match_shard(kgma_ctx *ctx, kgma_genome *g, const char *query, int64_t qlen, uint32_t flags,
                                      int shard_index, int shard_count, int64_t **starts, int64_t *n_out)
{
    if (!ctx || !g || !query || !starts || !n_out || shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return KGMA_E_ARG;
    *starts = nullptr; *n_out = 0;
    std::vector<unsigned long long> pos;
    int rc = exact_match_impl(ctx, g, query, qlen, flags, shard_index, shard_count, pos);
    if (rc) return rc;
    *n_out = (int64_t)pos.size();
    if (!pos.empty()) {
        *starts = (int64_t *)malloc(pos.size() * 8);
        if (!*starts) return set_err(ctx, KGMA_E_CAPACITY, "out of memory");
        memcpy(*starts, pos.data(), pos.size() * 8);
    }
    return KGMA_OK;
}

extern "C" int kgma_exact_match_merge(const kgma_genome *g, const int64_t *starts, int64_t n, int64_t qlen, int overlap,
                                      kgma_match **out, int64_t *n_out)
{
    if (!g || !out || !n_out || n < 0 || (n && !starts) || qlen < 1) return KGMA_E_ARG;
    *out = nullptr; *n_out = 0;
    std::vector<unsigned long long> pos((size_t)n);
    for (int64_t i = 0; i < n; i++) pos[(size_t)i] = (unsigned long long)starts[i];
    std::sort(pos.begin(), pos.end());
    pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
    return matches_from_starts(nullptr, g, pos, qlen, overlap, out, n_out);
}
