// exactMatch on the packed genome: every start position whose qlen symbols equal the query.
// Replaces exactMatch / FindAllOverlap / FindAll (src/ExactMatch.jl:89-121,33-43,20-30), i.e. repeated
// BioSequences.findfirst(ExactSearchQuery(query), view(seq, start:end)) calls, with one streaming pass:
// 2-bit codes are compared 16 bases (one 32-bit funnel-shifted word) at a time at every offset, survivors
// are verified over the whole query including the ambiguity plane (N equals only N), and match starts are
// appended with warp ballot/popc compaction.  Overlap / non-overlap selection is a host pass over the sorted
// starts (FindAll resumes at match_end+1, FindAllOverlap at match_start+1).
#include "kgma_internal.h"
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
        unsigned long long hits = 0;                     // bit o set: first 16 bases match at offset o
#pragma unroll
        for (int o = 0; o < 64; o++) {
            const int wi = o >> 4, s = (o & 15) * 2;
            uint32_t x = s ? __funnelshift_r(W[wi], W[wi + 1], s) : W[wi];
            if (((x ^ a.q0) & a.m0) == 0) hits |= 1ull << o;
        }
        // verify survivors (rare unless the query is low-complexity)
        unsigned long long ok = 0;
        while (hits) {
            int o = __ffsll((long long)hits) - 1; hits &= hits - 1;
            if (verify_match(a, blk * 64 + o)) ok |= 1ull << o;
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

}  // namespace kgma

using namespace kgma;

extern "C" int kgma_exact_match(kgma_ctx *ctx, kgma_genome *g, const char *query, int64_t qlen, int overlap,
                                uint32_t flags, kgma_match **out, int64_t *n_out)
{
    if (!ctx || !g || !query || !out || !n_out) return KGMA_E_ARG;
    *out = nullptr; *n_out = 0;
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
    int rc = dev_genome_prepare(ctx, g, true);
    if (rc) return rc;
    rc = genome_pin(ctx, g);
    if (rc) return rc;
    cudaStream_t st = ctx->s_compute;
    kgma_stats &S = ctx->stats; S = kgma_stats{};
    cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[1], e2 = ctx->ev[2];
    KGMA_CUDA(ctx, cudaEventRecord(e0, st));
    const size_t bases = (size_t)(g->G + TAIL_PAD);
    const bool have = (flags & KGMA_F_RESIDENT) && ctx->d_seq_valid && ctx->d_mask_valid && ctx->d_valid_lo == 0 && ctx->d_valid_hi >= (int64_t)bases;
    if (!have) {
        KGMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_seq2, g->seq2, bases / 4, cudaMemcpyHostToDevice, st));
        KGMA_CUDA(ctx, cudaMemcpyAsync(ctx->d_mask, g->mask, bases / 8, cudaMemcpyHostToDevice, st));
        S.h2d_bytes += bases / 4 + bases / 8;
        ctx->d_seq_valid = ctx->d_mask_valid = true; ctx->d_valid_lo = 0; ctx->d_valid_hi = (int64_t)bases;
        ctx->d_have_lo = 0; ctx->d_have_hi = (int64_t)bases;
    }
    KGMA_CUDA(ctx, cudaEventRecord(e1, st));
    const uint32_t cap = 1u << 24;
    size_t o = 0;
    auto carve = [&](size_t b) { size_t r = o; o += (b + 255) / 256 * 256; return r; };
    size_t o_q2 = carve(q2.size() * 4), o_qm = carve(qm.size() * 4), o_c = carve(256), o_out = carve((size_t)cap * 8);
    void *dv = nullptr;
    rc = dev_scratch(ctx, o, &dv);
    if (rc) return rc;
    unsigned char *d = (unsigned char *)dv;
    KGMA_CUDA(ctx, cudaMemcpyAsync(d + o_q2, q2.data(), q2.size() * 4, cudaMemcpyHostToDevice, st));
    KGMA_CUDA(ctx, cudaMemcpyAsync(d + o_qm, qm.data(), qm.size() * 4, cudaMemcpyHostToDevice, st));
    KGMA_CUDA(ctx, cudaMemsetAsync(d + o_c, 0, 256, st));
    MatchArgs a{};
    a.seq4 = (const uint4 *)ctx->d_seq2; a.seq = ctx->d_seq2; a.mask = ctx->d_mask;
    a.q2 = (const uint32_t *)(d + o_q2); a.qm = (const uint32_t *)(d + o_qm); a.qlen = (int)qlen;
    const int f = (int)std::min<int64_t>(qlen, 16);
    a.q0 = q2[0]; a.m0 = f == 16 ? 0xFFFFFFFFu : ((1u << (2 * f)) - 1);
    a.blk_begin = 0; a.blk_end = g->G / FBLOCK;
    a.out = (unsigned long long *)(d + o_out); a.out_cap = cap; a.out_count = (uint32_t *)(d + o_c);
    kgma_exact_match_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(a);
    KGMA_CUDA(ctx, cudaGetLastError());
    S.launches++;
    KGMA_CUDA(ctx, cudaEventRecord(e2, st));
    uint32_t cnt = 0;
    KGMA_CUDA(ctx, cudaMemcpyAsync(&cnt, d + o_c, 4, cudaMemcpyDeviceToHost, st));
    KGMA_CUDA(ctx, cudaStreamSynchronize(st));
    if (cnt > cap) return set_err(ctx, KGMA_E_CAPACITY, "more than %u exact matches", cap);
    std::vector<unsigned long long> pos(cnt);
    if (cnt) KGMA_CUDA(ctx, cudaMemcpy(pos.data(), d + o_out, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    S.d2h_bytes += 4 + (size_t)cnt * 8;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); S.h2d_ms = ms;
    cudaEventElapsedTime(&ms, e1, e2); S.filter_ms = ms;
    cudaEventElapsedTime(&ms, e0, e2); S.total_ms = ms;
    S.bases_scanned = g->total_len;
    if (!(flags & KGMA_F_RESIDENT)) ctx->d_seq_valid = ctx->d_mask_valid = false;
    std::sort(pos.begin(), pos.end());
    // map to records; keep matches wholly inside a record; apply FindAll's non-overlap rule per record
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
