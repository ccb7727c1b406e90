// Host-side genome container: FASTA parsing, 2-bit packing + ambiguity mask, record table.
// Replaces FASTX.FASTA.Reader + getSeq (src/Consts.jl:37-39) + the per-base NUCLEOTIDE_BITS
// dictionary lookups (src/Consts.jl:22-28) of the reference's hot loops.
#include "kgma_internal.h"
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <thread>
#include <algorithm>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <chrono>
#if defined(__SSE2__)
#include <emmintrin.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#endif

namespace kgma {

static thread_local std::string g_create_err;

int set_err(kgma_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}
const char *create_err() { return g_create_err.c_str(); }

// code table: low 2 bits = 2-bit code, bit 2 = masked (N or other IUPAC), bit 3 = not N (ambiguous IUPAC), 0xFF = invalid
struct CodeTab {
    uint8_t t[256];
    CodeTab()
    {
        memset(t, 0xFF, sizeof t);
        auto set = [&](char c, uint8_t v) { t[(uint8_t)c] = v; t[(uint8_t)(c | 0x20)] = v; };
        set('A', 0); set('C', 1); set('G', 2); set('T', 3);
        set('N', 3 | 4);                                       // Consts.jl:27  N -> 3
        for (const char *p = "RYSWKMBDHVU"; *p; ++p) set(*p, 0 | 4 | 8);   // other IUPAC: masked + ambiguous
        t[(uint8_t)'-'] = 0 | 4 | 8;
    }
};
static const CodeTab CT;

// The planes live in anonymous mappings marked MADV_HUGEPAGE: pages arrive zeroed and lazily, so a bulk ingest faults
// them in from its packing threads instead of one memset, and with 2 MB pages both those faults and the later
// cudaHostRegister (which pins page by page) cost a fraction of what they do on 4 KB pages (measured on the B200 box:
// registering 62 MB 39 ms -> 3.6 ms).  Growth is mremap: contents kept, new pages zero.
static void *plane_map(void *old, size_t old_bytes, size_t bytes)
{
    void *p = old ? mremap(old, old_bytes, bytes, MREMAP_MAYMOVE)
                  : mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    madvise(p, bytes, MADV_HUGEPAGE);                      // advisory: a kernel without THP just keeps 4 KB pages
    return p;
}

int genome_reserve(kgma_genome *g, int64_t bases)
{
    if (bases <= g->cap_bases) return KGMA_OK;
    int64_t nc = std::max<int64_t>(bases, g->cap_bases + g->cap_bases / 2);
    nc = (nc + 4095) / 4096 * 4096;
    if (g->pinned || g->host_alloc) return KGMA_E_STATE;
    void *s = plane_map(g->seq2, g->seq_map, (size_t)nc / 4);
    if (!s) return KGMA_E_CAPACITY;
    g->seq2 = (uint32_t *)s; g->seq_map = (size_t)nc / 4;
    void *m = plane_map(g->mask, g->mask_map, (size_t)nc / 8);
    if (!m) return KGMA_E_CAPACITY;
    g->mask = (uint32_t *)m; g->mask_map = (size_t)nc / 8;
    g->cap_bases = nc;
    return KGMA_OK;
}

int genome_pin(kgma_ctx *ctx, kgma_genome *g)
{
    if (g->pinned) return KGMA_OK;
    // page-lock in place so cudaMemcpyAsync is a true async DMA (north_star: pinned, double-buffered)
    // (only the 2-bit plane: the ambiguity plane reaches the device as a list of masked runs, and as a plain pageable copy
    // in the one case that needs it whole - exact match with a query shorter than 31 nt)
    cudaError_t e1 = cudaHostRegister(g->seq2, (size_t)g->cap_bases / 4, cudaHostRegisterDefault);
    if (e1 != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, KGMA_E_CUDA, "cudaHostRegister of the packed genome failed");
    }
    g->pinned = true;
    return KGMA_OK;
}

static std::atomic<uint64_t> g_uid{1};

// runs of set bits in mask words [w0, w1) as [start,end) bit positions, appended to out (runs are clipped to the range)
static void mask_runs_range(const uint32_t *mask, int64_t w0, int64_t w1, std::vector<int64_t> &out)
{
    int64_t open = -1;
    int64_t w = w0;
    while (w < w1) {
        if (open < 0) {
            // skip clear words, eight bytes at a time where alignment allows
            while (w + 1 < w1 && (w & 1) == 0 && *(const uint64_t *)(mask + w) == 0) w += 2;
            if (w >= w1) break;
        }
        uint32_t m = mask[w];
        if (m == 0) { if (open >= 0) { out.push_back(open); out.push_back(w * 32); open = -1; } w++; continue; }
        if (m == 0xFFFFFFFFu) { if (open < 0) open = w * 32; w++; continue; }
        int b = 0;
        while (b < 32) {
            if (open < 0) {
                const uint32_t rest = m >> b;
                if (!rest) break;
                b += __builtin_ctz(rest); open = w * 32 + b;
            } else {
                const uint32_t rest = ~m >> b;            // bits above 31-b read as 0 after the shift: no false "clear" bit
                if (!rest) break;
                b += __builtin_ctz(rest);
                if (b >= 32) break;
                out.push_back(open); out.push_back(w * 32 + b); open = -1;
            }
        }
        w++;
    }
    if (open >= 0) { out.push_back(open); out.push_back(w1 * 32); }
}

template <typename F> static void parallel_for(size_t n, int nthreads, F f);

// maximal runs of masked bases as [start,end) pairs in global coordinates; rebuilt when the genome changed.  A whole
// genome's ambiguity plane is hundreds of MB, so large planes are scanned in slices on several threads and stitched.
const std::vector<int64_t> &genome_nruns(kgma_genome *g)
{
    if (g->nruns_uid == g->uid) return g->nruns;
    g->nruns.clear();
    if (g->any_mask) {
        const int64_t nwords = (g->G + 31) / 32;
        const int64_t SL = (int64_t)1 << 20;               // words per slice (4 MB)
        const size_t nsl = (size_t)((nwords + SL - 1) / SL);
        std::vector<std::vector<int64_t>> part(nsl);
        const int nt = (int)std::min<size_t>(nsl, std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
        parallel_for(nsl, nt, [&](size_t i) { mask_runs_range(g->mask, (int64_t)i * SL, std::min(nwords, (int64_t)(i + 1) * SL), part[i]); });
        for (auto &pv : part)
            for (size_t j = 0; j + 1 < pv.size(); j += 2) {
                if (!g->nruns.empty() && g->nruns.back() == pv[j]) g->nruns.back() = pv[j + 1];     // a run continuing across a slice edge
                else { g->nruns.push_back(pv[j]); g->nruns.push_back(pv[j + 1]); }
            }
    }
    g->nruns_uid = g->uid;
    return g->nruns;
}

// start a new record at the next aligned offset; returns its index
static int begin_record(kgma_genome *g, const char *ident, const char *desc, int64_t len)
{
    kgma::Record r;
    r.ident = ident ? ident : "";
    r.desc = desc ? desc : r.ident;
    r.len = len;
    int64_t end = g->recs.empty() ? 0 : g->recs.back().off + g->recs.back().len;
    r.off = (end + REC_ALIGN - 1) / REC_ALIGN * REC_ALIGN;
    g->recs.push_back(std::move(r));
    return (int)g->recs.size() - 1;
}

// pack [lo,hi) of an ASCII record into the planes; returns first bad position or -1; sets flags
static void pack_ascii_range(kgma_genome *g, int64_t off, const char *s, int64_t lo, int64_t hi,
                             int64_t *bad, bool *amb, int64_t *amb_pos, bool *anymask)
{
    // lo is a multiple of 32 (except possibly 0 handled the same) so words are owned by one thread
    for (int64_t i = lo; i < hi;) {
        int64_t gp = off + i;
        uint32_t w = 0, mbits = 0;
        int n = (int)std::min<int64_t>(16, hi - i);
        for (int j = 0; j < n; j++) {
            uint8_t c = CT.t[(uint8_t)s[i + j]];
            if (c == 0xFF) { if (*bad < 0) *bad = i + j; c = 0; }
            w |= (uint32_t)(c & 3) << (2 * j);
            if (c & 4) { mbits |= 1u << j; if (c & 8) { if (!*amb) { *amb = true; *amb_pos = i + j; } } }
        }
        g->seq2[gp >> 4] = w;           // off is 128-aligned and i advances by 16: word aligned
        if (mbits) { *anymask = true; g->mask[gp >> 5] |= mbits << (gp & 31); }
        i += n;
    }
}

static int append_ascii_impl(kgma_genome *g, const char *ident, const char *desc, const char *seq, int64_t len)
{
    if (g->sealed) return KGMA_E_STATE;
    if (len < 0 || (len > 0 && !seq)) return KGMA_E_ARG;
    int r = begin_record(g, ident, desc, len);
    int64_t off = g->recs[r].off;
    int rc = genome_reserve(g, off + len + REC_ALIGN + TAIL_PAD + FGROUP);
    if (rc) return rc;
    int nt = (int)std::min<int64_t>(std::max<int64_t>(1, len >> 22), std::max(1u, std::thread::hardware_concurrency()));
    nt = std::min(nt, 32);
    std::vector<int64_t> bad(nt, -1), apos(nt, -1);
    std::vector<char> amb(nt, 0), anym(nt, 0);
    auto work = [&](int t) {
        int64_t chunk = ((len + nt - 1) / nt + 31) / 32 * 32;
        int64_t lo = std::min<int64_t>(len, chunk * t), hi = std::min<int64_t>(len, chunk * (t + 1));
        bool a = false, m = false;
        pack_ascii_range(g, off, seq, lo, hi, &bad[t], &a, &apos[t], &m);
        amb[t] = a; anym[t] = m;
    };
    if (nt == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
    for (int t = 0; t < nt; t++) {
        if (bad[t] >= 0) { g->err = "record " + std::to_string(r) + ": invalid character at position " + std::to_string(bad[t] + 1); return KGMA_E_SYMBOL; }
        if (amb[t] && !g->ambiguous) { g->ambiguous = true; g->amb_record = r; g->amb_pos = apos[t] + 1; }
        if (anym[t]) g->any_mask = true;
    }
    g->total_len += len;
    return KGMA_OK;
}

}  // namespace kgma

using namespace kgma;

extern "C" {

int kgma_version(void) { return 100; }

int kgma_genome_create(kgma_genome **out)
{
    if (!out) return KGMA_E_ARG;
    *out = new kgma_genome();
    (*out)->uid = g_uid.fetch_add(1);
    return KGMA_OK;
}

void kgma_genome_destroy(kgma_genome *g)
{
    if (!g) return;
    if (g->host_alloc) { cudaFreeHost(g->seq2); cudaFreeHost(g->mask); }
    else {
        if (g->pinned) cudaHostUnregister(g->seq2);
        if (g->seq2) munmap(g->seq2, g->seq_map);
        if (g->mask) munmap(g->mask, g->mask_map);
    }
    delete g;
}

int kgma_genome_append_ascii(kgma_genome *g, const char *identifier, const char *description,
                             const char *seq, int64_t len)
{
    if (!g) return KGMA_E_ARG;
    return append_ascii_impl(g, identifier, description, seq, len);
}

int kgma_genome_append_packed(kgma_genome *g, const char *identifier, const char *description,
                              const uint32_t *seq2, const uint32_t *mask, int64_t len)
{
    if (!g || g->sealed) return g ? KGMA_E_STATE : KGMA_E_ARG;
    if (len < 0 || (len > 0 && !seq2)) return KGMA_E_ARG;
    int r = begin_record(g, identifier, description, len);
    int64_t off = g->recs[r].off;
    int rc = genome_reserve(g, off + len + REC_ALIGN + TAIL_PAD + FGROUP);
    if (rc) return rc;
    int64_t nw = (len + 15) / 16;
    memcpy(g->seq2 + (off >> 4), seq2, (size_t)nw * 4);
    if (len & 15) g->seq2[(off >> 4) + nw - 1] &= (1u << (2 * (len & 15))) - 1;     // clear bits past the end
    if (mask) {
        int64_t mw = (len + 31) / 32;
        memcpy(g->mask + (off >> 5), mask, (size_t)mw * 4);
        if (len & 31) g->mask[(off >> 5) + mw - 1] &= (1u << (len & 31)) - 1;
        for (int64_t i = 0; i < mw && !g->any_mask; i++) if (g->mask[(off >> 5) + i]) g->any_mask = true;
    }
    g->total_len += len;
    return KGMA_OK;
}

int kgma_genome_append_bio4(kgma_genome *g, const char *identifier, const char *description,
                            const uint64_t *data, int64_t len)
{
    if (!g || g->sealed) return g ? KGMA_E_STATE : KGMA_E_ARG;
    if (len < 0 || (len > 0 && !data)) return KGMA_E_ARG;
    int r = begin_record(g, identifier, description, len);
    int64_t off = g->recs[r].off;
    int rc = genome_reserve(g, off + len + REC_ALIGN + TAIL_PAD + FGROUP);
    if (rc) return rc;
    for (int64_t i = 0; i < len; i += 16) {
        uint64_t w = data[i >> 4];
        uint32_t o = 0, mb = 0;
        int n = (int)std::min<int64_t>(16, len - i);
        for (int j = 0; j < n; j++) {
            unsigned nib = (unsigned)(w >> (4 * j)) & 15u;
            unsigned code;
            if (nib == 1) code = 0; else if (nib == 2) code = 1; else if (nib == 4) code = 2; else if (nib == 8) code = 3;
            else if (nib == 15) { code = 3; mb |= 1u << j; }
            else { code = 0; mb |= 1u << j; if (!g->ambiguous) { g->ambiguous = true; g->amb_record = r; g->amb_pos = i + j + 1; } }
            o |= code << (2 * j);
        }
        int64_t gp = off + i;
        g->seq2[gp >> 4] = o;
        if (mb) { g->any_mask = true; g->mask[gp >> 5] |= mb << (gp & 31); }
    }
    g->total_len += len;
    return KGMA_OK;
}

// Zero-copy ingest for a caller that packs the genome itself (north_star: "the Julia host code packs the genome to 2 bits
// per base plus an N/ambiguity mask"): the library lays the records out (every record starts at a multiple of 128 bases, so at
// a whole word of both planes) in page-locked planes of its own; kgma_genome_record_planes hands out where record r's words
// go; the caller writes them in place (same word format as kgma_genome_append_packed) and seals.  Scans then stream straight
// from these planes -- the path bench.py's e2e tier measures -- instead of staging a pageable copy.
int kgma_genome_create_pinned(kgma_ctx *ctx, int n_records, const int64_t *rec_len, kgma_genome **out)
{
    if (!ctx || !out || n_records < 1 || !rec_len) return KGMA_E_ARG;
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    kgma_genome *g = nullptr; kgma_genome_create(&g);
    int64_t end = 0;
    for (int r = 0; r < n_records; r++) {
        if (rec_len[r] < 0) { delete g; return set_err(ctx, KGMA_E_ARG, "record %d has a negative length", r); }
        kgma::Record R; R.ident = "record" + std::to_string(r + 1); R.desc = R.ident; R.len = rec_len[r];
        R.off = (end + REC_ALIGN - 1) / REC_ALIGN * REC_ALIGN; end = R.off + R.len;
        g->recs.push_back(R); g->total_len += R.len;
    }
    const int64_t G = (end + FGROUP - 1) / FGROUP * FGROUP + FGROUP;
    const int64_t cap = (G + TAIL_PAD + 4095) / 4096 * 4096;
    if (cudaHostAlloc((void **)&g->seq2, (size_t)cap / 4, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&g->mask, (size_t)cap / 8, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        if (g->seq2) cudaFreeHost(g->seq2);
        g->seq2 = g->mask = nullptr; delete g;
        return set_err(ctx, KGMA_E_CUDA, "cudaHostAlloc of %lld bases failed", (long long)cap);
    }
    g->host_alloc = true; g->pinned = true; g->cap_bases = cap; g->G = G;
    memset(g->seq2, 0, (size_t)cap / 4); memset(g->mask, 0, (size_t)cap / 8);
    *out = g;
    return KGMA_OK;
}

int kgma_genome_record_planes(kgma_genome *g, int record, uint32_t **seq2, uint32_t **mask)
{
    if (!g || record < 0 || record >= (int)g->recs.size() || !seq2 || !mask || !g->seq2) return KGMA_E_ARG;
    const int64_t off = g->recs[(size_t)record].off;
    *seq2 = g->seq2 + (off >> 4); *mask = g->mask + (off >> 5);
    return KGMA_OK;
}

// A genome of its own holding copies of `records` of g (in that order) in page-locked planes: what one device of a
// contig-partitioned multi-GPU scan works on.  Records start at multiples of 128 bases in both genomes, so the packed words and
// the mask words are copied as they are.
int kgma_genome_subset(kgma_ctx *ctx, const kgma_genome *g, const int32_t *records, int n_records, kgma_genome **out)
{
    if (!ctx || !g || !records || n_records < 1 || !out || !g->sealed) return KGMA_E_ARG;
    std::vector<int64_t> lens((size_t)n_records);
    for (int i = 0; i < n_records; i++) {
        if (records[i] < 0 || records[i] >= (int)g->recs.size()) return set_err(ctx, KGMA_E_ARG, "record %d out of range", records[i]);
        lens[(size_t)i] = g->recs[(size_t)records[i]].len;
    }
    kgma_genome *sub = nullptr;
    int rc = kgma_genome_create_pinned(ctx, n_records, lens.data(), &sub);
    if (rc) return rc;
    for (int i = 0; i < n_records; i++) {
        const kgma::Record &S = g->recs[(size_t)records[i]]; kgma::Record &D = sub->recs[(size_t)i];
        memcpy(sub->seq2 + (D.off >> 4), g->seq2 + (S.off >> 4), (size_t)((S.len + 15) / 16) * 4);
        memcpy(sub->mask + (D.off >> 5), g->mask + (S.off >> 5), (size_t)((S.len + 31) / 32) * 4);
        D.ident = S.ident; D.desc = S.desc;
    }
    if (g->ambiguous) {                                  // a symbol outside A,C,G,T,N somewhere in g: scans of the subset are refused as well
        sub->ambiguous = true; sub->amb_record = -1; sub->amb_pos = g->amb_pos;
        for (int i = 0; i < n_records; i++) if (records[i] == g->amb_record) sub->amb_record = i;
    }
    rc = kgma_genome_seal(sub);
    if (rc) { kgma_genome_destroy(sub); return rc; }
    *out = sub;
    return KGMA_OK;
}

int kgma_genome_set_names(kgma_genome *g, int record, const char *identifier, const char *description)
{
    if (!g || record < 0 || record >= (int)g->recs.size()) return KGMA_E_ARG;
    if (identifier) g->recs[(size_t)record].ident = identifier;
    if (description) g->recs[(size_t)record].desc = description;
    return KGMA_OK;
}

int kgma_genome_seal(kgma_genome *g)
{
    if (!g) return KGMA_E_ARG;
    if (g->sealed) return KGMA_OK;
    if (g->host_alloc) {
        // planes filled in place (kgma_genome_create_pinned): clear whatever the caller left behind the records' last
        // bases, find out whether anything is masked, and publish the new contents
        for (const kgma::Record &R : g->recs) {
            const int64_t e = R.off + R.len;
            if (e & 15) g->seq2[e >> 4] &= (1u << (2 * (e & 15))) - 1;
            if (e & 31) g->mask[e >> 5] &= (1u << (e & 31)) - 1;
        }
        const uint64_t *mw = (const uint64_t *)g->mask;
        for (int64_t i = 0, n = g->G / 64; i < n && !g->any_mask; i++) if (mw[i]) g->any_mask = true;
        g->uid = g_uid.fetch_add(1);
        g->sealed = true;
        return KGMA_OK;
    }
    int64_t end = g->recs.empty() ? 0 : g->recs.back().off + g->recs.back().len;
    int64_t G = (end + FGROUP - 1) / FGROUP * FGROUP + FGROUP;     // whole warp groups + one spare group
    int rc = genome_reserve(g, G + TAIL_PAD);
    if (rc) return rc;
    g->G = G;
    g->sealed = true;
    return KGMA_OK;
}

}  // extern "C"

// ---- parallel FASTA ingest -------------------------------------------------------------------------------------
// The file is mmap'ed; header lines are located, every record's residue bytes are cut into ~4 MB tasks, a first
// parallel pass counts residues per task (so every task knows its base offset), a second packs straight from the
// mapping into the 2-bit and ambiguity planes.  Words shared by two tasks are merged with atomic ORs.
namespace kgma {

struct FastaTask { int rec; const char *b, *e; int64_t base0 = 0, nres = 0, bad = -1, amb = -1; bool anymask = false; };

static inline bool is_ws(unsigned char c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }

static void count_task(FastaTask &t)
{
    int64_t ws = 0;
    const char *p = t.b;
#if defined(__SSE2__)
    const __m128i c1 = _mm_set1_epi8('\n'), c2 = _mm_set1_epi8('\r'), c3 = _mm_set1_epi8(' '), c4 = _mm_set1_epi8('\t');
    for (; p + 16 <= t.e; p += 16) {
        const __m128i v = _mm_loadu_si128((const __m128i *)p);
        const __m128i m = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, c1), _mm_cmpeq_epi8(v, c2)),
                                       _mm_or_si128(_mm_cmpeq_epi8(v, c3), _mm_cmpeq_epi8(v, c4)));
        ws += __builtin_popcount((unsigned)_mm_movemask_epi8(m));
    }
#endif
    for (; p < t.e; ++p) ws += is_ws((unsigned char)*p);
    t.nres = (int64_t)(t.e - t.b) - ws;
}

static void pack_task(kgma_genome *g, int64_t rec_off, FastaTask &t)
{
    int64_t gp = rec_off + t.base0, i = t.base0;            // global position / position in the record of the next residue
    const int64_t gp_end = gp + t.nres;
    uint32_t w = 0, mw = 0;
    auto flush_seq = [&](int64_t word, bool shared) {
        if (!w) return;
        if (shared) __atomic_fetch_or(&g->seq2[word], w, __ATOMIC_RELAXED); else g->seq2[word] = w;
        w = 0;
    };
    auto flush_mask = [&](int64_t word, bool shared) {
        if (!mw) return;
        if (shared) __atomic_fetch_or(&g->mask[word], mw, __ATOMIC_RELAXED); else g->mask[word] |= mw;
        mw = 0;
    };
    const int64_t first_w = gp >> 4, last_w = (gp_end - 1) >> 4, first_m = gp >> 5, last_m = (gp_end - 1) >> 5;
    for (const char *p = t.b; p < t.e; ++p) {
        // fast path: 8 plain A/C/G/T residues at a time (no whitespace, N or IUPAC among them), appended as 16 bits
        while (p + 8 <= t.e) {
            const unsigned char *u = (const unsigned char *)p;
            const uint8_t c0 = CT.t[u[0]], c1 = CT.t[u[1]], c2 = CT.t[u[2]], c3 = CT.t[u[3]],
                          c4 = CT.t[u[4]], c5 = CT.t[u[5]], c6 = CT.t[u[6]], c7 = CT.t[u[7]];
            if ((c0 | c1 | c2 | c3 | c4 | c5 | c6 | c7) & 0xFC) break;
            const uint32_t bits = (uint32_t)c0 | ((uint32_t)c1 << 2) | ((uint32_t)c2 << 4) | ((uint32_t)c3 << 6) |
                                  ((uint32_t)c4 << 8) | ((uint32_t)c5 << 10) | ((uint32_t)c6 << 12) | ((uint32_t)c7 << 14);
            const int sh = 2 * (int)(gp & 15);
            const uint64_t wide = (uint64_t)w | ((uint64_t)bits << sh);
            const int64_t gp_new = gp + 8;
            if ((gp_new >> 4) != (gp >> 4)) {                 // the word [gp>>4] is complete
                const int64_t wd = gp >> 4;
                w = (uint32_t)wide;
                flush_seq(wd, wd == first_w || wd == last_w);
                w = (uint32_t)(wide >> 32);
            } else w = (uint32_t)wide;
            if ((gp_new >> 5) != (gp >> 5)) { const int64_t wd = gp >> 5; flush_mask(wd, wd == first_m || wd == last_m); }
            gp = gp_new; i += 8; p += 8;
        }
        if (p >= t.e) break;
        const unsigned char ch = (unsigned char)*p;
        if (is_ws(ch)) continue;
        uint8_t c = CT.t[ch];
        if (c == 0xFF) { if (t.bad < 0) t.bad = i; c = 0; }
        w |= (uint32_t)(c & 3) << (2 * (gp & 15));
        if (c & 4) { mw |= 1u << (gp & 31); t.anymask = true; if ((c & 8) && t.amb < 0) t.amb = i; }
        ++gp; ++i;
        if ((gp & 15) == 0) { const int64_t wd = (gp >> 4) - 1; flush_seq(wd, wd == first_w || wd == last_w); }
        if ((gp & 31) == 0) { const int64_t wd = (gp >> 5) - 1; flush_mask(wd, wd == first_m || wd == last_m); }
    }
    if (gp & 15) flush_seq(gp >> 4, true);
    if (gp & 31) flush_mask(gp >> 5, true);
}

// ---- AVX2 forms of the two passes (chosen at run time; the SSE2/scalar forms above are the portable fallback and the
//      behavioural definition: identical planes, counts, first bad / ambiguous positions) -------------------------------
#if defined(__x86_64__)
__attribute__((target("avx2,popcnt")))
static void count_task_avx2(FastaTask &t)
{
    int64_t ws = 0;
    const char *p = t.b;
    const __m256i c1 = _mm256_set1_epi8('\n'), c2 = _mm256_set1_epi8('\r'), c3 = _mm256_set1_epi8(' '), c4 = _mm256_set1_epi8('\t');
    for (; p + 32 <= t.e; p += 32) {
        const __m256i v = _mm256_loadu_si256((const __m256i *)p);
        const __m256i m = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, c1), _mm256_cmpeq_epi8(v, c2)),
                                          _mm256_or_si256(_mm256_cmpeq_epi8(v, c3), _mm256_cmpeq_epi8(v, c4)));
        ws += __builtin_popcount((unsigned)_mm256_movemask_epi8(m));
    }
    for (; p < t.e; ++p) ws += is_ws((unsigned char)*p);
    t.nres = (int64_t)(t.e - t.b) - ws;
}

// 32 characters at a time: classify (A/C/G/T in either case, N, anything else), turn the leading run of plain or N
// residues into 2-bit codes with two multiply-adds and a byte shuffle, and append them to a bit accumulator that is
// written out in whole 32-bit words.  The first character that is neither (line ends, IUPAC codes, invalid bytes) goes
// through the scalar table, exactly as in pack_task.  With 80-column lines that is 3 vector steps per 81 bytes.
__attribute__((target("avx2,bmi2")))
static void pack_task_avx2(kgma_genome *g, int64_t rec_off, FastaTask &t)
{
    int64_t gp = rec_off + t.base0;                         // global position of the next residue
    const int64_t gp_end = gp + t.nres;
    const int64_t first_w = gp >> 4, last_w = (gp_end - 1) >> 4;
    uint64_t acc = 0; int nb = 0;                           // pending bits of word gp>>4 (nb = 2 * (gp & 15) once the task is under way)
    int64_t wd = gp >> 4;                                   // the word acc belongs to
    nb = 2 * (int)(gp & 15);                                // the task's first word may start in the middle: those low bits stay 0 here
    auto put_word = [&](uint32_t w) {
        if (wd == first_w || wd == last_w) { if (w) __atomic_fetch_or(&g->seq2[wd], w, __ATOMIC_RELAXED); }
        else g->seq2[wd] = w;
        wd++;
    };
    auto append = [&](int n, uint64_t payload, uint32_t nmask) {   // n <= 32 residues, payload = their codes, nmask = which are masked
        if (nmask) {
            t.anymask = true;
            const int sh = (int)(gp & 31);
            __atomic_fetch_or(&g->mask[gp >> 5], nmask << sh, __ATOMIC_RELAXED);
            if (sh && (nmask >> (32 - sh))) __atomic_fetch_or(&g->mask[(gp >> 5) + 1], nmask >> (32 - sh), __ATOMIC_RELAXED);
        }
        const unsigned __int128 wide = (unsigned __int128)acc | ((unsigned __int128)payload << nb);
        int tot = nb + 2 * n;
        uint64_t lo = (uint64_t)wide, hi = (uint64_t)(wide >> 64);
        while (tot >= 32) {
            put_word((uint32_t)lo);
            lo = (lo >> 32) | (hi << 32); hi >>= 32; tot -= 32;
        }
        acc = lo; nb = tot; gp += n;
    };
    const __m256i lower = _mm256_set1_epi8(0x20), three = _mm256_set1_epi8(3), one = _mm256_set1_epi8(1);
    const __m256i la = _mm256_set1_epi8('a'), lc = _mm256_set1_epi8('c'), lg = _mm256_set1_epi8('g'), lt = _mm256_set1_epi8('t'), ln = _mm256_set1_epi8('n');
    const __m256i mul4 = _mm256_set1_epi16(0x0401), mul16 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const char *p = t.b;
    while (p < t.e) {
        if (p + 32 <= t.e) {
            const __m256i v = _mm256_loadu_si256((const __m256i *)p);
            const __m256i u = _mm256_or_si256(v, lower);
            const __m256i vn = _mm256_cmpeq_epi8(u, ln);
            const __m256i va = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, la), _mm256_cmpeq_epi8(u, lc)),
                                               _mm256_or_si256(_mm256_cmpeq_epi8(u, lg), _mm256_cmpeq_epi8(u, lt)));
            const uint32_t mn = (uint32_t)_mm256_movemask_epi8(vn);
            const uint32_t ok = (uint32_t)_mm256_movemask_epi8(va) | mn;
            const int nz = ok == 0xFFFFFFFFu ? 32 : __builtin_ctz(~ok);
            if (nz) {
                // A=0x41 C=0x43 G=0x47 T=0x54 (and lower case): x = (c >> 1) & 3 gives 0,1,3,2; x ^ (x >> 1) swaps the last two.  N -> 3.
                const __m256i x = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
                __m256i code = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), one));
                code = _mm256_or_si256(code, _mm256_and_si256(vn, three));
                const __m256i b16 = _mm256_maddubs_epi16(code, mul4);              // c0 + 4 c1 per 16-bit lane
                const __m256i b32 = _mm256_madd_epi16(b16, mul16);                 // + 16 (c2 + 4 c3): one byte of packed codes per 32-bit lane
                const __m256i by = _mm256_shuffle_epi8(b32, gather);
                uint64_t payload = (uint64_t)(uint32_t)_mm256_extract_epi32(by, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(by, 4) << 32);
                uint32_t nm = mn;
                if (nz < 32) { payload &= ((uint64_t)1 << (2 * nz)) - 1; nm &= (1u << nz) - 1; }
                append(nz, payload, nm);
                p += nz;
                if (nz == 32) continue;
            }
        }
        // one character through the table: white space, IUPAC, invalid bytes, and the last few bytes of the task
        const unsigned char ch = (unsigned char)*p++;
        if (is_ws(ch)) continue;
        uint8_t c = CT.t[ch];
        const int64_t i = gp - rec_off;
        if (c == 0xFF) { if (t.bad < 0) t.bad = i; c = 0; }
        if ((c & 8) && t.amb < 0) t.amb = i;
        append(1, (uint64_t)(c & 3), (c & 4) ? 1u : 0u);
    }
    if (nb) { const uint32_t w = (uint32_t)acc; if (w) __atomic_fetch_or(&g->seq2[wd], w, __ATOMIC_RELAXED); }
}

static bool have_avx2() { static const bool v = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && !getenv("KGMA_NO_AVX2"); return v; }
#else
static bool have_avx2() { return false; }
static void count_task_avx2(FastaTask &) {}
static void pack_task_avx2(kgma_genome *, int64_t, FastaTask &) {}
#endif

template <typename F> static void parallel_for(size_t n, int nthreads, F f)
{
    if (n == 0) return;
    nthreads = (int)std::min<size_t>((size_t)std::max(1, nthreads), n);
    if (nthreads == 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto &x : th) x.join();
}

}  // namespace kgma

extern "C" {

int kgma_genome_from_fasta(const char *path, kgma_genome **out)
{
    if (!path || !out) return KGMA_E_ARG;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return KGMA_E_IO;
    struct stat st; if (fstat(fd, &st) != 0) { close(fd); return KGMA_E_IO; }
    size_t sz = (size_t)st.st_size;
    const char *buf = sz ? (const char *)mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    if (sz && buf == MAP_FAILED) { close(fd); return KGMA_E_IO; }
    if (sz) madvise((void *)buf, sz, MADV_SEQUENTIAL | MADV_WILLNEED);
    const int nthreads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    kgma_genome *g = nullptr; kgma_genome_create(&g);
    const bool trace = getenv("KGMA_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    auto lap = [&](const char *what) { if (trace) { double t1 = now(); fprintf(stderr, "[kgma ingest] %-10s %8.2f ms\n", what, t1 - t0); t0 = t1; } };

    // ---- header lines: '>' at the start of a line (found in parallel over 8 MB slices)
    const size_t SL = (size_t)8 << 20, nsl = (sz + SL - 1) / SL;
    std::vector<std::vector<size_t>> hdrs(nsl);
    parallel_for(nsl, nthreads, [&](size_t si) {
        const size_t lo = si * SL, hi = std::min(sz, lo + SL);
        const char *p = buf + lo;
        while (p < buf + hi) {
            const char *q = (const char *)memchr(p, '>', (size_t)(buf + hi - p));
            if (!q) break;
            if (q == buf || q[-1] == '\n') hdrs[si].push_back((size_t)(q - buf));
            p = q + 1;
        }
    });
    std::vector<size_t> hpos;
    for (auto &v : hdrs) hpos.insert(hpos.end(), v.begin(), v.end());
    lap("headers");

    // ---- records and tasks
    const size_t TASK = (size_t)4 << 20;
    std::vector<FastaTask> tasks;
    struct RecSpan { std::string ident, desc; size_t first_task, n_tasks; };
    std::vector<RecSpan> spans(hpos.size());
    for (size_t r = 0; r < hpos.size(); r++) {
        const size_t hs = hpos[r] + 1;
        const char *nl = (const char *)memchr(buf + hs, '\n', sz - hs);
        size_t he = nl ? (size_t)(nl - buf) : sz;
        const size_t ss = nl ? he + 1 : sz, se = (r + 1 < hpos.size()) ? hpos[r + 1] : sz;
        if (he > hs && buf[he - 1] == '\r') he--;
        spans[r].desc.assign(buf + hs, he - hs);
        size_t ie = 0; while (ie < spans[r].desc.size() && !isspace((unsigned char)spans[r].desc[ie])) ie++;
        spans[r].ident = spans[r].desc.substr(0, ie);
        spans[r].first_task = tasks.size();
        for (size_t a = ss; a < se || a == ss; a += TASK) {
            FastaTask t; t.rec = (int)r; t.b = buf + a; t.e = buf + std::min(se, a + TASK);
            tasks.push_back(t);
            if (a + TASK >= se) break;
        }
        spans[r].n_tasks = tasks.size() - spans[r].first_task;
    }
    const bool avx2 = have_avx2();
    parallel_for(tasks.size(), nthreads, [&](size_t i) { if (avx2) count_task_avx2(tasks[i]); else count_task(tasks[i]); });
    lap("count");

    // ---- layout: record lengths, offsets, one reservation
    int rc = KGMA_OK;
    for (size_t r = 0; r < spans.size(); r++) {
        int64_t len = 0;
        for (size_t i = 0; i < spans[r].n_tasks; i++) { FastaTask &t = tasks[spans[r].first_task + i]; t.base0 = len; len += t.nres; }
        begin_record(g, spans[r].ident.c_str(), spans[r].desc.c_str(), len);
        g->total_len += len;
    }
    if (!g->recs.empty()) rc = genome_reserve(g, g->recs.back().off + g->recs.back().len + REC_ALIGN + TAIL_PAD + FGROUP);
    lap("reserve");
    if (rc == KGMA_OK) {
        parallel_for(tasks.size(), nthreads, [&](size_t i) {
            if (!tasks[i].nres) return;
            if (avx2) pack_task_avx2(g, g->recs[(size_t)tasks[i].rec].off, tasks[i]); else pack_task(g, g->recs[(size_t)tasks[i].rec].off, tasks[i]);
        });
        lap("pack");
        for (const FastaTask &t : tasks) {
            if (t.bad >= 0 && rc == KGMA_OK) { g->err = "record " + std::to_string(t.rec) + ": invalid character at position " + std::to_string(t.bad + 1); rc = KGMA_E_SYMBOL; }
            if (t.amb >= 0 && !g->ambiguous) { g->ambiguous = true; g->amb_record = t.rec; g->amb_pos = t.amb + 1; }
            if (t.anymask) g->any_mask = true;
        }
    }
    if (sz) munmap((void *)buf, sz);
    close(fd);
    if (rc == KGMA_OK) rc = kgma_genome_seal(g);
    if (rc != KGMA_OK) { kgma_genome_destroy(g); return rc; }
    *out = g;
    return KGMA_OK;
}

int kgma_genome_n_records(const kgma_genome *g) { return g ? (int)g->recs.size() : 0; }
int64_t kgma_genome_record_len(const kgma_genome *g, int r) { return (g && r >= 0 && r < (int)g->recs.size()) ? g->recs[r].len : -1; }
int64_t kgma_genome_total_len(const kgma_genome *g) { return g ? g->total_len : 0; }
const char *kgma_genome_identifier(const kgma_genome *g, int r) { return (g && r >= 0 && r < (int)g->recs.size()) ? g->recs[r].ident.c_str() : nullptr; }
const char *kgma_genome_description(const kgma_genome *g, int r) { return (g && r >= 0 && r < (int)g->recs.size()) ? g->recs[r].desc.c_str() : nullptr; }

int64_t kgma_genome_record_offset(const kgma_genome *g, int r) { return (g && r >= 0 && r < (int)g->recs.size()) ? g->recs[r].off : -1; }

int kgma_genome_masked_runs(kgma_genome *g, int64_t **out, int64_t *n_runs)
{
    if (!g || !out || !n_runs) return KGMA_E_ARG;
    if (!g->sealed) return KGMA_E_STATE;
    const std::vector<int64_t> &v = genome_nruns(g);
    *n_runs = (int64_t)v.size() / 2;
    *out = (int64_t *)malloc(std::max<size_t>(1, v.size()) * sizeof(int64_t));
    if (!*out) return KGMA_E_CAPACITY;
    if (!v.empty()) memcpy(*out, v.data(), v.size() * sizeof(int64_t));
    return KGMA_OK;
}

int kgma_genome_get_seq(const kgma_genome *g, int r, int64_t first, int64_t last, char *out)
{
    if (!g || !out || r < 0 || r >= (int)g->recs.size()) return KGMA_E_ARG;
    const auto &R = g->recs[r];
    if (first < 1 || last > R.len || last < first - 1) return KGMA_E_ARG;
    static const char sym[4] = { 'A', 'C', 'G', 'T' };
    for (int64_t p = first; p <= last; p++) {
        int64_t gp = R.off + p - 1;
        char c = sym[base_code(g, gp)];
        if (base_masked(g, gp)) c = (c == 'T') ? 'N' : '?';      // '?' = an IUPAC symbol other than N was ingested
        out[p - first] = c;
    }
    out[last - first + 1] = 0;
    return KGMA_OK;
}

int kgma_genome_put_seq(kgma_genome *g, int r, int64_t first, const char *seq, int64_t len)
{
    if (!g || !seq || r < 0 || r >= (int)g->recs.size()) return KGMA_E_ARG;
    const auto &R = g->recs[r];
    if (first < 1 || first + len - 1 > R.len) return KGMA_E_ARG;
    g->uid = g_uid.fetch_add(1);                         // any device copy of the old contents is stale now
    for (int64_t i = 0; i < len; i++) {
        uint8_t c = CT.t[(uint8_t)seq[i]];
        if (c == 0xFF || (c & 8)) return KGMA_E_SYMBOL;
        int64_t gp = R.off + first - 1 + i;
        uint32_t &w = g->seq2[gp >> 4];
        w = (w & ~(3u << (2 * (gp & 15)))) | ((uint32_t)(c & 3) << (2 * (gp & 15)));
        uint32_t &m = g->mask[gp >> 5];
        if (c & 4) { m |= 1u << (gp & 31); g->any_mask = true; } else m &= ~(1u << (gp & 31));
    }
    return KGMA_OK;
}

}  // extern "C"
