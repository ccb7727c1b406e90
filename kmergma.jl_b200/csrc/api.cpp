// Context life-cycle and result accessors of the C ABI (include/kmergma.h).
#include "kgma_internal.h"
#include <sys/mman.h>
#include <cmath>
#include <cstdlib>
#include <algorithm>

namespace kgma { const char *create_err(); }
using namespace kgma;

#include <mutex>

static std::mutex g_pool_mu;
static std::vector<kgma_result *> g_pool;

kgma_result *result_acquire()
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool.empty()) { kgma_result *r = g_pool.back(); g_pool.pop_back(); return r; }
    }
    return new kgma_result();
}

void result_release(kgma_result *r)
{
    if (!r) return;
    // keep the capacity of the flat vectors, drop the per-window distance vectors (they can be GBs)
    r->hits.clear(); r->runs.clear(); r->first_D.clear(); r->cigar_ops.clear(); r->cigar_cnt.clear(); r->align_events.clear();
    std::vector<std::vector<double>>().swap(r->dists);
    const size_t keep = r->hits.capacity() * sizeof(kgma_hit) + r->runs.capacity() * sizeof(kgma_run) + r->cigar_ops.capacity() + r->cigar_cnt.capacity() * 4;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < 4 && keep < ((size_t)64 << 20)) { g_pool.push_back(r); return; }
    }
    delete r;
}

extern "C" {

int kgma_create(int device, kgma_ctx **out)
{
    if (!out) return KGMA_E_ARG;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_err(nullptr, KGMA_E_CUDA, "no CUDA device available (%s); libkmergma_cuda has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return set_err(nullptr, KGMA_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return set_err(nullptr, KGMA_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return set_err(nullptr, KGMA_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10) return set_err(nullptr, KGMA_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    kgma_ctx *c = new kgma_ctx();
    c->device = device; c->num_sms = prop.multiProcessorCount; c->smem_optin = prop.sharedMemPerBlockOptin;
    bool ok = cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_align, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 8 && ok; i++) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++) ok = cudaEventCreate(&c->a_ev0[i]) == cudaSuccess && cudaEventCreate(&c->a_ev1[i]) == cudaSuccess &&
                                           cudaEventCreate(&c->a_done[i]) == cudaSuccess;
    if (!ok) { kgma_destroy(c); return set_err(nullptr, KGMA_E_CUDA, "stream/event creation failed"); }
    *out = c;
    return KGMA_OK;
}

void kgma_destroy(kgma_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->s_compute) cudaStreamSynchronize(c->s_compute);
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
    if (c->s_align) cudaStreamSynchronize(c->s_align);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->a_ev0[i]) cudaEventDestroy(c->a_ev0[i]);
        if (c->a_ev1[i]) cudaEventDestroy(c->a_ev1[i]);
        if (c->a_done[i]) cudaEventDestroy(c->a_done[i]);
        if (c->a_dev[i]) cudaFree(c->a_dev[i]);
        if (c->a_host[i]) cudaFreeHost(c->a_host[i]);
    }
    for (auto &e : c->chunk_ev) cudaEventDestroy(e);
    for (auto &e : c->stage_ev) if (e) cudaEventDestroy(e);
    if (c->stage) { cudaHostUnregister(c->stage); munmap(c->stage, c->stage_bytes); }
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_align) cudaStreamDestroy(c->s_align);
    if (c->d_seq2) cudaFree(c->d_seq2);
    if (c->d_mask) cudaFree(c->d_mask);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    delete c;
}

const char *kgma_last_error(const kgma_ctx *c) { return c ? c->err.c_str() : create_err(); }

int kgma_get_stats(const kgma_ctx *c, kgma_stats *out) { if (!c || !out) return KGMA_E_ARG; *out = c->stats; return KGMA_OK; }

int64_t kgma_result_n_hits(const kgma_result *r) { return r ? (int64_t)r->hits.size() : 0; }
const kgma_hit *kgma_result_hits(const kgma_result *r) { return r && !r->hits.empty() ? r->hits.data() : nullptr; }
int64_t kgma_result_n_runs(const kgma_result *r) { return r ? (int64_t)r->runs.size() : 0; }
const kgma_run *kgma_result_runs(const kgma_result *r) { return r && !r->runs.empty() ? r->runs.data() : nullptr; }
const int64_t *kgma_result_first_D(const kgma_result *r) { return r && !r->first_D.empty() ? r->first_D.data() : nullptr; }
int64_t kgma_result_n_dists(const kgma_result *r, int p) { return (r && p >= 0 && p < (int)r->dists.size()) ? (int64_t)r->dists[p].size() : 0; }
const double *kgma_result_dists(const kgma_result *r, int p) { return (r && p >= 0 && p < (int)r->dists.size() && !r->dists[p].empty()) ? r->dists[p].data() : nullptr; }
const char *kgma_result_cigar_ops(const kgma_result *r) { return r && !r->cigar_ops.empty() ? r->cigar_ops.data() : nullptr; }
const int32_t *kgma_result_cigar_counts(const kgma_result *r) { return r && !r->cigar_cnt.empty() ? r->cigar_cnt.data() : nullptr; }
int64_t kgma_result_n_align_events(const kgma_result *r) { return r ? (int64_t)r->align_events.size() : 0; }
const kgma_align_event *kgma_result_align_events(const kgma_result *r) { return r && !r->align_events.empty() ? r->align_events.data() : nullptr; }
void kgma_result_free(kgma_result *r) { result_release(r); }

// ---- result formatting and writing: append_hit!'s header (Alignment.jl:57-81; OmnGenomeMiner.jl:141-149 in cluster mode) and
// write_results (API.jl:234-241) natively, for callers without Julia's string(round(d, digits = 2)).
// string(round(d, digits = 2)): round-half-even of d*100 then /100 (Base.round), printed as the shortest decimal that reads
// back to the same Float64, always with a fractional part ("8.1", "24.87", "9.0").
static std::string julia_round2_string(double d)
{
    const double y = std::nearbyint(d * 100.0) / 100.0;      // (default rounding mode: ties to even, like Base.round)
    if (std::isnan(y)) return "NaN";
    if (std::isinf(y)) return y > 0 ? "Inf" : "-Inf";
    if (y == 0) return std::signbit(y) ? "-0.0" : "0.0";
    char buf[64];
    for (int p = 0; p <= 16; p++) {                           // shortest digit string that reads back to y
        snprintf(buf, sizeof buf, "%.*e", p, y);
        if (strtod(buf, nullptr) == y) break;
    }
    std::string t(buf), digits;
    const size_t e = t.find('e');
    const int ex = atoi(t.c_str() + e + 1);
    const bool neg = t[0] == '-';
    for (size_t i = neg ? 1 : 0; i < e; i++) if (t[i] != '.') digits += t[i];
    std::string s = neg ? "-" : "";
    if (ex >= -5 && ex < 21) {                                // Base.show(::Float64): fixed notation in [1e-5, 1e21)
        if (ex < 0) { s += "0." + std::string((size_t)(-ex - 1), '0') + digits; }
        else if ((size_t)ex + 1 >= digits.size()) { s += digits + std::string((size_t)ex + 1 - digits.size(), '0') + ".0"; }
        else { s += digits.substr(0, (size_t)ex + 1) + "." + digits.substr((size_t)ex + 1); }
    } else {                                                  // 1.0e21, 1.5e-7
        s += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "e" + std::to_string(ex);
    }
    return s;
}

static std::string hit_header(const kgma_genome *g, const kgma_hit &h, bool cluster, bool with_genome_pos)
{
    const std::string id = (h.record >= 0 && h.record < (int)g->recs.size()) ? g->recs[(size_t)h.record].ident : std::string("?");
    std::string s = id + (cluster ? " | Dist = " : " | dist = ") + julia_round2_string(h.dist);
    if (cluster) s += " | KFV = " + std::to_string(h.profile);
    s += " | MatchPos = " + std::to_string(h.first) + ":" + std::to_string(h.last);
    if (cluster || with_genome_pos) s += " | GenomePos = " + std::to_string(h.genome_pos);     // record_KmerGMA! prints none (MultiThread/GenomeMiner.jl:88-91)
    s += " | Len = " + std::to_string(h.last - h.first + 1);
    return s;
}

int64_t kgma_hit_header(const kgma_genome *g, const kgma_hit *h, int cluster, int with_genome_pos, char *buf, int64_t cap)
{
    if (!g || !h) return KGMA_E_ARG;
    const std::string s = hit_header(g, *h, cluster != 0, with_genome_pos != 0);
    if (buf && cap > (int64_t)s.size()) memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

int kgma_result_write_fasta(const kgma_result *r, const kgma_genome *g, const char *path, int cluster, int with_genome_pos,
                            int width, int64_t *n_written)
{
    if (!r || !g || !path || width < 1) return KGMA_E_ARG;
    FILE *fp = fopen(path, "ab");                        // API.jl:235 open(file_path, "a")
    if (!fp) return KGMA_E_IO;
    std::string seq;
    int64_t n = 0;
    for (const kgma_hit &h : r->hits) {
        const std::string hd = hit_header(g, h, cluster != 0, with_genome_pos != 0);
        fputc('>', fp); fwrite(hd.data(), 1, hd.size(), fp); fputc('\n', fp);
        const int64_t len = h.last >= h.first ? h.last - h.first + 1 : 0;
        seq.assign((size_t)len + 1, '\0');
        if (len > 0 && kgma_genome_get_seq(g, h.record, h.first, h.last, &seq[0]) != KGMA_OK) { fclose(fp); return KGMA_E_ARG; }
        for (int64_t o = 0; o < len; o += width) {
            fwrite(seq.data() + o, 1, (size_t)std::min<int64_t>(width, len - o), fp); fputc('\n', fp);
        }
        n++;
    }
    if (fclose(fp) != 0) return KGMA_E_IO;
    if (n_written) *n_written = n;
    return KGMA_OK;
}

// fasta_id_to_cumulative_len_dict (ExactMatch.jl:146-158): the summed length of the records in front of `record`
int64_t kgma_genome_cumulative_len(const kgma_genome *g, int record)
{
    if (!g || record < 0 || record >= (int)g->recs.size()) return KGMA_E_ARG;
    int64_t c = 0;
    for (int r = 0; r < record; r++) c += g->recs[(size_t)r].len;
    return c;
}
// Contig-partitioned multi-GPU scan, the merge on rank 0: block b = [int64 n][8 bytes][kgma_hit x n] as written by rank b (records
// numbered within that rank's sub-genome); rec_map[rec_off[b] + local] = global record index; genome_pos_of[global] = GenomePos of
// that record in the whole genome.  Records are independent (GenomePos is a running sum of record lengths, GenomeMiner.jl:106),
// so the result is the blocks' hits renumbered and ordered by global record (a counting sort; hits of one record keep their order).
int kgma_hits_merge_partition(const void *blocks, int n_blocks, int64_t stride, const int32_t *rec_map, const int32_t *rec_off,
                              int32_t n_records, const int64_t *genome_pos_of, kgma_hit **out, int64_t *n_out)
{
    if (!blocks || n_blocks < 1 || stride < 16 || !rec_map || !rec_off || n_records < 1 || !genome_pos_of || !out || !n_out) return KGMA_E_ARG;
    int64_t total = 0;
    for (int b = 0; b < n_blocks; b++) {
        int64_t n; memcpy(&n, (const char *)blocks + (size_t)b * (size_t)stride, 8);
        if (n < 0 || 16 + n * (int64_t)sizeof(kgma_hit) > stride) return KGMA_E_ARG;
        total += n;
    }
    kgma_hit *res = (kgma_hit *)malloc((size_t)std::max<int64_t>(total, 1) * sizeof(kgma_hit));
    if (!res) return KGMA_E_CAPACITY;
    std::vector<int64_t> start((size_t)n_records + 1, 0);
    for (int pass = 0; pass < 2; pass++) {
        for (int b = 0; b < n_blocks; b++) {
            const char *p = (const char *)blocks + (size_t)b * (size_t)stride;
            int64_t n; memcpy(&n, p, 8);
            const kgma_hit *h = (const kgma_hit *)(p + 16);
            const int32_t lo = rec_off[b], hi = rec_off[b + 1];
            for (int64_t i = 0; i < n; i++) {
                if (h[i].record < 0 || h[i].record >= hi - lo) { free(res); return KGMA_E_ARG; }
                const int32_t gr = rec_map[lo + h[i].record];
                if (gr < 0 || gr >= n_records) { free(res); return KGMA_E_ARG; }
                if (pass == 0) start[(size_t)gr + 1]++;
                else { kgma_hit &o = res[start[(size_t)gr]++]; o = h[i]; o.record = gr; o.genome_pos = genome_pos_of[gr]; }
            }
        }
        if (pass == 0) for (int32_t r = 0; r < n_records; r++) start[(size_t)r + 1] += start[(size_t)r];
    }
    *out = res; *n_out = total;
    return KGMA_OK;
}

void kgma_free(void *p) { free(p); }

}  // extern "C"
