// Reference-family -> profile generation in exact integer form.
// Replaces gen_ref_ws_cons (src/ReferenceGeneration.jl:4-41), cluster_ref_API (:75-138),
// get_cluster_index (:50-57), eliminate_null_params (:152-168) and Profile/consensus_seq
// (src/Consensus.jl:6-48).  The k-mer profile is kept as integer sums S and the family size N
// (RV = S/N) so that the device distance D = sum (N c_i - S_i)^2 is exact.
#include "kgma_internal.h"
#include <cmath>
#include <algorithm>

namespace kgma {

static inline int code_of(char c)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1;
    case 'G': case 'g': return 2; case 'T': case 't': return 3;
    case 'N': case 'n': return 3;                       // Consts.jl:27
    default: return -1;
    }
}

// Kmers.jl:33-44 in integers; returns false on a symbol outside A,C,G,T,N
static bool count_kmers(const std::string &s, int k, int32_t *bins)
{
    uint32_t mask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1), km = 0;
    for (size_t i = 0; i < s.size(); i++) {
        int b = code_of(s[i]); if (b < 0) return false;
        km = ((km << 2) & mask) | (uint32_t)b;
        if ((int64_t)i >= k - 1) bins[km] += 1;
    }
    return true;
}

struct ColumnVotes {                                    // Consensus.jl Profile
    std::vector<int64_t> v[4];
    void ensure(size_t n) { for (auto &x : v) if (x.size() < n) x.resize(n, 0); }
    bool add(const std::string &s)
    {
        ensure(s.size());
        for (size_t i = 0; i < s.size(); i++) { int b = code_of(s[i]); if (b < 0) return false; v[b][i]++; }
        return true;
    }
    std::string consensus() const                       // strict '>' over A,C,G,T: ties keep the earlier symbol
    {
        static const char sym[4] = { 'A', 'C', 'G', 'T' };
        std::string out(v[0].size(), 'A');
        for (size_t j = 0; j < out.size(); j++) {
            int64_t best = v[0][j];
            for (int b = 1; b < 4; b++) if (v[b][j] > best) { best = v[b][j]; out[j] = sym[b]; }
        }
        return out;
    }
};

static inline int64_t round_half_even(double x) { return (int64_t)std::nearbyint(x); }

}  // namespace kgma

using namespace kgma;

extern "C" {

int kgma_refs_create(kgma_refs **out) { if (!out) return KGMA_E_ARG; *out = new kgma_refs(); return KGMA_OK; }
void kgma_refs_destroy(kgma_refs *r) { delete r; }
int kgma_refs_count(const kgma_refs *r) { return r ? (int)r->seqs.size() : 0; }
int64_t kgma_refs_maxlen(const kgma_refs *r)
{
    int64_t m = 0; if (r) for (auto &s : r->seqs) m = std::max<int64_t>(m, (int64_t)s.size());
    return m;
}

int kgma_refs_append_ascii(kgma_refs *r, const char *seq, int64_t len)
{
    if (!r || len < 0 || (len && !seq)) return KGMA_E_ARG;
    std::string s(seq, (size_t)len);
    for (auto &c : s) { c = (char)toupper((unsigned char)c); if (code_of(c) < 0) return KGMA_E_SYMBOL; }
    r->seqs.push_back(std::move(s));
    return KGMA_OK;
}

int kgma_refs_from_fasta(const char *path, kgma_refs **out)
{
    if (!path || !out) return KGMA_E_ARG;
    kgma_genome *g = nullptr;
    int rc = kgma_genome_from_fasta(path, &g);
    if (rc) return rc;
    kgma_refs *r = new kgma_refs();
    for (int i = 0; i < kgma_genome_n_records(g) && rc == KGMA_OK; i++) {
        int64_t L = kgma_genome_record_len(g, i);
        std::string s((size_t)L, 'A');
        if (L) rc = kgma_genome_get_seq(g, i, 1, L, &s[0]);
        if (rc == KGMA_OK && s.find('?') != std::string::npos) rc = KGMA_E_SYMBOL;
        r->seqs.push_back(std::move(s));
    }
    kgma_genome_destroy(g);
    if (rc) { delete r; return rc; }
    *out = r;
    return KGMA_OK;
}

int kgma_refs_profile(const kgma_refs *r, int k, int32_t *S, int32_t *n_refs, int64_t *window, char *consensus)
{
    if (!r || !S || k < 1 || k > 12 || r->seqs.empty()) return KGMA_E_ARG;
    size_t nb = (size_t)1 << (2 * k);
    memset(S, 0, nb * sizeof(int32_t));
    int64_t cum = 0; ColumnVotes cv;
    for (auto &s : r->seqs) {
        if (!count_kmers(s, k, S)) return KGMA_E_SYMBOL;
        cum += (int64_t)s.size();
        if (!cv.add(s)) return KGMA_E_SYMBOL;
    }
    int N = (int)r->seqs.size();
    if (n_refs) *n_refs = N;
    // ReferenceGeneration.jl:33,40: len = 1/len; Int(round(cumulative_nts*len)) — Float64 product, half-even
    if (window) *window = round_half_even((double)cum * (1.0 / (double)N));
    if (consensus) { std::string c = cv.consensus(); memcpy(consensus, c.c_str(), c.size() + 1); }
    return KGMA_OK;
}

// gen_ref_ws_cons of the strobemer path (src/StrobemerGMA/StrobeRefGen.jl:4-42): S[code] = summed counts of the gap-free
// 2-randstrobe codes (ungapped_strobe_2_mer_count!, Strobemers.jl:105-115) over all references; window and consensus as above.
// get_strobe_2_mer (Strobemers.jl:45-65): first strobe = bases 1..s, second strobe at the LAST start in w_min..w_max whose
// randstrobe_score (as_UInt(first) + as_UInt(candidate)) % q is 0 -- the running minimum starts at `2 << 63` == 0 -- else w_min.
int kgma_refs_strobe_profile(const kgma_refs *r, int s, int w_min, int w_max, int q, int32_t *S, int32_t *n_refs, int64_t *window, char *consensus)
{
    if (!r || !S || s < 1 || s > 6 || w_min < 1 || w_max < w_min || q < 1 || r->seqs.empty()) return KGMA_E_ARG;
    const int k = w_max + s - 1;
    size_t nb = (size_t)1 << (4 * s);
    memset(S, 0, nb * sizeof(int32_t));
    int64_t cum = 0; ColumnVotes cv;
    auto code = [](char c) -> int { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': case 'N': return 3; default: return -1; } };
    for (auto &sq : r->seqs) {
        const int64_t L = (int64_t)sq.size();
        for (int64_t i = 0; i < L; i++) if (code(sq[(size_t)i]) < 0) return KGMA_E_SYMBOL;
        for (int64_t i = 0; i + k <= L; i++) {
            auto smer = [&](int i1) { uint32_t v = 0; for (int j = 0; j < s; j++) v = (v << 2) | (uint32_t)code(sq[(size_t)(i + i1 - 1 + j)]); return v; };
            const uint32_t f = smer(1);
            int min_ind = w_min;
            for (int w = w_min; w <= w_max; w++) if ((f + smer(w)) % (uint32_t)q == 0) min_ind = w;
            S[((size_t)f << (2 * s)) | smer(min_ind)] += 1;
        }
        cum += L;
        if (!cv.add(sq)) return KGMA_E_SYMBOL;
    }
    int N = (int)r->seqs.size();
    if (n_refs) *n_refs = N;
    if (window) *window = round_half_even((double)cum * (1.0 / (double)N));
    if (consensus) { std::string c = cv.consensus(); memcpy(consensus, c.c_str(), c.size() + 1); }
    return KGMA_OK;
}

int kgma_refs_cluster(const kgma_refs *r, int k, const double *cutoffs, int n_cutoffs, int include_avg,
                      int drop_empty, int32_t *S, int32_t *n_members, int64_t *windows,
                      char *consensus, int64_t cons_stride, int32_t *invalid, double *ref_dists)
{
    if (!r || !S || !windows || !consensus || k < 1 || k > 12 || r->seqs.empty() || n_cutoffs < 0) return KGMA_E_ARG;
    size_t nb = (size_t)1 << (2 * k);
    int N = (int)r->seqs.size();
    int64_t maxlen = kgma_refs_maxlen(r);
    if (cons_stride < maxlen + 1) return KGMA_E_CAPACITY;
    std::vector<int32_t> Savg(nb); int64_t ws_avg = 0; std::string cons_avg((size_t)maxlen + 1, 0);
    int rc = kgma_refs_profile(r, k, Savg.data(), nullptr, &ws_avg, &cons_avg[0]);
    if (rc) return rc;
    // average KFV exactly as the reference forms it: Float64(count sum) * (1/N)   (ReferenceGeneration.jl:35,40)
    std::vector<double> avg(nb); double inv = 1.0 / (double)N;
    for (size_t i = 0; i < nb; i++) avg[i] = (double)Savg[i] * inv;

    int nc = n_cutoffs + 1;
    std::vector<std::vector<int32_t>> Sc(nc, std::vector<int32_t>(nb, 0));
    std::vector<ColumnVotes> cv(nc);
    for (auto &c : cv) c.ensure((size_t)maxlen);
    std::vector<int64_t> wsum(nc, 0); std::vector<int> lens(nc, 0);
    std::vector<int32_t> cnt(nb);
    for (int j = 0; j < N; j++) {
        const std::string &s = r->seqs[j];
        std::fill(cnt.begin(), cnt.end(), 0);
        if (!count_kmers(s, k, cnt.data())) return KGMA_E_SYMBOL;
        // kmer_dist(seq, average_KFV, k) = (1/(2k)) * sqeuclidean(kmer_count(seq), KFV)   (Kmers.jl:58-60)
        double acc = 0.0;
        for (size_t i = 0; i < nb; i++) { double d = (double)cnt[i] - avg[i]; acc += d * d; }
        double dist = (1.0 / (double)(2 * k)) * acc;
        if (ref_dists) ref_dists[j] = dist;
        int ci = 0; while (ci < n_cutoffs && !(dist <= cutoffs[ci])) ci++;      // get_cluster_index :50-57
        cv[ci].add(s); wsum[ci] += (int64_t)s.size(); lens[ci] += 1;
        for (size_t i = 0; i < nb; i++) Sc[ci][i] += cnt[i];
    }
    int out = 0;
    for (int c = 0; c < nc; c++) {
        bool inval = lens[c] == 0;
        if (invalid) invalid[c] = inval;
        if (inval && drop_empty) continue;
        memcpy(S + (size_t)out * nb, Sc[c].data(), nb * sizeof(int32_t));
        if (n_members) n_members[out] = lens[c];
        char *cd = consensus + (size_t)out * (size_t)cons_stride;
        if (!inval) {
            windows[out] = round_half_even((double)wsum[c] / (double)lens[c]);             // :119
            std::string cs = cv[c].consensus();
            size_t w = (size_t)std::min<int64_t>(windows[out], (int64_t)cs.size());
            memcpy(cd, cs.data(), w); cd[w] = 0;                                              // :120 [1:ws_i]
        } else { windows[out] = 0; cd[0] = 0; }
        out++;
    }
    if (include_avg) {                                                                        // :127-132
        if (invalid) invalid[nc] = 0;
        memcpy(S + (size_t)out * nb, Savg.data(), nb * sizeof(int32_t));
        if (n_members) n_members[out] = N;
        windows[out] = ws_avg;
        strcpy(consensus + (size_t)out * (size_t)cons_stride, cons_avg.c_str());
        out++;
    }
    return out;
}

int kgma_profile_from_kfv(const double *kfv, int64_t n_bins, int32_t max_n, int32_t *S, int32_t *n_refs)
{
    if (!kfv || !S || !n_refs || n_bins <= 0) return KGMA_E_ARG;
    if (max_n <= 0) max_n = 100000;
    for (int32_t N = 1; N <= max_n; N++) {
        bool ok = true;
        for (int64_t i = 0; i < n_bins && ok; i++) {
            double x = kfv[i] * (double)N, rx = std::nearbyint(x);
            if (std::fabs(x - rx) > 1e-9 * std::max(1.0, std::fabs(x)) || rx < 0 || rx > 2.0e9) ok = false;
        }
        if (!ok) continue;
        for (int64_t i = 0; i < n_bins; i++) S[i] = (int32_t)std::nearbyint(kfv[i] * (double)N);
        *n_refs = N;
        return KGMA_OK;
    }
    return KGMA_E_UNSUPPORTED;
}

}  // extern "C"
