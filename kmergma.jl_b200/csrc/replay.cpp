// Host replay of the reference's sequential hit state machine from run summaries
// (SURVEY Appendix B).  Replaces the minima finder / hit processing of
//   ac_gma_testing!  src/GenomeMiner.jl:57,82-104   and   Omn_KmerGMA!  src/OmnGenomeMiner.jl:59,114-156
// which carry (currminim, CMI, stop, goal_ind / prev_hit_range) base to base.  A run is a maximal
// stretch of loop steps with d < thr; (t_last, D_min, first argmin) per run is a sufficient
// statistic, so the replay is O(#runs) and independent of how the device segmented the genome.
#include "kgma_internal.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <climits>

namespace kgma {

static void merge_runs_slow(std::vector<kgma_run> &runs, std::vector<kgma_run_ext> *ext);

// `ext` (optional, parallel to `runs`): the extension result a shard computed for each run's own first-argmin window
// (kgma_scan_shard).  A merged run keeps the result of the piece that supplies its argmin.
void merge_runs(std::vector<kgma_run> &runs, std::vector<kgma_run_ext> *ext)
{
    static thread_local std::vector<uint32_t> order;
    bool sorted_by_index = false;
    auto tnow_ = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool trace_ = getenv("KGMA_TRACE_MERGE") != nullptr; const double tm0 = trace_ ? tnow_() : 0; double tm1 = 0, tm2 = 0;
    // order by (profile, record, t_first, marker) - the device appends runs in arbitrary order.  A few thousand runs per
    // genome: the key is squeezed to the bits that can differ ((profile, record) as one group number, t_first relative to
    // the smallest one) and sorted with an LSD radix sort of (key, index) pairs, 11 bits per pass -- three or four passes
    // for a 3 Gb genome -- then the 48-byte runs are read once, in order (a comparison sort of the runs themselves spends
    // ~0.1 ms in mispredicted branches, 5 % of a resident scan).
    {
        auto less = [](const kgma_run &a, const kgma_run &b) {
            if (a.profile != b.profile) return a.profile < b.profile;
            if (a.record != b.record) return a.record < b.record;
            if (a.t_first != b.t_first) return a.t_first < b.t_first;
            return (a.flags & KGMA_RUN_MARKER) < (b.flags & KGMA_RUN_MARKER);
        };
        const size_t n = runs.size();
        bool ok = n > 1 && n < ((size_t)1 << 31);
        int32_t maxrec = 0, maxprof = 0; int64_t tmin = INT64_MAX, tmax = INT64_MIN;
        for (const kgma_run &r : runs) {
            if (r.profile < 0 || r.record < 0 || r.t_first < 0) { ok = false; break; }
            maxrec = std::max(maxrec, r.record); maxprof = std::max(maxprof, r.profile);
            tmin = std::min(tmin, r.t_first); tmax = std::max(tmax, r.t_first);
        }
        auto bits_of = [](uint64_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; };
        int tbits = 0, gbits = 0;
        if (ok) {
            tbits = bits_of((uint64_t)(tmax - tmin));
            gbits = bits_of((uint64_t)maxprof * ((uint64_t)maxrec + 1) + (uint64_t)maxrec);
            if (tbits + gbits + 1 > 62 || (uint64_t)maxprof * ((uint64_t)maxrec + 1) > ((uint64_t)1 << 40)) ok = false;
        }
        if (!ok && ext) { merge_runs_slow(runs, ext); return; }
        if (!ok) std::sort(runs.begin(), runs.end(), less);
        else {
            struct KI { uint64_t key; uint32_t idx; };
            static thread_local std::vector<KI> ka, kb;               // (scratch kept per thread: no allocation per call)
            ka.resize(n); kb.resize(n);
            const uint64_t nrec1 = (uint64_t)maxrec + 1;
            for (size_t i = 0; i < n; i++) {
                const kgma_run &r = runs[i];
                const uint64_t grp = (uint64_t)r.profile * nrec1 + (uint64_t)r.record;
                ka[i] = { (grp << (tbits + 1)) | ((uint64_t)(r.t_first - tmin) << 1) | ((r.flags & KGMA_RUN_MARKER) ? 1u : 0u), (uint32_t)i };
            }
            const int total = gbits + tbits + 1;
            if (trace_) tm1 = tnow_();
            KI *src = ka.data(), *dst = kb.data();
            // digits of 8 bits for short lists (the prefix over the buckets is a fixed cost per pass: 256 steps instead of 2048),
            // of 11 bits otherwise; either way ceil(total / digit) passes
            const int dig = (n < 16384 && (total + 7) / 8 <= (total + 10) / 11) ? 8 : 11;
            const uint32_t dmask = (1u << dig) - 1u;
            for (int sh = 0; sh < total; sh += dig) {
                uint32_t cnt[2049];
                memset(cnt, 0, ((size_t)dmask + 2) * sizeof(uint32_t));
                for (size_t i = 0; i < n; i++) cnt[((src[i].key >> sh) & dmask) + 1]++;
                if (cnt[((src[0].key >> sh) & dmask) + 1] == n) continue;       // this digit is the same everywhere
                for (uint32_t d = 0; d <= dmask; d++) cnt[d + 1] += cnt[d];
                for (size_t i = 0; i < n; i++) dst[cnt[(src[i].key >> sh) & dmask]++] = src[i];
                std::swap(src, dst);
            }
            order.resize(n);
            for (size_t i = 0; i < n; i++) order[i] = src[i].idx;         // (LSD passes are stable: equal keys keep device order)
            sorted_by_index = true;
        }
    }
    if (trace_) tm2 = tnow_();
    static thread_local std::vector<kgma_run> out;
    static thread_local std::vector<kgma_run_ext> out_ext;
    out.clear(); out_ext.clear();
    out.reserve(runs.size());
    if (!ext && !runs.empty()) {
        // the common case (no per-run extension results to carry along): the open run is kept in registers and the updates are
        // written without data-dependent branches -- whether two neighbours join is a coin flip the branch predictor loses
        out.resize(runs.size());
        kgma_run *o = out.data(); size_t no = 0;
        kgma_run cur = runs[sorted_by_index ? order[0] : 0];
        for (size_t ri = 1; ri < runs.size(); ri++) {
            const kgma_run &r = runs[sorted_by_index ? order[ri] : ri];
            const bool same = cur.profile == r.profile && cur.record == r.record;
            const bool pm = (cur.flags & KGMA_RUN_MARKER) != 0, rm = (r.flags & KGMA_RUN_MARKER) != 0;
            if (same && pm && rm && cur.t_first == r.t_first) continue;              // the same marker reported twice
            if (same && !pm && !rm && r.t_first <= cur.t_last + 1) {
                // pieces of one maximal run (see below): min of minima, earlier argmin on ties
                const bool lt = r.D_min < cur.D_min, eq = r.D_min == cur.D_min;
                const uint32_t tie_new = lt ? (r.flags & KGMA_HIT_ARGMIN_TIE)
                                            : ((cur.flags & KGMA_HIT_ARGMIN_TIE) | (eq ? ((r.flags & KGMA_HIT_ARGMIN_TIE) | (r.t_argmin != cur.t_argmin ? KGMA_HIT_ARGMIN_TIE : 0u)) : 0u));
                const int64_t arg_new = lt ? r.t_argmin : (eq ? std::min(cur.t_argmin, r.t_argmin) : cur.t_argmin);
                cur.D_min = lt ? r.D_min : cur.D_min;
                cur.t_argmin = arg_new;
                uint32_t fl = (cur.flags & ~KGMA_HIT_ARGMIN_TIE) | tie_new | (r.flags & KGMA_HIT_NEAR_THR);
                const bool ext_right = r.t_last >= cur.t_last;
                fl = ext_right ? ((fl & ~KGMA_RUN_OPEN_RIGHT) | (r.flags & KGMA_RUN_OPEN_RIGHT)) : fl;
                cur.t_last = ext_right ? r.t_last : cur.t_last;
                cur.flags = fl;
                continue;
            }
            o[no++] = cur; cur = r;
        }
        o[no++] = cur;
        out.resize(no);
        runs.swap(out);
        if (trace_) fprintf(stderr, "[kgma merge] keys %.3f ms, sort %.3f ms, merge %.3f ms\n", tm1 - tm0, tm2 - tm1, tnow_() - tm2);
        return;
    }
    for (size_t ri = 0; ri < runs.size(); ri++) {
        const size_t src = sorted_by_index ? order[ri] : ri;
        if (sorted_by_index && ri + 12 < runs.size()) __builtin_prefetch(&runs[order[ri + 12]]);   // (the reads hop around a 200 KB array)
        const kgma_run &r = runs[src];                                      // (merged straight out of the unsorted array: no gather copy)
        if (!out.empty() && out.back().profile == r.profile && out.back().record == r.record) {
            kgma_run &p = out.back();
            const bool pm = (p.flags & KGMA_RUN_MARKER) != 0, rm = (r.flags & KGMA_RUN_MARKER) != 0;
            if (pm && rm && p.t_first == r.t_first) continue;            // the same marker reported twice
            if (!pm && !rm && r.t_first <= p.t_last + 1) {
                // pieces of one maximal run: cut by a span / chunk / shard boundary (adjacent), or reported twice because a
                // span was evaluated again (overlapping, same D values).  Min of minima, earlier argmin on ties.
                if (r.D_min < p.D_min) {
                    p.D_min = r.D_min; p.t_argmin = r.t_argmin; p.flags = (p.flags & ~KGMA_HIT_ARGMIN_TIE) | (r.flags & KGMA_HIT_ARGMIN_TIE);
                    if (ext) out_ext.back() = (*ext)[src];
                }
                else if (r.D_min == p.D_min) {
                    if (r.t_argmin != p.t_argmin) p.flags |= KGMA_HIT_ARGMIN_TIE;
                    if (ext && (r.t_argmin < p.t_argmin || (r.t_argmin == p.t_argmin && out_ext.back().lo == 0))) out_ext.back() = (*ext)[src];
                    p.t_argmin = std::min(p.t_argmin, r.t_argmin);
                    p.flags |= (r.flags & KGMA_HIT_ARGMIN_TIE);
                }
                p.flags |= (r.flags & KGMA_HIT_NEAR_THR);
                if (r.t_last >= p.t_last) { p.flags = (p.flags & ~KGMA_RUN_OPEN_RIGHT) | (r.flags & KGMA_RUN_OPEN_RIGHT); p.t_last = r.t_last; }
                continue;
            }
        }
        out.push_back(r);
        if (ext) out_ext.push_back((*ext)[src]);
    }
    runs.swap(out);                                      // (the scratch vector keeps the old buffer for the next call)
    if (ext) ext->swap(out_ext);
    if (trace_) fprintf(stderr, "[kgma merge] keys %.3f ms, sort %.3f ms, merge %.3f ms\n", tm1 - tm0, tm2 - tm1, tnow_() - tm2);
}

// merge with extension results for run lists the radix path does not take (out-of-range keys, fewer than two runs):
// sort an index array with the same order, permute both arrays, then run the linear merge on the sorted input
static void merge_runs_slow(std::vector<kgma_run> &runs, std::vector<kgma_run_ext> *ext)
{
    std::vector<size_t> idx(runs.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) {
        const kgma_run &a = runs[x], &b = runs[y];
        if (a.profile != b.profile) return a.profile < b.profile;
        if (a.record != b.record) return a.record < b.record;
        if (a.t_first != b.t_first) return a.t_first < b.t_first;
        return (a.flags & KGMA_RUN_MARKER) < (b.flags & KGMA_RUN_MARKER);
    });
    std::vector<kgma_run> out; std::vector<kgma_run_ext> out_ext;
    for (size_t src : idx) {
        const kgma_run &r = runs[src];
        if (!out.empty() && out.back().profile == r.profile && out.back().record == r.record) {
            kgma_run &p = out.back();
            const bool pm = (p.flags & KGMA_RUN_MARKER) != 0, rm = (r.flags & KGMA_RUN_MARKER) != 0;
            if (pm && rm && p.t_first == r.t_first) continue;
            if (!pm && !rm && r.t_first <= p.t_last + 1) {
                if (r.D_min < p.D_min) { p.D_min = r.D_min; p.t_argmin = r.t_argmin; p.flags = (p.flags & ~KGMA_HIT_ARGMIN_TIE) | (r.flags & KGMA_HIT_ARGMIN_TIE); out_ext.back() = (*ext)[src]; }
                else if (r.D_min == p.D_min) {
                    if (r.t_argmin != p.t_argmin) p.flags |= KGMA_HIT_ARGMIN_TIE;
                    if (r.t_argmin < p.t_argmin || (r.t_argmin == p.t_argmin && out_ext.back().lo == 0)) out_ext.back() = (*ext)[src];
                    p.t_argmin = std::min(p.t_argmin, r.t_argmin);
                    p.flags |= (r.flags & KGMA_HIT_ARGMIN_TIE);
                }
                p.flags |= (r.flags & KGMA_HIT_NEAR_THR);
                if (r.t_last >= p.t_last) { p.flags = (p.flags & ~KGMA_RUN_OPEN_RIGHT) | (r.flags & KGMA_RUN_OPEN_RIGHT); p.t_last = r.t_last; }
                continue;
            }
        }
        out.push_back(r); out_ext.push_back((*ext)[src]);
    }
    runs.swap(out); ext->swap(out_ext);
}

static uint32_t round_half_flag(double d)
{
    double x = d * 100.0, f = x - std::floor(x);
    return std::fabs(f - 0.5) < 1e-7 ? KGMA_HIT_ROUND_HALF : 0u;
}

// Alignment.jl:49-51 / OmnGenomeMiner.jl:135-136: remap the aligned sub-range into record coordinates
static void remap(int64_t first, int64_t L, const AlignRes &a, int64_t *nf, int64_t *nl)
{
    int64_t f = std::max<int64_t>(1, first + a.lo - 1);
    int64_t l = std::min<int64_t>(first + a.hi - 1, L);
    if (l < f - 1) l = f - 1;                            // Julia UnitRange normalisation
    *nf = f; *nl = l;
}

void apply_extensions(kgma_genome *g, std::vector<kgma_hit> &hits, const std::vector<Pending> &pend, const std::vector<AlignRes> &ares)
{
    for (const Pending &p : pend) {
        kgma_hit &h = hits[p.hit];
        remap(p.first, g->recs[h.record].len, ares[p.req], &h.first, &h.last);
        h.align_score = ares[p.req].score; h.cigar_off = ares[p.req].cig_off; h.cigar_len = ares[p.req].cig_len;
    }
}

// ac_gma_testing! state machine (GenomeMiner.jl:57,82-104) over records [r0, r1) from merged, sorted single-profile runs.
// Appends hits (unextended ranges) and, when KGMA_F_ALIGN, the extension requests; *genome_pos carries GenomePos across
// calls, so a genome can be replayed in record ranges as their run lists become available (pipelined streaming scan).
int replay_single_range(kgma_ctx *ctx, kgma_genome *g, const ProfTab &t, const kgma_scan_params &P,
                        std::vector<kgma_run> &runs, const std::vector<int64_t> &first_D, int r0, int r1,
                        int64_t *genome_pos_io, std::vector<kgma_hit> &hits, std::vector<AlignReq> &reqs, std::vector<Pending> &pend)
{
    const int k = t.k; const int64_t ws = t.ws, buff = P.buff;
    const bool do_align = (P.flags & KGMA_F_ALIGN) != 0;
    int64_t genome_pos = *genome_pos_io;
    size_t i = 0;
    // runs are ordered by (profile, record, t_first): skip to the first record of the range
    while (i < runs.size() && runs[i].record < r0) i++;
    for (int r = r0; r < r1; r++) {
        const int64_t L = g->recs[r].len;
        size_t b = i; while (i < runs.size() && runs[i].record == r) i++;
        if (L < ws) continue;                             // GenomeMiner.jl:37-39 (genome_pos not advanced)
        const int64_t steps = (P.only_record >= 0 && r != P.only_record) ? 0 : std::max<int64_t>(0, L - ws - (t.strobe ? 1 : 0));   // StrobeGenomeMiner.jl:45
        if (steps > 0 && i > b) {
            int64_t cur = first_D[r];                     // :57 currminim = kmerDist of the first window
            if (cur == INT64_MIN) return set_err(ctx, KGMA_E_STATE, "first-window distance of record %d missing", r);
            int64_t CMI = 2, goal = 0; bool stop = true;
            for (size_t j = b; j < i; j++) {
                kgma_run &ru = runs[j];
                if (ru.flags & KGMA_RUN_MARKER) continue;
                uint32_t hflags = ru.flags & (KGMA_HIT_NEAR_THR | KGMA_HIT_ARGMIN_TIE);
                // A run whose minimum EQUALS the minimum the state machine still carries (from the first window or from a run
                // whose hit goal_ind suppressed: currminim is only reset when a hit is emitted, :100-102) does not lower it here
                // -- exact repeats in a genome do this.  The reference's Float64 accumulator has drifted by ~1e-13 between the
                // two windows, so whether ITS `kmerDist < currminim` (:83) fires depends on its rounding history: flag the run.
                if (ru.D_min == cur) ru.flags |= KGMA_HIT_ARGMIN_TIE;
                if (ru.D_min < cur) { cur = ru.D_min; CMI = (t.strobe ? 0 : k - 1) + ru.t_argmin; stop = false; }   // :82-87 CMI = i_left (StrobeGMA!: CMI = i)
                if (ru.t_last >= steps) break;            // run still open at the record end: never emitted (A.1 step 5)
                if (!stop) {                              // :90-104 at step t_last+1
                    stop = true; CMI += 1;
                    if (CMI > goal) {
                        goal = CMI + ws - 1;
                        int64_t a = std::max<int64_t>(CMI - buff, 1), bb = std::min<int64_t>(CMI + ws - 1 + buff, L);
                        kgma_hit h{};
                        h.record = r; h.profile = 0; h.cmi = CMI; h.first = a; h.last = bb; h.genome_pos = genome_pos;
                        h.D = cur; h.dist = (double)cur / t.denom; h.flags = hflags | round_half_flag(h.dist);
                        if (do_align) { pend.push_back({ hits.size(), reqs.size(), a, j }); reqs.push_back({ r, 0, a, bb, align_hint(cur, t.T) }); }
                        hits.push_back(h);
                        cur = INT64_MAX;                  // :102 currminim = kmerDist (some value >= thr)
                    }
                }
            }
        }
        genome_pos += L;                                  // :106
    }
    *genome_pos_io = genome_pos;
    return KGMA_OK;
}

int replay(kgma_ctx *ctx, kgma_genome *g, const std::vector<ProfTab> &tabs, const kgma_profile *profiles,
           const kgma_scan_params &P, std::vector<kgma_run> &runs, const std::vector<int64_t> &first_D,
           kgma_result *res, std::vector<kgma_run_ext> *ext)
{
    const int C = (int)tabs.size(), nr = (int)g->recs.size(), k = tabs[0].k;
    const bool cluster = P.mode == KGMA_MODE_CLUSTER;
    const bool do_align = (P.flags & KGMA_F_ALIGN) != 0;
    const bool want_cig = (P.flags & KGMA_F_WANT_CIGARS) != 0;
    const int64_t buff = P.buff;
    int64_t maxws = 0; for (auto &t : tabs) maxws = std::max(maxws, t.ws);
    const bool trace = getenv("KGMA_TRACE") != nullptr;
    auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tr0 = tnow();
    merge_runs(runs, ext);
    const double tr1 = tnow();
    res->hits.clear(); res->cigar_ops.clear(); res->cigar_cnt.clear(); res->align_events.clear();

    // index runs per (profile, record)
    struct Span { size_t b = 0, e = 0; };
    std::vector<Span> span((size_t)C * nr);
    for (size_t i = 0; i < runs.size();) {
        size_t j = i;
        while (j < runs.size() && runs[j].profile == runs[i].profile && runs[j].record == runs[i].record) j++;
        if (runs[i].profile < 0 || runs[i].profile >= C || runs[i].record < 0 || runs[i].record >= nr)
            return set_err(ctx, KGMA_E_ARG, "run with invalid profile/record index");
        span[(size_t)runs[i].profile * nr + runs[i].record] = { i, j };
        i = j;
    }
    auto steps_of = [&](int r) -> int64_t {
        if (P.only_record >= 0 && r != P.only_record) return 0;
        int64_t L = g->recs[r].len;
        return std::max<int64_t>(0, cluster ? L - maxws - k + 2 : L - maxws - (tabs[0].strobe ? 1 : 0));
    };

    // (scratch kept per thread: a few hundred KB per call otherwise, each a fresh mmap)
    static thread_local std::vector<AlignReq> reqs; static thread_local std::vector<AlignRes> ares;
    static thread_local std::vector<Pending> pend;        // single mode: hits waiting for their extension
    reqs.clear(); ares.clear(); pend.clear();

    if (!cluster) {
        // ---------------- ac_gma_testing! ----------------
        int64_t genome_pos = 0;
        int rc = replay_single_range(ctx, g, tabs[0], P, runs, first_D, 0, nr, &genome_pos, res->hits, reqs, pend);
        if (rc) return rc;
        const double tr2 = tnow();
        if (trace) fprintf(stderr, "[kgma replay] merge %.3f ms, state machine %.3f ms (%zu runs, %zu requests)\n", tr1 - tr0, tr2 - tr1, runs.size(), reqs.size());
        if (do_align && !reqs.empty() && ext) {
            // the shards extended every run's own candidate window already (kgma_scan_shard): look the results up
            ares.assign(reqs.size(), AlignRes{ 1, 0, 0, 0, 0 });
            for (const Pending &p : pend) {
                const kgma_run_ext &e = (*ext)[p.run];
                if (e.lo == 0) return set_err(ctx, KGMA_E_STATE, "run %zu emits a hit but carries no extension result", p.run);
                ares[p.req] = AlignRes{ e.lo, e.hi, e.score, 0, 0 };
            }
            apply_extensions(g, res->hits, pend, ares);
        } else if (do_align && !reqs.empty()) {
            rc = align_batch_device(ctx, g, reqs, profiles, 1, true, P.gap_open, P.gap_extend,
                                    (P.flags & KGMA_F_TIE_OPEN) != 0, want_cig, ares,
                                    want_cig ? &res->cigar_ops : nullptr, want_cig ? &res->cigar_cnt : nullptr);
            if (rc) return rc;
            apply_extensions(g, res->hits, pend, ares);
        }
        if (tabs[0].strobe && do_align) {
            // process_hit! (Alignment.jl): `if aligned_obj.value < score_threshold; return end` -- the hit is dropped AFTER
            // goal_ind was advanced, so the state machine above is unaffected
            const int64_t thr_score = P.score_threshold;
            res->hits.erase(std::remove_if(res->hits.begin(), res->hits.end(), [&](const kgma_hit &h) { return h.align_score < thr_score; }), res->hits.end());
        }
        if (ctx) ctx->stats.n_align = (int64_t)reqs.size();
        return KGMA_OK;
    }

    // ---------------- Omn_KmerGMA! ----------------
    // Extension results feed back into the state machine (prev_hit_range, :139,:152), and only a fraction of the terminated
    // runs is ever extended by the reference (most are suppressed by `CMI in prev_hit_range`).  Extending every terminated
    // run up front costs several times the necessary work, so the replay runs in rounds: a pass over the events uses the
    // extension results it has and SPECULATES (unextended range) where one is missing, collecting exactly those requests;
    // they are extended in one batch and the pass is repeated.  A pass that needed no speculation is the reference's
    // sequential result.  Everything before a record's first missing result is exact, so every round makes progress; after a
    // few rounds the remaining terminated runs are simply all extended.
    struct Ev { int64_t end_step; int q; size_t run; };
    std::vector<Ev> evs;                                  // all records, in order
    std::vector<size_t> ev_begin((size_t)nr + 1, 0);
    for (int r = 0; r < nr; r++) {
        ev_begin[(size_t)r] = evs.size();
        if (steps_of(r) <= 0) continue;
        // run-end events of all profiles ordered by (end step, profile index): the reference visits profiles in index order
        // inside each loop step (:95).  Every profile's runs are already ordered, so this is a C-way merge.
        size_t head[MAX_PROFILES], tail[MAX_PROFILES];
        for (int q = 0; q < C; q++) { Span sp = span[(size_t)q * nr + r]; head[q] = sp.b; tail[q] = sp.e; }
        for (;;) {
            int best = -1; int64_t be = 0;
            for (int q = 0; q < C; q++) {
                while (head[q] < tail[q] && (runs[head[q]].flags & KGMA_RUN_MARKER)) head[q]++;
                if (head[q] >= tail[q]) continue;
                const int64_t e = runs[head[q]].t_last + 1;
                if (best < 0 || e < be) { best = q; be = e; }
            }
            if (best < 0) break;
            evs.push_back({ be, best, head[best] });
            head[best]++;
        }
    }
    ev_begin[(size_t)nr] = evs.size();

    const double tr2c = tnow();
    if (trace) fprintf(stderr, "[kgma replay] cluster: merge %.3f ms, event lists %.3f ms (%zu runs, %zu events)\n", tr1 - tr0, tr2c - tr1, runs.size(), evs.size());
    std::vector<char> have(runs.size(), 0);               // extension result of this run's candidate is known
    std::vector<AlignRes> res_of_run(runs.size());
    if (ext) for (size_t i = 0; i < runs.size(); i++) if ((*ext)[i].lo != 0) { have[i] = 1; res_of_run[i] = AlignRes{ (*ext)[i].lo, (*ext)[i].hi, (*ext)[i].score, 0, 0 }; }
    std::vector<size_t> missing;
    int64_t n_align_total = 0;
    std::vector<std::vector<kgma_hit>> rec_hits((size_t)nr);   // per record: a record whose pass needed no speculation is final
    std::vector<char> rec_done((size_t)nr, 0);
    std::vector<std::vector<kgma_align_event>> rec_events(want_cig ? (size_t)nr : 0);   // get_aligns: every extension, emitted or not (:133)
    for (int round = 0;; round++) {
        const double trr = tnow();
        missing.clear();
        int64_t genome_pos = 0;
        std::vector<int64_t> cur(C), CMIs(C, 1); std::vector<char> stop(C, 1);
        for (int r = 0; r < nr; r++) {
            const int64_t L = g->recs[r].len, steps = steps_of(r);
            if (steps > 0 && !rec_done[(size_t)r]) {
                const size_t missing_before = missing.size();
                std::vector<kgma_hit> &hits_r = rec_hits[(size_t)r];
                hits_r.clear();
                if (want_cig) rec_events[(size_t)r].clear();
                for (int q = 0; q < C; q++) { cur[q] = first_D[(size_t)q * nr + r]; CMIs[q] = 1; stop[q] = 1; }   // :73 curr_mins = first-window distance
                int64_t prev_a = 0, prev_b = 0;                                          // :59 prev_hit_range = 0:0
                for (size_t ei = ev_begin[(size_t)r]; ei < ev_begin[(size_t)r + 1]; ei++) {
                    const Ev &e = evs[ei];
                    kgma_run &ru = runs[e.run]; const int q = e.q;
                    if (cur[q] == INT64_MIN) return set_err(ctx, KGMA_E_STATE, "first-window distance of record %d profile %d missing", r, q);
                    if (ru.D_min == cur[q]) ru.flags |= KGMA_HIT_ARGMIN_TIE;            // tie with the carried minimum: see replay_single_range
                    if (ru.D_min < cur[q]) { cur[q] = ru.D_min; CMIs[q] = ru.t_argmin; stop[q] = 0; }   // :114-119 CMI = i
                    if (ru.t_last >= steps) continue;                                   // open at the end of the loop
                    if (stop[q]) continue;
                    stop[q] = 1;                                                        // :122
                    const int64_t CMI = CMIs[q];
                    if (CMI >= prev_a && CMI <= prev_b) continue;                       // :126
                    if (CMI != ru.t_argmin) return set_err(ctx, KGMA_E_STATE, "internal: extension request mismatch");
                    const int64_t wsq = tabs[q].ws;
                    const int64_t hl = std::max<int64_t>(CMI - buff, 1), hr = std::min<int64_t>(CMI + wsq - 1 + buff, L);
                    int64_t a = hl, b = hr; int64_t score = 0; uint32_t co = 0, cl = 0;
                    if (do_align) {
                        if (have[e.run]) {
                            const AlignRes &ar = res_of_run[e.run];
                            remap(hl, L, ar, &a, &b);
                            score = ar.score; co = ar.cig_off; cl = ar.cig_len;
                        } else missing.push_back(e.run);                               // speculate with the unextended range
                    }
                    const bool emit = b < prev_a || a > prev_b;                         // :139
                    if (want_cig && do_align && have[e.run])
                        rec_events[(size_t)r].push_back(kgma_align_event{ r, q + 1, CMI, score, co, cl, emit ? 1u : 0u, 0u });
                    if (emit) {
                        kgma_hit h{};
                        h.record = r; h.profile = q + 1; h.cmi = CMI; h.first = a; h.last = b; h.genome_pos = genome_pos;
                        h.D = cur[q]; h.dist = (double)cur[q] / tabs[q].denom;
                        h.flags = (ru.flags & (KGMA_HIT_NEAR_THR | KGMA_HIT_ARGMIN_TIE)) | round_half_flag(h.dist);
                        h.align_score = score; h.cigar_off = co; h.cigar_len = cl;
                        hits_r.push_back(h);
                        prev_a = a; prev_b = b;                                         // :152
                        cur[q] = INT64_MAX;                                             // :153 curr_mins[ind] = kmerDist
                    }
                }
                if (missing.size() == missing_before) rec_done[(size_t)r] = 1;
            }
            genome_pos += L;                                                            // :159 - every record
        }
        if (trace) fprintf(stderr, "[kgma replay] cluster: pass %d took %.3f ms, %zu results missing\n", round, tnow() - trr, missing.size());
        if (missing.empty()) break;                       // no speculation: this pass is the reference's result
        if (!ctx) return set_err(ctx, KGMA_E_STATE, "%zu runs emit hits but carry no extension result (host-only replay)", missing.size());
        if (round >= 3) {                                 // stop chasing: extend every terminated run that is still unknown
            missing.clear();
            for (size_t i = 0; i < runs.size(); i++) {
                const kgma_run &ru = runs[i];
                if ((ru.flags & KGMA_RUN_MARKER) || have[i] || ru.t_last >= steps_of(ru.record)) continue;
                missing.push_back(i);
            }
        }
        reqs.clear();
        for (size_t i : missing) {
            const kgma_run &ru = runs[i];
            const int64_t L = g->recs[ru.record].len, CMI = ru.t_argmin, wsq = tabs[ru.profile].ws;
            reqs.push_back({ ru.record, ru.profile, std::max<int64_t>(CMI - buff, 1), std::min<int64_t>(CMI + wsq - 1 + buff, L), align_hint(ru.D_min, tabs[ru.profile].T) });
        }
        int rc = align_batch_device(ctx, g, reqs, profiles, C, false, P.gap_open, P.gap_extend,
                                    (P.flags & KGMA_F_TIE_OPEN) != 0, want_cig, ares,
                                    want_cig ? &res->cigar_ops : nullptr, want_cig ? &res->cigar_cnt : nullptr);
        if (rc) return rc;
        for (size_t j = 0; j < missing.size(); j++) { res_of_run[missing[j]] = ares[j]; have[missing[j]] = 1; }
        if (getenv("KGMA_TRACE")) fprintf(stderr, "[kgma replay] round %d: %zu extensions\n", round, missing.size());
        n_align_total += (int64_t)reqs.size();
    }
    res->hits.clear();
    for (int r = 0; r < nr; r++) res->hits.insert(res->hits.end(), rec_hits[(size_t)r].begin(), rec_hits[(size_t)r].end());
    res->align_events.clear();
    if (want_cig) for (int r = 0; r < nr; r++) res->align_events.insert(res->align_events.end(), rec_events[(size_t)r].begin(), rec_events[(size_t)r].end());
    if (ctx) ctx->stats.n_align = n_align_total;
    return KGMA_OK;
}

}  // namespace kgma
