/*
 * kmergma_oracle.c — CPU restatement of KmerGMA.jl's homology-scan hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the timed CPU
 * baseline ("port").  Nothing in the product path (kmergma.jl_b200/) may import,
 * link or call it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * It follows the reference's own sequential algorithm and Float64 operation
 * order ("faithful" arithmetic), NOT the exact-integer / run-replay formulation
 * the CUDA product uses, so that comparing the two is a real parity check.
 *
 * Parity status: PINNED by every deterministic golden vector of the reference's
 * test-suite (test/test_folder/test-KmerGMA.jl, see tests/test_oracle_golden.py).
 * UNPINNED (no Julia / BioAlignments / Distances source in this container):
 *   - BioAlignments' choice between equal-score end columns and between
 *     gap-open and gap-extend on exact ties (alignment section below);
 *   - Distances.sqeuclidean's @simd summation order (last-ulp of first window);
 *   - matching of IUPAC symbols other than N in exactMatch.
 *
 * Citations are path:line in /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <ctype.h>

#define ORC_OK 0
#define ORC_E_IO (-1)
#define ORC_E_SYMBOL (-2)
#define ORC_E_CAP (-3)
#define ORC_E_ARG (-4)

/* ------------------------------------------------------------------ */
/* src/Consts.jl:22-28  NUCLEOTIDE_BITS: A0 C1 G2 T3, N -> 3.          */
/* Any other symbol is a KeyError in the reference -> -1 here.         */
static inline int nt_bits(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'N': case 'n': return 3;
    default: return -1;
    }
}

/* ------------------------------------------------------------------ */
/* FASTA container (stands in for FASTX.FASTA.Reader / Record).        */
typedef struct {
    int      n;
    char   **ident;   /* FASTA.identifier: header up to first whitespace */
    char   **desc;    /* FASTA.description: whole header line            */
    char   **seq;     /* upper-cased residues                            */
    int64_t *len;
    int      borrowed; /* seq[] point into caller-owned memory (orc_fasta_wrap): not freed here */
} orc_fasta;

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) abort(); return q; }

void orc_fasta_free(orc_fasta *f)
{
    if (!f) return;
    for (int i = 0; i < f->n; i++) { free(f->ident[i]); free(f->desc[i]); if (!f->borrowed) free(f->seq[i]); }
    free(f->ident); free(f->desc); free(f->seq); free(f->len); free(f);
}

/* One record over caller-owned, upper-case residues (no copy): lets the full-size parity tests hand a 250 Mb
 * contig to the scan functions below without writing it to disk first. */
orc_fasta *orc_fasta_wrap(const char *desc, char *seq, int64_t len)
{
    orc_fasta *f = (orc_fasta *)calloc(1, sizeof *f);
    f->n = 1; f->borrowed = 1;
    f->ident = (char **)xrealloc(NULL, sizeof(char *)); f->desc = (char **)xrealloc(NULL, sizeof(char *));
    f->seq = (char **)xrealloc(NULL, sizeof(char *)); f->len = (int64_t *)xrealloc(NULL, sizeof(int64_t));
    f->desc[0] = strdup(desc);
    size_t ie = 0; while (desc[ie] && !isspace((unsigned char)desc[ie])) ie++;
    f->ident[0] = strndup(desc, ie);
    f->seq[0] = seq; f->len[0] = len;
    return f;
}

orc_fasta *orc_fasta_read(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    char *buf = (char *)xrealloc(NULL, (size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, fp) != (size_t)sz) { fclose(fp); free(buf); return NULL; }
    fclose(fp); buf[sz] = 0;
    orc_fasta *f = (orc_fasta *)calloc(1, sizeof *f);
    int cap = 0; long i = 0;
    while (i < sz) {
        if (buf[i] != '>') { /* skip stray line */ while (i < sz && buf[i] != '\n') i++; i++; continue; }
        long hs = i + 1; while (i < sz && buf[i] != '\n') i++;
        long he = i; if (he > hs && buf[he - 1] == '\r') he--;
        if (f->n == cap) {
            cap = cap ? cap * 2 : 16;
            f->ident = (char **)xrealloc(f->ident, cap * sizeof(char *));
            f->desc = (char **)xrealloc(f->desc, cap * sizeof(char *));
            f->seq = (char **)xrealloc(f->seq, cap * sizeof(char *));
            f->len = (int64_t *)xrealloc(f->len, cap * sizeof(int64_t));
        }
        int r = f->n++;
        f->desc[r] = strndup(buf + hs, (size_t)(he - hs));
        long ie = hs; while (ie < he && !isspace((unsigned char)buf[ie])) ie++;
        f->ident[r] = strndup(buf + hs, (size_t)(ie - hs));
        /* sequence lines until next '>' at line start */
        long ss = i + 1, j = ss; size_t L = 0;
        char *s = (char *)xrealloc(NULL, (size_t)(sz - ss) + 2);
        while (j < sz) {
            if (buf[j] == '>' && (j == 0 || buf[j - 1] == '\n')) break;
            char c = buf[j++];
            if (c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
            s[L++] = (char)toupper((unsigned char)c);
        }
        s[L] = 0; f->seq[r] = (char *)xrealloc(s, L + 1); f->len[r] = (int64_t)L;
        i = j;
    }
    free(buf);
    return f;
}

int orc_fasta_n(const orc_fasta *f) { return f->n; }
int64_t orc_fasta_len(const orc_fasta *f, int r) { return f->len[r]; }
const char *orc_fasta_seq(const orc_fasta *f, int r) { return f->seq[r]; }
const char *orc_fasta_ident(const orc_fasta *f, int r) { return f->ident[r]; }
const char *orc_fasta_desc(const orc_fasta *f, int r) { return f->desc[r]; }

/* ------------------------------------------------------------------ */
/* src/Kmers.jl:33-44  kmer_count!: prime k-1 symbols, then for i=k..len
 * kmer = ((kmer<<2)&mask)|code ; bins[kmer+1] += 1                     */
int orc_kmer_count_add(const char *s, int64_t len, int k, double *bins)
{
    uint64_t mask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1), kmer = 0;
    for (int64_t i = 0; i < len; i++) {
        int b = nt_bits(s[i]); if (b < 0) return ORC_E_SYMBOL;
        if (i < k - 1) kmer = (kmer << 2) | (uint64_t)b;
        else { kmer = ((kmer << 2) & mask) | (uint64_t)b; bins[kmer] += 1.0; }
    }
    return ORC_OK;
}

/* src/Kmers.jl:14-28 kmer_count (fresh zero bins). */
int orc_kmer_count(const char *s, int64_t len, int k, double *bins)
{
    memset(bins, 0, sizeof(double) * ((size_t)1 << (2 * k)));
    return orc_kmer_count_add(s, len, k, bins);
}

/* Distances.sqeuclidean(a,b) = sum (a_i-b_i)^2, restated as a plain left-to-right loop
 * (the @simd order of Distances 0.10 is unpinned). */
static double sqeuclidean(const double *a, const double *b, size_t n)
{
    double s = 0.0;
    for (size_t i = 0; i < n; i++) { double d = a[i] - b[i]; s += d * d; }
    return s;
}

/* src/Kmers.jl:58-60  kmer_dist(seq, KFV, k) = (1/(2k)) * sqeuclidean(kmer_count(seq), KFV) */
int orc_kmer_dist_kfv(const char *s, int64_t len, const double *kfv, int k, double *out)
{
    size_t nb = (size_t)1 << (2 * k);
    double *bins = (double *)calloc(nb, sizeof(double));
    int rc = orc_kmer_count_add(s, len, k, bins);
    if (rc == ORC_OK) *out = (1.0 / (double)(2 * k)) * sqeuclidean(bins, kfv, nb);
    free(bins);
    return rc;
}

/* src/Kmers.jl:54-56 kmer_dist(seq1, seq2, k) */
int orc_kmer_dist_seq(const char *s1, int64_t l1, const char *s2, int64_t l2, int k, double *out)
{
    size_t nb = (size_t)1 << (2 * k);
    double *b2 = (double *)calloc(nb, sizeof(double));
    int rc = orc_kmer_count_add(s2, l2, k, b2);
    if (rc == ORC_OK) rc = orc_kmer_dist_kfv(s1, l1, b2, k, out);
    free(b2);
    return rc;
}

/* ------------------------------------------------------------------ */
/* src/Consensus.jl:6-48  Profile / add_consensus! / lengthen! / consensus_seq.
 * vecs[b][j] counts symbol b at column j; consensus = strict '>' scan over A,C,G,T. */
typedef struct { int64_t *v[4]; int64_t len; } orc_profile;

static void profile_init(orc_profile *p, int64_t len)
{
    p->len = len;
    for (int b = 0; b < 4; b++) p->v[b] = (int64_t *)calloc((size_t)(len > 0 ? len : 1), sizeof(int64_t));
}
static void profile_free(orc_profile *p) { for (int b = 0; b < 4; b++) free(p->v[b]); }
static void profile_lengthen(orc_profile *p, int64_t nl)          /* Consensus.jl:24-33 */
{
    if (nl <= p->len) return;
    for (int b = 0; b < 4; b++) {
        p->v[b] = (int64_t *)xrealloc(p->v[b], (size_t)nl * sizeof(int64_t));
        for (int64_t j = p->len; j < nl; j++) p->v[b][j] = 0;
    }
    p->len = nl;
}
static int profile_add(orc_profile *p, const char *s, int64_t len)  /* Consensus.jl:16-20 */
{
    for (int64_t i = 0; i < len; i++) {
        int b = nt_bits(s[i]); if (b < 0) return ORC_E_SYMBOL;
        if (i >= p->len) return ORC_E_ARG;     /* BoundsError in the reference */
        p->v[b][i] += 1;
    }
    return ORC_OK;
}
static void profile_consensus(const orc_profile *p, char *out)      /* Consensus.jl:37-48 */
{
    static const char sym[4] = { 'A', 'C', 'G', 'T' };
    for (int64_t j = 0; j < p->len; j++) {
        int64_t best = p->v[0][j]; char c = 'A';
        for (int b = 1; b < 4; b++) if (p->v[b][j] > best) { best = p->v[b][j]; c = sym[b]; }
        out[j] = c;
    }
    out[p->len] = 0;
}

/* exported thin wrappers so the Consensus.jl golden test can be replayed */
orc_profile *orc_profile_new(int64_t len) { orc_profile *p = (orc_profile *)calloc(1, sizeof *p); profile_init(p, len); return p; }
void orc_profile_del(orc_profile *p) { profile_free(p); free(p); }
int orc_profile_add(orc_profile *p, const char *s, int64_t len) { return profile_add(p, s, len); }
void orc_profile_lengthen(orc_profile *p, int64_t nl) { profile_lengthen(p, nl); }
int64_t orc_profile_len(const orc_profile *p) { return p->len; }
void orc_profile_counts(const orc_profile *p, int b, int64_t *out) { memcpy(out, p->v[b], (size_t)p->len * sizeof(int64_t)); }
void orc_profile_consensus(const orc_profile *p, char *out) { profile_consensus(p, out); }

/* ------------------------------------------------------------------ */
/* src/ReferenceGeneration.jl:4-41  gen_ref_ws_cons
 * RV = (sum of counts over all refs) .* (1/N); ws = Int(round(sum_len*(1/N))) (half-even);
 * consensus over a profile lengthened to the longest ref.             */
int orc_gen_ref_ws_cons(const orc_fasta *refs, int k, double *rv, int64_t *ws,
                        char *consensus /* cap maxlen+1 */, int64_t *maxlen_out)
{
    size_t nb = (size_t)1 << (2 * k);
    memset(rv, 0, nb * sizeof(double));
    int64_t cum = 0, maxlen = 0; int len = 0;
    orc_profile p; profile_init(&p, 1);
    for (int r = 0; r < refs->n; r++) {
        len += 1; cum += refs->len[r]; if (refs->len[r] > maxlen) maxlen = refs->len[r];
        int rc = orc_kmer_count_add(refs->seq[r], refs->len[r], k, rv);
        if (rc) { profile_free(&p); return rc; }
        profile_lengthen(&p, refs->len[r]);
        rc = profile_add(&p, refs->seq[r], refs->len[r]);
        if (rc) { profile_free(&p); return rc; }
    }
    double inv = 1.0 / (double)len;                       /* :33  len = 1/len */
    for (size_t i = 0; i < nb; i++) rv[i] = rv[i] * inv;  /* :35/:40 answer .* len */
    *ws = (int64_t)nearbyint((double)cum * inv);          /* Int(round(x)) half-even */
    if (consensus) profile_consensus(&p, consensus);
    if (maxlen_out) *maxlen_out = p.len > maxlen ? p.len : maxlen;
    profile_free(&p);
    return ORC_OK;
}

/* ReferenceGeneration.jl:50-57 get_cluster_index: first cutoff with inp <= cutoff (1-based), else n+1 */
int orc_get_cluster_index(double inp, const double *cutoffs, int n)
{
    int answer = 1;
    for (int i = 0; i < n; i++) { if (inp <= cutoffs[i]) return answer; answer++; }
    return answer;
}

/* ReferenceGeneration.jl:75-138 cluster_ref_API + :152-168 eliminate_null_params.
 * Outputs the surviving clusters compacted in order; with include_avg the overall
 * mean profile (full-length consensus) is appended last.
 * kfvs: [maxC][4^k]; wss: [maxC]; cons: [maxC][maxlen+1]; members: [maxC]; invalid: [maxC] (pre-elimination)
 * Returns number of surviving profiles (>=0) or error (<0). */
int orc_cluster_ref(const orc_fasta *refs, int k, const double *cutoffs, int ncut, int include_avg,
                    int eliminate_null,
                    double *kfvs, int64_t *wss, char *cons, int64_t cons_stride,
                    int *members, int *invalid_out, double *dists_out)
{
    size_t nb = (size_t)1 << (2 * k);
    int nc = ncut + 1;
    double *avg = (double *)calloc(nb, sizeof(double));
    int64_t avg_ws = 0, maxlen = 0;
    char *avg_cons = (char *)calloc((size_t)cons_stride + 1, 1);
    int rc = orc_gen_ref_ws_cons(refs, k, avg, &avg_ws, NULL, &maxlen);
    if (rc) { free(avg); free(avg_cons); return rc; }
    if (maxlen + 1 > cons_stride) { free(avg); free(avg_cons); return ORC_E_CAP; }
    orc_gen_ref_ws_cons(refs, k, avg, &avg_ws, avg_cons, &maxlen);

    double **K = (double **)calloc((size_t)nc, sizeof(double *));
    orc_profile *P = (orc_profile *)calloc((size_t)nc, sizeof(orc_profile));
    int64_t *wsum = (int64_t *)calloc((size_t)nc, sizeof(int64_t));
    int *lens = (int *)calloc((size_t)nc, sizeof(int));
    for (int c = 0; c < nc; c++) { K[c] = (double *)calloc(nb, sizeof(double)); profile_init(&P[c], maxlen); }

    for (int r = 0; r < refs->n && rc == ORC_OK; r++) {            /* :97-112 */
        double d;
        rc = orc_kmer_dist_kfv(refs->seq[r], refs->len[r], avg, k, &d);
        if (rc) break;
        if (dists_out) dists_out[r] = d;
        int ci = orc_get_cluster_index(d, cutoffs, ncut) - 1;
        rc = profile_add(&P[ci], refs->seq[r], refs->len[r]);
        wsum[ci] += refs->len[r]; lens[ci] += 1;
        if (rc == ORC_OK) rc = orc_kmer_count_add(refs->seq[r], refs->len[r], k, K[ci]);
    }
    int out = 0;
    if (rc == ORC_OK) {
        char *tmp = (char *)calloc((size_t)maxlen + 1, 1);
        for (int c = 0; c < nc; c++) {                             /* :114-125 */
            int inval = (lens[c] == 0);
            if (invalid_out) invalid_out[c] = inval;
            if (inval && eliminate_null) continue;
            double *dst = kfvs + (size_t)out * nb;
            char *cdst = cons + (size_t)out * (size_t)cons_stride;
            if (!inval) {
                for (size_t i = 0; i < nb; i++) dst[i] = K[c][i] / (double)lens[c];      /* ./= lens */
                wss[out] = (int64_t)nearbyint((double)wsum[c] / (double)lens[c]);
                profile_consensus(&P[c], tmp);
                memcpy(cdst, tmp, (size_t)wss[out]); cdst[wss[out]] = 0;                 /* [1:ws_i] */
            } else {
                memset(dst, 0, nb * sizeof(double)); wss[out] = 0; cdst[0] = 0;
            }
            if (members) members[out] = lens[c];
            out++;
        }
        free(tmp);
        if (include_avg) {                                          /* :127-132 */
            if (invalid_out) invalid_out[nc] = 0;
            memcpy(kfvs + (size_t)out * nb, avg, nb * sizeof(double));
            wss[out] = avg_ws;
            strcpy(cons + (size_t)out * (size_t)cons_stride, avg_cons);
            if (members) members[out] = refs->n;
            out++;
        }
    }
    for (int c = 0; c < nc; c++) { free(K[c]); profile_free(&P[c]); }
    free(K); free(P); free(wsum); free(lens); free(avg); free(avg_cons);
    return rc ? rc : out;
}

/* ------------------------------------------------------------------ */
/* Hit extension.  src/Alignment.jl:33-52 calls BioAlignments
 *   pairalign(SemiGlobalAlignment(), a = consensus[1:ws], b = seq[range],
 *             AffineGapScoreModel(EDNAFULL, gap_open, gap_extend))
 * BioAlignments is a third-party dependency NOT vendored in /root/reference
 * (Project.toml has no compat bound for it).  Its published algorithm is
 * restated here: Gotoh affine-gap DP; `a` aligned end to end, gaps that only
 * consume `b` before a[1] and after a[m] are free; a gap of length L costs
 * gap_open + L*gap_extend; EDNAFULL = NCBI NUC.4.4 restricted to A,C,G,T,N.
 * Traceback priority in the H state: match/mismatch > delete (consumes b) >
 * insert (consumes a).  Anchoring: Alignment.jl's own tests (test-KmerGMA.jl:
 * 129-145,188-192,222-225,245-248,259-262,267-270).
 * UNPINNED: on an exact tie between opening and extending a gap the restatement
 * extends (prefer_extend=1); this also selects the leftmost of equal-score end
 * columns in the free last row.                                           */
static inline int ednafull(char x, char y)
{
    int nx = (x == 'N'), ny = (y == 'N');
    if (nx && ny) return -1;
    if (nx || ny) return -2;
    return x == y ? 5 : -4;
}

#define TR_MATCH 1
#define TR_DEL   2
#define TR_INS   4
#define TR_EXTF  8   /* deletion state came from extension */
#define TR_EXTE  16  /* insertion state came from extension */

/* ops_out: chars '=','X','I','D'; cnt_out: run lengths; returns #ops or <0 */
int orc_semiglobal(const char *a, int m, const char *b, int n, int gap_open, int gap_extend,
                   int prefer_extend, char *ops_out, int32_t *cnt_out, int cap, int64_t *score_out)
{
    const int64_t NEG = INT64_MIN / 4;
    int64_t go = -(int64_t)gap_open, ge = -(int64_t)gap_extend;
    size_t W = (size_t)n + 1;
    uint8_t *tr = (uint8_t *)calloc((size_t)(m + 1) * W, 1);
    int64_t *H = (int64_t *)malloc(W * sizeof(int64_t));   /* row i-1 then row i */
    int64_t *E = (int64_t *)malloc(W * sizeof(int64_t));   /* insertion state per column */
    if (!tr || !H || !E) abort();
    for (int j = 0; j <= n; j++) { H[j] = 0; E[j] = NEG; tr[j] = TR_DEL; }   /* free leading deletions */
    tr[0] = 0;
    for (int i = 1; i <= m; i++) {
        int64_t hdiag = H[0];
        H[0] = -(go + (int64_t)i * ge);
        tr[(size_t)i * W] = TR_INS | (i > 1 ? TR_EXTE : 0);
        int64_t F = NEG, hleft = H[0];
        int last = (i == m);
        for (int j = 1; j <= n; j++) {
            uint8_t t = 0;
            /* insertion: consumes a_i, from row i-1 same column */
            int64_t eo = H[j] - go - ge, ee = E[j] - ge;
            int64_t e;
            if (prefer_extend ? (ee >= eo) : (ee > eo)) { e = ee; t |= TR_EXTE; } else e = eo;
            /* deletion: consumes b_j, from same row column j-1; free in the last row */
            int64_t fo = last ? hleft : hleft - go - ge, fe = last ? F : F - ge;
            int64_t f;
            if (prefer_extend ? (fe >= fo) : (fe > fo)) { f = fe; t |= TR_EXTF; } else f = fo;
            int64_t mm = hdiag + ednafull(a[i - 1], b[j - 1]);
            int64_t h = mm; if (f > h) h = f; if (e > h) h = e;
            if (mm == h) t |= TR_MATCH;
            if (f == h) t |= TR_DEL;
            if (e == h) t |= TR_INS;
            tr[(size_t)i * W + (size_t)j] = t;
            hdiag = H[j]; H[j] = h; E[j] = e; F = f; hleft = h;
        }
    }
    if (score_out) *score_out = H[n];
    /* traceback from (m,n); ops collected reversed */
    int nops = 0; int rc = 0;
    char *rops = (char *)malloc((size_t)(m + n + 2));
    int i = m, j = n, state = 0; size_t L = 0;
    while (i > 0 || j > 0) {
        if (i == 0) { rops[L++] = 'D'; j--; continue; }
        if (j == 0) { rops[L++] = 'I'; i--; continue; }
        uint8_t t = tr[(size_t)i * W + (size_t)j];
        if (state == 0) {
            if (t & TR_MATCH) { rops[L++] = (a[i - 1] == b[j - 1]) ? '=' : 'X'; i--; j--; }
            else if (t & TR_DEL) state = 1;
            else state = 2;
        } else if (state == 1) {
            rops[L++] = 'D'; if (!(t & TR_EXTF)) state = 0; j--;
        } else {
            rops[L++] = 'I'; if (!(t & TR_EXTE)) state = 0; i--;
        }
    }
    for (size_t p = L; p-- > 0;) {
        char c = rops[p];
        if (nops > 0 && ops_out[nops - 1] == c) cnt_out[nops - 1]++;
        else { if (nops >= cap) { rc = ORC_E_CAP; break; } ops_out[nops] = c; cnt_out[nops] = 1; nops++; }
    }
    free(rops); free(tr); free(H); free(E);
    return rc ? rc : nops;
}

/* src/Alignment.jl:13-30 cigar_to_UnitRange: lower = count of the FIRST op,
 * num_sum = sum of counts of every op except the LAST; returns (lower+1):num_sum.
 * (The loop returns when it reaches the final character, i.e. before adding the last op.) */
void orc_cigar_to_unitrange(const int32_t *cnt, int nops, int64_t *lo, int64_t *hi)
{
    int64_t lower = 0, num_sum = 0;
    for (int i = 0; i + 1 < nops; i++) { if (i == 0) lower = cnt[0]; num_sum += cnt[i]; }
    *lo = lower + 1; *hi = num_sum;
}

/* src/Alignment.jl:33-52 align_unitrange; positions 1-based inclusive. */
int orc_align_unitrange(const char *seq, int64_t L, int64_t first, int64_t last,
                        const char *cons, int cons_len, int gap_open, int gap_extend,
                        int prefer_extend, int64_t *nfirst, int64_t *nlast)
{
    int n = (int)(last - first + 1);
    int cap = cons_len + n + 2;
    char *ops = (char *)malloc((size_t)cap); int32_t *cnt = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    int nops = orc_semiglobal(cons, cons_len, seq + (first - 1), n, gap_open, gap_extend, prefer_extend, ops, cnt, cap, NULL);
    if (nops < 0) { free(ops); free(cnt); return nops; }
    int64_t lo, hi; orc_cigar_to_unitrange(cnt, nops, &lo, &hi);
    int64_t a = first + lo - 1; if (a < 1) a = 1;
    int64_t b = first + hi - 1; if (b > L) b = L;
    if (b < a - 1) b = a - 1;                      /* Julia UnitRange normalisation */
    *nfirst = a; *nlast = b;
    free(ops); free(cnt);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
typedef struct {
    int32_t record;      /* 0-based record index in file order */
    int32_t kfv;         /* 1-based profile index (cluster mode), 0 in single mode */
    double  dist;        /* currminim at emission, unrounded */
    int64_t first, last; /* 1-based MatchPos (post-alignment if aligned) */
    int64_t genome_pos;
    int64_t cmi;         /* CMI used for the hit (diagnostics) */
} orc_hit;

/* ------------------------------------------------------------------ */
/* EXACT-ARITHMETIC MODE (test switch).  The reference accumulates the distance in Float64 over the whole
 * record, so `d < thr` at an exact tie d == thr, and the choice between two windows attaining the same run
 * minimum, depend on its rounding history (and on Distances.sqeuclidean's unpinned summation order).  With
 * RV = S/N the same state machine can be run without any rounding: D = sum (N c_i - S_i)^2 = d * 2kN^2 is an
 * integer < 2^53, so holding D in a double and sliding it with integer increments is exact.  When enabled
 * (orc_set_exact), orc_ac_gma / orc_omn_gma run the identical control flow on D and T = ceil(thr * 2kN^2)
 * and report dist = D / (2kN^2).  Disabled (the default), they are the faithful Float64 restatement. */
static int g_exact_n = 0;
static int32_t g_exact_N[64];
void orc_set_exact(int n, const int32_t *N)
{
    g_exact_n = (n > 64) ? 64 : (n < 0 ? 0 : n);
    for (int i = 0; i < g_exact_n; i++) g_exact_N[i] = N[i];
}
/* exact ceil(x * m), x >= 0 a double, m > 0 */
static int64_t ceil_times(double x, int64_t m)
{
    if (!(x > 0)) return 0;
    int e; double f = frexp(x, &e);
    int64_t mant = (int64_t)ldexp(f, 53); e -= 53;          /* x = mant * 2^e exactly */
    __int128 p = (__int128)mant * (__int128)m;
    if (e >= 0) return (int64_t)(p << e);
    int sft = -e;
    if (sft > 120) return 1;
    __int128 q = p >> sft;
    if ((q << sft) != p) q += 1;
    return (int64_t)q;
}

/* src/GenomeMiner.jl:4-109 ac_gma_testing!  (also MultiThread/GenomeMiner.jl:8-98
 * record_KmerGMA! — same loop for one record; CMI=i_left+1 folded at :73).
 * dist_out (optional): L-ws values per scanned record, first window excluded (:79). */
int orc_ac_gma(const orc_fasta *g, const double *RV, const char *cons, int cons_len,
               int k, int64_t ws, double thr, int64_t buff, int do_align,
               int gap_open, int gap_extend, int prefer_extend,
               orc_hit *hits, int64_t hit_cap, int64_t *nhits,
               double *dist_out, int64_t dist_cap, int64_t *ndist,
               int only_record /* -1 = all */)
{
    size_t nb = (size_t)1 << (2 * k);
    uint64_t mask = nb - 1;
    double SF = 1.0 / (double)k;                 /* API.jl:86  ScaleFactor = 1/k */
    double SF0 = SF * 0.5;                       /* GenomeMiner.jl:29 */
    int64_t *c = (int64_t *)malloc(nb * sizeof(int64_t));
    double *cf = (double *)malloc(nb * sizeof(double));
    int64_t genome_pos = 0, nh = 0, nd = 0; int rc = ORC_OK;
    const int exact = g_exact_n > 0;                         /* test switch, see orc_set_exact */
    double EN = 0, Eden = 1, thr_eff = thr, *ES = NULL;
    if (exact) {
        EN = (double)g_exact_N[0]; Eden = 2.0 * (double)k * EN * EN;
        thr_eff = (double)ceil_times(thr, (int64_t)Eden);
        ES = (double *)malloc(nb * sizeof(double));
        for (size_t i = 0; i < nb; i++) ES[i] = (double)llround(RV[i] * EN);
    }
    for (int r = 0; r < g->n && rc == ORC_OK; r++) {
        if (only_record >= 0 && r != only_record) continue;
        const char *s = g->seq[r]; int64_t L = g->len[r];
        if (L < ws) continue;                    /* :37-39 — genome_pos NOT advanced */
        memset(c, 0, nb * sizeof(int64_t));
        {   /* :42-44 kmer_count! on the first window */
            uint64_t km = 0;
            for (int64_t i = 0; i < ws; i++) {
                int b = nt_bits(s[i]); if (b < 0) { rc = ORC_E_SYMBOL; break; }
                if (i < k - 1) km = (km << 2) | (uint64_t)b;
                else { km = ((km << 2) & mask) | (uint64_t)b; c[km] += 1; }
            }
            if (rc) break;
        }
        for (size_t i = 0; i < nb; i++) cf[i] = (double)c[i];
        double d = SF0 * sqeuclidean(RV, cf, nb);            /* :46-47 */
        if (exact) { d = 0; for (size_t i = 0; i < nb; i++) { double x = EN * cf[i] - ES[i]; d += x * x; } }
        uint64_t lk = 0, rk = 0;
        for (int64_t i = 0; i < k - 1; i++) lk = (lk << 2) | (uint64_t)nt_bits(s[i]);           /* :49-51 */
        for (int64_t i = ws - k + 1; i < ws; i++) rk = (rk << 2) | (uint64_t)nt_bits(s[i]);     /* :53-55 */
        int64_t CMI = 2, goal = 0; int stop = 1; double cur = d;                                  /* :57 */
        for (int64_t t = 1; t <= L - ws; t++) {              /* :60 i_left=k-1+t, i_right=ws+t (1-based) */
            int64_t il = k - 1 + t, ir = ws + t;
            int bl = nt_bits(s[il - 1]), br = nt_bits(s[ir - 1]);
            if (bl < 0 || br < 0) { rc = ORC_E_SYMBOL; break; }
            lk = ((lk << 2) & mask) | (uint64_t)bl;
            rk = ((rk << 2) & mask) | (uint64_t)br;
            if (lk != rk) {                                  /* :69-77, left-to-right Float64 */
                double x = (double)(1 + c[rk]);
                if (exact) d += 2.0 * EN * (EN * (x - (double)c[lk]) + ES[lk] - ES[rk]);
                else { x = x + RV[lk]; x = x - RV[rk]; x = x - (double)c[lk]; d += SF * x; }
                c[lk] -= 1; c[rk] += 1;
            }
            if (dist_out) { if (nd >= dist_cap) { rc = ORC_E_CAP; break; } dist_out[nd] = exact ? d / Eden : d; }
            nd++;
            if (d < thr_eff) {                               /* :82-87 */
                if (d < cur) { cur = d; CMI = il; stop = 0; }
            } else if (!stop) {                              /* :90-104 */
                stop = 1; CMI += 1;
                if (CMI > goal) {
                    goal = CMI + ws - 1;
                    int64_t a = CMI - buff; if (a < 1) a = 1;
                    int64_t b = CMI + ws - 1 + buff; if (b > L) b = L;
                    if (do_align) {
                        rc = orc_align_unitrange(s, L, a, b, cons, (int)ws, gap_open, gap_extend, prefer_extend, &a, &b);
                        if (rc) break;
                    }
                    if (nh >= hit_cap) { rc = ORC_E_CAP; break; }
                    orc_hit *h = &hits[nh++];
                    h->record = r; h->kfv = 0; h->dist = exact ? cur / Eden : cur; h->first = a; h->last = b;
                    h->genome_pos = genome_pos; h->cmi = CMI;
                    cur = d;
                }
            }
        }
        genome_pos += L;                                     /* :106 */
    }
    (void)cons_len;
    free(c); free(cf); free(ES);
    *nhits = nh; if (ndist) *ndist = nd;
    return rc;
}

/* ------------------------------------------------------------------ */
/* Experimental strobemer path, src/StrobemerGMA/.  PARITY UNPINNED for the scan itself: the reference's tests
 * (test-StrobemerGMA.jl) pin randstrobe_score, get_strobe_2_mer and ungapped_strobe_2_mer_count only; StrobeGMA! /
 * Strobemer_findGenes have no test and no golden.  What follows restates them line by line from those pinned pieces.
 *
 * Strobemers.jl:12-14 randstrobe_score = (as_UInt(s1) + as_UInt(s2)) % q.
 * Strobemers.jl:45-65 get_strobe_2_mer(seq[1:k], s, w_min, w_max, q): first strobe seq[1:s]; the running minimum is seeded
 * with `2 << 63`, which is 0 in Int64, and updated on `<=`: the second strobe starts at the LAST i in w_min..w_max whose
 * score is 0, else at w_min.  Returns as_UInt(first * second) (the gap-free strobemer, first base most significant). */
static int strobe_code(const char *s, int sl, int wmin, int wmax, int q, uint64_t *out)
{
    uint64_t f = 0;
    for (int i = 0; i < sl; i++) { int b = nt_bits(s[i]); if (b < 0) return ORC_E_SYMBOL; f = (f << 2) | (uint64_t)b; }
    int64_t min_score = 0; int min_ind = wmin;                       /* 2 << 63 == 0 */
    for (int i = wmin; i <= wmax; i++) {
        uint64_t v = 0;
        for (int j = 0; j < sl; j++) { int b = nt_bits(s[i - 1 + j]); if (b < 0) return ORC_E_SYMBOL; v = (v << 2) | (uint64_t)b; }
        int64_t cur = (int64_t)((f + v) % (uint64_t)q);
        if (cur <= min_score) { min_score = cur; min_ind = i; }
    }
    uint64_t v = 0;
    for (int j = 0; j < sl; j++) v = (v << 2) | (uint64_t)nt_bits(s[min_ind - 1 + j]);
    *out = (f << (2 * sl)) | v;
    return ORC_OK;
}

/* Strobemers.jl:105-115 ungapped_strobe_2_mer_count!: bins[code of every k = w_max+s-1 window] += 1 */
int orc_strobe_count_add(const char *s, int64_t len, int sl, int wmin, int wmax, int q, double *bins)
{
    int k = wmax + sl - 1;
    for (int64_t i = 0; i + k <= len; i++) {
        uint64_t c; int rc = strobe_code(s + i, sl, wmin, wmax, q, &c);
        if (rc) return rc;
        bins[c] += 1.0;
    }
    return ORC_OK;
}

/* StrobeRefGen.jl:4-42 gen_ref_ws_cons(refs; s, w_min, w_max, q): 2 << (4s - 1) = 4^(2s) bins */
int orc_strobe_gen_ref_ws_cons(const orc_fasta *refs, int sl, int wmin, int wmax, int q, double *rv, int64_t *ws,
                               char *consensus, int64_t *maxlen_out)
{
    size_t nb = (size_t)1 << (4 * sl);
    memset(rv, 0, nb * sizeof(double));
    int64_t cum = 0, maxlen = 0; int len = 0;
    orc_profile p; profile_init(&p, 1);
    for (int r = 0; r < refs->n; r++) {
        len += 1; cum += refs->len[r]; if (refs->len[r] > maxlen) maxlen = refs->len[r];
        int rc = orc_strobe_count_add(refs->seq[r], refs->len[r], sl, wmin, wmax, q, rv);
        if (rc) { profile_free(&p); return rc; }
        profile_lengthen(&p, refs->len[r]);
        rc = profile_add(&p, refs->seq[r], refs->len[r]);
        if (rc) { profile_free(&p); return rc; }
    }
    double inv = 1.0 / (double)len;
    for (size_t i = 0; i < nb; i++) rv[i] = rv[i] * inv;
    *ws = (int64_t)nearbyint((double)cum * inv);
    if (consensus) profile_consensus(&p, consensus);
    if (maxlen_out) *maxlen_out = p.len > maxlen ? p.len : maxlen;
    profile_free(&p);
    return ORC_OK;
}

/* StrobeGenomeMiner.jl:5-95 StrobeGMA!  (Strobemer_findGenes :119-158 calls it with ScaleFactor = 1/(w_max+s-1)).
 * Differences from ac_gma_testing!: codes are strobemer codes; the loop runs i = 1 .. L-ws-1 (:49); the ENTERING code is the
 * one at i+ws-k (:55), i.e. the last code of the window that is being left, so the tracked table is "the ws-k codes from i+1
 * on, plus a permanent copy of the first window's last code"; CMI = i (:79); a hit whose alignment scores below
 * score_threshold is dropped after goal_ind was advanced (Alignment.jl process_hit!). */
int orc_strobe_gma(const orc_fasta *g, const double *RV, const char *cons, int cons_len,
                   int sl, int wmin, int wmax, int q, int64_t ws, double thr, int64_t buff, int do_align,
                   int gap_open, int gap_extend, int prefer_extend, int64_t score_threshold,
                   orc_hit *hits, int64_t hit_cap, int64_t *nhits,
                   double *dist_out, int64_t dist_cap, int64_t *ndist)
{
    const int k = wmax + sl - 1;
    size_t nb = (size_t)1 << (4 * sl);
    double SF = 1.0 / (double)k;                                   /* :139 ScaleFactor = 1/(w_max+s-1) */
    double *cf = (double *)calloc(nb, sizeof(double));
    int64_t genome_pos = 0, nh = 0, nd = 0; int rc = ORC_OK;
    const int exact = g_exact_n > 0;
    double EN = 0, Eden = 1, thr_eff = thr, *ES = NULL;
    if (exact) {
        EN = (double)g_exact_N[0]; Eden = 2.0 * (double)k * EN * EN;
        thr_eff = (double)ceil_times(thr, (int64_t)Eden);
        ES = (double *)malloc(nb * sizeof(double));
        for (size_t i = 0; i < nb; i++) ES[i] = (double)llround(RV[i] * EN);
    }
    for (int r = 0; r < g->n && rc == ORC_OK; r++) {
        const char *s = g->seq[r]; int64_t L = g->len[r];
        if (L < ws) continue;                                      /* :33 */
        memset(cf, 0, nb * sizeof(double));
        rc = orc_strobe_count_add(s, ws, sl, wmin, wmax, q, cf);   /* :36-38 */
        if (rc) break;
        double d = (1.0 / (double)(2 * k)) * sqeuclidean(RV, cf, nb);   /* :40 */
        if (exact) { d = 0; for (size_t i = 0; i < nb; i++) { double x = EN * cf[i] - ES[i]; d += x * x; } }
        int64_t CMI = 2, goal = 0; int stop = 1; double cur = d;   /* :43 */
        for (int64_t i = 1; i <= L - ws - 1; i++) {                /* :45 */
            uint64_t lk, rk;
            rc = strobe_code(s + (i - 1), sl, wmin, wmax, q, &lk);             /* :47-49 view(seq, i:i+k-1) */
            if (!rc) rc = strobe_code(s + (i + ws - k - 1), sl, wmin, wmax, q, &rk);   /* :52-54 view(seq, i+ws-k:i+ws) */
            if (rc) break;
            if (lk != rk) {                                        /* :57-65 */
                double x = 1.0 + cf[rk];
                if (exact) d += 2.0 * EN * (EN * (x - cf[lk]) + ES[lk] - ES[rk]);
                else { x = x + RV[lk]; x = x - RV[rk]; x = x - cf[lk]; d += SF * x; }
                cf[lk] -= 1.0; cf[rk] += 1.0;
            }
            if (dist_out) { if (nd >= dist_cap) { rc = ORC_E_CAP; break; } dist_out[nd] = exact ? d / Eden : d; }
            nd++;
            if (d < thr_eff) {                                     /* :70-75 */
                if (d < cur) { cur = d; CMI = i; stop = 0; }
            } else if (!stop) {                                    /* :78-88 */
                stop = 1; CMI += 1;
                if (CMI > goal) {
                    goal = CMI + ws - 1;
                    int64_t a = CMI - buff; if (a < 1) a = 1;
                    int64_t b = CMI + ws - 1 + buff; if (b > L) b = L;
                    int keep = 1;
                    if (do_align) {                                /* Alignment.jl process_hit! */
                        int n = (int)(b - a + 1), cap = (int)ws + n + 2;
                        char *ops = (char *)malloc((size_t)cap); int32_t *cnt = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
                        int64_t score = 0;
                        int nops = orc_semiglobal(cons, (int)ws, s + (a - 1), n, gap_open, gap_extend, prefer_extend, ops, cnt, cap, &score);
                        if (nops < 0) { free(ops); free(cnt); rc = nops; break; }
                        if (score < score_threshold) keep = 0;
                        else {
                            int64_t lo, hi; orc_cigar_to_unitrange(cnt, nops, &lo, &hi);
                            int64_t na = a + lo - 1; if (na < 1) na = 1;
                            int64_t nbb = a + hi - 1; if (nbb > L) nbb = L;
                            if (nbb < na - 1) nbb = na - 1;
                            a = na; b = nbb;
                        }
                        free(ops); free(cnt);
                    }
                    if (keep) {
                        if (nh >= hit_cap) { rc = ORC_E_CAP; break; }
                        orc_hit *h = &hits[nh++];
                        h->record = r; h->kfv = 0; h->dist = exact ? cur / Eden : cur; h->first = a; h->last = b;
                        h->genome_pos = genome_pos; h->cmi = CMI;
                    }
                    cur = d;                                       /* :87 (after process_hit!, kept or not) */
                }
            }
        }
        genome_pos += L;                                           /* :91 */
    }
    (void)cons_len;
    free(cf); free(ES);
    *nhits = nh; if (ndist) *ndist = nd;
    return rc;
}

/* src/OmnGenomeMiner.jl:7-162 Omn_KmerGMA!  — C profiles; count tables are Float64 (:46);
 * prev_hit_range shared by all profiles, reset per record (:59).
 * cons: C strings at stride cons_stride; the WHOLE consensus_seqs[ind] is aligned (:131). */
/* Test hook for `get_aligns` (OmnGenomeMiner.jl:131-133): while a sink is set, orc_omn_gma records every extension it performs,
 * in order, as 6 int64: record (0-based), profile (1-based), CMI, hit_left, hit_right (the range handed to align_unitrange),
 * emitted (1 when the second overlap test :139 then lets the hit through). */
static int64_t *g_ev_buf = NULL, g_ev_cap = 0, *g_ev_n = NULL;
void orc_set_event_sink(int64_t *buf, int64_t cap, int64_t *n) { g_ev_buf = buf; g_ev_cap = cap; g_ev_n = n; if (n) *n = 0; }

int orc_omn_gma(const orc_fasta *g, const double *RVs, const int64_t *wss, int C,
                const char *cons, int64_t cons_stride,
                int k, const double *thr, int64_t buff, int align_hits,
                int gap_open, int gap_extend, int prefer_extend,
                orc_hit *hits, int64_t hit_cap, int64_t *nhits,
                double *dist_out /* [C][dist_cap] */, int64_t dist_cap, int64_t *ndist /* [C] */)
{
    size_t nb = (size_t)1 << (2 * k);
    uint64_t mask = nb - 1;
    double SF = 1.0 / (double)k;
    double *cnt = (double *)calloc(nb * (size_t)C, sizeof(double));
    double *kd = (double *)calloc((size_t)C, sizeof(double));
    double *curmin = (double *)malloc((size_t)C * sizeof(double));
    int64_t *CMIs = (int64_t *)malloc((size_t)C * sizeof(int64_t));
    int *stops = (int *)malloc((size_t)C * sizeof(int));
    uint64_t *rk = (uint64_t *)calloc((size_t)C, sizeof(uint64_t));
    int64_t maxws = 0, nh = 0, genome_pos = 0; int rc = ORC_OK;
    for (int q = 0; q < C; q++) { curmin[q] = 10000.0; CMIs[q] = 1; stops[q] = 1; if (wss[q] > maxws) maxws = wss[q]; if (ndist) ndist[q] = 0; }
    const int exact = g_exact_n >= C;                         /* test switch, see orc_set_exact */
    double *ES = NULL, *EN = NULL, *Eden = NULL, *thr_eff = (double *)malloc((size_t)C * sizeof(double));
    for (int q = 0; q < C; q++) thr_eff[q] = thr[q];
    if (exact) {
        ES = (double *)malloc(nb * (size_t)C * sizeof(double)); EN = (double *)malloc((size_t)C * sizeof(double)); Eden = (double *)malloc((size_t)C * sizeof(double));
        for (int q = 0; q < C; q++) {
            EN[q] = (double)g_exact_N[q]; Eden[q] = 2.0 * (double)k * EN[q] * EN[q];
            thr_eff[q] = (double)ceil_times(thr[q], (int64_t)Eden[q]);
            for (size_t i = 0; i < nb; i++) ES[(size_t)q * nb + i] = (double)llround(RVs[(size_t)q * nb + i] * EN[q]);
        }
    }
    for (int r = 0; r < g->n && rc == ORC_OK; r++) {
        const char *s = g->seq[r]; int64_t L = g->len[r];
        int64_t prev_a = 0, prev_b = 0;                       /* :59 prev_hit_range = 0:0 */
        for (int q = 0; q < C; q++) {                         /* :61-82 */
            if (L < wss[q]) continue;
            double *c = cnt + (size_t)q * nb;
            memset(c, 0, nb * sizeof(double));
            rc = orc_kmer_count_add(s, wss[q], k, c); if (rc) break;
            kd[q] = curmin[q] = SF * 0.5 * sqeuclidean(RVs + (size_t)q * nb, c, nb);
            if (exact) { double a = 0; for (size_t i = 0; i < nb; i++) { double x = EN[q] * c[i] - ES[(size_t)q * nb + i]; a += x * x; } kd[q] = curmin[q] = a; }
            CMIs[q] = 1; stops[q] = 1; rk[q] = 0;
            for (int64_t i = wss[q] - k + 1; i < wss[q]; i++) rk[q] = (rk[q] << 2) | (uint64_t)nt_bits(s[i]);
        }
        if (rc) break;
        uint64_t lk = 0;
        if (L >= k - 1) for (int64_t i = 0; i < k - 1; i++) { int b = nt_bits(s[i]); if (b < 0) { rc = ORC_E_SYMBOL; break; } lk = (lk << 2) | (uint64_t)b; }
        if (rc) break;
        int64_t nsteps = L - maxws + 1 - k + 1;               /* :89 view(seq, k:L-maxws+1) */
        for (int64_t i = 1; i <= nsteps && rc == ORC_OK; i++) {
            int bl = nt_bits(s[k - 1 + i - 1]); if (bl < 0) { rc = ORC_E_SYMBOL; break; }
            lk = ((lk << 2) & mask) | (uint64_t)bl;            /* :92-93 */
            for (int q = 0; q < C; q++) {
                double *c = cnt + (size_t)q * nb; const double *RV = RVs + (size_t)q * nb;
                int br = nt_bits(s[i + wss[q] - 1]); if (br < 0) { rc = ORC_E_SYMBOL; break; }
                rk[q] = ((rk[q] << 2) & mask) | (uint64_t)br;  /* :97 */
                uint64_t rr = rk[q];
                if (lk != rr) {                                /* :101-108 */
                    double x = 1.0 + c[rr];
                    if (exact) kd[q] += 2.0 * EN[q] * (EN[q] * (x - c[lk]) + ES[(size_t)q * nb + lk] - ES[(size_t)q * nb + rr]);
                    else { x = x + RV[lk]; x = x - RV[rr]; x = x - c[lk]; kd[q] += SF * x; }
                    c[lk] -= 1.0; c[rr] += 1.0;
                }
                double d = kd[q];
                if (dist_out) { if (ndist[q] >= dist_cap) { rc = ORC_E_CAP; break; } dist_out[(size_t)q * (size_t)dist_cap + (size_t)ndist[q]] = exact ? d / Eden[q] : d; }
                if (ndist) ndist[q]++;
                if (d < thr_eff[q]) {                          /* :114-119 */
                    if (d < curmin[q]) { curmin[q] = d; CMIs[q] = i; stops[q] = 0; }
                } else if (!stops[q]) {                        /* :122-156 */
                    stops[q] = 1;
                    int64_t CMI = CMIs[q];
                    if (!(CMI >= prev_a && CMI <= prev_b)) {   /* :126 */
                        int64_t hl = CMI - buff; if (hl < 1) hl = 1;
                        int64_t hr = CMI + wss[q] - 1 + buff; if (hr > L) hr = L;
                        int64_t a = hl, b = hr;
                        if (align_hits) {                      /* :130-136 whole consensus */
                            const char *cq = cons + (size_t)q * (size_t)cons_stride;
                            rc = orc_align_unitrange(s, L, hl, hr, cq, (int)strlen(cq), gap_open, gap_extend, prefer_extend, &a, &b);
                            if (rc) break;
                        }
                        if (align_hits && g_ev_buf && g_ev_n && *g_ev_n < g_ev_cap) {
                            int64_t *e = g_ev_buf + 6 * (*g_ev_n)++;
                            e[0] = r; e[1] = q + 1; e[2] = CMI; e[3] = hl; e[4] = hr; e[5] = (b < prev_a || a > prev_b) ? 1 : 0;
                        }
                        if (b < prev_a || a > prev_b) {        /* :139 */
                            if (nh >= hit_cap) { rc = ORC_E_CAP; break; }
                            orc_hit *h = &hits[nh++];
                            h->record = r; h->kfv = q + 1; h->dist = exact ? curmin[q] / Eden[q] : curmin[q]; h->first = a; h->last = b;
                            h->genome_pos = genome_pos; h->cmi = CMI;
                            prev_a = a; prev_b = b;            /* :152 */
                            curmin[q] = d;                     /* :153 */
                        }
                    }
                }
            }
        }
        genome_pos += L;                                       /* :159 — every record */
    }
    free(cnt); free(kd); free(curmin); free(CMIs); free(stops); free(rk); free(ES); free(EN); free(Eden); free(thr_eff);
    *nhits = nh;
    return rc;
}

/* ------------------------------------------------------------------ */
/* src/ExactMatch.jl:20-43,89-98.  BioSequences.findfirst(ExactSearchQuery(q), seq)
 * (third-party, not vendored) returns the first range where every symbol isequal;
 * FindAllOverlap resumes at match_start+1 (:38), FindAll at match_end+1 (:27).
 * Symbols are compared as upper-cased characters, so N only equals N.
 * starts_out: 1-based match starts. */
int orc_exact_match(const char *q, int64_t ql, const char *s, int64_t L, int overlap,
                    int64_t *starts_out, int64_t cap, int64_t *n_out)
{
    int64_t n = 0, start = 0;
    if (ql <= 0) { *n_out = 0; return ORC_E_ARG; }
    while (start + ql <= L) {
        int64_t p = -1;
        for (int64_t i = start; i + ql <= L; i++) {
            if (s[i] == q[0] && memcmp(s + i, q, (size_t)ql) == 0) { p = i; break; }
        }
        if (p < 0) break;
        if (n >= cap) return ORC_E_CAP;
        starts_out[n++] = p + 1;
        start = overlap ? p + 1 : p + ql;
    }
    *n_out = n;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* Bench helper: the ac_gma hot loop (GenomeMiner.jl:60-87 + hit bookkeeping without
 * alignment) over an in-memory upper-case sequence; used by bench.py's cpu_baseline
 * so that FASTA parsing is outside the timed region.  Returns number of hits. */
int64_t orc_ac_gma_seq(const char *s, int64_t L, const double *RV, int k, int64_t ws,
                       double thr, int64_t buff, orc_hit *hits, int64_t hit_cap)
{
    orc_fasta f; char *sq = (char *)s; int64_t len = L; char *id = (char *)"seq";
    f.n = 1; f.ident = &id; f.desc = &id; f.seq = &sq; f.len = &len;
    int64_t nh = 0;
    int rc = orc_ac_gma(&f, RV, NULL, 0, k, ws, thr, buff, 0, -69, -1, 1, hits, hit_cap, &nh, NULL, 0, NULL, -1);
    return rc ? rc : nh;
}

/* ------------------------------------------------------------------ */
/* Synthetic genome generator shared with the device generator (kgma_genome_synth,
 * SURVEY §8d): base(p) = splitmix64(seed ^ p) & 3 over the global (padded) base
 * coordinate p.  Writes n upper-case residues for p0 .. p0+n-1. */
static inline uint64_t orc_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void orc_synth(uint64_t seed, int64_t p0, int64_t n, char *out)
{
    static const char nt[4] = { 'A', 'C', 'G', 'T' };
    for (int64_t i = 0; i < n; i++) out[i] = nt[orc_splitmix64(seed ^ (uint64_t)(p0 + i)) & 3];
}
