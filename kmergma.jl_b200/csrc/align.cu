// Batched semi-global affine-gap extension of candidate hits, one warp per candidate.
// Replaces align_unitrange (src/Alignment.jl:33-52), i.e. BioAlignments'
//   pairalign(SemiGlobalAlignment(), consensus, view(seq, range), AffineGapScoreModel(EDNAFULL, gap_open, gap_extend))
// followed by cigar_to_UnitRange (src/Alignment.jl:13-30).
//
// DP (Gotoh): a = consensus is aligned end to end; gaps that consume only b (= genome slice) before a[1]
// and after a[m] are free; a gap of length L costs gap_open + L*gap_extend; EDNAFULL restricted to
// A,C,G,T,N (match 5, mismatch -4, N/base -2, N/N -1).  Traceback priority in the H state:
// match > delete (consumes b) > insert (consumes a); on exact open/extend ties the gap is extended
// unless KGMA_F_TIE_OPEN (BioAlignments' own tie rule is not pinned by any reference test).
//
// Three kernels.  kgma_align_tagged<C> (the default): one 32-bit word per DP state (score | tie tag | path counters), lanes
// own subject columns, a DPX three-way max per cell; it hands the rare alignment without a leading or trailing deletion run
// to kgma_align_summary<R>: each lane owns R = 10 consecutive DP rows in registers, lanes are
// skewed by one column (operands of the row above arrive by __shfl_up), so one sweep of n+31 steps covers 320
// consensus rows; instead of a trace matrix a packed summary of the canonical optimal path is carried with the
// scores, from which cigar_to_UnitRange's two numbers follow; the subject is read from the packed genome on the
// device.  kgma_align (only when CIGARs are requested): 32 rows per sweep, trace bytes packed four columns at a time
// into 32-bit global stores, lane-0 traceback.
#include "kgma_internal.h"
#include <chrono>
#include <algorithm>

namespace kgma {

#define TR_MATCH 1u
#define TR_DEL   2u
#define TR_INS   4u
#define TR_EXTF  8u
#define TR_EXTE  16u

struct AlignJob {
    int64_t tr_off;      // byte offset of this job's trace matrix
    int32_t a_off, m;    // consensus codes
    int32_t b_off, n;    // subject codes
    int32_t cig_off;     // offset into the cigar buffer (uint32 entries), capacity m+n+2
    int32_t pad;
};


struct AlignArgs {
    const uint8_t *a, *b;        // codes 0..3, 4 = N
    const AlignJob *jobs; int njobs;
    int *next_job;
    uint8_t *trace;
    uint32_t *cigar;             // (count << 8) | op, reversed order; may be null
    AlignOut *out;
    int go, ge;                  // positive penalties
    int tie_open;
    int ncol_cap;                // shared-memory columns per warp
};

__device__ __forceinline__ int edna(int x, int y)
{
    if ((x | y) & 4) return (x & y & 4) ? -1 : -2;
    return x == y ? 5 : -4;
}

__global__ void __launch_bounds__(128) kgma_align(AlignArgs A)
{
    extern __shared__ int s_bound[];                     // per warp: Hb[ncol_cap], Eb[ncol_cap]
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *Hb = s_bound + (size_t)wid * 2 * A.ncol_cap, *Eb = Hb + A.ncol_cap;
    const int NEG = -(1 << 29);
    const int go = A.go, ge = A.ge;

    for (;;) {
        int ji = 0;
        if (lane == 0) ji = atomicAdd(A.next_job, 1);
        ji = __shfl_sync(FULL, ji, 0);
        if (ji >= A.njobs) break;
        const AlignJob J = A.jobs[ji];
        const int m = J.m, n = J.n;
        const uint8_t *a = A.a + J.a_off, *b = A.b + J.b_off;
        const int stride = (n + 1 + 3) & ~3;
        uint8_t *tr = A.trace + J.tr_off;
        int score = 0;
        const int nrb = (m + 31) >> 5;
        for (int rb = 0; rb < nrb; rb++) {
            const int i = rb * 32 + lane + 1;
            const bool row_ok = i <= m;
            const int ai = row_ok ? a[i - 1] : 0;
            const bool last = (i == m);
            int Hleft = -(go + i * ge), F = NEG;
            int Hdiag = (i == 1) ? 0 : -(go + (i - 1) * ge);      // H[i-1][0]
            int Hout = NEG, Eout = NEG;
            uint32_t tp = 0;
            uint8_t *trow = tr + (size_t)i * stride;
            for (int s = 1; s <= n + 31; s++) {
                const int j = s - lane;
                int upH = __shfl_up_sync(FULL, Hout, 1), upE = __shfl_up_sync(FULL, Eout, 1);
                if (lane == 0 && j <= n) {
                    if (rb == 0) { upH = 0; upE = NEG; }           // row 0: free leading deletions
                    else { upH = Hb[j]; upE = Eb[j]; }
                }
                if (row_ok && j >= 1 && j <= n) {
                    uint32_t t = 0;
                    const int eo = upH - go - ge, ee = upE - ge;
                    int e;
                    if (A.tie_open ? (ee > eo) : (ee >= eo)) { e = ee; t |= TR_EXTE; } else e = eo;
                    const int fo = last ? Hleft : Hleft - go - ge, fe = last ? F : F - ge;
                    int f;
                    if (A.tie_open ? (fe > fo) : (fe >= fo)) { f = fe; t |= TR_EXTF; } else f = fo;
                    const int mm = Hdiag + edna(ai, b[j - 1]);
                    int h = max(mm, max(f, e));
                    if (mm == h) t |= TR_MATCH;
                    if (f == h) t |= TR_DEL;
                    if (e == h) t |= TR_INS;
                    Hdiag = upH; Hout = h; Eout = e; Hleft = h; F = f;
                    tp |= t << (8 * (j & 3));
                    if ((j & 3) == 3 || j == n) { *reinterpret_cast<uint32_t *>(trow + (j & ~3)) = tp; tp = 0; }
                    if (lane == 31) { Hb[j] = h; Eb[j] = e; }      // bottom row of this block -> next block's row above
                }
            }
            if (rb == nrb - 1) score = __shfl_sync(FULL, Hleft, (m - 1) & 31);
            __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) {
            // traceback from (m,n); ops come out in reverse order
            int i = m, j = n, state = 0;
            int nops = 0, cur_op = 0, cur_cnt = 0, c_first = 0, c_last = 0; long long total = 0;
            uint32_t *cg = A.cigar ? A.cigar + J.cig_off : nullptr;
            auto flush = [&]() {
                if (cur_cnt > 0) {
                    if (nops == 0) c_last = cur_cnt;
                    c_first = cur_cnt;
                    if (cg) cg[nops] = ((uint32_t)cur_cnt << 8) | (uint32_t)cur_op;
                    nops++; total += cur_cnt;
                }
            };
            auto push = [&](int op) { if (op == cur_op) cur_cnt++; else { flush(); cur_op = op; cur_cnt = 1; } };
            while (i > 0 || j > 0) {
                if (i == 0) { push('D'); j--; continue; }
                if (j == 0) { push('I'); i--; continue; }
                const uint32_t t = tr[(size_t)i * stride + j];
                if (state == 0) {
                    if (t & TR_MATCH) { push(a[i - 1] == b[j - 1] ? '=' : 'X'); i--; j--; }
                    else if (t & TR_DEL) state = 1;
                    else state = 2;
                } else if (state == 1) { push('D'); if (!(t & TR_EXTF)) state = 0; j--; }
                else { push('I'); if (!(t & TR_EXTE)) state = 0; i--; }
            }
            flush();
            // cigar_to_UnitRange (Alignment.jl:13-30): lower = count of the first op, num_sum = all ops but the last
            AlignOut o;
            o.score = score; o.nops = nops; o.cig_n = cg ? nops : 0;
            o.lower = nops >= 2 ? c_first : 0;
            o.num_sum = nops >= 2 ? (int)(total - c_last) : 0;
            A.out[ji] = o;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Trace-free variant (the default; the trace kernel above is only used when CIGARs are requested).
// cigar_to_UnitRange needs three numbers of the optimal path only: the length of its first run, the length of
// its last run and its total number of columns.  Those are carried forward with the scores as a packed
// "path summary" per DP state (H, E, F), choosing the predecessor with exactly the priorities the traceback
// uses (match > delete > insert; extend-vs-open as configured), so no trace matrix is written or walked.
// The subject is read straight from the packed genome on the device (2-bit plane + ambiguity plane).
// path summary, two 32-bit words:
//   lo = total columns (bits 0-15) | length of the first run << 16   (the first-run field stays 0 while the path is
//        still ONE run and is captured from `total` at the first change of op)
//   hi = last-run length (0-15) | last op << 16 | (>= 2 runs) << 18 | non-empty << 19
// Gap states do not append per cell: E / F keep the summary of the H cell the gap was opened from plus the gap
// length, and the run is appended only when H actually selects the gap (rare with EDNAFULL and gap_open << 0).
#define PS_EQ 0u
#define PS_X  1u
#define PS_I  2u
#define PS_D  3u
#define PS_MULTI (1u << 18)
#define PS_NE    (1u << 19)
struct PathSum { uint32_t lo, hi; };
__device__ __forceinline__ PathSum ps_run(uint32_t len, uint32_t op)
{
    PathSum r; r.lo = len; r.hi = len ? (len | (op << 16) | PS_NE) : 0u; return r;
}
__device__ __forceinline__ PathSum ps_append(PathSum s, uint32_t op, uint32_t len)
{
    PathSum r;
    const bool same = ((s.hi ^ (op << 16)) & ((3u << 16) | PS_NE)) == PS_NE;          // non-empty and same op
    const uint32_t cap = (same || (s.hi & PS_MULTI)) ? 0u : (s.lo << 16);           // first change of op: first run = total so far
    r.lo = s.lo + len + cap;
    r.hi = same ? s.hi + len : (len | (op << 16) | PS_NE | ((s.hi >> 1) & PS_MULTI));
    return r;
}

// Each lane owns R consecutive DP rows (register blocked), the warp 32*R rows per sweep; neighbouring lanes are
// skewed by one column, so a whole 289-row alignment is ONE sweep of n+31 steps with R = 10 (instead of ten
// sweeps of 32 rows): 7 shuffles per step are amortised over R cells and the per-step bookkeeping shrinks 10x.
template <int R, bool TIE_OPEN>
__global__ void __launch_bounds__(128, 3) kgma_align_summary(AlignArgs2 A)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t per_warp = (size_t)A.ncol_cap * (A.need_boundary ? (6 * 4 + 1) : 1);
    unsigned char *wb = s_raw + (size_t)wid * ((per_warp + 15) & ~(size_t)15);
    // bottom row of a sweep, parked for the next sweep (only when the consensus is longer than 32*R)
    int *Hb = reinterpret_cast<int *>(wb), *Eb = Hb + A.ncol_cap;
    uint32_t *sHlo = reinterpret_cast<uint32_t *>(Eb + A.ncol_cap), *sHhi = sHlo + A.ncol_cap;
    uint32_t *bElo = sHhi + A.ncol_cap, *bEhi = bElo + A.ncol_cap;
    uint8_t *bs = A.need_boundary ? reinterpret_cast<uint8_t *>(bEhi + A.ncol_cap) : wb;
    const int NEG = -(1 << 29);
    const int go = A.go, ge = A.ge;

    for (;;) {
        int ji = 0;
        if (lane == 0) ji = atomicAdd(A.next_job, 1);
        ji = __shfl_sync(FULL, ji, 0);
        if (ji >= A.njobs) break;
        const AlignJob2 J = A.jobs[ji];
        const int m = J.m, n = J.n;
        const uint8_t *a = A.a + J.a_off;
        // stage the subject codes: 2-bit code from the packed genome, 4 = N inside a masked run
        if (J.b_off >= 0) { for (int j = lane; j < n; j += 32) bs[j] = A.b[J.b_off + j]; }
        else for (int j = lane; j < n; j += 32) {
            const long long gp = J.gpos + j;
            uint32_t c = (__ldg(A.seq + (gp >> 4)) >> (2 * (int)(gp & 15))) & 3u;
            if (A.n_nruns) {
                int lo = -1, hi = A.n_nruns;                       // last run with start <= gp
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (A.nruns[2 * mid] <= gp) lo = mid; else hi = mid; }
                if (lo >= 0 && gp < A.nruns[2 * lo + 1]) c = 4u;
            }
            bs[j] = (uint8_t)c;
        }
        __syncwarp();
        int score = 0; PathSum sfin = { 0u, 0u };
        const int nrb = (m + 32 * R - 1) / (32 * R);
        for (int rb = 0; rb < nrb; rb++) {
            const int i0 = rb * 32 * R + lane * R;                                // this lane owns rows i0+1 .. i0+R
            // per row: scores, H-path summary, and the deletion-gap state (summary at gap open; its hi word carries the
            // gap length in bits 20-31, which the summary itself never uses)
            int H[R], F[R]; PathSum sH[R], bF[R]; uint32_t sc[R]; int gof[R], gef[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = i0 + r + 1;
                H[r] = -(go + i * ge); F[r] = NEG;                                // H[i][0], F[i][0]
                sH[r] = ps_run((uint32_t)i, PS_I); bF[r].lo = bF[r].hi = 0;
                const int ai = i <= m ? a[i - 1] : 0;
                // EDNAFULL row for this consensus symbol: 4 bits per subject code (score + 4), the symbol itself in bits 20-22
                uint32_t t = (uint32_t)ai << 20;
#pragma unroll
                for (int c = 0; c < 5; c++) t |= (uint32_t)(edna(ai, c) + 4) << (4 * c);
                sc[r] = t;
                gof[r] = (i == m) ? 0 : go + ge; gef[r] = (i == m) ? 0 : ge;      // trailing deletions after the last row are free
            }
            int Hdiag0 = (i0 == 0) ? 0 : -(go + i0 * ge);                         // H[i0][0]: diagonal of my first row at column 1
            PathSum sHdiag0 = ps_run((uint32_t)i0, PS_I);
            int Hout = NEG, Eout = NEG; PathSum sHout = { 0u, 0u }, bEout = { 0u, 0u };
            const bool lane_ok = i0 < m;
            const int gog = go + ge;
            const uint32_t GL1 = 1u << 20, SUMM = GL1 - 1;                        // gap-length unit / summary bits of a hi word
            for (int s = 1; s <= n + 31; s++) {
                const int j = s - lane;
                int upH = __shfl_up_sync(FULL, Hout, 1), upE = __shfl_up_sync(FULL, Eout, 1);
                PathSum supH, ubE;
                supH.lo = __shfl_up_sync(FULL, sHout.lo, 1); supH.hi = __shfl_up_sync(FULL, sHout.hi, 1);
                ubE.lo = __shfl_up_sync(FULL, bEout.lo, 1); ubE.hi = __shfl_up_sync(FULL, bEout.hi, 1);
                if (lane == 0 && j >= 1 && j <= n) {
                    if (rb == 0) { upH = 0; upE = NEG; supH = ps_run((uint32_t)j, PS_D); ubE.lo = ubE.hi = 0; }   // row 0: free leading deletions
                    else { upH = Hb[j]; upE = Eb[j]; supH.lo = sHlo[j]; supH.hi = sHhi[j]; ubE.lo = bElo[j]; ubE.hi = bEhi[j]; }
                }
                if (lane_ok && j >= 1 && j <= n) {
                    const uint32_t bj = bs[j - 1];
                    const int bsh = 4 * (int)bj;
                    int hU = upH, eU = upE; PathSum shU = supH, beU = ubE;                        // row above, column j
                    int hD = Hdiag0; PathSum shD = sHdiag0;                                       // row above, column j-1
                    Hdiag0 = upH; sHdiag0 = supH;
                    // rows past the end of the consensus (last lane only) compute harmless garbage nobody reads
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int eo = hU - gog, ee = eU - ge;
                        const bool xe = TIE_OPEN ? (ee > eo) : (ee >= eo);
                        const int e = xe ? ee : eo;
                        PathSum bE; bE.lo = xe ? beU.lo : shU.lo; bE.hi = (xe ? beU.hi : shU.hi) + GL1;   // extend: +1; open: summary | 1
                        const int fo = H[r] - gof[r], fe = F[r] - gef[r];
                        const bool xf = TIE_OPEN ? (fe > fo) : (fe >= fo);
                        const int f = xf ? fe : fo;
                        bF[r].lo = xf ? bF[r].lo : sH[r].lo; bF[r].hi = (xf ? bF[r].hi : sH[r].hi) + GL1;
                        const int mm = hD + (int)((sc[r] >> bsh) & 15u) - 4;
                        const int h = max(mm, max(f, e));
                        PathSum sHn = ps_append(shD, (sc[r] >> 20) == bj ? PS_EQ : PS_X, 1u);
                        if (mm != h) {
                            PathSum g = (f == h) ? bF[r] : bE;
                            const uint32_t glen = g.hi >> 20; g.hi &= SUMM;
                            sHn = ps_append(g, (f == h) ? PS_D : PS_I, glen);
                        }
                        hD = H[r]; shD = sH[r];                                    // my old value (column j-1) is the next row's diagonal
                        H[r] = h; sH[r] = sHn; F[r] = f;
                        hU = h; eU = e; shU = sHn; beU = bE;
                    }
                    Hout = hU; Eout = eU; sHout = shU; bEout = beU;               // my last row at column j -> next lane
                    if (lane == 31 && rb + 1 < nrb) { Hb[j] = hU; Eb[j] = eU; sHlo[j] = shU.lo; sHhi[j] = shU.hi; bElo[j] = beU.lo; bEhi[j] = beU.hi; }
                }
            }
            if (rb == nrb - 1) {
                int hs = 0; PathSum ss = { 0u, 0u };
#pragma unroll
                for (int r = 0; r < R; r++) if (i0 + r + 1 == m) { hs = H[r]; ss = sH[r]; }
                const int src = ((m - 1) - rb * 32 * R) / R;
                score = __shfl_sync(FULL, hs, src); sfin.lo = __shfl_sync(FULL, ss.lo, src); sfin.hi = __shfl_sync(FULL, ss.hi, src);
            }
            __syncwarp();
        }
        if (lane == 0) {
            // cigar_to_UnitRange (Alignment.jl:13-30): lower = count of the first op, num_sum = all ops but the last
            AlignOut o;
            const bool multi = (sfin.hi & PS_MULTI) != 0;
            o.score = score; o.nops = multi ? 2 : ((sfin.hi & PS_NE) ? 1 : 0); o.cig_n = 0;
            o.lower = multi ? (int)(sfin.lo >> 16) : 0;
            o.num_sum = multi ? (int)((sfin.lo & 0xFFFFu) - (sfin.hi & 0xFFFFu)) : 0;
            A.out[ji] = o;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kgma_align_tagged: the default extension kernel (prefer-extend tie rule, consensus <= 400, subject <= 511 bases).
//
// One 32-bit word per DP state carries everything the state machine of the traceback would need:
//     [31:20] score (signed)   [19:18] priority tag   [17:9] insert columns so far   [8:0] start column
// so that ONE signed integer max both selects the better predecessor and moves its bookkeeping along.  The tag
// makes equal scores resolve exactly as BioAlignments' traceback does (match 2 > delete 1 > insert 0 in the H state,
// extend 1 > open 0 inside a gap); it is rewritten after every max, so it never reaches the fields below it.
// The H state is a three-way max: one DPX instruction (__vimax3_s32, VIMNMX3 on sm_90+/sm_100a).  12 instructions per
// cell against ~79 for the path-summary kernel above.
//
// Orientation: lanes own COLUMNS (C consecutive subject bases each, in registers, read straight from the packed genome)
// and the consensus rows stream by, lane l working on row s-l at step s; the only traffic between lanes is two
// shuffles per step (H and F of the column to the left).  All rows use the ordinary gap costs; the free trailing
// deletions of the semi-global alignment are taken at the end: the optimal path leaves the last row at the FIRST
// column attaining max_j H[m][j] (the traceback extends the zero-cost trailing gap on ties), so with
//   x = start column of that cell's path (leading free deletions), I = its insert columns, j* = that column:
//   cigar_to_UnitRange (Alignment.jl:13-30)  lower = x,  num_sum = j* + I      (first run xD, last run (n-j*)D).
// That leaves two kinds of alignment open.  (1) The maximum of the last row is (also) attained at column n: the path may end
// at (m, n) -- the consensus overhangs the subject slice, or ends exactly on its last base -- and its last CIGAR run is then
// not a deletion run.  The same warp sweeps the alignment a second time with the OTHER payload
//     [17:9] trailing insert ops of the path   [8:0] trailing diagonal ops   (both 0: the last op is a deletion, or none yet)
// under identical score / tag arithmetic (so every max picks the same predecessor), and finishes it:
//   the path ends at (m, n) iff the maximum is first attained at column n, or is attained there by a diagonal step (the
//   traceback tests match before the free deletion); its last run is the insert run the word counts, or the =/X run
//   min(trailing diagonal steps, length of the equal-kind stretch of the diagonal through (m, n)), read off the sequences;
//   cigar_to_UnitRange: lower = x of THAT path, num_sum = n + I - last run.
// (2) Paths that start at column 0 of the slice (x = 0: the first run is no deletion run; hits at a record edge, buff = 0)
// are flagged (nops = -1) and redone by kgma_align_summary.
#define TG_SC_SH   20
#define TG_TAG1    (1 << 18)
#define TG_TAG2    (2 << 18)
#define TG_TAGMASK (3 << 18)
#define TG_IC1     (1 << 9)
#define TG_MAX_M   400
#define TG_MAX_N   511
#define TG_PAY     0x3FFFF               /* the 18 payload bits below the tag */

// (v & keep) | tag in ONE LOP3: with the two masks as literals the compiler emits an AND and an OR (an instruction has
// room for one immediate), so they are handed over in registers it cannot see through
__device__ __forceinline__ int tg_retag(int v, int keep, int tag)
{
    int r; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(v), "r"(keep), "r"(tag)); return r;
}

// byte permute with the sign-replicating selector bit (PTX prmt default mode: selector nibble bit 3 set = fill the result byte with
// the sign of the selected byte).  __byte_perm cannot be used: it masks the selector with 0x7777.
__device__ __forceinline__ int tg_prmt(uint32_t lo, uint32_t hi, uint32_t sel)
{
    int r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel)); return r;
}

// a + b on the FMA pipe (IMAD with a multiplier the compiler cannot see through).  The cell update is all integer ALU work
// (LOP3, VIADDMNMX, VIMNMX3, PRMT), and that pipe issues one warp instruction every two cycles per scheduler; the plain adds
// that feed the add-max instructions are the part that can move to the other pipe.
__device__ __forceinline__ int tg_add(int a, int b, int one)
{
    int r; asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b)); return r;
}

struct TgConst {
    int go, ge, S0, K_ee, K_eo, K_fe, K_mfo, K_efo, K_fo, KEEP, T1, T2, KHI, KDG, P_I1, ONE, SH20, NOT1, C512;
};

template <int PAY>
__device__ __forceinline__ TgConst tg_consts(int go, int ge)
{
    TgConst k;
    const int gog = go + ge;
    k.go = go; k.ge = ge;
    // sentinel for "no gap open yet": below every reachable score, and one more extension must not wrap the 12-bit field
    k.S0 = -2047 + ge;
    // first payload: an insert column more per E step.  Second payload: an E step is one more trailing insert op (an opening
    // resets the payload to "1 I" before the constant is added)
    // Third payload (paths from (0,0) only): nothing to count in a gap
    k.K_ee = (int)((unsigned)(-ge) << TG_SC_SH) + TG_TAG1 + (PAY == 2 ? 0 : TG_IC1);    // E stored with tag 0: extend -> 1
    k.K_eo = (int)((unsigned)(-gog) << TG_SC_SH) - TG_TAG2 + (PAY == 0 ? TG_IC1 : 0);   // H stored with tag 2: open   -> 0
    // The deletion gap F[i][j+1] = max(F[i][j] - ge, H[i][j] - gog) is evaluated as max(F - ge, mm - gog, e - gog): opening
    // from an H that itself came out of F is never better than extending that F (and carries the same counters), so H drops
    // out of the recurrence and the chain along a row is one add-max and one retag per cell.  Order on equal scores:
    // extend 2 > open from a match 1 > open from an insertion 0, the traceback's.
    k.K_fe  = (int)((unsigned)(-ge) << TG_SC_SH) + TG_TAG1;                             // F stored with tag 1 -> 2
    k.K_mfo = (int)((unsigned)(-gog) << TG_SC_SH) - TG_TAG1;                            // mm carries tag 2 -> 1
    k.K_efo = (int)((unsigned)(-gog) << TG_SC_SH);                                      // e carries tag 0
    k.K_fo  = (int)((unsigned)(-gog) << TG_SC_SH) - TG_TAG2;                            // first column of a lane: open from the H handed over
    asm volatile("mov.b32 %0, %4; mov.b32 %1, %5; mov.b32 %2, %6; mov.b32 %3, %7;" : "=r"(k.KEEP), "=r"(k.T1), "=r"(k.T2), "=r"(k.ONE)
                 : "n"(~TG_TAGMASK), "n"(TG_TAG1), "n"(TG_TAG2), "n"(1));
    asm volatile("mov.b32 %0, %4; mov.b32 %1, %5; mov.b32 %2, %6; mov.b32 %3, %7;" : "=r"(k.KHI), "=r"(k.KDG), "=r"(k.P_I1), "=r"(k.SH20)
                 : "n"(~TG_PAY), "n"(~(0x1FF << 9)), "n"(TG_IC1), "n"(1 << TG_SC_SH));
    asm volatile("mov.b32 %0, %2; mov.b32 %1, %3;" : "=r"(k.NOT1), "=r"(k.C512) : "n"(~1), "n"(512));
    return k;
}

__device__ __forceinline__ uint32_t tg_subject_code(const AlignArgs2 &A, const AlignJob2 &J, int j)      // code of subject base j (1-based)
{
    if (J.b_off >= 0) return A.b[J.b_off + j - 1];
    const long long gp = J.gpos + j - 1;
    uint32_t code = (__ldg(A.seq + (gp >> 4)) >> (2 * (int)(gp & 15))) & 3u;
    if (A.n_nruns) {
        int lo = -1, hi = A.n_nruns;                              // last masked run with start <= gp
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (A.nruns[2 * mid] <= gp) lo = mid; else hi = mid; }
        if (lo >= 0 && gp < A.nruns[2 * lo + 1]) code = 4u;
    }
    return code;
}

// One sweep of the DP by one warp; leaves row m of H in Hst (lane l: columns l*C+1 .. l*C+C).  PAY selects the payload:
// 0 = start column / insert columns, 1 = trailing ops (above), 2 = the head of a path that starts at (0,0):
//     [17:9] leading diagonal steps, or the leading insert ops when bit 1 is set   [1] starts with inserts   [0] only diagonal steps so far
// from which the first CIGAR run of a path without leading deletions follows (the insert run, or the equal-kind stretch of the
// main diagonal cut at the number of leading diagonal steps).
template <int C, int PAY>
__device__ __forceinline__ void tg_sweep(const AlignArgs2 &A, const AlignJob2 &J, int lane, int (&Hst)[C])
{
    const unsigned FULL = 0xFFFFFFFFu;
    constexpr bool CHAIN_B = PAY == 1, CHAIN_C = PAY == 2;
    const TgConst k = tg_consts<PAY>(A.go, A.ge);
    const int go = k.go, ge = k.ge, S0 = k.S0;
    const int K_ee = k.K_ee, K_eo = k.K_eo, K_fe = k.K_fe, K_mfo = k.K_mfo, K_efo = k.K_efo, K_fo = k.K_fo;
    const int KEEP = k.KEEP, T1 = k.T1, T2 = k.T2, KHI = k.KHI, KDG = k.KDG, P_I1 = k.P_I1, ONE = k.ONE, SH20 = k.SH20, NOT1 = k.NOT1, C512 = k.C512;
    const int m = J.m, n = J.n;
    const uint8_t *a = A.a + J.a_off;
    const int j0 = lane * C;                                                  // this lane owns columns j0+1 .. j0+C
    int Est[C]; uint32_t colsel[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int j = j0 + c + 1;
        const uint32_t code = j <= n ? tg_subject_code(A, J, j) : 0u;
        // byte permute selector: byte 0 = table[code], bytes 1-3 = its sign  ->  a sign-extended score in one PRMT
        colsel[c] = code | ((code | 8u) << 4) | ((code | 8u) << 8) | ((code | 8u) << 12);
        Hst[c] = TG_TAG2 | (PAY == 0 ? j : 0);                                // row 0: free leading deletions, path = jD
        Est[c] = (int)((unsigned)S0 << TG_SC_SH);
    }
    // H[0][j0]: diagonal of my first column at row 1
    int prevIn = TG_TAG2 | (PAY == 0 ? j0 : (CHAIN_C && j0 == 0) ? 1 : 0);    // (third payload: the empty path at (0,0) is "only diagonal steps so far")
    int Hlast = 0, Flast = 0;
    for (int s = 1; s <= m + 31; s++) {
        const int i = s - lane;
        const int inH = __shfl_up_sync(FULL, Hlast, 1), inF = __shfl_up_sync(FULL, Flast, 1);
        if (i >= 1 && i <= m) {
            int left, F, diag = prevIn;
            if (lane == 0) {                                                  // column 0: H[i][0] = -(go + i ge), path = iI; no deletion gap yet
                left = (int)((unsigned)(-(go + i * ge)) << TG_SC_SH) | TG_TAG2 | (i << 9) | (CHAIN_C ? 2 : 0);     // i insert columns so far / trailing / leading insert ops
                F = (int)((unsigned)S0 << TG_SC_SH) | TG_TAG1;
            } else { left = inH; F = inF; }
            prevIn = left;
            F = tg_retag(max(F + K_fe, (CHAIN_B ? (left & KHI) : CHAIN_C ? (left & NOT1) : left) + K_fo), KEEP, T1);   // F[i][j0+1]
            const int ai = a[i - 1];
            // EDNAFULL row of this consensus symbol as signed bytes: vs A,C,G,T in tlo, vs N in thi
            const uint32_t tlo = ai < 4 ? ((0xFCFCFCFCu & ~(0xFFu << (8 * ai))) | (5u << (8 * ai))) : 0xFEFEFEFEu;
            const uint32_t thi = ai < 4 ? 0xFEu : 0xFFu;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int up = Hst[c];
                const int sub = tg_prmt(tlo, thi, colsel[c]);
                int e, mm;
                if (PAY == 0) {
                    e = max(Est[c] + K_ee, tg_add(up, K_eo, ONE)) & KEEP;
                    mm = tg_add(sub, diag, SH20);                                                          // diag + (sub << 20), one IMAD
                } else if (CHAIN_C) {
                    e = max(Est[c] + K_ee, tg_add(up & NOT1, K_eo, ONE)) & KEEP;                           // a gap ends the leading diagonal run
                    mm = tg_add(sub, tg_add(diag & 1, diag, C512), SH20);                                  // one more leading diagonal step while it lasts
                } else {
                    e = max(Est[c] + K_ee, tg_add(tg_retag(up, KHI, P_I1), K_eo, ONE)) & KEEP;             // opening: the path's last run is now 1 I
                    mm = tg_add(sub, (diag & KDG) + 1, SH20);                                              // one more trailing diagonal op, no trailing insert
                }
                const int h = tg_retag(__vimax3_s32(mm, F, e), KEEP, T2);
                diag = up; Hst[c] = h; Est[c] = e;
                if (c == C - 1) { Hlast = h; Flast = F; }
                else if (PAY == 0) F = tg_retag(max(F + K_fe, max(mm + K_mfo, tg_add(e, K_efo, ONE))), KEEP, T1);   // F[i][j+1]
                else F = tg_retag(max(F + K_fe, max(mm + K_mfo, tg_add(e, K_efo, ONE)) & (CHAIN_B ? KHI : NOT1)), KEEP, T1);   // (an opened deletion: payload 0 / run closed)
            }
        }
    }
}

// One warp per queue entry.  An entry is an alignment (first-payload sweep; when the last-row maximum is also attained at
// column n and no twin is queued, the same warp adds the second-payload sweep; when the path starts at column 0, the
// third-payload sweep) or the twin of an alignment the host expects to end at (m, n) (second-payload sweep only: its window
// sits at the edge of a run, where the consensus overhangs the slice) -- the two sweeps of such an alignment then run side by
// side on two warps instead of one after the other.  The host combines the result records (align_collect):
//   out[slot]              score | x and start-to-first-maximum columns | (x, I) of the path to (m, n), is n the first maximum
//   out[nslots + slot]     trailing diagonal ops, trailing insert ops of the path to (m, n), equal-kind stretch ending there
//   out[2 nslots + slot]   head words of the paths to the first maximum and to (m, n), equal-kind stretch from (1, 1)
template <int C>
__global__ void __launch_bounds__(128) kgma_align_tagged(AlignArgs2 A)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int ji = 0;
        if (lane == 0) ji = atomicAdd(A.next_job, 1);
        ji = __shfl_sync(FULL, ji, 0);
        if (ji >= A.njobs) break;
        const AlignJob2 J = A.jobs[ji];
        const int m = J.m, n = J.n, j0 = lane * C;
        const int ln = (n - 1) / C;                                           // lane that owns column n
        const uint8_t *a = A.a + J.a_off;
        bool second = (J.mode & 2) != 0, third = (J.mode & 8) != 0;         // twins: one sweep only
        const bool twin = second || third;
        int jstar = 0;
        // last row of a sweep: first column attaining the maximum (ties -> lowest column; column 0, the all-insert path, included)
        auto first_max = [&](const int (&Hst)[C], int w0, int &bv, int &w, int &wlast) {
            int best = -(A.go + m * A.ge), bestw = w0, bestj = 0, wn = 0;               // H[m][0]: m inserts, payload w0
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int j = j0 + c + 1;
                const int v = Hst[c] >> TG_SC_SH;
                if (j <= n && v > best) { best = v; bestw = Hst[c]; bestj = j; }
                if (j == n) wn = Hst[c];
            }
            // warp argmax over (score, lowest column): columns grow with the lane index, so on equal scores the lower lane wins
            bv = best; int bl = lane;
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bv, d), ol = __shfl_xor_sync(FULL, bl, d);
                if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
            }
            w = __shfl_sync(FULL, bestw, bl); jstar = __shfl_sync(FULL, bestj, bl);
            wlast = __shfl_sync(FULL, wn, ln);
        };
        if (!twin) {
            int w, bv, wlast;
            {
                int Hst[C];
                tg_sweep<C, 0>(A, J, lane, Hst);
                first_max(Hst, m << 9, bv, w, wlast);
            }
            const bool tie = (wlast >> TG_SC_SH) == bv;                       // the path may end at (m, n): the second payload decides
            const int x = w & 0x1FF, ic = (w >> 9) & 0x1FF;
            const bool head = x == 0 || (tie && (wlast & 0x1FF) == 0);        // a path from column 0 may be the one: the third payload
            const bool off = A.tail_mode == 2;
            if (lane == 0) {
                AlignOut o;
                o.score = bv; o.lower = x; o.num_sum = jstar + ic;
                o.nops = (off && (tie || head)) ? -1 : tie ? 3 : 2;
                o.cig_n = (wlast & 0x3FFFF) | ((jstar == n) << 18);            // x and I of the path to (m, n); is column n the first maximum?
                A.out[J.slot] = o;
            }
            second = tie && !off && !(J.mode & 1);
            third = head && !off && !(J.mode & 4);
        }
        if (second) {
            int w2;
            {
                int Hs2[C];
                tg_sweep<C, 1>(A, J, lane, Hs2);
                int wn2 = 0;
#pragma unroll
                for (int c = 0; c < C; c++) if (j0 + c + 1 == n) wn2 = Hs2[c];
                w2 = __shfl_sync(FULL, wn2, ln);
            }
            // length of the equal-kind stretch ending at (m, n) on its diagonal: '=' where the symbols are the same, 'X' elsewhere
            const int ndiag = w2 & 0x1FF;
            const int lim = min(ndiag, min(m, n));
            const bool k0 = tg_subject_code(A, J, n) == (uint32_t)a[m - 1];
            int r = lim;
            for (int t0 = 0; t0 < lim; t0 += 32) {                            // 32 diagonal steps per round
                const int t = t0 + lane;
                const unsigned diff = __ballot_sync(FULL, t < lim && (tg_subject_code(A, J, n - t) == (uint32_t)a[m - t - 1]) != k0);
                if (diff) { r = t0 + __ffs((int)diff) - 1; break; }
            }
            if (lane == 0) {
                AlignOut o;
                o.score = 0; o.lower = ndiag; o.num_sum = (w2 >> 9) & 0x1FF; o.nops = r; o.cig_n = 1;
                A.out[A.nslots + J.slot] = o;
            }
        }
        if (third) {
            int w3, w3n, bv3;
            {
                int Hs3[C];
                tg_sweep<C, 2>(A, J, lane, Hs3);
                first_max(Hs3, (m << 9) | 2, bv3, w3, w3n);                   // the same scores, hence the same first maximum
            }
            // length of the equal-kind stretch of the main diagonal from (1, 1), as far as either path follows it
            const int d0 = max((w3 & 2) ? 0 : (w3 >> 9) & 0x1FF, (w3n & 2) ? 0 : (w3n >> 9) & 0x1FF);
            const int lim = min(d0, min(m, n));
            int r = lim;
            if (lim > 0) {
                const bool k0 = tg_subject_code(A, J, 1) == (uint32_t)a[0];
                for (int t0 = 0; t0 < lim; t0 += 32) {
                    const int t = t0 + lane;
                    const unsigned diff = __ballot_sync(FULL, t < lim && (tg_subject_code(A, J, 1 + t) == (uint32_t)a[t]) != k0);
                    if (diff) { r = t0 + __ffs((int)diff) - 1; break; }
                }
            }
            if (lane == 0) {
                AlignOut o;
                o.score = 0; o.lower = w3 & 0x3FFFF; o.num_sum = w3n & 0x3FFFF; o.nops = r; o.cig_n = 1;
                A.out[2 * A.nslots + J.slot] = o;
            }
        }
        __syncwarp();
    }
}

static inline uint8_t sym_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
    case 'T': case 't': return 3; case 'N': case 'n': return 4; default: return 255;
    }
}

// Small uploads of the extension batches go through this kernel (reads of the page-locked staging block over the bus)
// rather than the copy engine: during a streamed scan the engine is busy with 32 MB genome chunks, and a 50 KB copy queued
// behind one of those waits ~0.6 ms -- per round of the cluster-mode replay.
__global__ void kgma_fetch_host(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int slot_scratch(kgma_ctx *ctx, int slot, size_t dbytes, size_t hbytes, void **d, void **h)
{
    if (dbytes > ctx->a_dev_bytes[slot]) {
        if (ctx->a_dev[slot]) cudaFree(ctx->a_dev[slot]);
        ctx->a_dev[slot] = nullptr; ctx->a_dev_bytes[slot] = 0;
        const size_t nb = std::max(dbytes + dbytes / 4, (size_t)1 << 20);
        KGMA_CUDA(ctx, cudaMalloc(&ctx->a_dev[slot], nb));
        ctx->a_dev_bytes[slot] = nb;
    }
    if (hbytes > ctx->a_host_bytes[slot]) {
        if (ctx->a_host[slot]) cudaFreeHost(ctx->a_host[slot]);
        ctx->a_host[slot] = nullptr; ctx->a_host_bytes[slot] = 0;
        const size_t nb = std::max(hbytes + hbytes / 4, (size_t)1 << 20);
        KGMA_CUDA(ctx, cudaHostAlloc(&ctx->a_host[slot], nb, cudaHostAllocMapped));   // the kernels read / write it in place
        ctx->a_host_bytes[slot] = nb;
    }
    *d = ctx->a_dev[slot]; *h = ctx->a_host[slot];
    return KGMA_OK;
}

// Queue the trace-free extension of `reqs` on stream `st` (own scratch per slot, so it can run next to a scan that is
// still streaming); align_collect waits for it, redoes the few alignments the tagged kernel handed back, and converts.
int align_enqueue(kgma_ctx *ctx, kgma_genome *g, const std::vector<AlignReq> &reqs, const kgma_profile *profiles, int n_profiles,
                  bool single_mode_truncate, int gap_open, int gap_extend, bool tie_open, cudaStream_t st, int slot, AlignTicket *t)
{
    *t = AlignTicket{};
    if (reqs.empty()) return KGMA_OK;
    const bool trace = getenv("KGMA_TRACE") != nullptr;
    const double te0 = trace ? now_ms() : 0;
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    // (scratch kept per thread: the queue of a cfg2 batch is 160 KB, which malloc would serve with a fresh mmap per call)
    static thread_local std::vector<uint8_t> acodes, bcodes; static thread_local std::vector<AlignJob2> jobs, q;
    acodes.clear(); bcodes.clear();
    std::vector<int32_t> a_off(n_profiles), a_len(n_profiles);
    for (int q = 0; q < n_profiles; q++) {
        const kgma_profile &p = profiles[q];
        if (!p.consensus) return set_err(ctx, KGMA_E_ARG, "profile %d has no consensus sequence to align against", q);
        int len = p.consensus_len;
        if (single_mode_truncate) {                      // Alignment.jl:42 view(consensus_seq, 1:windowsize)
            if (len < p.window) return set_err(ctx, KGMA_E_ARG, "BoundsError: consensus (%d) shorter than the window (%lld)", len, (long long)p.window);
            len = (int)p.window;
        }
        a_off[q] = (int32_t)acodes.size(); a_len[q] = len;
        for (int i = 0; i < len; i++) {
            uint8_t c = sym_code(p.consensus[i]);
            if (c == 255) return set_err(ctx, KGMA_E_SYMBOL, "consensus of profile %d holds a symbol outside A,C,G,T,N", q);
            acodes.push_back(c);
        }
    }
    const std::vector<int64_t> &nruns = genome_nruns(g);
    const bool on_dev = ctx->dg_uid == g->uid && ctx->d_seq2 && ctx->d_have_hi > ctx->d_have_lo;
    jobs.resize(reqs.size()); int maxn = 0;
    for (size_t i = 0; i < reqs.size(); i++) {
        const AlignReq &rq = reqs[i];
        if (rq.record < 0 || rq.record >= (int)g->recs.size() || rq.profile < 0 || rq.profile >= n_profiles)
            return set_err(ctx, KGMA_E_ARG, "alignment request %zu out of range", i);
        const kgma::Record &R = g->recs[rq.record];
        if (rq.first < 1 || rq.last > R.len || rq.last < rq.first) return set_err(ctx, KGMA_E_ARG, "alignment range %lld:%lld invalid", (long long)rq.first, (long long)rq.last);
        AlignJob2 &J = jobs[i];
        J.gpos = R.off + rq.first - 1; J.n = (int)(rq.last - rq.first + 1); J.a_off = a_off[rq.profile]; J.m = a_len[rq.profile]; J.b_off = -1;
        J.mode = 0; J.slot = (int32_t)i;
        maxn = std::max(maxn, J.n);
        if (J.m + J.n >= 4095) return set_err(ctx, KGMA_E_UNSUPPORTED, "alignment of %d x %d too long for the extension kernel (12-bit gap lengths)", J.m, J.n);
        if (!(on_dev && J.gpos >= ctx->d_have_lo && J.gpos + J.n + 16 <= ctx->d_have_hi)) {      // slice lives on another shard's device: ship codes
            J.b_off = (int32_t)bcodes.size();
            for (int64_t p = rq.first; p <= rq.last; p++) {
                int64_t gp = R.off + p - 1;
                uint8_t c = (uint8_t)base_code(g, gp);
                if (base_masked(g, gp)) { if (c != 3) return set_err(ctx, KGMA_E_SYMBOL, "subject holds a symbol outside A,C,G,T,N"); c = 4; }
                bcodes.push_back(c);
            }
        }
    }
    if (g->ambiguous) return set_err(ctx, KGMA_E_SYMBOL, "subject holds a symbol outside A,C,G,T,N");
    const int nj = (int)jobs.size();                           // requests = result slots
    const int ncol = (maxn + 1 + 31) & ~31;
    const int warps_per_block = 4;
    constexpr int ROWS = 10;                                   // kgma_align_summary: DP rows per lane, one sweep covers 320 consensus rows
    int maxm = 0; for (int q = 0; q < n_profiles; q++) maxm = std::max(maxm, a_len[q]);
    const bool need_boundary = maxm > 32 * ROWS;
    const size_t per_warp = (((size_t)ncol * (need_boundary ? (6 * 4 + 1) : 1)) + 15) & ~(size_t)15;
    const size_t smem = (size_t)warps_per_block * per_warp;    // of the path-summary kernel (the tagged kernel uses none)
    if (smem > ctx->smem_optin) return set_err(ctx, KGMA_E_UNSUPPORTED, "subject slice of %d bases too long for the extension kernel", maxn);
    // The tagged kernel (one word per DP state, DPX three-way max) takes the batch when its 12-bit score field and 9-bit
    // counters hold every reachable value and the default tie rule is asked for; KGMA_ALIGN_KERNEL=summary forces the
    // path-summary kernel (tests compare the two).
    const int go = -gap_open, ge = -gap_extend;
    const char *kern_env = getenv("KGMA_ALIGN_KERNEL");
    const bool tagged = !(kern_env && !strcmp(kern_env, "summary")) && !tie_open && go >= 0 && ge >= 0 && maxm >= 1 && maxm <= TG_MAX_M && maxn <= TG_MAX_N &&
                        5 * maxm < 2040 && 2LL * go + (long long)(maxm + 3) * ge + 8 < 2040;
    // KGMA_ALIGN_TAIL = all | off: a second-payload twin for every alignment / no second sweep at all (tests compare the routes)
    int tail_mode = 0;
    { const char *tv = getenv("KGMA_ALIGN_TAIL"); tail_mode = tv && !strcmp(tv, "all") ? 1 : tv && !strcmp(tv, "off") ? 2 : 0; }
    // The queue: twins first (second-payload twins are the longest entries, and the host waits for their results), then the
    // alignments that have twins, then the rest.  A twin costs a sweep of its own, so who gets one depends on how full the
    // machine is: with room for three warps per alignment (small batches: a shard of a multi-GPU scan, the tail batch of a
    // streamed one) every alignment gets both twins and the batch takes the time of ONE sweep; otherwise the marked
    // alignments get a second-payload twin, and the third sweep stays with the warp that finds it necessary.
    if (tagged && tail_mode != 2) {
        const size_t room = (size_t)ctx->num_sms * 12;                       // resident warps of the tagged kernel (3 CTAs per SM)
        size_t n_marked = 0;
        for (const AlignReq &rq : reqs) n_marked += rq.hint & 1;
        const bool all3 = 3 * jobs.size() <= room;
        const bool marked3 = !all3 && jobs.size() + 2 * n_marked <= room;
        q.clear(); q.reserve(jobs.size() * 3);
        for (int pass = 0; pass < 4; pass++)
            for (size_t i = 0; i < jobs.size(); i++) {
                const bool marked = tail_mode == 1 || all3 || (reqs[i].hint & 1);
                const bool with3 = marked && (all3 || marked3);
                AlignJob2 t2 = jobs[i];
                if (pass == 0 && marked) { t2.mode = 2; q.push_back(t2); }
                if (pass == 1 && with3) { t2.mode = 8; q.push_back(t2); }
                if (pass == 2 && marked) { t2.mode = 1 | (with3 ? 4 : 0); q.push_back(t2); }
                if (pass == 3 && !marked) q.push_back(t2);
            }
        jobs.swap(q);
    }
    const int nq = (int)jobs.size();                           // queue entries
    const int nout = tagged ? 3 * nj : nj;                     // result records: [nj] first, [nj] second, [nj] third sweep
    size_t o = 0;
    auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    const size_t o_a = carve(acodes.size()), o_b = carve(bcodes.size()), o_j = carve((size_t)nq * sizeof(AlignJob2));
    const size_t o_n = carve(nruns.size() * 8 + 8);
    const size_t up = o;
    const size_t o_c = carve(256), o_o = carve((size_t)nout * sizeof(AlignOut));
    const size_t o_j2 = carve(tagged ? (size_t)nj * sizeof(AlignJob2) : 0), o_o2 = carve(tagged ? (size_t)nj * sizeof(AlignOut) : 0);   // second pass
    void *dv = nullptr, *hv = nullptr;
    int rc = slot_scratch(ctx, slot, o, up + (size_t)nout * sizeof(AlignOut) + (tagged ? (size_t)nj * (sizeof(AlignJob2) + sizeof(AlignOut)) : 0), &dv, &hv);
    if (rc) return rc;
    unsigned char *d = (unsigned char *)dv, *h = (unsigned char *)hv;
    memcpy(h + o_a, acodes.data(), acodes.size());
    if (!bcodes.empty()) memcpy(h + o_b, bcodes.data(), bcodes.size());
    memcpy(h + o_j, jobs.data(), (size_t)nq * sizeof(AlignJob2));
    if (!nruns.empty()) memcpy(h + o_n, nruns.data(), nruns.size() * 8);
    // Two ways to move the batch.  Next to a streaming scan (the pipelined parts run on s_align) the copy engine is busy with
    // 32 MB genome chunks and a small copy queued behind one waits ~0.6 ms, so the jobs are fetched by a kernel and the
    // results written in place into the mapped block.  Otherwise plain asynchronous copies are at least as fast.
    const bool in_place = st == ctx->s_align;
    void *h_dev = nullptr;                                     // device-side address of the staging block (same as h under UVA)
    if (in_place) KGMA_CUDA(ctx, cudaHostGetDevicePointer(&h_dev, h, 0));
    if (!in_place) KGMA_CUDA(ctx, cudaMemcpyAsync(d, h, up, cudaMemcpyHostToDevice, st));
    else {
        const size_t n16 = up / 16;                            // (carve rounds every piece to 256 bytes)
        const int cgrid = (int)std::min<size_t>((n16 + 255) / 256, (size_t)ctx->num_sms * 4);
        kgma_fetch_host<<<std::max(cgrid, 1), 256, 0, st>>>((uint4 *)d, (const uint4 *)h_dev, n16);
        KGMA_CUDA(ctx, cudaGetLastError());
        ctx->stats.launches++;
    }
    KGMA_CUDA(ctx, cudaMemsetAsync(d + o_c, 0, 256, st));
    ctx->stats.h2d_bytes += up;
    AlignArgs2 A{};
    A.a = d + o_a; A.seq = ctx->d_seq2; A.nruns = (const long long *)(d + o_n); A.n_nruns = (int)(nruns.size() / 2);
    A.b = d + o_b; A.jobs = (const AlignJob2 *)(d + o_j); A.njobs = nq; A.nslots = nj; A.next_job = (int *)(d + o_c);
    // results are written straight into the page-locked block (16 bytes per alignment, posted writes): no copy back
    A.out = in_place ? (AlignOut *)((unsigned char *)h_dev + up) : (AlignOut *)(d + o_o); A.go = go; A.ge = ge; A.tie_open = tie_open ? 1 : 0; A.ncol_cap = ncol;
    A.need_boundary = need_boundary ? 1 : 0;
    A.tail_mode = tail_mode;
    KGMA_CUDA(ctx, cudaFuncSetAttribute(kgma_align_summary<ROWS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KGMA_CUDA(ctx, cudaFuncSetAttribute(kgma_align_summary<ROWS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // tagged kernel: 3 CTAs (12 warps) per SM fetch entries from the queue -- with every entry resident at once (3.4 per scheduler
    // for the 2 022 entries of cfg2) the schedulers that drew four finish last; measured 0.257 ms against 0.280
    int ctas_per_sm = tagged ? 3 : 8;
    if (const char *e = getenv("KGMA_ALIGN_CTAS")) ctas_per_sm = std::max(1, std::min(8, atoi(e)));
    const int grid = std::min((nq + warps_per_block - 1) / warps_per_block, ctx->num_sms * ctas_per_sm);
    if (tagged && in_place) memset(h + up, 0, (size_t)nout * sizeof(AlignOut));
    else if (tagged) KGMA_CUDA(ctx, cudaMemsetAsync(d + o_o, 0, (size_t)nout * sizeof(AlignOut), st));   // (cig_n = 0: no second-sweep record)
    KGMA_CUDA(ctx, cudaEventRecord(ctx->a_ev0[slot], st));
    if (tagged) {
        if (maxn <= 13 * 32) kgma_align_tagged<13><<<grid, warps_per_block * 32, 0, st>>>(A);
        else kgma_align_tagged<16><<<grid, warps_per_block * 32, 0, st>>>(A);
    }
    else if (tie_open) kgma_align_summary<ROWS, true><<<grid, warps_per_block * 32, smem, st>>>(A);
    else kgma_align_summary<ROWS, false><<<grid, warps_per_block * 32, smem, st>>>(A);
    KGMA_CUDA(ctx, cudaGetLastError());
    KGMA_CUDA(ctx, cudaEventRecord(ctx->a_ev1[slot], st));
    ctx->stats.launches++;
    AlignOut *ho = (AlignOut *)(h + up);
    if (!in_place) KGMA_CUDA(ctx, cudaMemcpyAsync(ho, d + o_o, (size_t)nout * sizeof(AlignOut), cudaMemcpyDeviceToHost, st));
    KGMA_CUDA(ctx, cudaEventRecord(ctx->a_done[slot], st));
    ctx->stats.d2h_bytes += (size_t)nout * sizeof(AlignOut);
    t->active = true; t->slot = slot; t->nj = nj; t->ho = ho;
    t->tagged = tagged; t->st = st; t->args = A; t->smem = smem; t->maxn = maxn; t->nq = nq;
    t->d_jobs2 = d + o_j2; t->d_out2 = d + o_o2; t->d_cnt = d + o_c;
    t->h_jobs = h + o_j; t->h_jobs2 = h + up + (size_t)nout * sizeof(AlignOut); t->h_out2 = (unsigned char *)t->h_jobs2 + (size_t)nj * sizeof(AlignJob2);
    if (trace) fprintf(stderr, "[kgma align] batch of %d queued in %.3f ms of host time (%s kernel)\n", nj, now_ms() - te0, tagged ? "tagged" : "summary");
    return KGMA_OK;
}

int align_collect(kgma_ctx *ctx, AlignTicket *t, std::vector<AlignRes> &out)
{
    out.assign((size_t)t->nj, AlignRes{ 1, 0, 0, 0, 0 });
    if (!t->active) return KGMA_OK;
    KGMA_CUDA(ctx, cudaEventSynchronize(ctx->a_done[t->slot]));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->a_ev0[t->slot], ctx->a_ev1[t->slot]);
    ctx->stats.align_ms += ms;
    AlignOut *ho = (AlignOut *)t->ho;
    t->active = false;
    if (t->tagged) {
        const AlignJob2 *hj = (const AlignJob2 *)t->h_jobs;           // in queue order; .slot = index of the request
        AlignJob2 *hj2 = (AlignJob2 *)t->h_jobs2;
        std::vector<int> redo;
        for (int j = 0; j < t->nq; j++) {
            const AlignJob2 &J = hj[j];
            if (J.mode & (2 | 8)) continue;                           // a twin: its record is read below
            AlignOut &A1 = ho[J.slot];
            if (A1.nops >= 2) {
                // combine the sweeps' records (kgma_align_tagged)
                const int x_n = A1.cig_n & 0x1FF, i_n = (A1.cig_n >> 9) & 0x1FF; const bool first_max_at_n = (A1.cig_n >> 18) & 1;
                bool at_end = false; int lastrun = 0;
                if (A1.nops == 3) {                                   // the last-row maximum is also attained at column n
                    const AlignOut &B = ho[t->nj + J.slot];
                    if (B.cig_n != 1) return set_err(ctx, KGMA_E_STATE, "extension: second-sweep record of alignment %d missing", J.slot);
                    const bool diag_last = B.lower > 0;               // the path to (m, n) ends with a diagonal step: match before free deletion
                    at_end = first_max_at_n || diag_last;
                    lastrun = diag_last ? std::min(B.lower, B.nops) : B.num_sum;
                    ctx->stats.n_align_redo++;
                }
                int lower = at_end ? x_n : A1.lower, num_sum = at_end ? J.n + i_n - lastrun : A1.num_sum;
                if (lower == 0) {                                     // no leading deletions: the first run comes from the head payload
                    const AlignOut &Cc = ho[2 * t->nj + J.slot];
                    if (Cc.cig_n != 1) return set_err(ctx, KGMA_E_STATE, "extension: third-sweep record of alignment %d missing", J.slot);
                    const int w3 = at_end ? Cc.num_sum : Cc.lower, d0 = (w3 >> 9) & 0x1FF;
                    lower = (w3 & 2) ? d0 : std::min(d0, Cc.nops);
                    if (at_end && lower == J.n + i_n) { lower = 0; num_sum = 0; }    // the whole CIGAR is one run: (0 + 1):0
                    ctx->stats.n_align_head++;
                }
                A1.lower = lower; A1.num_sum = num_sum; A1.nops = 2;
            }
            if (A1.nops < 0) { hj2[redo.size()] = J; redo.push_back(J.slot); }
        }
        // second pass: the alignments the tagged kernel cannot finish (paths that start at column 0 of the slice: hits at a
        // record edge, or buff = 0) go through the path-summary kernel
        if (!redo.empty()) {
            const int nf = (int)redo.size();
            cudaStream_t st = (cudaStream_t)t->st;
            AlignArgs2 A = t->args;
            KGMA_CUDA(ctx, cudaMemcpyAsync(t->d_jobs2, hj2, (size_t)nf * sizeof(AlignJob2), cudaMemcpyHostToDevice, st));
            KGMA_CUDA(ctx, cudaMemsetAsync(t->d_cnt, 0, 256, st));
            A.jobs = (const AlignJob2 *)t->d_jobs2; A.njobs = nf; A.out = (AlignOut *)t->d_out2; A.next_job = (int *)t->d_cnt;
            const int grid = std::min((nf + 3) / 4, ctx->num_sms * 8);
            KGMA_CUDA(ctx, cudaEventRecord(ctx->a_ev0[t->slot], st));
            kgma_align_summary<10, false><<<grid, 128, t->smem, st>>>(A);
            KGMA_CUDA(ctx, cudaGetLastError());
            KGMA_CUDA(ctx, cudaEventRecord(ctx->a_ev1[t->slot], st));
            KGMA_CUDA(ctx, cudaMemcpyAsync(t->h_out2, t->d_out2, (size_t)nf * sizeof(AlignOut), cudaMemcpyDeviceToHost, st));
            KGMA_CUDA(ctx, cudaStreamSynchronize(st));
            cudaEventElapsedTime(&ms, ctx->a_ev0[t->slot], ctx->a_ev1[t->slot]);
            ctx->stats.align_ms += ms;
            ctx->stats.launches++;
            ctx->stats.n_align_summary += nf;
            const AlignOut *ho2 = (const AlignOut *)t->h_out2;
            for (int i = 0; i < nf; i++) ho[redo[(size_t)i]] = ho2[i];
            if (getenv("KGMA_TRACE")) fprintf(stderr, "[kgma align] %d of %d alignments redone by the path-summary kernel\n", nf, t->nj);
        }
    }
    for (int q = 0; q < t->nj; q++) { out[(size_t)q].lo = (int64_t)ho[q].lower + 1; out[(size_t)q].hi = ho[q].num_sum; out[(size_t)q].score = ho[q].score; }
    return KGMA_OK;
}

int align_batch_device(kgma_ctx *ctx, kgma_genome *g, const std::vector<AlignReq> &reqs,
                       const kgma_profile *profiles, int n_profiles, bool single_mode_truncate,
                       int gap_open, int gap_extend, bool tie_open, bool want_cigars,
                       std::vector<AlignRes> &out, std::vector<char> *cig_ops, std::vector<int32_t> *cig_cnt)
{
    out.assign(reqs.size(), AlignRes{ 1, 0, 0, 0, 0 });
    if (reqs.empty()) return KGMA_OK;
    KGMA_CUDA(ctx, cudaSetDevice(ctx->device));
    // consensus codes per profile
    std::vector<uint8_t> acodes; std::vector<int32_t> a_off(n_profiles), a_len(n_profiles);
    for (int q = 0; q < n_profiles; q++) {
        const kgma_profile &p = profiles[q];
        if (!p.consensus) return set_err(ctx, KGMA_E_ARG, "profile %d has no consensus sequence to align against", q);
        int len = p.consensus_len;
        if (single_mode_truncate) {                      // Alignment.jl:42 view(consensus_seq, 1:windowsize)
            if (len < p.window) return set_err(ctx, KGMA_E_ARG, "BoundsError: consensus (%d) shorter than the window (%lld)", len, (long long)p.window);
            len = (int)p.window;
        }
        a_off[q] = (int32_t)acodes.size(); a_len[q] = len;
        for (int i = 0; i < len; i++) {
            uint8_t c = sym_code(p.consensus[i]);
            if (c == 255) return set_err(ctx, KGMA_E_SYMBOL, "consensus of profile %d holds a symbol outside A,C,G,T,N", q);
            acodes.push_back(c);
        }
    }
    cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[1];
    double align_ms = 0;
    cudaStream_t st = ctx->s_extend ? ctx->s_extend : ctx->s_compute;
    if (!want_cigars) {
        // ---- trace-free path: one launch for everything; subjects come from the packed genome already on the device
        AlignTicket t;
        int rc = align_enqueue(ctx, g, reqs, profiles, n_profiles, single_mode_truncate, gap_open, gap_extend, tie_open, st, 0, &t);
        if (rc) return rc;
        return align_collect(ctx, &t, out);
    }
    const size_t TRACE_BUDGET = (size_t)768 << 20;
    size_t done = 0;
    while (done < reqs.size()) {
        // ---- carve a batch that fits the trace budget
        std::vector<AlignJob> jobs; std::vector<uint8_t> bcodes;
        size_t tr_bytes = 0, cig_entries = 0; int maxn = 0;
        size_t i = done;
        for (; i < reqs.size(); i++) {
            const AlignReq &rq = reqs[i];
            if (rq.record < 0 || rq.record >= (int)g->recs.size() || rq.profile < 0 || rq.profile >= n_profiles)
                return set_err(ctx, KGMA_E_ARG, "alignment request %zu out of range", i);
            const kgma::Record &R = g->recs[rq.record];
            if (rq.first < 1 || rq.last > R.len || rq.last < rq.first) return set_err(ctx, KGMA_E_ARG, "alignment range %lld:%lld invalid", (long long)rq.first, (long long)rq.last);
            int n = (int)(rq.last - rq.first + 1), m = a_len[rq.profile];
            size_t need = (size_t)(m + 1) * (size_t)((n + 1 + 3) & ~3);
            need = (need + 15) & ~(size_t)15;
            if (!jobs.empty() && tr_bytes + need > TRACE_BUDGET) break;
            AlignJob J{};
            J.tr_off = (int64_t)tr_bytes; J.a_off = a_off[rq.profile]; J.m = m;
            J.b_off = (int32_t)bcodes.size(); J.n = n; J.cig_off = (int32_t)cig_entries;
            tr_bytes += need; cig_entries += (size_t)(m + n + 2); maxn = std::max(maxn, n);
            for (int64_t p = rq.first; p <= rq.last; p++) {
                int64_t gp = R.off + p - 1;
                uint8_t c = (uint8_t)base_code(g, gp);
                if (base_masked(g, gp)) { if (c != 3) return set_err(ctx, KGMA_E_SYMBOL, "subject holds a symbol outside A,C,G,T,N"); c = 4; }
                bcodes.push_back(c);
            }
            jobs.push_back(J);
        }
        const int nj = (int)jobs.size();                           // requests = result slots
        const int ncol = (maxn + 1 + 31) & ~31;
        const int warps_per_block = 4;
        size_t smem = (size_t)warps_per_block * 2 * ncol * sizeof(int);
        if (smem > ctx->smem_optin) return set_err(ctx, KGMA_E_UNSUPPORTED, "subject slice of %d bases too long for the extension kernel", maxn);
        // ---- device buffers
        size_t o = 0;
        auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
        size_t o_a = carve(acodes.size()), o_b = carve(bcodes.size()), o_j = carve((size_t)nj * sizeof(AlignJob));
        size_t o_o = carve((size_t)nj * sizeof(AlignOut)), o_c = carve(256);
        size_t o_cg = carve(want_cigars ? cig_entries * 4 : 0), o_tr = carve(tr_bytes);
        void *dv = nullptr;
        int rc = dev_scratch(ctx, o, &dv);
        ctx->tab_sig = 0;                                  // (the arena is carved anew: the scan's resident tables are gone)
        if (rc) return rc;
        unsigned char *d = (unsigned char *)dv;
        KGMA_CUDA(ctx, cudaMemcpyAsync(d + o_a, acodes.data(), acodes.size(), cudaMemcpyHostToDevice, st));
        KGMA_CUDA(ctx, cudaMemcpyAsync(d + o_b, bcodes.data(), bcodes.size(), cudaMemcpyHostToDevice, st));
        KGMA_CUDA(ctx, cudaMemcpyAsync(d + o_j, jobs.data(), (size_t)nj * sizeof(AlignJob), cudaMemcpyHostToDevice, st));
        KGMA_CUDA(ctx, cudaMemsetAsync(d + o_c, 0, 256, st));
        ctx->stats.h2d_bytes += acodes.size() + bcodes.size() + (size_t)nj * sizeof(AlignJob);
        AlignArgs A{};
        A.a = d + o_a; A.b = d + o_b; A.jobs = (const AlignJob *)(d + o_j); A.njobs = nj; A.next_job = (int *)(d + o_c);
        A.trace = d + o_tr; A.cigar = want_cigars ? (uint32_t *)(d + o_cg) : nullptr; A.out = (AlignOut *)(d + o_o);
        A.go = -gap_open; A.ge = -gap_extend; A.tie_open = tie_open ? 1 : 0; A.ncol_cap = ncol;
        KGMA_CUDA(ctx, cudaFuncSetAttribute(kgma_align, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = std::min((nj + warps_per_block - 1) / warps_per_block, ctx->num_sms * 4);
        KGMA_CUDA(ctx, cudaEventRecord(e0, st));
        kgma_align<<<grid, warps_per_block * 32, smem, st>>>(A);
        KGMA_CUDA(ctx, cudaGetLastError());
        KGMA_CUDA(ctx, cudaEventRecord(e1, st));
        ctx->stats.launches++;
        std::vector<AlignOut> ho(nj);
        KGMA_CUDA(ctx, cudaMemcpyAsync(ho.data(), d + o_o, (size_t)nj * sizeof(AlignOut), cudaMemcpyDeviceToHost, st));
        std::vector<uint32_t> hc;
        if (want_cigars) { hc.resize(cig_entries); KGMA_CUDA(ctx, cudaMemcpyAsync(hc.data(), d + o_cg, cig_entries * 4, cudaMemcpyDeviceToHost, st)); }
        KGMA_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += (size_t)nj * sizeof(AlignOut) + (want_cigars ? cig_entries * 4 : 0);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); align_ms += ms;
        for (int q = 0; q < nj; q++) {
            AlignRes &r = out[done + q];
            r.lo = (int64_t)ho[q].lower + 1; r.hi = ho[q].num_sum; r.score = ho[q].score;
            if (want_cigars && cig_ops && cig_cnt) {
                r.cig_off = (uint32_t)cig_ops->size(); r.cig_len = (uint32_t)ho[q].cig_n;
                for (int t = ho[q].cig_n - 1; t >= 0; t--) {         // device wrote them end-to-start
                    uint32_t v = hc[(size_t)jobs[q].cig_off + t];
                    cig_ops->push_back((char)(v & 0xFF)); cig_cnt->push_back((int32_t)(v >> 8));
                }
            }
        }
        done = i;
    }
    ctx->stats.align_ms += align_ms;
    return KGMA_OK;
}

}  // namespace kgma

using namespace kgma;

extern "C" int kgma_align_batch(kgma_ctx *ctx, kgma_genome *g, const char *consensus, int32_t cons_len,
                                int32_t gap_open, int32_t gap_extend, uint32_t flags, int64_t n,
                                const int32_t *record, const int64_t *first, const int64_t *last,
                                int64_t *out_first, int64_t *out_last, int64_t *out_score)
{
    if (!ctx || !g || !consensus || n < 0 || (n && (!record || !first || !last || !out_first || !out_last))) return KGMA_E_ARG;
    kgma_profile p{};
    p.k = 1; p.n_refs = 1; p.window = cons_len; p.S = nullptr; p.consensus = consensus; p.consensus_len = cons_len; p.thr = 0;
    std::vector<AlignReq> reqs((size_t)n);
    for (int64_t i = 0; i < n; i++) reqs[(size_t)i] = { record[i], 0, first[i], last[i] };
    std::vector<AlignRes> res;
    ctx->stats = kgma_stats{};
    int rc = align_batch_device(ctx, g, reqs, &p, 1, false, gap_open, gap_extend, (flags & KGMA_F_TIE_OPEN) != 0, false, res, nullptr, nullptr);
    if (rc) return rc;
    for (int64_t i = 0; i < n; i++) {
        int64_t L = g->recs[record[i]].len;
        int64_t f = std::max<int64_t>(1, first[i] + res[(size_t)i].lo - 1), l = std::min<int64_t>(first[i] + res[(size_t)i].hi - 1, L);
        if (l < f - 1) l = f - 1;
        out_first[i] = f; out_last[i] = l;
        if (out_score) out_score[i] = res[(size_t)i].score;
    }
    return KGMA_OK;
}
