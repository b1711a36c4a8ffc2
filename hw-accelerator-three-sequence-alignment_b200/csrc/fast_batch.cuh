// fast_batch.cuh -- K1: the batched kernel for many short triples (len <= 64), as one CTA-wide systolic array.
//
// Mapping (the GPU counterpart of the RTL's P x P PE array, src/TriAlign_1cyc.v:116-125, with P = 16 threads x 4
// cells per side and the whole 64x64 (j,k) plane resident, so no slicing / face SRAM is needed at this length):
//   * one CTA = TJ x TK threads; thread (tj,tk) owns the RJ x RK cells j in (tj*RJ, tj*RJ+RJ], k likewise, for ALL i;
//   * the plane-(i-1) state of those cells (all, wX, u1, u3 -- see fast_core.cuh) lives in registers;
//   * thread (tj,tk) works on plane i at global step s = i + LAG*(tj + tk) (wavefront skew, like PE(j,k) at step
//     i+j+k; LAG: see BatchCfg);
//   * neighbours exchange tile edges through shared memory (2 * LAG edge slots), one CTA barrier per step
//     (the counterpart of the PE-to-PE wires; the RTL's _d1/_d2 delay registers are the "old" values published
//     one step late -- measured: keeping them in the consumer's registers instead, so that half of the halo loads
//     leave the critical path, is slower, because it moves all stores to the end of the step);
//   * two triples ride in the two 16-bit halves of every word; consecutive pairs of triples stream through the array
//     back to back, so the (TJ+TK-2)-step fill/drain is paid once per CTA, not once per triple.
// Per-pair score tables (a vs b, a vs c for all 16 combinations of the two a symbols) are rebuilt in shared memory
// while the previous pair is still in flight.
#pragma once

#include "fast_core.cuh"

namespace trialign {
namespace fast {

// LAG = steps between a thread and its top / left neighbour.  LAG 1: edges published at step s-1 are consumed at
// step s, so every step ends with a CTA barrier that everybody waits for.  LAG 2: edges are consumed TWO steps after
// they were published, so a thread only has to know that everybody finished step s-2 before it starts step s; the
// barrier of step s-1 is never waited for directly and warps drift up to one step apart, which spreads the
// shared-memory bursts at the start / end of a step over the other warps' tiles.
template <int RJ_, int RK_, int TJ_, int TK_, int LAG_ = 1>
struct BatchCfg {
    static constexpr int RJ = RJ_, RK = RK_, TJ = TJ_, TK = TK_, LAG = LAG_;
    static constexpr int NSLOT = 2 * LAG;        // edge buffers: written at step s, read at step s+LAG
    static constexpr int NBUF = LAG + 1;         // per-pair sequence / table buffers
    static constexpr int BJ = RJ * TJ, BK = RK * TK, NT = TJ * TK;
    static constexpr int L = 64;                 // planes per pair (sequences are padded to 64)
    // shared-memory strides (32-bit words), chosen so that the 8 / 16 lanes of one access wavefront (see thread_pos)
    // fall into distinct banks
    static constexpr int ROW_STRIDE = RK == 4 ? BK : BK + 16;   // 64-bit row accesses: two tj per wavefront
    static constexpr bool DIAG = TJ == 16 && TK == 16;          // lanes of an access group along an anti-diagonal, see thread_pos
    static constexpr int COL_STRIDE = DIAG ? BJ + 8 : BJ + 4;   // 128-bit column accesses: lanes differ in tk (and tj: DIAG)
    // corner relay, 32-bit accesses by whole warps.  DIAG: [tk+1][tj] with stride 48 -- the 32 lanes of a warp (two
    // wrapped anti-diagonals) and the border writers (consecutive tj of row 0) all hit distinct banks; otherwise
    // [tj][tk+1] (a warp spans 4 tj x 8 tk)
    static constexpr int CORNER_ROWS = DIAG ? TK + 1 : TJ;
    static constexpr int CORNER_STRIDE = DIAG ? 48 : TK + 8;
    static constexpr int TABJ_STRIDE = 68;                      // j-indexed tables: lanes of a wavefront differ in code only
    static constexpr int TABK_STRIDE = 128;                     // k-indexed tables: {cAC, tC} per code, see tabk_index
    static constexpr int SKEW = LAG * (TJ + TK - 2);
    // pair q is staged into buffer q % NBUF while earlier pairs drain: LAG 1 (2 buffers): after pair q-2 has left the
    // last thread, SKEW steps after pair q-1 entered; LAG 2 (3 buffers): pair q-3 left long ago
    static constexpr int LOAD_AT = LAG == 1 ? SKEW + 2 : 8, BUILD_AT = LOAD_AT + 4;
    static_assert(LAG == 1 || LAG == 2, "NSLOT must be a power of two");
    static_assert(BJ == 64 && BK == 64, "K1 keeps the whole 64x64 (j,k) plane in one CTA");
    static_assert(RJ == 4 && (RK == 4 || RK == 2), "edge exchange is laid out for 128-bit / 64-bit groups");
    static_assert(TK % 8 == 0 && TJ % 4 == 0 && BUILD_AT < L - 1, "thread_pos / staging schedule");
};

// position of the k-indexed constants (which = 0: cAC, 1: tC) inside a code's TABK_STRIDE words: a thread's RK values
// of both kinds are fetched with 128-bit loads whose bank group depends on tk only (the codes differ between lanes)
template <class C>
TA_HD int tabk_index(int which, int k)
{
    return C::RK == 4 ? which * 64 + k : (k / C::RK) * (2 * C::RK) + which * C::RK + (k % C::RK);
}

struct BatchArgs {
    const uint8_t *a, *b, *c;  // [n][len] device
    long long n;               // triples
    int len;                   // <= 64
    int32_t *scores;           // [n]
    long long pairs_total;     // ceil(n/2)
    int pairs_per_cta;         // contiguous chunk per CTA
    int base;                  // potential base (common.cuh envelope), < 2^15
};

// Range-checked variant (CHECK kernels, used when wrap-freedom at score_bits cannot be proved from the length): every
// sampled `all` must lie in [check_lo, check_hi]; pair_flags[pair] gets bit 0 / bit 1 set when the low / high triple of
// the pair came too close to the score_bits limits (the dispatcher re-runs those on the literal kernel).  A separate
// kernel parameter, so that the unchecked kernels' parameter block (and ptxas's allocation) stays as it was.
struct BatchCheck {
    int check_lo, check_hi;
    int *pair_flags;
};

// ---- run-time scoring parameters on the host side (fast_core.cuh: ScoreGen) ----
template <bool SOP>
inline ScoreGen<SOP> make_score_gen(int M, int X, int GO, int GE)
{
    ScoreGen<SOP> g;
    g.pi_ = (uint32_t)(2 * GE);
    g.pi2_ = rep((uint32_t)(2 * GE));
    g.dec_ = 0u - rep((uint32_t)(GO - GE));
    g.dech_ = rep((uint32_t)(0x10000 - (GO - GE)));
    g.pair_eq = rep((uint32_t)(M + 3 * GE));
    g.pair_d = rep((uint32_t)(M - X));
    g.row_eq = rep((uint32_t)(3 * M + 6 * GE));
    g.row_d = rep((uint32_t)(3 * M - 3 * X));
    g.col_eq = rep((uint32_t)(3 * M + 6 * GE));
    g.col_d = rep((uint32_t)(3 * M - 2 * (M + X)));
    g.sop_sub = rep((uint32_t)(3 * GE));
    return g;
}
// margins of the range check in true score units: a candidate exceeds the largest `all` of its predecessors by at most
// add_max and undercuts the smallest by at most 2 * (2 GO - add_min) (a state is at least all(pred) - 2 GO + add_min)
struct ScoreMargins {
    int add_max, add_min;
};
inline ScoreMargins score_margins(int M, int X)
{
    const int v[7] = {3 * M, 2 * (M + X), 3 * X, M + 2 * X, M, X, 0};
    ScoreMargins m{0, 0};
    for (int x : v) {
        m.add_max = x > m.add_max ? x : m.add_max;
        m.add_min = x < m.add_min ? x : m.add_min;
    }
    return m;
}
// potential base: every state is at least -(2 GO + 2 GE m) + 3 min(X, 0) (a gap chain from the zero border)
inline int score_base_len(int X, int GO, int GE, int len) { return 2 * GE * len + 8 * GO + 16 + 6 * (X < 0 ? -X : 0); }
inline int score_base(int X, int GO, int GE) { return score_base_len(X, GO, GE, 64); }
// every step is sampled; score_bits > 16 cannot wrap at these lengths and parameter sizes: limits out of reach
inline BatchCheck score_check(int M, int X, int GO, int check_bits, int *pair_flags)
{
    const ScoreMargins m = score_margins(M, X);
    const long long lim = check_bits <= 16 ? (1ll << (check_bits - 1)) : (1ll << 20);
    BatchCheck chk;
    chk.check_hi = (int)(lim - 1 - m.add_max);
    chk.check_lo = (int)(-lim + 2 * (2 * GO - m.add_min));
    chk.pair_flags = pair_flags;
    return chk;
}

constexpr int kBatchCheckEvery = 8;  // steps between two range samples: `all` moves by at most +3 / -4 per plane

template <class C>
struct BatchSmem {
    // Index 0 of the tj / tk dimension is the zero BORDER plane (src/TriAlign_1cyc.v:156-181: all 7 states are 0, i.e.
    // every reduced value equals the potential), written ahead of the step that reads it, so that every thread loads its
    // halos with the same instructions -- no border branches, no per-lane register shuffling in the step (5 of 8 warps
    // used to run the border code on top of the common path).  Writers: all threads of the CTA, one 128-bit group each
    // (batch2_write_borders, batch_write_borders_all: the default kernels); the edge threads for themselves
    // (batch_write_borders: the LAG 1 A/B variants); the lanes of the reader's own warp (batch2_write_borders_warp).
    // Every [slot][array] plane is followed by four padding words: plane strides of 4 / 12 (mod 32) words put the eight
    // (array, plane-of-step) groups that one access group of the border writers touches into distinct banks.
    struct RowPlane {
        uint32_t v[C::TJ + 1][C::ROW_STRIDE];
        uint32_t pad[4];
    };
    struct ColPlane {
        uint32_t v[C::TK + 1][C::COL_STRIDE];
        uint32_t pad[4];
    };
    RowPlane row[C::NSLOT][4];  // [slot][wY,u2,u1(old),all(old)].v[tj+1][k]   bottom rows
    ColPlane col[C::NSLOT][4];  // [slot][wZ,u2,u3(old),all(old)].v[tk+1][j]   right columns
    uint32_t corner[C::NSLOT][2][C::CORNER_ROWS][C::CORNER_STRIDE];  // [slot][u2,all(old)] x (tj, tk+1), see corner_at: relay of the cell above-right of a tile
    uint32_t tabAB[C::NBUF][16][C::TABJ_STRIDE];     // [buf][a code][j]  pair constant {4,2}
    uint32_t tabMB[C::NBUF][16][C::TABJ_STRIDE];     // [buf][a code][j]  triple row constant {9,3}
    uint32_t tabK[C::NBUF][16][C::TABK_STRIDE];      // [buf][a code][tabk_index]  pair constant {4,2} and triple column constant {9,6}
    uint32_t bpk[C::NBUF][64], cpk[C::NBUF][64];           // [buf][pos]  b / c symbols of the two triples, one per 16-bit half
    uint8_t seq[C::NBUF][3][2][64];                  // [buf][A,B,C][half][pos]
    uint8_t acode[C::NBUF][64];                      // a_lo | a_hi << 2
};

// corner relay entry of tile row tj, column index x (0 = the k = 0 border plane, tk + 1 = written by thread (tj,tk))
template <class C>
TA_HD uint32_t &corner_at(BatchSmem<C> &S, int slot, int arr, int tj, int x)
{
    return C::DIAG ? S.corner[slot][arr][x][tj] : S.corner[slot][arr][tj][x];
}
template <class C>
TA_HD const uint32_t &corner_at(const BatchSmem<C> &S, int slot, int arr, int tj, int x)
{
    return C::DIAG ? S.corner[slot][arr][x][tj] : S.corner[slot][arr][tj][x];
}

template <class C>
struct BatchRegs {
    uint32_t all[C::RJ][C::RK], wX[C::RJ][C::RK], u1[C::RJ][C::RK], u3[C::RJ][C::RK];  // plane i-1
    uint32_t cBC[C::RJ][C::RK];                                                        // per-cell constant
};

// Constants of the step a thread is about to run (a_i against the tile's rows / columns), fetched one step ahead in
// the shadow of the barrier (batch_thread_prepare).
template <class C>
struct BatchConsts {
    uint32_t cAB[C::RJ], tB[C::RJ], cAC[C::RK], tC[C::RK];
};

// thread id -> position in the TJ x TK thread grid.
// 256 threads (DIAG): eight consecutive lanes (the unit of a 128-bit shared-memory access) hold eight consecutive tj of
// one WRAPPED anti-diagonal (tj + tk) mod 16 = dd; a warp holds the wrapped diagonals 2w and 2w+1.  Lanes on one
// diagonal are on the same plane, so a warp sees at most 4 distinct skews instead of the 11 of a 4x8 patch: the
// per-pair reset (117 instructions, executed by the whole warp whenever one lane enters a new pair) runs 4 instead
// of 11 times per pair, and the a-code / score-table look-ups of an access group mostly hit one address.  Bank
// groups: rows 64 tj + 4 tk -> tk mod 8, columns 72 tk + 4 tj -> (2 tk + tj) mod 8 = (2 dd - tj) mod 8, k-indexed
// tables -> tk mod 8, j-indexed tables -> (code + tj) mod 8: all distinct within the eight lanes.
// 512 threads: a quarter-warp holds 8 consecutive tk of one tj, a warp spans 4 tj x 8 tk.
template <class C>
TA_HD void thread_pos(int tid, int &tj, int &tk)
{
    if (C::DIAG) {
        const int oct = tid >> 3, dd = oct >> 1;
        tj = 8 * (oct & 1) + (tid & 7);
        tk = (dd - tj) & 15;
    } else {
        const int lane = tid & 31, warp = tid >> 5;
        tk = (lane & 7) + 8 * (warp % (C::TK / 8));
        tj = (lane >> 3) + 4 * (warp / (C::TK / 8));
    }
}

TA_HD uint32_t phi(const BatchArgs &A, int i, int jl, int kl) { return rep((uint32_t)(A.base + 2 * (i + jl + kl))); }
template <class SC>
TA_HD uint32_t phi(const BatchArgs &A, int i, int jl, int kl, const SC &sc)
{
    return rep((uint32_t)(A.base + (int)sc.pi() * (i + jl + kl)));
}

// ---- CTA-uniform actions, executed by every thread at fixed global steps ----
// PAD (the range-checked kernels): positions beyond len get codes that match nothing, see below.
template <class C, bool PAD = false>
TA_HD void load_pair_sequences(BatchSmem<C> &S, const BatchArgs &A, int tid, long long pair, int buf)
{
    for (int e = tid; e < 3 * 2 * 64; e += C::NT) {
        const int which = e >> 7, half = (e >> 6) & 1, pos = e & 63;
        const long long t = 2 * pair + half;
        const uint8_t *src = which == 0 ? A.a : (which == 1 ? A.b : A.c);
        // positions beyond len are padding: b gets code 4, c code 5 -- they match nothing, not even each other, so
        // padded cells decay instead of growing like a run of matches (keeps the range check free of false alarms;
        // real cells never depend on padded ones)
        uint8_t v = (!PAD || which == 0) ? 0 : (uint8_t)(3 + which);
        if (t < A.n && pos < A.len) v = src[t * A.len + pos] & 3;  // PE keeps bits [1:0] (src/PE_1cyc.v:63-66)
        S.seq[buf][which][half][pos] = v;
    }
}

template <class C, class SC = ScoreRtl>
TA_HD void build_pair_tables(BatchSmem<C> &S, int tid, int buf, const SC &sc = SC())
{
    for (int e = tid; e < 16 * 64; e += C::NT) {
        const int code = e >> 6, p = e & 63;
        const uint32_t ap = (uint32_t)(code & 3) | ((uint32_t)(code >> 2) << 16);
        const uint32_t bp = (uint32_t)S.seq[buf][1][0][p] | ((uint32_t)S.seq[buf][1][1][p] << 16);
        const uint32_t cp = (uint32_t)S.seq[buf][2][0][p] | ((uint32_t)S.seq[buf][2][1][p] << 16);
        const uint32_t nab = neq_flag(ap, bp), nac = neq_flag(ap, cp);
        S.tabAB[buf][code][p] = pair_const(nab, sc);
        S.tabMB[buf][code][p] = triple_row_const(nab, sc);
        S.tabK[buf][code][tabk_index<C>(0, p)] = pair_const(nac, sc);
        S.tabK[buf][code][tabk_index<C>(1, p)] = triple_col_const(nac, sc);
        if (code == 0) {
            S.bpk[buf][p] = bp;
            S.cpk[buf][p] = cp;
        }
    }
    for (int p = tid; p < 64; p += C::NT) S.acode[buf][p] = (uint8_t)(S.seq[buf][0][0][p] | (S.seq[buf][0][1][p] << 2));
}

// ---- the part of step s that does not depend on the neighbours' step s-1 (runs between arrive and wait) ----
// Zero-border halos of plane i into edge slot `slot` (the slot the step of plane i reads): index 0 of the tj / tk
// dimension, written and read by thread (0,tk) / (tj,0) only, so no barrier is involved.
template <class C>
TA_HD void batch_write_borders(BatchSmem<C> &S, const BatchArgs &A, int tj, int tk, int slot, int i)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    if (tj == 0) {  // j = 0 border plane (src/TriAlign_1cyc.v:156-165)
        const int k0 = tk * RK;
        const uint32_t e = phi(A, i, 0, k0 + 1);
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            const uint32_t now = e + (uint32_t)c * 0x00020002u, old = now - 0x00020002u;
            S.row[slot][0].v[0][k0 + c] = now;
            S.row[slot][1].v[0][k0 + c] = now;
            S.row[slot][2].v[0][k0 + c] = old;
            S.row[slot][3].v[0][k0 + c] = old;
        }
    }
    if (tk == 0) {  // k = 0 border plane (src/TriAlign_1cyc.v:166-174); the corner cell (j0, 0) lies on it too
        const int j0 = tj * RJ;
        const uint32_t e = phi(A, i, j0, 0);
        corner_at<C>(S, slot, 0, tj, 0) = e;
        corner_at<C>(S, slot, 1, tj, 0) = e - 0x00020002u;
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
            const uint32_t now = e + (uint32_t)(r + 1) * 0x00020002u, old = now - 0x00020002u;
            S.col[slot][0].v[0][j0 + r] = now;
            S.col[slot][1].v[0][j0 + r] = now;
            S.col[slot][2].v[0][j0 + r] = old;
            S.col[slot][3].v[0][j0 + r] = old;
        }
    }
}

// Core: pair p (CTA-local, -1 while the pipeline fills), plane i.  RESET = false: only the constants of plane i.
// code_in >= 0: the a code of plane i, fetched a step ahead by the caller (two-plane kernel).
template <class C, bool RESET = true, class SC = ScoreRtl>
TA_HD void batch_thread_prepare_at(BatchRegs<C> &R, BatchConsts<C> &K, const BatchSmem<C> &S, const BatchArgs &A, int tj,
                                   int tk, int p, int i, const SC &sc = SC(), int code_in = -1)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int buf = (p + C::NBUF) % C::NBUF;
    const int j0 = tj * RJ, k0 = tk * RK;  // block-local index of the cell left/above the tile

    if (RESET && i == 1) {
        // new pair enters this thread: plane 0 is the all-zero border (src/PE_1cyc.v:164-178 override), and the
        // tile's (b,c) constants change
        const uint32_t zb = phi(A, 0, j0 + 1, k0 + 1, sc);
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const uint32_t z = zb + (uint32_t)(r + c) * sc.pi2();
                R.all[r][c] = z; R.wX[r][c] = z; R.u1[r][c] = z; R.u3[r][c] = z;
                const uint32_t nbc = neq_flag(S.bpk[buf][j0 + r], S.cpk[buf][k0 + c]);
                R.cBC[r][c] = pair_const(nbc, sc);
            }
    }

    const int code = code_in >= 0 ? code_in : S.acode[buf][i - 1];
    TA_CHECK(code >= 0 && code < 16 && code == S.acode[buf][i - 1]);
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        K.cAB[r] = S.tabAB[buf][code][j0 + r];
        K.tB[r] = S.tabMB[buf][code][j0 + r];
    }
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        K.cAC[c] = S.tabK[buf][code][tabk_index<C>(0, k0 + c)];
        K.tC[c] = S.tabK[buf][code][tabk_index<C>(1, k0 + c)];
    }
}

template <class C, class SC = ScoreRtl>
TA_HD void batch_thread_prepare(BatchRegs<C> &R, BatchConsts<C> &K, BatchSmem<C> &S, const BatchArgs &A, int tj,
                                int tk, int s, const SC &sc = SC())
{
    static_assert(SC::kRtl || (C::LAG == 2 && C::DIAG), "run-time scoring: lagged-barrier 256-thread kernel only");
    const int u = s - C::LAG * (tj + tk);
    batch_thread_prepare_at<C, true, SC>(R, K, S, A, tj, tk, u >> 6, (u & 63) + 1, sc);  // pair u / 64, plane 1..64
    // the lagged-barrier 256-thread kernel spreads the border halos over all threads (batch_write_borders_all)
    if (!(C::LAG == 2 && C::DIAG) && (tj == 0 || tk == 0))
        batch_write_borders<C>(S, A, tj, tk, (s - C::LAG) & (C::NSLOT - 1), (u & 63) + 1);
}

// Lagged-barrier kernel (LAG 2, 256 threads): the zero-border halos that step st reads (edge slot (st-2) & 3, last read at
// step st-4), written by the first 160 threads -- one 128-bit group or one corner word each -- at the top of step st-2:
// the writer has seen everybody finish step st-4, and the readers start step st only after everybody's arrival for
// step st-2.  (With the anti-diagonal thread layout every warp holds edge threads; written by the edge threads for
// themselves, as in the LAG 1 kernels, every warp would run the border code with one or two lanes.)
// A reader (0,tk) / (tj,0) is at plane ((st - 2t) & 63) + 1 at step st (t = tk / tj).
template <class C, class SC = ScoreRtl>
TA_HD void batch_write_borders_all(BatchSmem<C> &S, const BatchArgs &A, int tid, int st, const SC &sc = SC())
{
    static_assert(C::LAG == 2 && C::DIAG && C::RJ == 4 && C::RK == 4, "one 128-bit group per thread");
    const int slot = (st - 2) & (C::NSLOT - 1), t = tid & 15;
    const int i = ((st - 2 * t) & 63) + 1;
    if (tid < 128) {
        const int arr = (tid >> 4) & 3, col = tid >> 6;
        const uint32_t e = rep((uint32_t)(A.base + (int)sc.pi() * (i + 4 * t + 1 - (arr >> 1)))), d = sc.pi2();
        uint32_t *dst = col ? &S.col[slot][arr].v[0][4 * t] : &S.row[slot][arr].v[0][4 * t];
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4 *>(dst) = make_uint4(e, e + d, e + 2 * d, e + 3 * d);
#else
        for (int c = 0; c < 4; ++c) dst[c] = e + (uint32_t)c * d;
#endif
    } else if (tid < 160) {  // the corner cell (j0, 0) of the k = 0 border plane
        const int arr = (tid >> 4) & 1;
        corner_at<C>(S, slot, arr, t, 0) = rep((uint32_t)(A.base + (int)sc.pi() * (i + 4 * t - arr)));
    }
}

// ---- one wavefront step of one thread: halos published at step s-1 -> tile -> publish for step s+1 ----
// R = plane i-1 state (input), Rn = plane i state (output; may be R itself): the lagged-barrier kernel alternates two
// register sets in a step loop unrolled by two, so that the loop-carried state needs no register moves.
// CHECK: the run-time range check (see BatchArgs), sampled every kBatchCheckEvery-th step (warp-uniform).
// `viol` (CHECK only) accumulates the violations of the pair this thread is working on; it lives outside BatchRegs so
// that the unchecked kernels compile exactly as without the check.
// The halos of one plane of a thread's tile (see batch_load_halos).
template <class C>
struct BatchHalo {
    uint32_t T_wY[C::RK], T_u1[C::RK], T_u2[C::RK + 1], T_all[C::RK + 1], L_wZ[C::RJ], L_u3[C::RJ], L_u2[C::RJ], L_all[C::RJ];
};

template <class C>
TA_HD void batch_load_halos(BatchHalo<C> &H, const BatchSmem<C> &S, const BatchArgs &A, int tj, int tk, int rs, int i)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
    uint32_t (&T_wY)[RK] = H.T_wY, (&T_u1)[RK] = H.T_u1, (&T_u2)[RK + 1] = H.T_u2, (&T_all)[RK + 1] = H.T_all;
    uint32_t (&L_wZ)[RJ] = H.L_wZ, (&L_u3)[RJ] = H.L_u3, (&L_u2)[RJ] = H.L_u2, (&L_all)[RJ] = H.L_all;

    // halos: top = bottom row of thread (tj-1,tk), left = right column of (tj,tk-1), both published at step s-1.
    // T_u2/T_all[c] belong to column k0+c (c = 0 is the corner cell, c = RK the cell above-right of the tile);
    // L_u2/L_all[r] belong to row j0+r.  The corner cell belongs to thread (tj-1,tk-1), which is TWO steps ahead,
    // so its value is relayed by the left neighbour (which read it one step ago as its own T_*[RK]).
    // "old" = plane i-1 values, which the neighbours publish one step late (the RTL's _d1/_d2 delay registers,
    // src/PE_1cyc.v:80-109): they are known at the START of the publishing step, so those stores leave early.
    // rs: the edge slot read (written by the neighbours LAG steps ago; index 0 = border, written by this thread itself)
    (void)A;
    (void)i;
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        T_wY[c] = S.row[rs][0].v[tj][k0 + c];
        T_u1[c] = S.row[rs][2].v[tj][k0 + c];
        T_u2[c + 1] = S.row[rs][1].v[tj][k0 + c];
        T_all[c + 1] = S.row[rs][3].v[tj][k0 + c];
    }
    T_u2[0] = corner_at<C>(S, rs, 0, tj, tk);
    T_all[0] = corner_at<C>(S, rs, 1, tj, tk);
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        L_wZ[r] = S.col[rs][0].v[tk][j0 + r];
        L_u3[r] = S.col[rs][2].v[tk][j0 + r];
    }
#pragma unroll
    for (int r = 1; r < RJ; ++r) {
        L_u2[r] = S.col[rs][1].v[tk][j0 + r - 1];
        L_all[r] = S.col[rs][3].v[tk][j0 + r - 1];
    }
    L_u2[0] = T_u2[0];
    L_all[0] = T_all[0];
}

// Halos in hand: relay the corner, publish the old half of the own edges, update the tile, publish the new half (slot ws).
template <class C, bool CHECK = false, class SC = ScoreRtl>
TA_HD void batch_thread_step_halo(const BatchRegs<C> &R, BatchRegs<C> &Rn, const BatchConsts<C> &K, const BatchHalo<C> &H,
                                  BatchSmem<C> &S, const BatchArgs &A, int tj, int tk, int ws, int u, int p, int i, int s,
                                  int P, long long pair_base, uint32_t *viol_io = nullptr, const BatchCheck *chk = nullptr,
                                  const SC &sc = SC())
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
    const uint32_t (&T_wY)[RK] = H.T_wY, (&T_u1)[RK] = H.T_u1, (&T_u2)[RK + 1] = H.T_u2, (&T_all)[RK + 1] = H.T_all;
    const uint32_t (&L_wZ)[RJ] = H.L_wZ, (&L_u3)[RJ] = H.L_u3, (&L_u2)[RJ] = H.L_u2, (&L_all)[RJ] = H.L_all;
    corner_at<C>(S, ws, 0, tj, tk + 1) = T_u2[RK];
    corner_at<C>(S, ws, 1, tj, tk + 1) = T_all[RK];
    // the plane i-1 half of this thread's own edges is already known
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        S.row[ws][2].v[tj + 1][k0 + c] = R.u1[RJ - 1][c];
        S.row[ws][3].v[tj + 1][k0 + c] = R.all[RJ - 1][c];
    }
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        S.col[ws][2].v[tk + 1][j0 + r] = R.u3[r][RK - 1];
        S.col[ws][3].v[tk + 1][j0 + r] = R.all[r][RK - 1];
    }

    // the tile, row-major (within a plane a cell needs its j-1, k-1 and (j-1,k-1) neighbours)
    uint32_t n_all[RJ][RK], n_wX[RJ][RK], n_u1[RJ][RK], n_u3[RJ][RK], n_wY[RJ][RK], n_wZ[RJ][RK], n_u2[RJ][RK];
#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            const uint32_t p_all = (r > 0 && c > 0) ? R.all[r - 1][c - 1] : (r == 0 ? T_all[c] : L_all[r]);
            const uint32_t p_wY = r > 0 ? n_wY[r - 1][c] : T_wY[c];
            const uint32_t p_wZ = c > 0 ? n_wZ[r][c - 1] : L_wZ[r];
            const uint32_t p_u1 = r > 0 ? R.u1[r - 1][c] : T_u1[c];
            const uint32_t p_u2 = (r > 0 && c > 0) ? n_u2[r - 1][c - 1] : (r == 0 ? T_u2[c] : L_u2[r]);
            const uint32_t p_u3 = c > 0 ? R.u3[r][c - 1] : L_u3[r];
            const uint32_t cM = triple_const_of(K.tB[r], K.tC[c], K.cAB[r], R.cBC[r][c], K.cAC[c], sc);
            const CellOut o = cell_update<false, SC>(p_all, R.wX[r][c], p_wY, p_wZ, p_u1, p_u2, p_u3, cM, K.cAB[r],
                                                     R.cBC[r][c], K.cAC[c], nullptr, sc);
            n_all[r][c] = o.all; n_wX[r][c] = o.wX; n_u1[r][c] = o.u1; n_u3[r][c] = o.u3;
            n_wY[r][c] = o.wY; n_wZ[r][c] = o.wZ; n_u2[r][c] = o.u2;
        }

    // publish the plane-i half of the tile's bottom row and right column for the neighbours' step s+1
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        S.row[ws][0].v[tj + 1][k0 + c] = n_wY[RJ - 1][c];
        S.row[ws][1].v[tj + 1][k0 + c] = n_u2[RJ - 1][c];
    }
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        S.col[ws][0].v[tk + 1][j0 + r] = n_wZ[r][RK - 1];
        S.col[ws][1].v[tk + 1][j0 + r] = n_u2[r][RK - 1];
    }

    if (CHECK) {
        uint32_t viol = i == 1 ? 0u : *viol_io;  // a new pair enters this thread
        // run-time scoring: every step is sampled (the +3 / -4 per plane bound belongs to the RTL's constants)
        if ((SC::kRtl ? (s & (kBatchCheckEvery - 1)) == 0 : true) && i <= A.len) {
            // only cells of the real triple count: padding (j or k beyond len) decays without bound of its own
            uint32_t mx = 0u, mn = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < RJ; ++r)
#pragma unroll
                for (int c = 0; c < RK; ++c)
                    if (j0 + r < A.len && k0 + c < A.len) {
                        // run-time scoring: pi can be large against a narrow score_bits, so the cell's own share of
                        // the potential is taken off first (FMA pipe) and the limits below are exact per cell
                        const uint32_t v = SC::kRtl ? n_all[r][c] : fadd(n_all[r][c], 0u - (uint32_t)(r + c) * sc.pi2());
                        mx = vmax2(mx, v);
                        mn = vmin2(mn, v);
                    }
            // true value = stored - phi(cell); phi grows by 2 per row / column inside the tile
            const int ph = A.base + (int)sc.pi() * (i + j0 + 1 + k0 + 1), slack = SC::kRtl ? 2 * (RJ + RK - 2) : 0;
            const int h = ph + chk->check_hi, l = ph + chk->check_lo + slack;
            const uint32_t hi_lim = rep((uint32_t)(h > 0xFFFF ? 0xFFFF : h)), lo_lim = rep((uint32_t)(l < 0 ? 0 : l));
            viol |= (vmax2(mx, hi_lim) ^ hi_lim) | (vmin2(mn, lo_lim) ^ lo_lim);
        }
        // the pair leaves this thread: report
        if (i == C::L && viol != 0 && u >= 0 && p < P) {
            const int bits = ((viol & 0xFFFFu) ? 1 : 0) | ((viol >> 16) ? 2 : 0);
#if defined(__CUDA_ARCH__)
            atomicOr(chk->pair_flags + (pair_base + p), bits);
#else
            chk->pair_flags[pair_base + p] |= bits;
#endif
        }
        *viol_io = viol;
    }

    // Score = max of the 7 states of cell (len,len,len) (src/TriAlign_1cyc.v:141-142) = `all` of that cell
    if (i == A.len && u >= 0 && p < P && tj == (A.len - 1) / RJ && tk == (A.len - 1) / RK) {
        const int rf = (A.len - 1) % RJ, cf = (A.len - 1) % RK;
        uint32_t v = 0;
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c)
                if (r == rf && c == cf) v = n_all[r][c];
        const int off = A.base + (int)sc.pi() * (3 * A.len);
        const long long t0 = 2 * (pair_base + p);
        if (t0 < A.n) A.scores[t0] = (int32_t)(v & 0xFFFFu) - off;
        if (t0 + 1 < A.n) A.scores[t0 + 1] = (int32_t)(v >> 16) - off;
    }

#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            Rn.all[r][c] = n_all[r][c]; Rn.wX[r][c] = n_wX[r][c]; Rn.u1[r][c] = n_u1[r][c]; Rn.u3[r][c] = n_u3[r][c];
            Rn.cBC[r][c] = R.cBC[r][c];
        }
}

// rs / ws: the edge slot read (written by the neighbours LAG steps ago) and the one written now
template <class C, bool CHECK = false, class SC = ScoreRtl>
TA_HD void batch_thread_step_at(const BatchRegs<C> &R, BatchRegs<C> &Rn, const BatchConsts<C> &K, BatchSmem<C> &S,
                                const BatchArgs &A, int tj, int tk, int rs, int ws, int u, int p, int i, int s, int P,
                                long long pair_base, uint32_t *viol_io = nullptr, const BatchCheck *chk = nullptr,
                                const SC &sc = SC())
{
    BatchHalo<C> H;
    batch_load_halos<C>(H, S, A, tj, tk, rs, i);
    batch_thread_step_halo<C, CHECK, SC>(R, Rn, K, H, S, A, tj, tk, ws, u, p, i, s, P, pair_base, viol_io, chk, sc);
}

template <class C, bool CHECK = false, class SC = ScoreRtl>
TA_HD void batch_thread_step(const BatchRegs<C> &R, BatchRegs<C> &Rn, const BatchConsts<C> &K, BatchSmem<C> &S,
                             const BatchArgs &A, int tj, int tk, int s, int P, long long pair_base, uint32_t *viol_io = nullptr,
                             const BatchCheck *chk = nullptr, const SC &sc = SC())
{
    const int u = s - C::LAG * (tj + tk);
    batch_thread_step_at<C, CHECK, SC>(R, Rn, K, S, A, tj, tk, (s - C::LAG) & (C::NSLOT - 1), s & (C::NSLOT - 1), u, u >> 6,
                                       (u & 63) + 1, s, P, pair_base, viol_io, chk, sc);
}

// ---------------------------------------------------------------------------------------------------------------
// Two planes per step (K1 "P2").  Thread (tj,tk) runs planes 2u+1 and 2u+2 of pair u / 32 at global step s, u = s - (tj+tk):
// one barrier, one spin, one round of index arithmetic and staging checks per 32 cells instead of per 16.  The four
// edge slots of the lagged-barrier layout hold {even step, odd step} x {first, second plane}: edges are consumed ONE step
// after publication (the barrier is waited for every step, but there are half as many), plane by plane, so the step is
// two plane updates back to back (R -> R2 -> R: the register sets alternate as in the unrolled single-plane loop).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBatch2Steps = 32;   // steps per pair: 64 planes, two per step
constexpr int kBatch2PrefetchAt = 2, kBatch2LoadAt = 8, kBatch2BuildAt = 12;

template <class C>
TA_HD int batch2_total_steps(int P) { return P * kBatch2Steps + (C::TJ + C::TK - 2); }

// the a codes of the two planes thread (tj,tk) runs at step s, as one 16-bit load (plane 2x+1 sits at an even offset)
template <class C>
TA_HD uint32_t batch2_fetch_codes(const BatchSmem<C> &S, int tj, int tk, int s)
{
    const int u = s - (tj + tk), p = u >> 5, x = u & 31;
    const uint8_t *src = &S.acode[(p + C::NBUF) % C::NBUF][2 * x];
    return (uint32_t)src[0] | ((uint32_t)src[1] << 8);
}

// Neighbour-only synchronisation ("ring", option k1=256r2).  With the anti-diagonal thread layout warp w holds the wrapped
// diagonals 2w and 2w+1: it reads the edges of diagonal 2w-1 (warp w-1) and its own, and the slots it writes were last
// read by diagonal 2w+2 (warp w+1) -- a warp only needs "warps w-1 and w+1 finished step s" before step s+1, not the
// whole CTA.  The border halos must then be warp-local too: lane l of warp w writes, for the edge threads of its own
// warp -- (0,t) [rows] and (t,0) [columns], t = 2w, 2w+1 -- array (l >> 1) & 3 of plane l & 1; the eight lanes of an
// access group hit distinct banks thanks to the plane padding of BatchSmem.
template <class C>
TA_HD void batch2_write_borders_warp(BatchSmem<C> &S, const BatchArgs &A, int tid, int st)
{
    static_assert(C::DIAG && C::NT == 256 && C::RJ == 4 && C::RK == 4, "one 128-bit group per thread");
    const int rs = ((st - 1) & 1) << 1, w = tid >> 5, l = tid & 31;
    {
        const int e = l >> 3, col = e >> 1, t = 2 * w + (e & 1), arr = (l >> 1) & 3, pl = l & 1;
        const int i = 2 * ((st - t) & 31) + 1 + pl;
        const uint32_t v = rep((uint32_t)(A.base + 2 * (i + 4 * t + 1 - (arr >> 1))));
        uint32_t *dst = col ? &S.col[rs + pl][arr].v[0][4 * t] : &S.row[rs + pl][arr].v[0][4 * t];
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4 *>(dst) = make_uint4(v, v + 0x00020002u, v + 0x00040004u, v + 0x00060006u);
#else
        for (int c = 0; c < 4; ++c) dst[c] = v + (uint32_t)c * 0x00020002u;
#endif
    }
    if (l < 8) {  // the corner cell (j0, 0) of the k = 0 border plane
        const int t = 2 * w + (l & 1), arr = (l >> 1) & 1, pl = l >> 2;
        const int i = 2 * ((st - t) & 31) + 1 + pl;
        corner_at<C>(S, rs + pl, arr, t, 0) = rep((uint32_t)(A.base + 2 * (i + 4 * t - arr)));
    }
}

// codes: batch2_fetch_codes(s), fetched one step ahead (the pair's codes are staged long before its first plane enters
// thread (0,0)), so that the table look-ups below do not wait for a shared-memory round trip; on return it holds the
// codes of step s + 1.
template <class C>
TA_HD void batch2_thread_prepare(BatchRegs<C> &R, BatchConsts<C> &K0, BatchConsts<C> &K1, BatchSmem<C> &S,
                                 const BatchArgs &A, int tj, int tk, int s, uint32_t &codes)
{
    const int u = s - (tj + tk), p = u >> 5, i0 = 2 * (u & 31) + 1;
    const uint32_t now = codes;
    codes = batch2_fetch_codes<C>(S, tj, tk, s + 1);
    batch_thread_prepare_at<C, true>(R, K0, S, A, tj, tk, p, i0, ScoreRtl(), (int)(now & 0xFFu));
    batch_thread_prepare_at<C, false>(R, K1, S, A, tj, tk, p, i0 + 1, ScoreRtl(), (int)(now >> 8));
}

// Zero-border halos (index 0 of the edge slots' tj / tk dimension, see BatchSmem) of the two planes that step st reads,
// written by ALL threads of the CTA -- one 128-bit group each, the first 64 threads one corner word on top -- instead of
// by the edge threads for themselves: the edge threads sit in 5 of the 8 warps, every step ends at a CTA barrier and the
// warp with the most duties paces it.  Called at the top of step st-1 (any time between "everybody finished step st-2",
// the last reader of these slots, and the caller's arrival at the barrier of step st-1).
// A reader (0,tk) / (tj,0) at step st is at planes 2((st-t)&31)+1 and +2 (t = tk / tj); the potential of its halo cells
// is the same expression of t for the row and the column: 2(i + 4t + 1 + c) [- 2 for the plane i-1 ("old") arrays].
template <class C>
TA_HD void batch2_write_borders(BatchSmem<C> &S, const BatchArgs &A, int tid, int st)
{
    static_assert(C::NT == 256 && C::RJ == 4 && C::RK == 4 && C::TJ == 16 && C::TK == 16, "one 128-bit group per thread");
    const int rs = ((st - 1) & 1) << 1;
    {
        const int t = tid & 15, arr = (tid >> 4) & 3, pl = (tid >> 6) & 1, col = tid >> 7;
        const int i = 2 * ((st - t) & 31) + 1 + pl;
        const uint32_t e = rep((uint32_t)(A.base + 2 * (i + 4 * t + 1 - (arr >> 1))));
        uint32_t *dst = col ? &S.col[rs + pl][arr].v[0][4 * t] : &S.row[rs + pl][arr].v[0][4 * t];
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4 *>(dst) = make_uint4(e, e + 0x00020002u, e + 0x00040004u, e + 0x00060006u);
#else
        for (int c = 0; c < 4; ++c) dst[c] = e + (uint32_t)c * 0x00020002u;
#endif
    }
    if (tid < 64) {  // the corner cell (j0, 0) of the k = 0 border plane
        const int t = tid & 15, arr = (tid >> 4) & 1, pl = tid >> 5;
        const int i = 2 * ((st - t) & 31) + 1 + pl;
        corner_at<C>(S, rs + pl, arr, t, 0) = rep((uint32_t)(A.base + 2 * (i + 4 * t - arr)));
    }
}

template <class C>
TA_HD void batch2_thread_step(BatchRegs<C> &R, BatchRegs<C> &R2, const BatchConsts<C> &K0, const BatchConsts<C> &K1,
                              BatchSmem<C> &S, const BatchArgs &A, int tj, int tk, int s, int P, long long pair_base)
{
    static_assert(C::NSLOT == 4, "two step parities x two planes");
    const int u = s - (tj + tk), p = u >> 5, i0 = 2 * (u & 31) + 1;
    const int rs = ((s - 1) & 1) << 1, ws = (s & 1) << 1;
    // both planes' halos first: they were published a step ago, and loading the second plane's only after the first
    // plane's stores would put a shared-memory round trip into the middle of the step
    BatchHalo<C> H0, H1;
    batch_load_halos<C>(H0, S, A, tj, tk, rs, i0);
    batch_load_halos<C>(H1, S, A, tj, tk, rs + 1, i0 + 1);
    batch_thread_step_halo<C, false>(R, R2, K0, H0, S, A, tj, tk, ws, u, p, i0, s, P, pair_base);
    batch_thread_step_halo<C, false>(R2, R, K1, H1, S, A, tj, tk, ws + 1, u, p, i0 + 1, s, P, pair_base);
}

// the six 64-byte sequences of a pair are read with byte loads a few steps later: pull their lines in now, so that the
// loads (and the stores that wait for them) do not stall on HBM in the middle of a step
template <class C>
TA_HD void prefetch_pair_sequences(const BatchArgs &A, int tid, long long pair)
{
#if defined(__CUDA_ARCH__)
    if (tid < 12) {
        const int which = tid >> 2, half = (tid >> 1) & 1, end = tid & 1;
        const long long t = 2 * pair + half;
        if (t < A.n) {
            const uint8_t *src = (which == 0 ? A.a : (which == 1 ? A.b : A.c)) + t * A.len + (end ? A.len - 1 : 0);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
        }
    }
#else
    (void)A; (void)tid; (void)pair;
#endif
}

// RING (neighbour-only synchronisation, see batch2_write_borders_warp): a warp sees another warp's step s - d - 1 for
// ring distance d <= 4, so the byte loads run six steps before the table build instead of four.
template <class C, bool RING = false>
TA_HD void batch2_uniform_actions(BatchSmem<C> &S, const BatchArgs &A, int tid, int s, int P, long long pair_base)
{
    const int q = s / kBatch2Steps + 1;  // pair that starts entering the array at step 32*q
    if (q < P) {
        if (s % kBatch2Steps == (RING ? 0 : kBatch2PrefetchAt)) prefetch_pair_sequences<C>(A, tid, pair_base + q);
        if (s % kBatch2Steps == (RING ? kBatch2LoadAt - 2 : kBatch2LoadAt)) load_pair_sequences<C>(S, A, tid, pair_base + q, q % C::NBUF);
        if (s % kBatch2Steps == kBatch2BuildAt) build_pair_tables<C>(S, tid, q % C::NBUF);
    }
}

// Steps a CTA needs for P pairs, and the CTA-uniform actions of global step s.
template <class C>
TA_HD int batch_total_steps(int P) { return P * C::L + C::SKEW; }

template <class C, bool PAD = false, class SC = ScoreRtl>
TA_HD void batch_uniform_actions(BatchSmem<C> &S, const BatchArgs &A, int tid, int s, int P, long long pair_base,
                                 const SC &sc = SC())
{
    const int q = (s >> 6) + 1;  // pair that starts entering the array at step 64*q
    if (q < P) {
        if ((s & 63) == C::LOAD_AT) load_pair_sequences<C, PAD>(S, A, tid, pair_base + q, q % C::NBUF);
        if ((s & 63) == C::BUILD_AT) build_pair_tables<C, SC>(S, tid, q % C::NBUF, sc);
    }
}

}  // namespace fast
}  // namespace trialign
