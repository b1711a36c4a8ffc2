// fast_long.cuh -- K2L: block-sweep kernel for ONE long triple (BASELINE configs 3 and 5), second generation.
//
// The RTL's slicing method (src/TriAlign_1cyc.v:127-140,289-335):
// the (j,k) plane is cut into blocks, every block is swept over all of A, block faces stream CTA -> CTA through
// HBM/L2 gated by per-task progress counters; the two 16-bit halves of a word are two k-ADJACENT 64x64 blocks, the
// high half running PLAG planes behind the low half (thread (tj,tk)'s high half is the logical thread (tj,tk+16) of a
// 16x32 array) -- on the step engine of K1 / K2S (fast_batch.cuh, fast_stream.cuh):
//   * split-phase mbarrier instead of __syncthreads: LAG 1 (arrive after publishing, wait before the next halo loads;
//     shortest block-to-block pipeline fill, used for short sweeps) or LAG 2 (edges consumed two steps after
//     publication, a warp only waits for "everybody finished step s-2"; used for long sweeps);
//   * block halos are private to the edge threads (0,tk) / (tj,0), which fetch them with cp.async TWO steps ahead of use
//     (or write the zero-border potential, src/TriAlign_1cyc.v:156-181): the step has no border branches, only the
//     left-edge threads merge "previous pair's high block" with "this pair's low block" (hi_lo);
//   * a symbols of both halves pre-combined into one code per step;
//   * tile patches dealt to the warps so that every scheduler carries the same edge duty (see stream_thread_pos).
// Task descriptors, face rings, slab hand-over layout: sweep_common.cuh (SweepArgs, SweepTask).
#pragma once

#include "fast_stream.cuh"

namespace trialign {
namespace fast {

template <int LAG_>
struct Long2Cfg {
    static constexpr bool DIST = false;  // block-edge duties stay with the edge threads (stream_thread_pos)
    static constexpr int RJ = 4, RK = 4, TJ = 16, TK = 16, LAG = LAG_;
    static constexpr int NSLOT = 2 * LAG;
    static constexpr int BJ = RJ * TJ, BK = RK * TK, NT = TJ * TK;
    static constexpr int PLAG = LAG * TK;              // planes the high half runs behind the low half
    static constexpr int DEPTH = LAG * (TJ + TK - 2);  // largest physical skew in steps
    static constexpr int KSPAN = 2 * BK;               // columns of C one task covers
    static constexpr int ROW_W = BK, COL_W = BJ + 4, CORNER_STRIDE = TK + 8;
    static constexpr int TABJ_STRIDE = 68, TABK_STRIDE = 128;
    static constexpr int LA_MAX = 8192 + 64;           // planes of one launch
    static constexpr int ACODE_MAX = LA_MAX + PLAG + 2 * DEPTH + 8;
    static constexpr uint32_t HI_OFF = (uint32_t)(2 * (BK - PLAG)) << 16;  // phi(high) - phi(low)
    static constexpr int FJ_ROW_WORDS = 4 * ROW_W;     // one plane of a J face
    static constexpr int FK_ROW_WORDS = 4 * COL_W;     // one plane of a K face
    static constexpr int HOP_J = LAG * (TJ - 1) + 1;          // producer steps a J-face row is behind its consumer's step
    static constexpr int HOP_K = LAG * (TK - 1) + PLAG + 1;   // the same for a K-face row
    static constexpr int AHEAD = LAG + 2;              // consumption steps the gate thread certifies ahead (see kernel)
    static_assert(LAG == 1 || LAG == 2, "NSLOT must be a power of two");
    static_assert(PLAG < BK, "potential offset of the high half");
};

template <class C>
struct Long2Smem {
    uint32_t row[C::NSLOT][4][C::TJ + 1][C::ROW_W];   // [slot][wY,u2,u1(old),all(old)][tj+1 (0 = block halo)][k]
    uint32_t col[C::NSLOT][4][C::TK + 1][C::COL_W];   // [slot][wZ,u2,u3(old),all(old)][tk+1 (0 = block halo)][j]
    uint32_t corner[C::NSLOT][2][C::TJ][C::CORNER_STRIDE];  // [slot][u2,all(old)][tj][tk+1 (0 = block halo)]
    uint32_t tabAB[16][C::TABJ_STRIDE];               // [a code][j]  pair constant {4,2}
    uint32_t tabMB[16][C::TABJ_STRIDE];               // [a code][j]  triple row constant {9,3}
    uint32_t tabK[16][C::TABK_STRIDE];                // [a code][k] pair constant, [64 + k] triple column constant {9,6}
    uint32_t bpk[64], cpk[64];                        // b / c symbols of the task's blocks, low | high << 16
    uint32_t tabBC[C::BJ][C::BK];                     // [j][k]  (b,c) pair constant {4,2} of every cell: read every step
                                                      // instead of living in 16 registers (this kernel is register-bound)
    uint8_t acode[C::ACODE_MAX];                      // [i0 - 1 + DEPTH]  a(i0) | a(i0 - PLAG) << 2, indices clamped
    SweepTask task;
};

template <class C>
struct Long2Regs {
    uint32_t all[C::RJ][C::RK], wX[C::RJ][C::RK], u1[C::RJ][C::RK], u3[C::RJ][C::RK];  // plane i-1
    uint32_t viol;                                                                     // range-check violations seen
};

template <class C>
struct Long2Consts {
    uint32_t cAB[C::RJ], tB[C::RJ], cAC[C::RK], tC[C::RK];
};

// potential of a cell in block-local coordinates; the high half is the same expression for (i - PLAG, kl + BK), which
// is a constant offset
template <class C>
TA_HD uint32_t long2_phi(const SweepArgs &A, int i0, int jl, int kl)
{
    return rep((uint32_t)(A.base + 2 * (i0 + jl + kl))) + C::HI_OFF;
}

template <class C>
TA_HD int long2_planes(const SweepArgs &A) { return A.la + C::PLAG; }              // low-half planes 1 .. this value
template <class C>
TA_HD int long2_task_steps(const SweepArgs &A) { return A.la + C::PLAG + C::DEPTH - 1; }  // steps 0 .. this value

// Middle phase of a task: steps long2_mid_begin .. long2_mid_end (may be empty).  There every thread has both halves
// live (PLAG + 2 <= i0 <= la - 3), no half starts or finishes, and every face row fetched (two steps ahead) or stored
// exists -- the step loop runs a compact instantiation (RARE = false) without the plane-0 reset, the range checks, the
// score / final-state extraction and the slab hand-over.
template <class C>
TA_HD int long2_mid_begin(const SweepArgs &) { return C::DEPTH + C::PLAG + 1; }
template <class C>
TA_HD int long2_mid_end(const SweepArgs &A) { return A.la - 4; }

// ---- once per CTA: the a symbols of both halves for every step offset ----
template <class C>
TA_HD void long2_load_a(Long2Smem<C> &S, const SweepArgs &A, int tid)
{
    const int n = long2_planes<C>(A) + 2 * C::DEPTH + 2;
    for (int x = tid; x < n; x += C::NT) {
        int ia0 = x - C::DEPTH, ia1 = ia0 - C::PLAG;  // i0 - 1 and i1 - 1
        ia0 = ia0 < 0 ? 0 : (ia0 >= A.la ? A.la - 1 : ia0);
        ia1 = ia1 < 0 ? 0 : (ia1 >= A.la ? A.la - 1 : ia1);
        S.acode[x] = (uint8_t)((A.a[ia0] & 3) | ((A.a[ia1] & 3) << 2));  // PE keeps bits [1:0] (src/PE_1cyc.v:63-66)
    }
}

// ---- per task (CTA-uniform; barriers between the phases are the caller's) ----
template <class C>
TA_HD void long2_load_bc(Long2Smem<C> &S, const SweepArgs &A, int tid)
{
    const SweepTask &T = S.task;
    for (int e = tid; e < 2 * 64; e += C::NT) {
        const int which = e >> 6, pos = e & 63;
        // cells beyond lb / lc are padding: codes 4 and 5 match nothing (not even each other), so padded cells never
        // create values above their real neighbours (keeps the run-time range check free of false alarms)
        if (which == 0) {  // B rows of this block: both halves share them
            const int j = T.bj * C::BJ + pos;
            const uint32_t v = j < A.lb ? (A.b[j] & 3) : 4;
            S.bpk[pos] = v | (v << 16);
        } else {           // C columns: the high half is the next 64 columns
            const int k = T.bkp * C::KSPAN + pos;
            const uint32_t v0 = k < A.lc ? (A.c[k] & 3) : 5, v1 = k + C::BK < A.lc ? (A.c[k + C::BK] & 3) : 5;
            S.cpk[pos] = v0 | (v1 << 16);
        }
    }
}

template <class C>
TA_HD void long2_build_tables(Long2Smem<C> &S, int tid)
{
    for (int e = tid; e < 16 * 64; e += C::NT) {
        const int code = e >> 6, p = e & 63;
        const uint32_t ap = (uint32_t)(code & 3) | ((uint32_t)(code >> 2) << 16);
        const uint32_t nab = neq_flag(ap, S.bpk[p]), nac = neq_flag(ap, S.cpk[p]);
        S.tabAB[code][p] = pair_const(nab);
        S.tabMB[code][p] = triple_row_const(nab);
        S.tabK[code][p] = pair_const(nac);
        S.tabK[code][64 + p] = triple_col_const(nac);
    }
}

template <class C>
TA_HD void long2_init_thread(Long2Regs<C> &R, Long2Smem<C> &S, int tj, int tk)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            S.tabBC[j0 + r][k0 + c] = pair_const(neq_flag(S.bpk[j0 + r], S.cpk[k0 + c]));
            R.all[r][c] = R.wX[r][c] = R.u1[r][c] = R.u3[r][c] = 0;
        }
    R.viol = 0;
}

// ---- block-edge halos of consumption step c, into the edge slot that step reads.  Only thread (0,tk) / (tj,0) writes
// and reads these entries (index 0 of the tj / tk dimension): cp.async + wait_group, no barrier involved. ----
// RARE = false: the caller guarantees that every face row asked for exists (steps of the middle phase, long2_mid_*).
template <class C, bool RARE = true>
TA_HD void long2_edge_fetch(Long2Smem<C> &S, const SweepArgs &A, int tj, int tk, int c)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const SweepTask &T = S.task;
    const int slot = (c + C::LAG) & (C::NSLOT - 1);  // == (c - LAG) mod NSLOT
    const int rows = long2_planes<C>(A);             // face rows 1..rows exist
    if (tj == 0) {
        const int i = c - C::LAG * tk + 1;  // low-half plane of thread (0,tk) at step c
        const int k0 = tk * RK;
        if (T.fj_r) {
            if (!RARE || (i >= 1 && i <= rows)) {
                const uint32_t *src = T.fj_r + (long long)i * C::FJ_ROW_WORDS + k0;
#pragma unroll
                for (int v = 0; v < 4; ++v) ta_copy16(&S.row[slot][v][0][k0], src + v * C::ROW_W);
            }
        } else {  // j = 0 border plane: all 7 states are 0 (src/TriAlign_1cyc.v:156-165), reduced values = phi
            const uint32_t e = long2_phi<C>(A, i, 0, k0 + 1);
#pragma unroll
            for (int q = 0; q < RK; ++q) {
                const uint32_t now = e + (uint32_t)q * 0x00020002u, old = now - 0x00020002u;
                S.row[slot][0][0][k0 + q] = now;
                S.row[slot][1][0][k0 + q] = now;
                S.row[slot][2][0][k0 + q] = old;
                S.row[slot][3][0][k0 + q] = old;
            }
        }
    }
    if (tk == 0) {
        const int i = c - C::LAG * tj + 1;  // low-half plane of thread (tj,0) at step c
        const int j0 = tj * RJ;
        if (T.fk_r) {
            // the producer's high half is PLAG planes behind its row index.  Beyond the last plane (i > la) the low half
            // is dead but must stay BOUNDED while the high half is still live (32-bit adds carry from the low lane into
            // the high one, fast_core.cuh): it keeps consuming the last row instead of whatever the slot held.
            int src_row = i + C::PLAG;
            if (RARE && src_row > rows) src_row = rows;
            if (!RARE || src_row >= 1) {
                const uint32_t *src = T.fk_r + (long long)src_row * C::FK_ROW_WORDS + j0;
#pragma unroll
                for (int v = 0; v < 4; ++v) ta_copy16(&S.col[slot][v][0][j0], src + v * C::COL_W);
                // corner cell (row above the tile, last column of the pair to the left): face entry j0-1, or the relay
                // entry for the block's first tile row
                const int cidx = tj > 0 ? j0 - 1 : C::BJ;
                ta_copy4(&S.corner[slot][0][tj][0], T.fk_r + (long long)src_row * C::FK_ROW_WORDS + 1 * C::COL_W + cidx);
                ta_copy4(&S.corner[slot][1][tj][0], T.fk_r + (long long)src_row * C::FK_ROW_WORDS + 3 * C::COL_W + cidx);
            }
        } else {  // k = 0 border plane (src/TriAlign_1cyc.v:166-174); the consumer takes the HIGH half of these words
            const uint32_t e = (uint32_t)(A.base + 2 * (i + j0));  // corner cell (j0, 0) lies on the border plane too
            S.corner[slot][0][tj][0] = e << 16;
            S.corner[slot][1][tj][0] = (e - 2) << 16;
#pragma unroll
            for (int r = 0; r < RJ; ++r) {
                const uint32_t now = e + 2u * (uint32_t)(r + 1), old = now - 2;
                S.col[slot][0][0][j0 + r] = now << 16;
                S.col[slot][1][0][j0 + r] = now << 16;
                S.col[slot][2][0][j0 + r] = old << 16;
                S.col[slot][3][0][j0 + r] = old << 16;
            }
        }
    }
}

// ---- the part of step s1 that does not depend on the neighbours: plane-0 reset, the step's constants, and the halo
// fetch of step s1 + 1 ----
// RARE = false: steps of the middle phase (long2_mid_*), where no half starts and every face row exists.
template <class C, bool RARE = true>
TA_HD void long2_thread_prepare(Long2Regs<C> &R, Long2Consts<C> &K, Long2Smem<C> &S, const SweepArgs &A, int tj, int tk,
                                int s1, int tid)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const SweepTask &T = S.task;
    const int i0 = s1 - C::LAG * (tj + tk) + 1;  // low-half plane (valid 1..la)
    const int i1 = i0 - C::PLAG;                 // high-half plane
    const int j0 = tj * RJ, k0 = tk * RK;

    // plane 0 is the all-zero border (PE override, src/PE_1cyc.v:164-178): reset the half whose plane 1 starts now
    if (RARE && (i0 == 1 || i1 == 1)) {
        const uint32_t keep = (i0 == 1 ? 0u : 0x0000FFFFu) | (i1 == 1 ? 0u : 0xFFFF0000u);
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                // long2_phi(i0-1,...) is phi(plane 0) in whichever half starts now: the low half starts at i0 = 1, the
                // high half at i0 = 1 + PLAG (its own plane index is i0 - PLAG)
                const uint32_t zz = long2_phi<C>(A, i0 - 1, j0 + r + 1, k0 + c + 1) & ~keep;
                if (A.iface_in) {  // continue from the previous slab instead of the zero border
                    const uint32_t *src = A.iface_in + (((long long)T.self_task * 4) * 16 + (r * RK + c)) * C::NT + tid;
                    R.all[r][c] = (R.all[r][c] & keep) | (src[0 * 16 * C::NT] & ~keep);
                    R.wX[r][c] = (R.wX[r][c] & keep) | (src[1 * 16 * C::NT] & ~keep);
                    R.u1[r][c] = (R.u1[r][c] & keep) | (src[2 * 16 * C::NT] & ~keep);
                    R.u3[r][c] = (R.u3[r][c] & keep) | (src[3 * 16 * C::NT] & ~keep);
                } else {
                    R.all[r][c] = (R.all[r][c] & keep) | zz;
                    R.wX[r][c] = (R.wX[r][c] & keep) | zz;
                    R.u1[r][c] = (R.u1[r][c] & keep) | zz;
                    R.u3[r][c] = (R.u3[r][c] & keep) | zz;
                }
            }
    }

    const int code = S.acode[i0 - 1 + C::DEPTH];
    TA_CHECK(i0 - 1 + C::DEPTH >= 0 && i0 - 1 + C::DEPTH < C::ACODE_MAX && code >= 0 && code < 16);
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        K.cAB[r] = S.tabAB[code][j0 + r];
        K.tB[r] = S.tabMB[code][j0 + r];
    }
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        K.cAC[c] = S.tabK[code][k0 + c];
        K.tC[c] = S.tabK[code][64 + k0 + c];
    }
    if (tj == 0 || tk == 0) long2_edge_fetch<C, RARE>(S, A, tj, tk, s1 + 1);
    ta_copy_commit();
}

// ---- one wavefront step of one thread ----
// R = plane i-1 state (input), Rn = plane i state (output; may be R itself): the kernel alternates two register sets in
// a step loop unrolled by two, so that the loop-carried state needs no register moves.
// RARE = false: steps of the middle phase (long2_mid_*): both halves live, every face row in range, no half finishes.
template <class C, bool CHECK, bool RARE = true>
TA_HD void long2_thread_step(const Long2Regs<C> &R, Long2Regs<C> &Rn, const Long2Consts<C> &K, Long2Smem<C> &S,
                             const SweepArgs &A, int tj, int tk, int s, int tid)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const SweepTask &T = S.task;
    const int i0 = s - C::LAG * (tj + tk) + 1;  // low-half plane (valid 1..la)
    const int i1 = i0 - C::PLAG;                // high-half plane
    const int j0 = tj * RJ, k0 = tk * RK;
    const int rs = (s + C::LAG) & (C::NSLOT - 1), ws = s & (C::NSLOT - 1);  // rs == (s - LAG) mod NSLOT

    uint32_t viol = R.viol;

    // halos; index 0 of the tj / tk dimension is the block halo, filled by long2_edge_fetch
    uint32_t T_wY[RK], T_u1[RK], T_u2[RK + 1], T_all[RK + 1], L_wZ[RJ], L_u3[RJ], L_u2[RJ], L_all[RJ];
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        T_wY[c] = S.row[rs][0][tj][k0 + c];
        T_u1[c] = S.row[rs][2][tj][k0 + c];
        T_u2[c + 1] = S.row[rs][1][tj][k0 + c];
        T_all[c + 1] = S.row[rs][3][tj][k0 + c];
    }
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        L_wZ[r] = S.col[rs][0][tk][j0 + r];
        L_u3[r] = S.col[rs][2][tk][j0 + r];
    }
#pragma unroll
    for (int r = 1; r < RJ; ++r) {
        L_u2[r] = S.col[rs][1][tk][j0 + r - 1];
        L_all[r] = S.col[rs][3][tk][j0 + r - 1];
    }
    T_u2[0] = S.corner[rs][0][tj][tk];
    T_all[0] = S.corner[rs][1][tj][tk];
    if (tk == 0) {
        // left edge of the pair: low half <- high half of the block halo (the previous pair's high block), high half
        // <- low half of thread (tj,15)'s column (this pair's low block, one logical thread to the left)
        T_u2[0] = hi_lo(T_u2[0], S.corner[rs][0][tj][C::TK]);
        T_all[0] = hi_lo(T_all[0], S.corner[rs][1][tj][C::TK]);
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
            L_wZ[r] = hi_lo(L_wZ[r], S.col[rs][0][C::TK][j0 + r]);
            L_u3[r] = hi_lo(L_u3[r], S.col[rs][2][C::TK][j0 + r]);
        }
#pragma unroll
        for (int r = 1; r < RJ; ++r) {
            L_u2[r] = hi_lo(L_u2[r], S.col[rs][1][C::TK][j0 + r - 1]);
            L_all[r] = hi_lo(L_all[r], S.col[rs][3][C::TK][j0 + r - 1]);
        }
    }
    L_u2[0] = T_u2[0];
    L_all[0] = T_all[0];
    S.corner[ws][0][tj][tk + 1] = T_u2[RK];
    S.corner[ws][1][tj][tk + 1] = T_all[RK];
    // the plane i-1 half of this thread's own edges is known already: these stores leave before the tile
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        S.row[ws][2][tj + 1][k0 + c] = R.u1[RJ - 1][c];
        S.row[ws][3][tj + 1][k0 + c] = R.all[RJ - 1][c];
    }
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        S.col[ws][2][tk + 1][j0 + r] = R.u3[r][RK - 1];
        S.col[ws][3][tk + 1][j0 + r] = R.all[r][RK - 1];
    }

    // the tile
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
            const uint32_t cM = triple_const(K.tB[r], K.tC[c]);
            const CellOut o = cell_update<false>(p_all, R.wX[r][c], p_wY, p_wZ, p_u1, p_u2, p_u3, cM, K.cAB[r],
                                                 S.tabBC[j0 + r][k0 + c], K.cAC[c], nullptr);
            n_all[r][c] = o.all; n_wX[r][c] = o.wX; n_u1[r][c] = o.u1; n_u3[r][c] = o.u3;
            n_wY[r][c] = o.wY; n_wZ[r][c] = o.wZ; n_u2[r][c] = o.u2;
        }

    // publish the tile's bottom row and right column for the neighbours' step s + LAG
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        S.row[ws][0][tj + 1][k0 + c] = n_wY[RJ - 1][c];
        S.row[ws][1][tj + 1][k0 + c] = n_u2[RJ - 1][c];
    }
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        S.col[ws][0][tk + 1][j0 + r] = n_wZ[r][RK - 1];
        S.col[ws][1][tk + 1][j0 + r] = n_u2[r][RK - 1];
    }

    // block faces to global memory, re-based to the consumer's frame (its halo row / column has local index 0)
    const int rows = long2_planes<C>(A);
    if (tj == C::TJ - 1 && T.fj_w && (!RARE || (i0 >= 1 && i0 <= rows))) {
        const uint32_t adj = rep(2 * C::BJ);
        uint32_t *dst = T.fj_w + (long long)i0 * C::FJ_ROW_WORDS + k0;
        ta_store4(dst + 0 * C::ROW_W, n_wY[RJ - 1][0] - adj, n_wY[RJ - 1][1] - adj, n_wY[RJ - 1][2] - adj, n_wY[RJ - 1][3] - adj);
        ta_store4(dst + 1 * C::ROW_W, n_u2[RJ - 1][0] - adj, n_u2[RJ - 1][1] - adj, n_u2[RJ - 1][2] - adj, n_u2[RJ - 1][3] - adj);
        ta_store4(dst + 2 * C::ROW_W, R.u1[RJ - 1][0] - adj, R.u1[RJ - 1][1] - adj, R.u1[RJ - 1][2] - adj, R.u1[RJ - 1][3] - adj);
        ta_store4(dst + 3 * C::ROW_W, R.all[RJ - 1][0] - adj, R.all[RJ - 1][1] - adj, R.all[RJ - 1][2] - adj, R.all[RJ - 1][3] - adj);
    }
    if (tk == C::TK - 1 && T.fk_w && (!RARE || (i0 >= 1 && i0 <= rows))) {
        // only the high half crosses to the next column pair (KSPAN columns to the right of this pair's origin)
        const uint32_t adj = (uint32_t)(2 * C::KSPAN) << 16;
        uint32_t *dst = T.fk_w + (long long)i0 * C::FK_ROW_WORDS + j0;
        ta_store4(dst + 0 * C::COL_W, n_wZ[0][RK - 1] - adj, n_wZ[1][RK - 1] - adj, n_wZ[2][RK - 1] - adj, n_wZ[3][RK - 1] - adj);
        ta_store4(dst + 1 * C::COL_W, n_u2[0][RK - 1] - adj, n_u2[1][RK - 1] - adj, n_u2[2][RK - 1] - adj, n_u2[3][RK - 1] - adj);
        ta_store4(dst + 2 * C::COL_W, R.u3[0][RK - 1] - adj, R.u3[1][RK - 1] - adj, R.u3[2][RK - 1] - adj, R.u3[3][RK - 1] - adj);
        ta_store4(dst + 3 * C::COL_W, R.all[0][RK - 1] - adj, R.all[1][RK - 1] - adj, R.all[2][RK - 1] - adj, R.all[3][RK - 1] - adj);
        if (tj == 0) {  // relay: the cell above-right of the block pair = corner of the next pair's first tile
            uint32_t *rel = T.fk_w + (long long)i0 * C::FK_ROW_WORDS + C::BJ;
            rel[1 * C::COL_W] = T_u2[RK] - adj;
            rel[3 * C::COL_W] = T_all[RK] - adj;
        }
    }

    // Run-time wrap check (used when wrap-freedom at score_bits cannot be proved a priori): `all`
    // moves by at most +3 / -4 per plane, so sampling every 16th plane of a cell with 45 / 60 margins bounds every
    // plane.  Sampled on every 16th STEP: warp-uniform (sampling by plane index made 11 of 16 steps of a warp diverge
    // into this block, its lanes having 11 different skews).
    if (CHECK && (s & 15) == 0) {
        uint32_t mx = n_all[0][0], mn = n_all[0][0];
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; c += 2) {
                mx = vmax3(mx, n_all[r][c], n_all[r][c + 1]);
                mn = vmin3(mn, n_all[r][c], n_all[r][c + 1]);
            }
        // true value = stored - phi(cell); phi grows by 2 per row / column inside the tile (at most 12)
        const uint32_t ph = long2_phi<C>(A, i0, j0 + 1, k0 + 1);
        const int p0 = (int)(ph & 0xFFFFu), p1 = (int)(ph >> 16), slack = 2 * (RJ + RK - 2);
        const int h0 = p0 + A.check_hi, h1 = p1 + A.check_hi, l0 = p0 + A.check_lo + slack, l1 = p1 + A.check_lo + slack;
        const uint32_t hi_lim = (uint32_t)(h0 > 0xFFFF ? 0xFFFF : h0) | ((uint32_t)(h1 > 0xFFFF ? 0xFFFF : h1) << 16);
        const uint32_t lo_lim = (uint32_t)(l0 < 0 ? 0 : l0) | ((uint32_t)(l1 < 0 ? 0 : l1) << 16);
        const uint32_t bad = (vmax2(mx, hi_lim) ^ hi_lim) | (vmin2(mn, lo_lim) ^ lo_lim);
        const uint32_t live = !RARE ? 0xFFFFFFFFu
                                    : ((i0 >= 1 && i0 <= A.la) ? 0x0000FFFFu : 0u) | ((i1 >= 1 && i1 <= A.la) ? 0xFFFF0000u : 0u);
        viol |= bad & live;
    }

    // Score = max of the 7 states of cell (la,lb,lc) (src/TriAlign_1cyc.v:141-142) = `all` of that cell
    if (RARE && T.final_here && A.last_slab) {
        const int jl = A.lb - T.bj * C::BJ;                 // 1..64
        const int kk = A.lc - T.bkp * C::KSPAN;             // 1..KSPAN
        const int hf = (kk - 1) / C::BK;                    // the half that holds column lc
        const int kl = kk - hf * C::BK;                     // 1..64
        const int iplane = hf ? i1 : i0;
        if (iplane == A.la && tj == (jl - 1) / RJ && tk == (kl - 1) / RK) {
            const int rf = (jl - 1) % RJ, cf = (kl - 1) % RK;
            uint32_t v = 0, st[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int r = 0; r < RJ; ++r)
#pragma unroll
                for (int c = 0; c < RK; ++c)
                    if (r == rf && c == cf) {
                        v = n_all[r][c];
                        // the 7 states of the cell, recomputed from its inputs (rare path)
                        const uint32_t p_all = (r > 0 && c > 0) ? R.all[r - 1][c - 1] : (r == 0 ? T_all[c] : L_all[r]);
                        const uint32_t p_wY = r > 0 ? n_wY[r - 1][c] : T_wY[c];
                        const uint32_t p_wZ = c > 0 ? n_wZ[r][c - 1] : L_wZ[r];
                        const uint32_t p_u1 = r > 0 ? R.u1[r - 1][c] : T_u1[c];
                        const uint32_t p_u2 = (r > 0 && c > 0) ? n_u2[r - 1][c - 1] : (r == 0 ? T_u2[c] : L_u2[r]);
                        const uint32_t p_u3 = c > 0 ? R.u3[r][c - 1] : L_u3[r];
                        cell_update<true>(p_all, R.wX[r][c], p_wY, p_wZ, p_u1, p_u2, p_u3,
                                          triple_const(K.tB[r], K.tC[c]), K.cAB[r], S.tabBC[j0 + r][k0 + c], K.cAC[c], st);
                    }
            // potential of the cell in its half's own frame (high half: kl + BK, plane la)
            const int off = A.base + 2 * (A.la + jl + kl) + (hf ? 2 * C::BK : 0);
            const int shv = hf ? 16 : 0;
            A.out[0] = (int32_t)((v >> shv) & 0xFFFFu) - off;
            if (A.out_states)
                for (int q = 0; q < 7; ++q) A.out_states[q] = (int32_t)((st[q] >> shv) & 0xFFFFu) - off;
        }
    }

#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            Rn.all[r][c] = n_all[r][c]; Rn.wX[r][c] = n_wX[r][c]; Rn.u1[r][c] = n_u1[r][c]; Rn.u3[r][c] = n_u3[r][c];
        }
    Rn.viol = viol;

    // i-slab hand-over: the half that has just finished the slab's last plane stores its state for the next slab,
    // re-based by the slab's 2*la so that the reader's plane index restarts at 0 (16-bit stores: the halves finish
    // PLAG steps apart)
    if (RARE && A.iface_out && (i0 == A.la || i1 == A.la)) {
        const uint32_t adj = rep((uint32_t)(2 * A.la));
        uint16_t *dst16 = reinterpret_cast<uint16_t *>(A.iface_out + (((long long)T.self_task * 4) * 16) * C::NT + tid);
        const int hsel = (i0 == A.la) ? 0 : 1;
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const long long w = (long long)(r * RK + c) * C::NT;  // word offset of this cell within a value plane
                const uint32_t v0 = Rn.all[r][c] - adj, v1 = Rn.wX[r][c] - adj, v2 = Rn.u1[r][c] - adj, v3 = Rn.u3[r][c] - adj;
                dst16[2 * (w + 0LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v0 >> 16 : v0);
                dst16[2 * (w + 1LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v1 >> 16 : v1);
                dst16[2 * (w + 2LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v2 >> 16 : v2);
                dst16[2 * (w + 3LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v3 >> 16 : v3);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Dependency tracking (evaluated by the gate thread).  progress[t] = number of wavefront steps of task t that every
// thread has completed and that have been published (kProgressDone once the task is finished).  Thread (0,tk) consumes
// at its step c the J-face row producer thread (15,tk) wrote at producer step c + LAG*15; thread (tj,0) consumes the
// K-face row producer thread (tj,15) wrote at step c + LAG*15 + PLAG.
//   long2_certified(): the last consumption step whose face rows are complete, given the producers' progress.
//   Publication rhythm: long2_publish_now().
// ---------------------------------------------------------------------------------------------------------------
template <class C>
TA_HD int long2_certified(const SweepArgs &A, const SweepTask &T, int top_prog, int left_prog)
{
    const int total = long2_task_steps<C>(A) + 1;  // steps 0..steps
    int cert = kCertifiedAll;
    if (T.top_task >= 0 && top_prog < total) cert = top_prog - C::HOP_J;
    if (T.left_task != -1 && left_prog < total) {  // <= -2: a block row of the left neighbour GPU (column pipeline)
        const int c2 = left_prog - C::HOP_K;
        cert = c2 < cert ? c2 : cert;
    }
    return cert;
}

// claim order: tasks sorted by their earliest possible start (both predecessors have a smaller key)
template <class C>
TA_HD int long2_order_key(int bj, int bkp, int gate) { return (C::HOP_J + C::LAG + gate / 2) * bj + (C::HOP_K + C::LAG + gate / 2) * bkp; }

// Does the gate thread publish the task's progress once `done` steps are complete?  Every A.gate steps, plus the two
// counts the tasks below / to the right are waiting for before they can start (their prologue needs HOP + 1 + LAG
// steps).  A release is a MEMBAR.GPU with the task's face stores in flight, so there are as few as the hops allow.
template <class C>
TA_HD bool long2_publish_now(const SweepArgs &A, int done)
{
    if (A.early) {
        const int gate = (done < kEarlySteps && A.gate > kEarlyGate) ? kEarlyGate : A.gate;
        return (done & (gate - 1)) == 0;
    }
    return done == C::HOP_J + 1 + C::LAG || done == C::HOP_K + 1 + C::LAG || (done & (A.gate - 1)) == 0;
}

}  // namespace fast
}  // namespace trialign
