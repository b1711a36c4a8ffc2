// fast_stream2.cuh -- K2S "P2": the streaming block-sweep kernel for batches of triples longer than 64 (BASELINE config 4)
// with K1's two-planes-per-step structure (fast_batch.cuh: fast_batch2_kernel).
//
// What the ncu profile of the one-plane form (fast_stream.cuh) showed: 31.9 warp-instructions per cell pair against
// K1's 26.0, 2.6 of them branches and reconvergence pairs of the per-plane bookkeeping; 7 % of the samples on
// `no_instructions` (the step loop unrolled over two register sets is 3,200 instructions); 7 % in the barrier spin.
// Here a thread runs planes 2x+1 and 2x+2 of its task at step x: one barrier, one position update, one round of
// staging checks and duty bookkeeping per TWO planes, and the two register sets alternate inside ONE loop body.
//   * a task = one 64x64 (j,k) block of one PAIR of triples swept over L2 = ceil(len/2) steps (an odd len gets one
//     padding plane: a code 0, results unused -- the next task resets the plane state, the score is taken at plane len);
//     thread (tj,tk) runs step x of the CTA's n-th task at global step  s = n*L2 + x + (tj+tk);
//   * edge slots: {even, odd step} x {first, second plane}; edges are consumed one step after publication (split-phase
//     CTA barrier, waited for every step);
//   * block-edge duties by thread id, one 128-bit group in and one out per thread and step (2 planes x 128 groups):
//     at the top of step s a thread brings in the halos that an edge thread reads at step s+1 (cp.async from the face
//     rings, or the zero-border potential) and copies out what an edge thread published at step s-1 (index 16 of the
//     edge slots) -- the step itself has no edge code, and the lanes of a warp lie along anti-diagonals as in K1.
// CHECK / SC template parameters: the range-checked and run-time-scoring forms (see fast_stream.cuh, fast_core.cuh).
#pragma once

#include "fast_stream.cuh"

namespace trialign {
namespace fast {

struct Stream2Cfg {
    static constexpr bool DIST = true;  // stream_thread_pos: anti-diagonal lanes
    static constexpr int RJ = 4, RK = 4, TJ = 16, TK = 16;
    static constexpr int NSLOT = 4;         // {step parity} x {plane of the step}
    static constexpr int NBUF = 3;          // task buffers: two tasks in flight + one being staged
    static constexpr int BJ = RJ * TJ, BK = RK * TK, NT = TJ * TK;
    static constexpr int ROW_STRIDE = BK, COL_STRIDE = BJ + 8, CORNER_STRIDE = TK + 8;
    static constexpr int TABJ_STRIDE = 68, TABK_STRIDE = 128;
    static constexpr int SKEW = TJ + TK - 2;   // steps between thread (0,0) and thread (15,15)
    static constexpr int LEN_MIN = 65;         // L2 >= 33 > SKEW: at most two tasks in flight
    static constexpr int LEN_MAX = 2048;
    static constexpr int LOAD_AT = 4, BUILD_AT = 8;  // steps after a task enters thread (0,0): stage the next one
    static constexpr int FJ_ROW_WORDS = 4 * ROW_STRIDE;
    static constexpr int FK_ROW_WORDS = 4 * COL_STRIDE;
    static_assert(BUILD_AT + 8 < (LEN_MIN + 1) / 2, "staging schedule");
};

// where a thread is in its stream of tasks: step x of task t covers planes 2x+1 and 2x+2
struct Stream2Pos {
    int t;    // CTA-local task index (-1 / tasks: fillers)
    int x;    // 0 .. L2-1
    int buf;  // (t + NBUF) % NBUF
};

TA_HD int stream2_steps_per_task(const StreamArgs &A) { return (A.len + 1) >> 1; }

template <class C>
TA_HD void stream2_advance(const StreamArgs &A, Stream2Pos &p)
{
    if (++p.x == stream2_steps_per_task(A)) {
        p.x = 0;
        ++p.t;
        p.buf = p.buf + 1 == C::NBUF ? 0 : p.buf + 1;
    }
}

// position at global step s of a thread with skew `back` (= tj + tk); s - back >= -L2
template <class C>
TA_HD Stream2Pos stream2_pos_at(const StreamArgs &A, int back, int s)
{
    const int L2 = stream2_steps_per_task(A), u = s - back;
    Stream2Pos p;
    p.t = u >= 0 ? u / L2 : -1;
    p.x = u - p.t * L2;
    p.buf = (p.t + C::NBUF) % C::NBUF;
    return p;
}

// ---- duties, by thread id: t = tid & 15, arr = (tid >> 4) & 3, pl = (tid >> 6) & 1, col = tid >> 7 ----
// IN: the halos that edge thread (0,t) [rows] / (t,0) [columns] reads at step st for its plane pl -- array arr, four
// words -- into index 0 of edge slot rs(st) + pl; p = that edge thread's position at step st.  The column threads of the
// u2 / all(old) arrays bring the corner word along.  Called at the top of step st - 1 (the slot's last readers, step
// st - 2, are done); the copies land (cp.async wait) before the caller arrives for step st - 1.
template <class C, class SC = ScoreRtl>
TA_HD void stream2_duty_in(StreamSmem<C> &S, const StreamArgs &A, int tid, const Stream2Pos &p, int st, const SC &sc = SC())
{
    const int t = tid & 15, arr = (tid >> 4) & 3, pl = (tid >> 6) & 1, col = tid >> 7;
    const int slot = (((st - 1) & 1) << 1) + pl, i = 2 * p.x + 1 + pl;
    const StreamTask &T = S.task[p.buf];
    const uint32_t *face = col ? T.fk_r : T.fj_r;
    uint32_t *dst = col ? &S.col[slot][arr][0][4 * t] : &S.row[slot][arr][0][4 * t];
    if (face) {
        const uint32_t *row = face + (long long)i * (col ? C::FK_ROW_WORDS : C::FJ_ROW_WORDS) + arr * (col ? C::COL_STRIDE : C::ROW_STRIDE);
        ta_copy16(dst, row + 4 * t);
        // corner cell (row above the tile, last column of the block to the left): face entry j0-1, or the relay entry
        // for the block's first tile row
        if (col && (arr & 1)) ta_copy4(&S.corner[slot][arr >> 1][t][0], row + (t > 0 ? 4 * t - 1 : C::BJ));
    } else {  // zero border plane (src/TriAlign_1cyc.v:156-181): all 7 states are 0, reduced values = the potential
        const uint32_t e = rep((uint32_t)(A.base + (int)sc.pi() * (i + 4 * t + 1 - (arr >> 1)))), d = sc.pi2();
        ta_store4(dst, e, e + d, e + 2 * d, e + 3 * d);
        if (col && (arr & 1)) S.corner[slot][arr >> 1][t][0] = e - d;  // cell (j0, 0)
    }
}

// OUT: the bottom row of thread (15,t) [rows] / right column of (t,15), array arr, of the plane pl it ran at step ss --
// index 16 of edge slot ws(ss) + pl -- to the face ring of the task that thread was in (p = its position at step ss),
// re-based to the consumer's frame.  Called at the top of step ss + 1 (everybody has finished step ss; the slot is
// written again at step ss + 2).  The u2 / all(old) column threads of tile row 0 copy the relay entry along.
template <class C, class SC = ScoreRtl>
TA_HD void stream2_duty_out(StreamSmem<C> &S, int tid, const Stream2Pos &p, int ss, const SC &sc = SC())
{
    const int t = tid & 15, arr = (tid >> 4) & 3, pl = (tid >> 6) & 1, col = tid >> 7;
    const int slot = ((ss & 1) << 1) + pl, i = 2 * p.x + 1 + pl;
    const StreamTask &T = S.task[p.buf];
    uint32_t *face = col ? T.fk_w : T.fj_w;
    if (!face) return;
    const uint32_t adj = rep(sc.pi() * (col ? C::BK : C::BJ));
    const uint32_t *src = col ? &S.col[slot][arr][C::TK][4 * t] : &S.row[slot][arr][C::TJ][4 * t];
    uint32_t *row = face + (long long)i * (col ? C::FK_ROW_WORDS : C::FJ_ROW_WORDS) + arr * (col ? C::COL_STRIDE : C::ROW_STRIDE);
    ta_store4(row + 4 * t, src[0] - adj, src[1] - adj, src[2] - adj, src[3] - adj);
    if (col && t == 0 && (arr & 1)) row[C::BJ] = S.corner[slot][arr >> 1][0][C::TK] - adj;
}

// ---- the part of a step that does not depend on the neighbours: per-task reset and the constants of both planes ----
template <class C, class SC = ScoreRtl>
TA_HD void stream2_thread_prepare(StreamRegs<C> &R, StreamConsts<C> &K0, StreamConsts<C> &K1, const StreamSmem<C> &S,
                                  const StreamArgs &A, int tj, int tk, const Stream2Pos &p, const SC &sc = SC())
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK, buf = p.buf;
    if (p.x == 0) {
        // a new task enters this thread: plane 0 is the all-zero border (src/PE_1cyc.v:164-178 override), and the
        // tile's (b,c) constants change
        const uint32_t zb = stream_phi(A, 0, j0 + 1, k0 + 1, sc);
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const uint32_t z = zb + (uint32_t)(r + c) * sc.pi2();
                R.all[r][c] = z; R.wX[r][c] = z; R.u1[r][c] = z; R.u3[r][c] = z;
                R.cBC[r][c] = pair_const(neq_flag(S.bpk[buf][j0 + r], S.cpk[buf][k0 + c]), sc);
            }
    }
    const uint8_t *codes = &S.acode[S.task[buf].abuf][2 * p.x];  // planes 2x+1, 2x+2 (a padding plane reads code 0)
    const int code0 = codes[0], code1 = codes[1];
    TA_CHECK(code0 >= 0 && code0 < 16 && code1 >= 0 && code1 < 16);
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        K0.cAB[r] = S.tabAB[buf][code0][j0 + r];
        K0.tB[r] = S.tabMB[buf][code0][j0 + r];
        K1.cAB[r] = S.tabAB[buf][code1][j0 + r];
        K1.tB[r] = S.tabMB[buf][code1][j0 + r];
    }
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        K0.cAC[c] = S.tabK[buf][code0][k0 + c];
        K0.tC[c] = S.tabK[buf][code0][64 + k0 + c];
        K1.cAC[c] = S.tabK[buf][code1][k0 + c];
        K1.tC[c] = S.tabK[buf][code1][64 + k0 + c];
    }
}

// the halos of one plane of a thread's tile (see fast_batch.cuh: batch_load_halos); index 0 of the tj / tk dimension is
// the block halo, filled by stream2_duty_in
template <class C>
struct Stream2Halo {
    uint32_t T_wY[C::RK], T_u1[C::RK], T_u2[C::RK + 1], T_all[C::RK + 1], L_wZ[C::RJ], L_u3[C::RJ], L_u2[C::RJ], L_all[C::RJ];
};

template <class C>
TA_HD void stream2_load_halos(Stream2Halo<C> &H, const StreamSmem<C> &S, int tj, int tk, int rs)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        H.T_wY[c] = S.row[rs][0][tj][k0 + c];
        H.T_u1[c] = S.row[rs][2][tj][k0 + c];
        H.T_u2[c + 1] = S.row[rs][1][tj][k0 + c];
        H.T_all[c + 1] = S.row[rs][3][tj][k0 + c];
    }
    H.T_u2[0] = S.corner[rs][0][tj][tk];
    H.T_all[0] = S.corner[rs][1][tj][tk];
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        H.L_wZ[r] = S.col[rs][0][tk][j0 + r];
        H.L_u3[r] = S.col[rs][2][tk][j0 + r];
    }
#pragma unroll
    for (int r = 1; r < RJ; ++r) {
        H.L_u2[r] = S.col[rs][1][tk][j0 + r - 1];
        H.L_all[r] = S.col[rs][3][tk][j0 + r - 1];
    }
    H.L_u2[0] = H.T_u2[0];
    H.L_all[0] = H.T_all[0];
}

// one plane of one thread: relay the corner, publish the old half of the own edges, update the tile, publish the new
// half (edge slot ws).  R = plane i-1 state, Rn = plane i state.
// CHECK: the run-time range check of every plane that belongs to the triple (i <= len: the padding plane of an odd len
// is not part of it); `viol` accumulates the violations of the task this thread is working on.
template <class C, bool CHECK = false, class SC = ScoreRtl>
TA_HD void stream2_thread_plane(const StreamRegs<C> &R, StreamRegs<C> &Rn, const StreamConsts<C> &K, const Stream2Halo<C> &H,
                                StreamSmem<C> &S, const StreamArgs &A, int tj, int tk, int ws, int buf, int i,
                                uint32_t *viol_io = nullptr, const RangeCheck *chk = nullptr, const SC &sc = SC())
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
    S.corner[ws][0][tj][tk + 1] = H.T_u2[RK];
    S.corner[ws][1][tj][tk + 1] = H.T_all[RK];
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

    uint32_t n_all[RJ][RK], n_wX[RJ][RK], n_u1[RJ][RK], n_u3[RJ][RK], n_wY[RJ][RK], n_wZ[RJ][RK], n_u2[RJ][RK];
#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            const uint32_t p_all = (r > 0 && c > 0) ? R.all[r - 1][c - 1] : (r == 0 ? H.T_all[c] : H.L_all[r]);
            const uint32_t p_wY = r > 0 ? n_wY[r - 1][c] : H.T_wY[c];
            const uint32_t p_wZ = c > 0 ? n_wZ[r][c - 1] : H.L_wZ[r];
            const uint32_t p_u1 = r > 0 ? R.u1[r - 1][c] : H.T_u1[c];
            const uint32_t p_u2 = (r > 0 && c > 0) ? n_u2[r - 1][c - 1] : (r == 0 ? H.T_u2[c] : H.L_u2[r]);
            const uint32_t p_u3 = c > 0 ? R.u3[r][c - 1] : H.L_u3[r];
            const uint32_t cM = triple_const_of(K.tB[r], K.tC[c], K.cAB[r], R.cBC[r][c], K.cAC[c], sc);
            const CellOut o = cell_update<false, SC>(p_all, R.wX[r][c], p_wY, p_wZ, p_u1, p_u2, p_u3, cM, K.cAB[r],
                                                     R.cBC[r][c], K.cAC[c], nullptr, sc);
            n_all[r][c] = o.all; n_wX[r][c] = o.wX; n_u1[r][c] = o.u1; n_u3[r][c] = o.u3;
            n_wY[r][c] = o.wY; n_wZ[r][c] = o.wZ; n_u2[r][c] = o.u2;
        }

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

    // Score = max of the 7 states of cell (len,len,len) (src/TriAlign_1cyc.v:141-142) = `all` of that cell
    if (i == A.len) {
        const int jl = A.len - (A.nbj - 1) * C::BJ, kl = A.len - (A.nbk - 1) * C::BK;  // 1..64, in the last block
        if (tj == (jl - 1) / RJ && tk == (kl - 1) / RK) {
            const StreamTask &T = S.task[buf];
            if (T.final_here) {
                const int rf = (jl - 1) % RJ, cf = (kl - 1) % RK;
                uint32_t v = 0;
#pragma unroll
                for (int r = 0; r < RJ; ++r)
#pragma unroll
                    for (int c = 0; c < RK; ++c)
                        if (r == rf && c == cf) v = n_all[r][c];
                const int off = A.base + (int)sc.pi() * (A.len + jl + kl);
                const long long t0 = 2 * T.pair;
                if (t0 < A.n) A.scores[t0] = (int32_t)(v & 0xFFFFu) - off;
                if (t0 + 1 < A.n) A.scores[t0 + 1] = (int32_t)(v >> 16) - off;
            }
        }
    }

    if (CHECK) {
        // RTL constants: every 8th plane is sampled (`all` moves by at most +3 / -4 per plane; the margins of the limits
        // cover the planes in between); run-time scoring: every plane, limits exact per cell
        uint32_t viol = i == 1 ? 0u : *viol_io;  // a new task enters this thread
        if ((SC::kRtl ? (i & 7) == 0 : true) && i <= A.len) {
            uint32_t mx = n_all[0][0], mn = n_all[0][0];
#pragma unroll
            for (int r = 0; r < RJ; ++r)
#pragma unroll
                for (int c = 0; c < RK; c += 2) {
                    const uint32_t v0 = SC::kRtl ? n_all[r][c] : fadd(n_all[r][c], 0u - (uint32_t)(r + c) * sc.pi2());
                    const uint32_t v1 = SC::kRtl ? n_all[r][c + 1] : fadd(n_all[r][c + 1], 0u - (uint32_t)(r + c + 1) * sc.pi2());
                    mx = vmax3(mx, v0, v1);
                    mn = vmin3(mn, v0, v1);
                }
            const int ph = A.base + (int)sc.pi() * (i + j0 + 1 + k0 + 1), slack = SC::kRtl ? 2 * (RJ + RK - 2) : 0;
            const int h = ph + chk->check_hi, l = ph + chk->check_lo + slack;
            const uint32_t hi_lim = rep((uint32_t)(h > 0xFFFF ? 0xFFFF : h)), lo_lim = rep((uint32_t)(l < 0 ? 0 : l));
            viol |= (vmax2(mx, hi_lim) ^ hi_lim) | (vmin2(mn, lo_lim) ^ lo_lim);
        }
        if (i == A.len && viol != 0) {  // the task's last real plane leaves this thread: report
            const StreamTask &T = S.task[buf];
            if (T.valid) {
                const int bits = ((viol & 0xFFFFu) ? 1 : 0) | ((viol >> 16) ? 2 : 0);
#if defined(__CUDA_ARCH__)
                atomicOr(chk->pair_flags + T.pair, bits);
#else
                chk->pair_flags[T.pair] |= bits;
#endif
            }
        }
        *viol_io = viol;
    }

#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            Rn.all[r][c] = n_all[r][c]; Rn.wX[r][c] = n_wX[r][c]; Rn.u1[r][c] = n_u1[r][c]; Rn.u3[r][c] = n_u3[r][c];
            Rn.cBC[r][c] = R.cBC[r][c];
        }
}

// one step of one thread: both planes' halos first (published a step ago), then the two plane updates back to back,
// R -> R2 -> R
template <class C, bool CHECK = false, class SC = ScoreRtl>
TA_HD void stream2_thread_step(StreamRegs<C> &R, StreamRegs<C> &R2, const StreamConsts<C> &K0, const StreamConsts<C> &K1,
                               StreamSmem<C> &S, const StreamArgs &A, int tj, int tk, int s, const Stream2Pos &p,
                               uint32_t *viol_io = nullptr, const RangeCheck *chk = nullptr, const SC &sc = SC())
{
    const int rs = ((s - 1) & 1) << 1, ws = (s & 1) << 1, i0 = 2 * p.x + 1;
    Stream2Halo<C> H0, H1;
    stream2_load_halos<C>(H0, S, tj, tk, rs);
    stream2_load_halos<C>(H1, S, tj, tk, rs + 1);
    stream2_thread_plane<C, CHECK, SC>(R, R2, K0, H0, S, A, tj, tk, ws, p.buf, i0, viol_io, chk, sc);
    stream2_thread_plane<C, CHECK, SC>(R2, R, K1, H1, S, A, tj, tk, ws + 1, p.buf, i0 + 1, viol_io, chk, sc);
}

// CTA-uniform actions of the global step that is step g_off of task g_t for thread (0,0): stage task g_t + 1
template <class C, class SC = ScoreRtl>
TA_HD void stream2_uniform_actions(StreamSmem<C> &S, const StreamArgs &A, int tid, int cta, int P, long long pair_base,
                                   int g_t, int g_off, const SC &sc = SC())
{
    if (g_off == C::LOAD_AT) stream_stage_load<C>(S, A, tid, cta, P, pair_base, g_t + 1);
    if (g_off == C::BUILD_AT) stream_stage_build<C, SC>(S, tid, g_t + 1, sc);
}

// global steps a CTA runs for P pairs: steps 0 .. this value
template <class C>
TA_HD long long stream2_total_steps(const StreamArgs &A, int P)
{
    return (long long)P * A.nbj * A.nbk * stream2_steps_per_task(A) + C::SKEW;
}

}  // namespace fast
}  // namespace trialign
