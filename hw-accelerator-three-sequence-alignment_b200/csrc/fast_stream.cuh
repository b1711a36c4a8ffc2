// fast_stream.cuh -- K2S: the streaming block-sweep kernel for batches of triples longer than 64 (BASELINE config 4).
//
// The RTL's slicing method (src/TriAlign_1cyc.v:127-140,289-335: the (j,k) plane is cut into P x P blocks, every block
// is swept over all of A, block faces are parked in the z/y face SRAMs) on K1's CTA-wide systolic array:
//   * a task = one 64x64 (j,k) block of one PAIR of triples (one triple per 16-bit half), swept over i = 1..len;
//     thread (tj,tk) owns a 4x4 tile for all i and runs plane i of task t at global step
//         s = t*len + (i-1) + LAG*(tj+tk)
//     -- the tasks of a CTA (all blocks of its pairs, j-blocks fastest) stream through the array BACK TO BACK, so the
//     LAG*(TJ+TK-2)-step fill/drain is paid once per CTA, not once per block (a per-block kernel pays it per
//     block: 287 steps for 256 planes), and there is no per-task set-up phase: the next task's descriptor, (b,c)
//     symbols and score tables are staged into a third buffer while two tasks are in flight;
//   * K1's lagged barrier (LAG 2): tile edges are consumed two steps after publication, so a warp only waits for
//     "everybody finished step s-2";
//   * block faces live in CTA-private rings in global memory as PACKED words in the consumer's frame.  Only the edge
//     thread (0,tk) / (tj,0) ever touches the block-halo entries (index 0) of the edge slots: it fetches them with
//     cp.async TWO steps ahead of use, or writes the zero-border potential there (src/TriAlign_1cyc.v:156-181), so the
//     step itself has no border branches at all;
//   * a block's J face (bottom rows) is read by the next task, its K face (right columns) nbj tasks later (the RTL's
//     z-face ping-pong and y-face ring, TriAlign_1cyc.v:157-181).
// Exactness domain and potentials: fast_core.cuh (block-local j,k; absolute i; faces re-based when written).
#pragma once

#include "sweep_common.cuh"  // task descriptor, cp.async / 128-bit store helpers

namespace trialign {
namespace fast {

#if defined(__CUDA_ARCH__)
TA_HD void ta_copy_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// all but the most recent group of this thread are complete
TA_HD void ta_copy_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
#else
TA_HD void ta_copy_commit() {}
TA_HD void ta_copy_wait_prev() {}
#endif

// DIST = false (default): the block-edge duties (face fetch / border halos, face stores) belong to the edge threads, a
// warp is a 4x8 patch of tiles.
// DIST = true (option k2s=dist, A/B form: what made K1 16 % faster makes K2S 2 % slower -- its step never had border
// branches, and the duty bookkeeping now runs in every warp): the duties are dealt to ALL threads by thread id -- one 128-bit group per thread and step: warps 0-3 move
// the halos of the step two ahead into the edge slots (face rows with cp.async, or the zero-border potential), warps
// 4-7 copy the bottom rows / right columns published two steps ago from the edge slots to the face rings -- so the step
// itself has no edge code at all and every warp runs the same instruction stream (the step ends at a CTA barrier: the
// warp with the most duties paced all eight); with that the lanes of a warp can lie along anti-diagonals as in K1
// (fast_batch.cuh: thread_pos), which needs the column stride 72.
template <bool DIST_>
struct StreamCfgT {
    static constexpr bool DIST = DIST_;
    static constexpr int RJ = 4, RK = 4, TJ = 16, TK = 16, LAG = 2;
    static constexpr int NSLOT = 2 * LAG;   // edge buffers: written at step s, read at step s+LAG
    static constexpr int NBUF = 3;          // task buffers: two tasks in flight + one being staged
    static constexpr int BJ = RJ * TJ, BK = RK * TK, NT = TJ * TK;
    static constexpr int ROW_STRIDE = BK, COL_STRIDE = DIST ? BJ + 8 : BJ + 4, CORNER_STRIDE = TK + 8;
    static constexpr int TABJ_STRIDE = 68, TABK_STRIDE = 128;
    static constexpr int SKEW = LAG * (TJ + TK - 2);
    static constexpr int LEN_MIN = SKEW + 1;   // at most two tasks in flight; staging fits inside a task
    static constexpr int LEN_MAX = 2048;       // a-codes of two pairs kept in shared memory
    static constexpr int LOAD_AT = 8, BUILD_AT = 16;  // steps after a task enters thread (0,0): stage the next one
    static constexpr int FJ_ROW_WORDS = 4 * ROW_STRIDE;  // one plane of a J face: {wY,u2,u1(old),all(old)} x 64 columns
    static constexpr int FK_ROW_WORDS = 4 * COL_STRIDE;  // one plane of a K face: the same of 64 rows + relay entry
    static_assert(BUILD_AT + 8 < LEN_MIN, "staging schedule");
};
using StreamCfg = StreamCfgT<false>;
using StreamCfgD = StreamCfgT<true>;

struct StreamTask {
    int valid;                    // 0: filler (pipeline fill / drain): border halos, no faces, no score
    int bj, bk;
    int abuf;                     // a-code buffer of the task's pair
    int final_here;               // this block holds cell (len,len)
    long long pair;               // global pair index
    uint32_t *fj_w, *fk_w;        // faces this task writes (nullptr: nobody needs them)
    const uint32_t *fj_r, *fk_r;  // faces this task reads  (nullptr: zero border plane)
};

struct StreamArgs {
    const uint8_t *a, *b, *c;     // [n][len] device
    long long n;
    int len;
    int nbj, nbk;                 // blocks per side
    long long pairs_total;        // ceil(n/2)
    int pairs_per_cta;            // contiguous chunk per CTA
    uint32_t *face_j, *face_k;    // per-CTA rings: 2 J slots, fk_ring K slots
    long long fj_slot_words, fk_slot_words;
    int fk_ring;                  // nbj + 1
    int32_t *scores;              // [n]
    int base;                     // potential base
};

// Range-checked variant (CHECK kernels), as in fast_batch.cuh: sampled `all` in [check_lo, check_hi], else
// pair_flags[pair] gets bit 0 / 1 set for the low / high triple.  A separate kernel parameter: the unchecked kernel's
// parameter block -- and with it ptxas's register allocation (223 registers; 194 and 3 % slower with three more words
// in StreamArgs) -- stays exactly as it was.
struct RangeCheck {
    int check_lo, check_hi;
    int *pair_flags;
};

constexpr int kStreamCheckEvery = 16;  // steps between two range samples (`all` moves by at most +3 / -4 per plane)

template <class C>
struct StreamSmem {
    uint32_t row[C::NSLOT][4][C::TJ + 1][C::ROW_STRIDE];  // [slot][wY,u2,u1(old),all(old)][tj+1 (0 = block halo)][k]
    uint32_t col[C::NSLOT][4][C::TK + 1][C::COL_STRIDE];  // [slot][wZ,u2,u3(old),all(old)][tk+1 (0 = block halo)][j]
    uint32_t corner[C::NSLOT][2][C::TJ][C::CORNER_STRIDE];  // [slot][u2,all(old)][tj][tk+1 (0 = block halo)]
    uint32_t tabAB[C::NBUF][16][C::TABJ_STRIDE];  // [buf][a code][j]  pair constant {4,2}
    uint32_t tabMB[C::NBUF][16][C::TABJ_STRIDE];  // [buf][a code][j]  triple row constant {9,3}
    uint32_t tabK[C::NBUF][16][C::TABK_STRIDE];   // [buf][a code][k] pair constant, [64 + k] triple column constant {9,6}
    uint32_t bpk[C::NBUF][64], cpk[C::NBUF][64];  // [buf][pos]  b / c symbols of the block, one triple per half
    uint8_t acode[C::NBUF][C::LEN_MAX];           // [pair % NBUF][i]  a_lo | a_hi << 2 (a pair may be a single task)
    StreamTask task[C::NBUF];
};

template <class C>
struct StreamRegs {
    uint32_t all[C::RJ][C::RK], wX[C::RJ][C::RK], u1[C::RJ][C::RK], u3[C::RJ][C::RK];  // plane i-1
    uint32_t cBC[C::RJ][C::RK];                                                        // per-cell (b,c) constant
};

template <class C>
struct StreamConsts {
    uint32_t cAB[C::RJ], tB[C::RJ], cAC[C::RK], tC[C::RK];
};

// where a thread is in its stream of tasks
struct StreamPos {
    int t;    // CTA-local task index (-1 / tasks: fillers)
    int i;    // plane 1..len
    int buf;  // (t + NBUF) % NBUF
};

// Thread id -> tile.  A warp is a 4 (tj) x 8 (tk) patch.  Edge duty: patches of the first / last patch row fetch /
// store the J face (8 lanes each), every patch has 4 lanes on the left (fetch K) or right (store K) block edge;
// hardware warp w issues from scheduler w % 4, so the patches are dealt such that every scheduler gets one row duty
// and both column duties: {TL, R}, {TR, L}, {BL, R}, {BR, L}.
template <class C>
TA_HD void stream_thread_pos(int tid, int &tj, int &tk)
{
    if (C::DIST) {  // eight consecutive lanes = eight consecutive tj of one wrapped anti-diagonal (see fast_batch.cuh)
        const int oct = tid >> 3, dd = oct >> 1;
        tj = 8 * (oct & 1) + (tid & 7);
        tk = (dd - tj) & 15;
        return;
    }
    const int lane = tid & 31;
    const int patch = (0x45237610u >> (4 * (tid >> 5))) & 7;  // warps 0..7 -> patches 0,1,6,7,3,2,5,4
    tk = (lane & 7) + 8 * (patch & 1);
    tj = (lane >> 3) + 4 * (patch >> 1);
}

template <class SC = ScoreRtl>
TA_HD uint32_t stream_phi(const StreamArgs &A, int i, int jl, int kl, const SC &sc = SC())
{
    return rep((uint32_t)(A.base + (int)sc.pi() * (i + jl + kl)));
}

TA_HD void stream_advance(const StreamArgs &A, StreamPos &p, int nbuf)
{
    if (++p.i > A.len) {
        p.i = 1;
        ++p.t;
        p.buf = p.buf + 1 == nbuf ? 0 : p.buf + 1;
    }
}

// position of thread (tj,tk) at global step 0
template <class C>
TA_HD StreamPos stream_start_pos(const StreamArgs &A, int tj, int tk)
{
    StreamPos p;
    const int back = C::LAG * (tj + tk);  // < len
    p.t = back ? -1 : 0;
    p.i = back ? A.len - back + 1 : 1;
    p.buf = back ? C::NBUF - 1 : 0;
    return p;
}

// ---- staging of task t (CTA-local index) into buffer (t + NBUF) % NBUF; P = pairs of this CTA ----
template <class C>
TA_HD void stream_stage_load(StreamSmem<C> &S, const StreamArgs &A, int tid, int cta, int P, long long pair_base, int t)
{
    const int nblk = A.nbj * A.nbk;
    const int buf = (t + C::NBUF) % C::NBUF;
    const bool valid = t >= 0 && t < P * nblk;
    const int p = valid ? t / nblk : 0, q = valid ? t % nblk : 0;
    const int bk = q / A.nbj, bj = q % A.nbj;
    const long long pair = pair_base + p;
    if (tid == 0) {
        StreamTask &T = S.task[buf];
        T.valid = valid ? 1 : 0;
        T.bj = bj; T.bk = bk;
        T.pair = pair;
        T.abuf = valid ? (p % C::NBUF) : 0;
        T.final_here = valid && q == nblk - 1;
        uint32_t *fj = A.face_j + (long long)cta * 2 * A.fj_slot_words;
        uint32_t *fk = A.face_k + (long long)cta * A.fk_ring * A.fk_slot_words;
        T.fj_w = (valid && bj + 1 < A.nbj) ? fj + (t & 1) * A.fj_slot_words : nullptr;
        T.fk_w = (valid && bk + 1 < A.nbk) ? fk + (t % A.fk_ring) * A.fk_slot_words : nullptr;
        T.fj_r = (valid && bj > 0) ? fj + ((t - 1) & 1) * A.fj_slot_words : nullptr;
        T.fk_r = (valid && bk > 0) ? fk + ((t - A.nbj) % A.fk_ring) * A.fk_slot_words : nullptr;
    }
    // (b,c) symbols of the block, one triple per half.  Cells beyond len are padding: codes 4 and 5 match nothing (not
    // even each other), so padded cells never feed a real cell anything a real neighbour would not.
    for (int e = tid; e < 2 * 64; e += C::NT) {
        const int which = e >> 6, pos = e & 63;
        const int x = (which == 0 ? bj : bk) * 64 + pos;
        const uint8_t *src = which == 0 ? A.b : A.c;
        uint32_t v[2];
        for (int h = 0; h < 2; ++h) {
            const long long tr = 2 * pair + h;
            v[h] = 4 + which;
            if (!valid) v[h] = 0;
            else if (x < A.len) v[h] = (tr < A.n) ? (src[tr * A.len + x] & 3) : 0;  // PE keeps bits [1:0] (src/PE_1cyc.v:63-66)
        }
        (which == 0 ? S.bpk[buf] : S.cpk[buf])[pos] = v[0] | (v[1] << 16);
    }
    if (valid && q == 0) {  // a new pair starts: its a symbols, both halves combined
        const long long t0 = 2 * pair, t1 = 2 * pair + 1;
        for (int e = tid; e < A.len; e += C::NT) {
            const uint32_t a0 = (t0 < A.n) ? (A.a[t0 * A.len + e] & 3) : 0;
            const uint32_t a1 = (t1 < A.n) ? (A.a[t1 * A.len + e] & 3) : 0;
            S.acode[p % C::NBUF][e] = (uint8_t)(a0 | (a1 << 2));
        }
        // the two-planes-per-step kernel (fast_stream2.cuh) runs one padding plane behind an odd len
        if (tid == 0 && A.len < C::LEN_MAX) S.acode[p % C::NBUF][A.len] = 0;
    }
}

template <class C, class SC = ScoreRtl>
TA_HD void stream_stage_build(StreamSmem<C> &S, int tid, int t, const SC &sc = SC())
{
    const int buf = (t + C::NBUF) % C::NBUF;
    for (int e = tid; e < 16 * 64; e += C::NT) {
        const int code = e >> 6, p = e & 63;
        const uint32_t ap = (uint32_t)(code & 3) | ((uint32_t)(code >> 2) << 16);
        const uint32_t nab = neq_flag(ap, S.bpk[buf][p]), nac = neq_flag(ap, S.cpk[buf][p]);
        S.tabAB[buf][code][p] = pair_const(nab, sc);
        S.tabMB[buf][code][p] = triple_row_const(nab, sc);
        S.tabK[buf][code][p] = pair_const(nac, sc);
        S.tabK[buf][code][64 + p] = triple_col_const(nac, sc);
    }
}

// CTA-uniform actions of global step s = g_t*len + g_off (every thread tracks g_t / g_off in registers)
template <class C, class SC = ScoreRtl>
TA_HD void stream_uniform_actions(StreamSmem<C> &S, const StreamArgs &A, int tid, int cta, int P, long long pair_base,
                                  int g_t, int g_off, const SC &sc = SC())
{
    if (g_off == C::LOAD_AT) stream_stage_load<C>(S, A, tid, cta, P, pair_base, g_t + 1);
    if (g_off == C::BUILD_AT) stream_stage_build<C, SC>(S, tid, g_t + 1, sc);
}

// ---- block-edge halos of the step that runs plane p.i of task p.t, into edge slot `slot` (the slot that step reads).
// Only thread (0,tk) / (tj,0) writes and reads these entries, so no barrier is involved: cp.async + wait_group. ----
template <class C, class SC = ScoreRtl>
TA_HD void stream_edge_fetch(StreamSmem<C> &S, const StreamArgs &A, int tj, int tk, const StreamPos &p, int slot,
                             const SC &sc = SC())
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const StreamTask &T = S.task[p.buf];
    const int i = p.i;
    if (tj == 0) {
        const int k0 = tk * RK;
        const uint32_t *fj = T.fj_r;
        if (fj) {
            const uint32_t *src = fj + (long long)i * C::FJ_ROW_WORDS + k0;
#pragma unroll
            for (int v = 0; v < 4; ++v) ta_copy16(&S.row[slot][v][0][k0], src + v * C::ROW_STRIDE);
        } else {  // j = 0 border plane: all 7 states are 0 (src/TriAlign_1cyc.v:156-165), reduced values = phi
            const uint32_t e = stream_phi(A, i, 0, k0 + 1, sc);
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const uint32_t now = e + (uint32_t)c * sc.pi2(), old = now - sc.pi2();
                S.row[slot][0][0][k0 + c] = now;
                S.row[slot][1][0][k0 + c] = now;
                S.row[slot][2][0][k0 + c] = old;
                S.row[slot][3][0][k0 + c] = old;
            }
        }
    }
    if (tk == 0) {
        const int j0 = tj * RJ;
        const uint32_t *fk = T.fk_r;
        if (fk) {
            const uint32_t *src = fk + (long long)i * C::FK_ROW_WORDS + j0;
#pragma unroll
            for (int v = 0; v < 4; ++v) ta_copy16(&S.col[slot][v][0][j0], src + v * C::COL_STRIDE);
            // corner cell (row above the tile, last column of the block to the left): face entry j0-1, or the relay
            // entry for the block's first tile row
            const int cidx = tj > 0 ? j0 - 1 : C::BJ;
            ta_copy4(&S.corner[slot][0][tj][0], fk + (long long)i * C::FK_ROW_WORDS + 1 * C::COL_STRIDE + cidx);
            ta_copy4(&S.corner[slot][1][tj][0], fk + (long long)i * C::FK_ROW_WORDS + 3 * C::COL_STRIDE + cidx);
        } else {  // k = 0 border plane (src/TriAlign_1cyc.v:166-174); the corner cell (j0, 0) lies on it too
            const uint32_t e = stream_phi(A, i, j0, 0, sc);
            S.corner[slot][0][tj][0] = e;
            S.corner[slot][1][tj][0] = e - sc.pi2();
#pragma unroll
            for (int r = 0; r < RJ; ++r) {
                const uint32_t now = e + (uint32_t)(r + 1) * sc.pi2(), old = now - sc.pi2();
                S.col[slot][0][0][j0 + r] = now;
                S.col[slot][1][0][j0 + r] = now;
                S.col[slot][2][0][j0 + r] = old;
                S.col[slot][3][0][j0 + r] = old;
            }
        }
    }
}

// ---- the part of a step that does not depend on the neighbours: move on to the next plane, per-task reset, the
// step's constants, and the halo fetch of the step after it.  s1 = the global step being prepared. ----
template <class C, class SC = ScoreRtl>
TA_HD void stream_thread_prepare(StreamRegs<C> &R, StreamConsts<C> &K, StreamSmem<C> &S, const StreamArgs &A, int tj,
                                 int tk, StreamPos &p, int s1, bool advance, const SC &sc = SC())
{
    constexpr int RJ = C::RJ, RK = C::RK;
    if (advance) stream_advance(A, p, C::NBUF);
    const int j0 = tj * RJ, k0 = tk * RK;
    const int buf = p.buf;
    if (p.i == 1) {
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
    const int code = S.acode[S.task[buf].abuf][p.i - 1];
    TA_CHECK(code >= 0 && code < 16);
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        K.cAB[r] = S.tabAB[buf][code][j0 + r];
        K.tB[r] = S.tabMB[buf][code][j0 + r];
    }
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        K.cAC[c] = S.tabK[buf][code][k0 + c];
        K.tC[c] = S.tabK[buf][code][64 + k0 + c];
    }
    // block-edge halos of step s1 + 1 (two steps ahead of the step that has just run)
    if (!C::DIST) {
        if (tj == 0 || tk == 0) {
            StreamPos q = p;
            stream_advance(A, q, C::NBUF);
            stream_edge_fetch<C, SC>(S, A, tj, tk, q, (s1 + 1 + C::LAG) & (C::NSLOT - 1), sc);
        }
        ta_copy_commit();
    }
}

// ---- DIST: the block-edge duties by thread id ----
// position of a thread with skew `back` (= LAG * (tj + tk)) at global step x >= back - len
template <class C>
TA_HD StreamPos stream_pos_at(const StreamArgs &A, int back, int x)
{
    StreamPos p;
    const int u = x - back;
    p.t = u >= 0 ? u / A.len : -1;
    p.i = u - p.t * A.len + 1;
    p.buf = (p.t + C::NBUF) % C::NBUF;
    return p;
}

// Threads 0..127: the halos that the edge thread (0,t) [rows, tid < 64] / (t,0) [columns] reads at step r -- array
// arr = (tid >> 4) & 3, its four words -- into index 0 of edge slot (r - LAG); p = that edge thread's position at step r.
// The column threads of the u2 / all(old) arrays bring the corner word along.  Called at the top of step r - 2: the
// slot's last readers (step r - 4) are known to be done, and the caller's copies land (cp.async wait) before it
// arrives for step r - 2, which every reader of step r waits for.
template <class C, class SC = ScoreRtl>
TA_HD void stream_duty_fetch(StreamSmem<C> &S, const StreamArgs &A, int tid, const StreamPos &p, int r, const SC &sc = SC())
{
    const int t = tid & 15, arr = (tid >> 4) & 3, col = tid >> 6, slot = (r - C::LAG) & (C::NSLOT - 1);
    const StreamTask &T = S.task[p.buf];
    const uint32_t *face = col ? T.fk_r : T.fj_r;
    uint32_t *dst = col ? &S.col[slot][arr][0][4 * t] : &S.row[slot][arr][0][4 * t];
    if (face) {
        const uint32_t *row = face + (long long)p.i * (col ? C::FK_ROW_WORDS : C::FJ_ROW_WORDS) + arr * (col ? C::COL_STRIDE : C::ROW_STRIDE);
        ta_copy16(dst, row + 4 * t);
        // corner cell (row above the tile, last column of the block to the left): face entry j0-1, or the relay entry
        // for the block's first tile row
        if (col && (arr & 1)) ta_copy4(&S.corner[slot][arr >> 1][t][0], row + (t > 0 ? 4 * t - 1 : C::BJ));
    } else {  // zero border plane (src/TriAlign_1cyc.v:156-181): all 7 states are 0, reduced values = the potential
        const uint32_t e = rep((uint32_t)(A.base + (int)sc.pi() * (p.i + 4 * t + 1 - (arr >> 1)))), d = sc.pi2();
        ta_store4(dst, e, e + d, e + 2 * d, e + 3 * d);
        if (col && (arr & 1)) S.corner[slot][arr >> 1][t][0] = e - d;  // cell (j0, 0): one column before the tile's first
    }
}

// Threads 128..255: the bottom row of thread (15,t) [tid < 192] / right column of (t,15), array arr, published at step
// q into index 16 of edge slot q & 3, to the face ring of the task that thread was in (p = its position at step q),
// re-based to the consumer's frame.  Called at the top of step q + 2 (everybody has finished step q; the slot is
// written again at step q + 4).  The u2 / all(old) column threads of tile row 0 copy the relay entry along.
template <class C, class SC = ScoreRtl>
TA_HD void stream_duty_store(StreamSmem<C> &S, const StreamArgs &A, int tid, const StreamPos &p, int q, const SC &sc = SC())
{
    (void)A;
    const int t = tid & 15, arr = (tid >> 4) & 3, col = (tid >> 6) & 1, slot = q & (C::NSLOT - 1);
    const StreamTask &T = S.task[p.buf];
    uint32_t *face = col ? T.fk_w : T.fj_w;
    if (!face) return;
    const uint32_t adj = rep(sc.pi() * (col ? C::BK : C::BJ));
    const uint32_t *src = col ? &S.col[slot][arr][C::TK][4 * t] : &S.row[slot][arr][C::TJ][4 * t];
    uint32_t *row = face + (long long)p.i * (col ? C::FK_ROW_WORDS : C::FJ_ROW_WORDS) + arr * (col ? C::COL_STRIDE : C::ROW_STRIDE);
    ta_store4(row + 4 * t, src[0] - adj, src[1] - adj, src[2] - adj, src[3] - adj);
    if (col && t == 0 && (arr & 1)) row[C::BJ] = S.corner[slot][arr >> 1][0][C::TK] - adj;
}

// ---- one wavefront step of one thread: halos published at step s-LAG -> tile -> publish for step s+LAG ----
// R = plane i-1 state (input), Rn = plane i state (output; may be R itself).  The kernel alternates two register sets
// (R -> Rn, Rn -> R) in a step loop unrolled by two, so that the loop-carried state needs no register moves.
// CHECK: the run-time range check (see StreamArgs), sampled every kStreamCheckEvery-th step (warp-uniform).
// `viol` (CHECK only) accumulates the violations of the task this thread is working on; it lives outside StreamRegs so
// that the unchecked kernel compiles exactly as without the check.
template <class C, bool CHECK = false, class SC = ScoreRtl>
TA_HD void stream_thread_step(const StreamRegs<C> &R, StreamRegs<C> &Rn, const StreamConsts<C> &K, StreamSmem<C> &S,
                              const StreamArgs &A, int tj, int tk, int s, const StreamPos &p, uint32_t *viol_io = nullptr,
                              const RangeCheck *chk = nullptr, const SC &sc = SC())
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
    const int rs = (s - C::LAG) & (C::NSLOT - 1), ws = s & (C::NSLOT - 1);

    // halos (see fast_batch.cuh); index 0 of the tj / tk dimension is the block halo, filled by stream_edge_fetch
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
    L_u2[0] = T_u2[0];
    L_all[0] = T_all[0];
    S.corner[ws][0][tj][tk + 1] = T_u2[RK];
    S.corner[ws][1][tj][tk + 1] = T_all[RK];
    // the plane i-1 half of this thread's own edges is already known: these stores leave before the tile
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

    // publish the plane-i half of the tile's bottom row and right column for the neighbours' step s+LAG
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
    if (!C::DIST && (tj == C::TJ - 1 || tk == C::TK - 1)) {
        const StreamTask &T = S.task[p.buf];
        if (tj == C::TJ - 1) {
            uint32_t *fj = T.fj_w;
            if (fj) {
                const uint32_t adj = rep(sc.pi() * C::BJ);
                uint32_t *dst = fj + (long long)p.i * C::FJ_ROW_WORDS + k0;
                ta_store4(dst + 0 * C::ROW_STRIDE, n_wY[RJ - 1][0] - adj, n_wY[RJ - 1][1] - adj, n_wY[RJ - 1][2] - adj, n_wY[RJ - 1][3] - adj);
                ta_store4(dst + 1 * C::ROW_STRIDE, n_u2[RJ - 1][0] - adj, n_u2[RJ - 1][1] - adj, n_u2[RJ - 1][2] - adj, n_u2[RJ - 1][3] - adj);
                ta_store4(dst + 2 * C::ROW_STRIDE, R.u1[RJ - 1][0] - adj, R.u1[RJ - 1][1] - adj, R.u1[RJ - 1][2] - adj, R.u1[RJ - 1][3] - adj);
                ta_store4(dst + 3 * C::ROW_STRIDE, R.all[RJ - 1][0] - adj, R.all[RJ - 1][1] - adj, R.all[RJ - 1][2] - adj, R.all[RJ - 1][3] - adj);
            }
        }
        if (tk == C::TK - 1) {
            uint32_t *fk = T.fk_w;
            if (fk) {
                const uint32_t adj = rep(sc.pi() * C::BK);
                uint32_t *dst = fk + (long long)p.i * C::FK_ROW_WORDS + j0;
                ta_store4(dst + 0 * C::COL_STRIDE, n_wZ[0][RK - 1] - adj, n_wZ[1][RK - 1] - adj, n_wZ[2][RK - 1] - adj, n_wZ[3][RK - 1] - adj);
                ta_store4(dst + 1 * C::COL_STRIDE, n_u2[0][RK - 1] - adj, n_u2[1][RK - 1] - adj, n_u2[2][RK - 1] - adj, n_u2[3][RK - 1] - adj);
                ta_store4(dst + 2 * C::COL_STRIDE, R.u3[0][RK - 1] - adj, R.u3[1][RK - 1] - adj, R.u3[2][RK - 1] - adj, R.u3[3][RK - 1] - adj);
                ta_store4(dst + 3 * C::COL_STRIDE, R.all[0][RK - 1] - adj, R.all[1][RK - 1] - adj, R.all[2][RK - 1] - adj, R.all[3][RK - 1] - adj);
                if (tj == 0) {  // relay: the cell above-right of the block = corner of the next column's first tile row
                    uint32_t *rel = fk + (long long)p.i * C::FK_ROW_WORDS + C::BJ;
                    rel[1 * C::COL_STRIDE] = T_u2[RK] - adj;
                    rel[3 * C::COL_STRIDE] = T_all[RK] - adj;
                }
            }
        }
    }

    if (CHECK) {
        uint32_t viol = p.i == 1 ? 0u : *viol_io;  // a new task enters this thread
        // run-time scoring: every step is sampled (the +3 / -4 per plane bound belongs to the RTL's constants) and the
        // cell's own share of the potential is taken off first, so that the limits are exact per cell
        if (SC::kRtl ? (s & (kStreamCheckEvery - 1)) == 0 : true) {
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
            const int ph = A.base + (int)sc.pi() * (p.i + j0 + 1 + k0 + 1), slack = SC::kRtl ? 2 * (RJ + RK - 2) : 0;
            const int h = ph + chk->check_hi, l = ph + chk->check_lo + slack;
            const uint32_t hi_lim = rep((uint32_t)(h > 0xFFFF ? 0xFFFF : h)), lo_lim = rep((uint32_t)(l < 0 ? 0 : l));
            viol |= (vmax2(mx, hi_lim) ^ hi_lim) | (vmin2(mn, lo_lim) ^ lo_lim);
        }
        if (p.i == A.len && viol != 0) {  // the task leaves this thread: report
            const StreamTask &T = S.task[p.buf];
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

    // Score = max of the 7 states of cell (len,len,len) (src/TriAlign_1cyc.v:141-142) = `all` of that cell
    if (p.i == A.len) {
        const int jl = A.len - (A.nbj - 1) * C::BJ, kl = A.len - (A.nbk - 1) * C::BK;  // 1..64, in the last block
        if (tj == (jl - 1) / RJ && tk == (kl - 1) / RK) {
            const StreamTask &T = S.task[p.buf];
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

#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            Rn.all[r][c] = n_all[r][c]; Rn.wX[r][c] = n_wX[r][c]; Rn.u1[r][c] = n_u1[r][c]; Rn.u3[r][c] = n_u3[r][c];
            Rn.cBC[r][c] = R.cBC[r][c];
        }
}

// global steps a CTA runs for P pairs: steps 0 .. this value
template <class C>
TA_HD long long stream_total_steps(const StreamArgs &A, int P) { return (long long)P * A.nbj * A.nbk * A.len + C::SKEW; }

}  // namespace fast
}  // namespace trialign
