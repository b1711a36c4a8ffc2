// fast_sweep.cuh -- K2: block-sweep kernel for long triples (the GPU form of the RTL's slicing method,
// src/TriAlign_1cyc.v:127-140,289-335: the (j,k) plane is cut into blocks, every block is swept over all of A, and the
// faces between blocks are parked in memory -- face SRAMs there, HBM/L2 here).
//
//   * a task = one 64x64 (j,k) block per 16-bit half, swept over i = 1..la by the same 16x16-thread systolic array
//     as K1 (thread (tj,tk) owns a 4x4 tile, state of plane i-1 in registers, wavefront skew tj+tk, edge exchange
//     through shared memory, one __syncthreads per step);
//   * LONG mode (one long triple): the two halves of a word are two k-ADJACENT blocks; the high half runs 16 planes
//     behind the low half, which makes thread (tj,tk)'s high half the logical thread (tj,tk+16) of a 16x32 array:
//     its left halo is the low half's right column of thread (tj,15), published one step earlier;
//   * BATCH mode (many triples of equal length): the two halves are two triples, as in K1;
//   * block faces live in global memory as PACKED words, exactly as the producing edge thread holds them, so the
//     consuming edge thread of the next block copies them straight into the shared-memory halo slots with cp.async
//     one step ahead of use (the RTL's z-face ping-pong and y-face ring, TriAlign_1cyc.v:157-181);
//   * LONG mode: tasks are claimed in (k-block-pair, j-block) order from a global counter and synchronise through
//     per-task progress counters (rows of face data completely written), release/acquire at GPU scope;
//   * potentials are block-local in j,k and absolute in i (see fast_core.cuh); a face value is re-based when it is
//     written (minus 2*block extent), so readers use it as is.
#pragma once

#include "fast_core.cuh"

namespace trialign {
namespace fast {

template <bool LONG_>
struct SweepCfg {
    static constexpr bool LONG = LONG_;
    static constexpr int RJ = 4, RK = 4, TJ = 16, TK = 16;
    static constexpr int BJ = RJ * TJ, BK = RK * TK, NT = TJ * TK;
    static constexpr int LAG = LONG ? TK : 0;          // planes the high half runs behind the low half
    static constexpr int DEPTH = TJ + TK - 2;          // largest physical skew
    static constexpr int KSPAN = LONG ? 2 * BK : BK;   // columns of C one task covers
    static constexpr int ROW_W = BK;                   // words per face-J row and value
    static constexpr int COL_W = BJ + 4;               // words per face-K row and value (index BJ = relay, padded)
    static constexpr int TAB_STRIDE = 68;              // j-indexed tables: lanes of a wavefront differ in the code only
    static constexpr int TABK_STRIDE = 128;            // k-indexed tables {cAC[64], tC[64]}: bank group depends on tk only
    static constexpr int CORNER_STRIDE = TK + 8;       // a warp spans 4 tj x 8 tk: distinct banks
    static constexpr int LA_MAX = 8192 + 64;           // A symbols kept in shared memory
    static constexpr uint32_t HI_OFF = LONG ? ((uint32_t)(2 * (BK - LAG)) << 16) : 0u;  // phi(high) - phi(low)
    static constexpr int FJ_ROW_WORDS = 4 * ROW_W;     // one plane of a J face
    static constexpr int FK_ROW_WORDS = 4 * COL_W;     // one plane of a K face
};

// Everything a CTA needs to know about one task (CTA-uniform, lives in shared memory).
struct SweepTask {
    int valid;
    int bj, bkp;                  // block row, block column (LONG: column PAIR)
    long long pair;               // BATCH: index of the pair of triples
    uint32_t *fj_w, *fk_w;        // faces this task writes (nullptr: nobody needs them)
    const uint32_t *fj_r, *fk_r;  // faces this task reads  (nullptr: zero border plane)
    int top_task, left_task;      // LONG: tasks whose progress gates fj_r / fk_r (-1: none)
    int self_task;                // LONG: index into progress[]
    int final_here;               // this task holds cell (lb,lc)
};

struct SweepArgs {
    const uint8_t *a, *b, *c;     // BATCH: [n][la],[n][lb],[n][lc]   LONG: one triple
    long long n;
    int la, lb, lc;
    int nbj, nbkp;                // blocks per problem: rows, columns (LONG: column pairs)
    long long tasks_total;        // LONG: nbj*nbkp
    long long pairs_total;        // BATCH: ceil(n/2)
    uint32_t *face_j, *face_k;    // ring storage
    long long fj_slot_words, fk_slot_words;
    int fj_ring, fk_ring;         // LONG: global ring sizes; BATCH: per-CTA ring sizes
    int gate;                     // LONG: steps between two progress releases in steady state (power of two)
    int rgate;                    // K2L column pipeline: steps between two system-scope releases to the neighbour GPU
    int gate_tid;                 // K2L: the thread that polls / releases the progress counters
    int early;                    // K2L: 1 = fine publication rhythm during a task's first steps (experiments)
    int pingpong;                 // K2L: middle phase unrolled by two over two alternating register sets
    int *progress;                // LONG: [tasks_total] indexed by task id bj + nbj*bkp, zero-initialised
    const int *order;             // LONG: claim position -> task id (sorted by earliest possible start, sweep_order_key)
    const int *pos_of;            // LONG: task id -> claim position
    unsigned long long *counter;  // LONG: task counter, zero-initialised
    // i-slab pipeline (LONG, several GPUs): this launch computes planes i_begin+1 .. i_begin+la of the triple (a already
    // points at symbol i_begin).  The plane-i_begin state of every task comes from the previous slab (iface_in, on
    // this GPU) and the plane-(i_begin+la) state goes to the next slab's GPU (iface_out, possibly a peer pointer).
    // Layout: [task][all,wX,u1,u3][cell 0..15][thread 0..255] packed words, re-based so that the reader sees them as
    // its plane 0.  Flags: one int per task, 1 = state complete (release/acquire at system scope).
    const uint32_t *iface_in;
    uint32_t *iface_out;
    const int *iface_in_flag;
    int *iface_out_flag;
    // column pipeline (K2L, several GPUs): this launch sweeps a contiguous range of block-column pairs of the triple (c
    // already points at the range's first column, lc = its columns).  The K faces crossing the range's left edge come
    // from the left neighbour GPU (cface_in, cprog_in: on this GPU, written by the peer), those crossing its right edge go
    // to the right neighbour (cface_out, cprog_out: peer pointers): one face slot and one progress word per block row.
    // Progress words are tagged (epoch << 20 | steps, 0xFFFFF = done), so that runs repeat without clearing.
    const uint32_t *cface_in;
    const int *cprog_in;
    uint32_t *cface_out;
    int *cprog_out;
    int last_slab;                // emit the score (the launch holds cell (la,lb,lc) of the whole triple)
    int epoch;                    // value the hand-over flags take in this run (lets runs repeat without clearing)
    int32_t *out;                 // BATCH: scores[n]; LONG: out[0] = score
    int32_t *out_states;          // LONG: 7 final-cell states (or nullptr)
    int base;                     // potential base
    int check_lo, check_hi;       // range check (CHECK kernels): every sampled `all` must lie in [check_lo, check_hi]
    int *viol_flag;               // set non-zero when the range check fails
};

constexpr int kProgressDone = 1 << 30;

template <class C>
struct SweepSmem {
    uint32_t row[2][4][C::TJ + 1][C::BK];     // [slot][wY,u2,u1(old),all(old)][tj+1 (0 = block halo)][k]
    uint32_t col[2][4][C::TK + 1][C::COL_W];  // [slot][wZ,u2,u3(old),all(old)][tk+1 (0 = block halo)][j (+relay)]
    uint32_t corner[2][2][C::TJ][C::CORNER_STRIDE];  // [slot][u2,all(old)][tj][tk]  relay of the cell above-right of a tile
    uint32_t cornerL[2][2][C::TJ];            // [slot][u2,all(old)][tj]  corner cell of the tiles on the block's left edge
    uint32_t tabAB[16][C::TAB_STRIDE];        // [a code][j]  pair constant {4,2}
    uint32_t tabMB[16][C::TAB_STRIDE];        // [a code][j]  triple row constant {9,3}
    uint32_t tabK[16][C::TABK_STRIDE];        // [a code][k] pair constant {4,2}, [a code][64 + k] triple column constant {9,6}
    uint32_t tabBC[C::BJ][C::BK];             // [j][k]  (b,c) pair constant {4,2} of every cell of the block: read every step
                                              // instead of living in 16 registers (the LONG kernel is register-bound)
    uint8_t seqA[2][C::LA_MAX];               // [half][i]   (LONG: only [0] is used)
    uint8_t seqB[2][64], seqC[2][64];         // [half][pos in block]
    SweepTask task;
};

template <class C>
struct SweepRegs {
    uint32_t all[C::RJ][C::RK], wX[C::RJ][C::RK], u1[C::RJ][C::RK], u3[C::RJ][C::RK];  // plane i-1
    uint32_t viol;                                                                     // range-check violations seen
};

// Constants of the step a thread is about to run (the a symbols of both halves against the tile's rows / columns),
// fetched one step ahead, before the barrier (sweep_thread_prepare).
template <class C>
struct SweepConsts {
    uint32_t cAB[C::RJ], tB[C::RJ], cAC[C::RK], tC[C::RK];
};

template <class C>
TA_HD void sweep_thread_pos(int tid, int &tj, int &tk)
{
    const int lane = tid & 31, warp = tid >> 5;
    tk = (lane & 7) + 8 * (warp & 1);
    tj = (lane >> 3) + 4 * (warp >> 1);
}

// potential of a cell in block-local coordinates; the high half is the same expression for (i - LAG, kl + BK) in LONG
// mode, which is a constant offset
template <class C>
TA_HD uint32_t sphi(const SweepArgs &A, int i0, int jl, int kl)
{
    return rep((uint32_t)(A.base + 2 * (i0 + jl + kl))) + C::HI_OFF;
}

// ---------------------------------------------------------------------------------------------------------------
// Task geometry
// ---------------------------------------------------------------------------------------------------------------
template <class C>
TA_HD int sweep_task_steps(const SweepArgs &A) { return A.la + C::LAG + C::DEPTH; }  // steps 0 .. this value

// LONG mode: tasks are claimed in the order of their earliest possible start: block (bj,bkp) can start about
// (TJ+gate) steps after (bj-1,bkp) and (TK+LAG+gate) steps after (bj,bkp-1), so with short tasks (i-slabs, L~1024) a
// column-major order would leave most CTAs waiting on a chain.  Both predecessors have a smaller key.
TA_HD int sweep_order_key(int bj, int bkp) { return 3 * bj + 5 * bkp; }

// claim position -> descriptor.  Face ring slots are indexed by the claim position.
template <class C>
TA_HD void sweep_describe_long(const SweepArgs &A, long long pos, SweepTask &T)
{
    const int t = A.order[pos];
    T.valid = 1;
    T.pair = 0;
    T.bkp = t / A.nbj;
    T.bj = t % A.nbj;
    T.self_task = t;
    T.top_task = T.bj > 0 ? t - 1 : -1;
    T.left_task = T.bkp > 0 ? t - A.nbj : -1;
    T.fj_w = (T.bj + 1 < A.nbj) ? A.face_j + (pos % A.fj_ring) * A.fj_slot_words : nullptr;
    T.fk_w = (T.bkp + 1 < A.nbkp) ? A.face_k + (pos % A.fk_ring) * A.fk_slot_words : nullptr;
    T.fj_r = T.bj > 0 ? A.face_j + (A.pos_of[t - 1] % A.fj_ring) * A.fj_slot_words : nullptr;
    T.fk_r = T.bkp > 0 ? A.face_k + (A.pos_of[t - A.nbj] % A.fk_ring) * A.fk_slot_words : nullptr;
    T.final_here = (T.bj == (A.lb - 1) / C::BJ) && (T.bkp == (A.lc - 1) / C::KSPAN);
}

// BATCH mode: the q-th task of this CTA's current pair; faces are private to the CTA (region `cta`).
template <class C>
TA_HD void sweep_describe_batch(const SweepArgs &A, int cta, long long pair, int q, SweepTask &T)
{
    T.valid = 1;
    T.pair = pair;
    T.bkp = q / A.nbj;
    T.bj = q % A.nbj;
    T.self_task = T.top_task = T.left_task = -1;
    uint32_t *fj = A.face_j + (long long)cta * A.fj_ring * A.fj_slot_words;
    uint32_t *fk = A.face_k + (long long)cta * A.fk_ring * A.fk_slot_words;
    T.fj_w = (T.bj + 1 < A.nbj) ? fj + (q % A.fj_ring) * A.fj_slot_words : nullptr;
    T.fk_w = (T.bkp + 1 < A.nbkp) ? fk + (q % A.fk_ring) * A.fk_slot_words : nullptr;
    T.fj_r = T.bj > 0 ? fj + ((q - 1) % A.fj_ring) * A.fj_slot_words : nullptr;
    T.fk_r = T.bkp > 0 ? fk + ((q - A.nbj) % A.fk_ring) * A.fk_slot_words : nullptr;
    T.final_here = (T.bj == (A.lb - 1) / C::BJ) && (T.bkp == (A.lc - 1) / C::KSPAN);
}

// ---------------------------------------------------------------------------------------------------------------
// Task setup (CTA-uniform; barriers between the phases are the caller's)
// ---------------------------------------------------------------------------------------------------------------
template <class C>
TA_HD void sweep_load_a(SweepSmem<C> &S, const SweepArgs &A, int tid, long long pair)
{
    for (int h = 0; h < (C::LONG ? 1 : 2); ++h) {
        const long long t = C::LONG ? 0 : 2 * pair + h;
        for (int e = tid; e < A.la; e += C::NT) S.seqA[h][e] = (t < A.n) ? (A.a[t * A.la + e] & 3) : 0;
    }
}

template <class C>
TA_HD void sweep_load_bc(SweepSmem<C> &S, const SweepArgs &A, int tid)
{
    const SweepTask &T = S.task;
    for (int e = tid; e < 2 * 2 * 64; e += C::NT) {
        const int which = e >> 7, h = (e >> 6) & 1, pos = e & 63;
        const long long t = C::LONG ? 0 : 2 * T.pair + h;
        // cells beyond lb / lc are padding: codes 4 and 5 match nothing (not even each other), so padded cells
        // never create values above their real neighbours (keeps the run-time range check free of false alarms)
        if (which == 0) {  // B rows of this block (LONG: both halves share them)
            const int j = T.bj * C::BJ + pos;
            uint8_t v = 4;
            if (t < A.n && j < A.lb) v = A.b[t * A.lb + j] & 3;
            S.seqB[h][pos] = v;
        } else {           // C columns (LONG: high half is the next 64 columns)
            const int k = T.bkp * C::KSPAN + (C::LONG ? h * C::BK : 0) + pos;
            uint8_t v = 5;
            if (t < A.n && k < A.lc) v = A.c[t * A.lc + k] & 3;
            S.seqC[h][pos] = v;
        }
    }
}

template <class C>
TA_HD void sweep_build_tables(SweepSmem<C> &S, int tid)
{
    for (int e = tid; e < 16 * 64; e += C::NT) {
        const int code = e >> 6, p = e & 63;
        const uint32_t ap = (uint32_t)(code & 3) | ((uint32_t)(code >> 2) << 16);
        const uint32_t bp = (uint32_t)S.seqB[0][p] | ((uint32_t)S.seqB[1][p] << 16);
        const uint32_t cp = (uint32_t)S.seqC[0][p] | ((uint32_t)S.seqC[1][p] << 16);
        const uint32_t nab = neq_flag(ap, bp), nac = neq_flag(ap, cp);
        S.tabAB[code][p] = pair_const(nab);
        S.tabMB[code][p] = triple_row_const(nab);
        S.tabK[code][p] = pair_const(nac);
        S.tabK[code][64 + p] = triple_col_const(nac);
    }
}

template <class C>
TA_HD void sweep_init_thread(SweepRegs<C> &R, SweepSmem<C> &S, int tj, int tk)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const int j0 = tj * RJ, k0 = tk * RK;
#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            const uint32_t bp = (uint32_t)S.seqB[0][j0 + r] | ((uint32_t)S.seqB[1][j0 + r] << 16);
            const uint32_t cp = (uint32_t)S.seqC[0][k0 + c] | ((uint32_t)S.seqC[1][k0 + c] << 16);
            const uint32_t nbc = neq_flag(bp, cp);
            S.tabBC[j0 + r][k0 + c] = pair_const(nbc);
            R.all[r][c] = R.wX[r][c] = R.u1[r][c] = R.u3[r][c] = 0;
        }
    R.viol = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Block-edge halo traffic.  Called at step s (s = -1 in the task prologue) for the data thread (0,tk) / (tj,0) will
// consume at step s+1; writes only the block-halo entries (index 0) of slot (s & 1).
// Global faces are copied with TA_COPY16 / TA_COPY4 (cp.async on the device, plain copies in the host emulation).
// ---------------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
TA_HD void ta_copy16(uint32_t *dst_smem, const uint32_t *src_gmem)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_gmem) : "memory");
}
TA_HD void ta_copy4(uint32_t *dst_smem, const uint32_t *src_gmem)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src_gmem) : "memory");
}
TA_HD void ta_copy_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
TA_HD uint32_t ta_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
#else
TA_HD void ta_copy16(uint32_t *d, const uint32_t *s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3]; }
TA_HD void ta_copy4(uint32_t *d, const uint32_t *s) { d[0] = s[0]; }
TA_HD void ta_copy_wait() {}
TA_HD uint32_t ta_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int n = 0; n < 4; ++n) r |= (uint32_t)((v >> (8 * ((sel >> (4 * n)) & 7))) & 0xFF) << (8 * n);
    return r;
}
#endif

// 16-byte store of four packed words (dst is 16-byte aligned: face rows are 64 / 68 words, tiles start at multiples of 4)
#if defined(__CUDA_ARCH__)
TA_HD void ta_store4(uint32_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    *reinterpret_cast<uint4 *>(dst) = make_uint4(a, b, c, d);
}
#else
TA_HD void ta_store4(uint32_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}
#endif

// low half <- high half of x ; high half <- low half of y
TA_HD uint32_t hi_lo(uint32_t x, uint32_t y) { return ta_prmt(x, y, 0x5432); }

template <class C>
TA_HD void sweep_edge_fetch(SweepSmem<C> &S, const SweepArgs &A, int tj, int tk, int s)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const SweepTask &T = S.task;
    const int ws = s & 1;
    const int rows = A.la + C::LAG;  // face rows 1..rows exist
    if (tj == 0) {
        const int i = s + 2 - tk;  // low-half plane of thread (0,tk) at step s+1
        const int k0 = tk * RK;
        if (T.fj_r) {
            if (i >= 1 && i <= rows) {
                const uint32_t *src = T.fj_r + (long long)i * C::FJ_ROW_WORDS + k0;
#pragma unroll
                for (int v = 0; v < 4; ++v) ta_copy16(&S.row[ws][v][0][k0], src + v * C::ROW_W);
            }
        } else {  // j = 0 border plane: all 7 states are 0 (src/TriAlign_1cyc.v:156-165), reduced values = phi
            const uint32_t e = sphi<C>(A, i, 0, k0 + 1);  // phi is linear: one evaluation, carry-free offsets
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const uint32_t now = e + (uint32_t)c * 0x00020002u, old = now - 0x00020002u;
                S.row[ws][0][0][k0 + c] = now;
                S.row[ws][1][0][k0 + c] = now;
                S.row[ws][2][0][k0 + c] = old;
                S.row[ws][3][0][k0 + c] = old;
            }
        }
    }
    if (tk == 0) {
        const int i = s + 2 - tj;  // low-half plane of thread (tj,0) at step s+1
        const int j0 = tj * RJ;
        if (T.fk_r) {
            // LONG: the producer's high half is LAG planes behind its row index; beyond the last plane the dead low half
            // keeps consuming the last row, so that it stays bounded while the high half is live (fast_long.cuh)
            int src_row = i + C::LAG;
            if (C::LONG && src_row > rows) src_row = rows;
            if (src_row >= 1 && src_row <= rows) {
                const uint32_t *src = T.fk_r + (long long)src_row * C::FK_ROW_WORDS + j0;
#pragma unroll
                for (int v = 0; v < 4; ++v) ta_copy16(&S.col[ws][v][0][j0], src + v * C::COL_W);
                // corner cell (row above the tile, column k0 of the block) for THIS thread's plane: face row entry j0-1,
                // or the relay entry for the block's first tile row
                const int cidx = tj > 0 ? j0 - 1 : C::BJ;
                ta_copy4(&S.cornerL[ws][0][tj], T.fk_r + (long long)src_row * C::FK_ROW_WORDS + 1 * C::COL_W + cidx);
                ta_copy4(&S.cornerL[ws][1][tj], T.fk_r + (long long)src_row * C::FK_ROW_WORDS + 3 * C::COL_W + cidx);
            }
        } else {  // k = 0 border plane (src/TriAlign_1cyc.v:166-174).  LONG: the consumer takes the HIGH half.
            const int sh = C::LONG ? 16 : 0;
#pragma unroll
            for (int r = 0; r < RJ; ++r) {
                const uint32_t now = (uint32_t)(A.base + 2 * (i + j0 + r + 1)), old = now - 2;
                S.col[ws][0][0][j0 + r] = C::LONG ? (now << sh) : rep(now);
                S.col[ws][1][0][j0 + r] = C::LONG ? (now << sh) : rep(now);
                S.col[ws][2][0][j0 + r] = C::LONG ? (old << sh) : rep(old);
                S.col[ws][3][0][j0 + r] = C::LONG ? (old << sh) : rep(old);
            }
            {  // corner cell (j0, 0) lies on the border plane too
                const uint32_t now = (uint32_t)(A.base + 2 * (i + j0)), old = now - 2;
                S.cornerL[ws][0][tj] = C::LONG ? (now << sh) : rep(now);
                S.cornerL[ws][1][tj] = C::LONG ? (old << sh) : rep(old);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The part of step s that does not depend on the neighbours' step s-1: plane-0 reset and the step's constants.
// Runs at the end of step s-1, before its barrier, so that the table look-ups are off the critical path.
// ---------------------------------------------------------------------------------------------------------------
template <class C>
TA_HD void sweep_thread_prepare(SweepRegs<C> &R, SweepConsts<C> &K, const SweepSmem<C> &S, const SweepArgs &A, int tj,
                                int tk, int s, int tid)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const SweepTask &T = S.task;
    const int i0 = s - (tj + tk) + 1;  // low-half plane (valid 1..la)
    const int i1 = i0 - C::LAG;        // high-half plane
    const int j0 = tj * RJ, k0 = tk * RK;

    // plane 0 is the all-zero border (PE override, src/PE_1cyc.v:164-178): reset the half whose plane 1 starts now
    if (i0 == 1 || i1 == 1) {
        const uint32_t keep = (i0 == 1 ? 0u : 0x0000FFFFu) | (i1 == 1 ? 0u : 0xFFFF0000u);
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                // sphi(i0-1,...) is phi(plane 0) in whichever half starts now: the low half starts at i0 = 1, the
                // high half at i0 = 1 + LAG (its own plane index is i0 - LAG)
                const uint32_t zz = sphi<C>(A, i0 - 1, j0 + r + 1, k0 + c + 1) & ~keep;
                if (C::LONG && A.iface_in) {  // continue from the previous slab instead of the zero border
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

    // per-step constants: the a symbols of both halves against the tile's rows / columns
    int ia0 = i0 - 1, ia1 = i1 - 1;
    ia0 = ia0 < 0 ? 0 : (ia0 >= A.la ? A.la - 1 : ia0);
    ia1 = ia1 < 0 ? 0 : (ia1 >= A.la ? A.la - 1 : ia1);
    const int code = S.seqA[0][ia0] | (S.seqA[C::LONG ? 0 : 1][ia1] << 2);
    TA_CHECK(code >= 0 && code < 16);
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
}

// ---------------------------------------------------------------------------------------------------------------
// One wavefront step of one thread
// ---------------------------------------------------------------------------------------------------------------
template <class C, bool CHECK>
TA_HD void sweep_thread_step(SweepRegs<C> &R, const SweepConsts<C> &K, SweepSmem<C> &S, const SweepArgs &A, int tj,
                             int tk, int s, int tid)
{
    constexpr int RJ = C::RJ, RK = C::RK;
    const SweepTask &T = S.task;
    const int i0 = s - (tj + tk) + 1;  // low-half plane (valid 1..la)
    const int i1 = i0 - C::LAG;        // high-half plane
    const int j0 = tj * RJ, k0 = tk * RK;
    const int ws = s & 1, rs = ws ^ 1;

    // halos (see fast_batch.cuh); index 0 of the first dimension of row/col is the block halo
    uint32_t T_wY[RK], T_u1[RK], T_u2[RK + 1], T_all[RK + 1], L_wZ[RJ], L_u3[RJ], L_u2[RJ], L_all[RJ];
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        T_wY[c] = S.row[rs][0][tj][k0 + c];
        T_u1[c] = S.row[rs][2][tj][k0 + c];
        T_u2[c + 1] = S.row[rs][1][tj][k0 + c];
        T_all[c + 1] = S.row[rs][3][tj][k0 + c];
    }
    if (tk > 0) {
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
        T_u2[0] = S.corner[rs][0][tj][tk - 1];
        T_all[0] = S.corner[rs][1][tj][tk - 1];
    } else {
        // left edge of the block.  BATCH: the block halo column as is.  LONG: low half <- high half of the block halo
        // (the previous column pair's high block), high half <- low half of thread (tj,15)'s column (this pair's
        // low block, one logical thread to the left).
        uint32_t cu2 = S.cornerL[rs][0][tj], call = S.cornerL[rs][1][tj];
        if (C::LONG) {
            cu2 = hi_lo(cu2, S.corner[rs][0][tj][C::TK - 1]);
            call = hi_lo(call, S.corner[rs][1][tj][C::TK - 1]);
        }
        T_u2[0] = cu2;
        T_all[0] = call;
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
            uint32_t wz = S.col[rs][0][0][j0 + r], u3v = S.col[rs][2][0][j0 + r];
            if (C::LONG) {
                wz = hi_lo(wz, S.col[rs][0][C::TK][j0 + r]);
                u3v = hi_lo(u3v, S.col[rs][2][C::TK][j0 + r]);
            }
            L_wZ[r] = wz;
            L_u3[r] = u3v;
        }
#pragma unroll
        for (int r = 1; r < RJ; ++r) {
            uint32_t u2v = S.col[rs][1][0][j0 + r - 1], av = S.col[rs][3][0][j0 + r - 1];
            if (C::LONG) {
                u2v = hi_lo(u2v, S.col[rs][1][C::TK][j0 + r - 1]);
                av = hi_lo(av, S.col[rs][3][C::TK][j0 + r - 1]);
            }
            L_u2[r] = u2v;
            L_all[r] = av;
        }
    }
    L_u2[0] = T_u2[0];
    L_all[0] = T_all[0];
    S.corner[ws][0][tj][tk] = T_u2[RK];
    S.corner[ws][1][tj][tk] = T_all[RK];
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

    // publish the tile's bottom row and right column for the neighbours' next step
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
    const int rows = A.la + C::LAG;
    if (tj == C::TJ - 1 && T.fj_w && i0 >= 1 && i0 <= rows) {
        const uint32_t adj = rep(2 * C::BJ);
        uint32_t *dst = T.fj_w + (long long)i0 * C::FJ_ROW_WORDS + k0;
        ta_store4(dst + 0 * C::ROW_W, n_wY[RJ - 1][0] - adj, n_wY[RJ - 1][1] - adj, n_wY[RJ - 1][2] - adj, n_wY[RJ - 1][3] - adj);
        ta_store4(dst + 1 * C::ROW_W, n_u2[RJ - 1][0] - adj, n_u2[RJ - 1][1] - adj, n_u2[RJ - 1][2] - adj, n_u2[RJ - 1][3] - adj);
        ta_store4(dst + 2 * C::ROW_W, R.u1[RJ - 1][0] - adj, R.u1[RJ - 1][1] - adj, R.u1[RJ - 1][2] - adj, R.u1[RJ - 1][3] - adj);
        ta_store4(dst + 3 * C::ROW_W, R.all[RJ - 1][0] - adj, R.all[RJ - 1][1] - adj, R.all[RJ - 1][2] - adj, R.all[RJ - 1][3] - adj);
    }
    if (tk == C::TK - 1 && T.fk_w && i0 >= 1 && i0 <= rows) {
        // LONG: only the high half crosses to the next column pair (KSPAN columns to the right of this pair's origin)
        const uint32_t adj = C::LONG ? ((uint32_t)(2 * C::KSPAN) << 16) : rep(2 * C::BK);
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

    // Run-time wrap check (used when wrap-freedom at score_bits cannot be proved a priori).  `all` moves by at most
    // +3 / -4 from one plane to the next (a cell's `all` is at most 3 above the best `all` of plane i-1 and at least
    // all(i-1,j,k)-4), so sampling every 16th plane with a 45 / 60 margin in check_hi / check_lo bounds every plane.
    if (CHECK && (i0 & 15) == 0) {
        uint32_t mx = n_all[0][0], mn = n_all[0][0];
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; c += 2) {
                mx = vmax3(mx, n_all[r][c], n_all[r][c + 1]);
                mn = vmin3(mn, n_all[r][c], n_all[r][c + 1]);
            }
        // true value = stored - phi(cell); phi grows by 2 per row / column inside the tile (at most 12)
        // (limits are formed per half in scalar arithmetic: check_lo is negative, a packed add would carry)
        const uint32_t ph = sphi<C>(A, i0, j0 + 1, k0 + 1);
        const int p0 = (int)(ph & 0xFFFFu), p1 = (int)(ph >> 16), slack = 2 * (RJ + RK - 2);
        const int h0 = p0 + A.check_hi, h1 = p1 + A.check_hi, l0 = p0 + A.check_lo + slack, l1 = p1 + A.check_lo + slack;
        const uint32_t hi_lim = (uint32_t)(h0 > 0xFFFF ? 0xFFFF : h0) | ((uint32_t)(h1 > 0xFFFF ? 0xFFFF : h1) << 16);
        const uint32_t lo_lim = (uint32_t)(l0 < 0 ? 0 : l0) | ((uint32_t)(l1 < 0 ? 0 : l1) << 16);
        const uint32_t bad = (vmax2(mx, hi_lim) ^ hi_lim) | (vmin2(mn, lo_lim) ^ lo_lim);
        const uint32_t live = ((i0 >= 1 && i0 <= A.la) ? 0x0000FFFFu : 0u) | ((i1 >= 1 && i1 <= A.la) ? 0xFFFF0000u : 0u);
        R.viol |= bad & live;
    }

    // Score = max of the 7 states of cell (la,lb,lc) (src/TriAlign_1cyc.v:141-142) = `all` of that cell
    if (T.final_here && (!C::LONG || A.last_slab)) {
        const int jl = A.lb - T.bj * C::BJ;                 // 1..64
        const int kk = A.lc - T.bkp * C::KSPAN;             // 1..KSPAN
        const int hf = C::LONG ? (kk - 1) / C::BK : 0;      // LONG: the half that holds column lc
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
            // potential of the cell in each half's own frame
            const int off_lo = A.base + 2 * (A.la + jl + kl);
            const int off_hi = off_lo + (C::LONG ? 2 * C::BK : 0);  // high half: kl + BK, plane la (iplane already = la)
            if (C::LONG) {
                const int off = hf ? off_hi : off_lo;
                const int shv = hf ? 16 : 0;
                A.out[0] = (int32_t)((v >> shv) & 0xFFFFu) - off;
                if (A.out_states)
                    for (int q = 0; q < 7; ++q) A.out_states[q] = (int32_t)((st[q] >> shv) & 0xFFFFu) - off;
            } else {
                const long long t0 = 2 * T.pair;
                if (t0 < A.n) A.out[t0] = (int32_t)(v & 0xFFFFu) - off_lo;
                if (t0 + 1 < A.n) A.out[t0 + 1] = (int32_t)(v >> 16) - off_lo;
            }
        }
    }

#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            R.all[r][c] = n_all[r][c]; R.wX[r][c] = n_wX[r][c]; R.u1[r][c] = n_u1[r][c]; R.u3[r][c] = n_u3[r][c];
        }

    // i-slab hand-over: the half that has just finished the slab's last plane stores its state for the next slab,
    // re-based by the slab's 2*la so that the reader's plane index restarts at 0 (16-bit stores: the halves finish
    // LAG steps apart)
    if (C::LONG && A.iface_out && (i0 == A.la || i1 == A.la)) {
        const uint32_t adj = rep((uint32_t)(2 * A.la));
        uint16_t *dst16 = reinterpret_cast<uint16_t *>(A.iface_out + (((long long)T.self_task * 4) * 16) * C::NT + tid);
        const int hsel = (i0 == A.la) ? 0 : 1;
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const long long w = (long long)(r * RK + c) * C::NT;  // word offset of this cell within a value plane
                const uint32_t v0 = R.all[r][c] - adj, v1 = R.wX[r][c] - adj, v2 = R.u1[r][c] - adj, v3 = R.u3[r][c] - adj;
                dst16[2 * (w + 0LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v0 >> 16 : v0);
                dst16[2 * (w + 1LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v1 >> 16 : v1);
                dst16[2 * (w + 2LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v2 >> 16 : v2);
                dst16[2 * (w + 3LL * 16 * C::NT) + hsel] = (uint16_t)(hsel ? v3 >> 16 : v3);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LONG mode dependency tracking (evaluated by one thread).  progress[t] = number of wavefront steps task t has completed
// AND published (kProgressDone once it is finished).  Writers and readers of a face are skewed alike: thread (0,tk)
// fetches at its step f the row that producer thread (15,tk) wrote at producer step f+16, and thread (tj,0) fetches the
// row that producer thread (tj,15) wrote at step f+16+LAG.
//   consumer: sweep_certified() turns the producers' progress into the last fetch step that is safe; the kernel samples
//     the counters only when its next fetch is not certified yet, whatever the producers' publication rhythm;
//   producer: sweep_publish_now() -- every 2 steps during the first kEarlySteps steps of a task (tasks waiting to start
//     need only 16 / 32 rows: the pipeline fill per block hop is 18 / 34 steps), every A.gate steps afterwards (a
//     release is a MEMBAR.GPU that holds one warp, and hence the next CTA barrier, for hundreds of cycles).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kEarlySteps = 64, kEarlyGate = 2;
constexpr int kCertifiedAll = 1 << 30;

template <class C>
TA_HD bool sweep_publish_now(const SweepArgs &A, int s)
{
    const int gate = (s < kEarlySteps && A.gate > kEarlyGate) ? kEarlyGate : A.gate;
    return ((s + 1) & (gate - 1)) == 0;
}

// last fetch step (the fetch at step f brings the halos of step f+1; the prologue fetch is f = -1) whose face rows are
// complete, given the progress of the task above (top) and of the task to the left
template <class C>
TA_HD int sweep_certified(const SweepArgs &A, const SweepTask &T, int top_prog, int left_prog)
{
    const int total = sweep_task_steps<C>(A) + 1;  // steps 0..steps
    int cert = kCertifiedAll;
    if (T.top_task >= 0 && top_prog < total) cert = top_prog - (C::TJ + 1);
    if (T.left_task >= 0 && left_prog < total) {
        const int c2 = left_prog - (C::TK + C::LAG + 1);
        cert = c2 < cert ? c2 : cert;
    }
    return cert;
}

}  // namespace fast
}  // namespace trialign
