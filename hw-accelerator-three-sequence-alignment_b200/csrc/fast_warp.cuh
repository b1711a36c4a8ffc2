// fast_warp.cuh -- K1w: the batched kernel for short triples (len <= 64) as eight ASYNCHRONOUS warp-wide systolic arrays.
//
// Same arithmetic, tiles and potentials as K1 (fast_batch.cuh / fast_core.cuh); what changes is the synchronisation:
//   * a warp = 8 x 4 lanes x (4 x 4 cells) = one 32 x 16 block of the (j,k) plane; lane (lj,lk) works on plane i at
//     the warp's own step w = i + lj + lk.  Neighbouring lanes exchange tile edges with warp shuffles at the end of
//     every step -- no shared memory, no barrier inside a warp;
//   * the 2 x 4 warps of a CTA cover the 64 x 64 plane.  Warp (a,b) receives the bottom rows of warp (a-1,b) and the
//     right columns of warp (a,b-1) through shared-memory RINGS (depth kRing producer steps) with produced / consumed
//     counters, so there is no __syncthreads at all: a warp that is late (per-pair reset, table rebuild, edge work)
//     delays only its consumers, and only once the ring slack is used up;
//   * every warp keeps private per-pair score tables for its own 32 rows / 16 columns and rebuilds them itself.
// The RTL analogue is unchanged: lanes are PEs, shuffles are the PE-to-PE wires, the rings play the role of the face
// SRAM between slices (src/TriAlign_1cyc.v:127-140), only here the "slices" run concurrently.
#pragma once

#include "fast_core.cuh"

namespace trialign {
namespace fast {

struct WarpCfg {
    static constexpr int RJ = 4, RK = 4;      // cells per lane
    static constexpr int LJ = 8, LK = 4;      // lanes per warp
    static constexpr int WA = 2, WB = 4;      // warps per CTA
    static constexpr int WJ = RJ * LJ, WK = RK * LK;  // 32 x 16 cells per warp
    static constexpr int NW = WA * WB, NT = NW * 32;
    static constexpr int L = 64;              // planes per pair
    static constexpr int LANE_SKEW = LJ + LK - 2;
    static constexpr int kRing = 16;          // ring depth in producer steps
    static constexpr int TOP_WORDS = 4 * LK * RK;          // [val][lk][c]
    static constexpr int LEFT_WORDS = 4 * LJ * RJ + 2 * LJ;  // [val][lj][r] + relay [2][lj]
    static constexpr int TABJ = WJ + 4, TABK = WK + 4;     // padded table rows
    // warp (a,b) consumes what its top / left producer published LJ-1 / LK-1 steps "in the future" of its own counter
    static constexpr int TOP_LEAD = LJ - 1, LEFT_LEAD = LK - 1;
    static constexpr int TAIL = LANE_SKEW + (WA - 1) * LJ + (WB - 1) * LK;  // extra steps so that every consumer is fed
};

struct WarpArgs {
    const uint8_t *a, *b, *c;  // [n][len]
    long long n;
    int len;
    int32_t *scores;
    long long pairs_total;
    int pairs_per_cta;
    int base;
};

struct WarpShared {  // per warp
    uint32_t tabAB[2][16][WarpCfg::TABJ];
    uint32_t tabMB[2][16][WarpCfg::TABJ];
    uint32_t tabAC[2][16][WarpCfg::TABK];
    uint8_t seqB[2][2][WarpCfg::WJ], seqC[2][2][WarpCfg::WK];  // [buf][half][pos]
    uint8_t acode[2][64];
};

struct WarpSmem {
    uint32_t top[WarpCfg::WA][WarpCfg::WB][WarpCfg::kRing][WarpCfg::TOP_WORDS];    // ring read by warp (a,b), a >= 1
    uint32_t left[WarpCfg::WA][WarpCfg::WB][WarpCfg::kRing][WarpCfg::LEFT_WORDS];  // ring read by warp (a,b), b >= 1
    int produced[WarpCfg::WA][WarpCfg::WB];       // steps completed (both rings written)
    int consumed_top[WarpCfg::WA][WarpCfg::WB];   // top-ring entries consumed by warp (a,b): entries < this are free
    int consumed_left[WarpCfg::WA][WarpCfg::WB];
    WarpShared w[WarpCfg::NW];
};

// what a lane offers its neighbours after a step (shuffled to lane+4 / lane+1, or written to a ring)
struct LaneOut {
    uint32_t b_wY[4], b_u2[4], b_u1[4], b_all[4];  // bottom row: wY,u2 of this plane; u1,all of the previous plane
    uint32_t r_wZ[4], r_u2[4], r_u3[4], r_all[4];  // right column
    uint32_t relay_u2, relay_all;                  // the cell above-right of the tile (corner of the right neighbour)
};

// halos a lane holds for its next step (from shuffles; edge lanes overwrite them from a ring or with the border)
struct LaneHalo {
    uint32_t T_wY[4], T_u1[4], T_u2[5], T_all[5];
    uint32_t L_wZ[4], L_u3[4], L_u2[4], L_all[4];
};

struct LaneRegs {
    uint32_t all[4][4], wX[4][4], u1[4][4], u3[4][4];
    uint32_t cBC[4][4], U[4][4];
    LaneHalo H;
};

TA_HD uint32_t wphi(const WarpArgs &A, int i, int jl, int kl) { return rep((uint32_t)(A.base + 2 * (i + jl + kl))); }

TA_HD int warp_total_steps(int P) { return P * WarpCfg::L + WarpCfg::TAIL; }  // steps 0 .. this value

// ---- per-warp sequence loading / table build for pair `pair` into buffer buf (all 32 lanes take part) ----
TA_HD void warp_load_pair(WarpShared &W, const WarpArgs &A, int lane, int a, int b, long long pair, int buf)
{
    using C = WarpCfg;
    for (int e = lane; e < 2 * C::WJ; e += 32) {  // B rows of this warp
        const int h = e / C::WJ, p = e % C::WJ, j = a * C::WJ + p;
        const long long t = 2 * pair + h;
        W.seqB[buf][h][p] = (t < A.n && j < A.len) ? (A.b[t * A.len + j] & 3) : 0;
    }
    for (int e = lane; e < 2 * C::WK; e += 32) {  // C columns of this warp
        const int h = e / C::WK, p = e % C::WK, k = b * C::WK + p;
        const long long t = 2 * pair + h;
        W.seqC[buf][h][p] = (t < A.n && k < A.len) ? (A.c[t * A.len + k] & 3) : 0;
    }
    for (int p = lane; p < 64; p += 32) {
        const long long t0 = 2 * pair, t1 = 2 * pair + 1;
        const uint32_t a0 = (t0 < A.n && p < A.len) ? (A.a[t0 * A.len + p] & 3) : 0;
        const uint32_t a1 = (t1 < A.n && p < A.len) ? (A.a[t1 * A.len + p] & 3) : 0;
        W.acode[buf][p] = (uint8_t)(a0 | (a1 << 2));
    }
}

TA_HD void warp_build_tables(WarpShared &W, int lane, int buf)
{
    using C = WarpCfg;
    for (int e = lane; e < 16 * C::WJ; e += 32) {
        const int code = e / C::WJ, p = e % C::WJ;
        const uint32_t ap = (uint32_t)(code & 3) | ((uint32_t)(code >> 2) << 16);
        const uint32_t bp = (uint32_t)W.seqB[buf][0][p] | ((uint32_t)W.seqB[buf][1][p] << 16);
        const uint32_t nab = neq_flag(ap, bp);
        W.tabAB[buf][code][p] = pair_const(nab);
        W.tabMB[buf][code][p] = eq_mask(nab);
    }
    for (int e = lane; e < 16 * C::WK; e += 32) {
        const int code = e / C::WK, p = e % C::WK;
        const uint32_t ap = (uint32_t)(code & 3) | ((uint32_t)(code >> 2) << 16);
        const uint32_t cp = (uint32_t)W.seqC[buf][0][p] | ((uint32_t)W.seqC[buf][1][p] << 16);
        W.tabAC[buf][code][p] = pair_const(neq_flag(ap, cp));
    }
}

// ---- ring gating, evaluated by one lane: may warp (a,b) run its step w? ----
TA_HD bool warp_can_step(const WarpSmem &S, int a, int b, int w, int last_step, int prod_top, int prod_left, int cons_below,
                         int cons_right)
{
    using C = WarpCfg;
    (void)S;
    if (a > 0 && w + C::TOP_LEAD <= last_step && prod_top < w + C::TOP_LEAD + 1) return false;   // input not there yet
    if (b > 0 && w + C::LEFT_LEAD <= last_step && prod_left < w + C::LEFT_LEAD + 1) return false;
    if (a + 1 < C::WA && cons_below < w - C::kRing + 1) return false;                            // output slot not free
    if (b + 1 < C::WB && cons_right < w - C::kRing + 1) return false;
    return true;
}

// ---- one step of one lane.  H (in R) holds the halos received by shuffle; edge lanes replace them here. ----
TA_HD void warp_lane_step(LaneRegs &R, WarpSmem &S, const WarpArgs &A, int a, int b, int lj, int lk, int w, int P,
                          long long pair_base, int last_step, LaneOut &O)
{
    using C = WarpCfg;
    constexpr int RJ = C::RJ, RK = C::RK;
    WarpShared &W = S.w[a * C::WB + b];
    const int u = w - (lj + lk);
    const int p = u >> 6;
    const int i = (u & 63) + 1;
    const int buf = p & 1;
    const int j0w = lj * RJ, k0w = lk * RK;                 // position inside the warp block
    const int j0 = a * C::WJ + j0w, k0 = b * C::WK + k0w;   // position inside the 64 x 64 plane
    LaneHalo &H = R.H;

    // block-edge halos: from the producer warp's ring, or the zero border planes (src/TriAlign_1cyc.v:156-181)
    if (lj == 0) {
        if (a > 0) {
            if (w + C::TOP_LEAD <= last_step) {
                const uint32_t *e = S.top[a][b][(w + C::TOP_LEAD) % C::kRing];
#pragma unroll
                for (int c = 0; c < RK; ++c) {
                    H.T_wY[c] = e[(0 * C::LK + lk) * RK + c];
                    H.T_u2[c + 1] = e[(1 * C::LK + lk) * RK + c];
                    H.T_u1[c] = e[(2 * C::LK + lk) * RK + c];
                    H.T_all[c + 1] = e[(3 * C::LK + lk) * RK + c];
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                H.T_wY[c] = wphi(A, i, 0, k0 + c + 1);
                H.T_u1[c] = wphi(A, i - 1, 0, k0 + c + 1);
            }
#pragma unroll
            for (int c = 0; c <= RK; ++c) {
                H.T_u2[c] = wphi(A, i, 0, k0 + c);
                H.T_all[c] = wphi(A, i - 1, 0, k0 + c);
            }
        }
    }
    if (lk == 0) {
        if (b > 0) {
            if (w + C::LEFT_LEAD <= last_step) {
                const uint32_t *e = S.left[a][b][(w + C::LEFT_LEAD) % C::kRing];
#pragma unroll
                for (int r = 0; r < RJ; ++r) {
                    H.L_wZ[r] = e[(0 * C::LJ + lj) * RJ + r];
                    H.L_u2[r] = e[(1 * C::LJ + lj) * RJ + r];   // row j0+r+1 of the producer (see below for the shift)
                    H.L_u3[r] = e[(2 * C::LJ + lj) * RJ + r];
                    H.L_all[r] = e[(3 * C::LJ + lj) * RJ + r];
                }
                H.T_u2[0] = e[4 * C::LJ * RJ + lj];
                H.T_all[0] = e[4 * C::LJ * RJ + C::LJ + lj];
            }
        } else {
#pragma unroll
            for (int r = 0; r < RJ; ++r) {
                H.L_wZ[r] = wphi(A, i, j0 + r + 1, 0);
                H.L_u3[r] = wphi(A, i - 1, j0 + r + 1, 0);
                H.L_u2[r] = wphi(A, i, j0 + r + 1, 0);
                H.L_all[r] = wphi(A, i - 1, j0 + r + 1, 0);
            }
            if (!(lj == 0 && a == 0)) {  // corner (row j0, column 0) on the k = 0 border; for the plane's own corner the
                H.T_u2[0] = wphi(A, i, j0, 0);   // top-border code above already set it
                H.T_all[0] = wphi(A, i - 1, j0, 0);
            }
        }
    }

    if (i == 1) {  // a new pair enters this lane (src/PE_1cyc.v:164-178 override = all-zero plane 0)
        uint32_t bp[RJ], cp[RK];
#pragma unroll
        for (int r = 0; r < RJ; ++r) bp[r] = (uint32_t)W.seqB[buf][0][j0w + r] | ((uint32_t)W.seqB[buf][1][j0w + r] << 16);
#pragma unroll
        for (int c = 0; c < RK; ++c) cp[c] = (uint32_t)W.seqC[buf][0][k0w + c] | ((uint32_t)W.seqC[buf][1][k0w + c] << 16);
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c) {
                const uint32_t z = wphi(A, 0, j0 + r + 1, k0 + c + 1);
                R.all[r][c] = z; R.wX[r][c] = z; R.u1[r][c] = z; R.u3[r][c] = z;
                const uint32_t nbc = neq_flag(bp[r], cp[c]);
                R.cBC[r][c] = pair_const(nbc);
                R.U[r][c] = triple_u(nbc);
            }
    }

    const int code = W.acode[buf][i - 1];
    TA_CHECK(code >= 0 && code < 16);
    uint32_t cAB[RJ], mAB[RJ], cAC[RK];
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        cAB[r] = W.tabAB[buf][code][j0w + r];
        mAB[r] = W.tabMB[buf][code][j0w + r];
    }
#pragma unroll
    for (int c = 0; c < RK; ++c) cAC[c] = W.tabAC[buf][code][k0w + c];

    // In H, L_u2[r] / L_all[r] are the left neighbour's values at row j0+r+1 (as shuffled / stored); the cell at
    // row j0+r needs row j0+r-1+1, i.e. index r-1, and the corner for r = 0.
    uint32_t n_all[RJ][RK], n_wX[RJ][RK], n_u1[RJ][RK], n_u3[RJ][RK], n_wY[RJ][RK], n_wZ[RJ][RK], n_u2[RJ][RK];
#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            const uint32_t p_all = (r > 0 && c > 0) ? R.all[r - 1][c - 1] : (r == 0 ? H.T_all[c] : H.L_all[r - 1]);
            const uint32_t p_wY = r > 0 ? n_wY[r - 1][c] : H.T_wY[c];
            const uint32_t p_wZ = c > 0 ? n_wZ[r][c - 1] : H.L_wZ[r];
            const uint32_t p_u1 = r > 0 ? R.u1[r - 1][c] : H.T_u1[c];
            const uint32_t p_u2 = (r > 0 && c > 0) ? n_u2[r - 1][c - 1] : (r == 0 ? H.T_u2[c] : H.L_u2[r - 1]);
            const uint32_t p_u3 = c > 0 ? R.u3[r][c - 1] : H.L_u3[r];
            const uint32_t cM = triple_const(mAB[r], R.U[r][c]);
            const CellOut o = cell_update<false>(p_all, R.wX[r][c], p_wY, p_wZ, p_u1, p_u2, p_u3, cM, cAB[r],
                                                 R.cBC[r][c], cAC[c], nullptr);
            n_all[r][c] = o.all; n_wX[r][c] = o.wX; n_u1[r][c] = o.u1; n_u3[r][c] = o.u3;
            n_wY[r][c] = o.wY; n_wZ[r][c] = o.wZ; n_u2[r][c] = o.u2;
        }

    // what the neighbours need ("old" = plane i-1 values: the PE's _d1/_d2 delay registers, src/PE_1cyc.v:80-109)
#pragma unroll
    for (int c = 0; c < RK; ++c) {
        O.b_wY[c] = n_wY[RJ - 1][c]; O.b_u2[c] = n_u2[RJ - 1][c]; O.b_u1[c] = R.u1[RJ - 1][c]; O.b_all[c] = R.all[RJ - 1][c];
    }
#pragma unroll
    for (int r = 0; r < RJ; ++r) {
        O.r_wZ[r] = n_wZ[r][RK - 1]; O.r_u2[r] = n_u2[r][RK - 1]; O.r_u3[r] = R.u3[r][RK - 1]; O.r_all[r] = R.all[r][RK - 1];
    }
    O.relay_u2 = H.T_u2[RK];
    O.relay_all = H.T_all[RK];

    // rings for the warps below / to the right
    if (lj == C::LJ - 1 && a + 1 < C::WA) {
        uint32_t *e = S.top[a + 1][b][w % C::kRing];
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            e[(0 * C::LK + lk) * RK + c] = O.b_wY[c];
            e[(1 * C::LK + lk) * RK + c] = O.b_u2[c];
            e[(2 * C::LK + lk) * RK + c] = O.b_u1[c];
            e[(3 * C::LK + lk) * RK + c] = O.b_all[c];
        }
    }
    if (lk == C::LK - 1 && b + 1 < C::WB) {
        uint32_t *e = S.left[a][b + 1][w % C::kRing];
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
            e[(0 * C::LJ + lj) * RJ + r] = O.r_wZ[r];
            e[(1 * C::LJ + lj) * RJ + r] = O.r_u2[r];
            e[(2 * C::LJ + lj) * RJ + r] = O.r_u3[r];
            e[(3 * C::LJ + lj) * RJ + r] = O.r_all[r];
        }
        e[4 * C::LJ * RJ + lj] = O.relay_u2;
        e[4 * C::LJ * RJ + C::LJ + lj] = O.relay_all;
    }

    // Score = `all` of cell (len,len,len)
    if (i == A.len && u >= 0 && p < P && j0 / RJ == (A.len - 1) / RJ && k0 / RK == (A.len - 1) / RK) {
        const int rf = (A.len - 1) % RJ, cf = (A.len - 1) % RK;
        uint32_t v = 0;
#pragma unroll
        for (int r = 0; r < RJ; ++r)
#pragma unroll
            for (int c = 0; c < RK; ++c)
                if (r == rf && c == cf) v = n_all[r][c];
        const int off = A.base + 2 * (3 * A.len);
        const long long t0 = 2 * (pair_base + p);
        if (t0 < A.n) A.scores[t0] = (int32_t)(v & 0xFFFFu) - off;
        if (t0 + 1 < A.n) A.scores[t0 + 1] = (int32_t)(v >> 16) - off;
    }

#pragma unroll
    for (int r = 0; r < RJ; ++r)
#pragma unroll
        for (int c = 0; c < RK; ++c) {
            R.all[r][c] = n_all[r][c]; R.wX[r][c] = n_wX[r][c]; R.u1[r][c] = n_u1[r][c]; R.u3[r][c] = n_u3[r][c];
        }
}

// halos for the next step from what the upper (lane-4) and left (lane-1) neighbours offered
TA_HD void warp_take_halo(LaneHalo &H, const LaneOut &up, const LaneOut &lf)
{
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        H.T_wY[c] = up.b_wY[c]; H.T_u1[c] = up.b_u1[c]; H.T_u2[c + 1] = up.b_u2[c]; H.T_all[c + 1] = up.b_all[c];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        H.L_wZ[r] = lf.r_wZ[r]; H.L_u3[r] = lf.r_u3[r]; H.L_u2[r] = lf.r_u2[r]; H.L_all[r] = lf.r_all[r];
    }
    H.T_u2[0] = lf.relay_u2;
    H.T_all[0] = lf.relay_all;
}

// when does warp-local step w rebuild the tables of the next pair?  (buffer of pair q-1 is free once the last lane
// has left it: (w & 63) > LANE_SKEW; pair q+1 starts at the next multiple of 64)
TA_HD bool warp_rebuild_step(int w) { return (w & 63) == 40; }

}  // namespace fast
}  // namespace trialign
