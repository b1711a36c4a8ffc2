// fast_core.cuh -- the factored max-plus cell update shared by the fast kernels (K1 batch, K2 long).
//
// Exactness domain: no candidate of the RTL recurrence wraps at score_bits (proved a priori from the envelope in
// common.cuh, or established at run time by the range check).  Inside that domain the 49 wrapped candidates of
// src/PE_1cyc.v:164-218 can be regrouped, because every transition cost has the form
//     cost(q -> m) = GO*|G(m)| - (GO-GE)*|G(m) & G(q)|,   G(state) = set of sequences that state gaps
//     G(M)={}  G(Ix)={B,C}  G(Iy)={A,C}  G(Iz)={A,B}  G(Ixy)={C}  G(Iyz)={A}  G(Ixz)={B}
// (check against the cost table in SURVEY.md 8(a): e.g. Iy->Ix: C extends, B opens: GE+GO = 3).  With
//     all = max7(V)   hC = max(Ix,Iy,Ixy)   hA = max(Iy,Iz,Iyz)   hB = max(Ix,Iz,Ixz)
// a producer cell hands out seven reduced values instead of 7x7 candidates:
//     to M   (i+1,j+1,k+1): all                       (+ S3(a,b,c) at the consumer)
//     to Ixy (i+1,j+1,k)  : u1 = max(all-1, hC)   (= gXY+1)      to Ix: wX = max(max(u3,u1)-1, Ix)  (= gX+2)
//     to Iyz (i,j+1,k+1)  : u2 = max(all-1, hA)                   to Iy: wY = max(max(u2,u1)-1, Iy)
//     to Ixz (i+1,j,k+1)  : u3 = max(all-1, hB)                   to Iz: wZ = max(max(u2,u3)-1, Iz)
// 11 max-type instructions (VIMNMX3 / VIADDMNMX, .U16x2) per cell instead of 49 adds + 42 maxes: wX is evaluated as
// max3(u3-1, u1-1, Ix) with the decrements as plain 32-bit subtractions (see cell_update).
//
// Potential trick: values are stored as  v~ = v + 2*(i+j+k) + base  (offset binary, unsigned 16-bit lanes).  A gap
// move advances i+j+k by 1, so the "-2" of gX/gY/gZ is absorbed (Ix~ = wX directly); pair moves add (s+3) in {2,4}
// and the M move adds (S3+6) in {3,6,9}: all remaining adds are additions of small positive constants, which are
// carry-free on packed lanes and therefore plain 32-bit adds (IADD3 or IMAD, whichever pipe is free).
//
// Two independent problems ride in the two 16-bit halves of every 32-bit word.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define TA_HD __host__ __device__ __forceinline__
#else
#define TA_HD inline
#endif

// Index checks that are active only in the host emulation (tests/emu); free on the device.
#if !defined(__CUDACC__) && defined(TA_EMU_CHECKS)
#include <cstdio>
#include <cstdlib>
#define TA_CHECK(cond)                                                              \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::fprintf(stderr, "TA_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            std::abort();                                                           \
        }                                                                           \
    } while (0)
#else
#define TA_CHECK(cond) ((void)0)
#endif

#if defined(__CUDACC__)
static __constant__ uint32_t trialign_fma_one = 1;  // opaque multiplier of fast::fadd
#endif

namespace trialign {
namespace fast {

constexpr uint32_t kOnes = 0x00010001u;
TA_HD uint32_t rep(uint32_t v) { return (v & 0xFFFFu) * kOnes; }  // same 16-bit value in both halves

#if defined(__CUDA_ARCH__)
TA_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
TA_HD uint32_t vmin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
TA_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
TA_HD uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
// max(a - 1, c) per half
TA_HD uint32_t vdecmax(uint32_t a, uint32_t c) { return __viaddmax_u16x2(a, 0xFFFFFFFFu, c); }
// max(a + b, c) per half (the add wraps per half)
TA_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_u16x2(a, b, c); }
// a + b issued as IMAD on the FMA pipe: the multiplier is a __constant__ the compiler cannot fold, so ptxas cannot
// turn it back into an ALU-pipe IADD3 / VIADD (the ALU pipe is the bound, the FMA pipe idles otherwise)
TA_HD uint32_t fadd(uint32_t a, uint32_t b) { return a * ::trialign_fma_one + b; }
#else
TA_HD uint32_t fadd(uint32_t a, uint32_t b) { return a + b; }
TA_HD uint32_t vmax2(uint32_t a, uint32_t b)
{
    const uint32_t lo = (a & 0xFFFFu) > (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
TA_HD uint32_t vmin2(uint32_t a, uint32_t b)
{
    const uint32_t lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
TA_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
TA_HD uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) { return vmin2(vmin2(a, b), c); }
TA_HD uint32_t vdecmax(uint32_t a, uint32_t c)
{
    const uint32_t dec = (((a & 0xFFFFu) + 0xFFFFu) & 0xFFFFu) | ((((a >> 16) + 0xFFFFu) & 0xFFFFu) << 16);
    return vmax2(dec, c);
}
TA_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t sum = (((a & 0xFFFFu) + (b & 0xFFFFu)) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
    return vmax2(sum, c);
}
#endif

// ---- scoring parameters (the testbench's commented-out match / mismatch / gap ports) ----
// The factoring holds for any MATCH, MISMATCH, GO, GE with GO2 = 2 GO, GE2 = 2 GE, GOGE = GO + GE (src/PE_1cyc.v:57-59):
// the cost table keeps the form GO*|G(m)| - (GO-GE)*|G(m) & G(q)|.  With delta = GO - GE,
//     u. = max(all - delta, h.)      w. = max(max(u,u') - delta, I.)
// and the potential grows by pi = 2 GE per unit of i+j+k, so that a one-symbol gap move (two gapped sequences, 2 GE)
// still needs no add, a pair move adds s + 3 GE and the M move adds S3 + 6 GE -- all >= 0 inside the domain
// score_fast_domain() accepts, hence carry-free 32-bit adds.
// ScoreRtl: the RTL's localparams (src/PE_1cyc.v:53-56) as compile-time constants -- the default of every template
// below, so that the kernels of the RTL's scoring compile exactly as before.  ScoreGen: run-time values (a separate
// kernel parameter of the *_scored kernels).  SOP: the paper's sum-of-pairs rule instead of the RTL's temp_ABC.
struct ScoreRtl {
    static constexpr bool kRtl = true, kSop = false;
    TA_HD uint32_t pi() const { return 2; }
    TA_HD uint32_t pi2() const { return 0x00020002u; }  // pi in both halves
    TA_HD uint32_t dec() const { return 0u - kOnes; }   // -delta as ONE 32-bit addend: x + dec() = x - delta in both halves
    TA_HD uint32_t dech() const { return 0xFFFFFFFFu; } // -delta as a per-half addend (VIADDMNMX.U16x2 adds half by half)
};
template <bool SOP>
struct ScoreGen {
    static constexpr bool kRtl = false, kSop = SOP;
    uint32_t pi_, pi2_, dec_, dech_;  // 2 GE; the same in both halves; 0 - delta * 0x00010001; (-delta & 0xFFFF) in both halves
    uint32_t pair_eq, pair_d;       // pair constant: pair_eq - pair_d * [x != y]   (s + 3 GE, both halves)
    uint32_t row_eq, row_d;         // RTL rule: tB = row_eq - row_d * [a != b]    (3 MATCH + 6 GE  /  3 MISMATCH + 6 GE)
    uint32_t col_eq, col_d;         // RTL rule: tC = col_eq - col_d * [a != c]    (3 MATCH + 6 GE  /  2 (MATCH+MISMATCH) + 6 GE)
    uint32_t sop_sub;               // sum of pairs: cM = cAB + cBC + cAC - 3 GE (both halves)
    TA_HD uint32_t pi() const { return pi_; }
    TA_HD uint32_t pi2() const { return pi2_; }
    TA_HD uint32_t dec() const { return dec_; }
    TA_HD uint32_t dech() const { return dech_; }
};

// Reduced outputs of one cell, all in the cell's own potential.
struct CellOut {
    uint32_t all, u1, u2, u3, wX, wY, wZ;
};

// One cell update.  Inputs (each already in the potential of THIS cell's predecessors, see header):
//   p_all  = all(i-1,j-1,k-1)   p_wX = wX(i-1,j,k)   p_wY = wY(i,j-1,k)   p_wZ = wZ(i,j,k-1)
//   p_u1   = u1(i-1,j-1,k)      p_u2 = u2(i,j-1,k-1) p_u3 = u3(i-1,j,k-1)
//   cM in {3,6,9} = S3+6,  cAB/cBC/cAC in {2,4} = pair score + 3   (packed, per half)
// Returns the 7 states through st[] (M,Ix,Iy,Iz,Ixy,Iyz,Ixz in potential form) when WANT_STATES.
//
// Pipe balance (profiles/alu_peak.json: the integer-ALU pipe issues 2 warp-instr/clk/SM and is the bound; plain
// 32-bit adds can also issue as IMAD on the FMA pipe): 11 max-type instructions on the ALU pipe -- the M add is fused
// into the `all` reduction (VIADDMNMX) and the three "-1" of wX/wY/wZ are plain 32-bit subtractions of 0x00010001.
// Those are lane-exact as long as the LOW lane is >= 1, which holds for every live value (potentials are >= base -
// envelope > 0, common.cuh); a dead low lane can only disturb a dead high lane (the halves of a word start and stop
// together in BATCH mode, and in LONG mode the low half stays live -- clamped a index -- until the high half ends).
template <bool WANT_STATES, class SC = ScoreRtl>
TA_HD CellOut cell_update(uint32_t p_all, uint32_t p_wX, uint32_t p_wY, uint32_t p_wZ, uint32_t p_u1, uint32_t p_u2,
                          uint32_t p_u3, uint32_t cM, uint32_t cAB, uint32_t cBC, uint32_t cAC, uint32_t *st,
                          const SC &sc = SC())
{
    const uint32_t X = p_wX, Y = p_wY, Z = p_wZ;
    const uint32_t XY = fadd(p_u1, cAB);
    const uint32_t YZ = fadd(p_u2, cBC);
    const uint32_t XZ = fadd(p_u3, cAC);
    if (WANT_STATES) {
        st[0] = p_all + cM; st[1] = X; st[2] = Y; st[3] = Z; st[4] = XY; st[5] = YZ; st[6] = XZ;
    }
    CellOut o;
    const uint32_t hC = vmax3(X, Y, XY);
    const uint32_t hA = vmax3(Y, Z, YZ);
    const uint32_t hB = vmax3(X, Z, XZ);
    o.all = vaddmax(p_all, cM, vmax3(hC, hA, hB));
    if (SC::kRtl) {
        o.u1 = vdecmax(o.all, hC);
        o.u2 = vdecmax(o.all, hA);
        o.u3 = vdecmax(o.all, hB);
    } else {
        o.u1 = vaddmax(o.all, sc.dech(), hC);
        o.u2 = vaddmax(o.all, sc.dech(), hA);
        o.u3 = vaddmax(o.all, sc.dech(), hB);
    }
    const uint32_t d1 = fadd(o.u1, sc.dec()), d2 = fadd(o.u2, sc.dec()), d3 = fadd(o.u3, sc.dec());
    o.wX = vmax3(d3, d1, X);
    o.wY = vmax3(d2, d1, Y);
    o.wZ = vmax3(d2, d3, Z);
    return o;
}

// pair-score constants from 2-bit symbols packed one per half (x, y in {0..3} per half)
TA_HD uint32_t neq_flag(uint32_t x, uint32_t y) { return vmin2(x ^ y, kOnes); }       // 1 where x != y, per half
TA_HD uint32_t pair_const(uint32_t neq) { return 0x00040004u - 2u * neq; }             // {4 (equal), 2 (differ)}
// S3 + 6 = {9 (a=b=c), 6 (a=b!=c), 3 (a!=b)} (src/PE_1cyc.v:162) factorises into a row and a column constant:
//   cM = min(tB, tC),  tB = 9 - 6*[a!=b] in {9,3},  tC = 9 - 3*[a!=c] in {9,6}
// (a!=b -> 3 whatever c is; a=b -> b=c iff a=c, so tC decides between 9 and 6).  No per-cell (b,c) constant needed.
TA_HD uint32_t triple_row_const(uint32_t neq_ab) { return 0x00090009u - 6u * neq_ab; }
TA_HD uint32_t triple_col_const(uint32_t neq_ac) { return 0x00090009u - 3u * neq_ac; }
TA_HD uint32_t triple_const(uint32_t tB, uint32_t tC) { return vmin2(tB, tC); }

// the same from run-time parameters (neq has 0 / 1 per half, so the products stay inside their halves)
TA_HD uint32_t pair_const(uint32_t neq, const ScoreRtl &) { return pair_const(neq); }
TA_HD uint32_t triple_row_const(uint32_t neq, const ScoreRtl &) { return triple_row_const(neq); }
TA_HD uint32_t triple_col_const(uint32_t neq, const ScoreRtl &) { return triple_col_const(neq); }
template <bool SOP>
TA_HD uint32_t pair_const(uint32_t neq, const ScoreGen<SOP> &sc) { return sc.pair_eq - (sc.pair_d & 0xFFFFu) * neq; }
template <bool SOP>
TA_HD uint32_t triple_row_const(uint32_t neq, const ScoreGen<SOP> &sc) { return sc.row_eq - (sc.row_d & 0xFFFFu) * neq; }
template <bool SOP>
TA_HD uint32_t triple_col_const(uint32_t neq, const ScoreGen<SOP> &sc) { return sc.col_eq - (sc.col_d & 0xFFFFu) * neq; }
// cM of one cell: RTL rule min(tB, tC); sum of pairs cAB + cBC + cAC - 3 GE (three adds on the FMA pipe)
template <class SC>
TA_HD uint32_t triple_const_of(uint32_t tB, uint32_t tC, uint32_t cAB, uint32_t cBC, uint32_t cAC, const SC &sc)
{
    if constexpr (SC::kSop) return fadd(fadd(fadd(cAB, cAC), cBC), 0u - sc.sop_sub);
    else return triple_const(tB, tC);
}

}  // namespace fast
}  // namespace trialign
