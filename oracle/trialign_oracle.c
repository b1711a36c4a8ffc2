/*
 * trialign_oracle.c -- O1, the functional CPU oracle.  TEST INFRASTRUCTURE ONLY (see header).
 *
 * Follows, from /root/reference (file:line):
 *   src/PE_1cyc.v:55-61     scoring localparams MATCH, MISMATCH, GO, GE, GO2, GE2, GOGE
 *   src/PE_1cyc.v:63-66     symbols are 2-bit registers (inputs 4-bit, truncated)
 *   src/PE_1cyc.v:159-162   temp_AB / temp_BC / temp_AC / temp_ABC (incl. the precedence quirk)
 *   src/PE_1cyc.v:164-218   the 49 candidates
 *   src/PE_1cyc.v:22-28, 139-145   MAX7 (signed, value only -- `act` is commented out)
 *   src/PE_1cyc.v:127-133   candidate wires are `signed [wordsize-1:0]` -> wrap mod 2^W, then signed max
 *   src/PE_1cyc.v:164-178,196-202,212-218  EN_i&&!EN override -> plane i=0 is all-zero
 *   src/TriAlign_1cyc.v:156-181  slice_y==0 / slice_z==0 borders are 0 in all 7 states
 *   src/TriAlign_1cyc.v:141-142,342-345  Score = MAX7 of the 7 states of cell (A_idx,B_idx,C_idx)
 * PE_2cyc.v / TriAlign_2cyc.v compute the same function (one more pipeline stage).
 */
#include "trialign_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* PE_1cyc.v:55-61 */
#define MATCH 1
#define MISMATCH (-1)
#define GO 2
#define GE 1
#define GO2 (GO << 1)
#define GE2 (GE << 1)
#define GOGE (GO + GE)

/* Two's-complement truncation to W bits, read back as signed (PE_1cyc.v:127-133). sh = 32-W. */
static inline int32_t wrapw(uint32_t x, int sh) { return (int32_t)(x << sh) >> sh; }

/* PE_1cyc.v:22-28: three-level compare-select tree; value-neutral on ties. */
static inline int32_t max7(int32_t g4, int32_t g1a, int32_t g1b, int32_t g3a, int32_t g2a, int32_t g2b,
                           int32_t g3b)
{
    int32_t t1 = g1a > g1b ? g1a : g1b;
    int32_t t2 = g2a > g2b ? g2a : g2b;
    int32_t t3 = g3a > g3b ? g3a : g3b;
    int32_t t12 = t1 > t2 ? t1 : t2;
    int32_t t34 = t3 > g4 ? t3 : g4;
    return t12 > t34 ? t12 : t34;
}

/* PE_1cyc.v:159-162.  Verilog precedence: `+` binds tighter than `<<`, so
 *   MATCH+MISMATCH<<1  == (MATCH+MISMATCH)<<1 == 0      (A==B, B!=C)
 *   MATCH<<1+MISMATCH  is only reachable when A==B, B==C, A!=C, i.e. never. */
static inline void pe_scores(unsigned A, unsigned B, unsigned C, int *sAB, int *sBC, int *sAC, int *sABC)
{
    *sAB = (A == B) ? MATCH : MISMATCH;
    *sBC = (B == C) ? MATCH : MISMATCH;
    *sAC = (A == C) ? MATCH : MISMATCH;
    *sABC = (A == B) ? ((B == C) ? ((A == C) ? MATCH * 3 : (MATCH << (1 + MISMATCH)))
                                 : ((MATCH + MISMATCH) << 1))
                     : MISMATCH * 3;
}

/* One PE evaluation.  Predecessor vectors (each 7 states, order M,Ix,Iy,Iz,Ixy,Iyz,Ixz):
 *   p111=(i-1,j-1,k-1) [_1_d2]   p100=(i-1,j,k) [own regs]   p010=(i,j-1,k) [_2_i]   p001=(i,j,k-1) [_3_i]
 *   p110=(i-1,j-1,k) [_2_d1]     p011=(i,j-1,k-1) [_1_d1]    p101=(i-1,j,k-1) [_3_d1]          */
static inline void pe_cell(int32_t *o, const int32_t *p111, const int32_t *p100, const int32_t *p010,
                           const int32_t *p001, const int32_t *p110, const int32_t *p011,
                           const int32_t *p101, unsigned A, unsigned B, unsigned C, int sh)
{
    int sAB, sBC, sAC, sABC;
    pe_scores(A, B, C, &sAB, &sBC, &sAC, &sABC);
#define W_(x) wrapw((uint32_t)(x), sh)
    /* M: PE_1cyc.v:164-170, MAX7 port map :139 */
    o[TA_M] = max7(W_(p111[TA_M] + sABC), W_(p111[TA_IX] + sABC), W_(p111[TA_IY] + sABC),
                   W_(p111[TA_IZ] + sABC), W_(p111[TA_IXY] + sABC), W_(p111[TA_IYZ] + sABC),
                   W_(p111[TA_IXZ] + sABC));
    /* Ix: :172-178, port map :140 (G4=Ix, G1=M,Iyz, G3=Iy,Iz, G2=Ixy,Ixz) */
    o[TA_IX] = max7(W_(p100[TA_IX] - GE2), W_(p100[TA_M] - GO2), W_(p100[TA_IYZ] - GO2),
                    W_(p100[TA_IY] - GOGE), W_(p100[TA_IXY] - GOGE), W_(p100[TA_IXZ] - GOGE),
                    W_(p100[TA_IZ] - GOGE));
    /* Iy: :180-186, port map :141 */
    o[TA_IY] = max7(W_(p010[TA_IY] - GE2), W_(p010[TA_M] - GO2), W_(p010[TA_IXZ] - GO2),
                    W_(p010[TA_IX] - GOGE), W_(p010[TA_IXY] - GOGE), W_(p010[TA_IYZ] - GOGE),
                    W_(p010[TA_IZ] - GOGE));
    /* Iz: :188-194, port map :142 */
    o[TA_IZ] = max7(W_(p001[TA_IZ] - GE2), W_(p001[TA_M] - GO2), W_(p001[TA_IXY] - GO2),
                    W_(p001[TA_IX] - GOGE), W_(p001[TA_IXZ] - GOGE), W_(p001[TA_IYZ] - GOGE),
                    W_(p001[TA_IY] - GOGE));
    /* Ixy: :196-202, port map :143 */
    o[TA_IXY] = max7(W_(p110[TA_IY] - GE + sAB), W_(p110[TA_M] - GO + sAB), W_(p110[TA_IXZ] - GO + sAB),
                     W_(p110[TA_IYZ] - GO + sAB), W_(p110[TA_IXY] - GE + sAB), W_(p110[TA_IX] - GE + sAB),
                     W_(p110[TA_IZ] - GO + sAB));
    /* Iyz: :204-210, port map :144 */
    o[TA_IYZ] = max7(W_(p011[TA_IZ] - GE + sBC), W_(p011[TA_M] - GO + sBC), W_(p011[TA_IXY] - GO + sBC),
                     W_(p011[TA_IYZ] - GE + sBC), W_(p011[TA_IXZ] - GO + sBC), W_(p011[TA_IX] - GO + sBC),
                     W_(p011[TA_IY] - GE + sBC));
    /* Ixz: :212-218, port map :145 */
    o[TA_IXZ] = max7(W_(p101[TA_IZ] - GE + sAC), W_(p101[TA_M] - GO + sAC), W_(p101[TA_IXY] - GO + sAC),
                     W_(p101[TA_IYZ] - GO + sAC), W_(p101[TA_IXZ] - GE + sAC), W_(p101[TA_IX] - GE + sAC),
                     W_(p101[TA_IY] - GO + sAC));
#undef W_
}

static int bad_args(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int w)
{
    if (la < 0 || lb < 0 || lc < 0) return 1;
    if (w < 2 || w > 32) return 1;
    if ((la && !a) || (lb && !b) || (lc && !c)) return 1;
    return 0;
}

int oracle_score(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                 int score_bits, int32_t *score, int32_t final_states[7], int32_t range[2])
{
    if (bad_args(a, la, b, lb, c, lc, score_bits) || !score) return -1;
    const int sh = 32 - score_bits;
    int32_t lo = 0, hi = 0;
    int32_t fin[7] = {0, 0, 0, 0, 0, 0, 0};
    if (la > 0 && lb > 0 && lc > 0) {
        /* two rolling (j,k) planes, zero borders (TriAlign_1cyc.v:156-181; PE override for i=0) */
        const size_t row = (size_t)(lc + 1) * 7, plane = (size_t)(lb + 1) * row;
        int32_t *buf = (int32_t *)calloc(2 * plane, sizeof(int32_t));
        if (!buf) return -2;
        int32_t *prev = buf, *cur = buf + plane;
        for (int i = 1; i <= la; ++i) {
            const unsigned A = a[i - 1] & 3u;
            for (int j = 1; j <= lb; ++j) {
                const unsigned B = b[j - 1] & 3u;
                for (int k = 1; k <= lc; ++k) {
                    const unsigned C = c[k - 1] & 3u;
                    int32_t *o = cur + j * row + (size_t)k * 7;
                    pe_cell(o, prev + (j - 1) * row + (size_t)(k - 1) * 7, prev + j * row + (size_t)k * 7,
                            cur + (j - 1) * row + (size_t)k * 7, cur + j * row + (size_t)(k - 1) * 7,
                            prev + (j - 1) * row + (size_t)k * 7, cur + (j - 1) * row + (size_t)(k - 1) * 7,
                            prev + j * row + (size_t)(k - 1) * 7, A, B, C, sh);
                    for (int q = 0; q < 7; ++q) {
                        if (o[q] < lo) lo = o[q];
                        if (o[q] > hi) hi = o[q];
                    }
                }
            }
            int32_t *t = prev; prev = cur; cur = t;
        }
        memcpy(fin, prev + lb * row + (size_t)lc * 7, sizeof fin);
        free(buf);
    }
    /* FINAL_MAX: TriAlign_1cyc.v:141-142 (G1=M,Ix G2=Iy,Iz G3=Ixy,Iyz G4=Ixz) */
    *score = max7(fin[TA_IXZ], fin[TA_M], fin[TA_IX], fin[TA_IXY], fin[TA_IY], fin[TA_IZ], fin[TA_IYZ]);
    if (final_states) memcpy(final_states, fin, sizeof fin);
    if (range) { range[0] = lo; range[1] = hi; }
    return 0;
}

int oracle_cube(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                int score_bits, int32_t *out)
{
    if (bad_args(a, la, b, lb, c, lc, score_bits) || !out) return -1;
    const int sh = 32 - score_bits;
    const size_t sk = 7, sj = (size_t)(lc + 1) * sk, si = (size_t)(lb + 1) * sj;
    memset(out, 0, (size_t)(la + 1) * si * sizeof(int32_t));
    for (int i = 1; i <= la; ++i)
        for (int j = 1; j <= lb; ++j)
            for (int k = 1; k <= lc; ++k) {
#define AT(I, J, K) (out + (size_t)(I) * si + (size_t)(J) * sj + (size_t)(K) * sk)
                pe_cell(AT(i, j, k), AT(i - 1, j - 1, k - 1), AT(i - 1, j, k), AT(i, j - 1, k), AT(i, j, k - 1),
                        AT(i - 1, j - 1, k), AT(i, j - 1, k - 1), AT(i - 1, j, k - 1), a[i - 1] & 3u,
                        b[j - 1] & 3u, c[k - 1] & 3u, sh);
#undef AT
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Traceback.  The RTL ships score-only, but carries the hook: MAX7's commented-out `act` output
 * (PE_1cyc.v:12,30) would report WHICH of its seven inputs won.  The tie rule follows from the compare-select tree
 * (:22-28, `a > b ? a : b`: the second operand wins a tie), i.e. on equal values the ports win in the order
 * G4, G3_B, G3_A, G2_B, G2_A, G1_B, G1_A.  Which source state sits on which port differs per MAX7 instance (:139-145).
 * Tables below: ports in max7()'s argument order (G4, G1_A, G1_B, G3_A, G2_A, G2_B, G3_B).
 * ------------------------------------------------------------------------------------------------------------------ */
static const int kPortSrc[7][7] = {
    /* M   :139 */ {TA_M, TA_IX, TA_IY, TA_IZ, TA_IXY, TA_IYZ, TA_IXZ},
    /* Ix  :140 */ {TA_IX, TA_M, TA_IYZ, TA_IY, TA_IXY, TA_IXZ, TA_IZ},
    /* Iy  :141 */ {TA_IY, TA_M, TA_IXZ, TA_IX, TA_IXY, TA_IYZ, TA_IZ},
    /* Iz  :142 */ {TA_IZ, TA_M, TA_IXY, TA_IX, TA_IXZ, TA_IYZ, TA_IY},
    /* Ixy :143 */ {TA_IY, TA_M, TA_IXZ, TA_IYZ, TA_IXY, TA_IX, TA_IZ},
    /* Iyz :144 */ {TA_IZ, TA_M, TA_IXY, TA_IYZ, TA_IXZ, TA_IX, TA_IY},
    /* Ixz :145 */ {TA_IZ, TA_M, TA_IXY, TA_IYZ, TA_IXZ, TA_IX, TA_IY},
};
/* FINAL_MAX, TriAlign_1cyc.v:141-142 */
static const int kFinalPortSrc[7] = {TA_IXZ, TA_M, TA_IX, TA_IXY, TA_IY, TA_IZ, TA_IYZ};
/* cost subtracted by destination (row) and source (column), PE_1cyc.v:164-218 */
static const int kCost[7][7] = {
    {0, 0, 0, 0, 0, 0, 0},                         /* M   */
    {GO2, GE2, GOGE, GOGE, GOGE, GO2, GOGE},       /* Ix  */
    {GO2, GOGE, GE2, GOGE, GOGE, GOGE, GO2},       /* Iy  */
    {GO2, GOGE, GOGE, GE2, GO2, GOGE, GOGE},       /* Iz  */
    {GO, GE, GE, GO, GE, GO, GO},                  /* Ixy */
    {GO, GO, GE, GE, GO, GE, GO},                  /* Iyz */
    {GO, GE, GO, GE, GO, GO, GE},                  /* Ixz */
};
/* predecessor offset of a destination state, and the alignment column it emits (bit0 = A advances, bit1 = B, bit2 = C) */
static const int kDelta[7][3] = {{1, 1, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}};

/* index (in port order) of the input MAX7 selects */
static inline int max7_port(const int32_t g[7])
{
    const int i1 = g[1] > g[2] ? 1 : 2, i2 = g[4] > g[5] ? 4 : 5, i3 = g[3] > g[6] ? 3 : 6;
    const int i12 = g[i1] > g[i2] ? i1 : i2;
    const int i34 = g[i3] > g[0] ? i3 : 0;
    return g[i12] > g[i34] ? i12 : i34;
}

int oracle_align(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                 int32_t *score, int32_t start[3], uint8_t *moves, int cap, int *n_moves)
{
    if (bad_args(a, la, b, lb, c, lc, score_bits) || !score || !start || !n_moves || cap < 0 || (cap && !moves))
        return -1;
    const int sh = 32 - score_bits;
    *n_moves = 0;
    start[0] = la; start[1] = lb; start[2] = lc;
    if (la == 0 || lb == 0 || lc == 0) { *score = 0; return 0; }
    const size_t sk = 7, sj = (size_t)(lc + 1) * sk, si = (size_t)(lb + 1) * sj;
    int32_t *cube = (int32_t *)malloc((size_t)(la + 1) * si * sizeof(int32_t));
    if (!cube) return -2;
    int rc = oracle_cube(a, la, b, lb, c, lc, score_bits, cube);
    if (rc) { free(cube); return rc; }
#define AT(I, J, K) (cube + (size_t)(I) * si + (size_t)(J) * sj + (size_t)(K) * sk)
    int32_t g[7];
    const int32_t *fin = AT(la, lb, lc);
    for (int p = 0; p < 7; ++p) g[p] = fin[kFinalPortSrc[p]];
    int port = max7_port(g);
    *score = g[port];
    int state = kFinalPortSrc[port];
    int i = la, j = lb, k = lc, n = 0;
    while (i > 0 && j > 0 && k > 0) {
        if (n >= cap) { free(cube); return -4; }
        moves[n++] = (uint8_t)(kDelta[state][0] | (kDelta[state][1] << 1) | (kDelta[state][2] << 2));
        int sAB, sBC, sAC, sABC;
        pe_scores(a[i - 1] & 3u, b[j - 1] & 3u, c[k - 1] & 3u, &sAB, &sBC, &sAC, &sABC);
        const int add = state == TA_M ? sABC : state == TA_IXY ? sAB : state == TA_IYZ ? sBC : state == TA_IXZ ? sAC : 0;
        const int32_t own = AT(i, j, k)[state];
        i -= kDelta[state][0]; j -= kDelta[state][1]; k -= kDelta[state][2];
        const int32_t *pred = AT(i, j, k);  /* all-zero on the three border planes (TriAlign_1cyc.v:156-181) */
        for (int p = 0; p < 7; ++p) {
            const int q = kPortSrc[state][p];
            g[p] = wrapw((uint32_t)(pred[q] - kCost[state][q] + add), sh);
        }
        port = max7_port(g);
        if (g[port] != own) { free(cube); return -3; }  /* the tables above disagree with pe_cell(): a bug */
        state = kPortSrc[state][port];
    }
#undef AT
    free(cube);
    for (int x = 0; x < n / 2; ++x) { const uint8_t t = moves[x]; moves[x] = moves[n - 1 - x]; moves[n - 1 - x] = t; }
    start[0] = i; start[1] = j; start[2] = k;
    *n_moves = n;
    return 0;
}

int oracle_score_mt(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                    int score_bits, int threads, int32_t *score, int32_t final_states[7])
{
    if (bad_args(a, la, b, lb, c, lc, score_bits) || !score) return -1;
    const int sh = 32 - score_bits;
    int32_t fin[7] = {0, 0, 0, 0, 0, 0, 0};
    if (la > 0 && lb > 0 && lc > 0) {
        /* four rolling anti-diagonal planes d=i+j+k, indexed [j][k]; row 0 / column 0 stay zero */
        const size_t row = (size_t)(lc + 1) * 7, plane = (size_t)(lb + 1) * row;
        int32_t *buf = (int32_t *)calloc(4 * plane, sizeof(int32_t));
        if (!buf) return -2;
        static const int32_t zero7[7] = {0, 0, 0, 0, 0, 0, 0};
#ifdef _OPENMP
        if (threads <= 0) threads = omp_get_max_threads();
#else
        threads = 1;
#endif
        for (int d = 3; d <= la + lb + lc; ++d) {
            int32_t *P0 = buf + (size_t)(d & 3) * plane;        /* plane d   */
            const int32_t *P1 = buf + (size_t)((d - 1) & 3) * plane; /* d-1 */
            const int32_t *P2 = buf + (size_t)((d - 2) & 3) * plane; /* d-2 */
            const int32_t *P3 = buf + (size_t)((d - 3) & 3) * plane; /* d-3 */
            int jlo = d - la - lc; if (jlo < 1) jlo = 1;
            int jhi = d - 2; if (jhi > lb) jhi = lb;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
            for (int j = jlo; j <= jhi; ++j) {
                int klo = d - j - la; if (klo < 1) klo = 1;
                int khi = d - j - 1; if (khi > lc) khi = lc;
                const unsigned B = b[j - 1] & 3u;
                for (int k = klo; k <= khi; ++k) {
                    const int i = d - j - k;
                    const int first = (i == 1); /* (i-1) predecessors are the zero plane */
                    pe_cell(P0 + j * row + (size_t)k * 7,
                            first ? zero7 : P3 + (j - 1) * row + (size_t)(k - 1) * 7,
                            first ? zero7 : P1 + j * row + (size_t)k * 7,
                            P1 + (j - 1) * row + (size_t)k * 7, P1 + j * row + (size_t)(k - 1) * 7,
                            first ? zero7 : P2 + (j - 1) * row + (size_t)k * 7,
                            P2 + (j - 1) * row + (size_t)(k - 1) * 7,
                            first ? zero7 : P2 + j * row + (size_t)(k - 1) * 7, a[i - 1] & 3u, B,
                            c[k - 1] & 3u, sh);
                }
            }
        }
        memcpy(fin, buf + (size_t)((la + lb + lc) & 3) * plane + lb * row + (size_t)lc * 7, sizeof fin);
        free(buf);
    }
    *score = max7(fin[TA_IXZ], fin[TA_M], fin[TA_IX], fin[TA_IXY], fin[TA_IY], fin[TA_IZ], fin[TA_IYZ]);
    if (final_states) memcpy(final_states, fin, sizeof fin);
    return 0;
}

/* Same function again, scheduled like the RTL's slicing (TriAlign_1cyc.v:127-140,289-335): the (j,k) plane is cut into
 * blk x blk blocks, every block is swept over all i with two small rolling planes (cache resident), and the block's last
 * row / last column of every plane i are parked in face memories for the blocks below / to the right (the z-face and
 * y-face SRAMs; the corner cell the RTL routes through border_00, :149,175-181, travels as element 0 of a face row).
 * Blocks on one anti-diagonal bj+bk are independent -> OpenMP over them; faces of two anti-diagonals are live.
 * Every cell still goes through the literal pe_cell().  Used for the long golden vectors (8192^3 in ~20 min). */
int oracle_score_blocked(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                         int score_bits, int threads, int blk, int32_t *score, int32_t final_states[7])
{
    if (bad_args(a, la, b, lb, c, lc, score_bits) || !score || blk < 1) return -1;
    const int sh = 32 - score_bits;
    int32_t fin[7] = {0, 0, 0, 0, 0, 0, 0};
    int rc = 0;
    if (la > 0 && lb > 0 && lc > 0) {
        const int nbj = (lb + blk - 1) / blk, nbk = (lc + blk - 1) / blk;
        const size_t frow = (size_t)(blk + 1) * 7;             /* one face row: corner + blk cells, 7 states */
        const size_t fsz = (size_t)(la + 1) * frow;            /* one face: planes 0..la */
        /* J faces (last row of a block, read by block bj+1) and K faces (last column, read by bk+1), [parity][bj] */
        int32_t *jf = (int32_t *)calloc((size_t)2 * nbj * fsz, sizeof(int32_t));
        int32_t *kf = (int32_t *)calloc((size_t)2 * nbj * fsz, sizeof(int32_t));
        if (!jf || !kf) { free(jf); free(kf); return -2; }
#ifdef _OPENMP
        if (threads <= 0) threads = omp_get_max_threads();
#else
        threads = 1;
#endif
        const size_t prow = (size_t)(blk + 1) * 7, pplane = (size_t)(blk + 1) * prow;
        for (int d = 0; d <= nbj + nbk - 2; ++d) {
            int blo = d - (nbk - 1); if (blo < 0) blo = 0;
            int bhi = d < nbj - 1 ? d : nbj - 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
            for (int bj = blo; bj <= bhi; ++bj) {
                const int bk = d - bj;
                const int j0 = bj * blk, k0 = bk * blk;
                const int nj = lb - j0 < blk ? lb - j0 : blk, nk = lc - k0 < blk ? lc - k0 : blk;
                int32_t *pl = (int32_t *)calloc(2 * pplane, sizeof(int32_t));
                if (!pl) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                    rc = -2;
                    continue;
                }
                /* inbound: J face of (bj-1,bk) and K face of (bj,bk-1), both written on anti-diagonal d-1 */
                const int32_t *jin = bj > 0 ? jf + ((size_t)((d - 1) & 1) * nbj + (bj - 1)) * fsz : NULL;
                const int32_t *kin = bk > 0 ? kf + ((size_t)((d - 1) & 1) * nbj + bj) * fsz : NULL;
                int32_t *jout = jf + ((size_t)(d & 1) * nbj + bj) * fsz;
                int32_t *kout = kf + ((size_t)(d & 1) * nbj + bj) * fsz;
                int32_t *prev = pl, *cur = pl + pplane;   /* plane 0 is all zero (PE first-symbol override) */
                for (int i = 1; i <= la; ++i) {
                    const unsigned A = a[i - 1] & 3u;
                    /* halos of plane i: local row 0 = global row j0, local column 0 = global column k0 */
                    if (jin) memcpy(cur, jin + (size_t)i * frow, (size_t)(nk + 1) * 7 * sizeof(int32_t));
                    else memset(cur, 0, (size_t)(nk + 1) * 7 * sizeof(int32_t));
                    for (int j = 1; j <= nj; ++j) {
                        if (kin) memcpy(cur + j * prow, kin + (size_t)i * frow + (size_t)j * 7, 7 * sizeof(int32_t));
                        else memset(cur + j * prow, 0, 7 * sizeof(int32_t));
                    }
                    if (!jin && kin) memcpy(cur, kin + (size_t)i * frow, 7 * sizeof(int32_t)); /* corner via the K face */
                    for (int j = 1; j <= nj; ++j) {
                        const unsigned B = b[j0 + j - 1] & 3u;
                        for (int k = 1; k <= nk; ++k) {
                            const unsigned C = c[k0 + k - 1] & 3u;
                            pe_cell(cur + j * prow + (size_t)k * 7, prev + (j - 1) * prow + (size_t)(k - 1) * 7,
                                    prev + j * prow + (size_t)k * 7, cur + (j - 1) * prow + (size_t)k * 7,
                                    cur + j * prow + (size_t)(k - 1) * 7, prev + (j - 1) * prow + (size_t)k * 7,
                                    cur + (j - 1) * prow + (size_t)(k - 1) * 7, prev + j * prow + (size_t)(k - 1) * 7,
                                    A, B, C, sh);
                        }
                    }
                    /* outbound faces: last row incl. its column-0 halo, last column incl. its row-0 halo */
                    memcpy(jout + (size_t)i * frow, cur + (size_t)nj * prow, (size_t)(nk + 1) * 7 * sizeof(int32_t));
                    for (int j = 0; j <= nj; ++j)
                        memcpy(kout + (size_t)i * frow + (size_t)j * 7, cur + j * prow + (size_t)nk * 7,
                               7 * sizeof(int32_t));
                    int32_t *t = prev; prev = cur; cur = t;
                }
                if (bj == nbj - 1 && bk == nbk - 1) memcpy(fin, prev + nj * prow + (size_t)nk * 7, sizeof fin);
                free(pl);
            }
            if (rc) break;
        }
        free(jf); free(kf);
        if (rc) return rc;
    }
    *score = max7(fin[TA_IXZ], fin[TA_M], fin[TA_IX], fin[TA_IXY], fin[TA_IY], fin[TA_IZ], fin[TA_IYZ]);
    if (final_states) memcpy(final_states, fin, sizeof fin);
    return 0;
}

int oracle_score_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len,
                       int score_bits, int threads, int32_t *scores)
{
    if (n < 0 || len < 0 || score_bits < 2 || score_bits > 32 || !scores) return -1;
    if (n && len && (!a || !b || !c)) return -1;
    int rc = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
    for (int64_t t = 0; t < n; ++t) {
        int r = oracle_score(a + t * len, len, b + t * len, len, c + t * len, len, score_bits, &scores[t],
                             NULL, NULL);
        if (r) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            rc = r;
        }
    }
    return rc;
}

/* sym = mix(seed ^ ((3t+s) << 32) ^ p) >> 62, mix = splitmix64 finaliser. */
static inline uint64_t mix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_synth(uint64_t seed, int64_t triple, int which, int len, uint8_t *out)
{
    const uint64_t hi = ((uint64_t)(3 * triple + which)) << 32;
    for (int p = 0; p < len; ++p) out[p] = (uint8_t)(mix64(seed ^ hi ^ (uint64_t)p) >> 62);
}

int oracle_hw_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------------------------
 * Explicit scoring parameters: the values the testbench's commented-out .i_match / .i_mismatch /
 * .i_gap_open / .i_gap_extend ports would feed (TriAlign_tb.sv:220-224) in place of the PE localparams (PE_1cyc.v:53-59).
 * Table-driven and independent of pe_cell(): destination state m reads the predecessor cell at offset kDelta[m],
 * subtracts kCostSym[m][q] per source state q, adds the move's substitution score; every candidate is wrapped to
 * score_bits before the signed max (PE_1cyc.v:127-133).
 * ------------------------------------------------------------------------------------------------------------------ */
enum { C_0 = 0, C_GO, C_GE, C_GO2, C_GE2, C_GOGE };
static const int kCostSym[7][7] = {
    {C_0, C_0, C_0, C_0, C_0, C_0, C_0},                      /* M   PE_1cyc.v:164-170 */
    {C_GO2, C_GE2, C_GOGE, C_GOGE, C_GOGE, C_GO2, C_GOGE},    /* Ix  :172-178 */
    {C_GO2, C_GOGE, C_GE2, C_GOGE, C_GOGE, C_GOGE, C_GO2},    /* Iy  :180-186 */
    {C_GO2, C_GOGE, C_GOGE, C_GE2, C_GO2, C_GOGE, C_GOGE},    /* Iz  :188-194 */
    {C_GO, C_GE, C_GE, C_GO, C_GE, C_GO, C_GO},               /* Ixy :196-202 */
    {C_GO, C_GO, C_GE, C_GE, C_GO, C_GE, C_GO},               /* Iyz :204-210 */
    {C_GO, C_GE, C_GO, C_GE, C_GO, C_GO, C_GE},               /* Ixz :212-218 */
};

int oracle_score_ex(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                    const int32_t scoring[5], int32_t *score, int32_t final_states[7])
{
    if (bad_args(a, la, b, lb, c, lc, score_bits) || !score || !scoring) return -1;
    const int32_t M = scoring[0], X = scoring[1], go = scoring[2], ge = scoring[3], rule = scoring[4];
    const int32_t costv[6] = {0, go, ge, go << 1, ge << 1, go + ge};  /* PE_1cyc.v:57-59 */
    const int sh = 32 - score_bits;
    int32_t fin[7] = {0, 0, 0, 0, 0, 0, 0};
    if (la > 0 && lb > 0 && lc > 0) {
        const size_t sk = 7, sj = (size_t)(lc + 1) * sk, si = (size_t)(lb + 1) * sj;
        int32_t *buf = (int32_t *)calloc(2 * si, sizeof(int32_t)); /* planes i-1 and i; borders stay 0 */
        if (!buf) return -2;
        for (int i = 1; i <= la; ++i) {
            int32_t *cur = buf + (size_t)(i & 1) * si, *prev = buf + (size_t)((i - 1) & 1) * si;
            if (i == 1) memset(prev, 0, si * sizeof(int32_t)); /* PE override: plane 0 is all-zero (PE_1cyc.v:164-178) */
            for (int j = 1; j <= lb; ++j)
                for (int k = 1; k <= lc; ++k) {
                    const unsigned A = a[i - 1] & 3u, B = b[j - 1] & 3u, C = c[k - 1] & 3u;
                    const int32_t sAB = A == B ? M : X, sBC = B == C ? M : X, sAC = A == C ? M : X;
                    const int32_t s3 = rule ? sAB + sBC + sAC
                                            : (A == B ? (B == C ? M * 3 : (int32_t)((uint32_t)(M + X) << 1)) : X * 3);
                    const int32_t add[7] = {s3, 0, 0, 0, sAB, sBC, sAC};
                    for (int m = 0; m < 7; ++m) {
                        const int32_t *src = (kDelta[m][0] ? prev : cur) + (size_t)(j - kDelta[m][1]) * sj +
                                             (size_t)(k - kDelta[m][2]) * sk;
                        int32_t best = INT32_MIN;
                        for (int q = 0; q < 7; ++q) {
                            const int32_t cand = wrapw((uint32_t)src[q] - (uint32_t)costv[kCostSym[m][q]] + (uint32_t)add[m], sh);
                            if (cand > best) best = cand;
                        }
                        cur[(size_t)j * sj + (size_t)k * sk + m] = best;
                    }
                }
        }
        memcpy(fin, buf + (size_t)(la & 1) * si + (size_t)lb * sj + (size_t)lc * sk, sizeof fin);
        free(buf);
    }
    int32_t best = fin[0];
    for (int q = 1; q < 7; ++q)
        if (fin[q] > best) best = fin[q];
    *score = best;
    if (final_states) memcpy(final_states, fin, sizeof fin);
    return 0;
}
