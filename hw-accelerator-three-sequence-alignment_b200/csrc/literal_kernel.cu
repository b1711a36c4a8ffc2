// literal_kernel.cu -- KL: the literal kernel.  One launch per anti-diagonal plane d = i+j+k; every cell evaluates
// the RTL's 49 wrapped candidates and 7 signed maxes exactly as written (src/PE_1cyc.v:164-218, :22-28), for any
// score_bits in [2,32] and any lengths.  It is the always-exact path: the fast kernels fall back to it whenever
// wrap-freedom at score_bits cannot be established.
//
// HBM layout: four rolling planes (d, d-1, d-2, d-3), each [7 states][n triples][lb+1][lc+1], k fastest so that the
// 32 lanes of a warp read consecutive k of one state array (coalesced).  Row j=0 and column k=0 are the RTL's zero
// borders (src/TriAlign_1cyc.v:156-181) and are never written; the i=0 plane is handled by the `first` test, the
// counterpart of the PE's EN_i&&!EN override (src/PE_1cyc.v:164-178).
#include "kernels.cuh"

namespace trialign {

namespace {

// Scoring of one launch: the RTL's localparams (src/PE_1cyc.v:53-59) by default, or the values of the testbench's
// commented-out .i_match / .i_mismatch / .i_gap_open / .i_gap_extend ports (src/TriAlign_tb.sv:220-224).
// cost[m][q] = cost subtracted per source state q (M,Ix,Iy,Iz,Ixy,Iyz,Ixz) for destination Ix,Iy,Iz,Ixy,Iyz,Ixz; rows
// of SURVEY.md 8(a) with GO2 = GO<<1, GE2 = GE<<1, GOGE = GO+GE (src/PE_1cyc.v:57-59).
struct LitScoring {
    int cost[6][7];
    int match, mismatch;
    int s3_all, s3_ab, s3_ne;  // temp_ABC of src/PE_1cyc.v:162: a=b=c / a=b!=c / a!=b
    int sum_of_pairs;          // 1: S(a,b,c) = s(a,b)+s(b,c)+s(a,c) (pic/Algorithm.png) instead of temp_ABC
};

LitScoring make_scoring(const trialign_scoring *sc)
{
    const int M = sc ? sc->match : 1, X = sc ? sc->mismatch : -1, GO = sc ? sc->gap_open : 2, GE = sc ? sc->gap_extend : 1;
    const int GO2 = GO * 2, GE2 = GE * 2, GOGE = GO + GE;
    const int rows[6][7] = {
        {GO2, GE2, GOGE, GOGE, GOGE, GO2, GOGE},  // Ix  <- (i-1,j,k)      :172-178
        {GO2, GOGE, GE2, GOGE, GOGE, GOGE, GO2},  // Iy  <- (i,j-1,k)      :180-186
        {GO2, GOGE, GOGE, GE2, GO2, GOGE, GOGE},  // Iz  <- (i,j,k-1)      :188-194
        {GO, GE, GE, GO, GE, GO, GO},             // Ixy <- (i-1,j-1,k)    :196-202
        {GO, GO, GE, GE, GO, GE, GO},             // Iyz <- (i,j-1,k-1)    :204-210
        {GO, GE, GO, GE, GO, GO, GE},             // Ixz <- (i-1,j,k-1)    :212-218
    };
    LitScoring L;
    for (int m = 0; m < 6; ++m)
        for (int q = 0; q < 7; ++q) L.cost[m][q] = rows[m][q];
    L.match = M;
    L.mismatch = X;
    // src/PE_1cyc.v:162 with Verilog precedence (`+` binds tighter than `<<`): MATCH*3 / (MATCH+MISMATCH)<<1 / MISMATCH*3
    L.s3_all = 3 * M;
    L.s3_ab = (M + X) * 2;
    L.s3_ne = 3 * X;
    L.sum_of_pairs = (sc && sc->triple_rule == TRIALIGN_TRIPLE_SUM_OF_PAIRS) ? 1 : 0;
    return L;
}

__device__ __forceinline__ int wrapw(int x, int sh) { return (int)((unsigned)x << sh) >> sh; }

template <typename T>
__device__ __forceinline__ int max7_from(const T *__restrict__ plane, size_t qstride, size_t idx, bool zero,
                                         const int *cost, int add, int sh)
{
    int best = INT_MIN;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        const int p = zero ? 0 : (int)plane[q * qstride + idx];
        const unsigned raw = (unsigned)p - (unsigned)(cost ? cost[q] : 0) + (unsigned)add;
        const int cand = wrapw((int)raw, sh);  // wrap each candidate, then signed max
        best = cand > best ? cand : best;
    }
    return best;
}

template <typename T>
__global__ void __launch_bounds__(256)
literal_plane_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, const uint8_t *__restrict__ c,
                     int n, int la, int lb, int lc, int d, int sh, const LitScoring sc, T *__restrict__ P0,
                     const T *__restrict__ P1, const T *__restrict__ P2, const T *__restrict__ P3)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * lb * lc;
    if (e >= total) return;
    const int k = (int)(e % lc) + 1;
    const int j = (int)((e / lc) % lb) + 1;
    const int t = (int)(e / ((long long)lb * lc));
    const int i = d - j - k;
    if (i < 1 || i > la) return;

    const size_t row = (size_t)(lc + 1);
    const size_t tri = (size_t)(lb + 1) * row;
    const size_t qstride = (size_t)n * tri;
    const size_t at = (size_t)t * tri + (size_t)j * row + k;  // [t][j][k]

    const unsigned A = a[(size_t)t * la + (i - 1)] & 3u;
    const unsigned B = b[(size_t)t * lb + (j - 1)] & 3u;
    const unsigned C = c[(size_t)t * lc + (k - 1)] & 3u;
    // src/PE_1cyc.v:159-162 (with Verilog precedence: `+` binds tighter than `<<`)
    const int sAB = (A == B) ? sc.match : sc.mismatch;
    const int sBC = (B == C) ? sc.match : sc.mismatch;
    const int sAC = (A == C) ? sc.match : sc.mismatch;
    const int sABC = sc.sum_of_pairs ? sAB + sBC + sAC : ((A == B) ? ((B == C) ? sc.s3_all : sc.s3_ab) : sc.s3_ne);
    const bool first = (i == 1);

    const int vM = max7_from(P3, qstride, at - row - 1, first, nullptr, sABC, sh);    // (i-1,j-1,k-1)
    const int vIx = max7_from(P1, qstride, at, first, sc.cost[0], 0, sh);              // (i-1,j,k)
    const int vIy = max7_from(P1, qstride, at - row, false, sc.cost[1], 0, sh);        // (i,j-1,k)
    const int vIz = max7_from(P1, qstride, at - 1, false, sc.cost[2], 0, sh);          // (i,j,k-1)
    const int vIxy = max7_from(P2, qstride, at - row, first, sc.cost[3], sAB, sh);     // (i-1,j-1,k)
    const int vIyz = max7_from(P2, qstride, at - row - 1, false, sc.cost[4], sBC, sh); // (i,j-1,k-1)
    const int vIxz = max7_from(P2, qstride, at - 1, first, sc.cost[5], sAC, sh);       // (i-1,j,k-1)

    P0[0 * qstride + at] = (T)vM;
    P0[1 * qstride + at] = (T)vIx;
    P0[2 * qstride + at] = (T)vIy;
    P0[3 * qstride + at] = (T)vIz;
    P0[4 * qstride + at] = (T)vIxy;
    P0[5 * qstride + at] = (T)vIyz;
    P0[6 * qstride + at] = (T)vIxz;
}

// Score = MAX7 of the 7 states of cell (la,lb,lc)  (src/TriAlign_1cyc.v:141-142,342-345)
template <typename T>
__global__ void literal_final_kernel(const T *__restrict__ P, int n, int lb, int lc, int32_t *__restrict__ scores,
                                     int32_t *__restrict__ finals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t row = (size_t)(lc + 1), tri = (size_t)(lb + 1) * row, qstride = (size_t)n * tri;
    const size_t at = (size_t)t * tri + (size_t)lb * row + lc;
    int best = INT_MIN;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        const int v = (int)P[q * qstride + at];
        if (finals) finals[(size_t)t * 7 + q] = v;
        best = v > best ? v : best;
    }
    scores[t] = best;
}

template <typename T>
int literal_run_t(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                  int64_t n, int la, int lb, int lc, int score_bits, int32_t *d_scores, int32_t *d_finals,
                  const trialign_scoring *scoring)
{
    const int sh = 32 - score_bits;
    const LitScoring sc = make_scoring(scoring);
    const size_t tri = (size_t)(lb + 1) * (lc + 1);
    const size_t per_triple = 4 * 7 * tri * sizeof(T);
    // chunk the batch so that the four rolling planes stay within ~3 GiB and the flat cell index within int range
    int64_t chunk = (int64_t)((3ull << 30) / per_triple);
    const int64_t idx_cap = (int64_t)((1ull << 31) / ((size_t)lb * lc + 1));
    if (chunk > idx_cap) chunk = idx_cap;
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    int rc = ctx->planes.reserve((size_t)chunk * per_triple);
    if (rc) return rc;
    for (int64_t t0 = 0; t0 < n; t0 += chunk) {
        const int nc = (int)((n - t0 < chunk) ? (n - t0) : chunk);
        const size_t plane_elems = 7 * (size_t)nc * tri;
        T *base = (T *)ctx->planes.ptr;
        TA_CUDA(cudaMemsetAsync(base, 0, 4 * plane_elems * sizeof(T), st));
        const long long cells = (long long)nc * lb * lc;
        const unsigned blocks = (unsigned)((cells + 255) / 256);
        for (int d = 3; d <= la + lb + lc; ++d) {
            literal_plane_kernel<T><<<blocks, 256, 0, st>>>(
                d_a + (size_t)t0 * la, d_b + (size_t)t0 * lb, d_c + (size_t)t0 * lc, nc, la, lb, lc, d, sh, sc,
                base + (size_t)(d & 3) * plane_elems, base + (size_t)((d - 1) & 3) * plane_elems,
                base + (size_t)((d - 2) & 3) * plane_elems, base + (size_t)((d - 3) & 3) * plane_elems);
            count_launch();
        }
        TA_CUDA(cudaGetLastError());
        literal_final_kernel<T><<<(nc + 127) / 128, 128, 0, st>>>(
            base + (size_t)((la + lb + lc) & 3) * plane_elems, nc, lb, lc, d_scores + t0,
            d_finals ? d_finals + (size_t)t0 * 7 : nullptr);
        count_launch();
        TA_CUDA(cudaGetLastError());
    }
    return TRIALIGN_OK;
}

}  // namespace

int literal_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                int64_t n, int la, int lb, int lc, int score_bits, int32_t *d_scores, int32_t *d_finals,
                const trialign_scoring *scoring)
{
    if (n <= 0) return TRIALIGN_OK;
    if (score_bits <= 16)
        return literal_run_t<int16_t>(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, d_scores, d_finals, scoring);
    return literal_run_t<int32_t>(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, d_scores, d_finals, scoring);
}

}  // namespace trialign
