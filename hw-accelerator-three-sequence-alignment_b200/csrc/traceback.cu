// traceback.cu -- KT: optimal alignment (traceback) of one triple.  SURVEY.md 8(f).2: the RTL ships score-only but
// carries the hook, MAX7's commented-out `act` output (src/PE_1cyc.v:12,30), which would report WHICH of its seven
// inputs won.  This path defines the alignment as exactly what those `act` signals would select:
//   * forward: the literal plane sweep (one launch per anti-diagonal plane d = i+j+k, 49 wrapped candidates per cell)
//     with every MAX7 evaluated as the RTL's compare-select tree (src/PE_1cyc.v:22-28: `a > b ? a : b`, the second
//     operand wins a tie) over the instance's own port wiring (:139-145); the winning SOURCE STATE of each of the 7
//     destination states is packed into one 21-bit direction word per cell, kept for the whole cube in HBM
//     ([la][lb][lc] uint32, k fastest: coalesced 4-byte stores, 4 B of the ~200 B a cell moves anyway);
//   * backward: one thread walks from FINAL_MAX's choice at (la,lb,lc) (src/TriAlign_1cyc.v:141-142) along the
//     direction words until it reaches a zero border plane (i, j or k = 0; src/TriAlign_1cyc.v:156-181), <= la+lb+lc
//     dependent 4-byte loads.
// Works for any score_bits in [2,32] (it follows the wrapped values, like the RTL would).
// Checkpointed recomputation: when the whole cube (la*lb*lc*4 bytes; 4.3 GB at 1024^3) does not fit the memory budget,
// the sweep is cut into SEGMENTS of D anti-diagonal planes.  Pass A runs score-only and parks the four rolling planes
// at every segment boundary; the direction words of only ONE segment ([D][lb][lc], indexed by plane instead of by i)
// exist at a time: the last segment's come out of pass A itself, every earlier segment is re-swept from its checkpoint
// when the walk reaches it (from the last segment backwards; the walk's position and state live in device memory and
// it stops early on a border plane).  Twice the sweep for O(sqrt) of the memory.
#include <algorithm>
#include <climits>
#include <cmath>

#include "kernels.cuh"

namespace trialign {

namespace {

constexpr int kGO = 2, kGE = 1, kGO2 = 4, kGE2 = 2, kGOGE = 3;  // src/PE_1cyc.v:55-61

// The wiring tables are local constants of forceinline functions: with the destination state and the port unrolled,
// every look-up folds at compile time and a predecessor value is a plain register reference.  (As __constant__ arrays
// they were run-time loads followed by 7-way select chains: ~7x the instructions per cell.)
// state order M,Ix,Iy,Iz,Ixy,Iyz,Ixz; ports in the order G4, G1_A, G1_B, G3_A, G2_A, G2_B, G3_B
__host__ __device__ __forceinline__ int port_src(int dest, int port)
{
    const int t[8][7] = {
        {0, 1, 2, 3, 4, 5, 6},  // M   src/PE_1cyc.v:139
        {1, 0, 5, 2, 4, 6, 3},  // Ix  :140
        {2, 0, 6, 1, 4, 5, 3},  // Iy  :141
        {3, 0, 4, 1, 6, 5, 2},  // Iz  :142
        {2, 0, 6, 5, 4, 1, 3},  // Ixy :143
        {3, 0, 4, 5, 6, 1, 2},  // Iyz :144
        {3, 0, 4, 5, 6, 1, 2},  // Ixz :145
        {6, 0, 1, 4, 2, 3, 5},  // FINAL_MAX src/TriAlign_1cyc.v:141-142
    };
    return t[dest][port];
}
// cost subtracted by destination (row) and source (column); src/PE_1cyc.v:164-218, SURVEY.md 8(a)
__host__ __device__ __forceinline__ int move_cost(int dest, int q)
{
    const int t[7][7] = {
        {0, 0, 0, 0, 0, 0, 0},
        {kGO2, kGE2, kGOGE, kGOGE, kGOGE, kGO2, kGOGE},
        {kGO2, kGOGE, kGE2, kGOGE, kGOGE, kGOGE, kGO2},
        {kGO2, kGOGE, kGOGE, kGE2, kGO2, kGOGE, kGOGE},
        {kGO, kGE, kGE, kGO, kGE, kGO, kGO},
        {kGO, kGO, kGE, kGE, kGO, kGE, kGO},
        {kGO, kGE, kGO, kGE, kGO, kGO, kGE},
    };
    return t[dest][q];
}

__device__ __forceinline__ int wrapw(int x, int sh) { return (int)((unsigned)x << sh) >> sh; }

// index (in port order) of the input the MAX7 tree selects
__device__ __forceinline__ int max7_port(const int g[7])
{
    const int i1 = g[1] > g[2] ? 1 : 2, i2 = g[4] > g[5] ? 4 : 5, i3 = g[3] > g[6] ? 3 : 6;
    const int v1 = g[1] > g[2] ? g[1] : g[2], v2 = g[4] > g[5] ? g[4] : g[5], v3 = g[3] > g[6] ? g[3] : g[6];
    const int i12 = v1 > v2 ? i1 : i2, v12 = v1 > v2 ? v1 : v2;
    const int i34 = v3 > g[0] ? i3 : 0, v34 = v3 > g[0] ? v3 : g[0];
    return v12 > v34 ? i12 : i34;
}

// one destination state: value and winning source state
template <typename T>
__device__ __forceinline__ void dest_state(int dest, const T *__restrict__ plane, size_t qstride, size_t idx, bool zero,
                                           int add, int sh, int &value, unsigned &src)
{
    int p[7], g[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) p[q] = zero ? 0 : (int)plane[q * qstride + idx];
#pragma unroll
    for (int port = 0; port < 7; ++port) {
        const int q = port_src(dest, port);  // compile-time: dest and port are unrolled
        g[port] = wrapw((int)((unsigned)p[q] - (unsigned)move_cost(dest, q) + (unsigned)add), sh);
    }
    const int port = max7_port(g);
    int v = g[0];
    unsigned sq = (unsigned)port_src(dest, 0);
#pragma unroll
    for (int x = 1; x < 7; ++x) {
        v = (x == port) ? g[x] : v;
        sq = (x == port) ? (unsigned)port_src(dest, x) : sq;
    }
    value = v;
    src = sq;
}

// One anti-diagonal plane d = i+j+k.  The grid covers only the rows j that hold cells of the plane (blockIdx.y, from
// j_lo) and, per row, the k interval with 1 <= i <= la (blockIdx.x from the row's first k): about half of the lb x lc
// threads a full-plane launch would start.  T = int16 where score_bits <= 16 (halves the plane traffic).
// DIR = false: score-only (pass A of the checkpointed form).  dir_d0 < 0: direction words indexed [i-1][j-1][k-1] (whole
// cube); dir_d0 >= 0: indexed [d - dir_d0 - 1][j-1][k-1] (one segment of planes dir_d0 < d <= dir_d0 + D).
template <typename T, bool DIR>
__global__ void __launch_bounds__(256)
trace_plane_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, const uint8_t *__restrict__ c, int la,
                   int lb, int lc, int d, int sh, int j_lo, T *__restrict__ P0, const T *__restrict__ P1,
                   const T *__restrict__ P2, const T *__restrict__ P3, uint32_t *__restrict__ dir, int dir_d0)
{
    const int j = j_lo + (int)blockIdx.y;
    if (j > lb) return;
    int k_lo = d - j - la;  // i = d - j - k <= la
    if (k_lo < 1) k_lo = 1;
    const int k = k_lo + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k > lc) return;
    const int i = d - j - k;
    if (i < 1) return;
    const size_t row = (size_t)(lc + 1);
    const size_t qstride = (size_t)(lb + 1) * row;
    const size_t at = (size_t)j * row + k;
    const unsigned A = a[i - 1] & 3u, B = b[j - 1] & 3u, C = c[k - 1] & 3u;
    // src/PE_1cyc.v:159-162 (Verilog precedence: `+` binds tighter than `<<`)
    const int sAB = (A == B) ? 1 : -1, sBC = (B == C) ? 1 : -1, sAC = (A == C) ? 1 : -1;
    const int sABC = (A == B) ? ((B == C) ? 3 : 0) : -3;
    const bool first = (i == 1);  // the i = 0 plane is all-zero (PE override, src/PE_1cyc.v:164-178)
    int v[7];
    unsigned s[7];
    dest_state(0, P3, qstride, at - row - 1, first, sABC, sh, v[0], s[0]);  // M   <- (i-1,j-1,k-1)
    dest_state(1, P1, qstride, at, first, 0, sh, v[1], s[1]);               // Ix  <- (i-1,j,k)
    dest_state(2, P1, qstride, at - row, false, 0, sh, v[2], s[2]);         // Iy  <- (i,j-1,k)
    dest_state(3, P1, qstride, at - 1, false, 0, sh, v[3], s[3]);           // Iz  <- (i,j,k-1)
    dest_state(4, P2, qstride, at - row, first, sAB, sh, v[4], s[4]);       // Ixy <- (i-1,j-1,k)
    dest_state(5, P2, qstride, at - row - 1, false, sBC, sh, v[5], s[5]);   // Iyz <- (i,j-1,k-1)
    dest_state(6, P2, qstride, at - 1, first, sAC, sh, v[6], s[6]);         // Ixz <- (i-1,j,k-1)
    unsigned word = 0;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        P0[q * qstride + at] = (T)v[q];
        word |= s[q] << (3 * q);
    }
    if (DIR) dir[((size_t)(dir_d0 < 0 ? i - 1 : d - dir_d0 - 1) * lb + (j - 1)) * lc + (k - 1)] = word;
}

// out[0] = score, out[1..3] = current cell (the start cell once the walk is over), out[4] = number of columns so far,
// out[5] = current state, out[6] = 1 when a border plane has been reached.  `init`: start from FINAL_MAX's choice at
// (la,lb,lc) in plane buffer Pfin; otherwise continue where the previous segment's walk stopped.  The walk covers
// the cells whose plane i+j+k lies above dir_d0 (everything if dir_d0 < 0); moves[] is filled backwards.
template <typename T>
__global__ void trace_walk_kernel(const uint32_t *__restrict__ dir, const T *__restrict__ Pfin, int la, int lb, int lc,
                                  int init, int dir_d0, int32_t *__restrict__ out, uint8_t *__restrict__ moves)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int i, j, k, n, state;
    if (init) {
        const size_t row = (size_t)(lc + 1), qstride = (size_t)(lb + 1) * row, at = (size_t)lb * row + lc;
        int g[7];
#pragma unroll
        for (int port = 0; port < 7; ++port) g[port] = (int)Pfin[port_src(7, port) * qstride + at];
        const int port = max7_port(g);
        int best = g[0];
        state = port_src(7, 0);
#pragma unroll
        for (int x = 1; x < 7; ++x) {
            best = (x == port) ? g[x] : best;
            state = (x == port) ? port_src(7, x) : state;
        }
        out[0] = best;
        i = la; j = lb; k = lc; n = 0;
    } else {
        i = out[1]; j = out[2]; k = out[3]; n = out[4]; state = out[5];
    }
    while (i > 0 && j > 0 && k > 0 && i + j + k > dir_d0) {
        // M=7 (all advance), Ix=1 (A), Iy=2 (B), Iz=4 (C), Ixy=3, Iyz=6, Ixz=5
        const int mv = state == 0 ? 7 : state == 1 ? 1 : state == 2 ? 2 : state == 3 ? 4 : state == 4 ? 3 : state == 5 ? 6 : 5;
        moves[n++] = (uint8_t)mv;
        const uint32_t w = dir[((size_t)(dir_d0 < 0 ? i - 1 : i + j + k - dir_d0 - 1) * lb + (j - 1)) * lc + (k - 1)];
        i -= mv & 1;
        j -= (mv >> 1) & 1;
        k -= (mv >> 2) & 1;
        state = (int)((w >> (3 * state)) & 7u);
    }
    out[1] = i; out[2] = j; out[3] = k; out[4] = n; out[5] = state;
    out[6] = (i > 0 && j > 0 && k > 0) ? 0 : 1;
}

// the walk collected the columns last to first: forward order
__global__ void trace_reverse_kernel(const int32_t *__restrict__ out, uint8_t *__restrict__ moves)
{
    const int n = out[4];
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n / 2; x += gridDim.x * blockDim.x) {
        const uint8_t t = moves[x];
        moves[x] = moves[n - 1 - x];
        moves[n - 1 - x] = t;
    }
}

}  // namespace

size_t traceback_dir_bytes(int la, int lb, int lc) { return (size_t)la * lb * lc * sizeof(uint32_t); }

namespace {

// planes e0 < d <= e1 of the sweep on the four rolling plane buffers
template <typename T, bool DIR>
int sweep_planes(cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c, int lc, int sh,
                 T *base, size_t plane_elems, int e0, int e1, uint32_t *dir, int dir_d0)
{
    for (int d = e0 + 1; d <= e1; ++d) {
        // rows with cells of this plane: 1 <= i = d-j-k <= la for some k in [1,lc]
        const int j_lo = std::max(1, d - la - lc), j_hi = std::min(lb, d - 2);
        if (j_hi < j_lo) continue;
        // widest k interval of a row: [max(1, d-j-la), min(lc, d-j-1)]
        int width = 0;
        for (int j : {j_lo, j_hi, std::min(j_hi, std::max(j_lo, d - la - 1)), std::min(j_hi, std::max(j_lo, d - lc - 1))}) {
            const int k0 = std::max(1, d - j - la), k1 = std::min(lc, d - j - 1);
            width = std::max(width, k1 - k0 + 1);
        }
        if (width < 1) continue;
        const dim3 grid((unsigned)((width + 255) / 256), (unsigned)(j_hi - j_lo + 1));
        trace_plane_kernel<T, DIR><<<grid, 256, 0, st>>>(d_a, d_b, d_c, la, lb, lc, d, sh, j_lo, base + (size_t)(d & 3) * plane_elems,
                                                         base + (size_t)((d - 1) & 3) * plane_elems,
                                                         base + (size_t)((d - 2) & 3) * plane_elems,
                                                         base + (size_t)((d - 3) & 3) * plane_elems, dir, dir_d0);
        count_launch();
    }
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

template <typename T>
int traceback_run_t(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb,
                    const uint8_t *d_c, int lc, int score_bits, int32_t *d_out, uint8_t *d_moves)
{
    const int sh = 32 - score_bits, total = la + lb + lc;
    const size_t plane_elems = 7 * (size_t)(lb + 1) * (lc + 1), ckpt_bytes = 4 * plane_elems * sizeof(T);
    const size_t seg_plane_bytes = (size_t)lb * lc * sizeof(uint32_t);
    int rc = ctx->planes.reserve(ckpt_bytes);
    if (rc) return rc;
    // memory budget for direction words + checkpoints: option kt_budget_mb, else 85 % of what the device has free now
    // plus what this path already holds
    size_t budget;
    if (opt_int("kt_budget_mb", 0) > 0) {
        budget = (size_t)opt_int("kt_budget_mb", 0) << 20;
    } else if (traceback_dir_bytes(la, lb, lc) <= ctx->dirs.bytes) {
        budget = ctx->dirs.bytes;  // the whole cube fits what an earlier call reserved: no need to ask the driver
    } else {
        size_t free_b = 0, total_b = 0;
        TA_CUDA(cudaMemGetInfo(&free_b, &total_b));
        budget = (size_t)(0.85 * (double)(free_b + ctx->dirs.bytes + ctx->ckpt.bytes));
    }
    // planes per segment: option kt_seg forces it; otherwise the whole cube if it fits, else the D that balances one
    // segment of direction words against total/D checkpoints
    int D = opt_int("kt_seg", 0);
    if (D <= 0) {
        if (traceback_dir_bytes(la, lb, lc) <= budget) {
            D = total;
        } else {
            D = (int)std::ceil(std::sqrt((double)total * (double)ckpt_bytes / (double)seg_plane_bytes));
            D = std::max(D, 8);
        }
    }
    const int first = 2;  // planes 0..2 hold no cell: the state "after plane 2" is all-zero buffers
    const int nseg = D >= total - first ? 1 : (total - first + D - 1) / D;
    auto bound = [&](int s) { return std::min(total, first + s * D); };  // segment s (1..nseg) = planes (bound(s-1), bound(s)]
    T *base = (T *)ctx->planes.ptr;
    TA_CUDA(cudaMemsetAsync(base, 0, ckpt_bytes, st));
    if (nseg == 1) {
        if ((rc = ctx->dirs.reserve(traceback_dir_bytes(la, lb, lc)))) return rc;
        uint32_t *dir = (uint32_t *)ctx->dirs.ptr;
        if ((rc = sweep_planes<T, true>(st, d_a, la, d_b, lb, d_c, lc, sh, base, plane_elems, first, total, dir, -1))) return rc;
        trace_walk_kernel<T><<<1, 32, 0, st>>>(dir, base + (size_t)(total & 3) * plane_elems, la, lb, lc, 1, -1, d_out, d_moves);
        count_launch();
    } else {
        const int n_ckpt = nseg - 2;  // states after segments 1 .. nseg-2 (segment 1 restarts from zeros, the last is never re-swept)
        const size_t need = (size_t)D * seg_plane_bytes + (size_t)std::max(n_ckpt, 0) * ckpt_bytes;
        if (need > budget && opt_int("kt_seg", 0) <= 0) {
            tls().cuda_error = "traceback: direction words of one segment plus the checkpoints exceed the memory budget";
            return TRIALIGN_ENOMEM;
        }
        if ((rc = ctx->dirs.reserve((size_t)D * seg_plane_bytes))) return rc;
        if (n_ckpt > 0 && (rc = ctx->ckpt.reserve((size_t)n_ckpt * ckpt_bytes))) return rc;
        uint32_t *dir = (uint32_t *)ctx->dirs.ptr;
        unsigned char *ck = (unsigned char *)ctx->ckpt.ptr;
        // pass A: score-only up to the last segment, which already writes its direction words
        for (int s = 1; s <= nseg; ++s) {
            if (s < nseg) {
                if ((rc = sweep_planes<T, false>(st, d_a, la, d_b, lb, d_c, lc, sh, base, plane_elems, bound(s - 1), bound(s), nullptr, 0))) return rc;
                if (s <= n_ckpt) TA_CUDA(cudaMemcpyAsync(ck + (size_t)(s - 1) * ckpt_bytes, base, ckpt_bytes, cudaMemcpyDeviceToDevice, st));
            } else {
                if ((rc = sweep_planes<T, true>(st, d_a, la, d_b, lb, d_c, lc, sh, base, plane_elems, bound(s - 1), bound(s), dir, bound(s - 1)))) return rc;
            }
        }
        trace_walk_kernel<T><<<1, 32, 0, st>>>(dir, base + (size_t)(total & 3) * plane_elems, la, lb, lc, 1, bound(nseg - 1), d_out, d_moves);
        count_launch();
        // earlier segments, last to first, until the walk has reached a border plane
        for (int s = nseg - 1; s >= 1; --s) {
            int32_t done = 0;
            TA_CUDA(cudaMemcpyAsync(&done, d_out + 6, sizeof done, cudaMemcpyDeviceToHost, st));
            TA_CUDA(cudaStreamSynchronize(st));
            if (done) break;
            if (s == 1) TA_CUDA(cudaMemsetAsync(base, 0, ckpt_bytes, st));
            else TA_CUDA(cudaMemcpyAsync(base, ck + (size_t)(s - 2) * ckpt_bytes, ckpt_bytes, cudaMemcpyDeviceToDevice, st));
            if ((rc = sweep_planes<T, true>(st, d_a, la, d_b, lb, d_c, lc, sh, base, plane_elems, bound(s - 1), bound(s), dir, bound(s - 1)))) return rc;
            trace_walk_kernel<T><<<1, 32, 0, st>>>(dir, base, la, lb, lc, 0, bound(s - 1), d_out, d_moves);
            count_launch();
        }
    }
    trace_reverse_kernel<<<8, 256, 0, st>>>(d_out, d_moves);
    count_launch();
    TA_CUDA(cudaGetLastError());
    tls().last_segments = nseg;
    return TRIALIGN_OK;
}

}  // namespace

// d_out: 5 int32 (score, start i/j/k, columns); d_moves: la+lb+lc bytes
int traceback_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb,
                  const uint8_t *d_c, int lc, int score_bits, int32_t *d_out, uint8_t *d_moves)
{
    if (score_bits <= 16)
        return traceback_run_t<int16_t>(ctx, st, d_a, la, d_b, lb, d_c, lc, score_bits, d_out, d_moves);
    return traceback_run_t<int32_t>(ctx, st, d_a, la, d_b, lb, d_c, lc, score_bits, d_out, d_moves);
}

}  // namespace trialign
