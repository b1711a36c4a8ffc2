// fast_warp.cu -- K1w launcher: eight asynchronous warp-wide systolic arrays per CTA (see fast_warp.cuh).
#include <cstdio>
#include <cstdlib>

#include "fast_warp.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using C = fast::WarpCfg;

__device__ __forceinline__ int ld_vol(const int *p) { return *(const volatile int *)p; }

// Only the bottom row travels to lane+LK and only the right column (+ relay) to lane+1: 16 + 18 shuffles per step.
__device__ __forceinline__ void shfl_up_part(fast::LaneOut &dst, const fast::LaneOut &src, bool right_part)
{
    const unsigned m = 0xFFFFFFFFu;
    if (!right_part) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dst.b_wY[q] = __shfl_up_sync(m, src.b_wY[q], C::LK);
            dst.b_u2[q] = __shfl_up_sync(m, src.b_u2[q], C::LK);
            dst.b_u1[q] = __shfl_up_sync(m, src.b_u1[q], C::LK);
            dst.b_all[q] = __shfl_up_sync(m, src.b_all[q], C::LK);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dst.r_wZ[q] = __shfl_up_sync(m, src.r_wZ[q], 1);
            dst.r_u2[q] = __shfl_up_sync(m, src.r_u2[q], 1);
            dst.r_u3[q] = __shfl_up_sync(m, src.r_u3[q], 1);
            dst.r_all[q] = __shfl_up_sync(m, src.r_all[q], 1);
        }
        dst.relay_u2 = __shfl_up_sync(m, src.relay_u2, 1);
        dst.relay_all = __shfl_up_sync(m, src.relay_all, 1);
    }
}

__global__ void __launch_bounds__(C::NT, 1) fast_warp_kernel(fast::WarpArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::WarpSmem &S = *reinterpret_cast<fast::WarpSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int a = wid / C::WB, b = wid % C::WB;
    const int lj = lane / C::LK, lk = lane % C::LK;
    const long long pair_base = (long long)blockIdx.x * A.pairs_per_cta;
    long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    const int last = fast::warp_total_steps(P);
    fast::WarpShared &W = S.w[wid];

    if (lane == 0) {
        S.produced[a][b] = 0;
        S.consumed_top[a][b] = 0;
        S.consumed_left[a][b] = 0;
    }
    for (int e = lane; e < 64; e += 32) W.acode[1][e] = 0;  // lanes still filling the pipeline index buffer 1
    for (int e = lane; e < 2 * C::WJ; e += 32) (&W.seqB[1][0][0])[e] = 0;
    for (int e = lane; e < 2 * C::WK; e += 32) (&W.seqC[1][0][0])[e] = 0;
    fast::warp_load_pair(W, A, lane, a, b, pair_base, 0);
    __syncwarp();
    fast::warp_build_tables(W, lane, 0);
    __syncthreads();  // the only CTA-wide barrier: ring counters are initialised

    fast::LaneRegs R;
    fast::LaneOut O, up, lf;
    const int *prod_top = a > 0 ? &S.produced[a - 1][b] : nullptr;
    const int *prod_left = b > 0 ? &S.produced[a][b - 1] : nullptr;
    const int *cons_below = a + 1 < C::WA ? &S.consumed_top[a + 1][b] : nullptr;
    const int *cons_right = b + 1 < C::WB ? &S.consumed_left[a][b + 1] : nullptr;

#pragma unroll 1
    for (int w = 0; w <= last; ++w) {
        // every lane polls the same four counters (broadcast reads): the loop stays warp-uniform, so the warp does not
        // split into "lane 0" and "the rest" (a split warp pushes every shuffle below onto a slow per-lane path)
        while (!fast::warp_can_step(S, a, b, w, last, prod_top ? ld_vol(prod_top) : 0, prod_left ? ld_vol(prod_left) : 0,
                                    cons_below ? ld_vol(cons_below) : 0, cons_right ? ld_vol(cons_right) : 0)) {
        }
        __threadfence_block();  // acquire: ring data written before the producer's counter update
        __syncwarp();
        if (fast::warp_rebuild_step(w)) {
            const int q = (w >> 6) + 1;
            if (q < P) {
                fast::warp_load_pair(W, A, lane, a, b, pair_base + q, q & 1);
                __syncwarp();
                fast::warp_build_tables(W, lane, q & 1);
                __syncwarp();
            }
        }
        fast::warp_lane_step(R, S, A, a, b, lj, lk, w, P, pair_base, last, O);
        __syncwarp();  // every lane has read its ring entries and written its own; the warp is converged here, which
                       // keeps the shuffles below on the hardware path (a diverged warp falls back to a per-lane loop)
        shfl_up_part(up, O, false);  // from lane - LK: the tile above (bottom row)
        shfl_up_part(lf, O, true);   // from lane - 1: the tile to the left (right column + relay)
        fast::warp_take_halo(R.H, up, lf);
        __threadfence_block();  // release: the lanes' ring stores above (every lane fences its own stores)
        __syncwarp();
        // all lanes store the same values (benign): no divergent region around the counter update
        if (a > 0) *(volatile int *)&S.consumed_top[a][b] = w + C::TOP_LEAD + 1;
        if (b > 0) *(volatile int *)&S.consumed_left[a][b] = w + C::LEFT_LEAD + 1;
        *(volatile int *)&S.produced[a][b] = w + 1;
    }
}

}  // namespace

int fast_warp_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c, int64_t n,
                  int len, int32_t *d_scores)
{
    if (n <= 0) return TRIALIGN_OK;
    static_assert(sizeof(fast::WarpSmem) <= 227 * 1024, "shared memory budget");
    const size_t smem = sizeof(fast::WarpSmem);
    static bool attr_set[kMaxDevices] = {};
    if (!attr_set[ctx->device]) {
        TA_CUDA(cudaFuncSetAttribute(fast_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device] = true;
    }
    fast::WarpArgs A;
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = n; A.len = len; A.scores = d_scores;
    A.pairs_total = (n + 1) / 2;
    const long long slots = ctx->sm_count;
    long long ppc = (A.pairs_total + slots - 1) / slots;
    if (ppc < 1) ppc = 1;
    if (ppc > (1 << 20)) ppc = 1 << 20;
    A.pairs_per_cta = (int)ppc;
    A.base = 2 * 64 + 32;
    const long long grid = (A.pairs_total + ppc - 1) / ppc;
    if (std::getenv("TRIALIGN_DEBUG"))
        std::fprintf(stderr, "[trialign] K1w: grid=%lld pairs/cta=%lld smem=%zu\n", grid, ppc, smem);
    fast_warp_kernel<<<(unsigned)grid, C::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
