// fast_batch.cu -- K1 launcher: CTA-wide systolic kernel for batches of short triples (see fast_batch.cuh).
#include <cstdlib>

#include "fast_batch.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using Cfg = fast::BatchCfg<4, 4, 16, 16>;

// Split-phase CTA barrier (mbarrier, one arrival per warp): a thread ARRIVES as soon as its tile edges of step s are
// published and WAITS only when it needs its neighbours' edges for step s+1; the work in between (per-pair reset,
// next step's table look-ups, staging of the next pair) hides the barrier latency.
__device__ __forceinline__ void bar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive_warp(uint64_t *bar)
{
    __syncwarp();  // orders the other lanes' shared-memory stores before lane 0's release
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

template <bool SPLIT>
__global__ void __launch_bounds__(Cfg::NT, 1) fast_batch_kernel(fast::BatchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::BatchSmem<Cfg> &S = *reinterpret_cast<fast::BatchSmem<Cfg> *>(smem_raw);
    __shared__ __align__(8) uint64_t step_bar;
    const int tid = threadIdx.x;
    const long long pair_base = (long long)blockIdx.x * A.pairs_per_cta;
    long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    int tj, tk;
    fast::thread_pos<Cfg>(tid, tj, tk);

    if (tid == 0) bar_init(&step_bar, Cfg::NT / 32);
    // threads that are still filling the pipeline (local step < 0) index buffer 1: give it in-range contents
    for (int e = tid; e < 64; e += Cfg::NT) {
        S.acode[1][e] = 0;
        S.bpk[1][e] = 0;
        S.cpk[1][e] = 0;
    }
    for (int e = tid; e < 3 * 2 * 64; e += Cfg::NT) (&S.seq[1][0][0][0])[e] = 0;
    fast::load_pair_sequences<Cfg>(S, A, tid, pair_base, 0);
    __syncthreads();
    fast::build_pair_tables<Cfg>(S, tid, 0);
    __syncthreads();

    fast::BatchRegs<Cfg> R;
    fast::BatchConsts<Cfg> K;
    fast::batch_thread_prepare<Cfg>(R, K, S, A, tj, tk, 0);
    const int steps = fast::batch_total_steps<Cfg>(P);
#pragma unroll 1
    for (int s = 0; s <= steps; ++s) {
        fast::batch_thread_step<Cfg>(R, K, S, A, tj, tk, s, P, pair_base);
        if (SPLIT) bar_arrive_warp(&step_bar);
        fast::batch_uniform_actions<Cfg>(S, A, tid, s, P, pair_base);
        fast::batch_thread_prepare<Cfg>(R, K, S, A, tj, tk, s + 1);
        if (SPLIT) bar_wait(&step_bar, s & 1);
        else __syncthreads();
    }
}

}  // namespace

bool fast_batch_supports(int len) { return len >= 1 && len <= 64; }

int fast_batch_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                   int64_t n, int len, int32_t *d_scores)
{
    if (n <= 0) return TRIALIGN_OK;
    static_assert(sizeof(fast::BatchSmem<Cfg>) <= 227 * 1024, "shared memory budget");
    const size_t smem = sizeof(fast::BatchSmem<Cfg>);
    // One CTA per SM (the 128-register build that would allow two spills).  TRIALIGN_K1_BAR=sync selects the
    // variant with a plain __syncthreads per step for comparison.
    static const int variant = []() {
        const char *e = std::getenv("TRIALIGN_K1_BAR");
        return (e && e[0] == 's') ? 0 : 1;
    }();
    auto kernel = variant == 1 ? fast_batch_kernel<true> : fast_batch_kernel<false>;
    static bool attr_set[kMaxDevices][2] = {};
    if (!attr_set[ctx->device][variant]) {
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        attr_set[ctx->device][variant] = true;
    }
    fast::BatchArgs A;
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = n; A.len = len; A.scores = d_scores;
    A.pairs_total = (n + 1) / 2;
    // one persistent CTA per SM slot; contiguous chunks so that the fill/drain is paid once per CTA
    int occ = 1;
    TA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, Cfg::NT, smem));
    if (occ < 1) occ = 1;
    const long long slots = (long long)ctx->sm_count * occ;
    long long ppc = (A.pairs_total + slots - 1) / slots;
    if (ppc < 1) ppc = 1;
    if (ppc > (1 << 20)) ppc = 1 << 20;  // keep the step counter far from int overflow
    A.pairs_per_cta = (int)ppc;
    A.base = 2 * 64 + 32;  // >= 2*min(len)+8 plus the "-1" headroom of the reductions (common.cuh envelope)
    const long long grid = (A.pairs_total + ppc - 1) / ppc;
    if (std::getenv("TRIALIGN_DEBUG"))
        std::fprintf(stderr, "[trialign] K1: occ=%d grid=%lld pairs/cta=%lld smem=%zu\n", occ, grid, ppc, smem);
    kernel<<<(unsigned)grid, Cfg::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
