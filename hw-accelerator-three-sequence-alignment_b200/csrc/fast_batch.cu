// fast_batch.cu -- K1 launcher: CTA-wide systolic kernel for batches of short triples (see fast_batch.cuh).
#include <cstdlib>

#include "fast_batch.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using Cfg = fast::BatchCfg<4, 4, 16, 16>;

template <int MINBLOCKS>
__global__ void __launch_bounds__(Cfg::NT, MINBLOCKS) fast_batch_kernel(fast::BatchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::BatchSmem<Cfg> &S = *reinterpret_cast<fast::BatchSmem<Cfg> *>(smem_raw);
    const int tid = threadIdx.x;
    const long long pair_base = (long long)blockIdx.x * A.pairs_per_cta;
    long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    int tj, tk;
    fast::thread_pos<Cfg>(tid, tj, tk);

    // threads that are still filling the pipeline (local step < 0) index buffer 1: give it in-range contents
    for (int e = tid; e < 64; e += Cfg::NT) S.acode[1][e] = 0;
    for (int e = tid; e < 3 * 2 * 64; e += Cfg::NT) (&S.seq[1][0][0][0])[e] = 0;
    fast::load_pair_sequences<Cfg>(S, A, tid, pair_base, 0);
    __syncthreads();
    fast::build_pair_tables<Cfg>(S, tid, 0);
    __syncthreads();

    fast::BatchRegs<Cfg> R;
    const int steps = fast::batch_total_steps<Cfg>(P);
#pragma unroll 1
    for (int s = 0; s <= steps; ++s) {
        fast::batch_uniform_actions<Cfg>(S, A, tid, s, P, pair_base);
        fast::batch_thread_step<Cfg>(R, S, A, tj, tk, s, P, pair_base);
        __syncthreads();
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
    // One 190-register CTA per SM is the measured optimum (1.10 TCUPS); the 128-register build that allows two
    // CTAs per SM spills and measured 0.83 TCUPS.  TRIALIGN_K1_OCC=2 selects it for comparison.
    static const int want_occ = []() {
        const char *e = std::getenv("TRIALIGN_K1_OCC");
        return (e && e[0] == '2') ? 2 : 1;
    }();
    auto kernel = want_occ == 1 ? fast_batch_kernel<1> : fast_batch_kernel<2>;
    static bool attr_set[kMaxDevices][3] = {};
    if (!attr_set[ctx->device][want_occ]) {
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        attr_set[ctx->device][want_occ] = true;
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
