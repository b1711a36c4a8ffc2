// fast_stream.cu -- K2S launcher: streaming block-sweep kernel for batches of triples longer than 64
// (see fast_stream.cuh).  One persistent CTA per SM sweeps all blocks of its contiguous chunk of pairs back to back.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "fast_batch.cuh"  // host helpers of the run-time scoring parameters
#include "fast_stream.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using Cfg = fast::StreamCfg;

// Split-phase CTA barrier, one arrival per warp (as in fast_batch.cu): a warp ARRIVES once its tile edges of step s are
// published and WAITS, two steps later, for everybody's step s.
__device__ __forceinline__ void bar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive_warp(uint64_t *bar)
{
    __syncwarp();  // orders the other lanes' shared-memory / global stores before lane 0's release
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

template <class C, bool CHECK, class SC = fast::ScoreRtl>
__device__ __forceinline__ void fast_stream_body(const fast::StreamArgs &A, const fast::RangeCheck *chk, const SC &sc = SC())
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::StreamSmem<C> &S = *reinterpret_cast<fast::StreamSmem<C> *>(smem_raw);
    __shared__ __align__(8) uint64_t step_bar[2];
    const int tid = threadIdx.x, cta = blockIdx.x;
    const long long pair_base = (long long)cta * A.pairs_per_cta;
    const long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    const int tasks = P * A.nbj * A.nbk;
    int tj, tk;
    fast::stream_thread_pos<C>(tid, tj, tk);

    if (tid == 0) {
        bar_init(&step_bar[0], C::NT / 32);
        bar_init(&step_bar[1], C::NT / 32);
    }
    fast::stream_stage_load<C>(S, A, tid, cta, P, pair_base, -1);  // filler the pipeline fill runs on
    fast::stream_stage_load<C>(S, A, tid, cta, P, pair_base, 0);
    __syncthreads();
    fast::stream_stage_build<C, SC>(S, tid, -1, sc);
    fast::stream_stage_build<C, SC>(S, tid, 0, sc);
    __syncthreads();

    fast::StreamRegs<C> R, R2;
    uint32_t viol = 0;  // CHECK kernel only
    fast::StreamConsts<C> K;
    fast::StreamPos pos = fast::stream_start_pos<C>(A, tj, tk);
    // DIST: the edge thread this thread serves every step -- fetch (threads 0..127): its position two steps ahead,
    // store (threads 128..255): two steps back, starting with step 0 at the top of step 2
    fast::StreamPos duty = fast::stream_pos_at<C>(A, C::LAG * (tid < 128 ? (tid & 15) : C::TJ - 1 + (tid & 15)), 0);
    if constexpr (C::DIST) {
        if (tid < 128) {
            fast::stream_duty_fetch<C, SC>(S, A, tid, duty, 0, sc);
            fast::stream_advance(A, duty, C::NBUF);
            fast::stream_duty_fetch<C, SC>(S, A, tid, duty, 1, sc);
            fast::stream_advance(A, duty, C::NBUF);
        }
        fast::ta_copy_commit();
        fast::ta_copy_wait();
        __syncthreads();
    } else {
        if (tj == 0 || tk == 0) fast::stream_edge_fetch<C, SC>(S, A, tj, tk, pos, (0 + C::LAG) & (C::NSLOT - 1), sc);
        fast::ta_copy_commit();
    }
    fast::stream_thread_prepare<C, SC>(R, K, S, A, tj, tk, pos, 0, false, sc);
    const int steps = (int)fast::stream_total_steps<C>(A, P) | 1;  // an even number of steps: the loop runs two per turn
    int g_t = 0, g_off = 0;
    // one step: wait until everybody has finished step s-2 (even and odd steps use separate barriers) and this thread's
    // block-halo fetch for step s has landed | step | arrive | staging, next step's constants and halo fetch
#define TA_STREAM_STEP(RIN, ROUT, S_)                                                              \
    do {                                                                                          \
        const int s_ = (S_);                                                                      \
        if (s_ >= 2) bar_wait(&step_bar[s_ & 1], ((s_ - 2) >> 1) & 1);                            \
        if constexpr (C::DIST) {                                                                  \
            if (tid < 128) {                                                                      \
                fast::stream_duty_fetch<C, SC>(S, A, tid, duty, s_ + 2, sc);                      \
                fast::stream_advance(A, duty, C::NBUF);                                           \
            } else if (s_ >= 2) {                                                                 \
                fast::stream_duty_store<C, SC>(S, A, tid, duty, s_ - 2, sc);                      \
                fast::stream_advance(A, duty, C::NBUF);                                           \
            }                                                                                     \
            fast::ta_copy_commit();                                                               \
        } else {                                                                                  \
            fast::ta_copy_wait_prev();                                                            \
        }                                                                                         \
        fast::stream_thread_step<C, CHECK, SC>(RIN, ROUT, K, S, A, tj, tk, s_, pos, &viol, chk, sc); \
        if constexpr (C::DIST) fast::ta_copy_wait();                                              \
        bar_arrive_warp(&step_bar[s_ & 1]);                                                       \
        if (g_t < tasks) fast::stream_uniform_actions<C, SC>(S, A, tid, cta, P, pair_base, g_t, g_off, sc); \
        fast::stream_thread_prepare<C, SC>(ROUT, K, S, A, tj, tk, pos, s_ + 1, true, sc);         \
        if (++g_off == A.len) {                                                                   \
            g_off = 0;                                                                            \
            ++g_t;                                                                                \
        }                                                                                         \
    } while (0)
#pragma unroll 1
    for (int s = 0; s <= steps; s += 2) {
        // two register sets alternate (R -> R2 -> R), so the loop-carried plane state needs no register moves
        TA_STREAM_STEP(R, R2, s);
        TA_STREAM_STEP(R2, R, s + 1);
    }
#undef TA_STREAM_STEP
}

// Two plain kernels around the one body (the unchecked one is the default path; the checked one adds the range test).
__global__ void __launch_bounds__(Cfg::NT, 1) fast_stream_kernel(fast::StreamArgs A) { fast_stream_body<Cfg, false>(A, nullptr); }
__global__ void __launch_bounds__(Cfg::NT, 1) fast_stream_kernel_checked(fast::StreamArgs A, fast::RangeCheck chk)
{
    fast_stream_body<Cfg, true>(A, &chk);
}
// the block-edge duties dealt to all threads (fast_stream.cuh: StreamCfgT<true>)
__global__ void __launch_bounds__(Cfg::NT, 1) fast_stream_kernel_dist(fast::StreamArgs A)
{
    fast_stream_body<fast::StreamCfgD, false>(A, nullptr);
}
// Run-time scoring parameters (trialign_score_batch_ex): the range-checked kernel with the constants
// of fast::ScoreGen as a third kernel parameter.  SOP: sum-of-pairs triple score.
template <bool SOP>
__global__ void __launch_bounds__(Cfg::NT, 1)
fast_stream_kernel_scored(fast::StreamArgs A, fast::RangeCheck chk, fast::ScoreGen<SOP> sc)
{
    fast_stream_body<Cfg, true, fast::ScoreGen<SOP>>(A, &chk, sc);
}

// contiguous chunks of pairs, one persistent CTA per SM (fill/drain once per CTA); the step counter stays below 2^30
void stream_plan(int len, int64_t n, int sm_count, long long &ppc, long long &grid)
{
    const long long nb = (len + Cfg::BJ - 1) / Cfg::BJ, pairs = (n + 1) / 2;
    const long long steps_per_pair = nb * nb * len;
    ppc = (pairs + sm_count - 1) / sm_count;
    ppc = std::max<long long>(1, std::min<long long>(ppc, (1ll << 30) / steps_per_pair));
    grid = (pairs + ppc - 1) / ppc;
}

long long stream_face_bytes(int len, long long grid)
{
    const int nbj = (len + Cfg::BJ - 1) / Cfg::BJ;
    const long long fj = (long long)(len + 2) * Cfg::FJ_ROW_WORDS, fk = (long long)(len + 2) * fast::StreamCfgD::FK_ROW_WORDS;
    return grid * (2 * fj + (nbj + 1) * fk) * 4 + 1024;
}

}  // namespace

bool fast_stream_supports(int len)
{
    return len >= Cfg::LEN_MIN && len <= Cfg::LEN_MAX && fast_long_supports(len, len, len);
}

// 16-bit lane range with run-time parameters: the potential base, pi per unit of (absolute i + block-local j, k) and the
// largest value a cell can hold (three-way matches all along) must stay below 2^16
bool fast_stream_score_ok(const trialign_scoring &sc, int len, int score_bits)
{
    if (!score_fast_domain(sc, score_bits) || len < Cfg::LEN_MIN || len > Cfg::LEN_MAX) return false;
    const long long pi = 2 * sc.gap_extend;
    const long long top = fast::score_base_len(sc.mismatch, sc.gap_open, sc.gap_extend, len) + pi * (len + Cfg::BJ + Cfg::BK) +
                          3ll * std::max(sc.match, 0) * len + 64;
    return top < 65536;
}

int fast_stream_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                    int64_t n, int len, int32_t *d_scores, int check_bits, int *d_pair_flags, const trialign_scoring *scoring)
{
    using C = Cfg;
    if (n <= 0) return TRIALIGN_OK;
    // the default: two planes per step (fast_stream2.cu); options k2s=p1 / k2s=dist keep the one-plane forms
    if (fast_stream2_supports(len) && !opt_is("k2s", "p1") && !opt_is("k2s", "dist") &&
        (!scoring || fast_stream_score_ok(*scoring, len + 1 <= C::LEN_MAX ? len + 1 : len, check_bits)))
        return fast_stream2_run(ctx, st, d_a, d_b, d_c, n, len, d_scores, check_bits, d_pair_flags, scoring);
    static_assert(sizeof(fast::StreamSmem<C>) + 64 <= 227 * 1024, "shared memory budget");
    static_assert(sizeof(fast::StreamSmem<fast::StreamCfgD>) + 64 <= 227 * 1024, "shared memory budget");
    // option k2s=dist: the block-edge duties dealt to all threads (A/B form of the unchecked kernel; measured 2 % slower
    // than the edge-thread form, 1,268 vs 1,296 GCUPS on 32,768 x 256^3, so it is not the default)
    const bool dist = !scoring && check_bits <= 0 && opt_is("k2s", "dist");
    const size_t smem = dist ? sizeof(fast::StreamSmem<fast::StreamCfgD>) : sizeof(fast::StreamSmem<C>);
    static bool attr_set[kMaxDevices] = {};
    if (!attr_set[ctx->device]) {
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(fast::StreamSmem<C>)));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_dist, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(fast::StreamSmem<fast::StreamCfgD>)));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_dist, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_checked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(fast::StreamSmem<C>)));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_checked, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_scored<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(fast::StreamSmem<C>)));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_scored<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_scored<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(fast::StreamSmem<C>)));
        TA_CUDA(cudaFuncSetAttribute(fast_stream_kernel_scored<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        attr_set[ctx->device] = true;
    }
    fast::StreamArgs A{};
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = n; A.len = len; A.scores = d_scores;
    A.nbj = (len + C::BJ - 1) / C::BJ;
    A.nbk = (len + C::BK - 1) / C::BK;
    A.pairs_total = (n + 1) / 2;
    long long ppc, grid;
    stream_plan(len, n, ctx->sm_count, ppc, grid);
    A.pairs_per_cta = (int)ppc;
    // face rings are private to a CTA; CTAs beyond one wave reuse nothing, so the rings are sized by the grid
    A.fj_slot_words = (long long)(len + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(len + 2) * (dist ? fast::StreamCfgD::FK_ROW_WORDS : C::FK_ROW_WORDS);
    A.fk_ring = A.nbj + 1;
    const size_t fj_bytes = (size_t)grid * 2 * A.fj_slot_words * 4;
    const size_t fk_bytes = (size_t)grid * A.fk_ring * A.fk_slot_words * 4;
    int rc = ctx->faces.reserve(fj_bytes + fk_bytes + 1024);
    if (rc) return rc;
    A.face_j = (uint32_t *)ctx->faces.ptr;
    A.face_k = (uint32_t *)((char *)ctx->faces.ptr + ((fj_bytes + 255) & ~(size_t)255));
    const long long m = len;
    A.base = (int)(-envelope_lo(m) + 32);
    if (opt_debug())
        std::fprintf(stderr, "[trialign] K2S stream: len=%d blocks %dx%d grid=%lld pairs/cta=%lld faces=%.1f MB smem=%zu base=%d\n",
                     len, A.nbj, A.nbk, grid, ppc, (fj_bytes + fk_bytes) / 1e6, smem, A.base);
    if (scoring) {
        if (check_bits <= 0 || !d_pair_flags || !fast_stream_score_ok(*scoring, len, check_bits)) return TRIALIGN_EINVAL;
        const int M = scoring->match, X = scoring->mismatch, GO = scoring->gap_open, GE = scoring->gap_extend;
        A.base = fast::score_base_len(X, GO, GE, len);
        const fast::BatchCheck bc = fast::score_check(M, X, GO, check_bits, d_pair_flags);
        fast::RangeCheck chk;
        chk.check_hi = bc.check_hi; chk.check_lo = bc.check_lo; chk.pair_flags = d_pair_flags;
        if (scoring->triple_rule == TRIALIGN_TRIPLE_SUM_OF_PAIRS)
            fast_stream_kernel_scored<true><<<(unsigned)grid, C::NT, smem, st>>>(A, chk, fast::make_score_gen<true>(M, X, GO, GE));
        else
            fast_stream_kernel_scored<false><<<(unsigned)grid, C::NT, smem, st>>>(A, chk, fast::make_score_gen<false>(M, X, GO, GE));
    } else if (check_bits > 0) {
        const int lim = 1 << (check_bits - 1), gap = fast::kStreamCheckEvery - 1;
        fast::RangeCheck chk;
        chk.check_hi = lim - 1 - 3 - 3 * gap;
        chk.check_lo = -lim + 8 + 4 * gap;
        chk.pair_flags = d_pair_flags;
        fast_stream_kernel_checked<<<(unsigned)grid, C::NT, smem, st>>>(A, chk);
    } else {
        if (dist) fast_stream_kernel_dist<<<(unsigned)grid, C::NT, smem, st>>>(A);
        else fast_stream_kernel<<<(unsigned)grid, C::NT, smem, st>>>(A);
    }
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

// the streaming kernel is used while its face rings stay small next to the device memory
bool fast_stream_fits(int len, int64_t n, int sm_count)
{
    long long ppc, grid;
    stream_plan(len, n, sm_count, ppc, grid);
    return stream_face_bytes(len, grid) <= (8ll << 30);
}

}  // namespace trialign
