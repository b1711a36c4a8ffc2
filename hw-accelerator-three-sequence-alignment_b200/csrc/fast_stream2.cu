// fast_stream2.cu -- launcher of K2S "P2": two planes per step (see fast_stream2.cuh).  One persistent CTA per SM sweeps
// all blocks of its contiguous chunk of pairs back to back.
#include <algorithm>
#include <cstdio>

#include "fast_batch.cuh"  // host helpers of the run-time scoring parameters
#include "fast_stream2.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using C2 = fast::Stream2Cfg;

// split-phase CTA barrier, one arrival per warp (as in fast_batch.cu)
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

// The staging of the next task runs at 2 of every L2 steps: out of line, so that the hot loop body stays small
// (instruction cache) -- the call sits on the rarely taken side of its branch.
template <class SC>
__device__ __noinline__ void stage_task_cold(fast::StreamSmem<C2> &S, const fast::StreamArgs &A, int tid, int cta, int P,
                                             long long pair_base, int g_t, int g_off, const SC &sc)
{
    fast::stream2_uniform_actions<C2, SC>(S, A, tid, cta, P, pair_base, g_t, g_off, sc);
}

template <bool CHECK, class SC = fast::ScoreRtl>
__device__ __forceinline__ void fast_stream2_body(const fast::StreamArgs &A, const fast::RangeCheck *chk, const SC &sc = SC())
{
    using C = C2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::StreamSmem<C> &S = *reinterpret_cast<fast::StreamSmem<C> *>(smem_raw);
    __shared__ __align__(8) uint64_t step_bar[1];
    const int tid = threadIdx.x, cta = blockIdx.x;
    const long long pair_base = (long long)cta * A.pairs_per_cta;
    const long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    const int tasks = P * A.nbj * A.nbk, L2 = fast::stream2_steps_per_task(A);
    int tj, tk;
    fast::stream_thread_pos<C>(tid, tj, tk);

    if (tid == 0) bar_init(&step_bar[0], C::NT / 32);
    fast::stream_stage_load<C>(S, A, tid, cta, P, pair_base, -1);  // filler the pipeline fill runs on
    fast::stream_stage_load<C>(S, A, tid, cta, P, pair_base, 0);
    __syncthreads();
    fast::stream_stage_build<C, SC>(S, tid, -1, sc);
    fast::stream_stage_build<C, SC>(S, tid, 0, sc);
    __syncthreads();

    fast::StreamRegs<C> R, R2;
    uint32_t viol = 0;  // CHECK kernels only
    fast::StreamConsts<C> K0, K1;
    fast::Stream2Pos pos = fast::stream2_pos_at<C>(A, tj + tk, 0);
    // the edge threads this thread serves: IN -> (0,t) / (t,0), skew t, one step ahead; OUT -> (15,t) / (t,15), skew
    // 15 + t, one step back (starting with step 0 at the top of step 1)
    fast::Stream2Pos pin = fast::stream2_pos_at<C>(A, tid & 15, 0), pout = fast::stream2_pos_at<C>(A, C::TJ - 1 + (tid & 15), 0);
    fast::stream2_duty_in<C, SC>(S, A, tid, pin, 0, sc);
    fast::stream2_advance<C>(A, pin);
    fast::ta_copy_commit();
    fast::ta_copy_wait();
    __syncthreads();
    fast::stream2_thread_prepare<C, SC>(R, K0, K1, S, A, tj, tk, pos, sc);
    const int steps = (int)fast::stream2_total_steps<C>(A, P);
    int g_t = 0, g_off = 0;
#pragma unroll 1
    for (int s = 0; s <= steps; ++s) {
        // everybody has finished step s-1: halos of step s+1 in, faces of step s-1 out
        fast::stream2_duty_in<C, SC>(S, A, tid, pin, s + 1, sc);
        fast::stream2_advance<C>(A, pin);
        if (s >= 1) {
            fast::stream2_duty_out<C, SC>(S, tid, pout, s - 1, sc);
            fast::stream2_advance<C>(A, pout);
        }
        fast::ta_copy_commit();
        fast::stream2_thread_step<C, CHECK, SC>(R, R2, K0, K1, S, A, tj, tk, s, pos, &viol, chk, sc);
        fast::ta_copy_wait();
        bar_arrive_warp(&step_bar[0]);
        if (g_t < tasks && (g_off == C::LOAD_AT || g_off == C::BUILD_AT))
            stage_task_cold<SC>(S, A, tid, cta, P, pair_base, g_t, g_off, sc);
        if (++g_off == L2) {
            g_off = 0;
            ++g_t;
        }
        fast::stream2_advance<C>(A, pos);
        fast::stream2_thread_prepare<C, SC>(R, K0, K1, S, A, tj, tk, pos, sc);
        bar_wait(&step_bar[0], s & 1);  // everybody has finished step s
    }
}

__global__ void __launch_bounds__(C2::NT, 1) fast_stream2_kernel(fast::StreamArgs A) { fast_stream2_body<false>(A, nullptr); }
__global__ void __launch_bounds__(C2::NT, 1) fast_stream2_kernel_checked(fast::StreamArgs A, fast::RangeCheck chk)
{
    fast_stream2_body<true>(A, &chk);
}
// run-time scoring parameters (trialign_score_batch_ex): range-checked on every plane, constants as a kernel argument
template <bool SOP>
__global__ void __launch_bounds__(C2::NT, 1)
fast_stream2_kernel_scored(fast::StreamArgs A, fast::RangeCheck chk, fast::ScoreGen<SOP> sc)
{
    fast_stream2_body<true, fast::ScoreGen<SOP>>(A, &chk, sc);
}

template <class K>
int set_smem_attr(K kernel, size_t smem)
{
    TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    return TRIALIGN_OK;
}

}  // namespace

bool fast_stream2_supports(int len) { return len >= C2::LEN_MIN && len <= C2::LEN_MAX && fast_long_supports(len + 1, len, len); }

// pairs per CTA / grid as for the one-plane kernel (the step counter stays far below 2^30)
int fast_stream2_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                     int64_t n, int len, int32_t *d_scores, int check_bits, int *d_pair_flags, const trialign_scoring *scoring)
{
    using C = C2;
    if (n <= 0) return TRIALIGN_OK;
    static_assert(sizeof(fast::StreamSmem<C>) + 64 <= 227 * 1024, "shared memory budget");
    const size_t smem = sizeof(fast::StreamSmem<C>);
    static bool attr_set[kMaxDevices] = {};
    if (!attr_set[ctx->device]) {
        int rc0;
        if ((rc0 = set_smem_attr(fast_stream2_kernel, smem))) return rc0;
        if ((rc0 = set_smem_attr(fast_stream2_kernel_checked, smem))) return rc0;
        if ((rc0 = set_smem_attr(fast_stream2_kernel_scored<false>, smem))) return rc0;
        if ((rc0 = set_smem_attr(fast_stream2_kernel_scored<true>, smem))) return rc0;
        attr_set[ctx->device] = true;
    }
    fast::StreamArgs A{};
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = n; A.len = len; A.scores = d_scores;
    A.nbj = (len + C::BJ - 1) / C::BJ;
    A.nbk = (len + C::BK - 1) / C::BK;
    A.pairs_total = (n + 1) / 2;
    const long long nb = A.nbj, steps_per_pair = nb * nb * ((len + 1) / 2);
    long long ppc = (A.pairs_total + ctx->sm_count - 1) / ctx->sm_count;
    ppc = std::max<long long>(1, std::min<long long>(ppc, (1ll << 30) / steps_per_pair));
    const long long grid = (A.pairs_total + ppc - 1) / ppc;
    A.pairs_per_cta = (int)ppc;
    // face rings are private to a CTA; rows 0 .. len+1 (an odd len runs one padding plane)
    A.fj_slot_words = (long long)(len + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(len + 2) * C::FK_ROW_WORDS;
    A.fk_ring = A.nbj + 1;
    const size_t fj_bytes = (size_t)grid * 2 * A.fj_slot_words * 4;
    const size_t fk_bytes = (size_t)grid * A.fk_ring * A.fk_slot_words * 4;
    int rc = ctx->faces.reserve(fj_bytes + fk_bytes + 1024);
    if (rc) return rc;
    A.face_j = (uint32_t *)ctx->faces.ptr;
    A.face_k = (uint32_t *)((char *)ctx->faces.ptr + ((fj_bytes + 255) & ~(size_t)255));
    A.base = (int)(-envelope_lo((long long)len + 1) + 32);
    if (opt_debug())
        std::fprintf(stderr, "[trialign] K2S P2: len=%d blocks %dx%d grid=%lld pairs/cta=%lld faces=%.1f MB smem=%zu base=%d\n",
                     len, A.nbj, A.nbk, grid, ppc, (fj_bytes + fk_bytes) / 1e6, smem, A.base);
    if (scoring) {
        if (check_bits <= 0 || !d_pair_flags || !fast_stream_score_ok(*scoring, len + 1 <= C::LEN_MAX ? len + 1 : len, check_bits))
            return TRIALIGN_EINVAL;
        const int M = scoring->match, X = scoring->mismatch, GO = scoring->gap_open, GE = scoring->gap_extend;
        A.base = fast::score_base_len(X, GO, GE, len + 1);
        const fast::BatchCheck bc = fast::score_check(M, X, GO, check_bits, d_pair_flags);
        fast::RangeCheck chk;
        chk.check_hi = bc.check_hi; chk.check_lo = bc.check_lo; chk.pair_flags = d_pair_flags;
        if (scoring->triple_rule == TRIALIGN_TRIPLE_SUM_OF_PAIRS)
            fast_stream2_kernel_scored<true><<<(unsigned)grid, C::NT, smem, st>>>(A, chk, fast::make_score_gen<true>(M, X, GO, GE));
        else
            fast_stream2_kernel_scored<false><<<(unsigned)grid, C::NT, smem, st>>>(A, chk, fast::make_score_gen<false>(M, X, GO, GE));
    } else if (check_bits > 0) {
        // every 8th plane is sampled; the margins cover the planes in between (`all` moves by at most +3 / -4 per plane)
        const int lim = 1 << (check_bits - 1), gap = 7;
        fast::RangeCheck chk;
        chk.check_hi = lim - 1 - 3 - 3 * gap;
        chk.check_lo = -lim + 8 + 4 * gap;
        chk.pair_flags = d_pair_flags;
        fast_stream2_kernel_checked<<<(unsigned)grid, C::NT, smem, st>>>(A, chk);
    } else {
        fast_stream2_kernel<<<(unsigned)grid, C::NT, smem, st>>>(A);
    }
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
