// fast_batch.cu -- K1 launcher: CTA-wide systolic kernel for batches of short triples (see fast_batch.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fast_batch.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using Cfg256 = fast::BatchCfg<4, 4, 16, 16>;  // 256 threads, 4x4 cells each, <= 255 registers
using Cfg512 = fast::BatchCfg<4, 2, 16, 32>;  // 512 threads, 4x2 cells each, <= 128 registers: 4 warps per scheduler
using Cfg256L = fast::BatchCfg<4, 4, 16, 16, 2>;  // edges consumed two steps after publication (lagged barrier)

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

// MINB = CTAs per SM the register budget is set for: 2 (<= 128 registers, a few words spilled) lets two CTAs cover each
// other's barrier waits and fixed-latency stalls (4 warps per scheduler instead of 2).
template <class Cfg, bool SPLIT, int MINB, bool CHECK, class SC = fast::ScoreRtl>
__device__ __forceinline__ void fast_batch_body(const fast::BatchArgs &A, const fast::BatchCheck *chk, const SC &sc = SC())
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::BatchSmem<Cfg> &S = *reinterpret_cast<fast::BatchSmem<Cfg> *>(smem_raw);
    __shared__ __align__(8) uint64_t step_bar[2];
    const int tid = threadIdx.x;
    const long long pair_base = (long long)blockIdx.x * A.pairs_per_cta;
    long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    int tj, tk;
    fast::thread_pos<Cfg>(tid, tj, tk);

    if (tid == 0) {
        bar_init(&step_bar[0], Cfg::NT / 32);
        bar_init(&step_bar[1], Cfg::NT / 32);
    }
    // threads that are filling / draining the pipeline index buffers no pair was staged into: give them in-range
    // contents
    for (int fill = 1; fill < Cfg::NBUF; ++fill) {
        for (int e = tid; e < 64; e += Cfg::NT) {
            S.acode[fill][e] = 0;
            S.bpk[fill][e] = 0;
            S.cpk[fill][e] = 0;
        }
        for (int e = tid; e < 3 * 2 * 64; e += Cfg::NT) (&S.seq[fill][0][0][0])[e] = 0;
    }
    fast::load_pair_sequences<Cfg, CHECK>(S, A, tid, pair_base, 0);
    __syncthreads();
    fast::build_pair_tables<Cfg, SC>(S, tid, 0, sc);
    if constexpr (Cfg::LAG == 2 && Cfg::DIAG) {
        fast::batch_write_borders_all<Cfg, SC>(S, A, tid, 0, sc);
        fast::batch_write_borders_all<Cfg, SC>(S, A, tid, 1, sc);
    }
    __syncthreads();

    fast::BatchRegs<Cfg> R;
    uint32_t viol = 0;  // CHECK kernels only
    fast::BatchConsts<Cfg> K;
    fast::batch_thread_prepare<Cfg, SC>(R, K, S, A, tj, tk, 0, sc);
    const int steps = fast::batch_total_steps<Cfg>(P);
    if (Cfg::LAG == 2) {
        // Lagged barrier: everybody has finished step s-2 (even and odd steps use separate barriers, so the one waited
        // on is at most one phase ahead of this thread's own arrivals).  Two register sets alternate (R -> R2 -> R) in
        // a loop unrolled by two, so the loop-carried plane state needs no register moves; an odd extra step at the
        // end runs on drained threads only.
        fast::BatchRegs<Cfg> R2;
#define TA_BATCH_STEP(RIN, ROUT, S_)                                                        \
    do {                                                                                   \
        const int s_ = (S_);                                                               \
        if (s_ >= 2) bar_wait(&step_bar[s_ & 1], ((s_ - 2) >> 1) & 1);                     \
        if constexpr (Cfg::LAG == 2 && Cfg::DIAG) fast::batch_write_borders_all<Cfg, SC>(S, A, tid, s_ + 2, sc); \
        fast::batch_thread_step<Cfg, CHECK, SC>(RIN, ROUT, K, S, A, tj, tk, s_, P, pair_base, &viol, chk, sc); \
        bar_arrive_warp(&step_bar[s_ & 1]);                                                \
        fast::batch_uniform_actions<Cfg, CHECK, SC>(S, A, tid, s_, P, pair_base, sc);      \
        fast::batch_thread_prepare<Cfg, SC>(ROUT, K, S, A, tj, tk, s_ + 1, sc);            \
    } while (0)
#pragma unroll 1
        for (int s = 0; s <= steps; s += 2) {
            TA_BATCH_STEP(R, R2, s);
            TA_BATCH_STEP(R2, R, s + 1);
        }
#undef TA_BATCH_STEP
    } else {
#pragma unroll 1
        for (int s = 0; s <= steps; ++s) {
            fast::batch_thread_step<Cfg, CHECK>(R, R, K, S, A, tj, tk, s, P, pair_base, &viol, chk);
            if (SPLIT) bar_arrive_warp(&step_bar[0]);
            fast::batch_uniform_actions<Cfg, CHECK>(S, A, tid, s, P, pair_base);
            fast::batch_thread_prepare<Cfg>(R, K, S, A, tj, tk, s + 1);
            if (SPLIT) bar_wait(&step_bar[0], s & 1);
            else __syncthreads();
        }
    }
}

template <class Cfg, bool SPLIT, int MINB = 1>
__global__ void __launch_bounds__(Cfg::NT, MINB) fast_batch_kernel(fast::BatchArgs A)
{
    fast_batch_body<Cfg, SPLIT, MINB, false>(A, nullptr);
}

// The staging of the next pair runs at 2 of every 32 steps: out of line, so that the hot loop body stays small
// (instruction cache) -- the calls sit on the rarely taken side of their branches.
template <class Cfg>
__device__ __noinline__ void stage_pair_cold(fast::BatchSmem<Cfg> &S, const fast::BatchArgs &A, int tid, int s, int P, long long pair_base)
{
    fast::batch2_uniform_actions<Cfg>(S, A, tid, s, P, pair_base);
}

// Two planes per step (fast_batch.cuh, "P2"): the lagged-barrier layout's four edge slots and three pair buffers, one
// split-phase barrier per TWO planes.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 1) fast_batch2_kernel(fast::BatchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::BatchSmem<Cfg> &S = *reinterpret_cast<fast::BatchSmem<Cfg> *>(smem_raw);
    __shared__ __align__(8) uint64_t step_bar[1];
    const int tid = threadIdx.x;
    const long long pair_base = (long long)blockIdx.x * A.pairs_per_cta;
    long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    int tj, tk;
    fast::thread_pos<Cfg>(tid, tj, tk);
    if (tid == 0) bar_init(&step_bar[0], Cfg::NT / 32);
    for (int fill = 1; fill < Cfg::NBUF; ++fill) {
        for (int e = tid; e < 64; e += Cfg::NT) {
            S.acode[fill][e] = 0;
            S.bpk[fill][e] = 0;
            S.cpk[fill][e] = 0;
        }
        for (int e = tid; e < 3 * 2 * 64; e += Cfg::NT) (&S.seq[fill][0][0][0])[e] = 0;
    }
    fast::load_pair_sequences<Cfg>(S, A, tid, pair_base, 0);
    __syncthreads();
    fast::build_pair_tables<Cfg>(S, tid, 0);
    fast::batch2_write_borders<Cfg>(S, A, tid, 0);
    __syncthreads();
    fast::BatchRegs<Cfg> R, R2;
    fast::BatchConsts<Cfg> K0, K1;
    uint32_t codes = fast::batch2_fetch_codes<Cfg>(S, tj, tk, 0);
    fast::batch2_thread_prepare<Cfg>(R, K0, K1, S, A, tj, tk, 0, codes);
    const int steps = fast::batch2_total_steps<Cfg>(P);
#pragma unroll 1
    for (int s = 0; s <= steps; ++s) {
        fast::batch2_write_borders<Cfg>(S, A, tid, s + 1);  // everybody has finished step s-1, the slots' last reader
        fast::batch2_thread_step<Cfg>(R, R2, K0, K1, S, A, tj, tk, s, P, pair_base);
        bar_arrive_warp(&step_bar[0]);
        {
            const int m = s % fast::kBatch2Steps;
            if (m == fast::kBatch2PrefetchAt || m == fast::kBatch2LoadAt || m == fast::kBatch2BuildAt)
                stage_pair_cold<Cfg>(S, A, tid, s, P, pair_base);
        }
        fast::batch2_thread_prepare<Cfg>(R, K0, K1, S, A, tj, tk, s + 1, codes);
        bar_wait(&step_bar[0], s & 1);  // everybody has finished step s
    }
}
// Two planes per step with neighbour-only synchronisation (fast_batch.cuh: batch2_write_borders_warp): warp w arrives on
// its own barrier (two per warp, even / odd steps, so that a neighbour one step ahead cannot alias the phase parity)
// and waits for warps w-1 and w+1 only.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 1) fast_batch2r_kernel(fast::BatchArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::BatchSmem<Cfg> &S = *reinterpret_cast<fast::BatchSmem<Cfg> *>(smem_raw);
    __shared__ __align__(8) uint64_t ring_bar[Cfg::NT / 32][2];
    const int tid = threadIdx.x, w = tid >> 5;
    const long long pair_base = (long long)blockIdx.x * A.pairs_per_cta;
    long long left = A.pairs_total - pair_base;
    if (left <= 0) return;
    const int P = left < A.pairs_per_cta ? (int)left : A.pairs_per_cta;
    int tj, tk;
    fast::thread_pos<Cfg>(tid, tj, tk);
    if (tid < 2 * (Cfg::NT / 32)) bar_init(&ring_bar[tid >> 1][tid & 1], 1);
    for (int fill = 1; fill < Cfg::NBUF; ++fill) {
        for (int e = tid; e < 64; e += Cfg::NT) {
            S.acode[fill][e] = 0;
            S.bpk[fill][e] = 0;
            S.cpk[fill][e] = 0;
        }
        for (int e = tid; e < 3 * 2 * 64; e += Cfg::NT) (&S.seq[fill][0][0][0])[e] = 0;
    }
    fast::load_pair_sequences<Cfg>(S, A, tid, pair_base, 0);
    __syncthreads();
    fast::build_pair_tables<Cfg>(S, tid, 0);
    fast::batch2_write_borders_warp<Cfg>(S, A, tid, 0);
    __syncthreads();
    fast::BatchRegs<Cfg> R, R2;
    fast::BatchConsts<Cfg> K0, K1;
    uint32_t codes = fast::batch2_fetch_codes<Cfg>(S, tj, tk, 0);
    fast::batch2_thread_prepare<Cfg>(R, K0, K1, S, A, tj, tk, 0, codes);
    const int steps = fast::batch2_total_steps<Cfg>(P);
    uint64_t *const mine = &ring_bar[w][0], *const left_bar = &ring_bar[(w + Cfg::NT / 32 - 1) % (Cfg::NT / 32)][0],
                   *const right_bar = &ring_bar[(w + 1) % (Cfg::NT / 32)][0];
#pragma unroll 1
    for (int s = 0; s <= steps; ++s) {
        fast::batch2_write_borders_warp<Cfg>(S, A, tid, s + 1);  // read by this warp's own edge threads at step s+1
        fast::batch2_thread_step<Cfg>(R, R2, K0, K1, S, A, tj, tk, s, P, pair_base);
        bar_arrive_warp(mine + (s & 1));
        fast::batch2_uniform_actions<Cfg, true>(S, A, tid, s, P, pair_base);
        fast::batch2_thread_prepare<Cfg>(R, K0, K1, S, A, tj, tk, s + 1, codes);
        bar_wait(left_bar + (s & 1), (s >> 1) & 1);   // warp w-1 has finished step s: its edges are there
        bar_wait(right_bar + (s & 1), (s >> 1) & 1);  // warp w+1 has finished step s: it has read what step s+1 overwrites
    }
}
template <class Cfg, bool SPLIT>
__global__ void __launch_bounds__(Cfg::NT, 1) fast_batch_kernel_checked(fast::BatchArgs A, fast::BatchCheck chk)
{
    fast_batch_body<Cfg, SPLIT, 1, true>(A, &chk);
}
// Run-time scoring parameters (trialign_score_batch_ex): the range-checked lagged-barrier kernel with
// the constants of fast::ScoreGen as a third kernel parameter.  SOP: sum-of-pairs triple score.
template <class Cfg, bool SOP>
__global__ void __launch_bounds__(Cfg::NT, 1)
fast_batch_kernel_scored(fast::BatchArgs A, fast::BatchCheck chk, fast::ScoreGen<SOP> sc)
{
    fast_batch_body<Cfg, true, 1, true, fast::ScoreGen<SOP>>(A, &chk, sc);
}

}  // namespace

bool fast_batch_supports(int len) { return len >= 1 && len <= 64; }

namespace {

template <class Cfg, bool SPLIT, int MINB = 1, bool CHECK = false>
int launch_batch(DeviceCtx *ctx, cudaStream_t st, fast::BatchArgs A, int variant, fast::BatchCheck chk = {})
{
    static_assert(MINB * (sizeof(fast::BatchSmem<Cfg>) + 1024 + 64) <= 228 * 1024, "shared memory budget");
    const size_t smem = sizeof(fast::BatchSmem<Cfg>);
    const void *kernel;
    if constexpr (CHECK) kernel = (const void *)fast_batch_kernel_checked<Cfg, SPLIT>;
    else kernel = (const void *)fast_batch_kernel<Cfg, SPLIT, MINB>;
    static bool attr_set[kMaxDevices] = {};
    if (!attr_set[ctx->device]) {
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        attr_set[ctx->device] = true;
    }
    // one persistent CTA per SM; contiguous chunks so that the fill/drain is paid once per CTA
    const long long slots = (long long)ctx->sm_count * MINB;
    long long ppc = (A.pairs_total + slots - 1) / slots;
    if (ppc < 1) ppc = 1;
    if (ppc > (1 << 20)) ppc = 1 << 20;  // keep the step counter far from int overflow
    A.pairs_per_cta = (int)ppc;
    const long long grid = (A.pairs_total + ppc - 1) / ppc;
    if (opt_debug()) {
        int occ = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, Cfg::NT, smem);
        std::fprintf(stderr, "[trialign] K1: variant=%d threads=%d grid=%lld pairs/cta=%lld smem=%zu ctas/sm=%d\n", variant,
                     Cfg::NT, grid, ppc, smem, occ);
    }
    if constexpr (CHECK) fast_batch_kernel_checked<Cfg, SPLIT><<<(unsigned)grid, Cfg::NT, smem, st>>>(A, chk);
    else fast_batch_kernel<Cfg, SPLIT, MINB><<<(unsigned)grid, Cfg::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

int launch_batch2(DeviceCtx *ctx, cudaStream_t st, fast::BatchArgs A, bool ring)
{
    using Cfg = Cfg256L;  // its shared-memory layout: four edge slots, three pair buffers
    const size_t smem = sizeof(fast::BatchSmem<Cfg>);
    static bool attr_set[kMaxDevices] = {};
    if (!attr_set[ctx->device]) {
        TA_CUDA(cudaFuncSetAttribute(fast_batch2r_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TA_CUDA(cudaFuncSetAttribute(fast_batch2r_kernel<Cfg>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        TA_CUDA(cudaFuncSetAttribute(fast_batch2_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TA_CUDA(cudaFuncSetAttribute(fast_batch2_kernel<Cfg>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
        attr_set[ctx->device] = true;
    }
    long long ppc = (A.pairs_total + ctx->sm_count - 1) / ctx->sm_count;
    if (ppc < 1) ppc = 1;
    if (ppc > (1 << 20)) ppc = 1 << 20;
    A.pairs_per_cta = (int)ppc;
    const long long grid = (A.pairs_total + ppc - 1) / ppc;
    if (opt_debug())
        std::fprintf(stderr, "[trialign] K1 P2: grid=%lld pairs/cta=%lld smem=%zu\n", grid, ppc, smem);
    if (ring) fast_batch2r_kernel<Cfg><<<(unsigned)grid, Cfg::NT, smem, st>>>(A);
    else fast_batch2_kernel<Cfg><<<(unsigned)grid, Cfg::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace

bool score_fast_domain(const trialign_scoring &sc, int score_bits)
{
    const int M = sc.match, X = sc.mismatch, GO = sc.gap_open, GE = sc.gap_extend;
    if (GE < 0 || GE > 4 || GO < GE || GO > 8 || M > 8 || X < -8 || X > M) return false;
    if (X + 3 * GE < 0 || 3 * X + 6 * GE < 0) return false;  // pair / triple constants in potential form
    if (sc.triple_rule == TRIALIGN_TRIPLE_RTL) {
        // temp_ABC = min(row constant, column constant) needs 3X <= 2(M+X) <= 3M; all three + 6 GE non-negative
        if (!(3 * X <= 2 * (M + X) && 2 * (M + X) <= 3 * M) || 2 * (M + X) + 6 * GE < 0) return false;
    }
    if (score_bits < 6) return false;
    if (score_bits <= 16) {  // the window the range check accepts must not be empty
        const fast::ScoreMargins m = fast::score_margins(M, X);
        const int lim = 1 << (score_bits - 1);
        if (lim - 1 - m.add_max < 8 || -lim + 2 * (2 * GO - m.add_min) > -8) return false;
    }
    return true;
}

int fast_batch_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                   int64_t n, int len, int32_t *d_scores, int check_bits, int *d_pair_flags, const trialign_scoring *scoring)
{
    if (n <= 0) return TRIALIGN_OK;
    if (scoring) {
        if (check_bits <= 0 || !d_pair_flags || !score_fast_domain(*scoring, check_bits)) return TRIALIGN_EINVAL;
        fast::BatchArgs A;
        A.a = d_a; A.b = d_b; A.c = d_c;
        A.n = n; A.len = len; A.scores = d_scores;
        A.pairs_total = (n + 1) / 2;
        A.pairs_per_cta = 1;
        const int M = scoring->match, X = scoring->mismatch, GO = scoring->gap_open, GE = scoring->gap_extend;
        A.base = fast::score_base(X, GO, GE);
        const fast::BatchCheck chk = fast::score_check(M, X, GO, check_bits, d_pair_flags);
        using Cfg = Cfg256L;
        const size_t smem = sizeof(fast::BatchSmem<Cfg>);
        const bool sop = scoring->triple_rule == TRIALIGN_TRIPLE_SUM_OF_PAIRS;
        const void *kernel = sop ? (const void *)fast_batch_kernel_scored<Cfg, true> : (const void *)fast_batch_kernel_scored<Cfg, false>;
        static bool attr_set[kMaxDevices][2] = {};
        if (!attr_set[ctx->device][sop]) {
            TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
            attr_set[ctx->device][sop] = true;
        }
        long long ppc = (A.pairs_total + ctx->sm_count - 1) / ctx->sm_count;
        if (ppc < 1) ppc = 1;
        if (ppc > (1 << 20)) ppc = 1 << 20;
        A.pairs_per_cta = (int)ppc;
        const unsigned grid = (unsigned)((A.pairs_total + ppc - 1) / ppc);
        if (sop) fast_batch_kernel_scored<Cfg, true><<<grid, Cfg::NT, smem, st>>>(A, chk, fast::make_score_gen<true>(M, X, GO, GE));
        else fast_batch_kernel_scored<Cfg, false><<<grid, Cfg::NT, smem, st>>>(A, chk, fast::make_score_gen<false>(M, X, GO, GE));
        count_launch();
        TA_CUDA(cudaGetLastError());
        return TRIALIGN_OK;
    }
    // option k1 selects a variant for comparison: "256" / "512" threads per CTA; suffix "m" = split-phase mbarrier
    // instead of __syncthreads, "l" = lagged barrier (LAG 2), "p2" = two planes per step, e.g. k1=512m.  Default: 256p2.
    // (A two-CTAs-per-SM build, 128 registers, ran at exactly the speed of its one-CTA form in round 2 and is gone.)
    int variant = 16;
    {
        char e[8] = {0};
        if (trialign_get_option("k1", e, sizeof e) == TRIALIGN_OK && e[0] && std::strcmp(e, "stream") != 0) {
            variant = 0;
            if (e[0] == '5') variant = 2;
            if (e[1] && e[2] && e[3] == 'm') variant |= 1;
            if (e[1] && e[2] && e[3] == 'l') variant |= 4;
            if (e[1] && e[2] && e[3] == 'p' && e[4] == '2') variant = 16; // 256p2: two planes per step
            if (e[1] && e[2] && e[3] == 'r' && e[4] == '2') variant = 32; // 256r2: the same with neighbour-only synchronisation
        }
    }
    fast::BatchArgs A;
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = n; A.len = len; A.scores = d_scores;
    A.pairs_total = (n + 1) / 2;
    A.pairs_per_cta = 1;
    A.base = 2 * 64 + 32;  // >= 2*min(len)+8 plus the "-1" headroom of the reductions (common.cuh envelope)
    if (check_bits > 0) {
        // candidates reach all+3 and (min all)-8; the margins cover the unsampled planes between two samples
        // (`all` moves by at most +3 / -4 per plane)
        const int lim = 1 << (check_bits - 1), gap = fast::kBatchCheckEvery - 1;
        fast::BatchCheck chk;
        chk.check_hi = lim - 1 - 3 - 3 * gap;
        chk.check_lo = -lim + 8 + 4 * gap;
        chk.pair_flags = d_pair_flags;
        return launch_batch<Cfg256L, true, 1, true>(ctx, st, A, 100, chk);
    }
    if (variant == 16 || variant == 32) return launch_batch2(ctx, st, A, variant == 32);
    switch (variant) {
    case 1: return launch_batch<Cfg256, true>(ctx, st, A, variant);
    case 2: return launch_batch<Cfg512, false>(ctx, st, A, variant);
    case 3: return launch_batch<Cfg512, true>(ctx, st, A, variant);
    case 4: return launch_batch<Cfg256L, true>(ctx, st, A, variant);
    default: return launch_batch<Cfg256, false>(ctx, st, A, variant);
    }
}

}  // namespace trialign
