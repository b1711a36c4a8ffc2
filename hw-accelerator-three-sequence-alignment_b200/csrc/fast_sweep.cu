// fast_sweep.cu -- K2 launchers: block-sweep kernels for long triples (see fast_sweep.cuh).
//   sweep_long_kernel : one long triple; persistent CTAs claim (k-block-pair, j-block) tasks from a global counter
//                       and stream block faces through global memory, gated by per-task progress counters.
//   sweep_batch_kernel: many triples of equal length > 64; every CTA sweeps all blocks of its pairs one after the
//                       other, faces stay private to the CTA, no inter-CTA synchronisation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fast_sweep.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

using LongCfg = fast::SweepCfg<true>;
using BatchCfg = fast::SweepCfg<false>;

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// system scope: the flag / data may live in a peer GPU's memory (i-slab pipeline over NVLink)
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class C>
__device__ __forceinline__ void setup_task(fast::SweepSmem<C> &S, fast::SweepRegs<C> &R, const fast::SweepArgs &A,
                                           int tid, int tj, int tk)
{
    fast::sweep_load_bc<C>(S, A, tid);
    __syncthreads();
    fast::sweep_build_tables<C>(S, tid);
    fast::sweep_init_thread<C>(R, S, tj, tk);
    __syncthreads();
}

template <bool CHECK>
__global__ void __launch_bounds__(LongCfg::NT, 1) sweep_long_kernel(fast::SweepArgs A)
{
    using C = LongCfg;
    // The thread that polls / releases the progress counters sits in warp 3 (tiles tj 4..7, tk 8..15): that warp has no
    // block-edge duty (no halo fetch, no left-edge merge, few face stores), so the MEMBAR of a release eats its slack
    // instead of delaying the warp every barrier already waits for.
    constexpr int kGateTid = 96;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::SweepSmem<C> &S = *reinterpret_cast<fast::SweepSmem<C> *>(smem_raw);
    const int tid = threadIdx.x;
    int tj, tk;
    fast::sweep_thread_pos<C>(tid, tj, tk);
    fast::sweep_load_a<C>(S, A, tid, 0);
    fast::SweepRegs<C> R;
    fast::SweepConsts<C> K;
    const int steps = fast::sweep_task_steps<C>(A);

    for (;;) {
        __syncthreads();  // everybody is done with the previous task descriptor
        if (tid == kGateTid) {
            const unsigned long long t = atomicAdd(A.counter, 1ull);
            if (t < (unsigned long long)A.tasks_total) {
                fast::sweep_describe_long<C>(A, (long long)t, S.task);
                // ring re-use: the previous tenants of the face slots must have been consumed
                // (t is the claim position; the ring sizes exceed the largest producer->consumer distance in claim
                // order by more than the grid size, so every consumer waited for here was claimed long ago)
                if ((long long)t >= A.fj_ring) {
                    const int prev = A.order[t - A.fj_ring];
                    if (prev % A.nbj + 1 < A.nbj)
                        while (ld_acquire(A.progress + prev + 1) < fast::kProgressDone) __nanosleep(200);
                }
                if ((long long)t >= A.fk_ring) {
                    const int prev = A.order[t - A.fk_ring];
                    if (prev / A.nbj + 1 < A.nbkp)
                        while (ld_acquire(A.progress + prev + A.nbj) < fast::kProgressDone) __nanosleep(200);
                }
                // i-slab pipeline: the previous slab (another GPU) must have delivered this task's plane-0 state
                if (A.iface_in_flag)
                    while (ld_acquire_sys(A.iface_in_flag + S.task.self_task) != A.epoch) __nanosleep(500);
            } else {
                S.task.valid = 0;
            }
        }
        __syncthreads();
        if (!S.task.valid) break;
        setup_task<C>(S, R, A, tid, tj, tk);
        const fast::SweepTask &T = S.task;
        const int *top_p = T.top_task >= 0 ? A.progress + T.top_task : nullptr;
        const int *left_p = T.left_task >= 0 ? A.progress + T.left_task : nullptr;

        // prologue: halos of step 0.  cert = last fetch step whose face rows are known to be complete.
        int cert = -2;
        if (tid == kGateTid) {
            for (;;) {
                cert = fast::sweep_certified<C>(A, T, top_p ? ld_acquire(top_p) : 0, left_p ? ld_acquire(left_p) : 0);
                if (cert >= -1) break;
                __nanosleep(100);
            }
        }
        __syncthreads();
        fast::sweep_edge_fetch<C>(S, A, tj, tk, -1);
        fast::sweep_thread_prepare<C>(R, K, S, A, tj, tk, 0, tid);
        fast::ta_copy_wait();
        __syncthreads();

#pragma unroll 1
        for (int s = 0; s <= steps; ++s) {
            // the fetch of step s+1 must be certified before the barrier that ends step s; the producers' progress is
            // sampled at the top of the step, so the load latency hides under the tile
            const bool need = tid == kGateTid && cert < s + 1;
            int tp = 0, lp = 0;
            if (need) {
                tp = top_p ? ld_acquire(top_p) : 0;
                lp = left_p ? ld_acquire(left_p) : 0;
            }
            fast::sweep_edge_fetch<C>(S, A, tj, tk, s);  // rows certified during the previous steps
            fast::sweep_thread_step<C, CHECK>(R, K, S, A, tj, tk, s, tid);
            fast::sweep_thread_prepare<C>(R, K, S, A, tj, tk, s + 1, tid);
            fast::ta_copy_wait();
            if (need) {
                cert = fast::sweep_certified<C>(A, T, tp, lp);
                while (cert < s + 1) {
                    __nanosleep(40);
                    cert = fast::sweep_certified<C>(A, T, top_p ? ld_acquire(top_p) : 0, left_p ? ld_acquire(left_p) : 0);
                }
            }
            __syncthreads();
            // release: the edge threads' face stores of the steps so far are ordered before the barrier above
            if (tid == kGateTid && fast::sweep_publish_now<C>(A, s)) st_release(A.progress + T.self_task, s + 1);
        }
        if (CHECK && R.viol) atomicOr(A.viol_flag, 1);
        if (tid == kGateTid) {
            st_release(A.progress + T.self_task, fast::kProgressDone);
            if (A.iface_out_flag) {  // every thread's hand-over stores precede the last barrier of the step loop
                __threadfence_system();
                st_release_sys(A.iface_out_flag + T.self_task, A.epoch);
            }
        }
    }
}

__global__ void __launch_bounds__(BatchCfg::NT, 1) sweep_batch_kernel(fast::SweepArgs A)
{
    using C = BatchCfg;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::SweepSmem<C> &S = *reinterpret_cast<fast::SweepSmem<C> *>(smem_raw);
    const int tid = threadIdx.x;
    int tj, tk;
    fast::sweep_thread_pos<C>(tid, tj, tk);
    fast::SweepRegs<C> R;
    fast::SweepConsts<C> K;
    const int steps = fast::sweep_task_steps<C>(A);
    const int tasks = A.nbj * A.nbkp;

    for (long long pair = blockIdx.x; pair < A.pairs_total; pair += gridDim.x) {
        __syncthreads();
        fast::sweep_load_a<C>(S, A, tid, pair);
        for (int q = 0; q < tasks; ++q) {
            __syncthreads();
            if (tid == 0) fast::sweep_describe_batch<C>(A, (int)blockIdx.x, pair, q, S.task);
            __syncthreads();
            setup_task<C>(S, R, A, tid, tj, tk);
            fast::sweep_edge_fetch<C>(S, A, tj, tk, -1);
            fast::sweep_thread_prepare<C>(R, K, S, A, tj, tk, 0, tid);
            fast::ta_copy_wait();
            __syncthreads();
#pragma unroll 1
            for (int s = 0; s <= steps; ++s) {
                fast::sweep_edge_fetch<C>(S, A, tj, tk, s);
                fast::sweep_thread_step<C, false>(R, K, S, A, tj, tk, s, tid);
                fast::sweep_thread_prepare<C>(R, K, S, A, tj, tk, s + 1, tid);
                fast::ta_copy_wait();
                __syncthreads();
            }
        }
    }
}

template <class K>
int prepare_kernel(K kernel, size_t smem, bool &done)
{
    if (!done) {
        TA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    return TRIALIGN_OK;
}

}  // namespace

// 16-bit envelope of the fast kernels: stored value = v + 2*(i + jl + kl) + base, block-local jl,kl <= 64/128.
size_t fast_sweep_slab_state_bytes(int lb, int lc)
{
    const long long tasks = (long long)((lb + LongCfg::BJ - 1) / LongCfg::BJ) * ((lc + LongCfg::KSPAN - 1) / LongCfg::KSPAN);
    return (size_t)tasks * 4 * 16 * LongCfg::NT * sizeof(uint32_t);
}
size_t fast_sweep_slab_flag_bytes(int lb, int lc)
{
    const long long tasks = (long long)((lb + LongCfg::BJ - 1) / LongCfg::BJ) * ((lc + LongCfg::KSPAN - 1) / LongCfg::KSPAN);
    return (size_t)tasks * sizeof(int);
}

bool fast_long_range_ok(int planes, int la_total, int lb, int lc)
{
    if (planes < 1 || la_total < planes || lb < 1 || lc < 1) return false;
    const long long m = la_total < lb ? (la_total < lc ? la_total : lc) : (lb < lc ? lb : lc);
    const long long base = -envelope_lo(m) + 32;
    const long long top = base + envelope_hi(m) + 9 + 2 * ((long long)planes + LongCfg::LAG + LongCfg::DEPTH + 64 + 128) + 600;
    return top < 65535;
}

bool fast_sweep_supports(int la, int lb, int lc)
{
    if (la < 1 || lb < 1 || lc < 1 || la > LongCfg::LA_MAX - 64) return false;
    return fast_long_range_ok(la, la, lb, lc);
}

static int sweep_base(int la, int lb, int lc)
{
    const long long m = la < lb ? (la < lc ? la : lc) : (lb < lc ? lb : lc);
    return (int)(-envelope_lo(m) + 32);
}

int fast_sweep_long_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb,
                        const uint8_t *d_c, int lc, int32_t *d_score, int32_t *d_states, int check_bits, int *d_viol,
                        const SlabLink *slab)
{
    using C = LongCfg;
    const size_t smem = sizeof(fast::SweepSmem<C>);
    static bool prepared[kMaxDevices][2] = {};
    const bool check = check_bits > 0;
    int rc = check ? prepare_kernel(sweep_long_kernel<true>, smem, prepared[ctx->device][1])
                   : prepare_kernel(sweep_long_kernel<false>, smem, prepared[ctx->device][0]);
    if (rc) return rc;
    fast::SweepArgs A{};
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = 1; A.la = la; A.lb = lb; A.lc = lc;
    if (check) {
        // candidates reach all+3 and (min all)-8; 45 / 60 cover the 15 unsampled planes between two samples
        const int lim = 1 << (check_bits - 1);
        A.check_hi = lim - 1 - 3 - 45;
        A.check_lo = -lim + 8 + 60;
        A.viol_flag = d_viol;
    }
    A.nbj = (lb + C::BJ - 1) / C::BJ;
    A.nbkp = (lc + C::KSPAN - 1) / C::KSPAN;
    A.tasks_total = (long long)A.nbj * A.nbkp;
    // long sweeps release their progress less often (see fast_sweep.cuh); option k2_gate overrides for experiments
    A.gate = la >= 4096 ? 16 : 8;
    {
        const int g = opt_int("k2_gate", 0);
        if (g >= 1 && g <= 64 && (g & (g - 1)) == 0) A.gate = g;
    }
    int grid = ctx->sm_count;  // one persistent CTA per SM; all co-resident (spin-waits need that)
    if (grid > A.tasks_total) grid = (int)A.tasks_total;
    A.fj_slot_words = (long long)(la + C::LAG + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(la + C::LAG + 2) * C::FK_ROW_WORDS;
    // claim order (earliest possible start first) and the face-ring sizes it needs
    std::vector<int> order((size_t)A.tasks_total), pos_of((size_t)A.tasks_total);
    for (int t = 0; t < (int)A.tasks_total; ++t) order[t] = t;
    const int nbj = A.nbj;
    std::stable_sort(order.begin(), order.end(), [nbj](int x, int y) {
        return fast::sweep_order_key(x % nbj, x / nbj) < fast::sweep_order_key(y % nbj, y / nbj);
    });
    for (int p = 0; p < (int)A.tasks_total; ++p) pos_of[order[p]] = p;
    int dist_j = 0, dist_k = 0;  // largest producer -> consumer distance in claim order
    for (int t = 0; t < (int)A.tasks_total; ++t) {
        if (t % nbj + 1 < nbj) dist_j = std::max(dist_j, pos_of[t + 1] - pos_of[t]);
        if (t / nbj + 1 < A.nbkp) dist_k = std::max(dist_k, pos_of[t + nbj] - pos_of[t]);
    }
    A.fj_ring = (int)std::min<long long>(A.tasks_total, (long long)dist_j + 2 * grid + 2);
    A.fk_ring = (int)std::min<long long>(A.tasks_total, (long long)dist_k + 2 * grid + 2);
    const size_t fj_bytes = (size_t)A.fj_ring * A.fj_slot_words * 4, fk_bytes = (size_t)A.fk_ring * A.fk_slot_words * 4;
    const size_t prog_bytes = (((size_t)A.tasks_total * sizeof(int) + 255) & ~(size_t)255);
    const size_t ctl_bytes = 256 + prog_bytes;           // task counter, progress[]
    const size_t ord_bytes = 2 * prog_bytes;             // order[], pos_of[]
    if ((rc = ctx->faces.reserve(fj_bytes + fk_bytes + 512))) return rc;
    if ((rc = ctx->misc.reserve(ctl_bytes + ord_bytes))) return rc;
    A.face_j = (uint32_t *)ctx->faces.ptr;
    A.face_k = (uint32_t *)((char *)ctx->faces.ptr + ((fj_bytes + 255) & ~(size_t)255));
    A.counter = (unsigned long long *)ctx->misc.ptr;
    A.progress = (int *)((char *)ctx->misc.ptr + 256);
    int *d_order = (int *)((char *)ctx->misc.ptr + ctl_bytes);
    int *d_pos = (int *)((char *)ctx->misc.ptr + ctl_bytes + prog_bytes);
    A.order = d_order;
    A.pos_of = d_pos;
    TA_CUDA(cudaMemcpyAsync(d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_pos, pos_of.data(), pos_of.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    A.out = d_score;
    A.out_states = d_states;
    A.base = sweep_base(la, lb, lc);
    A.last_slab = 1;
    if (slab) {  // planes i_begin+1 .. i_end of a triple of slab->la_total planes
        A.base = sweep_base(slab->la_total, lb, lc);
        A.iface_in = (const uint32_t *)slab->in_state;
        A.iface_in_flag = (const int *)slab->in_flags;
        A.iface_out = (uint32_t *)slab->out_state;
        A.iface_out_flag = (int *)slab->out_flags;
        A.last_slab = slab->last ? 1 : 0;
        A.epoch = slab->epoch;
    }
    TA_CUDA(cudaMemsetAsync(ctx->misc.ptr, 0, ctl_bytes, st));
    if (opt_debug())
        std::fprintf(stderr, "[trialign] K2 long: la=%d blocks %dx%d(pairs) tasks=%lld grid=%d faces=%.1f MB base=%d\n", la,
                     A.nbj, A.nbkp, A.tasks_total, grid, (fj_bytes + fk_bytes) / 1e6, A.base);
    if (check) sweep_long_kernel<true><<<grid, C::NT, smem, st>>>(A);
    else sweep_long_kernel<false><<<grid, C::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

int fast_sweep_batch_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                         int64_t n, int len, int32_t *d_scores)
{
    using C = BatchCfg;
    if (n <= 0) return TRIALIGN_OK;
    const size_t smem = sizeof(fast::SweepSmem<C>);
    static bool prepared[kMaxDevices] = {};
    int rc = prepare_kernel(sweep_batch_kernel, smem, prepared[ctx->device]);
    if (rc) return rc;
    fast::SweepArgs A{};
    A.a = d_a; A.b = d_b; A.c = d_c;
    A.n = n; A.la = A.lb = A.lc = len;
    A.nbj = (len + C::BJ - 1) / C::BJ;
    A.nbkp = (len + C::KSPAN - 1) / C::KSPAN;
    A.pairs_total = (n + 1) / 2;
    int grid = (int)std::min<long long>(A.pairs_total, ctx->sm_count);
    A.fj_slot_words = (long long)(len + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(len + 2) * C::FK_ROW_WORDS;
    A.fj_ring = 2;
    A.fk_ring = A.nbj + 1;
    const size_t fj_bytes = (size_t)grid * A.fj_ring * A.fj_slot_words * 4;
    const size_t fk_bytes = (size_t)grid * A.fk_ring * A.fk_slot_words * 4;
    if ((rc = ctx->faces.reserve(fj_bytes + fk_bytes + 512))) return rc;
    A.face_j = (uint32_t *)ctx->faces.ptr;
    A.face_k = (uint32_t *)((char *)ctx->faces.ptr + ((fj_bytes + 255) & ~(size_t)255));
    A.out = d_scores;
    A.base = sweep_base(len, len, len);
    if (opt_debug())
        std::fprintf(stderr, "[trialign] K2 batch: len=%d blocks %dx%d grid=%d faces=%.1f MB base=%d\n", len, A.nbj, A.nbkp,
                     grid, (fj_bytes + fk_bytes) / 1e6, A.base);
    sweep_batch_kernel<<<grid, C::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
