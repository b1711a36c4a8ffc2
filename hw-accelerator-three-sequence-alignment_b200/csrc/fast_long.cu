// fast_long.cu -- K2L launcher: block-sweep kernel for one long triple on the K1/K2S step engine (see fast_long.cuh).
// Persistent CTAs claim (k-block-pair, j-block) tasks from a global counter in earliest-start order and stream block
// faces through global memory, gated by per-task progress counters (release/acquire at GPU scope); the i-slab pipeline
// over several GPUs hands the plane state of every task to the next GPU through peer memory (system scope).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fast_long.cuh"
#include "kernels.cuh"

namespace trialign {

namespace {

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

// Polling without ordering: a relaxed load does not hold back the memory operations that follow it, so the gate warp
// runs its step while the value is in flight (an acquire load is an LDG plus an L1 invalidate that waits for it: the
// gate warp -- and with it the CTA's next barrier -- stood still for an L2 round trip at every poll: 9.8 % of the
// stall samples of the 8192^3 sweep).  The acquire fence follows once, when a new value is actually relied upon.
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_relaxed_sys(const int *p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acquire_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

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

// progress word of the column pipeline: epoch << 20 | steps (0xFFFFF = task finished); a word of another epoch reads 0
__device__ __forceinline__ int col_decode(int v, int epoch)
{
    if ((v >> 20) != epoch) return 0;
    const int steps = v & 0xFFFFF;
    return steps == 0xFFFFF ? fast::kProgressDone : steps;
}
__device__ __forceinline__ int col_encode(int steps, int epoch)
{
    return (epoch << 20) | (steps >= fast::kProgressDone ? 0xFFFFF : steps);
}

// COLS: the column-pipeline variant (faces and progress words of the range edges in peer memory); a separate
// instantiation, so that the single-GPU / i-slab kernels compile exactly as without it.
template <int LAG, bool CHECK, bool COLS = false>
__global__ void __launch_bounds__(256, 1) long2_kernel(fast::SweepArgs A)
{
    using C = fast::Long2Cfg<LAG>;
    // The thread that polls / releases the progress counters (A.gate_tid, see fast_long2_run)
    const int kGateTid = A.gate_tid;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fast::Long2Smem<C> &S = *reinterpret_cast<fast::Long2Smem<C> *>(smem_raw);
    __shared__ __align__(8) uint64_t step_bar[2];
    const int tid = threadIdx.x;
    int tj, tk;
    fast::stream_thread_pos<C>(tid, tj, tk);
    if (tid == 0) {
        bar_init(&step_bar[0], C::NT / 32);
        bar_init(&step_bar[1], C::NT / 32);
    }
    fast::long2_load_a<C>(S, A, tid);
    fast::Long2Regs<C> R, R2;
    fast::Long2Consts<C> K;
    const int steps = fast::long2_task_steps<C>(A);
    const bool edge = tj == 0 || tk == 0;
    int g = 0;  // steps run by this CTA so far: the barrier phases continue across tasks

    for (;;) {
        __syncthreads();  // everybody is done with the previous task descriptor
        if (tid == kGateTid) {
            const unsigned long long t = atomicAdd(A.counter, 1ull);
            if (t < (unsigned long long)A.tasks_total) {
                fast::sweep_describe_long<C>(A, (long long)t, S.task);
                // column pipeline: the first / last local column pair exchanges its K face with a neighbour GPU
                if (COLS && A.cface_in && S.task.bkp == 0) {
                    S.task.fk_r = A.cface_in + (long long)S.task.bj * A.fk_slot_words;
                    S.task.left_task = -2 - S.task.bj;  // <= -2: gated by the peer-written progress word of block row bj
                }
                if (COLS && A.cface_out && S.task.bkp == A.nbkp - 1) S.task.fk_w = A.cface_out + (long long)S.task.bj * A.fk_slot_words;
                // ring re-use: the previous tenants of the face slots must have been consumed (the ring sizes exceed
                // the largest producer->consumer distance in claim order by more than the grid size, so every
                // consumer waited for here was claimed long ago)
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
        fast::long2_load_bc<C>(S, A, tid);
        __syncthreads();
        fast::long2_build_tables<C>(S, tid);
        fast::long2_init_thread<C>(R, S, tj, tk);
        const fast::SweepTask &T = S.task;
        const int *top_p = T.top_task >= 0 ? A.progress + T.top_task : nullptr;
        const bool left_remote = COLS && T.left_task <= -2;
        const int *left_p = T.left_task >= 0 ? A.progress + T.left_task : (left_remote ? A.cprog_in + (-2 - T.left_task) : nullptr);
        int *const right_p = (COLS && A.cprog_out && T.bkp == A.nbkp - 1) ? A.cprog_out + T.bj : nullptr;  // peer memory
        // producers' progress: the left one may be a word another GPU writes (system scope, epoch-tagged)
        auto load_left = [&]() {
            if (COLS && left_remote) return col_decode(ld_acquire_sys(left_p), A.epoch);
            return ld_acquire(left_p);
        };
        // the step loop polls relaxed and fences once per certification it relies on (see ld_relaxed)
        auto poll_top = [&]() { return ld_relaxed(top_p); };
        auto poll_left = [&]() {
            if (COLS && left_remote) return col_decode(ld_relaxed_sys(left_p), A.epoch);
            return ld_relaxed(left_p);
        };
        auto certify_fence = [&]() {
            if (COLS && left_remote) fence_acquire_sys();
            else fence_acquire_gpu();
        };
        // publication of this task's progress: locally, and to the right neighbour GPU for the last local column
        // The remote word is released at system scope -- a MEMBAR.SYS that waits for the acknowledgements of the face
        // rows stored into the neighbour's memory over NVLink, several microseconds with the CTA held at its next
        // barrier -- so it follows a coarser rhythm (A.rgate steps) than the local one; the count the neighbour's first
        // task is waiting for (HOP_K + 1 + LAG) goes out at once.
        auto publish = [&](int done) {
            st_release(A.progress + T.self_task, done);
            if (COLS && right_p &&
                (done >= fast::kProgressDone || (done & (A.rgate - 1)) == 0 || done == C::HOP_K + 1 + LAG))
                st_release_sys(right_p, col_encode(done, A.epoch));
        };

        // prologue: the halos of steps 0 and 1 are fetched here, and the first LAG iterations fetch before any barrier
        // of this task has been waited for.  cert = last consumption step whose face rows are known to be complete.
        int cert = -1000;
        if (tid == kGateTid) {
            for (;;) {
                cert = fast::long2_certified<C>(A, T, top_p ? ld_acquire(top_p) : 0, left_p ? load_left() : 0);
                if (cert >= 1 + LAG) break;
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (edge) fast::long2_edge_fetch<C>(S, A, tj, tk, 0);
        fast::ta_copy_commit();
        fast::long2_thread_prepare<C>(R, K, S, A, tj, tk, 0, tid);

        // One step.  The gate thread certifies consumption step s + AHEAD before it arrives for step s: every edge
        // thread fetches the halos of step x + 2 only after it has seen the arrivals of step x + 2 - AHEAD.  The
        // producers' progress is sampled at the top of the step, so the load latency hides under the tile.
#define TA_LONG_STEP(RARE, RIN, ROUT, S_)                                                                                     \
    do {                                                                                                                \
        const int s_ = (S_);                                                                                            \
        const int want = s_ + C::AHEAD < steps ? s_ + C::AHEAD : steps;                                                 \
        const bool need = tid == kGateTid && cert < want;                                                               \
        int tp = 0, lp = 0;                                                                                             \
        if (need) {                                                                                                     \
            tp = top_p ? poll_top() : 0;                                                                                \
            lp = left_p ? poll_left() : 0;                                                                              \
        }                                                                                                               \
        if (LAG == 2) {                                                                                                 \
            if (g >= 2) bar_wait(&step_bar[g & 1], ((g - 2) >> 1) & 1); /* everybody has finished step s-2 */           \
            /* release: the face stores of steps <= s-2 are ordered before the arrivals just observed */                \
            if (tid == kGateTid && s_ >= 2 && fast::long2_publish_now<C>(A, s_ - 1))                                    \
                publish(s_ - 1);                                                                                        \
        }                                                                                                               \
        fast::ta_copy_wait_prev(); /* this thread's block-halo fetch for step s has landed */                           \
        fast::long2_thread_step<C, CHECK, RARE>(RIN, ROUT, K, S, A, tj, tk, s_, tid);                                  \
        if (need) {                                                                                                     \
            cert = fast::long2_certified<C>(A, T, tp, lp);                                                              \
            while (cert < want) {                                                                                       \
                __nanosleep(40);                                                                                        \
                cert = fast::long2_certified<C>(A, T, top_p ? poll_top() : 0, left_p ? poll_left() : 0);                \
            }                                                                                                           \
            certify_fence(); /* the rows certified by the values just read are ordered after them */                    \
        }                                                                                                               \
        bar_arrive_warp(&step_bar[LAG == 2 ? (g & 1) : 0]);                                                             \
        fast::long2_thread_prepare<C, RARE>(ROUT, K, S, A, tj, tk, s_ + 1, tid);                                        \
        if (LAG == 1) {                                                                                                 \
            bar_wait(&step_bar[0], g & 1); /* everybody has finished step s */                                          \
            if (tid == kGateTid && fast::long2_publish_now<C>(A, s_ + 1)) publish(s_ + 1);                              \
        }                                                                                                               \
        ++g;                                                                                                            \
    } while (0)
        // Three phases (long2_mid_*): head and tail share the full step; the middle -- nearly all steps of a long
        // sweep -- runs a compact instantiation without the rare paths (plane-0 reset, range checks, score and final
        // states, slab hand-over), which more than halves the loop body.  (K1 / K2S also alternate two register sets
        // in a loop unrolled by two; here that doubled a 30 KB loop body and ran slower.)
        {
            int s = 0;
            const int mid_begin = fast::long2_mid_begin<C>(A);
            int mid_end = fast::long2_mid_end<C>(A);
            // Experiment (option k2_pp=1, unchecked kernel only): the middle loop unrolled by two over two alternating
            // register sets, which removes the register moves of the loop-carried state (+4.5 % in K1 / K2S).  Here it
            // gains nothing at L = 8,192 (1,131 vs 1,124 GCUPS) and loses a lot whenever tasks run in lock-step with
            // their producers (1,024^3: 392 vs 492; 4,096 x 1,024^2: 536 vs 763), so it stays off.
            const bool pingpong = !CHECK && A.pingpong;
            if (pingpong && mid_begin <= mid_end && ((mid_end - mid_begin) & 1) == 0) --mid_end;  // even count
#pragma unroll 1
            for (int phase = 0; phase < 3; ++phase) {
                if (phase == 1) {
#pragma unroll 1
                    if (!CHECK && pingpong) {
#pragma unroll 1
                        for (; s <= mid_end; s += 2) {
                            TA_LONG_STEP(false, R, R2, s);
                            TA_LONG_STEP(false, R2, R, s + 1);
                        }
                    } else {
#pragma unroll 1
                        for (; s <= mid_end; ++s) TA_LONG_STEP(false, R, R, s);
                    }
                } else {
                    const int end = (phase == 0 && mid_begin <= mid_end) ? mid_begin - 1 : steps;
#pragma unroll 1
                    for (; s <= end; ++s) TA_LONG_STEP(true, R, R, s);
                }
            }
        }
#undef TA_LONG_STEP
        fast::ta_copy_wait();
        __syncthreads();  // every thread's face / hand-over stores precede the final release
        if (CHECK && R.viol) atomicOr(A.viol_flag, 1);
        if (tid == kGateTid) {
            publish(fast::kProgressDone);
            if (A.iface_out_flag) {
                __threadfence_system();
                st_release_sys(A.iface_out_flag + T.self_task, A.epoch);
            }
        }
    }
}

template <int LAG>
int run_long2(DeviceCtx *ctx, cudaStream_t st, fast::SweepArgs A, bool check, int la, bool cols = false)
{
    using C = fast::Long2Cfg<LAG>;
    static_assert(sizeof(fast::Long2Smem<C>) + 64 <= 227 * 1024, "shared memory budget");
    const size_t smem = sizeof(fast::Long2Smem<C>);
    static bool prepared[kMaxDevices][3] = {};
    const int variant = cols ? 2 : (check ? 1 : 0);
    if (!prepared[ctx->device][variant]) {
        if (cols) {
            TA_CUDA(cudaFuncSetAttribute(long2_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        } else if (check) {
            TA_CUDA(cudaFuncSetAttribute(long2_kernel<LAG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        } else {
            TA_CUDA(cudaFuncSetAttribute(long2_kernel<LAG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        prepared[ctx->device][variant] = true;
    }
    A.nbj = (A.lb + C::BJ - 1) / C::BJ;
    A.nbkp = (A.lc + C::KSPAN - 1) / C::KSPAN;
    A.tasks_total = (long long)A.nbj * A.nbkp;
    int grid = ctx->sm_count;  // one persistent CTA per SM; all co-resident (spin-waits need that)
    if (grid > A.tasks_total) grid = (int)A.tasks_total;
    A.fj_slot_words = (long long)(la + C::PLAG + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(la + C::PLAG + 2) * C::FK_ROW_WORDS;
    // claim order (earliest possible start first) and the face-ring sizes it needs
    std::vector<int> order((size_t)A.tasks_total), pos_of((size_t)A.tasks_total);
    for (int t = 0; t < (int)A.tasks_total; ++t) order[t] = t;
    const int nbj = A.nbj, gate = A.gate, nbkp = A.nbkp;
    // Column pipeline: row-major claims, so that every GPU works on the same band of block rows at the same time and the
    // last local column of a row -- what the right neighbour waits for -- follows its first column within 8 claims.
    // Earliest-start order (option k2_cols_order=diag) is 2 % faster for a range by itself (65.9 vs 67.3 ms per eighth of
    // 8192^3), but tilts the sweep: column 7 trails column 0 by 0.7 wave, the neighbour trails by as much, and 8 GPUs
    // take 96 ms instead of 78 (tools/sim_cols_schedule.py models both).
    const bool row_major = cols && !opt_is("k2_cols_order", "diag");
    std::stable_sort(order.begin(), order.end(), [nbj, gate, nbkp, row_major](int x, int y) {
        if (row_major) return (x % nbj) * nbkp + x / nbj < (y % nbj) * nbkp + y / nbj;
        return fast::long2_order_key<C>(x % nbj, x / nbj, gate) < fast::long2_order_key<C>(y % nbj, y / nbj, gate);
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
    int rc;
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
    TA_CUDA(cudaMemsetAsync(ctx->misc.ptr, 0, ctl_bytes, st));
    if (opt_debug())
        std::fprintf(stderr, "[trialign] K2L: la=%d lag=%d gate=%d blocks %dx%d(pairs) tasks=%lld grid=%d faces=%.1f MB smem=%zu base=%d\n",
                     la, LAG, A.gate, A.nbj, A.nbkp, A.tasks_total, grid, (fj_bytes + fk_bytes) / 1e6, smem, A.base);
    if (cols) long2_kernel<1, false, true><<<grid, C::NT, smem, st>>>(A);
    else if (check) long2_kernel<LAG, true><<<grid, C::NT, smem, st>>>(A);
    else long2_kernel<LAG, false><<<grid, C::NT, smem, st>>>(A);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace

// ---- limits and hand-over sizes (host) ----
// 16-bit envelope of the fast kernels: stored value = v + 2*(i + jl + kl) + base, block-local jl, kl <= 64 / 128.
// `planes` = planes of ONE launch (an i-slab of a longer triple keeps only its own a symbols in shared memory and
// restarts its plane potential); the value envelope depends on the whole triple (la_total).
bool fast_long_range_ok(int planes, int la_total, int lb, int lc)
{
    using C = fast::Long2Cfg<1>;
    if (planes < 1 || la_total < planes || lb < 1 || lc < 1) return false;
    const long long m = la_total < lb ? (la_total < lc ? la_total : lc) : (lb < lc ? lb : lc);
    const long long base = -envelope_lo(m) + 32;
    const long long top = base + envelope_hi(m) + 9 + 2 * ((long long)planes + 2 * C::PLAG + 2 * C::DEPTH + 64 + 128) + 600;
    return top < 65535;
}

bool fast_long_supports(int la, int lb, int lc)
{
    using C = fast::Long2Cfg<1>;
    if (la < 1 || lb < 1 || lc < 1 || la > C::LA_MAX - 64) return false;
    return fast_long_range_ok(la, la, lb, lc);
}

// i-slab hand-over: per block task the plane state {all, wX, u1, u3} of 16 cells x 256 threads (packed words) + one flag
size_t fast_long_slab_state_bytes(int lb, int lc)
{
    using C = fast::Long2Cfg<1>;
    const long long tasks = (long long)((lb + C::BJ - 1) / C::BJ) * ((lc + C::KSPAN - 1) / C::KSPAN);
    return (size_t)tasks * 4 * 16 * C::NT * sizeof(uint32_t);
}
size_t fast_long_slab_flag_bytes(int lb, int lc)
{
    using C = fast::Long2Cfg<1>;
    const long long tasks = (long long)((lb + C::BJ - 1) / C::BJ) * ((lc + C::KSPAN - 1) / C::KSPAN);
    return (size_t)tasks * sizeof(int);
}

size_t fast_long2_col_face_bytes(int la, int lb)
{
    using C = fast::Long2Cfg<1>;
    const size_t nbj = (size_t)(lb + C::BJ - 1) / C::BJ;
    return nbj * (size_t)(la + C::PLAG + 2) * C::FK_ROW_WORDS * sizeof(uint32_t);
}
size_t fast_long2_col_prog_bytes(int lb)
{
    using C = fast::Long2Cfg<1>;
    return (size_t)((lb + C::BJ - 1) / C::BJ) * sizeof(int);
}

int fast_long2_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb,
                   const uint8_t *d_c, int lc, int32_t *d_score, int32_t *d_states, int check_bits, int *d_viol,
                   const SlabLink *slab, const ColLink *col)
{
    const bool check = check_bits > 0;
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
    // LAG 1 (split-phase barrier, hop 17 / 33 steps between dependent tasks) is the default: measured faster than the
    // lagged barrier (LAG 2) at L = 1,024 (416 vs 325 GCUPS, the sweep is dominated by the block-to-block pipeline
    // fill) and at L = 8,192 (943 vs 860).  Long sweeps release their progress less often (a release is a MEMBAR.GPU).
    // Options k2_lag / k2_gate (trialign_set_option) override for experiments.
    int lag = 1;
    A.gate = la >= 4096 ? 32 : 4;  // measured at L = 8,192: gate 8 1,101, 16 1,115, 32 1,126 GCUPS; at L = 1,024: 2 435, 4 461, 8 457
    {
        const int v = opt_int("k2_lag", 0);
        if (v == 1 || v == 2) lag = v;
    }
    A.gate_tid = opt_int("k2_gatetid", 128) & 255;
    A.pingpong = opt_int("k2_pp", 0) ? 1 : 0;
    A.early = opt_int("k2_early", 0) ? 1 : 0;
    {
        const int g = opt_int("k2_gate", 0);
        if (g >= 1 && g <= 64 && (g & (g - 1)) == 0) A.gate = g;
    }
    A.rgate = 128;  // column pipeline: steps between two system-scope releases of a boundary task's progress
    {
        const int g = opt_int("k2_rgate", 0);
        if (g >= A.gate && g <= 4096 && (g & (g - 1)) == 0) A.rgate = g;
    }
    if (A.rgate < A.gate) A.rgate = A.gate;
    A.out = d_score;
    A.out_states = d_states;
    const long long m0 = la < lb ? (la < lc ? la : lc) : (lb < lc ? lb : lc);
    A.base = (int)(-envelope_lo(m0) + 32);
    A.last_slab = 1;
    if (slab) {  // planes i_begin+1 .. i_end of a triple of slab->la_total planes
        const long long lt = slab->la_total;
        const long long m = lt < lb ? (lt < lc ? lt : lc) : (lb < lc ? lb : lc);
        A.base = (int)(-envelope_lo(m) + 32);
        A.iface_in = (const uint32_t *)slab->in_state;
        A.iface_in_flag = (const int *)slab->in_flags;
        A.iface_out = (uint32_t *)slab->out_state;
        A.iface_out_flag = (int *)slab->out_flags;
        A.last_slab = slab->last ? 1 : 0;
        A.epoch = slab->epoch;
    }
    if (col) {  // block-column pairs [first, first + lc/128) of a triple of col->lc_total columns (c points at them)
        const long long lt = col->lc_total;
        const long long m = la < lb ? (la < lt ? la : lt) : (lb < lt ? lb : lt);
        A.base = (int)(-envelope_lo(m) + 32);
        A.cface_in = (const uint32_t *)col->in_face;
        A.cprog_in = (const int *)col->in_prog;
        A.cface_out = (uint32_t *)col->out_face;
        A.cprog_out = (int *)col->out_prog;
        A.last_slab = col->last ? 1 : 0;
        A.epoch = col->epoch;
        if (check) return TRIALIGN_EINEXACT;  // the column pipeline needs an input that is provably wrap-free
        return run_long2<1>(ctx, st, A, false, la, true);  // LAG 1: the boundary faces are laid out for PLAG = 16
    }
    return lag == 2 ? run_long2<2>(ctx, st, A, check, la) : run_long2<1>(ctx, st, A, check, la);
}

}  // namespace trialign
