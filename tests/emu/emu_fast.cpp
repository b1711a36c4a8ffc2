// emu_fast.cpp -- host emulation of the fast kernels' CTA choreography (test infrastructure).
// The per-thread step functions in csrc/fast_*.cuh are __host__ __device__; here every "thread" of a CTA is run
// sequentially inside each barrier-delimited step, which is exactly the kernel's schedule (threads only read what
// was published before the previous barrier).  Lets the CPU test-suite check the skewing, halo exchange, pair
// hand-over and score extraction against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../hw-accelerator-three-sequence-alignment_b200/csrc/fast_batch.cuh"

using namespace trialign::fast;

// SC = ScoreRtl: the unchecked kernel.  SC = ScoreGen<..>: run-time scoring parameters on the range-checked kernel
// (PAD sequences, per-pair flags), as fast_batch_kernel_scored runs it.
template <class C, class SC = ScoreRtl>
static int emu_fast_batch_cfg(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                              int pairs_per_cta, int schedule_seed, int32_t *scores, const SC &sc = SC(),
                              const BatchCheck *chk = nullptr)
{
    constexpr bool CHECK = !SC::kRtl;
    std::vector<uint32_t> viol(C::NT, 0);
    if (len < 1 || len > 64 || n < 0) return -1;
    BatchArgs A;
    A.a = a; A.b = b; A.c = c; A.n = n; A.len = len; A.scores = scores;
    A.pairs_total = (n + 1) / 2;
    A.pairs_per_cta = pairs_per_cta;
    A.base = base;
    auto S = std::make_unique<BatchSmem<C>>();
    std::vector<BatchRegs<C>> R(C::NT);
    for (long long pair_base = 0; pair_base < A.pairs_total; pair_base += pairs_per_cta) {
        const int P = (int)std::min<long long>(pairs_per_cta, A.pairs_total - pair_base);
        std::memset(S.get(), 0xA5, sizeof(BatchSmem<C>));  // poison: nothing may be read before it is written
        std::memset(R.data(), 0x5A, R.size() * sizeof(BatchRegs<C>));
        for (int fill = 1; fill < C::NBUF; ++fill) {  // as the kernel prologue does
            std::memset(S->acode[fill], 0, sizeof S->acode[fill]);
            std::memset(S->seq[fill], 0, sizeof S->seq[fill]);
            std::memset(S->bpk[fill], 0, sizeof S->bpk[fill]);
            std::memset(S->cpk[fill], 0, sizeof S->cpk[fill]);
        }
        for (int tid = 0; tid < C::NT; ++tid) load_pair_sequences<C, CHECK>(*S, A, tid, pair_base, 0);
        for (int tid = 0; tid < C::NT; ++tid) build_pair_tables<C, SC>(*S, tid, 0, sc);
        if constexpr (C::LAG == 2 && C::DIAG)
            for (int tid = 0; tid < C::NT; ++tid) {
                batch_write_borders_all<C, SC>(*S, A, tid, 0, sc);
                batch_write_borders_all<C, SC>(*S, A, tid, 1, sc);
            }
        const int steps = batch_total_steps<C>(P);
        std::vector<BatchConsts<C>> K(C::NT);
        for (int tid = 0; tid < C::NT; ++tid) {
            int tj, tk;
            thread_pos<C>(tid, tj, tk);
            batch_thread_prepare<C, SC>(R[tid], K[tid], *S, A, tj, tk, 0, sc);
        }
        if (C::LAG == 1) {
            // the kernel's schedule: step s | arrive | uniform actions, prepare s+1 | wait.  Two passes per step: nothing
            // written after the arrive (pair staging, registers) may be needed by another thread's step s.
            for (int s = 0; s <= steps; ++s) {
                for (int tid = 0; tid < C::NT; ++tid) {
                    int tj, tk;
                    thread_pos<C>(tid, tj, tk);
                    batch_thread_step<C, CHECK, SC>(R[tid], R[tid], K[tid], *S, A, tj, tk, s, P, pair_base, &viol[tid], chk, sc);
                }
                for (int tid = C::NT - 1; tid >= 0; --tid) {
                    int tj, tk;
                    thread_pos<C>(tid, tj, tk);
                    batch_uniform_actions<C, CHECK, SC>(*S, A, tid, s, P, pair_base, sc);
                    batch_thread_prepare<C, SC>(R[tid], K[tid], *S, A, tj, tk, s + 1, sc);
                }
            }
        } else {
            // lagged barrier: a WARP may run step s as soon as every warp has finished step s-2, so warps drift up to
            // one step apart.  Warps are advanced in a pseudo-random order (schedule_seed), one whole step per turn
            // (step | arrive | uniform actions, prepare s+1), lanes in order -- lanes of a warp never read each other's
            // edges of the same step.
            constexpr int NW = C::NT / 32;
            std::vector<int> done(NW, -1);  // last finished step per warp
            uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)schedule_seed;
            int finished = 0;
            while (finished < NW) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                int w = (int)((rng >> 33) % NW);
                const int burst = schedule_seed ? 1 + (int)((rng >> 20) % 3) : 1;
                for (int b = 0; b < burst; ++b) {
                    const int s = done[w] + 1;
                    if (s > steps) break;
                    const int mn = *std::min_element(done.begin(), done.end());
                    if (mn < s - 2) break;  // bar_wait(step s-2) would block
                    if constexpr (C::LAG == 2 && C::DIAG)
                        for (int lane = 0; lane < 32; ++lane) batch_write_borders_all<C, SC>(*S, A, w * 32 + lane, s + 2, sc);
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = w * 32 + lane;
                        int tj, tk;
                        thread_pos<C>(tid, tj, tk);
                        batch_thread_step<C, CHECK, SC>(R[tid], R[tid], K[tid], *S, A, tj, tk, s, P, pair_base, &viol[tid], chk, sc);
                    }
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = w * 32 + lane;
                        int tj, tk;
                        thread_pos<C>(tid, tj, tk);
                        batch_uniform_actions<C, CHECK, SC>(*S, A, tid, s, P, pair_base, sc);
                        batch_thread_prepare<C, SC>(R[tid], K[tid], *S, A, tj, tk, s + 1, sc);
                    }
                    done[w] = s;
                    if (s == steps) ++finished;
                }
            }
        }
    }
    return 0;
}

// K1 "P2": two planes per step on the lagged-barrier layout (fast_batch2_kernel).  Schedule of the kernel: step s |
// arrive | uniform actions, prepare s+1 | wait for everybody's step s.  Threads run their step in forward and reverse
// order on alternate steps: nothing a step reads may have been written in the same step.
extern "C" int emu_fast_batch2(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                               int pairs_per_cta, int32_t *scores)
{
    using C = BatchCfg<4, 4, 16, 16, 2>;
    if (len < 1 || len > 64 || n < 0) return -1;
    BatchArgs A;
    A.a = a; A.b = b; A.c = c; A.n = n; A.len = len; A.scores = scores;
    A.pairs_total = (n + 1) / 2;
    A.pairs_per_cta = pairs_per_cta;
    A.base = base;
    auto S = std::make_unique<BatchSmem<C>>();
    std::vector<BatchRegs<C>> R(C::NT), R2(C::NT);
    std::vector<BatchConsts<C>> K0(C::NT), K1(C::NT);
    std::vector<uint32_t> codes(C::NT);
    for (long long pair_base = 0; pair_base < A.pairs_total; pair_base += pairs_per_cta) {
        const int P = (int)std::min<long long>(pairs_per_cta, A.pairs_total - pair_base);
        std::memset(S.get(), 0xA5, sizeof(BatchSmem<C>));
        std::memset(R.data(), 0x5A, R.size() * sizeof(BatchRegs<C>));
        std::memset(R2.data(), 0x5A, R2.size() * sizeof(BatchRegs<C>));
        for (int fill = 1; fill < C::NBUF; ++fill) {
            std::memset(S->acode[fill], 0, sizeof S->acode[fill]);
            std::memset(S->seq[fill], 0, sizeof S->seq[fill]);
            std::memset(S->bpk[fill], 0, sizeof S->bpk[fill]);
            std::memset(S->cpk[fill], 0, sizeof S->cpk[fill]);
        }
        for (int tid = 0; tid < C::NT; ++tid) load_pair_sequences<C>(*S, A, tid, pair_base, 0);
        for (int tid = 0; tid < C::NT; ++tid) build_pair_tables<C>(*S, tid, 0);
        for (int tid = 0; tid < C::NT; ++tid) batch2_write_borders<C>(*S, A, tid, 0);
        for (int tid = 0; tid < C::NT; ++tid) {
            int tj, tk;
            thread_pos<C>(tid, tj, tk);
            codes[tid] = batch2_fetch_codes<C>(*S, tj, tk, 0);
            batch2_thread_prepare<C>(R[tid], K0[tid], K1[tid], *S, A, tj, tk, 0, codes[tid]);
        }
        const int steps = batch2_total_steps<C>(P);
        for (int s = 0; s <= steps; ++s) {
            for (int x = 0; x < C::NT; ++x) {
                const int tid = (s & 1) ? C::NT - 1 - x : x;
                int tj, tk;
                thread_pos<C>(tid, tj, tk);
                batch2_write_borders<C>(*S, A, tid, s + 1);  // next step's border halos: not read by anybody's step s
                batch2_thread_step<C>(R[tid], R2[tid], K0[tid], K1[tid], *S, A, tj, tk, s, P, pair_base);
            }
            for (int x = 0; x < C::NT; ++x) {
                const int tid = (s & 1) ? x : C::NT - 1 - x;
                int tj, tk;
                thread_pos<C>(tid, tj, tk);
                batch2_uniform_actions<C>(*S, A, tid, s, P, pair_base);
                batch2_thread_prepare<C>(R[tid], K0[tid], K1[tid], *S, A, tj, tk, s + 1, codes[tid]);
            }
        }
    }
    return 0;
}

// Run-time scoring parameters (fast_batch_kernel_scored): lagged-barrier 256-thread kernel, range-checked at check_bits;
// flags[(n+1)/2] gets bit 0 / 1 for the low / high triple of a pair the check did not clear.
extern "C" int emu_fast_batch_ex(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int pairs_per_cta,
                                 int schedule_seed, int M, int X, int GO, int GE, int sop, int check_bits, int32_t *scores,
                                 int *flags)
{
    using C = BatchCfg<4, 4, 16, 16, 2>;
    const BatchCheck chk = score_check(M, X, GO, check_bits, flags);
    const int base = score_base(X, GO, GE);
    if (sop) return emu_fast_batch_cfg<C, ScoreGen<true>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, make_score_gen<true>(M, X, GO, GE), &chk);
    return emu_fast_batch_cfg<C, ScoreGen<false>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, make_score_gen<false>(M, X, GO, GE), &chk);
}

// K1 "P2" with neighbour-only synchronisation (fast_batch2r_kernel).  A warp's step has two parts: A = border halos of
// step s+1, step s, ARRIVE; B = uniform actions, prepare s+1.  Warp w may start part A of step s+1 once its own part B of
// step s is done and warps w-1 and w+1 have ARRIVED for step s -- their part B may still be pending, exactly as on the
// GPU, where the staging stores after the arrive are only released by the next one.  Warps advance in a pseudo-random
// order; poisoned shared memory.
extern "C" int emu_fast_batch2_ring(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                                    int pairs_per_cta, int schedule_seed, int32_t *scores)
{
    using C = BatchCfg<4, 4, 16, 16, 2>;
    constexpr int NW = C::NT / 32;
    if (len < 1 || len > 64 || n < 0) return -1;
    BatchArgs A;
    A.a = a; A.b = b; A.c = c; A.n = n; A.len = len; A.scores = scores;
    A.pairs_total = (n + 1) / 2;
    A.pairs_per_cta = pairs_per_cta;
    A.base = base;
    auto S = std::make_unique<BatchSmem<C>>();
    std::vector<BatchRegs<C>> R(C::NT), R2(C::NT);
    std::vector<BatchConsts<C>> K0(C::NT), K1(C::NT);
    std::vector<uint32_t> codes(C::NT);
    for (long long pair_base = 0; pair_base < A.pairs_total; pair_base += pairs_per_cta) {
        const int P = (int)std::min<long long>(pairs_per_cta, A.pairs_total - pair_base);
        std::memset(S.get(), 0xA5, sizeof(BatchSmem<C>));
        std::memset(R.data(), 0x5A, R.size() * sizeof(BatchRegs<C>));
        std::memset(R2.data(), 0x5A, R2.size() * sizeof(BatchRegs<C>));
        for (int fill = 1; fill < C::NBUF; ++fill) {
            std::memset(S->acode[fill], 0, sizeof S->acode[fill]);
            std::memset(S->seq[fill], 0, sizeof S->seq[fill]);
            std::memset(S->bpk[fill], 0, sizeof S->bpk[fill]);
            std::memset(S->cpk[fill], 0, sizeof S->cpk[fill]);
        }
        for (int tid = 0; tid < C::NT; ++tid) load_pair_sequences<C>(*S, A, tid, pair_base, 0);
        for (int tid = 0; tid < C::NT; ++tid) build_pair_tables<C>(*S, tid, 0);
        for (int tid = 0; tid < C::NT; ++tid) batch2_write_borders_warp<C>(*S, A, tid, 0);
        for (int tid = 0; tid < C::NT; ++tid) {
            int tj, tk;
            thread_pos<C>(tid, tj, tk);
            codes[tid] = batch2_fetch_codes<C>(*S, tj, tk, 0);
            batch2_thread_prepare<C>(R[tid], K0[tid], K1[tid], *S, A, tj, tk, 0, codes[tid]);
        }
        const int steps = batch2_total_steps<C>(P);
        std::vector<int> arrived(NW, -1), prepared(NW, -1);  // last step with part A / part B done
        uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)schedule_seed ^ ((uint64_t)pair_base << 17);
        int finished = 0;
        while (finished < NW) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const int w = (int)((rng >> 33) % NW);
            if (prepared[w] < arrived[w]) {  // part B of step arrived[w]
                const int s = arrived[w];
                for (int lane = 0; lane < 32; ++lane) {
                    const int tid = w * 32 + lane;
                    int tj, tk;
                    thread_pos<C>(tid, tj, tk);
                    batch2_uniform_actions<C, true>(*S, A, tid, s, P, pair_base);
                    batch2_thread_prepare<C>(R[tid], K0[tid], K1[tid], *S, A, tj, tk, s + 1, codes[tid]);
                }
                prepared[w] = s;
                if (s == steps) ++finished;
                continue;
            }
            const int s = arrived[w] + 1;
            if (s > steps) continue;
            if (arrived[(w + NW - 1) % NW] < s - 1 || arrived[(w + 1) % NW] < s - 1) continue;  // the two waits would block
            for (int lane = 0; lane < 32; ++lane) batch2_write_borders_warp<C>(*S, A, w * 32 + lane, s + 1);
            for (int lane = 0; lane < 32; ++lane) {
                const int tid = w * 32 + lane;
                int tj, tk;
                thread_pos<C>(tid, tj, tk);
                batch2_thread_step<C>(R[tid], R2[tid], K0[tid], K1[tid], *S, A, tj, tk, s, P, pair_base);
            }
            arrived[w] = s;
        }
    }
    return 0;
}

// threads = 256: 16x16 threads x 4x4 cells; threads = 512: 16x32 threads x 4x2 cells; lag = 1 / 2 (see BatchCfg);
// schedule_seed != 0 (lag 2 only): warps advance in a pseudo-random order with bursts
extern "C" int emu_fast_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                              int pairs_per_cta, int threads, int lag, int schedule_seed, int32_t *scores)
{
    (void)schedule_seed;
    if (threads == 256 && lag == 1) return emu_fast_batch_cfg<BatchCfg<4, 4, 16, 16>>(a, b, c, n, len, base, pairs_per_cta, 0, scores);
    if (threads == 512 && lag == 1) return emu_fast_batch_cfg<BatchCfg<4, 2, 16, 32>>(a, b, c, n, len, base, pairs_per_cta, 0, scores);
    if (threads == 256 && lag == 2) return emu_fast_batch_cfg<BatchCfg<4, 4, 16, 16, 2>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores);
    return -1;
}

// ---------------------------------------------------------------------------------------------------------------
// K2S (fast_stream.cuh): all blocks of a CTA's pairs streamed back to back, lagged barrier, faces in "global" memory.
// Warps advance in a pseudo-random order (schedule_seed != 0: with bursts), at most two steps apart, exactly what the
// kernel's mbarrier protocol allows.  Every emulated CTA gets poisoned shared memory and poisoned face rings.
// ---------------------------------------------------------------------------------------------------------------
#include "../../hw-accelerator-three-sequence-alignment_b200/csrc/fast_stream.cuh"

template <class SC, class C = StreamCfg>
static int emu_fast_stream_sc(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                              int pairs_per_cta, int schedule_seed, int32_t *scores, const SC &sc, const RangeCheck *chk)
{
    constexpr bool CHECK = !SC::kRtl;
    std::vector<uint32_t> viol(C::NT, 0);
    if (len < C::LEN_MIN || len > C::LEN_MAX || n < 0 || pairs_per_cta < 1) return -1;
    StreamArgs A{};
    A.a = a; A.b = b; A.c = c; A.n = n; A.len = len; A.scores = scores;
    A.nbj = (len + C::BJ - 1) / C::BJ;
    A.nbk = (len + C::BK - 1) / C::BK;
    A.pairs_total = (n + 1) / 2;
    A.pairs_per_cta = pairs_per_cta;
    A.fj_slot_words = (long long)(len + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(len + 2) * C::FK_ROW_WORDS;
    A.fk_ring = A.nbj + 1;
    A.base = base;
    const int n_ctas = (int)((A.pairs_total + pairs_per_cta - 1) / pairs_per_cta);
    std::vector<uint32_t> fj((size_t)n_ctas * 2 * A.fj_slot_words, 0xDEADBEEFu), fk((size_t)n_ctas * A.fk_ring * A.fk_slot_words, 0xDEADBEEFu);
    A.face_j = fj.data(); A.face_k = fk.data();
    auto S = std::make_unique<StreamSmem<C>>();
    std::vector<StreamRegs<C>> R(C::NT);
    std::vector<StreamConsts<C>> K(C::NT);
    std::vector<StreamPos> pos(C::NT), duty(C::NT);
    for (int cta = 0; cta < n_ctas; ++cta) {
        const long long pair_base = (long long)cta * pairs_per_cta;
        const int P = (int)std::min<long long>(pairs_per_cta, A.pairs_total - pair_base);
        const int tasks = P * A.nbj * A.nbk;
        std::memset(S.get(), 0xA5, sizeof(StreamSmem<C>));
        std::memset(R.data(), 0x5A, R.size() * sizeof(StreamRegs<C>));
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_load<C>(*S, A, tid, cta, P, pair_base, -1);
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_load<C>(*S, A, tid, cta, P, pair_base, 0);
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_build<C, SC>(*S, tid, -1, sc);
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_build<C, SC>(*S, tid, 0, sc);
        for (int tid = 0; tid < C::NT; ++tid) {
            int tj, tk;
            stream_thread_pos<C>(tid, tj, tk);
            pos[tid] = stream_start_pos<C>(A, tj, tk);
            duty[tid] = stream_pos_at<C>(A, C::LAG * (tid < 128 ? (tid & 15) : C::TJ - 1 + (tid & 15)), 0);
            if (C::DIST) {
                if (tid < 128) {
                    stream_duty_fetch<C, SC>(*S, A, tid, duty[tid], 0, sc);
                    stream_advance(A, duty[tid], C::NBUF);
                    stream_duty_fetch<C, SC>(*S, A, tid, duty[tid], 1, sc);
                    stream_advance(A, duty[tid], C::NBUF);
                }
            } else if (tj == 0 || tk == 0) stream_edge_fetch<C, SC>(*S, A, tj, tk, pos[tid], C::LAG & (C::NSLOT - 1), sc);
            stream_thread_prepare<C, SC>(R[tid], K[tid], *S, A, tj, tk, pos[tid], 0, false, sc);
        }
        const int steps = (int)stream_total_steps<C>(A, P);
        constexpr int NW = C::NT / 32;
        std::vector<int> done(NW, -1);  // last finished step per warp
        uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)schedule_seed ^ ((uint64_t)cta << 20);
        int finished = 0;
        while (finished < NW) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const int w = (int)((rng >> 33) % NW);
            const int burst = schedule_seed ? 1 + (int)((rng >> 20) % 3) : 1;
            for (int bi = 0; bi < burst; ++bi) {
                const int s = done[w] + 1;
                if (s > steps) break;
                if (*std::min_element(done.begin(), done.end()) < s - 2) break;  // bar_wait(step s-2) would block
                const int g_t = s / len, g_off = s % len;
                if (C::DIST)  // the duties at the top of the step: halos of step s+2 in, faces of step s-2 out
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = w * 32 + lane;
                        if (tid < 128) {
                            stream_duty_fetch<C, SC>(*S, A, tid, duty[tid], s + 2, sc);
                            stream_advance(A, duty[tid], C::NBUF);
                        } else if (s >= 2) {
                            stream_duty_store<C, SC>(*S, A, tid, duty[tid], s - 2, sc);
                            stream_advance(A, duty[tid], C::NBUF);
                        }
                    }
                for (int lane = 0; lane < 32; ++lane) {
                    const int tid = w * 32 + lane;
                    int tj, tk;
                    stream_thread_pos<C>(tid, tj, tk);
                    stream_thread_step<C, CHECK, SC>(R[tid], R[tid], K[tid], *S, A, tj, tk, s, pos[tid], &viol[tid], chk, sc);
                }
                for (int lane = 0; lane < 32; ++lane) {
                    const int tid = w * 32 + lane;
                    int tj, tk;
                    stream_thread_pos<C>(tid, tj, tk);
                    if (g_t < tasks) stream_uniform_actions<C, SC>(*S, A, tid, cta, P, pair_base, g_t, g_off, sc);
                    stream_thread_prepare<C, SC>(R[tid], K[tid], *S, A, tj, tk, pos[tid], s + 1, true, sc);
                }
                done[w] = s;
                if (s == steps) ++finished;
            }
        }
    }
    return 0;
}

extern "C" int emu_fast_stream(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                               int pairs_per_cta, int schedule_seed, int32_t *scores)
{
    return emu_fast_stream_sc<ScoreRtl>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, ScoreRtl(), nullptr);
}

// the block-edge duties dealt to all threads (fast_stream_kernel_dist)
extern "C" int emu_fast_stream_dist(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                                    int pairs_per_cta, int schedule_seed, int32_t *scores)
{
    return emu_fast_stream_sc<ScoreRtl, StreamCfgD>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, ScoreRtl(), nullptr);
}

// run-time scoring parameters on the range-checked kernel (fast_stream_kernel_scored)
extern "C" int emu_fast_stream_ex(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int pairs_per_cta,
                                  int schedule_seed, int M, int X, int GO, int GE, int sop, int check_bits, int32_t *scores,
                                  int *flags)
{
    const BatchCheck bc = score_check(M, X, GO, check_bits, flags);
    RangeCheck chk;
    chk.check_hi = bc.check_hi; chk.check_lo = bc.check_lo; chk.pair_flags = flags;
    const int base = score_base_len(X, GO, GE, len);
    if (sop) return emu_fast_stream_sc<ScoreGen<true>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, make_score_gen<true>(M, X, GO, GE), &chk);
    return emu_fast_stream_sc<ScoreGen<false>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, make_score_gen<false>(M, X, GO, GE), &chk);
}

// ---------------------------------------------------------------------------------------------------------------
// K2S "P2" (fast_stream2.cuh): two planes per step, block-edge duties by thread id.  Schedule of the kernel: duties
// (halos of step s+1 in, faces of step s-1 out) | step s | arrive | uniform actions, prepare s+1 | wait for everybody's
// step s.  Threads run forward and in reverse on alternate steps (schedule_seed: a rotated start), so nothing a step
// reads may have been written in the same step.  Poisoned shared memory and face rings.
// ---------------------------------------------------------------------------------------------------------------
#include "../../hw-accelerator-three-sequence-alignment_b200/csrc/fast_stream2.cuh"

template <bool CHECK, class SC>
static int emu_fast_stream2_sc(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                               int pairs_per_cta, int schedule_seed, int32_t *scores, const SC &sc, const RangeCheck *chk)
{
    using C = Stream2Cfg;
    std::vector<uint32_t> viol(C::NT, 0);
    if (len < C::LEN_MIN || len > C::LEN_MAX || n < 0 || pairs_per_cta < 1) return -1;
    StreamArgs A{};
    A.a = a; A.b = b; A.c = c; A.n = n; A.len = len; A.scores = scores;
    A.nbj = (len + C::BJ - 1) / C::BJ;
    A.nbk = (len + C::BK - 1) / C::BK;
    A.pairs_total = (n + 1) / 2;
    A.pairs_per_cta = pairs_per_cta;
    A.fj_slot_words = (long long)(len + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(len + 2) * C::FK_ROW_WORDS;
    A.fk_ring = A.nbj + 1;
    A.base = base;
    const int n_ctas = (int)((A.pairs_total + pairs_per_cta - 1) / pairs_per_cta), L2 = stream2_steps_per_task(A);
    std::vector<uint32_t> fj((size_t)n_ctas * 2 * A.fj_slot_words, 0xDEADBEEFu), fk((size_t)n_ctas * A.fk_ring * A.fk_slot_words, 0xDEADBEEFu);
    A.face_j = fj.data(); A.face_k = fk.data();
    auto S = std::make_unique<StreamSmem<C>>();
    std::vector<StreamRegs<C>> R(C::NT), R2(C::NT);
    std::vector<StreamConsts<C>> K0(C::NT), K1(C::NT);
    std::vector<Stream2Pos> pos(C::NT), pin(C::NT), pout(C::NT);
    for (int cta = 0; cta < n_ctas; ++cta) {
        const long long pair_base = (long long)cta * pairs_per_cta;
        const int P = (int)std::min<long long>(pairs_per_cta, A.pairs_total - pair_base);
        const int tasks = P * A.nbj * A.nbk;
        std::memset(S.get(), 0xA5, sizeof(StreamSmem<C>));
        std::memset(R.data(), 0x5A, R.size() * sizeof(StreamRegs<C>));
        std::memset(R2.data(), 0x5A, R2.size() * sizeof(StreamRegs<C>));
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_load<C>(*S, A, tid, cta, P, pair_base, -1);
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_load<C>(*S, A, tid, cta, P, pair_base, 0);
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_build<C, SC>(*S, tid, -1, sc);
        for (int tid = 0; tid < C::NT; ++tid) stream_stage_build<C, SC>(*S, tid, 0, sc);
        for (int tid = 0; tid < C::NT; ++tid) {
            int tj, tk;
            stream_thread_pos<C>(tid, tj, tk);
            pos[tid] = stream2_pos_at<C>(A, tj + tk, 0);
            pin[tid] = stream2_pos_at<C>(A, tid & 15, 0);
            pout[tid] = stream2_pos_at<C>(A, C::TJ - 1 + (tid & 15), 0);
            stream2_duty_in<C, SC>(*S, A, tid, pin[tid], 0, sc);
            stream2_advance<C>(A, pin[tid]);
        }
        for (int tid = 0; tid < C::NT; ++tid) {
            int tj, tk;
            stream_thread_pos<C>(tid, tj, tk);
            stream2_thread_prepare<C, SC>(R[tid], K0[tid], K1[tid], *S, A, tj, tk, pos[tid], sc);
        }
        const int steps = (int)stream2_total_steps<C>(A, P);
        int g_t = 0, g_off = 0;
        for (int s = 0; s <= steps; ++s) {
            const int rot = schedule_seed ? (int)(((unsigned)s * 2654435761u + (unsigned)schedule_seed * 40503u) >> 8) % C::NT : 0;
            for (int x = 0; x < C::NT; ++x) {
                const int tid = (((s & 1) ? C::NT - 1 - x : x) + rot) % C::NT;
                int tj, tk;
                stream_thread_pos<C>(tid, tj, tk);
                stream2_duty_in<C, SC>(*S, A, tid, pin[tid], s + 1, sc);
                stream2_advance<C>(A, pin[tid]);
                if (s >= 1) {
                    stream2_duty_out<C, SC>(*S, tid, pout[tid], s - 1, sc);
                    stream2_advance<C>(A, pout[tid]);
                }
                stream2_thread_step<C, CHECK, SC>(R[tid], R2[tid], K0[tid], K1[tid], *S, A, tj, tk, s, pos[tid], &viol[tid], chk, sc);
            }
            for (int x = 0; x < C::NT; ++x) {
                const int tid = (((s & 1) ? x : C::NT - 1 - x) + rot) % C::NT;
                int tj, tk;
                stream_thread_pos<C>(tid, tj, tk);
                if (g_t < tasks) stream2_uniform_actions<C, SC>(*S, A, tid, cta, P, pair_base, g_t, g_off, sc);
                stream2_advance<C>(A, pos[tid]);
                stream2_thread_prepare<C, SC>(R[tid], K0[tid], K1[tid], *S, A, tj, tk, pos[tid], sc);
            }
            if (++g_off == L2) {
                g_off = 0;
                ++g_t;
            }
        }
    }
    return 0;
}

extern "C" int emu_fast_stream2(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                                int pairs_per_cta, int schedule_seed, int32_t *scores)
{
    return emu_fast_stream2_sc<false, ScoreRtl>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, ScoreRtl(), nullptr);
}

// sop < 0: the RTL's constants, range-checked at check_bits (every 8th plane); sop = 0 / 1: run-time scoring parameters
extern "C" int emu_fast_stream2_ex(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int pairs_per_cta,
                                   int schedule_seed, int M, int X, int GO, int GE, int sop, int check_bits, int32_t *scores,
                                   int *flags)
{
    RangeCheck chk;
    chk.pair_flags = flags;
    if (sop < 0) {
        const int lim = 1 << (check_bits - 1), gap = 7;
        chk.check_hi = lim - 1 - 3 - 3 * gap;
        chk.check_lo = -lim + 8 + 4 * gap;
        return emu_fast_stream2_sc<true, ScoreRtl>(a, b, c, n, len, 2 * (len + 1) + 8 + 32, pairs_per_cta, schedule_seed, scores, ScoreRtl(), &chk);
    }
    const BatchCheck bc = score_check(M, X, GO, check_bits, flags);
    chk.check_hi = bc.check_hi; chk.check_lo = bc.check_lo;
    const int base = score_base_len(X, GO, GE, len + 1);
    if (sop) return emu_fast_stream2_sc<true, ScoreGen<true>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, make_score_gen<true>(M, X, GO, GE), &chk);
    return emu_fast_stream2_sc<true, ScoreGen<false>>(a, b, c, n, len, base, pairs_per_cta, schedule_seed, scores, make_score_gen<false>(M, X, GO, GE), &chk);
}

// ---------------------------------------------------------------------------------------------------------------
// K2L (fast_long.cuh): one long triple on several emulated CTAs.  Every CTA runs the kernel's own protocol: the warps of
// a CTA drift (a warp may run step s once every warp has finished step s-LAG), the gate warp cannot finish a step
// before the producers' progress certifies the halo fetches AHEAD steps later, progress is published with the
// kernel's rhythm.  CTAs and warps are advanced in a pseudo-random order.
// ---------------------------------------------------------------------------------------------------------------
#include "../../hw-accelerator-three-sequence-alignment_b200/csrc/fast_long.cuh"

namespace {

template <class C>
struct EmuCta2 {
    std::unique_ptr<Long2Smem<C>> S{new Long2Smem<C>};
    std::vector<Long2Regs<C>> R = std::vector<Long2Regs<C>>(C::NT);
    std::vector<Long2Consts<C>> K = std::vector<Long2Consts<C>>(C::NT);
    std::vector<int> done = std::vector<int>(C::NT / 32, -1);
    bool has_task = false, prologue_done = false;
};

template <int LAG>
int run_long2(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int base, int n_ctas,
              int want_states, int32_t *out8, int check_bits, int *viol_out, const uint32_t *iface_in, uint32_t *iface_out,
              int last_slab, int gate, int seed)
{
    using C = Long2Cfg<LAG>;
    constexpr int NW = C::NT / 32, kGateWarp = 128 / 32;
    if (la < 1 || lb < 1 || lc < 1 || la > C::LA_MAX || n_ctas < 1) return -1;
    SweepArgs A{};
    A.a = a; A.b = b; A.c = c; A.n = 1; A.la = la; A.lb = lb; A.lc = lc;
    A.nbj = (lb + C::BJ - 1) / C::BJ;
    A.nbkp = (lc + C::KSPAN - 1) / C::KSPAN;
    A.tasks_total = (long long)A.nbj * A.nbkp;
    A.gate = gate;
    A.early = seed & 1;  // both publication rhythms
    A.fj_slot_words = (long long)(la + C::PLAG + 2) * C::FJ_ROW_WORDS;
    A.fk_slot_words = (long long)(la + C::PLAG + 2) * C::FK_ROW_WORDS;
    A.fj_ring = (int)A.tasks_total;  // the emulation keeps every face (ring re-use is a launcher concern)
    A.fk_ring = (int)A.tasks_total;
    std::vector<uint32_t> fj((size_t)A.fj_ring * A.fj_slot_words, 0xDEADBEEFu), fk((size_t)A.fk_ring * A.fk_slot_words, 0xDEADBEEFu);
    std::vector<int> prog((size_t)A.tasks_total, 0);
    unsigned long long counter = 0;
    std::vector<int> order((size_t)A.tasks_total), pos_of((size_t)A.tasks_total);
    for (int t = 0; t < (int)A.tasks_total; ++t) order[t] = t;
    const int nbj = A.nbj;
    const int nbkp = A.nbkp;
    const bool row_major = (seed & 4) != 0;  // the column pipeline's claim order (fast_long.cu)
    std::stable_sort(order.begin(), order.end(), [nbj, gate, nbkp, row_major](int x, int y) {
        if (row_major) return (x % nbj) * nbkp + x / nbj < (y % nbj) * nbkp + y / nbj;
        return long2_order_key<C>(x % nbj, x / nbj, gate) < long2_order_key<C>(y % nbj, y / nbj, gate);
    });
    for (int p = 0; p < (int)A.tasks_total; ++p) pos_of[order[p]] = p;
    A.order = order.data(); A.pos_of = pos_of.data();
    A.face_j = fj.data(); A.face_k = fk.data(); A.progress = prog.data(); A.counter = &counter;
    A.out = out8; A.out_states = want_states ? out8 + 1 : nullptr; A.base = base;
    A.iface_in = iface_in; A.iface_out = iface_out; A.last_slab = last_slab;
    int viol = 0;
    A.viol_flag = &viol;
    const int lim = 1 << ((check_bits > 0 ? check_bits : 16) - 1);
    A.check_hi = lim - 1 - 3 - 45;
    A.check_lo = -lim + 8 + 60;
    std::vector<EmuCta2<C>> ctas(n_ctas);
    for (auto &cta : ctas) {
        std::memset(cta.S.get(), (seed & 2) ? 0xFF : 0x00, sizeof(Long2Smem<C>));  // nasty poison: a dead lane must never disturb a live one
        for (int tid = 0; tid < C::NT; ++tid) long2_load_a<C>(*cta.S, A, tid);
    }
    const int steps = long2_task_steps<C>(A);
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)seed;
    long long finished = 0, idle = 0;
    while (finished < A.tasks_total) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        auto &cta = ctas[(size_t)((rng >> 40) % (uint64_t)n_ctas)];
        bool progressed = false;
        if (!cta.has_task) {
            if ((long long)counter < A.tasks_total) {
                sweep_describe_long<C>(A, (long long)counter++, cta.S->task);
                cta.has_task = true;
                cta.prologue_done = false;
                for (int tid = 0; tid < C::NT; ++tid) long2_load_bc<C>(*cta.S, A, tid);
                for (int tid = 0; tid < C::NT; ++tid) long2_build_tables<C>(*cta.S, tid);
                for (int tid = 0; tid < C::NT; ++tid) {
                    int tj, tk;
                    stream_thread_pos<C>(tid, tj, tk);
                    long2_init_thread<C>(cta.R[tid], *cta.S, tj, tk);
                }
                std::fill(cta.done.begin(), cta.done.end(), -1);
                progressed = true;
            }
        } else {
            const SweepTask &T = cta.S->task;
            const int tp = T.top_task >= 0 ? prog[T.top_task] : 0, lp = T.left_task >= 0 ? prog[T.left_task] : 0;
            const int cert = long2_certified<C>(A, T, tp, lp);
            if (!cta.prologue_done) {
                if (cert >= 1 + LAG) {
                    for (int tid = 0; tid < C::NT; ++tid) {
                        int tj, tk;
                        stream_thread_pos<C>(tid, tj, tk);
                        if (tj == 0 || tk == 0) long2_edge_fetch<C>(*cta.S, A, tj, tk, 0);
                        long2_thread_prepare<C>(cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, 0, tid);
                    }
                    cta.prologue_done = true;
                    progressed = true;
                }
            } else {
                const int w = (int)((rng >> 20) % NW);
                const int s = cta.done[w] + 1;
                const int mn = *std::min_element(cta.done.begin(), cta.done.end());
                // LAG 2: step s needs everybody's step s-2; LAG 1: the wait at the end of step s-1 needs everybody's s-1
                bool can = s <= steps && mn >= s - LAG;
                const int want = s + C::AHEAD < steps ? s + C::AHEAD : steps;
                if (can && w == kGateWarp && cert < want) can = false;  // the gate thread spins before its arrival
                if (can) {
                    if (w == kGateWarp) {  // publication, as the kernel orders it around the barrier waits
                        if (LAG == 2 && s >= 2 && long2_publish_now<C>(A, s - 1)) prog[T.self_task] = s - 1;
                        if (LAG == 1 && s >= 1 && long2_publish_now<C>(A, s)) prog[T.self_task] = s;
                    }
                    // the kernel runs the compact instantiation (RARE = false) during the middle phase
                    const bool mid = s >= long2_mid_begin<C>(A) && s <= long2_mid_end<C>(A);
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = w * 32 + lane;
                        int tj, tk;
                        stream_thread_pos<C>(tid, tj, tk);
                        if (mid) long2_thread_step<C, true, false>(cta.R[tid], cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s, tid);
                        else long2_thread_step<C, true, true>(cta.R[tid], cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s, tid);
                    }
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = w * 32 + lane;
                        int tj, tk;
                        stream_thread_pos<C>(tid, tj, tk);
                        if (mid) long2_thread_prepare<C, false>(cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s + 1, tid);
                        else long2_thread_prepare<C, true>(cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s + 1, tid);
                    }
                    cta.done[w] = s;
                    progressed = true;
                    if (*std::min_element(cta.done.begin(), cta.done.end()) == steps) {
                        for (const auto &r : cta.R)
                            if (r.viol) viol = 1;
                        prog[T.self_task] = kProgressDone;
                        cta.has_task = false;
                        ++finished;
                    }
                }
            }
        }
        if (progressed) idle = 0;
        else if (++idle > 200000) return -2;  // deadlock
    }
    if (viol_out) *viol_out = viol;
    return 0;
}

}  // namespace

extern "C" int emu_fast_long2(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int base,
                              int n_ctas, int lag, int gate, int seed, int32_t *out8, int check_bits, int *viol_out)
{
    if (lag == 1) return run_long2<1>(a, la, b, lb, c, lc, base, n_ctas, 1, out8, check_bits, viol_out, nullptr, nullptr, 1, gate, seed);
    if (lag == 2) return run_long2<2>(a, la, b, lb, c, lc, base, n_ctas, 1, out8, check_bits, viol_out, nullptr, nullptr, 1, gate, seed);
    return -1;
}

// the i-slab pipeline, slabs run one after the other (what n_slabs GPUs do concurrently, task by task)
extern "C" int emu_fast_long2_slabs(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int base,
                                    int n_ctas, int n_slabs, int lag, int32_t *out8)
{
    using C = Long2Cfg<1>;
    const long long tasks = (long long)((lb + C::BJ - 1) / C::BJ) * ((lc + C::KSPAN - 1) / C::KSPAN);
    std::vector<uint32_t> st[2];
    st[0].assign((size_t)tasks * 4 * 16 * C::NT, 0xDEADBEEFu);
    st[1] = st[0];
    for (int g = 0; g < n_slabs; ++g) {
        const int i0 = (int)((long long)la * g / n_slabs), i1 = (int)((long long)la * (g + 1) / n_slabs);
        if (i1 <= i0) return -3;
        const uint32_t *in = g > 0 ? st[(g - 1) & 1].data() : nullptr;
        uint32_t *out = g + 1 < n_slabs ? st[g & 1].data() : nullptr;
        const int rc = lag == 2 ? run_long2<2>(a + i0, i1 - i0, b, lb, c, lc, base, n_ctas, 1, out8, 0, nullptr, in, out, g + 1 == n_slabs, 4, g + 1)
                                : run_long2<1>(a + i0, i1 - i0, b, lb, c, lc, base, n_ctas, 1, out8, 0, nullptr, in, out, g + 1 == n_slabs, 4, g + 1);
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Column pipeline (fast_long.cu, COLS kernels): several emulated GPUs run CONCURRENTLY, each sweeping a range of
// block-column pairs with its own task counter, claim order (row-major), face rings and progress array; the K faces
// and epoch-tagged progress words of the range edges live in the right neighbour's "memory".  The descriptor overrides,
// the tagged progress words and the publication mirror the kernel body (long2_kernel<1,false,true>).
// ---------------------------------------------------------------------------------------------------------------
namespace {

int emu_col_decode(int v, int epoch)
{
    if ((v >> 20) != epoch) return 0;
    const int steps = v & 0xFFFFF;
    return steps == 0xFFFFF ? kProgressDone : steps;
}
int emu_col_encode(int steps, int epoch) { return (epoch << 20) | (steps >= kProgressDone ? 0xFFFFF : steps); }

struct EmuGpu {
    using C = Long2Cfg<1>;
    SweepArgs A{};
    std::vector<uint32_t> fj, fk, cface;  // cface / cprog: boundary buffers of THIS gpu (written by the left neighbour)
    std::vector<int> prog, cprog, order, pos_of;
    unsigned long long counter = 0;
    std::vector<EmuCta2<C>> ctas;
    long long finished = 0;
};

}  // namespace

extern "C" int emu_fast_long2_cols(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int base,
                                   int n_ctas, int n_ranges, int seed, int epochs, int32_t *out8)
{
    using C = Long2Cfg<1>;
    constexpr int NW = C::NT / 32, kGateWarp = 128 / 32;
    const int pairs = (lc + 127) / 128;
    if (la < 1 || lb < 1 || lc < 1 || n_ranges < 1 || pairs < n_ranges || n_ctas < 1) return -1;
    const int nbj = (lb + C::BJ - 1) / C::BJ;
    const long long slot_k = (long long)(la + C::PLAG + 2) * C::FK_ROW_WORDS;
    std::vector<EmuGpu> gpus(n_ranges);
    for (int g = 0; g < n_ranges; ++g) {
        EmuGpu &G = gpus[g];
        const int p0 = pairs * g / n_ranges, p1 = pairs * (g + 1) / n_ranges;
        const int c0 = 128 * p0, c1 = std::min(lc, 128 * p1);
        SweepArgs &A = G.A;
        A.a = a; A.b = b; A.c = c + c0; A.n = 1; A.la = la; A.lb = lb; A.lc = c1 - c0;
        A.nbj = nbj;
        A.nbkp = (A.lc + C::KSPAN - 1) / C::KSPAN;
        A.tasks_total = (long long)A.nbj * A.nbkp;
        A.gate = 4;
        A.rgate = (seed & 4) ? 4 : 32;  // remote publication rhythm (fast_long.cu: 128 by default)
        A.fj_slot_words = (long long)(la + C::PLAG + 2) * C::FJ_ROW_WORDS;
        A.fk_slot_words = slot_k;
        A.fj_ring = A.fk_ring = (int)A.tasks_total;
        G.fj.assign((size_t)A.fj_ring * A.fj_slot_words, 0xDEADBEEFu);
        G.fk.assign((size_t)A.fk_ring * A.fk_slot_words, 0xDEADBEEFu);
        G.prog.assign((size_t)A.tasks_total, 0);
        G.order.resize((size_t)A.tasks_total);
        G.pos_of.resize((size_t)A.tasks_total);
        for (int t = 0; t < (int)A.tasks_total; ++t) G.order[t] = t;
        const int nbkp = A.nbkp;
        const bool row_major = (seed & 1) == 0;  // both claim orders of fast_long.cu (default: earliest start first)
        const int gate = A.gate;
        std::stable_sort(G.order.begin(), G.order.end(), [nbj, nbkp, row_major, gate](int x, int y) {
            if (row_major) return (x % nbj) * nbkp + x / nbj < (y % nbj) * nbkp + y / nbj;
            return long2_order_key<C>(x % nbj, x / nbj, gate) < long2_order_key<C>(y % nbj, y / nbj, gate);
        });
        for (int p = 0; p < (int)A.tasks_total; ++p) G.pos_of[G.order[p]] = p;
        A.order = G.order.data(); A.pos_of = G.pos_of.data();
        A.face_j = G.fj.data(); A.face_k = G.fk.data(); A.progress = G.prog.data(); A.counter = &G.counter;
        A.out = out8; A.out_states = out8 + 1; A.base = base;
        A.last_slab = g + 1 == n_ranges;
        if (g > 0) {
            G.cface.assign((size_t)nbj * slot_k, 0xDEADBEEFu);
            G.cprog.assign((size_t)nbj, 0);
        }
        G.ctas.resize(n_ctas);
    }
    for (int g = 0; g < n_ranges; ++g) {
        SweepArgs &A = gpus[g].A;
        if (g > 0) { A.cface_in = gpus[g].cface.data(); A.cprog_in = gpus[g].cprog.data(); }
        if (g + 1 < n_ranges) { A.cface_out = gpus[g + 1].cface.data(); A.cprog_out = gpus[g + 1].cprog.data(); }
    }
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)seed;
    for (int epoch = 1; epoch <= epochs; ++epoch) {
        for (auto &G : gpus) {
            G.A.epoch = epoch;
            G.counter = 0;
            G.finished = 0;
            std::fill(G.prog.begin(), G.prog.end(), 0);  // the launcher clears the local progress array
            for (auto &cta : G.ctas) {
                std::memset(cta.S.get(), (seed & 2) ? 0xFF : 0x00, sizeof(Long2Smem<C>));
                for (int tid = 0; tid < C::NT; ++tid) long2_load_a<C>(*cta.S, G.A, tid);
                cta.has_task = false;
            }
        }
        out8[0] = -99999;
        const int steps = long2_task_steps<C>(gpus[0].A);
        long long idle = 0;
        auto all_done = [&]() {
            for (auto &G : gpus)
                if (G.finished < G.A.tasks_total) return false;
            return true;
        };
        while (!all_done()) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            // skewed choice of the GPU: later ranges sometimes run far ahead of / behind their producers
            EmuGpu &G = gpus[(size_t)((rng >> 44) % (uint64_t)n_ranges)];
            SweepArgs &A = G.A;
            auto &cta = G.ctas[(size_t)((rng >> 40) % (uint64_t)n_ctas) % G.ctas.size()];
            bool progressed = false;
            if (!cta.has_task) {
                if ((long long)G.counter < A.tasks_total) {
                    SweepTask &T = cta.S->task;
                    sweep_describe_long<C>(A, (long long)G.counter++, T);
                    if (A.cface_in && T.bkp == 0) {
                        T.fk_r = A.cface_in + (long long)T.bj * A.fk_slot_words;
                        T.left_task = -2 - T.bj;
                    }
                    if (A.cface_out && T.bkp == A.nbkp - 1) T.fk_w = A.cface_out + (long long)T.bj * A.fk_slot_words;
                    cta.has_task = true;
                    cta.prologue_done = false;
                    for (int tid = 0; tid < C::NT; ++tid) long2_load_bc<C>(*cta.S, A, tid);
                    for (int tid = 0; tid < C::NT; ++tid) long2_build_tables<C>(*cta.S, tid);
                    for (int tid = 0; tid < C::NT; ++tid) {
                        int tj, tk;
                        stream_thread_pos<C>(tid, tj, tk);
                        long2_init_thread<C>(cta.R[tid], *cta.S, tj, tk);
                    }
                    std::fill(cta.done.begin(), cta.done.end(), -1);
                    progressed = true;
                }
            } else {
                const SweepTask &T = cta.S->task;
                const int tp = T.top_task >= 0 ? G.prog[T.top_task] : 0;
                const int lp = T.left_task >= 0 ? G.prog[T.left_task]
                                                : (T.left_task <= -2 ? emu_col_decode(G.cprog[-2 - T.left_task], A.epoch) : 0);
                int *right_p = (A.cprog_out && T.bkp == A.nbkp - 1) ? A.cprog_out + T.bj : nullptr;
                auto publish = [&](int done) {  // mirrors the kernel: the remote word follows the coarser rgate rhythm
                    G.prog[T.self_task] = done;
                    if (right_p && (done >= kProgressDone || (done & (A.rgate - 1)) == 0 || done == C::HOP_K + 1 + C::LAG))
                        *right_p = emu_col_encode(done, A.epoch);
                };
                const int cert = long2_certified<C>(A, T, tp, lp);
                if (!cta.prologue_done) {
                    if (cert >= 2) {
                        for (int tid = 0; tid < C::NT; ++tid) {
                            int tj, tk;
                            stream_thread_pos<C>(tid, tj, tk);
                            if (tj == 0 || tk == 0) long2_edge_fetch<C>(*cta.S, A, tj, tk, 0);
                            long2_thread_prepare<C>(cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, 0, tid);
                        }
                        cta.prologue_done = true;
                        progressed = true;
                    }
                } else {
                    const int w = (int)((rng >> 20) % NW);
                    const int s = cta.done[w] + 1;
                    const int mn = *std::min_element(cta.done.begin(), cta.done.end());
                    bool can = s <= steps && mn >= s - 1;
                    const int want = s + C::AHEAD < steps ? s + C::AHEAD : steps;
                    if (can && w == kGateWarp && cert < want) can = false;
                    if (can) {
                        if (w == kGateWarp && s >= 1 && long2_publish_now<C>(A, s)) publish(s);
                        const bool mid = s >= long2_mid_begin<C>(A) && s <= long2_mid_end<C>(A);
                        for (int lane = 0; lane < 32; ++lane) {
                            const int tid = w * 32 + lane;
                            int tj, tk;
                            stream_thread_pos<C>(tid, tj, tk);
                            if (mid) long2_thread_step<C, false, false>(cta.R[tid], cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s, tid);
                            else long2_thread_step<C, false, true>(cta.R[tid], cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s, tid);
                        }
                        for (int lane = 0; lane < 32; ++lane) {
                            const int tid = w * 32 + lane;
                            int tj, tk;
                            stream_thread_pos<C>(tid, tj, tk);
                            if (mid) long2_thread_prepare<C, false>(cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s + 1, tid);
                            else long2_thread_prepare<C, true>(cta.R[tid], cta.K[tid], *cta.S, A, tj, tk, s + 1, tid);
                        }
                        cta.done[w] = s;
                        progressed = true;
                        if (*std::min_element(cta.done.begin(), cta.done.end()) == steps) {
                            publish(kProgressDone);
                            cta.has_task = false;
                            ++G.finished;
                        }
                    }
                }
            }
            if (progressed) idle = 0;
            else if (++idle > 400000) return -2;  // deadlock
        }
        if (epoch < epochs) {
            // the face buffers are reused by the next epoch without clearing; poison them to catch a read that was not
            // certified in its own epoch
            for (auto &G : gpus) std::fill(G.cface.begin(), G.cface.end(), 0xDEADBEEFu);
        }
    }
    return 0;
}
