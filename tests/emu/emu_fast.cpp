// emu_fast.cpp -- host emulation of the fast kernels' CTA choreography (test infrastructure).
// The per-thread step functions in csrc/fast_*.cuh are __host__ __device__; here every "thread" of a CTA is run
// sequentially inside each barrier-delimited step, which is exactly the kernel's schedule (threads only read what
// was published before the previous barrier).  Lets the CPU test-suite check the skewing, halo exchange, pair
// hand-over and score extraction against the oracle without a GPU.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../hw-accelerator-three-sequence-alignment_b200/csrc/fast_batch.cuh"

using namespace trialign::fast;

extern "C" int emu_fast_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, long long n, int len, int base,
                              int pairs_per_cta, int32_t *scores)
{
    using C = BatchCfg<4, 4, 16, 16>;
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
        std::memset(S->acode[1], 0, sizeof S->acode[1]);  // as the kernel prologue does
        std::memset(S->seq[1], 0, sizeof S->seq[1]);
        for (int tid = 0; tid < C::NT; ++tid) load_pair_sequences<C>(*S, A, tid, pair_base, 0);
        for (int tid = 0; tid < C::NT; ++tid) build_pair_tables<C>(*S, tid, 0);
        const int steps = batch_total_steps<C>(P);
        for (int s = 0; s <= steps; ++s) {
            for (int tid = 0; tid < C::NT; ++tid) {
                int tj, tk;
                thread_pos<C>(tid, tj, tk);
                batch_uniform_actions<C>(*S, A, tid, s, P, pair_base);
                batch_thread_step<C>(R[tid], *S, A, tj, tk, s, P, pair_base);
            }
        }
    }
    return 0;
}
