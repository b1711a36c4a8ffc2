// kernels.cuh -- entry points of the CUDA kernel families behind the C ABI (all pointers are device memory).
#pragma once

#include <climits>
#include <cstdint>

#include "common.cuh"

namespace trialign {

// KL: literal plane-sweep kernel, exact for any score_bits in [2,32] (literal_kernel.cu).
// d_a/d_b/d_c are [n][la] / [n][lb] / [n][lc]; d_scores [n]; d_finals [n][7] or nullptr.
int literal_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                int64_t n, int la, int lb, int lc, int score_bits, int32_t *d_scores, int32_t *d_finals,
                const trialign_scoring *scoring = nullptr);  // nullptr: the RTL's constants

// K1: batched systolic kernel, two triples per 32-bit lane, len <= 64 (fast_batch.cu).  Exact only where no
// candidate wraps at score_bits -- the dispatcher checks provably_wrap_free() first.
bool fast_batch_supports(int len);
// check_bits > 0: the range-checked variant; d_pair_flags[(n+1)/2] (zeroed by the caller) gets bit 0 / 1 set for the
// low / high triple of a pair whose values came too close to the score_bits = check_bits limits.
// scoring != nullptr (always with check_bits > 0): run-time scoring parameters inside score_fast_domain().
int fast_batch_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                   int64_t n, int len, int32_t *d_scores, int check_bits = 0, int *d_pair_flags = nullptr,
                   const trialign_scoring *scoring = nullptr);
// Non-RTL scoring parameters the factored kernels can evaluate (fast_core.cuh): small non-negative gap costs with
// GE <= GO, MISMATCH <= MATCH, every potential-form constant (s + 3 GE, S3 + 6 GE) non-negative, the RTL triple rule
// factorisable as min(row, column) constant, and a score_bits whose limits leave room for the range check's margins.
bool score_fast_domain(const trialign_scoring &sc, int score_bits);

// K2L: block-sweep kernel for ONE long triple (fast_long.cu); same exactness condition as K1 plus the 16-bit envelope
// checked by fast_long_supports().
bool fast_long_supports(int la, int lb, int lc);
// 16-bit range of a launch over `planes` planes of a triple of la_total x lb x lc (i-slab pipeline: planes < la_total)
bool fast_long_range_ok(int planes, int la_total, int lb, int lc);
// i-slab pipeline of one long triple over several GPUs: a launch computes planes (i_begin, i_begin+la] and hands the
// plane state of every block task to the next slab.  in_* live on this GPU, out_* may be peer memory.
struct SlabLink {
    int la_total;          // planes of the whole triple
    int last;              // this slab ends at plane la_total: emit the score
    int epoch;             // flag value of this run
    const void *in_state;  // nullptr for the first slab
    const void *in_flags;
    void *out_state;       // nullptr for the last slab
    void *out_flags;
};
size_t fast_long_slab_state_bytes(int lb, int lc);
size_t fast_long_slab_flag_bytes(int lb, int lc);
// K2L runs on the K1 / K2S step engine (split-phase or lagged mbarrier, block halos prefetched by the edge threads).
// column pipeline of one long triple over several GPUs: a launch sweeps ALL planes of a contiguous range of block-column
// pairs; the K faces crossing the range's edges stream to / from the neighbour GPUs through peer memory, gated by
// epoch-tagged progress words at system scope.  in_* live on this GPU, out_* are peer pointers.
struct ColLink {
    int lc_total;          // columns of the whole triple
    int last;              // this range holds column lc_total: emit the score
    int epoch;             // tag of this run's progress words (1 .. 2047)
    const void *in_face;   // nullptr for the first range
    const void *in_prog;
    void *out_face;        // nullptr for the last range
    void *out_prog;
};
size_t fast_long2_col_face_bytes(int la, int lb);
size_t fast_long2_col_prog_bytes(int lb);
int fast_long2_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb,
                   const uint8_t *d_c, int lc, int32_t *d_score, int32_t *d_states, int check_bits, int *d_viol,
                   const SlabLink *slab = nullptr, const ColLink *col = nullptr);
// check_bits > 0 selects the range-checked kernel: *d_viol (zeroed by the caller) becomes non-zero if some value came
// close enough to the score_bits = check_bits limits that a wrap cannot be excluded.

// K2S: streaming block-sweep kernel for batches of equal-length triples longer than 64 (fast_stream.cu): all blocks of
// a CTA's pairs stream through the systolic array back to back.  Same exactness condition as K1 / K2.
bool fast_stream_supports(int len);
bool fast_stream_fits(int len, int64_t n, int sm_count);  // CTA-private face rings within the memory budget
int fast_stream_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                    int64_t n, int len, int32_t *d_scores, int check_bits = 0, int *d_pair_flags = nullptr,
                    const trialign_scoring *scoring = nullptr);
// K2S "P2" (fast_stream2.cu): two planes per step, block-edge duties by thread id; the default for len >= 65, also
// range-checked and with run-time scoring parameters (option k2s=p1 keeps the one-plane kernels)
bool fast_stream2_supports(int len);
int fast_stream2_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                     int64_t n, int len, int32_t *d_scores, int check_bits = 0, int *d_pair_flags = nullptr,
                     const trialign_scoring *scoring = nullptr);
// run-time scoring parameters on K2S: score_fast_domain() plus the 16-bit lane range at this length
bool fast_stream_score_ok(const trialign_scoring &sc, int len, int score_bits);

// flagged-triple plumbing of the checked batch path (batch_fixup.cu): compaction of the per-pair flags into a list of
// triple indices, gather of those triples into contiguous arrays, scatter of their re-computed scores
int fixup_compact(cudaStream_t st, const int *d_pair_flags, int64_t n, int *d_list, int *d_count);
int fixup_gather(cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c, int len, const int *d_list,
                 int count, uint8_t *g_a, uint8_t *g_b, uint8_t *g_c);
int fixup_scatter(cudaStream_t st, const int32_t *g_scores, const int *d_list, int count, int32_t *d_scores);

// KT: traceback of one triple (traceback.cu).  d_out: 5 int32 (score, start cell i/j/k, number of columns); d_moves:
// la+lb+lc bytes, forward order, bit0/1/2 = A/B/C advances.
size_t traceback_dir_bytes(int la, int lb, int lc);
int traceback_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb,
                  const uint8_t *d_c, int lc, int score_bits, int32_t *d_out, uint8_t *d_moves);

// all triples of a sequence set (all_triples.cu): lays triples first..first+count-1 (lexicographic x<y<z, roles
// A=x, B=y, C=z) out as [count][len] batch arrays on the device
long long all_triples_count(int nseq);
int all_triples_gather(cudaStream_t st, const uint8_t *d_seqs, int nseq, int len, long long first, long long count,
                       uint8_t *d_a, uint8_t *d_b, uint8_t *d_c);

// the testbench's packed 128-bit sequence RAM format, expanded on the device (seq_formats.cu): d_ram is
// [n][ceil(len/32)][16] bytes, d_out [n][len] symbol bytes
int unpack_ram128_run(cudaStream_t st, const uint8_t *d_ram, int64_t n, int len, uint8_t *d_out);

// synthetic generator on the device (synth_kernel.cu)
int synth_run(cudaStream_t st, uint64_t seed, int64_t first, int64_t n, int len, uint8_t *d_a, uint8_t *d_b,
              uint8_t *d_c);

}  // namespace trialign
