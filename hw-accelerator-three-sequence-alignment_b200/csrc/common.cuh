// common.cuh -- shared host-side plumbing of libtrialign: error capture, per-device context, launch counter.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/trialign.h"

namespace trialign {

// ---- error capture (thread-local, never throws across the C ABI) ----
struct ThreadState {
    std::string cuda_error;
    double last_seconds = -1.0;
    double last_cells = 0.0;
    int last_path = TRIALIGN_PATH_NONE;
    int last_devices = 0;         // GPUs that served the last compute call of this thread
    int64_t last_flagged = 0;     // triples the range-checked batch path re-scored on the literal kernel (last call)
    int last_segments = 0;        // segments of the last traceback (1 = whole direction cube in memory)
    bool scoring_active = false;  // a *_ex entry point with non-RTL scoring is in progress on this thread
    trialign_scoring scoring{};
};
ThreadState &tls();

int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define TA_CUDA(expr)                                                            \
    do {                                                                         \
        cudaError_t e__ = (expr);                                                \
        if (e__ != cudaSuccess) return ::trialign::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

// ---- run-time options (trialign_set_option): experiment switches live here, never in the environment ----
// Known names: see kOptionNames in api.cu.  opt_int() returns dflt when the option is unset or not a number.
int opt_int(const char *name, int dflt);
bool opt_is(const char *name, const char *value);
inline bool opt_debug() { return opt_int("debug", 0) != 0; }

// ---- keeps the calling thread's current CUDA device unchanged across an entry point ----
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// ---- NVTX ranges around the phases of the host-pointer entry points (H2D / kernels / D2H) ----
struct NvtxRange {
    explicit NvtxRange(const char *name);
    ~NvtxRange();
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

extern std::atomic<int64_t> g_launches;
inline void count_launch(int64_t k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

// ---- grow-only device scratch buffer ----
struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need);  // TRIALIGN_OK / ENOMEM (CUDA errors recorded)
    void release();
};

// ---- lazily created per-device context ----
struct DeviceCtx {
    int device = -1;
    bool ready = false;
    std::mutex mu;  // serialises host-pointer calls that share the scratch buffers
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    Scratch seq, out, planes, faces, misc;
    Scratch dirs;        // direction cube (or one segment of it) of the traceback path
    Scratch ckpt;        // checkpointed traceback: the rolling planes parked at the segment boundaries
    Scratch viol;        // range-check flag of the checked fast path (never reallocated by the kernels)
    Scratch flagged;     // per-triple range-check flags + compaction buffers of the checked batch path
    // Stream ordering of the shared scratch buffers: every enqueue that touches them records scratch_ev on its stream;
    // a later enqueue on ANOTHER stream first waits for that event (ScratchLease), so two calls in flight on different
    // streams never run kernels against the same faces / planes / progress words.  Guarded by mu.
    cudaEvent_t scratch_ev = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_used = false;
};

// Taken (with ctx->mu held) around every enqueue that uses the context's scratch buffers.
struct ScratchLease {
    DeviceCtx *ctx;
    cudaStream_t st;
    ScratchLease(DeviceCtx *c, cudaStream_t s) : ctx(c), st(s)
    {
        if (ctx->scratch_used && ctx->scratch_stream != st) cudaStreamWaitEvent(st, ctx->scratch_ev, 0);
    }
    ~ScratchLease()
    {
        cudaEventRecord(ctx->scratch_ev, st);
        ctx->scratch_stream = st;
        ctx->scratch_used = true;
    }
    ScratchLease(const ScratchLease &) = delete;
    ScratchLease &operator=(const ScratchLease &) = delete;
};

constexpr int kMaxDevices = 32;
int get_ctx(int device, DeviceCtx **ctx);  // sets the current device (callers hold a DeviceGuard), creates stream/events on first use

// wrap-free envelope of the recurrence: with zero borders every state of cell (i,j,k)
// lies in [-(2*m+5), 3*m] with m = min(i,j,k) (at most m three-way matches; gap chains from the zero border lose
// GE2 per step).  A margin is added (states in [-(2*m+8), 3*m+2]); candidates add at most +3 / -4 to a state.
inline int64_t envelope_hi(int64_t m) { return 3 * m + 2; }
inline int64_t envelope_lo(int64_t m) { return -(2 * m + 8); }
inline bool provably_wrap_free(int la, int lb, int lc, int score_bits)
{
    if (score_bits >= 32) return true;
    const int64_t m = la < lb ? (la < lc ? la : lc) : (lb < lc ? lb : lc);
    const int64_t hi = envelope_hi(m) + 3, lo = envelope_lo(m) - 4;
    const int64_t lim = (int64_t)1 << (score_bits - 1);
    return hi <= lim - 1 && lo >= -lim;
}

}  // namespace trialign
