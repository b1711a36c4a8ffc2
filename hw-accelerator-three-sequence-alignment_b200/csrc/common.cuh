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
    Scratch dirs;        // direction cube of the traceback path
    Scratch viol;        // range-check flag of the checked fast path (never reallocated by the kernels)
};

constexpr int kMaxDevices = 32;
int get_ctx(int device, DeviceCtx **ctx);  // sets the current device, creates stream/events on first use

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
