// api.cu -- the C ABI of libtrialign (include/trialign.h): argument checking, per-device contexts, host<->device
// staging, kernel-family dispatch.  No CPU compute path exists here: without a CUDA device the compute entry
// points fail with TRIALIGN_ENODEVICE.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "kernels.cuh"

namespace trialign {

std::atomic<int64_t> g_launches{0};

ThreadState &tls()
{
    thread_local ThreadState s;
    return s;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s: %s (%s:%d)", cudaGetErrorName(e), what, file, line);
    tls().cuda_error = buf;
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();  // clear the sticky-less OOM so later calls can proceed
        return TRIALIGN_ENOMEM;
    }
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
        return TRIALIGN_ENODEVICE;
    return TRIALIGN_ECUDA;
}

int Scratch::reserve(size_t need)
{
    if (need <= bytes) return TRIALIGN_OK;
    if (ptr) {
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    // grow geometrically to limit re-allocation churn, but never beyond what was asked by more than 25 %
    const size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&ptr, need);
        if (e != cudaSuccess) {
            ptr = nullptr;
            return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        }
        bytes = need;
        return TRIALIGN_OK;
    }
    bytes = want;
    return TRIALIGN_OK;
}

void Scratch::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

static DeviceCtx g_ctx[kMaxDevices];
static std::mutex g_ctx_mu;

int get_ctx(int device, DeviceCtx **out)
{
    if (device < 0 || device >= kMaxDevices) return TRIALIGN_EINVAL;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        tls().cuda_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorName(e);
        return TRIALIGN_ENODEVICE;
    }
    if (device >= count) return count == 0 ? TRIALIGN_ENODEVICE : TRIALIGN_EINVAL;
    TA_CUDA(cudaSetDevice(device));
    DeviceCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!c.ready) {
        c.device = device;
        TA_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        TA_CUDA(cudaEventCreate(&c.ev0));
        TA_CUDA(cudaEventCreate(&c.ev1));
        TA_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
        c.ready = true;
    }
    *out = &c;
    return TRIALIGN_OK;
}

namespace {

bool env_is(const char *name, const char *value)
{
    const char *e = std::getenv(name);
    return e && std::strcmp(e, value) == 0;
}

// one long triple: K2L (fast_long.cu); TRIALIGN_K2L=old selects the first-generation kernel (fast_sweep.cu) for A/B runs
int long_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
             int lc, int32_t *d_score, int32_t *d_states, int check_bits, int *d_viol, const SlabLink *slab = nullptr)
{
    if (env_is("TRIALIGN_K2L", "old"))
        return fast_sweep_long_run(ctx, st, d_a, la, d_b, lb, d_c, lc, d_score, d_states, check_bits, d_viol, slab);
    return fast_long2_run(ctx, st, d_a, la, d_b, lb, d_c, lc, d_score, d_states, check_bits, d_viol, slab);
}

bool backend_is_gpu(int backend)
{
    return backend == TRIALIGN_BACKEND_AUTO || backend == TRIALIGN_BACKEND_GPU ||
           backend == TRIALIGN_BACKEND_GPU_LITERAL || backend == TRIALIGN_BACKEND_GPU_FAST;
}

int check_backend(int backend)
{
    if (backend_is_gpu(backend)) return TRIALIGN_OK;
    if (backend == TRIALIGN_BACKEND_CPU_ORACLE || backend == TRIALIGN_BACKEND_RTL_1CYC ||
        backend == TRIALIGN_BACKEND_RTL_2CYC)
        return TRIALIGN_EUNSUPPORTED;
    return TRIALIGN_EINVAL;
}

// Dispatch of one uniform batch on one device; every pointer is device memory.  `synced_ok` tells whether the
// dispatcher may synchronise the stream (needed by the fast path's range check before a literal re-run).
int dispatch_batch(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c,
                   int64_t n, int la, int lb, int lc, int score_bits, int backend, int32_t *d_scores,
                   int32_t *d_finals)
{
    ThreadState &ts = tls();
    if (ts.scoring_active) {  // non-RTL scoring parameters (*_ex entry points): the literal kernel evaluates them as given
        if (backend == TRIALIGN_BACKEND_GPU_FAST) return TRIALIGN_EINEXACT;
        ts.last_path = TRIALIGN_PATH_LITERAL;
        return literal_run(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, d_scores, d_finals, &ts.scoring);
    }
    const bool want_fast = backend != TRIALIGN_BACKEND_GPU_LITERAL;
    // fast paths are used only where they are provably bit-exact: no candidate can wrap at score_bits
    if (want_fast && provably_wrap_free(la, lb, lc, score_bits)) {
        if (!d_finals && la == lb && lb == lc && la <= 64 && fast_stream_supports(la) && env_is("TRIALIGN_K1", "stream")) {
            ts.last_path = TRIALIGN_PATH_FAST_STREAM;  // experiment: K2S on a single block per pair (len 61..64)
            return fast_stream_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
        if (!d_finals && la == lb && lb == lc && fast_batch_supports(la)) {
            ts.last_path = TRIALIGN_PATH_FAST_BATCH;
            return fast_batch_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
        if (n == 1 && fast_sweep_supports(la, lb, lc)) {
            ts.last_path = TRIALIGN_PATH_FAST_LONG;
            return long_run(ctx, st, d_a, la, d_b, lb, d_c, lc, d_scores, d_finals, 0, nullptr);
        }
        if (!d_finals && la == lb && lb == lc && fast_stream_supports(la) && fast_stream_fits(la, n, ctx->sm_count) &&
            !env_is("TRIALIGN_K2", "sweep")) {  // TRIALIGN_K2=sweep: the per-block kernel below, for A/B runs
            ts.last_path = TRIALIGN_PATH_FAST_STREAM;
            return fast_stream_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
        if (!d_finals && la == lb && lb == lc && fast_sweep_supports(la, lb, lc)) {
            ts.last_path = TRIALIGN_PATH_FAST_LONG;
            return fast_sweep_batch_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
    }
    // One long triple whose wrap-freedom cannot be proved from the lengths: run the fast kernel with its run-time
    // range check, and fall back to the literal kernel only if some value came near the score_bits limits.
    if (want_fast && n == 1 && score_bits >= 10 && score_bits <= 16 && fast_sweep_supports(la, lb, lc)) {
        int rc = ctx->viol.reserve(256);
        if (rc) return rc;
        int *d_viol = (int *)ctx->viol.ptr;
        TA_CUDA(cudaMemsetAsync(d_viol, 0, sizeof(int), st));
        rc = long_run(ctx, st, d_a, la, d_b, lb, d_c, lc, d_scores, d_finals, score_bits, d_viol);
        if (rc) return rc;
        int viol = 0;
        TA_CUDA(cudaMemcpyAsync(&viol, d_viol, sizeof(int), cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
        if (!viol) {
            ts.last_path = TRIALIGN_PATH_FAST_LONG;
            return TRIALIGN_OK;
        }
        if (backend == TRIALIGN_BACKEND_GPU_FAST) return TRIALIGN_EINEXACT;
        ts.last_path = TRIALIGN_PATH_FAST_THEN_LITERAL;
        return literal_run(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, d_scores, d_finals);
    }
    if (backend == TRIALIGN_BACKEND_GPU_FAST) return TRIALIGN_EINEXACT;
    ts.last_path = TRIALIGN_PATH_LITERAL;
    return literal_run(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, d_scores, d_finals);
}

int first_device(unsigned devices)
{
    if (devices == 0) return 0;
    for (int d = 0; d < kMaxDevices; ++d)
        if (devices & (1u << d)) return d;
    return 0;
}

// host-pointer batch on ONE device: stage in, run, stage out; kernel time measured with events on the stream
int run_host_batch_on_device(int device, const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int la,
                             int lb, int lc, int score_bits, int backend, int32_t *scores, int32_t *finals,
                             double *seconds)
{
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    const size_t na = (size_t)n * la, nb = (size_t)n * lb, nc = (size_t)n * lc;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    if ((rc = ctx->seq.reserve(up(na) + up(nb) + up(nc) + 256))) return rc;
    const size_t n_pad = ((size_t)n + 63) & ~(size_t)63;  // finals start on a 256-byte boundary
    if ((rc = ctx->out.reserve(n_pad * sizeof(int32_t) + (finals ? (size_t)n * 7 * sizeof(int32_t) : 0)))) return rc;
    uint8_t *d_a = (uint8_t *)ctx->seq.ptr, *d_b = d_a + up(na), *d_c = d_b + up(nb);
    int32_t *d_scores = (int32_t *)ctx->out.ptr;
    int32_t *d_finals = finals ? d_scores + n_pad : nullptr;
    TA_CUDA(cudaMemcpyAsync(d_a, a, na, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_b, b, nb, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_c, c, nc, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaEventRecord(ctx->ev0, st));
    rc = dispatch_batch(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, backend, d_scores, d_finals);
    if (rc) {
        cudaStreamSynchronize(st);
        return rc;
    }
    TA_CUDA(cudaEventRecord(ctx->ev1, st));
    TA_CUDA(cudaMemcpyAsync(scores, d_scores, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (finals) TA_CUDA(cudaMemcpyAsync(finals, d_finals, (size_t)n * 7 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    TA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (seconds) *seconds = (double)ms * 1e-3;
    return TRIALIGN_OK;
}

}  // namespace
}  // namespace trialign

using namespace trialign;

extern "C" {

int trialign_version(void) { return TRIALIGN_VERSION; }

const char *trialign_strerror(int rc)
{
    switch (rc) {
    case TRIALIGN_OK: return "ok";
    case TRIALIGN_EINVAL: return "invalid argument";
    case TRIALIGN_ENOMEM: return "out of memory";
    case TRIALIGN_ECUDA: return "CUDA error (see trialign_last_cuda_error)";
    case TRIALIGN_EIO: return "file I/O error";
    case TRIALIGN_EPARSE: return "malformed input file";
    case TRIALIGN_EUNSUPPORTED: return "backend not part of libtrialign (CPU checkers live under oracle/)";
    case TRIALIGN_ENODEVICE: return "no usable CUDA device";
    case TRIALIGN_ECAPACITY: return "output buffer too small";
    case TRIALIGN_EINEXACT: return "fast kernels cannot be proven bit-exact for this input";
    default: return "unknown error";
    }
}

const char *trialign_last_cuda_error(void) { return tls().cuda_error.c_str(); }

int trialign_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count < 0 ? 0 : count;
}

int trialign_score_batch_dev(int device, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c, int64_t n,
                             int len, int score_bits, int backend, int32_t *d_scores, void *stream)
{
    if (n < 0 || len < 0 || score_bits < 2 || score_bits > 32) return TRIALIGN_EINVAL;
    if (n > 0 && !d_scores) return TRIALIGN_EINVAL;
    if (n > 0 && len > 0 && (!d_a || !d_b || !d_c)) return TRIALIGN_EINVAL;
    int rc = check_backend(backend);
    if (rc) return rc;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)n * len * len * len;
    ts.last_path = TRIALIGN_PATH_NONE;
    if (n == 0) return TRIALIGN_OK;
    DeviceCtx *ctx = nullptr;
    if ((rc = get_ctx(device, &ctx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (len == 0) {
        TA_CUDA(cudaMemsetAsync(d_scores, 0, (size_t)n * sizeof(int32_t), st));
        return TRIALIGN_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    return dispatch_batch(ctx, st, d_a, d_b, d_c, n, len, len, len, score_bits, backend, d_scores, nullptr);
}

int trialign_score_dev(int device, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
                       int lc, int score_bits, int backend, int32_t *d_out, void *stream)
{
    if (la < 0 || lb < 0 || lc < 0 || score_bits < 2 || score_bits > 32 || !d_out) return TRIALIGN_EINVAL;
    if ((la && !d_a) || (lb && !d_b) || (lc && !d_c)) return TRIALIGN_EINVAL;
    int rc = check_backend(backend);
    if (rc) return rc;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)la * lb * lc;
    ts.last_path = TRIALIGN_PATH_NONE;
    DeviceCtx *ctx = nullptr;
    if ((rc = get_ctx(device, &ctx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (la == 0 || lb == 0 || lc == 0) {
        TA_CUDA(cudaMemsetAsync(d_out, 0, 8 * sizeof(int32_t), st));
        return TRIALIGN_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    return dispatch_batch(ctx, st, d_a, d_b, d_c, 1, la, lb, lc, score_bits, backend, d_out, d_out + 1);
}

int trialign_score(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                   int backend, unsigned devices, int32_t *score, int32_t final_states[7])
{
    if (la < 0 || lb < 0 || lc < 0 || score_bits < 2 || score_bits > 32 || !score) return TRIALIGN_EINVAL;
    if ((la && !a) || (lb && !b) || (lc && !c)) return TRIALIGN_EINVAL;
    int rc = check_backend(backend);
    if (rc) return rc;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)la * lb * lc;
    ts.last_path = TRIALIGN_PATH_NONE;
    if (trialign_device_count() == 0) return TRIALIGN_ENODEVICE;
    if (la == 0 || lb == 0 || lc == 0) {
        *score = 0;
        if (final_states) std::memset(final_states, 0, 7 * sizeof(int32_t));
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    int32_t fin[7];
    double sec = -1.0;
    rc = run_host_batch_on_device(first_device(devices), a, b, c, 1, la, lb, lc, score_bits, backend, score, fin,
                                  &sec);
    if (rc) return rc;
    if (final_states) std::memcpy(final_states, fin, sizeof fin);
    ts.last_seconds = sec;
    return TRIALIGN_OK;
}

int64_t trialign_triple_count(int nseq) { return nseq < 3 ? 0 : (int64_t)all_triples_count(nseq); }

int trialign_score_all_triples(const uint8_t *seqs, int nseq, int len, int score_bits, int backend, unsigned devices,
                               int32_t *scores)
{
    if (nseq < 0 || len < 0 || score_bits < 2 || score_bits > 32) return TRIALIGN_EINVAL;
    const int64_t total = trialign_triple_count(nseq);
    if (total > 0 && (!scores || (len > 0 && !seqs))) return TRIALIGN_EINVAL;
    int rc = check_backend(backend);
    if (rc) return rc;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)total * len * len * len;
    ts.last_path = TRIALIGN_PATH_NONE;
    if (trialign_device_count() == 0) return TRIALIGN_ENODEVICE;
    if (total == 0) return TRIALIGN_OK;
    if (len == 0) {
        std::memset(scores, 0, (size_t)total * sizeof(int32_t));
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    DeviceCtx *ctx = nullptr;
    if ((rc = get_ctx(first_device(devices), &ctx))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    // chunks of at most 2^20 triples bound the staging arrays (3 x chunk x len bytes) whatever nseq is
    const int64_t chunk = total < (1 << 20) ? total : (1 << 20);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t set_bytes = up((size_t)nseq * len), arr = up((size_t)chunk * len);
    if ((rc = ctx->seq.reserve(set_bytes + 3 * arr + 256))) return rc;
    if ((rc = ctx->out.reserve((size_t)chunk * sizeof(int32_t)))) return rc;
    uint8_t *d_set = (uint8_t *)ctx->seq.ptr, *d_a = d_set + set_bytes, *d_b = d_a + arr, *d_c = d_b + arr;
    int32_t *d_scores = (int32_t *)ctx->out.ptr;
    TA_CUDA(cudaMemcpyAsync(d_set, seqs, (size_t)nseq * len, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaEventRecord(ctx->ev0, st));
    for (int64_t t0 = 0; t0 < total; t0 += chunk) {
        const int64_t cnt = total - t0 < chunk ? total - t0 : chunk;
        if ((rc = all_triples_gather(st, d_set, nseq, len, t0, cnt, d_a, d_b, d_c)) ||
            (rc = dispatch_batch(ctx, st, d_a, d_b, d_c, cnt, len, len, len, score_bits, backend, d_scores, nullptr))) {
            cudaStreamSynchronize(st);
            return rc;
        }
        TA_CUDA(cudaMemcpyAsync(scores + t0, d_scores, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    TA_CUDA(cudaEventRecord(ctx->ev1, st));
    TA_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    TA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ts.last_seconds = (double)ms * 1e-3;
    return TRIALIGN_OK;
}

int trialign_align(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                   unsigned devices, int32_t *score, int32_t start[3], uint8_t *moves, int cap, int *n_moves)
{
    if (la < 0 || lb < 0 || lc < 0 || score_bits < 2 || score_bits > 32 || !score || !start || !n_moves) return TRIALIGN_EINVAL;
    if ((la && !a) || (lb && !b) || (lc && !c) || cap < 0 || (cap && !moves)) return TRIALIGN_EINVAL;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)la * lb * lc;
    ts.last_path = TRIALIGN_PATH_NONE;
    if (trialign_device_count() == 0) return TRIALIGN_ENODEVICE;
    *n_moves = 0;
    start[0] = la; start[1] = lb; start[2] = lc;
    if (la == 0 || lb == 0 || lc == 0) {  // already on a border plane: empty path, score 0
        *score = 0;
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(first_device(devices), &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t total = (size_t)la + lb + lc;
    if ((rc = ctx->seq.reserve(up(la) + up(lb) + up(lc) + 256))) return rc;
    if ((rc = ctx->out.reserve(256 + up(total)))) return rc;
    uint8_t *d_a = (uint8_t *)ctx->seq.ptr, *d_b = d_a + up(la), *d_c = d_b + up(lb);
    int32_t *d_out = (int32_t *)ctx->out.ptr;
    uint8_t *d_moves = (uint8_t *)ctx->out.ptr + 256;
    TA_CUDA(cudaMemcpyAsync(d_a, a, la, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_b, b, lb, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaMemcpyAsync(d_c, c, lc, cudaMemcpyHostToDevice, st));
    TA_CUDA(cudaEventRecord(ctx->ev0, st));
    ts.last_path = TRIALIGN_PATH_TRACEBACK;
    rc = traceback_run(ctx, st, d_a, la, d_b, lb, d_c, lc, score_bits, d_out, d_moves);
    if (rc) {
        cudaStreamSynchronize(st);
        return rc;
    }
    TA_CUDA(cudaEventRecord(ctx->ev1, st));
    int32_t head[5];
    TA_CUDA(cudaMemcpyAsync(head, d_out, sizeof head, cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    TA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ts.last_seconds = (double)ms * 1e-3;
    *score = head[0];
    start[0] = head[1]; start[1] = head[2]; start[2] = head[3];
    *n_moves = head[4];
    if (head[4] > cap) return TRIALIGN_ECAPACITY;
    if (head[4] > 0) TA_CUDA(cudaMemcpy(moves, d_moves, (size_t)head[4], cudaMemcpyDeviceToHost));
    return TRIALIGN_OK;
}

int trialign_score_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len, int score_bits,
                         int backend, unsigned devices, int32_t *scores)
{
    if (n < 0 || len < 0 || score_bits < 2 || score_bits > 32) return TRIALIGN_EINVAL;
    if (n > 0 && !scores) return TRIALIGN_EINVAL;
    if (n > 0 && len > 0 && (!a || !b || !c)) return TRIALIGN_EINVAL;
    int rc = check_backend(backend);
    if (rc) return rc;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)n * len * len * len;
    ts.last_path = TRIALIGN_PATH_NONE;
    const int have = trialign_device_count();
    if (have == 0) return TRIALIGN_ENODEVICE;
    if (n == 0) return TRIALIGN_OK;
    if (len == 0) {
        std::memset(scores, 0, (size_t)n * sizeof(int32_t));
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    std::vector<int> devs;
    for (int d = 0; d < kMaxDevices; ++d)
        if (devices & (1u << d)) devs.push_back(d);
    if (devs.empty()) devs.push_back(0);
    for (int d : devs)
        if (d >= have) return TRIALIGN_EINVAL;
    if (devs.size() == 1) {
        double sec = -1.0;
        rc = run_host_batch_on_device(devs[0], a, b, c, n, len, len, len, score_bits, backend, scores, nullptr, &sec);
        ts.last_seconds = sec;
        return rc;
    }
    // contiguous split across devices, one host thread per device, no exchange
    const size_t G = devs.size();
    std::vector<int> rcs(G, 0), paths(G, 0);
    std::vector<double> secs(G, -1.0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
        const int64_t t0 = n * (int64_t)g / (int64_t)G, t1 = n * (int64_t)(g + 1) / (int64_t)G;
        th.emplace_back([&, g, t0, t1]() {
            tls().scoring_active = ts.scoring_active;  // the *_ex scoring of the calling thread
            tls().scoring = ts.scoring;
            if (t1 > t0)
                rcs[g] = run_host_batch_on_device(devs[g], a + (size_t)t0 * len, b + (size_t)t0 * len,
                                                  c + (size_t)t0 * len, t1 - t0, len, len, len, score_bits,
                                                  backend, scores + t0, nullptr, &secs[g]);
            paths[g] = tls().last_path;
            errs[g] = tls().cuda_error;
        });
    }
    for (auto &t : th) t.join();
    double worst = 0.0;
    for (size_t g = 0; g < G; ++g) {
        if (rcs[g]) {
            ts.cuda_error = errs[g];
            return rcs[g];
        }
        worst = std::max(worst, secs[g]);
        if (paths[g] > ts.last_path) ts.last_path = paths[g];
    }
    ts.last_seconds = worst;
    return TRIALIGN_OK;
}

// ---- explicit scoring parameters ----
void trialign_scoring_rtl(trialign_scoring *sc)
{
    if (!sc) return;
    sc->match = 1;       // src/PE_1cyc.v:53-56
    sc->mismatch = -1;
    sc->gap_open = 2;
    sc->gap_extend = 1;
    sc->triple_rule = TRIALIGN_TRIPLE_RTL;
}

namespace {

int check_scoring(const trialign_scoring *sc, bool *is_rtl)
{
    if (!sc) return TRIALIGN_EINVAL;
    if (sc->match < -1024 || sc->match > 1024 || sc->mismatch < -1024 || sc->mismatch > 1024) return TRIALIGN_EINVAL;
    if (sc->gap_open < 0 || sc->gap_open > 1024 || sc->gap_extend < 0 || sc->gap_extend > 1024) return TRIALIGN_EINVAL;
    if (sc->triple_rule != TRIALIGN_TRIPLE_RTL && sc->triple_rule != TRIALIGN_TRIPLE_SUM_OF_PAIRS) return TRIALIGN_EINVAL;
    trialign_scoring rtl;
    trialign_scoring_rtl(&rtl);
    *is_rtl = std::memcmp(sc, &rtl, sizeof rtl) == 0;
    return TRIALIGN_OK;
}

// sets the calling thread's scoring for the duration of one *_ex call
struct ScopedScoring {
    explicit ScopedScoring(const trialign_scoring &sc)
    {
        tls().scoring = sc;
        tls().scoring_active = true;
    }
    ~ScopedScoring() { tls().scoring_active = false; }
};

}  // namespace

int trialign_score_ex(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                      const trialign_scoring *scoring, unsigned devices, int32_t *score, int32_t final_states[7])
{
    bool is_rtl = false;
    const int rc = check_scoring(scoring, &is_rtl);
    if (rc) return rc;
    if (is_rtl) return trialign_score(a, la, b, lb, c, lc, score_bits, TRIALIGN_BACKEND_AUTO, devices, score, final_states);
    ScopedScoring scope(*scoring);
    return trialign_score(a, la, b, lb, c, lc, score_bits, TRIALIGN_BACKEND_AUTO, devices, score, final_states);
}

int trialign_score_batch_ex(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len, int score_bits,
                            const trialign_scoring *scoring, unsigned devices, int32_t *scores)
{
    bool is_rtl = false;
    const int rc = check_scoring(scoring, &is_rtl);
    if (rc) return rc;
    if (is_rtl) return trialign_score_batch(a, b, c, n, len, score_bits, TRIALIGN_BACKEND_AUTO, devices, scores);
    ScopedScoring scope(*scoring);
    return trialign_score_batch(a, b, c, n, len, score_bits, TRIALIGN_BACKEND_AUTO, devices, scores);
}

int trialign_slab_buffer_bytes(int lb, int lc, int64_t *state_bytes, int64_t *flag_bytes)
{
    if (lb < 1 || lc < 1 || !state_bytes || !flag_bytes) return TRIALIGN_EINVAL;
    *state_bytes = (int64_t)fast_sweep_slab_state_bytes(lb, lc);
    *flag_bytes = (int64_t)fast_sweep_slab_flag_bytes(lb, lc);
    return TRIALIGN_OK;
}

int trialign_slab_score_dev(int device, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
                            int lc, int i_begin, int i_end, const void *in_state, const void *in_flags,
                            void *out_state, void *out_flags, int epoch, int score_bits, int32_t *d_out, void *stream)
{
    if (la < 1 || lb < 1 || lc < 1 || score_bits < 2 || score_bits > 32 || !d_a || !d_b || !d_c) return TRIALIGN_EINVAL;
    if (i_begin < 0 || i_end <= i_begin || i_end > la) return TRIALIGN_EINVAL;
    if ((i_begin > 0) != (in_state != nullptr) || (in_state != nullptr) != (in_flags != nullptr)) return TRIALIGN_EINVAL;
    if ((i_end < la) != (out_state != nullptr) || (out_state != nullptr) != (out_flags != nullptr)) return TRIALIGN_EINVAL;
    if (i_end == la && !d_out) return TRIALIGN_EINVAL;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)(i_end - i_begin) * lb * lc;
    ts.last_path = TRIALIGN_PATH_NONE;
    if (!provably_wrap_free(la, lb, lc, score_bits) || !fast_sweep_supports(la, lb, lc)) return TRIALIGN_EINEXACT;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    SlabLink link;
    link.la_total = la;
    link.last = (i_end == la);
    link.epoch = epoch;
    link.in_state = in_state;
    link.in_flags = in_flags;
    link.out_state = out_state;
    link.out_flags = out_flags;
    ts.last_path = TRIALIGN_PATH_FAST_LONG;
    return long_run(ctx, (cudaStream_t)stream, d_a + i_begin, i_end - i_begin, d_b, lb, d_c, lc, d_out,
                               d_out ? d_out + 1 : nullptr, 0, nullptr, &link);
}

int trialign_cols_buffer_bytes(int la, int lb, int64_t *face_bytes, int64_t *prog_bytes)
{
    if (la < 1 || lb < 1 || !face_bytes || !prog_bytes) return TRIALIGN_EINVAL;
    *face_bytes = (int64_t)fast_long2_col_face_bytes(la, lb);
    *prog_bytes = (int64_t)fast_long2_col_prog_bytes(lb);
    return TRIALIGN_OK;
}

int trialign_cols_score_dev(int device, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
                            int lc, int c_begin, int c_end, const void *in_face, const void *in_prog, void *out_face,
                            void *out_prog, int epoch, int score_bits, int32_t *d_out, void *stream)
{
    if (la < 1 || lb < 1 || lc < 1 || score_bits < 2 || score_bits > 32 || !d_a || !d_b || !d_c) return TRIALIGN_EINVAL;
    if (c_begin < 0 || c_end <= c_begin || c_end > lc || (c_begin % 128) != 0) return TRIALIGN_EINVAL;
    if (c_end < lc && (c_end % 128) != 0) return TRIALIGN_EINVAL;  // ranges are whole block-column pairs
    if ((c_begin > 0) != (in_face != nullptr) || (in_face != nullptr) != (in_prog != nullptr)) return TRIALIGN_EINVAL;
    if ((c_end < lc) != (out_face != nullptr) || (out_face != nullptr) != (out_prog != nullptr)) return TRIALIGN_EINVAL;
    if (c_end == lc && !d_out) return TRIALIGN_EINVAL;
    if (epoch < 1 || epoch > 2047) return TRIALIGN_EINVAL;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)la * lb * (c_end - c_begin);
    ts.last_path = TRIALIGN_PATH_NONE;
    if (!provably_wrap_free(la, lb, lc, score_bits) || !fast_sweep_supports(la, lb, lc)) return TRIALIGN_EINEXACT;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ColLink link;
    link.lc_total = lc;
    link.last = (c_end == lc);
    link.epoch = epoch;
    link.in_face = in_face;
    link.in_prog = in_prog;
    link.out_face = out_face;
    link.out_prog = out_prog;
    ts.last_path = TRIALIGN_PATH_FAST_LONG;
    return fast_long2_run(ctx, (cudaStream_t)stream, d_a, la, d_b, lb, d_c + c_begin, c_end - c_begin, d_out,
                          d_out ? d_out + 1 : nullptr, 0, nullptr, nullptr, &link);
}

int trialign_device_alloc(int device, int64_t bytes, void **dptr)
{
    if (bytes <= 0 || !dptr) return TRIALIGN_EINVAL;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    TA_CUDA(cudaMalloc(dptr, (size_t)bytes));
    TA_CUDA(cudaMemset(*dptr, 0, (size_t)bytes));
    return TRIALIGN_OK;
}

int trialign_device_free(int device, void *dptr)
{
    if (!dptr) return TRIALIGN_EINVAL;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    TA_CUDA(cudaFree(dptr));
    return TRIALIGN_OK;
}

int trialign_ipc_export(const void *dptr, unsigned char handle[64])
{
    if (!dptr || !handle) return TRIALIGN_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    TA_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    std::memcpy(handle, &h, 64);
    return TRIALIGN_OK;
}

int trialign_ipc_open(int device, const unsigned char handle[64], void **dptr)
{
    if (!handle || !dptr) return TRIALIGN_EINVAL;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    TA_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TRIALIGN_OK;
}

int trialign_ipc_close(void *dptr)
{
    if (!dptr) return TRIALIGN_EINVAL;
    TA_CUDA(cudaIpcCloseMemHandle(dptr));
    return TRIALIGN_OK;
}

int trialign_synth_batch_dev(int device, uint64_t seed, int64_t first, int64_t n, int len, uint8_t *d_a,
                             uint8_t *d_b, uint8_t *d_c, void *stream)
{
    if (n < 0 || len < 0 || first < 0) return TRIALIGN_EINVAL;
    if (n == 0 || len == 0) return TRIALIGN_OK;
    if (!d_a || !d_b || !d_c) return TRIALIGN_EINVAL;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    return synth_run((cudaStream_t)stream, seed, first, n, len, d_a, d_b, d_c);
}

int trialign_last_timing(double *seconds, double *cells)
{
    ThreadState &ts = tls();
    if (seconds) *seconds = ts.last_seconds;
    if (cells) *cells = ts.last_cells;
    return ts.last_seconds < 0.0 ? TRIALIGN_EINVAL : TRIALIGN_OK;
}

int trialign_last_path(void) { return tls().last_path; }

int64_t trialign_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
