// api.cu -- the C ABI of libtrialign (include/trialign.h): argument checking, per-device contexts, host<->device
// staging, kernel-family dispatch.  No CPU compute path exists here: without a CUDA device the compute entry
// points fail with TRIALIGN_ENODEVICE.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

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

NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

// ---- run-time options (trialign_set_option) ----
// Every experiment switch of the kernels is an explicit option of the library; nothing is read from the environment.
static const char *const kOptionNames[] = {
    "debug",       // 1: launch geometry of the fast kernels on stderr
    "k1",          // K1 variant for A/B runs: 256 | 256m | 512 | 512m | 256l | 256p2 (default) | 256r2 | stream (K2S on len <= 64)
    "k2_lag", "k2_pp", "k2_gatetid", "k2_early", "k2_gate", "k2_rgate", "k2_cols_order",  // K2L experiment parameters (fast_long.cu)
    "k2s",         // K2S variant for A/B runs: "p1" = one plane per step, duties on the edge threads; "dist" = one plane
                   // per step, duties dealt to all threads (default: two planes per step, fast_stream2.cu)
    "multi",       // one long triple on several GPUs: "slab" (i-slab pipeline) | "cols" (column pipeline); default: auto
    "multi_min",   // smallest la (slab) / lc (cols) per GPU for which the multi-GPU pipeline is used (default 64 / 128)
    "kt_seg",      // traceback: anti-diagonal planes per segment of the checkpointed form (default: whole cube if it fits)
    "kt_budget_mb",  // traceback: memory budget for direction words + checkpoints (default: 85 % of the free device memory)
};
static std::mutex g_opt_mu;
static std::map<std::string, std::string> g_opts;

static bool option_known(const char *name)
{
    for (const char *n : kOptionNames)
        if (std::strcmp(n, name) == 0) return true;
    return false;
}

int opt_int(const char *name, int dflt)
{
    std::lock_guard<std::mutex> lk(g_opt_mu);
    auto it = g_opts.find(name);
    if (it == g_opts.end() || it->second.empty()) return dflt;
    char *end = nullptr;
    const long v = std::strtol(it->second.c_str(), &end, 10);
    return (end && *end == 0) ? (int)v : dflt;
}

bool opt_is(const char *name, const char *value)
{
    std::lock_guard<std::mutex> lk(g_opt_mu);
    auto it = g_opts.find(name);
    return it != g_opts.end() && it->second == value;
}

static std::string opt_str(const char *name)
{
    std::lock_guard<std::mutex> lk(g_opt_mu);
    auto it = g_opts.find(name);
    return it == g_opts.end() ? std::string() : it->second;
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
        TA_CUDA(cudaEventCreateWithFlags(&c.scratch_ev, cudaEventDisableTiming));
        TA_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
        c.ready = true;
    }
    *out = &c;
    return TRIALIGN_OK;
}

namespace {

// one long triple: K2L (fast_long.cu)
int long_run(DeviceCtx *ctx, cudaStream_t st, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
             int lc, int32_t *d_score, int32_t *d_states, int check_bits, int *d_viol, const SlabLink *slab = nullptr)
{
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
    const bool want_fast = backend != TRIALIGN_BACKEND_GPU_LITERAL;
    // non-RTL scoring parameters (*_ex entry points): batches of short triples inside score_fast_domain() run the
    // range-checked K1 with the constants as a kernel argument (below); everything else goes to the literal kernel,
    // which evaluates the parameters as given
    const trialign_scoring *scoring = ts.scoring_active ? &ts.scoring : nullptr;
    const bool scored_fast =
        scoring && want_fast && !d_finals && la == lb && lb == lc &&
        ((fast_batch_supports(la) && score_fast_domain(*scoring, score_bits)) ||
         (fast_stream_supports(la) && fast_stream_fits(la, n, ctx->sm_count) && fast_stream_score_ok(*scoring, la, score_bits)));
    if (scoring && !scored_fast) {
        if (backend == TRIALIGN_BACKEND_GPU_FAST) return TRIALIGN_EINEXACT;
        ts.last_path = TRIALIGN_PATH_LITERAL;
        return literal_run(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, d_scores, d_finals, scoring);
    }
    // fast paths are used only where they are provably bit-exact: no candidate can wrap at score_bits
    if (want_fast && !scoring && provably_wrap_free(la, lb, lc, score_bits)) {
        if (!d_finals && la == lb && lb == lc && la <= 64 && fast_stream_supports(la) && opt_is("k1", "stream")) {
            ts.last_path = TRIALIGN_PATH_FAST_STREAM;  // experiment: K2S on a single block per pair (len 61..64)
            return fast_stream_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
        if (!d_finals && la == lb && lb == lc && fast_batch_supports(la)) {
            ts.last_path = TRIALIGN_PATH_FAST_BATCH;
            return fast_batch_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
        if (n == 1 && fast_long_supports(la, lb, lc)) {
            ts.last_path = TRIALIGN_PATH_FAST_LONG;
            return long_run(ctx, st, d_a, la, d_b, lb, d_c, lc, d_scores, d_finals, 0, nullptr);
        }
        if (!d_finals && la == lb && lb == lc && fast_stream_supports(la) && fast_stream_fits(la, n, ctx->sm_count)) {
            ts.last_path = TRIALIGN_PATH_FAST_STREAM;
            return fast_stream_run(ctx, st, d_a, d_b, d_c, n, la, d_scores);
        }
        // batches of very long triples (beyond K2S's 2,048, or whose CTA-private face rings would not fit): every
        // triple is a full-GPU job of its own for K2L, one launch after the other
        if (!d_finals && fast_long_supports(la, lb, lc)) {
            ts.last_path = TRIALIGN_PATH_FAST_LONG;
            for (int64_t t = 0; t < n; ++t) {
                const int rc = long_run(ctx, st, d_a + t * la, la, d_b + t * lb, lb, d_c + t * lc, lc, d_scores + t,
                                        nullptr, 0, nullptr);
                if (rc) return rc;
            }
            return TRIALIGN_OK;
        }
    }
    // One long triple whose wrap-freedom cannot be proved from the lengths: run the fast kernel with its run-time
    // range check, and fall back to the literal kernel only if some value came near the score_bits limits.
    if (want_fast && !scoring && n == 1 && score_bits >= 10 && score_bits <= 16 && fast_long_supports(la, lb, lc)) {
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
    // A batch whose wrap-freedom cannot be proved from the length: the range-checked batch kernels flag, per triple,
    // values that came near the score_bits limits; only the flagged triples are re-scored by the literal kernel.
    if (scored_fast ||
        (want_fast && !d_finals && la == lb && lb == lc && score_bits >= 6 && score_bits <= 16 &&
         (fast_batch_supports(la) || (fast_stream_supports(la) && fast_stream_fits(la, n, ctx->sm_count))))) {
        const int64_t pairs = (n + 1) / 2;
        const size_t flag_bytes = (((size_t)pairs * sizeof(int)) + 255) & ~(size_t)255;
        const size_t list_bytes = (((size_t)n * sizeof(int)) + 255) & ~(size_t)255;
        int rc = ctx->flagged.reserve(256 + flag_bytes + list_bytes);
        if (rc) return rc;
        int *d_count = (int *)ctx->flagged.ptr;
        int *d_flags = (int *)((char *)ctx->flagged.ptr + 256);
        int *d_list = (int *)((char *)ctx->flagged.ptr + 256 + flag_bytes);
        TA_CUDA(cudaMemsetAsync(d_flags, 0, flag_bytes, st));
        const bool k1 = fast_batch_supports(la);
        rc = k1 ? fast_batch_run(ctx, st, d_a, d_b, d_c, n, la, d_scores, score_bits, d_flags, scoring)
                : fast_stream_run(ctx, st, d_a, d_b, d_c, n, la, d_scores, score_bits, d_flags, scoring);
        if (rc) return rc;
        if ((rc = fixup_compact(st, d_flags, n, d_list, d_count))) return rc;
        int count = 0;
        TA_CUDA(cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
        ts.last_flagged += count;
        if (count == 0) {
            ts.last_path = k1 ? TRIALIGN_PATH_FAST_BATCH : TRIALIGN_PATH_FAST_STREAM;
            return TRIALIGN_OK;
        }
        if (backend == TRIALIGN_BACKEND_GPU_FAST) return TRIALIGN_EINEXACT;
        // gather -> literal -> scatter; the staging arrays live behind the planes of the literal kernel's own scratch
        const size_t arr = (((size_t)count * la) + 255) & ~(size_t)255;
        const size_t sc_bytes = (((size_t)count * sizeof(int32_t)) + 255) & ~(size_t)255;
        if ((rc = ctx->misc.reserve(3 * arr + sc_bytes))) return rc;
        uint8_t *g_a = (uint8_t *)ctx->misc.ptr, *g_b = g_a + arr, *g_c = g_b + arr;
        int32_t *g_scores = (int32_t *)(g_c + arr);
        if ((rc = fixup_gather(st, d_a, d_b, d_c, la, d_list, count, g_a, g_b, g_c))) return rc;
        if ((rc = literal_run(ctx, st, g_a, g_b, g_c, count, la, la, la, score_bits, g_scores, nullptr, scoring))) return rc;
        if ((rc = fixup_scatter(st, g_scores, d_list, count, d_scores))) return rc;
        ts.last_path = TRIALIGN_PATH_FAST_THEN_LITERAL;
        return TRIALIGN_OK;
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
// packed: a/b/c are batches in the testbench's 128-bit RAM format ([n][ceil(len/32)][16] bytes, la = lb = lc = len); they
// are uploaded as they are and expanded on the device.
int run_host_batch_on_device(int device, const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int la,
                             int lb, int lc, int score_bits, int backend, int32_t *scores, int32_t *finals,
                             double *seconds, bool packed = false)
{
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    const size_t na = (size_t)n * la, nb = (size_t)n * lb, nc = (size_t)n * lc;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t ram = packed ? up((size_t)n * ((la + 31) / 32) * 16) : 0;  // one packed array
    if ((rc = ctx->seq.reserve(up(na) + up(nb) + up(nc) + 3 * ram + 256))) return rc;
    const size_t n_pad = ((size_t)n + 63) & ~(size_t)63;  // finals start on a 256-byte boundary
    if ((rc = ctx->out.reserve(n_pad * sizeof(int32_t) + (finals ? (size_t)n * 7 * sizeof(int32_t) : 0)))) return rc;
    uint8_t *d_a = (uint8_t *)ctx->seq.ptr, *d_b = d_a + up(na), *d_c = d_b + up(nb);
    int32_t *d_scores = (int32_t *)ctx->out.ptr;
    int32_t *d_finals = finals ? d_scores + n_pad : nullptr;
    ScratchLease lease(ctx, st);
    {
        NvtxRange r("trialign:h2d");
        if (packed) {
            uint8_t *r_a = d_c + up(nc), *r_b = r_a + ram, *r_c = r_b + ram;
            const size_t bytes = (size_t)n * ((la + 31) / 32) * 16;
            TA_CUDA(cudaMemcpyAsync(r_a, a, bytes, cudaMemcpyHostToDevice, st));
            TA_CUDA(cudaMemcpyAsync(r_b, b, bytes, cudaMemcpyHostToDevice, st));
            TA_CUDA(cudaMemcpyAsync(r_c, c, bytes, cudaMemcpyHostToDevice, st));
            if ((rc = unpack_ram128_run(st, r_a, n, la, d_a)) || (rc = unpack_ram128_run(st, r_b, n, la, d_b)) ||
                (rc = unpack_ram128_run(st, r_c, n, la, d_c)))
                return rc;
        } else {
            TA_CUDA(cudaMemcpyAsync(d_a, a, na, cudaMemcpyHostToDevice, st));
            TA_CUDA(cudaMemcpyAsync(d_b, b, nb, cudaMemcpyHostToDevice, st));
            TA_CUDA(cudaMemcpyAsync(d_c, c, nc, cudaMemcpyHostToDevice, st));
        }
    }
    TA_CUDA(cudaEventRecord(ctx->ev0, st));
    {
        NvtxRange r("trialign:kernels");
        rc = dispatch_batch(ctx, st, d_a, d_b, d_c, n, la, lb, lc, score_bits, backend, d_scores, d_finals);
    }
    if (rc) {
        cudaStreamSynchronize(st);
        return rc;
    }
    TA_CUDA(cudaEventRecord(ctx->ev1, st));
    {
        NvtxRange r("trialign:d2h");
        TA_CUDA(cudaMemcpyAsync(scores, d_scores, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (finals) TA_CUDA(cudaMemcpyAsync(finals, d_finals, (size_t)n * 7 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        TA_CUDA(cudaStreamSynchronize(st));
    }
    float ms = 0.f;
    TA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (seconds) *seconds = (double)ms * 1e-3;
    return TRIALIGN_OK;
}

std::vector<int> devices_of(unsigned mask)
{
    std::vector<int> devs;
    for (int d = 0; d < kMaxDevices; ++d)
        if (mask & (1u << d)) devs.push_back(d);
    if (devs.empty()) devs.push_back(0);
    return devs;
}

// How one long triple is spread over G GPUs: 0 = not at all (first GPU alone), 1 = i-slab pipeline, 2 = column pipeline.
// Measured on 8192^3 (W=16): 2 GPUs slab 255 ms / cols 269, 4 GPUs 143 / 140, 8 GPUs 81.3 / 78.3 -- the slab pipeline
// pays (G-1) task durations of fill plus the set-up of G times as many tasks, the column pipeline a lag of one block row
// per GPU plus coarser waves; columns win from 4 GPUs on when every GPU gets at least two block-column pairs.
bool long_launch_supports(int planes, int la_total, int lb, int lc);
int multi_plan(int la, int lb, int lc, int score_bits, int G)
{
    if (G < 2 || !provably_wrap_free(la, lb, lc, score_bits)) return 0;
    const bool cols_ok = lc >= opt_int("multi_min", 128) * G && lc >= 128 * G && long_launch_supports(la, la, lb, lc);
    const bool slab_ok = la >= opt_int("multi_min", 64) * G && la >= G && long_launch_supports((la + G - 1) / G, la, lb, lc);
    if (opt_is("multi", "cols")) return cols_ok ? 2 : 0;
    if (opt_is("multi", "slab")) return slab_ok ? 1 : 0;
    if (cols_ok && G >= 4 && lc >= 256 * G && la >= 2048) return 2;
    return slab_ok ? 1 : (cols_ok ? 2 : 0);
}

// Limits of one K2L launch over `planes` planes of a triple whose whole extent is (la_total, lb, lc): the launch keeps
// its own planes' a symbols in shared memory, the 16-bit potential range depends on the whole triple.
bool long_launch_supports(int planes, int la_total, int lb, int lc)
{
    return planes >= 1 && planes <= la_total && fast_long_supports(planes, lb, lc) &&
           fast_long_range_ok(planes, la_total, lb, lc);
}

// ---- ONE long triple on several GPUs from ONE process (the drop-in form of SURVEY.md 8(e)(2)) ----
// The calling thread drives every GPU of the mask: peer access g -> g+1, one stream per device, the hand-over buffers of
// the i-slab pipeline (or the boundary faces of the column pipeline) live on the receiving GPU and are written by the
// sending GPU's kernel through NVLink.  Kernels are launched upstream first; each spins only on flags its upstream
// neighbour -- already launched -- raises, so the chain always drains.  All streams are synchronised before returning,
// which is also what keeps a run from overwriting hand-over slots its predecessor's consumers have not read.
struct MultiPipe {
    std::vector<int> devs;
    std::vector<void *> state, flags;   // [g] buffers on device g (g >= 1), written by g-1
    std::vector<size_t> state_bytes, flag_bytes;
    int epoch = 0;
    bool peers_ready = false;
    int la = 0, lb = 0, lc = 0;         // shape of the last run: another shape maps tasks to other flag words
};
std::mutex g_multi_mu;
std::map<unsigned, std::unique_ptr<MultiPipe>> g_multi;

int multi_reserve(MultiPipe &mp, size_t g, size_t sbytes, size_t fbytes)
{
    if (mp.state_bytes[g] >= sbytes && mp.flag_bytes[g] >= fbytes) return TRIALIGN_OK;
    TA_CUDA(cudaSetDevice(mp.devs[g]));
    TA_CUDA(cudaDeviceSynchronize());
    if (mp.state[g]) cudaFree(mp.state[g]);
    if (mp.flags[g]) cudaFree(mp.flags[g]);
    mp.state[g] = mp.flags[g] = nullptr;
    mp.state_bytes[g] = mp.flag_bytes[g] = 0;
    TA_CUDA(cudaMalloc(&mp.state[g], sbytes));
    TA_CUDA(cudaMalloc(&mp.flags[g], fbytes));
    TA_CUDA(cudaMemset(mp.flags[g], 0, fbytes));
    mp.state_bytes[g] = sbytes;
    mp.flag_bytes[g] = fbytes;
    return TRIALIGN_OK;
}

int run_host_triple_multi(const std::vector<int> &devs, unsigned mask, bool cols, const uint8_t *a, int la,
                          const uint8_t *b, int lb, const uint8_t *c, int lc, int32_t *score, int32_t fin[7],
                          double *seconds)
{
    DeviceGuard guard;
    const size_t G = devs.size();
    std::lock_guard<std::mutex> mlk(g_multi_mu);
    std::unique_ptr<MultiPipe> &slot = g_multi[mask | (cols ? 0x80000000u : 0u)];
    if (!slot) {
        slot.reset(new MultiPipe);
        slot->devs = devs;
        slot->state.assign(G, nullptr);
        slot->flags.assign(G, nullptr);
        slot->state_bytes.assign(G, 0);
        slot->flag_bytes.assign(G, 0);
    }
    MultiPipe &mp = *slot;
    std::vector<DeviceCtx *> ctx(G, nullptr);
    int rc;
    for (size_t g = 0; g < G; ++g)
        if ((rc = get_ctx(devs[g], &ctx[g]))) return rc;
    std::vector<std::unique_lock<std::mutex>> locks;
    for (size_t g = 0; g < G; ++g) locks.emplace_back(ctx[g]->mu);  // ascending device order: no lock cycle
    if (!mp.peers_ready) {
        for (size_t g = 0; g + 1 < G; ++g) {
            int ok = 0;
            TA_CUDA(cudaDeviceCanAccessPeer(&ok, devs[g], devs[g + 1]));
            if (!ok) return TRIALIGN_EUNSUPPORTED;
            TA_CUDA(cudaSetDevice(devs[g]));
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[g + 1], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        }
        mp.peers_ready = true;
    }
    size_t sbytes, fbytes;
    if (cols) {
        sbytes = fast_long2_col_face_bytes(la, lb);
        fbytes = std::max<size_t>(fast_long2_col_prog_bytes(lb), 256);
    } else {
        sbytes = fast_long_slab_state_bytes(lb, lc);
        fbytes = fast_long_slab_flag_bytes(lb, lc);
    }
    for (size_t g = 1; g < G; ++g)
        if ((rc = multi_reserve(mp, g, sbytes, fbytes))) return rc;
    if (la != mp.la || lb != mp.lb || lc != mp.lc) {
        // Flags are compared with the epoch and never cleared; with another shape a word may still hold what a run
        // 2,047 epochs ago left there.  Nothing is in flight (every run ends synchronised): clear them.
        for (size_t g = 1; g < G; ++g) {
            TA_CUDA(cudaSetDevice(devs[g]));
            TA_CUDA(cudaMemset(mp.flags[g], 0, mp.flag_bytes[g]));
        }
        mp.la = la; mp.lb = lb; mp.lc = lc;
    }
    mp.epoch = mp.epoch % 2047 + 1;  // 1..2047: the column pipeline tags its progress words with 11 bits
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const int pairs = (lc + 127) / 128;
    std::vector<int32_t *> d_out(G, nullptr);
    for (size_t g = 0; g < G; ++g) {
        TA_CUDA(cudaSetDevice(devs[g]));
        cudaStream_t st = ctx[g]->stream;
        if ((rc = ctx[g]->seq.reserve(up(la) + up(lb) + up(lc) + 256))) return rc;
        if ((rc = ctx[g]->out.reserve(256))) return rc;
        uint8_t *d_a = (uint8_t *)ctx[g]->seq.ptr, *d_b = d_a + up(la), *d_c = d_b + up(lb);
        d_out[g] = (int32_t *)ctx[g]->out.ptr;
        ScratchLease lease(ctx[g], st);
        {
            NvtxRange r("trialign:h2d");
            TA_CUDA(cudaMemcpyAsync(d_a, a, la, cudaMemcpyHostToDevice, st));
            TA_CUDA(cudaMemcpyAsync(d_b, b, lb, cudaMemcpyHostToDevice, st));
            TA_CUDA(cudaMemcpyAsync(d_c, c, lc, cudaMemcpyHostToDevice, st));
        }
        TA_CUDA(cudaEventRecord(ctx[g]->ev0, st));
        NvtxRange r("trialign:kernels");
        const bool last = g + 1 == G;
        if (cols) {
            const int c0 = 128 * (int)((size_t)pairs * g / G), c1 = last ? lc : 128 * (int)((size_t)pairs * (g + 1) / G);
            ColLink link;
            link.lc_total = lc;
            link.last = last;
            link.epoch = mp.epoch;
            link.in_face = g ? mp.state[g] : nullptr;
            link.in_prog = g ? mp.flags[g] : nullptr;
            link.out_face = last ? nullptr : mp.state[g + 1];
            link.out_prog = last ? nullptr : mp.flags[g + 1];
            rc = fast_long2_run(ctx[g], st, d_a, la, d_b, lb, d_c + c0, c1 - c0, d_out[g], d_out[g] + 1, 0, nullptr,
                                nullptr, &link);
        } else {
            const int i0 = (int)((long long)la * g / G), i1 = (int)((long long)la * (g + 1) / G);
            SlabLink link;
            link.la_total = la;
            link.last = last;
            link.epoch = mp.epoch;
            link.in_state = g ? mp.state[g] : nullptr;
            link.in_flags = g ? mp.flags[g] : nullptr;
            link.out_state = last ? nullptr : mp.state[g + 1];
            link.out_flags = last ? nullptr : mp.flags[g + 1];
            rc = fast_long2_run(ctx[g], st, d_a + i0, i1 - i0, d_b, lb, d_c, lc, d_out[g], d_out[g] + 1, 0, nullptr,
                                &link);
        }
        if (rc) break;
        TA_CUDA(cudaEventRecord(ctx[g]->ev1, st));
    }
    if (rc) {  // drain whatever was launched (upstream kernels finish on their own)
        for (size_t g = 0; g < G; ++g) {
            cudaSetDevice(devs[g]);
            cudaStreamSynchronize(ctx[g]->stream);
        }
        return rc;
    }
    int32_t res[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    {
        NvtxRange r("trialign:d2h");
        TA_CUDA(cudaSetDevice(devs[G - 1]));
        TA_CUDA(cudaMemcpyAsync(res, d_out[G - 1], sizeof res, cudaMemcpyDeviceToHost, ctx[G - 1]->stream));
        for (size_t g = 0; g < G; ++g) {
            TA_CUDA(cudaSetDevice(devs[g]));
            TA_CUDA(cudaStreamSynchronize(ctx[g]->stream));
        }
    }
    double worst = 0.0;
    for (size_t g = 0; g < G; ++g) {
        float ms = 0.f;
        TA_CUDA(cudaSetDevice(devs[g]));
        TA_CUDA(cudaEventElapsedTime(&ms, ctx[g]->ev0, ctx[g]->ev1));
        worst = std::max(worst, (double)ms * 1e-3);
    }
    *score = res[0];
    std::memcpy(fin, res + 1, 7 * sizeof(int32_t));
    if (seconds) *seconds = worst;
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
    case TRIALIGN_EUNSUPPORTED: return "not supported: a backend outside libtrialign (CPU checkers live under oracle/), or a size beyond a kernel's limits";
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

int trialign_set_option(const char *name, const char *value)
{
    if (!name || !option_known(name)) return TRIALIGN_EINVAL;
    std::lock_guard<std::mutex> lk(g_opt_mu);
    if (!value || !value[0]) g_opts.erase(name);
    else g_opts[name] = value;
    return TRIALIGN_OK;
}

int trialign_get_option(const char *name, char *buf, int cap)
{
    if (!name || !option_known(name) || !buf || cap < 1) return TRIALIGN_EINVAL;
    const std::string v = opt_str(name);
    if ((int)v.size() + 1 > cap) return TRIALIGN_ECAPACITY;
    std::memcpy(buf, v.c_str(), v.size() + 1);
    return TRIALIGN_OK;
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
    ts.last_flagged = 0;
    ts.last_devices = 1;
    if (n == 0) return TRIALIGN_OK;
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    if ((rc = get_ctx(device, &ctx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (len == 0) {
        TA_CUDA(cudaMemsetAsync(d_scores, 0, (size_t)n * sizeof(int32_t), st));
        return TRIALIGN_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScratchLease lease(ctx, st);
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
    ts.last_flagged = 0;
    ts.last_devices = 1;
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    if ((rc = get_ctx(device, &ctx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (la == 0 || lb == 0 || lc == 0) {
        TA_CUDA(cudaMemsetAsync(d_out, 0, 8 * sizeof(int32_t), st));
        return TRIALIGN_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScratchLease lease(ctx, st);
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
    ts.last_flagged = 0;
    if (trialign_device_count() == 0) return TRIALIGN_ENODEVICE;
    if (la == 0 || lb == 0 || lc == 0) {
        *score = 0;
        if (final_states) std::memset(final_states, 0, 7 * sizeof(int32_t));
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    int32_t fin[7];
    double sec = -1.0;
    const std::vector<int> devs = devices_of(devices);
    const int have = trialign_device_count();
    for (int d : devs)
        if (d >= have) return TRIALIGN_EINVAL;
    ts.last_devices = 1;
    // several GPUs in the mask: the triple is split as an i-slab (or column) pipeline when it is long enough for every
    // GPU to get whole blocks and the fast kernels are provably bit-exact for it; otherwise the first GPU serves it alone
    // (trialign_last_devices() tells which happened)
    if (devs.size() > 1 && backend != TRIALIGN_BACKEND_GPU_LITERAL && !ts.scoring_active) {
        const int G = (int)devs.size();
        const int plan = multi_plan(la, lb, lc, score_bits, G);
        if (plan) {
            rc = run_host_triple_multi(devs, devices, plan == 2, a, la, b, lb, c, lc, score, fin, &sec);
            if (rc != TRIALIGN_EUNSUPPORTED) {  // EUNSUPPORTED: no peer access between neighbours -> one GPU below
                if (rc) return rc;
                if (final_states) std::memcpy(final_states, fin, sizeof fin);
                ts.last_seconds = sec;
                ts.last_path = TRIALIGN_PATH_FAST_LONG;
                ts.last_devices = G;
                return TRIALIGN_OK;
            }
        }
    }
    rc = run_host_batch_on_device(devs[0], a, b, c, 1, la, lb, lc, score_bits, backend, score, fin, &sec);
    if (rc) return rc;
    if (final_states) std::memcpy(final_states, fin, sizeof fin);
    ts.last_seconds = sec;
    return TRIALIGN_OK;
}

int64_t trialign_triple_count(int nseq) { return nseq < 3 ? 0 : (int64_t)all_triples_count(nseq); }

namespace {
// triples [t_begin, t_end) of the set on one device: upload the set once, gather chunk by chunk, score, copy back
int all_triples_on_device(int device, const uint8_t *seqs, int nseq, int len, int score_bits, int backend,
                          int64_t t_begin, int64_t t_end, int32_t *scores, double *seconds)
{
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    int rc;
    if ((rc = get_ctx(device, &ctx))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    const int64_t total = t_end - t_begin;
    // chunks of at most 2^20 triples bound the staging arrays (3 x chunk x len bytes) whatever nseq is
    const int64_t chunk = total < (1 << 20) ? total : (1 << 20);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t set_bytes = up((size_t)nseq * len), arr = up((size_t)chunk * len);
    if ((rc = ctx->seq.reserve(set_bytes + 3 * arr + 256))) return rc;
    if ((rc = ctx->out.reserve((size_t)chunk * sizeof(int32_t)))) return rc;
    uint8_t *d_set = (uint8_t *)ctx->seq.ptr, *d_a = d_set + set_bytes, *d_b = d_a + arr, *d_c = d_b + arr;
    int32_t *d_scores = (int32_t *)ctx->out.ptr;
    ScratchLease lease(ctx, st);
    {
        NvtxRange r("trialign:h2d");
        TA_CUDA(cudaMemcpyAsync(d_set, seqs, (size_t)nseq * len, cudaMemcpyHostToDevice, st));
    }
    TA_CUDA(cudaEventRecord(ctx->ev0, st));
    for (int64_t t0 = t_begin; t0 < t_end; t0 += chunk) {
        const int64_t cnt = t_end - t0 < chunk ? t_end - t0 : chunk;
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
    if (seconds) *seconds = (double)ms * 1e-3;
    return TRIALIGN_OK;
}
}  // namespace

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
    ts.last_flagged = 0;
    ts.last_devices = 1;
    const int have = trialign_device_count();
    if (have == 0) return TRIALIGN_ENODEVICE;
    if (total == 0) return TRIALIGN_OK;
    if (len == 0) {
        std::memset(scores, 0, (size_t)total * sizeof(int32_t));
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    std::vector<int> devs = devices_of(devices);
    for (int d : devs)
        if (d >= have) return TRIALIGN_EINVAL;
    if ((int64_t)devs.size() > total) devs.resize((size_t)total);
    if (devs.size() == 1) {
        double sec = -1.0;
        rc = all_triples_on_device(devs[0], seqs, nseq, len, score_bits, backend, 0, total, scores, &sec);
        ts.last_seconds = sec;
        return rc;
    }
    // contiguous ranges of the triple list, one host thread per device, no exchange
    const size_t G = devs.size();
    std::vector<int> rcs(G, 0), paths(G, 0);
    std::vector<double> secs(G, -1.0);
    std::vector<std::string> errs(G);
    std::vector<int64_t> flagged(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
        const int64_t t0 = total * (int64_t)g / (int64_t)G, t1 = total * (int64_t)(g + 1) / (int64_t)G;
        th.emplace_back([&, g, t0, t1]() {
            rcs[g] = all_triples_on_device(devs[g], seqs, nseq, len, score_bits, backend, t0, t1, scores, &secs[g]);
            paths[g] = tls().last_path;
            errs[g] = tls().cuda_error;
            flagged[g] = tls().last_flagged;
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
        ts.last_flagged += flagged[g];
    }
    ts.last_seconds = worst;
    ts.last_devices = (int)G;
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
    ts.last_flagged = 0;
    if (trialign_device_count() == 0) return TRIALIGN_ENODEVICE;
    *n_moves = 0;
    start[0] = la; start[1] = lb; start[2] = lc;
    if (la == 0 || lb == 0 || lc == 0) {  // already on a border plane: empty path, score 0
        *score = 0;
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    // the traceback runs on ONE GPU, the first of the mask (the direction cube is not partitioned)
    if (first_device(devices) >= trialign_device_count()) return TRIALIGN_EINVAL;
    ts.last_devices = 1;
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(first_device(devices), &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    ScratchLease lease(ctx, st);
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
    ts.last_segments = 0;
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

static int score_batch_impl(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len, int score_bits,
                            int backend, unsigned devices, int32_t *scores, bool packed)
{
    const size_t stride = packed ? (size_t)((len + 31) / 32) * 16 : (size_t)len;  // host bytes per sequence
    if (n < 0 || len < 0 || score_bits < 2 || score_bits > 32) return TRIALIGN_EINVAL;
    if (n > 0 && !scores) return TRIALIGN_EINVAL;
    if (n > 0 && len > 0 && (!a || !b || !c)) return TRIALIGN_EINVAL;
    int rc = check_backend(backend);
    if (rc) return rc;
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)n * len * len * len;
    ts.last_path = TRIALIGN_PATH_NONE;
    ts.last_flagged = 0;
    const int have = trialign_device_count();
    if (have == 0) return TRIALIGN_ENODEVICE;
    if (n == 0) return TRIALIGN_OK;
    if (len == 0) {
        std::memset(scores, 0, (size_t)n * sizeof(int32_t));
        ts.last_seconds = 0.0;
        return TRIALIGN_OK;
    }
    std::vector<int> devs = devices_of(devices);
    for (int d : devs)
        if (d >= have) return TRIALIGN_EINVAL;
    if ((int64_t)devs.size() > n) devs.resize((size_t)n);
    ts.last_devices = (int)devs.size();
    if (devs.size() == 1) {
        double sec = -1.0;
        rc = run_host_batch_on_device(devs[0], a, b, c, n, len, len, len, score_bits, backend, scores, nullptr, &sec,
                                      packed);
        ts.last_seconds = sec;
        return rc;
    }
    // contiguous split across devices, one host thread per device, no exchange
    const size_t G = devs.size();
    std::vector<int> rcs(G, 0), paths(G, 0);
    std::vector<double> secs(G, -1.0);
    std::vector<std::string> errs(G);
    std::vector<int64_t> flagged(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
        const int64_t t0 = n * (int64_t)g / (int64_t)G, t1 = n * (int64_t)(g + 1) / (int64_t)G;
        th.emplace_back([&, g, t0, t1]() {
            tls().scoring_active = ts.scoring_active;  // the *_ex scoring of the calling thread
            tls().scoring = ts.scoring;
            if (t1 > t0)
                rcs[g] = run_host_batch_on_device(devs[g], a + (size_t)t0 * stride, b + (size_t)t0 * stride,
                                                  c + (size_t)t0 * stride, t1 - t0, len, len, len, score_bits,
                                                  backend, scores + t0, nullptr, &secs[g], packed);
            paths[g] = tls().last_path;
            errs[g] = tls().cuda_error;
            flagged[g] = tls().last_flagged;
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
        ts.last_flagged += flagged[g];
    }
    ts.last_seconds = worst;
    return TRIALIGN_OK;
}

int trialign_score_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len, int score_bits,
                         int backend, unsigned devices, int32_t *scores)
{
    return score_batch_impl(a, b, c, n, len, score_bits, backend, devices, scores, false);
}

// ---- the testbench's packed sequence RAMs as a direct input (src/TriAlign_tb.sv:94-96,150-169) ----
int trialign_score_batch_ram128(const uint8_t *ram_a, const uint8_t *ram_b, const uint8_t *ram_c, int64_t n, int len,
                                int score_bits, int backend, unsigned devices, int32_t *scores)
{
    return score_batch_impl(ram_a, ram_b, ram_c, n, len, score_bits, backend, devices, scores, true);
}

int trialign_score_ram128(const uint8_t *ram_a, int la, const uint8_t *ram_b, int lb, const uint8_t *ram_c, int lc,
                          int score_bits, int backend, unsigned devices, int32_t *score, int32_t final_states[7])
{
    if (la < 0 || lb < 0 || lc < 0) return TRIALIGN_EINVAL;
    if ((la && !ram_a) || (lb && !ram_b) || (lc && !ram_c)) return TRIALIGN_EINVAL;
    // one triple: the three sequences are expanded on the host (a few KB) and take the ordinary path, multi-GPU included
    std::vector<uint8_t> a((size_t)la), b((size_t)lb), c((size_t)lc);
    int rc;
    if ((rc = trialign_unpack_seq_ram128(ram_a, la, a.data(), la)) || (rc = trialign_unpack_seq_ram128(ram_b, lb, b.data(), lb)) ||
        (rc = trialign_unpack_seq_ram128(ram_c, lc, c.data(), lc)))
        return rc;
    return trialign_score(a.data(), la, b.data(), lb, c.data(), lc, score_bits, backend, devices, score, final_states);
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

int trialign_multi_plan(int la, int lb, int lc, int score_bits, int n_devices)
{
    if (la < 0 || lb < 0 || lc < 0 || score_bits < 2 || score_bits > 32 || n_devices < 1) return TRIALIGN_EINVAL;
    if (la == 0 || lb == 0 || lc == 0) return 0;
    return multi_plan(la, lb, lc, score_bits, n_devices);
}

int trialign_slab_buffer_bytes(int lb, int lc, int64_t *state_bytes, int64_t *flag_bytes)
{
    if (lb < 1 || lc < 1 || !state_bytes || !flag_bytes) return TRIALIGN_EINVAL;
    *state_bytes = (int64_t)fast_long_slab_state_bytes(lb, lc);
    *flag_bytes = (int64_t)fast_long_slab_flag_bytes(lb, lc);
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
    if (epoch < 1) return TRIALIGN_EINVAL;  // 0 is what freshly allocated flags hold
    ThreadState &ts = tls();
    ts.last_seconds = -1.0;
    ts.last_cells = (double)(i_end - i_begin) * lb * lc;
    ts.last_path = TRIALIGN_PATH_NONE;
    ts.last_flagged = 0;
    ts.last_devices = 1;
    if (!provably_wrap_free(la, lb, lc, score_bits)) return TRIALIGN_EINEXACT;
    // size limits apply to the slab's own planes (their a symbols live in shared memory) and to the 16-bit range of the
    // whole triple -- not to la, so the pipeline takes triples longer than one launch could
    if (!long_launch_supports(i_end - i_begin, la, lb, lc)) return TRIALIGN_EUNSUPPORTED;
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScratchLease lease(ctx, (cudaStream_t)stream);
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
    ts.last_flagged = 0;
    ts.last_devices = 1;
    if (!provably_wrap_free(la, lb, lc, score_bits)) return TRIALIGN_EINEXACT;
    if (!long_launch_supports(la, la, lb, lc)) return TRIALIGN_EUNSUPPORTED;
    DeviceGuard guard;
    DeviceCtx *ctx = nullptr;
    int rc = get_ctx(device, &ctx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ScratchLease lease(ctx, (cudaStream_t)stream);
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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
    DeviceGuard guard;
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

int trialign_last_devices(void) { return tls().last_devices; }

int64_t trialign_last_flagged(void) { return tls().last_flagged; }

int trialign_last_segments(void) { return tls().last_segments; }

int64_t trialign_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
