// batch_fixup.cu -- plumbing of the range-checked batch path: the fast kernels (K1 / K2S CHECK variants) flag the
// triples whose values came too close to the score_bits limits for wrap-freedom to be certain; those -- and only those
// -- are gathered into contiguous arrays, re-scored by the literal kernel (exact for any input) and scattered back.
#include "kernels.cuh"

namespace trialign {

namespace {

// flags[pair]: bit 0 / 1 = low / high triple of the pair.  list[] receives the flagged triple indices (< n) in no
// particular order, *count their number.
__global__ void compact_kernel(const int *flags, long long n, int *list, int *count)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if ((flags[t >> 1] >> (t & 1)) & 1) list[atomicAdd(count, 1)] = (int)t;
}

__global__ void gather_kernel(const uint8_t *a, const uint8_t *b, const uint8_t *c, int len, const int *list, int count,
                              uint8_t *ga, uint8_t *gb, uint8_t *gc)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)count * len) return;
    const int q = (int)(e / len), pos = (int)(e % len);
    const long long src = (long long)list[q] * len + pos;
    ga[e] = a[src];
    gb[e] = b[src];
    gc[e] = c[src];
}

__global__ void scatter_kernel(const int32_t *g_scores, const int *list, int count, int32_t *scores)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < count) scores[list[q]] = g_scores[q];
}

}  // namespace

int fixup_compact(cudaStream_t st, const int *d_pair_flags, int64_t n, int *d_list, int *d_count)
{
    TA_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
    const unsigned grid = (unsigned)((n + 255) / 256);
    compact_kernel<<<grid, 256, 0, st>>>(d_pair_flags, (long long)n, d_list, d_count);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

int fixup_gather(cudaStream_t st, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c, int len, const int *d_list,
                 int count, uint8_t *g_a, uint8_t *g_b, uint8_t *g_c)
{
    const long long total = (long long)count * len;
    gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_a, d_b, d_c, len, d_list, count, g_a, g_b, g_c);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

int fixup_scatter(cudaStream_t st, const int32_t *g_scores, const int *d_list, int count, int32_t *d_scores)
{
    scatter_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(g_scores, d_list, count, d_scores);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
