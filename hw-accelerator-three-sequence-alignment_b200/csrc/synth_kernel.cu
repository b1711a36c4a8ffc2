// synth_kernel.cu -- the synthetic ACGT generator of SURVEY.md 8(d) on the device, so that benchmark inputs can be
// created resident in HBM.  Bit-identical to trialign_synth() in host_io.cpp.
#include "kernels.cuh"

namespace trialign {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_kernel(uint64_t seed, int64_t first, int64_t n, int len, uint8_t *__restrict__ a,
                             uint8_t *__restrict__ b, uint8_t *__restrict__ c)
{
    const int64_t total = n * len;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / len;
        const uint64_t p = (uint64_t)(e - t * len);
        const uint64_t base = (uint64_t)(3 * (first + t));
        a[e] = (uint8_t)(mix64(seed ^ ((base + 0) << 32) ^ p) >> 62);
        b[e] = (uint8_t)(mix64(seed ^ ((base + 1) << 32) ^ p) >> 62);
        c[e] = (uint8_t)(mix64(seed ^ ((base + 2) << 32) ^ p) >> 62);
    }
}

}  // namespace

int synth_run(cudaStream_t st, uint64_t seed, int64_t first, int64_t n, int len, uint8_t *d_a, uint8_t *d_b,
              uint8_t *d_c)
{
    const int64_t total = n * len;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    synth_kernel<<<blocks, 256, 0, st>>>(seed, first, n, len, d_a, d_b, d_c);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
