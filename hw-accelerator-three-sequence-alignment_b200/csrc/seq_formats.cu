// seq_formats.cu -- device side of the testbench's packed sequence format (src/TriAlign_tb.sv:94-96,150-169): 128-bit
// RAM words of 32 four-bit symbols, symbol n in word n>>5 at bits [4*(n&31)+3 : 4*(n&31)].  A batch in this format is
// uploaded as it is (half the bytes of the one-symbol-per-byte layout) and expanded here into the [n][len] byte arrays
// the kernels read; the PE's truncation to bits [1:0] (src/PE_1cyc.v:63-66) is applied by the kernels themselves.
#include "kernels.cuh"

namespace trialign {

namespace {

// one thread per 16 symbols: an 8-byte load, sixteen nibbles out
__global__ void unpack_ram128_kernel(const uint8_t *__restrict__ ram, long long n, int len, int words, uint8_t *__restrict__ out)
{
    const int chunks = (len + 15) / 16;  // per sequence
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * chunks) return;
    const long long t = e / chunks;
    const int q = (int)(e % chunks);
    const unsigned long long v = *reinterpret_cast<const unsigned long long *>(ram + (t * words) * 16 + q * 8);
    uint8_t *dst = out + t * len + q * 16;
    const int m = len - q * 16 < 16 ? len - q * 16 : 16;
    for (int x = 0; x < m; ++x) dst[x] = (uint8_t)((v >> (4 * x)) & 15);
}

}  // namespace

int unpack_ram128_run(cudaStream_t st, const uint8_t *d_ram, int64_t n, int len, uint8_t *d_out)
{
    if (n <= 0 || len <= 0) return TRIALIGN_OK;
    const int words = (len + 31) / 32, chunks = (len + 15) / 16;
    const long long total = (long long)n * chunks;
    unpack_ram128_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_ram, (long long)n, len, words, d_out);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
