// alu_microbench.cu -- measures the integer-ALU peaks the 3D-DP kernels are bounded by on B200 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/alu_microbench tools/alu_microbench.cu
// Each kernel runs ILP independent dependent-chains per thread of one instruction kind and reports
// warp-instructions per clock per SM (x32 = lanes/clk/SM) from clock64() deltas, plus the SM clock derived
// from cudaEvent time.  Output: one JSON object per line.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 4096;

enum Op { OP_VMAX_S16X2, OP_VMAX_U16X2, OP_VMAX3_U16X2, OP_VADDMAX_U16X2, OP_VADD_16X2, OP_IADD3, OP_IMAD, OP_LOP3,
          OP_PRMT, OP_VMAX3_S32, OP_MIX_VMAX3_IADD3, OP_MIX_VMAX3_IMAD, OP_MIX_VADDMAX_VMAX3, OP_SHFL, OP_LDS32,
          OP_LDS128, OP_MIX_VMAX3_LDS128, OP_MIX_VMAX3_PRMT, OP_COUNT };

static const char* op_name[OP_COUNT] = { "vimnmx.s16x2", "vimnmx.u16x2", "vimnmx3.u16x2", "viaddmnmx.u16x2", "viadd.16x2",
  "iadd3", "imad", "lop3", "prmt", "vimnmx3.s32", "mix vimnmx3.u16x2+iadd3 (1:1)", "mix vimnmx3.u16x2+imad (1:1)",
  "mix viaddmnmx+vimnmx3 (1:1)", "shfl.up", "lds.32", "lds.128", "mix vimnmx3.u16x2+lds.128 (4:1)", "mix vimnmx3.u16x2+prmt (1:1)" };
// instructions issued per inner step per chain
static const int op_instr[OP_COUNT] = {1,1,1,1,1,1,1,1,1,1,2,2,2,1,1,1,5,2};

__device__ __forceinline__ uint4 lds128(const uint4* base, unsigned idx)
{
    uint4 v;
    unsigned addr = (unsigned)__cvta_generic_to_shared(base + idx);
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds32(const uint4* base, unsigned idx)
{
    unsigned v;
    unsigned addr = (unsigned)__cvta_generic_to_shared((const unsigned*)base + idx);
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(unsigned* out, unsigned y0, unsigned z0, long long* cyc)
{
    __shared__ uint4 sm[1024];
    unsigned x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 2654435761u + i; y[i] = y0 + i * 0x00010001u; z[i] = z0 ^ (i * 0x01010101u); }
    sm[threadIdx.x] = make_uint4(x[0], x[1], x[2], x[3]);
    __syncthreads();
    const unsigned lane_off = (threadIdx.x & 1023);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == OP_VMAX_S16X2) x[i] = __vmaxs2(x[i], y[i]) ^ 0; 
                if (OP == OP_VMAX_S16X2) { unsigned t = y[i]; y[i] = x[i]; x[i] = t; }
                if (OP == OP_VMAX_U16X2) { unsigned t = __vmaxu2(x[i], y[i]); y[i] = x[i]; x[i] = t; }
                if (OP == OP_VMAX3_U16X2) { unsigned t = __vimax3_u16x2(x[i], y[i], z[i]); z[i] = y[i]; y[i] = x[i]; x[i] = t; }
                if (OP == OP_VADDMAX_U16X2) { unsigned t = __viaddmax_u16x2(x[i], 0xFFFFFFFFu, y[i]); y[i] = x[i]; x[i] = t; }
                if (OP == OP_VADD_16X2) { x[i] = __vadd2(x[i], y[i]); }
                if (OP == OP_IADD3) { x[i] = x[i] + y[i] + z[i]; unsigned t = y[i]; y[i] = x[i]; x[i] = t; }
                if (OP == OP_IMAD) { x[i] = x[i] * y0 + z[i]; }
                if (OP == OP_LOP3) { x[i] = (x[i] & y[i]) ^ z[i]; unsigned t = y[i]; y[i] = x[i]; x[i] = t; }
                if (OP == OP_PRMT) { x[i] = __byte_perm(x[i], y[i], 0x5410 + (z0 & 0x1111)); unsigned t = y[i]; y[i] = x[i]; x[i] = t; }
                if (OP == OP_VMAX3_S32) { unsigned t = __vimax3_s32(x[i], y[i], z[i]); z[i] = y[i]; y[i] = x[i]; x[i] = t; }
                if (OP == OP_MIX_VMAX3_IADD3) { unsigned t = __vimax3_u16x2(x[i], y[i], z[i]); z[i] = y[i] + x[i] + t; y[i] = x[i]; x[i] = t; }
                if (OP == OP_MIX_VMAX3_IMAD) { unsigned t = __vimax3_u16x2(x[i], y[i], z[i]); z[i] = y[i] * y0 + t; y[i] = x[i]; x[i] = t; }
                if (OP == OP_MIX_VADDMAX_VMAX3) { unsigned t = __viaddmax_u16x2(x[i], 0xFFFEFFFEu, y[i]); z[i] = __vimax3_u16x2(z[i], y[i], t); y[i] = x[i]; x[i] = t; }
                if (OP == OP_SHFL) { x[i] = __shfl_up_sync(0xffffffffu, x[i], 1) + 1; }
                if (OP == OP_LDS32) { x[i] = lds32(sm, (x[i] + lane_off) & 4095); }
                if (OP == OP_LDS128) { uint4 v = lds128(sm, (x[i] + lane_off) & 1023); x[i] = v.x ^ v.y ^ v.z ^ v.w; }
                if (OP == OP_MIX_VMAX3_PRMT) { unsigned t = __vimax3_u16x2(x[i], y[i], z[i]); z[i] = __byte_perm(y[i], t, 0x5410 + (z0 & 0x1111)); y[i] = x[i]; x[i] = t; }
            }
            if (OP == OP_MIX_VMAX3_LDS128) {
                // per u: 32 vimnmx3 (4 per chain) + 8 lds.128?  keep 4:1 by doing 4 max steps per chain and one lds per chain
#pragma unroll
                for (int i = 0; i < ILP; ++i) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) { unsigned t = __vimax3_u16x2(x[i], y[i], z[i]); z[i] = y[i]; y[i] = x[i]; x[i] = t; }
                    uint4 v = lds128(sm, (threadIdx.x + i * 32 + u) & 1023);
                    z[i] = v.x;
                }
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i] ^ z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, int threads, unsigned* d_out, long long* d_cyc)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = sms;
    bench<OP><<<blocks, threads>>>(d_out, 0x7fff7fffu, 0x12345678u, d_cyc); // warm
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    bench<OP><<<blocks, threads>>>(d_out, 0x7fff7fffu, 0x12345678u, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; long long mx = 0;
    for (auto c : cyc) { mean += (double)c; if (c > mx) mx = c; }
    mean /= blocks;
    const double warps = threads / 32.0;
    const double instr_per_warp = (double)ITERS * ILP * op_instr[OP];
    const double wi_per_clk_sm = warps * instr_per_warp / mean;   // warp-instructions / clk / SM
    const double mhz = (double)mx / (ms * 1e-3) / 1e6;
    printf("{\"op\": \"%s\", \"threads_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm\": %.1f, "
           "\"cycles\": %.0f, \"ms\": %.4f, \"sm_mhz_est\": %.0f}\n",
           op_name[OP], threads, wi_per_clk_sm, wi_per_clk_sm * 32.0, mean, ms, mhz);
    fflush(stdout);
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    unsigned* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, (size_t)sms * 1024 * sizeof(unsigned)));
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long)));
    for (int threads : {256, 512, 1024}) {
        run<OP_VMAX_S16X2>(sms, threads, d_out, d_cyc);
        run<OP_VMAX_U16X2>(sms, threads, d_out, d_cyc);
        run<OP_VMAX3_U16X2>(sms, threads, d_out, d_cyc);
        run<OP_VADDMAX_U16X2>(sms, threads, d_out, d_cyc);
        run<OP_VADD_16X2>(sms, threads, d_out, d_cyc);
        run<OP_IADD3>(sms, threads, d_out, d_cyc);
        run<OP_IMAD>(sms, threads, d_out, d_cyc);
        run<OP_LOP3>(sms, threads, d_out, d_cyc);
        run<OP_PRMT>(sms, threads, d_out, d_cyc);
        run<OP_VMAX3_S32>(sms, threads, d_out, d_cyc);
        run<OP_MIX_VMAX3_IADD3>(sms, threads, d_out, d_cyc);
        run<OP_MIX_VMAX3_IMAD>(sms, threads, d_out, d_cyc);
        run<OP_MIX_VADDMAX_VMAX3>(sms, threads, d_out, d_cyc);
        run<OP_MIX_VMAX3_PRMT>(sms, threads, d_out, d_cyc);
        run<OP_SHFL>(sms, threads, d_out, d_cyc);
        run<OP_LDS32>(sms, threads, d_out, d_cyc);
        run<OP_LDS128>(sms, threads, d_out, d_cyc);
        run<OP_MIX_VMAX3_LDS128>(sms, threads, d_out, d_cyc);
    }
    return 0;
}
