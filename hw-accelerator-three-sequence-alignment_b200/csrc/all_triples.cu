// all_triples.cu -- the caller that produces "a million short triples" (the progressive three-way
// alignment of pic/Motivation.png needs the score of every triple of a sequence set): the N sequences are uploaded
// once (N x len bytes), a gather kernel lays the C(N,3) triples out as the [n][len] batch arrays the batch kernels
// read -- in HBM, never on the host -- and K1 / K2 score them.  Triple t is the t-th (x < y < z) in lexicographic
// order, with roles A = x, B = y, C = z (the RTL's score is not symmetric in the roles, src/PE_1cyc.v:162).
#include "kernels.cuh"

namespace trialign {
namespace {

__host__ __device__ inline long long choose2(long long m) { return m < 2 ? 0 : m * (m - 1) / 2; }
__host__ __device__ inline long long choose3(long long m) { return m < 3 ? 0 : m * (m - 1) * (m - 2) / 6; }

// lexicographic unranking: triples whose first index is < x number C(n,3) - C(n-x,3)
__device__ inline void unrank3(int n, long long t, int &x, int &y, int &z)
{
    const long long all = choose3(n);
    int lo = 0, hi = n - 3;  // largest x with (all - C(n-x,3)) <= t
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (all - choose3(n - mid) <= t) lo = mid; else hi = mid - 1;
    }
    x = lo;
    long long r = t - (all - choose3(n - x));
    const int m = n - x - 1;  // y, z are chosen among the m indices above x
    const long long all2 = choose2(m);
    int lo2 = 0, hi2 = m - 2;
    while (lo2 < hi2) {
        const int mid = (lo2 + hi2 + 1) >> 1;
        if (all2 - choose2(m - mid) <= r) lo2 = mid; else hi2 = mid - 1;
    }
    y = x + 1 + lo2;
    r -= all2 - choose2(m - lo2);
    z = y + 1 + (int)r;
}

// one warp per triple: lane 0 unranks, all lanes copy the three rows
__global__ void __launch_bounds__(256)
gather_triples_kernel(const uint8_t *__restrict__ seqs, int nseq, int len, long long first, long long count,
                      uint8_t *__restrict__ a, uint8_t *__restrict__ b, uint8_t *__restrict__ c)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < count; t += nwarps) {
        int x = 0, y = 0, z = 0;
        if (lane == 0) unrank3(nseq, first + t, x, y, z);
        x = __shfl_sync(0xFFFFFFFFu, x, 0);
        y = __shfl_sync(0xFFFFFFFFu, y, 0);
        z = __shfl_sync(0xFFFFFFFFu, z, 0);
        const uint8_t *sx = seqs + (size_t)x * len, *sy = seqs + (size_t)y * len, *sz = seqs + (size_t)z * len;
        uint8_t *da = a + (size_t)t * len, *db = b + (size_t)t * len, *dc = c + (size_t)t * len;
        for (int p = lane; p < len; p += 32) {
            da[p] = sx[p];
            db[p] = sy[p];
            dc[p] = sz[p];
        }
    }
}

}  // namespace

long long all_triples_count(int nseq) { return choose3(nseq); }

int all_triples_gather(cudaStream_t st, const uint8_t *d_seqs, int nseq, int len, long long first, long long count,
                       uint8_t *d_a, uint8_t *d_b, uint8_t *d_c)
{
    if (count <= 0) return TRIALIGN_OK;
    const long long blocks = (count * 32 + 255) / 256;
    gather_triples_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, st>>>(d_seqs, nseq, len, first, count, d_a,
                                                                                             d_b, d_c);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TRIALIGN_OK;
}

}  // namespace trialign
