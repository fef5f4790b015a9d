// scan.cuh — exclusive prefix sum of an int array on the device (set-up code: list compaction, per-brick offsets).
// Three levels of 1024-element blocks cover 2^30 elements.  Sums must stay below 2^31 (they index arrays that do).
#pragma once
#include <cuda_runtime.h>

namespace mflbm {

constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

// out[i] = sum of in[first .. i) inside each block of SCAN_BLOCK elements; block_sums[b] = total of block b
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ block_sums, const long long n) {
    __shared__ int warp_tot[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? in[base + k] : 0; tot += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += up; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; w++) off += warp_tot[w];
    int run = off + inc - tot;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (block_sums && threadIdx.x == SCAN_THREADS - 1) block_sums[blockIdx.x] = off + inc;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int* __restrict__ out, const int* __restrict__ block_off, const long long n) {
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    const int o = block_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) out[base + k] += o;
}

// out[0 .. n) = exclusive scan of in[0 .. n) (in == out allowed); returns the total (synchronises the stream)
inline cudaError_t exclusive_scan(const int* d_in, int* d_out, long long n, cudaStream_t stream, long long* total) {
    if (total) *total = 0;
    if (n <= 0) return cudaSuccess;
    const long long nb1 = (n + SCAN_BLOCK - 1) / SCAN_BLOCK, nb2 = (nb1 + SCAN_BLOCK - 1) / SCAN_BLOCK, nb3 = (nb2 + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nb3 > 1) return cudaErrorInvalidValue;
    int *s1 = nullptr, *s2 = nullptr, *s3 = nullptr;
    cudaError_t e = cudaMalloc((void**)&s1, sizeof(int) * (size_t)(nb1 + nb2 + 2));
    if (e != cudaSuccess) return e;
    s2 = s1 + nb1; s3 = s2 + nb2;
    k_scan_blocks<<<(unsigned)nb1, SCAN_THREADS, 0, stream>>>(d_in, d_out, s1, n);
    k_scan_blocks<<<(unsigned)nb2, SCAN_THREADS, 0, stream>>>(s1, s1, s2, nb1);
    k_scan_blocks<<<1, SCAN_THREADS, 0, stream>>>(s2, s2, s3, nb2);
    if (nb2 > 1) k_scan_add<<<(unsigned)nb2, SCAN_THREADS, 0, stream>>>(s1, s2, nb1);
    if (nb1 > 1) k_scan_add<<<(unsigned)nb1, SCAN_THREADS, 0, stream>>>(d_out, s1, n);
    int tot = 0;
    e = cudaMemcpyAsync(&tot, s3, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(s1);
    if (total) *total = tot;
    return e;
}

}  // namespace mflbm
