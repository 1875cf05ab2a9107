// Research micro-benchmark (not part of the product): latency of the three-value warp sum that heads every recursion
// step -- 5-level shuffle butterfly (what the kernels use) against fp64 MMA (m8n8k4, ones matrix) as a reduction
// network.  Built and run by hand:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/wsum warp_sum_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ double sum_bfly(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// A = data (row = quad, k = lane in quad), B = ones: every lane gets its quad's sum; then three butterfly levels
__device__ __forceinline__ double sum_mma1(double v) {
  double d0, d1;
  mma884(d0, d1, v, 1.0, 0.0, 0.0);
#pragma unroll
  for (int o = 16; o >= 4; o >>= 1) d0 += __shfl_xor_sync(0xffffffffu, d0, o);
  return d0;
}
// two MMAs: quad sums, one permuting shuffle, sums of four quads, one butterfly level
__device__ __forceinline__ double sum_mma2(double v, int lane) {
  double d0, d1;
  mma884(d0, d1, v, 1.0, 0.0, 0.0);                       // d0 = sum of my quad
  const int src = 4 * (4 * ((lane >> 2) >> 2) + (lane & 3));   // lane (n, k) fetches the sum of quad 4*(n>>2) + k
  const double s = __shfl_sync(0xffffffffu, d0, src);
  double e0, e1;
  mma884(e0, e1, 1.0, s, 0.0, 0.0);                        // A = ones, B[k][n] = s: e = sum over k of the four quad sums
  e0 += __shfl_xor_sync(0xffffffffu, e0, 16);
  return e0;
}

template <int MODE>
__global__ void bench(double *out, long long *cyc, int iters) {
  const int lane = threadIdx.x & 31;
  double a = 1.0 + lane * 1e-3, b = 2.0 + lane * 1e-3, c = 3.0 + lane * 1e-3;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double sa, sb, sc;
    if (MODE == 0) { sa = sum_bfly(a); sb = sum_bfly(b); sc = sum_bfly(c); }
    else if (MODE == 1) { sa = sum_mma1(a); sb = sum_mma1(b); sc = sum_mma1(c); }
    else { sa = sum_mma2(a, lane); sb = sum_mma2(b, lane); sc = sum_mma2(c, lane); }
    // dependent chain like the recursion: next inputs depend on the sums
    a = a * 0.5 + sa * 1e-3; b = b * 0.5 + sb * 1e-3; c = c * 0.5 + sc * 1e-3;
  }
  const long long t1 = clock64();
  if (lane == 0) { cyc[blockIdx.x] = t1 - t0; }
  out[blockIdx.x * 32 + lane] = a + b + c;
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 32 * 1024); cudaMalloc(&cyc, 8 * 1024);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) bench<0><<<1, 32>>>(out, cyc, iters);
      if (mode == 1) bench<1><<<1, 32>>>(out, cyc, iters);
      if (mode == 2) bench<2><<<1, 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    long long h; double o[32];
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 256, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): %.1f cycles per three-value reduction step, check %.12g %s\n", mode,
           mode == 0 ? "5-level butterfly" : mode == 1 ? "1 MMA + 3 levels" : "2 MMA + permute + 1 level", (double)h / iters, o[0],
           cudaGetErrorString(cudaGetLastError()));
  }
  // throughput with many warps: 148 x 8 blocks of 128 threads
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (mode == 0) bench<0><<<1024, 128>>>(out, cyc, 2000);
    if (mode == 1) bench<1><<<1024, 128>>>(out, cyc, 2000);
    if (mode == 2) bench<2><<<1024, 128>>>(out, cyc, 2000);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d throughput: %.3f ms for 4096 warps x 2000 steps\n", mode, ms);
  }
  return 0;
}
