// Micro-benchmarks of the fp64 paths the reduced solve depends on (run on the B200 box):
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o /tmp/mb && /tmp/mb
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int MODE> __global__ void k(double *out, long long *cyc, int iters) {
  double x = threadIdx.x * 1e-3 + 1.0, y = 1.0000001, a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
  float fx = (float)x, fy = 1.0000001f;
  __shared__ double sm[1024];
  sm[threadIdx.x] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }            // dependent DFMA x4
    if (MODE == 1) { fx = fmaf(fx, fy, 1e-9f); fx = fmaf(fx, fy, 1e-9f); fx = fmaf(fx, fy, 1e-9f); fx = fmaf(fx, fy, 1e-9f); }
    if (MODE == 2) { a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9); }    // independent DFMA x4
    if (MODE == 3) { dmma(a0, a1, x, y, a0, a1); dmma(a0, a1, x, y, a0, a1); dmma(a0, a1, x, y, a0, a1); dmma(a0, a1, x, y, a0, a1); }  // dependent DMMA x4
    if (MODE == 4) { dmma(a0, a1, x, y, a0, a1); dmma(a2, a3, x, y, a2, a3); dmma(x, fx == 0 ? y : a0, a0, y, x, a0); dmma(a0, a1, x, y, a0, a1); }
    if (MODE == 5) { x = (double)rsqrtf((float)x) + 1.0; x = (double)rsqrtf((float)x) + 1.0; x = (double)rsqrtf((float)x) + 1.0; x = (double)rsqrtf((float)x) + 1.0; }
    if (MODE == 6) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
    if (MODE == 7) { int idx = (int)x & 1023; x = sm[idx] + 1.0; idx = (int)x & 1023; x = sm[idx] + 1.0; idx = (int)x & 1023; x = sm[idx] + 1.0; idx = (int)x & 1023; x = sm[idx] + 1.0; }
    if (MODE == 8) { x = sqrt(x) + 1.0; x = 1.0 / x + 1.0; x = sqrt(x) + 1.0; x = 1.0 / x + 1.0; }
    if (MODE == 9) { x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0; x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0; x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0; x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + a0 + a1 + a2 + a3 + fx;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char *name, int threads, int blocks, double ops_per_iter_thread) {
  double *out; long long *cyc; cudaMalloc(&out, sizeof(double) * threads * blocks); cudaMalloc(&cyc, 8);
  int iters = 20000;
  k<MODE><<<blocks, threads>>>(out, cyc, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s threads %4d blocks %3d : %.1f cycles/iter (4 ops) => %.1f cyc/op ; %.3f ms ; %.2f Gop/s\n", name, threads, blocks,
         (double)c / iters, (double)c / iters / 4, ms, ops_per_iter_thread * threads * blocks * iters / ms / 1e6);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("dependent DFMA, 1 warp", 32, 1, 4);
  run<1>("dependent FFMA, 1 warp", 32, 1, 4);
  run<2>("independent DFMA x4, 1 warp", 32, 1, 4);
  run<2>("independent DFMA x4, 8 warps", 256, 1, 4);
  run<2>("independent DFMA x4, 32 warps", 1024, 1, 4);
  run<2>("indep DFMA x4, 148x1024", 1024, 148, 4);
  run<3>("dependent DMMA, 1 warp", 32, 1, 4);
  run<3>("dependent DMMA, 8 warps", 256, 1, 4);
  run<3>("dependent DMMA, 32 warps", 1024, 1, 4);
  run<3>("dep DMMA 148x1024 (x512 flop)", 1024, 148, 4);
  run<5>("rsqrtf via cvt chain, 1 warp", 32, 1, 4);
  run<6>("__syncthreads x4, 256 thr", 256, 1, 4);
  run<6>("__syncthreads x4, 1024 thr", 1024, 1, 4);
  run<7>("LDS.64 dependent, 1 warp", 32, 1, 4);
  run<8>("sqrt+div fp64 chain, 1 warp", 32, 1, 4);
  run<9>("shfl double + add, 1 warp", 32, 1, 4);
  return 0;
}
