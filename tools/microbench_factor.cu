// Micro-benchmark of the factor warp's [A] step (8x8 Cholesky + W = L^-1 columns + forward substitution) in isolation:
// one warp alone on an SM, and the same warp next to warps that keep the FP64 pipe busy with DMMAs.
// nvcc -O3 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a tools/microbench_factor.cu -o /tmp/mbf && /tmp/mbf
#include <cstdio>
#include <cuda_runtime.h>

constexpr int tri8(int a, int b) { return a * (a + 1) / 2 + b; }
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// MODE 0: rsqrt.approx.f64 seed + 1 Newton (the kernel's code); 1: fp32 rsqrtf seed + 2 Newton; 2: no rsqrt (inv = piv * c):
// the pure DFMA chain; 3: MODE 0 without the W / z substitutions (factor only); 4: MUFU only chain
template <int MODE> __device__ __forceinline__ double inv_sqrt(double piv) {
  if (MODE == 0 || MODE == 3) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
    return fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
  } else if (MODE == 1) {
    double y = (double)rsqrtf((float)piv);
    y = fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
    return fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
  } else if (MODE == 4) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
    return y;
  }
  return piv * 0.011;
}

template <int MODE> __global__ void k_factor(const double *A, double *out, long long *cyc, int iters, int busy_warps) {
  __shared__ double Dsm[64], zs[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 64) Dsm[threadIdx.x] = A[threadIdx.x];
  if (threadIdx.x < 8) zs[threadIdx.x] = 1.0 + threadIdx.x;
  __syncthreads();
  if (warp > 0) {                                   // background: dependent-free DMMA stream on every scheduler
    if (warp > busy_warps) return;
    double a0 = lane, a1 = 1, a2 = 2, a3 = 3, x = 1e-3, y = 1e-3;
    for (int i = 0; i < iters * 24; ++i) { dmma(a0, a1, x, y, a0, a1); dmma(a2, a3, x, y, a2, a3); }
    out[64 + threadIdx.x] = a0 + a1 + a2 + a3;
    return;
  }
  double acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double a[36];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) a[tri8(i, j)] = Dsm[i * 8 + j] + (i == j ? acc * 1e-30 : 0.0);
    double zr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) zr[k] = zs[k];
    bool ok = true;
    double wv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double piv = a[tri8(k, k)];
      ok = ok && ((float)piv > 0.0f);
      const double inv = inv_sqrt<MODE>(piv);
      if (MODE != 3) {
        double sv = lane == 8 ? zr[k] : (lane == k ? 1.0 : 0.0);
#pragma unroll
        for (int j = 0; j < k; ++j) sv -= a[tri8(k, j)] * wv[j];
        wv[k] = sv * inv;
      } else wv[k] = inv;
#pragma unroll
      for (int i = k + 1; i < 8; ++i) a[tri8(i, k)] *= inv;
#pragma unroll
      for (int j = k + 1; j < 8; ++j)
#pragma unroll
        for (int i = j; i < 8; ++i) a[tri8(i, j)] -= a[tri8(i, k)] * a[tri8(j, k)];
    }
    acc += ok ? wv[7] + wv[0] + wv[3] : 1.0;
  }
  long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) cyc[0] = t1 - t0;
}

template <int MODE> void run(const char *name, int busy) {
  double h[64], *A, *out; long long *cyc;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) h[i * 8 + j] = (i == j ? 20.0 : 0.0) + 1.0 / (1 + i + j);
  cudaMalloc(&A, 512); cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 8);
  cudaMemcpy(A, h, 512, cudaMemcpyHostToDevice);
  const int iters = 2000;
  k_factor<MODE><<<1, 32 * (1 + busy)>>>(A, out, cyc, 10, busy);
  k_factor<MODE><<<1, 32 * (1 + busy)>>>(A, out, cyc, iters, busy);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s busy DMMA warps %2d : %.0f cycles per 8x8 block (%.0f per pivot)  [%s]\n", name, busy, (double)c / iters, (double)c / iters / 8,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(A); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("rsqrt.approx.f64 + 1 Newton, W + z", 0);
  run<3>("rsqrt.approx.f64 + 1 Newton, factor only", 0);
  run<1>("rsqrtf seed + 2 Newton, W + z", 0);
  run<2>("no rsqrt (DFMA chain only), W + z", 0);
  run<4>("MUFU.RSQ64H only (no Newton), W + z", 0);
  run<0>("rsqrt.approx.f64 + 1 Newton, W + z", 3);      // warps 1-3: other schedulers busy
  run<0>("rsqrt.approx.f64 + 1 Newton, W + z", 4);      // warp 4 shares the factor warp's scheduler
  run<0>("rsqrt.approx.f64 + 1 Newton, W + z", 8);      // two DMMA warps on every scheduler
  run<3>("rsqrt.approx.f64 + 1 Newton, factor only", 8);
  return 0;
}
