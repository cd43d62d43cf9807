// Micro-benchmarks behind the back substitution's chain warp (ba_solve_diag.cu): what a dependent DMMA costs when the
// result feeds the A operand (not the accumulator), an mbarrier probe, and spinning neighbours.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_chain.cu -o /tmp/mbc && /tmp/mbc
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ unsigned mb_test(unsigned mb, unsigned par) {
  unsigned done;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(mb), "r"(par) : "memory");
  return done;
}
__device__ __forceinline__ unsigned mb_try(unsigned mb, unsigned par) {
  unsigned done;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(mb), "r"(par) : "memory");
  return done;
}
// MODE 0: D -> C chain; 1: D -> A chain; 2: the chain warp's step (2 DMMA, DADD, 2 DMMA + 6 independent DMMAs);
// 3: mbarrier test_wait, result feeds the next address; 4: MODE 2 + a probe per step; SPIN: warps 1.. spin on try_wait
template <int MODE, int SPIN> __global__ void k(double *out, long long *cyc, int iters) {
  __shared__ __align__(8) unsigned long long mbar[2];
  __shared__ double sm[64];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar[1])));
    stop = 0;
  }
  if (threadIdx.x < 64) sm[threadIdx.x] = 1.0 + threadIdx.x * 1e-3;
  __syncthreads();
  const unsigned mb0 = (unsigned)__cvta_generic_to_shared(&mbar[0]), mb1 = (unsigned)__cvta_generic_to_shared(&mbar[1]);
  if (warp > 0) {
    if (SPIN) { while (!stop) { if (mb_try(mb1, 0)) break; } }
    return;
  }
  double x0 = 1.0 + lane * 1e-3, x1 = 1.0 - lane * 1e-3, b0 = 1e-3, b1 = 2e-3, a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { dmma(x0, x1, b0, b1, x0, x1); dmma(x0, x1, b0, b1, x0, x1); dmma(x0, x1, b0, b1, x0, x1); dmma(x0, x1, b0, b1, x0, x1); }
    if (MODE == 1) { dmma(x0, x1, x0, b1, 0.0, 0.0); dmma(x0, x1, x0, b1, 0.0, 0.0); dmma(x0, x1, x0, b1, 0.0, 0.0); dmma(x0, x1, x0, b1, 0.0, 0.0); }
    if (MODE == 2 || MODE == 4) {
      unsigned fd = 1;
      if (MODE == 4) fd = mb_test(mb0, 1);
      double t0v, t1v, e0, e1, n0, n1;
      dmma(e0, e1, x0, b0, a0, a1); dmma(t0v, t1v, x1, b1, e0, e1);
      dmma(e0, e1, x0, b0, a2, a3); dmma(a0, a1, x1, b1, e0, e1);
      if (MODE == 4 && !fd) { while (!mb_try(mb0, 1)) {} }
      t0v = sm[lane & 7] - t0v; t1v = sm[8 + (lane & 7)] - t1v;
      dmma(e0, e1, t0v, b0, 0.0, 0.0); dmma(n0, n1, t1v, b1, e0, e1);
      dmma(e0, e1, x0, b0, a4, a5); dmma(a2, a3, x1, b1, e0, e1);
      dmma(e0, e1, x0, b0, 0.0, 0.0); dmma(a4, a5, x1, b1, e0, e1);
      x0 = n0; x1 = n1;
    }
    if (MODE == 3) { acc += mb_test(mb0 + (acc & 8), 1); acc += mb_test(mb0 + (acc & 8), 1); acc += mb_test(mb0 + (acc & 8), 1); acc += mb_test(mb0 + (acc & 8), 1); }
  }
  long long t1 = clock64();
  stop = 1;
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb1));
  out[threadIdx.x] = x0 + x1 + a0 + a1 + a2 + a3 + a4 + a5 + acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE, int SPIN> void run(const char *name, int threads, double per) {
  double *out; long long *cyc; cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<MODE, SPIN><<<1, threads>>>(out, cyc, 100);
  k<MODE, SPIN><<<1, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-64s %4d threads: %.1f cycles/iter = %.1f per %s\n", name, threads, (double)c / iters, (double)c / iters / per, per == 1 ? "step" : "op");
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 0>("DMMA chain through the accumulator (D -> C)", 32, 4);
  run<1, 0>("DMMA chain through the A operand (D -> A)", 32, 4);
  run<3, 0>("mbarrier.test_wait, dependent", 32, 4);
  run<2, 0>("chain step: 2 DMMA, DADD, 2 DMMA (+ 6 independent DMMAs)", 32, 1);
  run<4, 0>("chain step + one mbarrier probe", 32, 1);
  run<2, 0>("chain step, 12 warps resident (11 exited)", 384, 1);
  run<2, 1>("chain step, 3 warps spinning on mbarrier.try_wait", 128, 1);
  run<2, 1>("chain step, 11 warps spinning on mbarrier.try_wait", 384, 1);
  return 0;
}
