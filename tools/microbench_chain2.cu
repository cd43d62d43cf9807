// Replica of the back substitution's chain warp + helper warps (ba_solve_diag.cu) with the parts switchable:
// finds which part of the step costs what. nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int OFF> __device__ __forceinline__ double lds64o(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ double2 lds128o(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ void sts128o(unsigned a, double x, double y) { asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(x), "d"(y) : "memory"); }
template <int OFF> __device__ __forceinline__ unsigned mbar_test_o(unsigned a, unsigned par) {
  unsigned done;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1+%2], %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(a), "n"(OFF), "r"(par) : "memory");
  return done;
}
template <int OFF> __device__ __forceinline__ void mbar_wait_o(unsigned a, unsigned par) {
  unsigned done = 0;
  while (!done) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1+%2], %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(a), "n"(OFF), "r"(par) : "memory");
}
template <int OFF> __device__ __forceinline__ void mbar_arrive_o(unsigned a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1];" ::"r"(a), "n"(OFF) : "memory"); }
constexpr int kStRow = 1056, kStSlot = 8 * kStRow, kNear = 3;
// F bits: 1 operand loads, 2 x store, 4 arrive, 8 probe of fdone (needs helpers), 16 helpers run, 32 helpers: lane 0 waits only
template <int F> __global__ void __launch_bounds__(384, 1) k(double *out, long long *cyc, int rounds, const double *__restrict__ Lg) {
  extern __shared__ __align__(16) double dsm[];
  __shared__ __align__(8) unsigned long long s_mb[48];
  double *z = dsm, *xsol = dsm + 2048, *Lst = dsm + 4096;
  const int hw = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  for (int i = threadIdx.x; i < 4096 + 16 * kStSlot / 8; i += blockDim.x) dsm[i] = 1e-3 * (i & 63);
  if (threadIdx.x == 0)
    for (int k2 = 0; k2 < 48; ++k2) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mb[k2])), "r"(k2 >= 32 ? 3 : 1) : "memory");
  __syncthreads();
  const unsigned mbb = (unsigned)__cvta_generic_to_shared(s_mb);
  if (hw == 0) {
    const unsigned stg = (unsigned)__cvta_generic_to_shared(Lst) + (unsigned)(2 * q * kStRow + g * 8);
    unsigned zq = (unsigned)__cvta_generic_to_shared(z) + (unsigned)(2047 * 8 - 56 + 2 * q * 8), xq = zq + 2048 * 8;
    unsigned parR = 0;
    bool later = false;
    double x0 = 1.0, x1 = 0.5, a00 = 0, a01 = 0, a10 = 0, a11 = 0, b00 = 1e-3, b01 = 1e-3, b10 = 1e-3, b11 = 1e-3, b20 = 1e-3, b21 = 1e-3, w0 = 1e-3, w1 = 1e-3;
    auto step = [&](auto kc) {
      constexpr int k = decltype(kc)::value, k1 = (k + 1) & 15, kf = (k - kNear - 1) & 15;
      constexpr int so = k * kStSlot, so1 = k1 * kStSlot;
      const unsigned pfd = k >= kNear + 1 ? parR : parR ^ 1u;
      unsigned fd = 1, fl = 1;
      const unsigned pf1 = k == 15 ? parR ^ 1u : parR;
      if (F & 64) fl = mbar_test_o<8 * k1>(mbb, pf1);
      if ((F & 8) && !(F & 256) && (k >= kNear + 1 || later)) fd = mbar_test_o<8 * (32 + kf)>(mbb, pfd);
      double t0, t1, e0, e1;
      dmma884(e0, e1, x0, b00, a00, a01); dmma884(t0, t1, x1, b01, e0, e1);
      dmma884(e0, e1, x0, b10, a10, a11); dmma884(a00, a01, x1, b11, e0, e1);
      if (!(fd & fl)) { if ((F & 8) && !(F & 256)) mbar_wait_o<8 * (32 + kf)>(mbb, pfd); if (F & 64) mbar_wait_o<8 * k1>(mbb, pf1); }
      const double2 zz = lds128o<-64 * k>(zq);
      t0 = zz.x - t0; t1 = zz.y - t1;
      double xn0, xn1;
      dmma884(e0, e1, t0, w0, 0.0, 0.0); dmma884(xn0, xn1, t1, w1, e0, e1);
      dmma884(e0, e1, x0, b20, 0.0, 0.0); dmma884(a10, a11, x1, b21, e0, e1);
      if (F & 1) {
        w0 = lds64o<so1 + 120 * 8>(stg); w1 = lds64o<so1 + 120 * 8 + kStRow>(stg);
        b00 = lds64o<so + 112 * 8>(stg); b01 = lds64o<so + 112 * 8 + kStRow>(stg);
        b10 = lds64o<so + 104 * 8>(stg); b11 = lds64o<so + 104 * 8 + kStRow>(stg);
        b20 = lds64o<so + 96 * 8>(stg); b21 = lds64o<so + 96 * 8 + kStRow>(stg);
      }
      if ((F & 2) && g == 0) sts128o<-64 * k>(xq, xn0, xn1);
      if (F & 4) { __syncwarp(); if (lane == 0) mbar_arrive_o<8 * (16 + k)>(mbb); }
      x0 = xn0; x1 = xn1;
    };
    if (F & 64) mbar_wait_o<0>(mbb, 0);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      step(std::integral_constant<int, 0>{}); step(std::integral_constant<int, 1>{}); step(std::integral_constant<int, 2>{}); step(std::integral_constant<int, 3>{});
      step(std::integral_constant<int, 4>{}); step(std::integral_constant<int, 5>{}); step(std::integral_constant<int, 6>{}); step(std::integral_constant<int, 7>{});
      step(std::integral_constant<int, 8>{}); step(std::integral_constant<int, 9>{}); step(std::integral_constant<int, 10>{}); step(std::integral_constant<int, 11>{});
      step(std::integral_constant<int, 12>{}); step(std::integral_constant<int, 13>{}); step(std::integral_constant<int, 14>{}); step(std::integral_constant<int, 15>{});
      parR ^= 1u; later = true;
      if ((r & 7) == 7) { zq += 7 * 1024; xq += 7 * 1024; } else { zq -= 1024; xq -= 1024; }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = x0 + x1 + a00 + a01 + a10 + a11;
  } else if (hw <= 3 && (F & 16)) {
    const int tau = threadIdx.x - 32, i = tau & 7;
    const unsigned mbx = mbb + 128u, st0 = (unsigned)__cvta_generic_to_shared(Lst) + (unsigned)i * 8u;
    unsigned xa = (unsigned)__cvta_generic_to_shared(xsol) + 2040 * 8, slot = 0, ring = 0, par = 0;
    int m = tau >> 3, cnt = 0;
    double facc = 0.0;
    for (int n = 0; n < rounds * 16; ++n) {
      const int d = 15 - m;
      const unsigned st = st0 + slot * (unsigned)kStSlot + (unsigned)(120 - 8 * d) * 8u;
      if (F & 32) { if (lane == 0) mbar_wait_o<0>(mbx + ring * 8u, par); __syncwarp(); } else mbar_wait_o<0>(mbx + ring * 8u, par);
      const double2 xa0 = lds128o<0>(xa), xb0 = lds128o<16>(xa), xc0 = lds128o<32>(xa), xd0 = lds128o<48>(xa);
      const double h0 = lds64o<0>(st), h1 = lds64o<kStRow>(st), h2 = lds64o<2 * kStRow>(st), h3 = lds64o<3 * kStRow>(st);
      const double h4 = lds64o<4 * kStRow>(st), h5 = lds64o<5 * kStRow>(st), h6 = lds64o<6 * kStRow>(st), h7 = lds64o<7 * kStRow>(st);
      const double c0 = fma(h1, xa0.y, h0 * xa0.x), c1s = fma(h3, xb0.y, h2 * xb0.x), c2 = fma(h5, xc0.y, h4 * xc0.x), c3 = fma(h7, xd0.y, h6 * xd0.x);
      facc += (c0 + c1s) + (c2 + c3);
      if (d == kNear + 1) { z[(8 * (2047 - cnt) + i) & 2047] -= facc; facc = 0.0; }
      m = m == 11 ? 0 : m + 1;
      __syncwarp();
      if (lane == 0) mbar_arrive_o<128>(mbx + ring * 8u);
      ++cnt; if (cnt == 128) { cnt = 0; xa += 127 * 64; } else xa -= 64u;
      slot = (slot + 1) & 15u; ring = (ring + 1) & 15u; par ^= (ring == 0);
    }
    out[64 + threadIdx.x] = facc;
  } else if (hw == 5 && lane == 0 && (F & 64)) {
    auto issue = [&](int n) {
      const unsigned mb = mbb + (n & 15) * 8, dst = (unsigned)__cvta_generic_to_shared(Lst) + (unsigned)(n & 15) * kStSlot;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(8192) : "memory");
      if (F & 128) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(Lg + (size_t)(n & 127) * 1024), "r"(8192), "r"(mb) : "memory");
      } else {
      for (int rr = 0; rr < 8; ++rr)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + rr * kStRow), "l"(Lg + (size_t)(n & 127) * 1024 + rr * 128), "r"(1024), "r"(mb) : "memory");
      }
    };
    for (int n = 0; n < 16; ++n) issue(n);
    for (int n = 16; n < rounds * 16 + 1; ++n) {
      if (!(F & 256)) mbar_wait_o<0>(mbb + 256 + ((n - 16) & 15) * 8, ((n - 16) >> 4) & 1);
      mbar_wait_o<0>(mbb + 128 + ((n - 15) & 15) * 8, ((n - 15) >> 4) & 1);
      issue(n);
    }
  }
}
template <int F> void run(const char *name) {
  double *out, *Lg; long long *cyc; cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, 8); cudaMalloc(&Lg, 128 * 8192); cudaMemset(Lg, 0, 128 * 8192);
  const int smem = (4096 + 16 * kStSlot / 8) * 8;
  cudaFuncSetAttribute(k<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<F><<<1, 384, smem>>>(out, cyc, 8, Lg);
  k<F><<<1, 384, smem>>>(out, cyc, 64, Lg);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-72s %.1f cycles/step  (%s)\n", name, (double)c / (64 * 16), cudaGetErrorString(e));
  fflush(stdout);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("8 DMMA + z load + 2 DADD, unrolled x16");
  run<1>("+ 8 operand loads");
  run<3>("+ x store");
  run<7>("+ syncwarp, arrive");
  run<7 + 16 + 8>("+ 3 helper warps (96 threads wait for x, far field, arrive), chain probes fdone");
  run<7 + 16 + 8 + 64>("+ loader thread (8 bulk copies of 1 KB per step), chain probes full");
  run<7 + 16 + 8 + 64 + 128>("loader thread: 1 bulk copy of 8 KB per step");
  run<7 + 8 + 64 + 128 + 256>("no helpers; loader: 1 bulk copy of 8 KB per step, free-running after xrdy only");
  return 0;
}
