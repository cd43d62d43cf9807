// ba_solve_mma.cu — reduced camera solve on the FP64 tensor cores (ba.py:60-70 block_solve, :5-19
// CholeskySolver, :323-325 NaN retry).
//
//   A = S + (ep + lm diag S) I ;  A = L L^T ;  dX = A^-1 y          band half-width bw <= 120
//
// Blocked right-looking band Cholesky with 8x8 tiles, one CTA of 8 warps. The active window is the
// 16x16-tile square [J, J+15]^2 (128 scalar rows); every lower tile of it lives in REGISTERS as the
// C fragment of an m8n8k4 DMMA, owned by a fixed warp chosen from the circular tile positions
// (a mod 16, b mod 16) so that each warp holds 17 tiles, 15 of which are updated per step. Per tile
// column J:
//   [D] the owner of tile (J,J) drops it to shared memory; warp 0 factors the 8x8 block in registers
//       (fp64, rsqrt seed + Newton), inverts the factor (W = L_JJ^-1), and does the forward substitution
//       of the right-hand side for the block;
//   [P] the 15 panel tiles become L_aJ = A_aJ W^T with two DMMAs each (the C fragment is turned into A
//       fragments with warp shuffles), go to shared memory + global L, update the right-hand side, and
//       their registers are refilled from global with the tile row that enters the window;
//   [U] the 120 trailing tiles get  C_ab -= L_aJ L_bJ^T  with two DMMAs each, operands from shared memory.
// The backward substitution streams the tile rows of L back through shared memory.
// Why fp64: the reduced system of a short window is ill-conditioned (kappa ~ 1e3..1e4); solving it in
// fp32 puts the result at the reference's own fp32 noise floor (~1e-4), see DESIGN.md §Precision.
#include "ba_internal.h"

namespace ba {

constexpr int kMmaThreads = 256, kMmaWarps = 8;
constexpr int kTilesPerWarp = 17;
constexpr int kPs = 12;            // row stride (doubles) of the shared 8x8 tiles: conflict-free fragment loads

// the 136 unordered pairs {x <= y} of the 16 circular tile positions, diagonal-major; pair idx belongs
// to warp idx % 8 as its (idx / 8)-th tile. Every position appears in exactly 2 tiles of every warp.
__constant__ unsigned char c_px[136] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 0, 1, 2,
    3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9,
    10, 11, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 0, 1,
    2, 3, 4, 5, 6, 7, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 1, 2, 0, 1, 0};
__constant__ unsigned char c_py[136] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 2, 3, 4,
    5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13,
    14, 15, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 7, 8, 9, 10, 11, 12, 13, 14, 15, 8, 9,
    10, 11, 12, 13, 14, 15, 9, 10, 11, 12, 13, 14, 15, 10, 11, 12, 13, 14, 15, 11, 12, 13, 14, 15, 12, 13, 14, 15, 13, 14, 15, 14, 15, 15};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

__device__ __forceinline__ double rsqrt64(double x) {
  double y = (double)rsqrtf((float)x);
  y = y * (1.5 - 0.5 * x * y * y);
  y = y * (1.5 - 0.5 * x * y * y);
  y = y * (1.5 - 0.5 * x * y * y);
  return y;
}

__device__ __forceinline__ void cp_async8_d(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

constexpr int tri8(int a, int b) { return a * (a + 1) / 2 + b; }

__global__ void __launch_bounds__(kMmaThreads, 1) k_solve_band_mma(CallView cv, int allow_retry, double *__restrict__ Wg) {
  extern __shared__ double dsm[];
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int M = cv.M, bw = cv.bw, ld = cv.ld, off = cv.off;
  const int NT8 = (M + 7) >> 3, Mp = NT8 * 8;
  double *z = dsm;                         // [Mp]   right-hand side -> forward solution -> solution
  double *Psm = z + Mp;                    // [16][8][kPs]  panel tiles L_aJ by circular position
  double *Dsm = Psm + 16 * 8 * kPs;        // [8][kPs]      raw diagonal tile
  double *Wsm = Dsm + 8 * kPs;             // [8][kPs]      W = L_JJ^-1
  double *zJ = Wsm + 8 * kPs;              // [8]
  double *Lst = zJ + 8;                    // [2][8][128 + 64] back-substitution stages: 8 rows of L + W_J
  __shared__ int s_fail, s_nan;
  const double *__restrict__ S = cv.S;
  double *__restrict__ L = cv.L;
  const double ep = (double)cv.ep;
  auto Sg = [&](int r, int c) { return (size_t)r * ld + c + off; };
  int status = 0;

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    // value of the damped matrix at (r, c), r >= c, with identity padding beyond M
    auto Aval = [&](int r, int c) -> double {
      if (r >= M) return r == c ? 1.0 : 0.0;
      if (c > r || r - c > bw) return 0.0;
      double v = S[Sg(r, c)];
      if (r == c) v = v + (ep + lm * v);                        // ba.py:67
      return v;
    };
    auto load_tile = [&](int a, int b, double &c0, double &c1) {   // tile (a, b), a >= b, C-fragment layout
      const int r = 8 * a + g, c = 8 * b + 2 * q;
      c0 = Aval(r, c);
      c1 = Aval(r, c + 1);
    };

    for (int r = tau; r < Mp; r += kMmaThreads) z[r] = r < M ? cv.y[r] : 0.0;
    if (tau == 0) { s_fail = 0; s_nan = 0; }
    double ct[kTilesPerWarp][2];
#pragma unroll
    for (int t = 0; t < kTilesPerWarp; ++t) {
      const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];   // x <= y: initial window holds tile (y, x)
      ct[t][0] = ct[t][1] = 0.0;
      if (y < NT8) load_tile(y, x, ct[t][0], ct[t][1]);
    }
    bool failed = false;

    for (int J = 0; J < NT8; ++J) {
      const int e = J & 15;
      __syncthreads();                                             // previous [U] done with Psm
      // ---- [D] diagonal tile -> shared ----
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];
        if (x == e && y == e) { Dsm[g * kPs + 2 * q] = ct[t][0]; Dsm[g * kPs + 2 * q + 1] = ct[t][1]; }
      }
      __syncthreads();
      if (warp == 0) {
        // every lane factors the 8x8 block redundantly in registers (no divergence, no extra exchange)
        double a[36];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) a[tri8(i, j)] = Dsm[i * kPs + j];
        bool ok = true;
        double invd[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double piv = a[tri8(k, k)];
          ok = ok && (piv > 0.0);                                  // potrf info != 0 (also NaN), ba.py:11
          const bool fast = piv > 1e-30 && piv < 1e30;
          const double inv = fast ? rsqrt64(piv) : 1.0 / sqrt(piv);
          invd[k] = inv;
          a[tri8(k, k)] = piv * inv;
#pragma unroll
          for (int i = k + 1; i < 8; ++i) a[tri8(i, k)] *= inv;
#pragma unroll
          for (int j = k + 1; j < 8; ++j)
#pragma unroll
            for (int i = j; i < 8; ++i) a[tri8(i, j)] -= a[tri8(i, k)] * a[tri8(j, k)];
        }
        if (!ok) { if (lane == 0) s_fail = 1; }
        else {
          // L_JJ to global (band storage), rows < M only
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int j = 0; j <= i; ++j)
                if (8 * J + i < M) L[Sg(8 * J + i, 8 * J + j)] = a[tri8(i, j)];
          }
          // forward substitution of the block with L_JJ: zJ = L_JJ^-1 z_J
          double zo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            double s = z[8 * J + i];
#pragma unroll
            for (int j = 0; j < i; ++j) s -= a[tri8(i, j)] * zo[j];
            zo[i] = s * invd[i];
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { z[8 * J + i] = zo[i]; zJ[i] = zo[i]; }
          }
          // W = L_JJ^-1 (lower triangular), one column at a time straight to shared / global memory
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            double wc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) wc[i] = 0.0;
            wc[j] = invd[j];
#pragma unroll
            for (int i = j + 1; i < 8; ++i) {
              double s = 0.0;
#pragma unroll
              for (int m2 = j; m2 < i; ++m2) s += a[tri8(i, m2)] * wc[m2];
              wc[i] = -s * invd[i];
            }
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { Wsm[i * kPs + j] = wc[i]; Wg[(size_t)J * 64 + i * 8 + j] = wc[i]; }
            }
          }
        }
      }
      __syncthreads();
      if (s_fail) { failed = true; break; }
      // ---- [P] panel tiles: L_aJ = A_aJ W^T; refill the registers with the entering tile row ----
      {
        const double wb0 = Wsm[g * kPs + q], wb1 = Wsm[g * kPs + 4 + q];   // B[k][n] = W[n][k], n = g, k = 4s + q
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];
          if (x != e && y != e) continue;
          const int xo = (x == e) ? y : x;                         // the other position (== e for the diagonal tile)
          const int a = J + ((xo - e) & 15);                       // global tile row held at position xo
          if (xo != e && a < NT8) {
            // C fragment (cols 2q, 2q+1 of row g) -> A fragments (col 4s + q of row g)
            const int src0 = (lane & ~3) | (q >> 1), src1 = (lane & ~3) | (2 + (q >> 1));
            const double v00 = __shfl_sync(0xffffffffu, ct[t][0], src0), v01 = __shfl_sync(0xffffffffu, ct[t][1], src0);
            const double v10 = __shfl_sync(0xffffffffu, ct[t][0], src1), v11 = __shfl_sync(0xffffffffu, ct[t][1], src1);
            const double a0 = (q & 1) ? v01 : v00, a1 = (q & 1) ? v11 : v10;
            double p0, p1;
            dmma884(p0, p1, a0, wb0, 0.0, 0.0);
            dmma884(p0, p1, a1, wb1, p0, p1);
            double *pt = Psm + (xo * 8 + g) * kPs + 2 * q;
            pt[0] = p0; pt[1] = p1;
            const int r = 8 * a + g, c = 8 * J + 2 * q;
            if (r < M) {                                           // columns of tile J are < M whenever a row below is
              if (r - c <= bw) L[Sg(r, c)] = p0;
              if (r - c - 1 <= bw) L[Sg(r, c + 1)] = p1;
            }
            // right-hand side: z_a -= L_aJ zJ
            double part = p0 * zJ[2 * q] + p1 * zJ[2 * q + 1];
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0) z[r] -= part;
          } else if (xo != e) {
            // warp-uniform branch: nothing to do for tiles below the matrix
          }
          // refill: position e now stands for tile index J + 16
          const int an = J + 16, bn = (xo == e) ? J + 16 : a;
          ct[t][0] = ct[t][1] = 0.0;
          if (an < NT8) load_tile(an, bn, ct[t][0], ct[t][1]);
        }
      }
      __syncthreads();
      // ---- [U] trailing update: C_ab -= L_aJ L_bJ^T ----
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];
        if (x == e || y == e) continue;
        const int ax = J + ((x - e) & 15), ay = J + ((y - e) & 15);
        const int pa = ax > ay ? x : y, pb = ax > ay ? y : x;       // row tile = larger global index
        if (max(ax, ay) >= NT8) continue;
        const double *A = Psm + (pa * 8 + g) * kPs + q, *B = Psm + (pb * 8 + g) * kPs + q;
        dmma884(ct[t][0], ct[t][1], -A[0], B[0], ct[t][0], ct[t][1]);
        dmma884(ct[t][0], ct[t][1], -A[4], B[4], ct[t][0], ct[t][1]);
      }
    }
    __syncthreads();
    if (failed) {                                                   // dX = 0 (ba.py:12-13); no NaN -> no retry
      for (int r = tau; r < M; r += kMmaThreads) cv.dX[r] = 0.0;
      status |= (attempt == 0) ? 1 : 4;
      break;
    }

    // ---- backward substitution L^T x = z by tile rows, descending. Stage s holds rows 8J..8J+7 of L
    //      restricted to columns [8(J-15), 8J) plus W_J. ----
    auto stage_load = [&](int J, int sidx) {
      double *dst = Lst + (size_t)sidx * (8 * 128 + 64);
      for (int o = tau; o < 8 * 120 + 64; o += kMmaThreads) {
        if (o < 8 * 120) {
          const int gg = o / 120, xcol = o - gg * 120;
          const int r = 8 * J + gg, c = 8 * (J - 15) + xcol;
          if (r < M && c >= 0 && r - c <= bw) cp_async8_d(dst + gg * 128 + xcol, L + Sg(r, c));
          else dst[gg * 128 + xcol] = 0.0;
        } else {
          cp_async8_d(dst + 8 * 128 + (o - 8 * 120), Wg + (size_t)J * 64 + (o - 8 * 120));
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __threadfence_block();
    stage_load(NT8 - 1, (NT8 - 1) & 1);
    for (int J = NT8 - 1; J >= 0; --J) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                             // stage J landed; iteration J+1 fully retired
      if (J > 0) stage_load(J - 1, (J - 1) & 1);                   // overlaps this iteration
      const double *st = Lst + (size_t)(J & 1) * (8 * 128 + 64);
      const double *Wj = st + 8 * 128;                             // W_J row-major 8x8
      // x_J = W_J^T z_J. The 120 column threads need all 8 values (registers, static indices); the 8
      // writer threads compute their own component.
      double xJ[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = i; k < 8; ++k) s += Wj[k * 8 + i] * z[8 * J + k];
        xJ[i] = s;
      }
      double xmine = 0.0;
      if (tau < 8) for (int k = tau; k < 8; ++k) xmine += Wj[k * 8 + tau] * z[8 * J + k];
      __syncthreads();                                             // everyone has read z_J
      if (tau < 8) z[8 * J + tau] = xmine;
      if (tau >= 32 && tau < 32 + 120) {
        const int xcol = tau - 32, c = 8 * (J - 15) + xcol;
        if (c >= 0) {
          double s = 0.0;
#pragma unroll
          for (int gg = 0; gg < 8; ++gg) s += st[gg * 128 + xcol] * xJ[gg];
          z[c] -= s;
        }
      }
    }
    __syncthreads();
    int nan_local = 0;
    for (int r = tau; r < M; r += kMmaThreads) { const double v = z[r]; cv.dX[r] = v; nan_local |= (v != v); }
    if (nan_local) s_nan = 1;
    __syncthreads();
    if (s_nan && allow_retry && attempt == 0) { status |= 2; __syncthreads(); continue; }   // ba.py:324-325
    break;
  }
  if (tau == 0) cv.status[0] = status;
}

size_t solve_mma_smem_bytes(int M) {
  const int Mp = ((M + 7) / 8) * 8;
  return ((size_t)Mp + 16 * 8 * kPs + 2 * 8 * kPs + 8 + 2 * (8 * 128 + 64)) * sizeof(double);
}

int launch_solve_band_mma(const CallView &cv, int allow_retry, double *Wg, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    BA_CUDA(cudaFuncSetAttribute(k_solve_band_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
    attr_set = true;
  }
  k_solve_band_mma<<<1, kMmaThreads, solve_mma_smem_bytes(cv.M), s>>>(cv, allow_retry, Wg);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
