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

constexpr int kMmaWarps = 8;                 // tile warps; warp 8 is the factor warp
constexpr int kMmaThreads = 32 * (kMmaWarps + 1);
constexpr int kTilesPerWarp = 17;
constexpr int kBackStages = 4;     // tile rows of L in flight during the back substitution
constexpr int kPs = 12;            // row stride (doubles) of the shared 8x8 tiles: conflict-free fragment loads
constexpr int kTs = 8 * kPs;       // doubles per shared tile

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
  // fp32 seed (rel. error ~2e-7) + one Newton step -> ~6e-14. The factor is then exact for a matrix that
  // differs from A by 1e-13 relative, five orders below what the fp32 edge terms carry.
  const double hy = 0.5 * y;
  return fma(fma(-x * y, hy, 0.5), y, y);  // y + y * (0.5 - 0.5 x y^2)

}

// named barriers (id 0 is __syncthreads): 1 = panel tiles complete (tile warps), 2 = diagonal tile published
// (owner warp -> factor warp), 3 = W_J / zJ ready (factor warp -> tile warps; also closes the previous [U])
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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
  double *Psm = z + Mp;                    // [17][8][kPs]  panel tiles L_aJ by circular position; tile 16 = zeros
  double *Nsm = Psm + 17 * kTs;            // [17][8][kPs]  the same, negated (A operand of C -= L L^T)
  double *Dsm = Nsm + 17 * kTs;            // [8][kPs]      raw diagonal tile
  double *Wsm = Dsm + 8 * kPs;             // [8][kPs]      W = L_JJ^-1
  double *zJ = Wsm + 8 * kPs;              // [8]
  double *dd = zJ + 8;                     // [Mp]  damping ep + lm * S_rr, added when a diagonal tile is factored
  double *Lst = dd + Mp;                   // [kBackStages][8][128 + 64] back-substitution stages: 8 rows of L + W_J
  unsigned *tabU = reinterpret_cast<unsigned *>(Lst + kBackStages * (8 * 128 + 64));   // [16][136] operand offsets of [U]
  unsigned *tabP = tabU + 16 * 136;        // [16][8] which (two) of a warp's 17 tiles touch position e
  unsigned *tabX = tabP + 16 * 8;          // [16][8] the other position of those two tiles, 8 bits each
  double *xs = reinterpret_cast<double *>(tabX + 16 * 8);          // [8] x_J of the back substitution
  __shared__ int s_fail, s_nan;
  const double *__restrict__ S = cv.S;
  double *__restrict__ L = cv.L;
  const double ep = (double)cv.ep;
  auto Sg = [&](int r, int c) { return (size_t)r * ld + c + off; };
  int status = 0;

  // ---- step tables: which operand tiles every register tile needs when position e leaves the window ----
  for (int o = tau; o < 16 * 136; o += kMmaThreads) {
    const int e = o / 136, idx = o - e * 136;
    const int x = c_px[idx], y = c_py[idx];
    unsigned offA = 16 * kTs, offB = 16 * kTs;                      // inactive tile: both operands = the zero tile
    if (x != e && y != e) {
      const int ax = (x - e) & 15, ay = (y - e) & 15;
      offA = (ax > ay ? x : y) * kTs;                               // row tile = larger global index
      offB = (ax > ay ? y : x) * kTs;
    }
    tabU[o] = offA | (offB << 16);
  }
  for (int o = tau; o < 16 * 8; o += kMmaThreads) {
    const int e = o >> 3, w = o & 7;
    unsigned m = 0, xo = 0;
    int k = 0;
    for (int t = 0; t < kTilesPerWarp; ++t) {
      const int x = c_px[t * 8 + w], y = c_py[t * 8 + w];
      if (x == e || y == e) { m |= 1u << t; xo |= (unsigned)(x == e ? y : x) << (8 * k++); }
    }
    tabP[o] = m;
    tabX[o] = xo;
  }
  for (int o = tau; o < kTs; o += kMmaThreads) { Psm[16 * kTs + o] = 0.0; Nsm[16 * kTs + o] = 0.0; }

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    // value of S at (r, c), r >= c, with identity padding beyond M. The damping of the diagonal
    // (ba.py:67) is added from `dd` when the diagonal tile is factored, so that nothing here consumes the
    // loaded value and the global latency of a refill hides behind the rest of the step.
    auto Aval = [&](int r, int c) -> double {
      if (r >= M) return r == c ? 1.0 : 0.0;
      if (c > r || r - c > bw) return 0.0;
      return S[Sg(r, c)];
    };
    auto load_tile = [&](int a, int b, double &c0, double &c1) {   // tile (a, b), a >= b, C-fragment layout
      const int r = 8 * a + g, c = 8 * b + 2 * q;
      c0 = Aval(r, c);
      c1 = Aval(r, c + 1);
    };

    for (int r = tau; r < Mp; r += kMmaThreads) {
      z[r] = r < M ? cv.y[r] : 0.0;
      dd[r] = r < M ? ep + lm * S[Sg(r, r)] : 0.0;                 // A = S + (ep + lm * S) .* I, ba.py:67
    }
    if (tau == 0) { s_fail = 0; s_nan = 0; }
    __syncthreads();                                               // z, dd, tables, zero tiles visible
    bool failed = false;
    if (warp == kMmaWarps) {
      // =================== factor warp: [A] for column J while the tile warps still update column J-1 ==========
      for (int J = 0; J < NT8; ++J) {
        bar_sync(2, 64);                                           // tile (J,J) (+ damping) is in Dsm, z_J is final
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
#pragma unroll
          for (int i = k + 1; i < 8; ++i) a[tri8(i, k)] *= inv;
#pragma unroll
          for (int j = k + 1; j < 8; ++j)
#pragma unroll
            for (int i = j; i < 8; ++i) a[tri8(i, j)] -= a[tri8(i, k)] * a[tri8(j, k)];
        }
        if (ok) {
          // (L_JJ itself is never needed again: the back substitution uses W_J.)
          // One forward substitution per lane, same instruction stream, different right-hand side:
          // lanes 0..7 solve L_JJ w = e_lane (column `lane` of W = L_JJ^-1), lane 8 solves L_JJ zJ = z_J.
          double wv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            double sv = lane == 8 ? z[8 * J + i] : (lane == i ? 1.0 : 0.0);
#pragma unroll
            for (int j = 0; j < i; ++j) sv -= a[tri8(i, j)] * wv[j];
            wv[i] = sv * invd[i];
          }
          if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) Wsm[i * kPs + lane] = wv[i];
          } else if (lane == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { z[8 * J + i] = wv[i]; zJ[i] = wv[i]; }
          }
        } else if (lane == 0) {
          s_fail = 1;
        }
        bar_arrive(3, kMmaThreads);                                // W_J, zJ (or the failure flag) published
        if (!ok) { failed = true; break; }
      }
    } else {
      // =================== tile warps ===========================================================================
      double ct[kTilesPerWarp][2];
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];   // x <= y: initial window holds tile (y, x)
        ct[t][0] = ct[t][1] = 0.0;
        if (y < NT8) load_tile(y, x, ct[t][0], ct[t][1]);
      }
      if (warp == 0) {                                              // tile (0,0) is tile 0 of warp 0
        const double dmp = dd[g];
        Dsm[g * kPs + 2 * q] = ct[0][0] + (2 * q == g ? dmp : 0.0);
        Dsm[g * kPs + 2 * q + 1] = ct[0][1] + (2 * q + 1 == g ? dmp : 0.0);
        bar_arrive(2, 64);
      }
      // The two tiles of this warp that touch the retiring position e are worked on in fixed registers
      // (et) so that the unrolled code can interleave them; their refills (rf) are loaded during [P] and
      // merged back into the tile registers one step later, when the loads have long landed.
      double rf[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      unsigned pm_prev = 0;

      for (int J = 0; J < NT8; ++J) {
        const int e = J & 15;
        const unsigned pm = tabP[e * 8 + warp], xo2 = tabX[e * 8 + warp];
        double et[2][2];
        {
          int kp = 0, kc = 0;
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            if ((pm_prev >> t) & 1u) { ct[t][0] = kp ? rf[1][0] : rf[0][0]; ct[t][1] = kp ? rf[1][1] : rf[0][1]; ++kp; }
            if ((pm >> t) & 1u) {
              if (kc == 0) { et[0][0] = ct[t][0]; et[0][1] = ct[t][1]; } else { et[1][0] = ct[t][0]; et[1][1] = ct[t][1]; }
              ++kc;
            }
          }
        }
        pm_prev = pm;
        const int xo0 = xo2 & 0xff, xo1 = (xo2 >> 8) & 0xff;
        bar_sync(3, kMmaThreads);                                  // W_J, zJ ready; every tile warp is past [U](J-1)
        if (s_fail) { failed = true; break; }
        // ---- [P] panel tiles: L_aJ = A_aJ W^T; both e-tiles in one straight-line block. A tile below the
        //      matrix is all zeros and yields zeros; the diagonal tile only skips its stores. ----
        if (warp == kMmaWarps - 1) {                                // W_J -> global for the back substitution
          Wg[(size_t)J * 64 + lane] = Wsm[(lane >> 3) * kPs + (lane & 7)];
          Wg[(size_t)J * 64 + 32 + lane] = Wsm[(4 + (lane >> 3)) * kPs + (lane & 7)];
        }
        {
          const double wb0 = Wsm[g * kPs + q], wb1 = Wsm[g * kPs + 4 + q];   // B[k][n] = W[n][k], n = g, k = 4s + q
          const double zq0 = zJ[2 * q], zq1 = zJ[2 * q + 1];
          const int lo = g * kPs + 2 * q;
          const int src0 = (lane & ~3) | (q >> 1), src1 = src0 + 2;
          const int cJ = 8 * J + 2 * q;
          const int an = J + 16;
          double p[2][2], part[2];
          int xo[2] = {xo0, xo1}, ag[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            ag[k] = J + ((xo[k] - e) & 15);                        // global tile row held at position xo
            // C fragment (cols 2q, 2q+1 of row g) -> A fragments (col 4s + q of row g)
            const double v00 = __shfl_sync(0xffffffffu, et[k][0], src0), v01 = __shfl_sync(0xffffffffu, et[k][1], src0);
            const double v10 = __shfl_sync(0xffffffffu, et[k][0], src1), v11 = __shfl_sync(0xffffffffu, et[k][1], src1);
            const double a0 = (q & 1) ? v01 : v00, a1 = (q & 1) ? v11 : v10;
            dmma884(p[k][0], p[k][1], a0, wb0, 0.0, 0.0);
            dmma884(p[k][0], p[k][1], a1, wb1, p[k][0], p[k][1]);
            part[k] = p[k][0] * zq0 + p[k][1] * zq1;               // right-hand side: z_a -= L_aJ zJ
            part[k] += __shfl_xor_sync(0xffffffffu, part[k], 1);
            part[k] += __shfl_xor_sync(0xffffffffu, part[k], 2);
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const bool panel = xo[k] != e;
            const int r = 8 * ag[k] + g;
            if (panel) {
              Psm[xo[k] * kTs + lo] = p[k][0]; Psm[xo[k] * kTs + lo + 1] = p[k][1];
              Nsm[xo[k] * kTs + lo] = -p[k][0]; Nsm[xo[k] * kTs + lo + 1] = -p[k][1];
              if (q == 0 && r < Mp) z[r] -= part[k];
              if (r < M) {                                         // columns of tile J are < M whenever a row below is
                double *lp = L + Sg(r, cJ);
                if (r - cJ <= bw) lp[0] = p[k][0];
                if (r - cJ - 1 <= bw) lp[1] = p[k][1];
              }
            }
            // refill: position e now stands for tile index J + 16
            rf[k][0] = rf[k][1] = 0.0;
            if (an < NT8) load_tile(an, panel ? ag[k] : an, rf[k][0], rf[k][1]);
          }
        }
        bar_sync(1, 32 * kMmaWarps);                               // all panel tiles (and z updates) of column J done
        // ---- [U] trailing update: C_ab -= L_aJ L_bJ^T, branch-free; operand offsets from the step table
        //      (e-tiles, whose registers are stale until the refill is merged, and tiles below the matrix read
        //      zeros). Look-ahead: the owner of tile (J+1,J+1) updates it first and hands it to the factor
        //      warp, so that the 8x8 factorisation of the next column overlaps the rest of this update. ----
        {
          const unsigned *tu = tabU + e * 136 + warp;
          const double *An = Nsm + g * kPs + q, *Bp = Psm + g * kPs + q;
          const int en = (J + 1) & 15;
          const bool own_next = (J + 1 < NT8) && warp == (en & 7);
          const int tn = en >> 3;                                  // pair {en,en} is tile en/8 of warp en%8
          if (own_next) {
            const unsigned o = tu[tn * 8];
            const double *A = An + (o & 0xffffu), *B = Bp + (o >> 16);
            double c0 = tn ? ct[1][0] : ct[0][0], c1 = tn ? ct[1][1] : ct[0][1];
            dmma884(c0, c1, A[0], B[0], c0, c1);
            dmma884(c0, c1, A[4], B[4], c0, c1);
            if (tn) { ct[1][0] = c0; ct[1][1] = c1; } else { ct[0][0] = c0; ct[0][1] = c1; }
            const double dmp = dd[8 * (J + 1) + g];
            Dsm[g * kPs + 2 * q] = c0 + (2 * q == g ? dmp : 0.0);
            Dsm[g * kPs + 2 * q + 1] = c1 + (2 * q + 1 == g ? dmp : 0.0);
            bar_arrive(2, 64);
          }
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            unsigned o = tu[t * 8];
            if (t < 2 && own_next && t == tn) o = (16u * kTs) | ((16u * kTs) << 16);    // already applied above
            const double *A = An + (o & 0xffffu), *B = Bp + (o >> 16);
            dmma884(ct[t][0], ct[t][1], A[0], B[0], ct[t][0], ct[t][1]);
            dmma884(ct[t][0], ct[t][1], A[4], B[4], ct[t][0], ct[t][1]);
          }
        }
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
    // kBackStages - 1 stages in flight: one (possibly empty) cp.async group per tile row
    for (int k = 0; k < kBackStages - 1; ++k) {
      if (NT8 - 1 - k >= 0) stage_load(NT8 - 1 - k, (NT8 - 1 - k) % kBackStages);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int J = NT8 - 1; J >= 0; --J) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kBackStages - 2) : "memory");
      __syncthreads();                                             // stage J landed; iteration J+1 fully retired
      if (J - (kBackStages - 1) >= 0) stage_load(J - (kBackStages - 1), (J - (kBackStages - 1)) % kBackStages);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      const double *st = Lst + (size_t)(J % kBackStages) * (8 * 128 + 64);
      const double *Wj = st + 8 * 128;                             // W_J row-major 8x8
      // x_J = W_J^T z_J : 8 threads, one component each (two partial sums to halve the FMA chain)
      if (tau < 8) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          s0 += (k >= tau ? Wj[k * 8 + tau] : 0.0) * z[8 * J + k];
          s1 += (k + 1 >= tau ? Wj[(k + 1) * 8 + tau] : 0.0) * z[8 * J + k + 1];
        }
        __syncwarp(0xffu);                                         // all 8 lanes have read z_J
        xs[tau] = s0 + s1;
        z[8 * J + tau] = s0 + s1;
      }
      __syncthreads();
      if (tau >= 32 && tau < 32 + 120) {
        const int xcol = tau - 32, c = 8 * (J - 15) + xcol;
        if (c >= 0) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int gg = 0; gg < 8; gg += 2) { s0 += st[gg * 128 + xcol] * xs[gg]; s1 += st[(gg + 1) * 128 + xcol] * xs[gg + 1]; }
          z[c] -= s0 + s1;
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
  return ((size_t)2 * Mp + 2 * 17 * kTs + 2 * kTs + 8 + kBackStages * (8 * 128 + 64)) * sizeof(double) +
         (16 * 136 + 2 * 16 * 8) * sizeof(unsigned) + 8 * sizeof(double);
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
