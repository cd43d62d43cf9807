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
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ba_internal.h"

namespace ba {

constexpr int kMmaWarps = 8;                 // tile warps
// 9 hardware warps: 0..7 = tile warps (two per scheduler: the DMMAs of [U] occupy a scheduler's FP64 unit for
// 16 cycles each, 68 per scheduler and column, and that unit is what bounds the step), 8 = factor warp.
// (Measured alternative: factor warp alone on scheduler 0 and 3+3+2 tile warps on the others: the factorisation
// drops to 1.8k cycles but [U] rises to 3.1k, slower overall.)
constexpr int kMmaThreads = 32 * (kMmaWarps + 1);
// "isolated" launch shape: 12 hardware warps, of which warp 0 (alone on scheduler 0) is the factor warp, warps
// 1-3, 5-7, 9-10 are the tile warps (3 + 3 + 2 on schedulers 1-3) and warps 4, 8, 11 exit at once — the dependent
// DFMA chain of the 8x8 factorisation then never queues behind a 16-cycle DMMA on its scheduler's FP64 unit.
constexpr int kMmaHwThreads = 32 * 12;
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
  // not volatile: pure function of its operands, so the compiler may interleave the DMMAs of independent tiles
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
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
// (owner warp -> factor warp), 3 = W_J / zJ ready (factor warp -> tile warps; also closes the previous [U]),
// 4 = panel tile L_{J+1,J} stored (its producer -> owner of tile (J+1,J+1))
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void cp_async8_d(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

constexpr int tri8(int a, int b) { return a * (a + 1) / 2 + b; }

// optional phase trace (BA_TRACE=1): clock64 stamps of tile warp 0 and of the factor warp per tile column
#define BA_TR(slot) do { if (trace && lane == 0) trace[(size_t)J * 16 + (slot)] = clock64(); } while (0)
// the same, but the clock is read only once `dep` is available (barrier waits are deferred to the first dependent
// instruction, so a plain clock read right after bar.sync would not see them)
__device__ __forceinline__ long long clk_after(double dep) { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep)); return t; }
__device__ __forceinline__ long long clk_after(int dep) { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep)); return t; }
#define BA_TRD(slot, dep) do { if (trace && lane == 0) trace[(size_t)J * 16 + (slot)] = clk_after(dep); } while (0)

// Step tables of the solver (which operand tiles every register tile needs when circular position e leaves the
// window, which two tiles of a warp touch e, and the other position of those tiles). They depend on nothing but
// the tile ownership, so they are built once per process into global memory and copied to shared memory by
// every launch. Layout: tabU [16][136], tabP [16][8], tabX [16][8].
__global__ void k_build_step_tables(unsigned *gtab) {
  const int tau = threadIdx.x;
  constexpr int kMmaThreadsLocal = 256;
  unsigned *tabU = gtab, *tabP = gtab + 16 * 136, *tabX = tabP + 16 * 8;
  for (int o = tau; o < 16 * 136; o += kMmaThreadsLocal) {
    const int e = o / 136, idx = o - e * 136;
    const int x = c_px[idx], y = c_py[idx];
    unsigned offA = 16 * kTs, offB = 16 * kTs;                      // inactive tile: both operands = the zero tile
    if (x != e && y != e) {
      const int ax = (x - e) & 15, ay = (y - e) & 15;
      offA = (ax > ay ? x : y) * kTs;                               // row tile = larger global index
      offB = (ax > ay ? y : x) * kTs;
    }
    // bits 0-11 offA, 12-23 offB, 24 this pair touches e (+25: which of the warp's two), 26 it touches e+1 (+27)
    const int w = idx & 7, en = (e + 1) & 15;
    unsigned fl = 0;
    int ke = 0, kn = 0;
    for (int t2 = 0; t2 < idx / 8; ++t2) {
      const int x2 = c_px[t2 * 8 + w], y2 = c_py[t2 * 8 + w];
      ke += (x2 == e || y2 == e);
      kn += (x2 == en || y2 == en);
    }
    if (x == e || y == e) fl |= 1u | ((unsigned)ke << 1);
    if (x == en || y == en) fl |= 4u | ((unsigned)kn << 3);
    tabU[o] = offA | (offB << 12) | (fl << 24);
  }
  for (int o = tau; o < 16 * 8; o += kMmaThreadsLocal) {
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
}

// cluster-wide barrier (both CTAs of the twisted factorisation); release/acquire orders global memory
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// twist != 0: launched as a cluster of two CTAs. CTA 0 eliminates tile columns [0, Jm0) of the matrix, CTA 1
// the last Jm1 tile columns, working on the index-reversed matrix (same code, reversed coordinates); CTA 1
// then hands CTA 0 what its eliminations contributed to the 16 middle tile columns, CTA 0 finishes the
// middle, solves it, and both back-substitute their side in parallel ("burn at both ends": the serial chain
// of a banded Cholesky is halved). Exchange through global scratch XD + cluster barriers.
__global__ void __launch_bounds__(kMmaHwThreads, 1) k_solve_band_mma(CallView cv, int allow_retry, double *__restrict__ Wg_all,
                                                                   double *__restrict__ L_all, double *__restrict__ XD,
                                                                   int *__restrict__ gfl, int twist,
                                                                   const unsigned *__restrict__ gtab,
                                                                   long long *__restrict__ trace) {
  extern __shared__ double dsm[];
  const int lane = threadIdx.x & 31;
  int hw = threadIdx.x >> 5;                                        // logical warp: 0..7 tile warps, 8 factor warp
  if (blockDim.x == kMmaHwThreads) {
    hw = (int)(0xf76f543f2108ull >> (4 * hw)) & 15;                        // {8,0,1,2,-,3,4,5,-,6,7,-}
    if (hw == 15) return;
  }
  const int tau = 32 * hw + lane;
  const bool is_factor = hw == kMmaWarps, is_tile = hw < kMmaWarps;
  const int warp = hw;                                              // tile warp index 0..7 (meaningful if is_tile)
  const int g = lane >> 2, q = lane & 3;
  const int M = cv.M, bw = cv.bw, ld = cv.ld, off = cv.off;
  const int NT8 = (M + 7) >> 3, Mp = NT8 * 8;
  const int side = twist ? (int)blockIdx.x : 0;
  // columns eliminated from the top / from the bottom: equal shares, because the 16 middle columns can only start
  // once BOTH sides are done (measured: giving side 0 fewer columns to balance its extra middle work is slower)
  const int Jm0 = (NT8 - 16) / 2, Jm1 = NT8 - 16 - Jm0;
  const int c1 = twist ? (side ? Jm1 : Jm0) : NT8;                  // end of this side's first segment
  const int NTloc = twist ? c1 + 16 : NT8;                          // tiles this side ever sees (local coordinates)
  const int nseg = (twist && side == 0) ? 2 : 1;
  double *__restrict__ Wg = Wg_all + (size_t)side * NT8 * 64;
  double *__restrict__ L = L_all + (size_t)side * Mp * (bw + 1);    // this side's factor, local band storage
  auto Lg = [&](int rl, int cl) { return (size_t)rl * bw + cl + bw; };
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
  double *Esm = reinterpret_cast<double *>(tabX + 16 * 8);         // [8 warps][2 slots][64] next column's e-tiles
  double *xsol = dd;                                               // solution of the back substitution (dd is dead then)
  __shared__ int s_fail, s_nan, s_abort;
  const double *__restrict__ S = cv.S;
  const double ep = (double)cv.ep;
  auto Sg = [&](int r, int c) { return (size_t)r * ld + c + off; };
  int status = 0;
  long long *phase = trace ? trace + 16 * 4096 + side * 8 : nullptr;   // coarse phase stamps of this side (BA_TRACE)
  if (side) trace = nullptr;                                           // per-column stamps: side 0 only
  if (phase && tau == 0) phase[0] = clock64();

  // ---- step tables (static, built once per process by k_build_step_tables): global -> shared ----
  for (int o = tau; o < 16 * 136 + 2 * 16 * 8; o += kMmaThreads) tabU[o] = gtab[o];
  for (int o = tau; o < kTs; o += kMmaThreads) { Psm[16 * kTs + o] = 0.0; Nsm[16 * kTs + o] = 0.0; }

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    // value of S at (r, c), r >= c, with identity padding beyond M. The damping of the diagonal
    // (ba.py:67) is added from `dd` when the diagonal tile is factored, so that nothing here consumes the
    // loaded value and the global latency of a refill hides behind the rest of the step.
    auto Aval = [&](int rl, int cl) -> double {                  // local coordinates (reversed on side 1)
      if (cl > rl) return 0.0;
      const int r = side ? Mp - 1 - cl : rl, c = side ? Mp - 1 - rl : cl;   // global, r >= c
      if (r >= M) return r == c ? 1.0 : 0.0;
      if (r - c > bw) return 0.0;
      return S[Sg(r, c)];
    };
    auto load_tile = [&](int a, int b, double &c0, double &c1) {   // tile (a, b), a >= b, C-fragment layout
      const int r = 8 * a + g, c = 8 * b + 2 * q;
      c0 = Aval(r, c);
      c1 = Aval(r, c + 1);
    };

    for (int rl = tau; rl < Mp; rl += kMmaThreads) {
      const int r = side ? Mp - 1 - rl : rl;
      z[rl] = r < M ? cv.y[r] : 0.0;
      dd[rl] = r < M ? ep + lm * S[Sg(r, r)] : 0.0;                // A = S + (ep + lm * S) .* I, ba.py:67
    }
    if (tau == 0) { s_fail = 0; s_nan = 0; s_abort = 0; }
    int *gf = gfl + 4 * attempt;                                   // [0] side 0 failed, [1] side 1 failed, [2] NaN
    __syncthreads();                                               // z, dd, tables, zero tiles visible
    bool failed = false;
    const int nsegs = twist ? 2 : 1;
    if (is_factor) {
      // =================== factor warp: [A] for column J while the tile warps still update column J-1 ==========
      for (int seg = 0; seg < nsegs; ++seg) {
      if (seg == 1) {                                              // twist hand-over (see the tile-warp branch)
        __syncthreads();
        cluster_sync();
        __syncthreads();
      }
      const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
      for (int J = jb; J < je; ++J) {
        BA_TR(8);
        bar_sync(2, 64);                                           // tile (J,J) (+ damping) is in Dsm, z_J is final
        // every lane factors the 8x8 block redundantly in registers (no divergence, no extra exchange)
        double a[36];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) a[tri8(i, j)] = Dsm[i * kPs + j];
        BA_TRD(9, a[35]);
        // Right-looking 8x8 Cholesky fused with the forward substitutions: one right-hand side per lane, same
        // instruction stream — lanes 0..7 solve L_JJ w = e_lane (column `lane` of W = L_JJ^-1), lane 8 solves
        // L_JJ zJ = z_J. wv[k] is formed right after pivot k, so only (one FMA + one multiply) per row sits on
        // the dependent chain. (L_JJ itself is never needed again: the back substitution uses W_J.)
        bool ok = true;
        double wv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double piv = a[tri8(k, k)];
          const float pf = (float)piv;                             // range / sign / NaN test on the fp32 copy
          ok = ok && (pf > 0.0f);                                  // potrf info != 0 (also NaN), ba.py:11
          const bool fast = pf > 1e-30f && pf < 1e30f;
          double inv;
          if (fast) {                                              // fp32 seed (~2e-7) + one Newton step -> ~6e-14
            const double y = (double)rsqrtf(pf);
            inv = fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
          } else {
            inv = 1.0 / sqrt(piv);
          }
          double sv = lane == 8 ? z[8 * J + k] : (lane == k ? 1.0 : 0.0);
#pragma unroll
          for (int j = 0; j < k; ++j) sv -= a[tri8(k, j)] * wv[j];
          wv[k] = sv * inv;
#pragma unroll
          for (int i = k + 1; i < 8; ++i) a[tri8(i, k)] *= inv;
#pragma unroll
          for (int j = k + 1; j < 8; ++j)
#pragma unroll
            for (int i = j; i < 8; ++i) a[tri8(i, j)] -= a[tri8(i, k)] * a[tri8(j, k)];
        }
        if (ok) {
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
        BA_TR(10);
        bar_arrive(3, 32 * (kMmaWarps + 1));                       // W_J, zJ (or the failure flag) published
        if (!ok) { failed = true; break; }
      }
      }
    } else if (is_tile) {
      // =================== tile warps ===========================================================================
      double ct[kTilesPerWarp][2];
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];   // x <= y: initial window holds tile (y, x)
        ct[t][0] = ct[t][1] = 0.0;
        if (y < NTloc) load_tile(y, x, ct[t][0], ct[t][1]);
      }
      double *myE = Esm + warp * 128 + 2 * lane;                    // this warp's two e-tile slots (fragment layout)
      {                                                             // e-tiles of column 0 = "next" tiles of position 15
        const unsigned *tu = tabU + 15 * 136 + warp;
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          const unsigned o = tu[t * 8];
          if (o & (4u << 24)) { double *d = myE + ((o >> 27) & 1u) * 64; d[0] = ct[t][0]; d[1] = ct[t][1]; }
        }
      }
      if (warp == 0) {                                              // tile (0,0) is tile 0 of tile warp 0
        const double dmp = dd[g];
        Dsm[g * kPs + 2 * q] = ct[0][0] + (2 * q == g ? dmp : 0.0);
        Dsm[g * kPs + 2 * q + 1] = ct[0][1] + (2 * q + 1 == g ? dmp : 0.0);
        bar_arrive(2, 64);
      }
      __syncwarp();
      // Per column J, the two tiles of this warp that touch the retiring position e ("e-tiles") are read from
      // the warp's shared slots (written by the previous [U]) into fixed registers, so the unrolled code never
      // indexes the tile registers dynamically. Their refills are loaded into fixed registers at the top of the
      // step and moved into the tile registers by predicated moves inside the [U] loop.
      for (int seg = 0; seg < nsegs; ++seg) {
      if (seg == 1) {
        // ---- twist hand-over. Side 1: what its eliminations did to the 16 middle tile columns (window minus
        //      the untouched matrix) and to the right-hand side goes to XD in its local coordinates. Side 0 adds
        //      it to its window, refreshes the e-tile slots and publishes the diagonal tile of column c1. ----
        __syncthreads();
        const int ec = c1 & 15;
        if (side == 1 && !s_fail) {
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];
            const int ax = c1 + ((x - ec) & 15), ay = c1 + ((y - ec) & 15);
            const int rl = 8 * max(ax, ay) + g, cl = 8 * min(ax, ay) + 2 * q;
            double *d = XD + (size_t)(rl - 8 * c1) * 128 + (cl - 8 * c1);
            d[0] = ct[t][0] - Aval(rl, cl);
            d[1] = ct[t][1] - Aval(rl, cl + 1);
          }
          const int i = warp * 32 + lane;
          if (i < 128) { const int r = Mp - 1 - (8 * c1 + i); XD[16384 + i] = z[8 * c1 + i] - (r < M ? cv.y[r] : 0.0); }
        }
        if (tau == 0) gf[side] = s_fail;
        cluster_sync();
        if (tau == 0) s_abort = side == 0 ? (gf[1] | s_fail) : s_fail;
        __syncthreads();
        if (side == 0 && !s_abort) {
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int x = c_px[t * 8 + warp], y = c_py[t * 8 + warp];
            const int ax = c1 + ((x - ec) & 15), ay = c1 + ((y - ec) & 15);
            const int r = 8 * max(ax, ay) + g, c = 8 * min(ax, ay) + 2 * q;
            const double *d = XD + (size_t)(Mp - 1 - c - 8 * Jm1) * 128 + (Mp - 1 - r - 8 * Jm1);
            ct[t][0] += d[0];
            ct[t][1] += d[-128];
          }
          const int i = warp * 32 + lane;
          if (i < 128) z[8 * c1 + i] += XD[16384 + 127 - i];
          {
            const unsigned *tu = tabU + ((ec - 1) & 15) * 136 + warp;
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) {
              const unsigned o = tu[t * 8];
              if (o & (4u << 24)) { double *d = myE + ((o >> 27) & 1u) * 64; d[0] = ct[t][0]; d[1] = ct[t][1]; }
            }
          }
          bar_sync(1, 32 * kMmaWarps);                             // z of the middle complete before the factor warp reads it
          if (warp == (ec & 7)) {
            const double dmp = dd[8 * c1 + g];
            const double d0 = (ec >> 3) ? ct[1][0] : ct[0][0], d1 = (ec >> 3) ? ct[1][1] : ct[0][1];
            Dsm[g * kPs + 2 * q] = d0 + (2 * q == g ? dmp : 0.0);
            Dsm[g * kPs + 2 * q + 1] = d1 + (2 * q + 1 == g ? dmp : 0.0);
            bar_arrive(2, 64);
          }
          __syncwarp();
        }
      }
      const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
      for (int J = jb; J < je; ++J) {
        const int e = J & 15;
        const unsigned xo2 = tabX[e * 8 + warp];
        const int xo0 = xo2 & 0xff, xo1 = (xo2 >> 8) & 0xff;
        if (warp == 0) BA_TR(0);
        double et[2][2], rf[2][2];
        et[0][0] = myE[0]; et[0][1] = myE[1]; et[1][0] = myE[64]; et[1][1] = myE[65];
        int ag[2];
        {
          const int an = J + 16;
          const int xo[2] = {xo0, xo1};
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            ag[k] = J + ((xo[k] - e) & 15);                        // global tile row held at the other position
            rf[k][0] = rf[k][1] = 0.0;                             // position e next stands for tile index J + 16
            if (an < NTloc) load_tile(an, xo[k] == e ? an : ag[k], rf[k][0], rf[k][1]);
          }
        }
        if (warp == 0) BA_TR(1);
        bar_sync(3, 32 * (kMmaWarps + 1));                         // W_J, zJ ready; every tile warp is past [U](J-1)
        const int sf = s_fail;
        if (warp == 0) BA_TRD(2, sf);
        if (sf) { failed = true; break; }
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
          double p[2][2], part[2];
          const int xo[2] = {xo0, xo1};
#pragma unroll
          for (int k = 0; k < 2; ++k) {
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
            const int r = 8 * ag[k] + g;
            if (xo[k] != e) {
              Psm[xo[k] * kTs + lo] = p[k][0]; Psm[xo[k] * kTs + lo + 1] = p[k][1];
              Nsm[xo[k] * kTs + lo] = -p[k][0]; Nsm[xo[k] * kTs + lo + 1] = -p[k][1];
              if (r < 8 * NTloc) {                                 // tiles below this side's matrix are all zero
                if (q == 0) z[r] -= part[k];
                double *lp = L + Lg(r, cJ);
                if (r - cJ <= bw) lp[0] = p[k][0];
                if (r - cJ - 1 <= bw) lp[1] = p[k][1];
              }
            }
          }
        }
        // ---- look-ahead: tile (J+1,J+1) only needs L_{J+1,J}. The warp that produced that panel tile signals
        //      the owner of (J+1,J+1) directly (barrier 4), which updates its tile, adds the damping and hands it
        //      to the factor warp (barrier 2) — the 8x8 factorisation of column J+1 then overlaps the rest of
        //      [P] and all of [U] of column J. ----
        const unsigned *tu = tabU + e * 136 + warp;
        const double *An = Nsm + g * kPs + q, *Bp = Psm + g * kPs + q;
        const int en = (J + 1) & 15;
        const bool have_next = J + 1 < je;
        const bool own_next = have_next && warp == (en & 7);
        const bool is_prod = have_next && (xo0 == en || xo1 == en);
        const int tn = en >> 3;                                    // pair {en,en} is tile en/8 of tile warp en%8
        if (is_prod && !own_next) bar_arrive(4, 64);
        if (own_next) {
          if (is_prod) __syncwarp(); else bar_sync(4, 64);
          const unsigned o = tu[tn * 8];
          const double *A = An + (o & 0xfffu), *B = Bp + ((o >> 12) & 0xfffu);
          double c0 = tn ? ct[1][0] : ct[0][0], c1 = tn ? ct[1][1] : ct[0][1];
          dmma884(c0, c1, A[0], B[0], c0, c1);
          dmma884(c0, c1, A[4], B[4], c0, c1);
          if (tn) { ct[1][0] = c0; ct[1][1] = c1; } else { ct[0][0] = c0; ct[0][1] = c1; }
          const double dmp = dd[8 * (J + 1) + g];
          Dsm[g * kPs + 2 * q] = c0 + (2 * q == g ? dmp : 0.0);
          Dsm[g * kPs + 2 * q + 1] = c1 + (2 * q + 1 == g ? dmp : 0.0);
          bar_arrive(2, 64);
        }
        if (warp == 0) BA_TR(3);
        bar_sync(1, 32 * kMmaWarps);                               // all panel tiles (and z updates) of column J done
        if (warp == 0) BA_TRD(4, Psm[16 * kTs]);
        // ---- [U] trailing update: C_ab -= L_aJ L_bJ^T, branch-free; operand offsets and flags from the step
        //      table (e-tiles and tiles below the matrix read zeros). Inside the loop, an e-tile takes its refill,
        //      and a tile that touches position e+1 is copied to the warp's shared slots for the next column. ----
        {
          unsigned ot[kTilesPerWarp];
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) ot[t] = tu[t * 8];
          {                                                         // the look-ahead tile was already updated above
            const unsigned zo = (16u * kTs) | ((16u * kTs) << 12);
            if (own_next && tn == 0) ot[0] = (ot[0] & 0xff000000u) | zo;
            if (own_next && tn == 1) ot[1] = (ot[1] & 0xff000000u) | zo;
          }
          // two passes (k = 0..3, then k = 4..7): the 17 DMMAs of a pass are independent of each other, so the warp
          // never sits on the 26-cycle DMMA -> DMMA dependency of one tile, and the operand loads of a pass are
          // issued together ahead of its DMMAs
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double fa[kTilesPerWarp], fb[kTilesPerWarp];
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) {
              const unsigned o = ot[t];
              fa[t] = An[(o & 0xfffu) + 4 * h];
              fb[t] = Bp[((o >> 12) & 0xfffu) + 4 * h];
            }
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) dmma884(ct[t][0], ct[t][1], fa[t], fb[t], ct[t][0], ct[t][1]);
          }
          // an e-tile's registers take its refill; a tile that touches position e+1 goes to the warp's slots
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const unsigned fl = ot[t] >> 24;
            const bool ise = fl & 1u, sel = fl & 2u;
            ct[t][0] = ise ? (sel ? rf[1][0] : rf[0][0]) : ct[t][0];
            ct[t][1] = ise ? (sel ? rf[1][1] : rf[0][1]) : ct[t][1];
            if (fl & 4u) { double *d = myE + ((fl >> 3) & 1u) * 64; d[0] = ct[t][0]; d[1] = ct[t][1]; }
          }
        }
        if (warp == 0) BA_TR(5);
        __syncwarp();                                              // slots written by this warp, read by it next step
      }
      }
    }
    __syncthreads();
    (void)failed;
    if (phase && tau == 0) phase[1] = clock64();
    const bool bad = s_fail || s_abort;                             // this side could not factor (or was told to stop)
    const int ncols = (twist && side == 1) ? c1 : NTloc;            // tile columns of L (and W_J) this side owns

    // ---- backward substitution L^T x = z by tile rows, descending, in local coordinates. Stage s holds rows
    //      8J..8J+7 of L restricted to columns [8(J-15), 8J) plus W_J. On side 1 the middle rows J >= c1 were not
    //      factored here: their solution is given (x_mid from side 0) and only their coupling to the columns this
    //      side owns (which it did compute) is applied. ----
    // Each thread copies the same (row-in-tile, column) elements of every stage; their source moves by a fixed
    // stride per tile row, so everything but the two edge tests is computed once here.
    int sl_kind[4], sl_doff[4], sl_xc[4];                          // 0 none, 1 band element, 2 always zero, 3 element of W_J
    long long sl_src[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o = tau + k * kMmaThreads;
      sl_kind[k] = 0; sl_doff[k] = 0; sl_xc[k] = 0; sl_src[k] = 0;
      if (o < 8 * 120) {
        const int gg = o / 120, xcol = o - gg * 120;
        sl_doff[k] = gg * 128 + xcol;
        sl_xc[k] = xcol - 120;                                     // column = 8 J + sl_xc
        sl_kind[k] = (120 + gg - xcol <= bw) ? 1 : 2;              // r - c does not depend on J
        sl_src[k] = (long long)gg * bw + (xcol - 120) + bw;        // Lg(8J + gg, 8J + xcol - 120) - J (8 bw + 8)
      } else if (o < 8 * 120 + 64) {
        sl_kind[k] = 3; sl_doff[k] = 8 * 128 + (o - 8 * 120); sl_src[k] = o - 8 * 120;
      }
    }
    auto stage_load = [&](int J, int sidx) {
      double *dst = Lst + (size_t)sidx * (8 * 128 + 64);
      const double *Lj = L + (long long)J * (8 * bw + 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (sl_kind[k] == 1) {
          const int c = 8 * J + sl_xc[k];
          if (c >= 0 && c < 8 * ncols) cp_async8_d(dst + sl_doff[k], Lj + sl_src[k]);
          else dst[sl_doff[k]] = 0.0;
        } else if (sl_kind[k] == 2) {
          dst[sl_doff[k]] = 0.0;
        } else if (sl_kind[k] == 3) {
          if (J < ncols) cp_async8_d(dst + sl_doff[k], Wg + (size_t)J * 64 + sl_src[k]);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __threadfence_block();
    bool anybad = bad;
    if (twist && side == 1) {                                      // wait for x_mid
      cluster_sync();
      anybad = bad || gf[0] != 0 || gf[1] != 0;
      if (!anybad) {
        for (int i = tau; i < 128; i += kMmaThreads) xsol[8 * c1 + i] = XD[16384 + 128 + (Mp - 1 - (8 * c1 + i) - 8 * Jm0)];
      }
      __syncthreads();
    }
    bool synced2 = !(twist && side == 0);                          // side 0 owes the cluster one barrier (x_mid hand-over)
    if (!anybad) {
      // kBackStages - 1 stages in flight: one (possibly empty) cp.async group per tile row
      for (int k = 0; k < kBackStages - 1; ++k) {
        if (NTloc - 1 - k >= 0) stage_load(NTloc - 1 - k, (NTloc - 1 - k) % kBackStages);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      for (int J = NTloc - 1; J >= 0; --J) {
        if (!synced2 && J == c1 - 1) {                             // middle solved: publish it, then carry on downwards
          __syncthreads();
          for (int i = tau; i < 128; i += kMmaThreads) XD[16384 + 128 + i] = xsol[8 * c1 + i];
          if (tau == 0) gf[0] = 0;
          cluster_sync();
          synced2 = true;
          if (gf[1] != 0) { anybad = true; break; }
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(kBackStages - 2) : "memory");
        __syncthreads();                                           // stage J landed; iteration J+1 fully retired
        if (J - (kBackStages - 1) >= 0) stage_load(J - (kBackStages - 1), (J - (kBackStages - 1)) % kBackStages);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const double *st = Lst + (size_t)(J % kBackStages) * (8 * 128 + 64);
        const double *Wj = st + 8 * 128;                           // W_J row-major 8x8
        // x_J = W_J^T z_J, computed redundantly by every thread that uses it (broadcast loads, no exchange, one
        // barrier per tile row); the solution goes to xsol so that z_J stays readable during the iteration.
        if (tau < 32 + 120) {
          double xJ[8];
          if (J < ncols) {
            double zz[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) zz[k] = z[8 * J + k];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              double s0 = 0.0, s1 = 0.0;
#pragma unroll
              for (int k = i; k < 8; k += 2) { s0 += Wj[k * 8 + i] * zz[k]; if (k + 1 < 8) s1 += Wj[(k + 1) * 8 + i] * zz[k + 1]; }
              xJ[i] = s0 + s1;
            }
            if (tau == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) xsol[8 * J + i] = xJ[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) xJ[i] = xsol[8 * J + i];   // given (side 1, middle rows)
          }
          if (tau >= 32) {
            const int xcol = tau - 32, c = 8 * (J - 15) + xcol;
            if (c >= 0) {
              double s0 = 0.0, s1 = 0.0;
#pragma unroll
              for (int gg = 0; gg < 8; gg += 2) { s0 += st[gg * 128 + xcol] * xJ[gg]; s1 += st[(gg + 1) * 128 + xcol] * xJ[gg + 1]; }
              z[c] -= s0 + s1;
            }
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (!synced2) {                                                // failed before the hand-over: still meet the barrier
      if (tau == 0) gf[0] = 1;
      cluster_sync();
      synced2 = true;
      anybad = true;
    }
    __syncthreads();
    if (phase && tau == 0) phase[2] = clock64();
    // ---- write this side's rows of dX (zeros if any factorisation failed, ba.py:12-13), NaN check ----
    const int nrows = 8 * ((twist && side == 1) ? c1 : NTloc);
    int nan_local = 0;
    for (int rl = tau; rl < nrows; rl += kMmaThreads) {
      const int r = side ? Mp - 1 - rl : rl;
      if (r < M) { const double v = anybad ? 0.0 : xsol[rl]; cv.dX[r] = v; nan_local |= (v != v); }
    }
    if (nan_local) s_nan = 1;
    __syncthreads();
    int any_nan = s_nan;
    if (twist) {
      if (tau == 0 && s_nan) atomicOr(&gf[2], 1);
      cluster_sync();
      any_nan = gf[2];
      anybad = anybad || gf[0] != 0 || gf[1] != 0;
    }
    if (anybad) { status |= (attempt == 0) ? 1 : 4; break; }         // no NaN possible -> no retry
    if (any_nan && allow_retry && attempt == 0) { status |= 2; __syncthreads(); continue; }   // ba.py:324-325
    break;
  }
  if (tau == 0 && side == 0) cv.status[0] = status;
  if (phase && tau == 0) phase[3] = clock64();
}

size_t solve_mma_smem_bytes(int M) {
  const int Mp = ((M + 7) / 8) * 8;
  return ((size_t)2 * Mp + 8 * 2 * 64 + 2 * 17 * kTs + 2 * kTs + 8 + kBackStages * (8 * 128 + 64)) * sizeof(double) +
         (16 * 136 + 2 * 16 * 8) * sizeof(unsigned);
}

// doubles of scratch the solver needs besides [S | y]: both sides' L (local band storage) and W tiles, the
// hand-over block, and a few ints of flags
size_t solve_mma_scratch_doubles(int M, int bw) {
  const size_t Mp = ((size_t)(M + 7) / 8) * 8;
  return 2 * Mp * (bw + 1) + 2 * (Mp / 8) * 64 + (128 * 128 + 256) + 8;
}

int launch_solve_band_mma(const CallView &cv, int allow_retry, double *scratch, cudaStream_t s) {
  static bool attr_set = false;
  static long long *trace = nullptr;
  static int trace_left = 0;
  static int twist_min = 64;
  static int iso = 0;        // measured on B200 (cfg3): [A] 3322 -> 1870 cycles but [U] 2464 -> 3121 with 3 tile warps per scheduler: slower overall
  static unsigned *gtab = nullptr;
  if (!attr_set) {
    BA_CUDA(cudaMalloc(&gtab, (16 * 136 + 2 * 16 * 8) * sizeof(unsigned)));
    k_build_step_tables<<<1, 256, 0, s>>>(gtab);
    BA_LAUNCH_CHECK();
    BA_CUDA(cudaFuncSetAttribute(k_solve_band_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
    if (const char *e = getenv("BA_TRACE")) { trace_left = atoi(e); BA_CUDA(cudaMalloc(&trace, 16 * 8 * 4096 + 256)); }
    if (const char *e = getenv("BA_TWIST_MIN")) twist_min = atoi(e);        // tile columns from which two CTAs are used
    if (const char *e = getenv("BA_SOLVE_ISO")) iso = atoi(e);              // 1: 12 hardware warps, factor warp alone on scheduler 0
    attr_set = true;
  }
  const size_t Mp = ((size_t)(cv.M + 7) / 8) * 8;
  const int nt = (int)(Mp / 8);
  double *L_all = scratch, *W_all = L_all + 2 * Mp * (cv.bw + 1), *XD = W_all + 2 * (Mp / 8) * 64;
  int *gfl = reinterpret_cast<int *>(XD + 128 * 128 + 256);
  const int twist = nt >= twist_min ? 1 : 0;
  const bool tr = trace && trace_left > 0 && nt <= 4096;
  BA_CUDA(cudaMemsetAsync(gfl, 0, 8 * sizeof(int), s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(twist ? 2 : 1);
  cfg.blockDim = dim3(iso ? kMmaHwThreads : kMmaThreads);
  cfg.dynamicSmemBytes = solve_mma_smem_bytes(cv.M);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = twist ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BA_CUDA(cudaLaunchKernelEx(&cfg, k_solve_band_mma, cv, allow_retry, W_all, L_all, XD, gfl, twist, (const unsigned *)gtab, tr ? trace : (long long *)nullptr));
  BA_LAUNCH_CHECK();
  if (tr && --trace_left == 0) {       // debug only: synchronises and prints mean phase lengths in SM cycles
    std::vector<long long> h((size_t)16 * 4096 + 32);
    BA_CUDA(cudaStreamSynchronize(s));
    BA_CUDA(cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost));
    double d[8] = {0};
    const int ncol = twist ? (nt - 16) / 2 : nt;                       // side 0, first segment
    for (int J = 1; J + 1 < ncol; ++J) {
      const long long *a = &h[(size_t)J * 16], *n = &h[(size_t)(J + 1) * 16];
      d[0] += a[1] - a[0]; d[1] += a[2] - a[1]; d[2] += a[3] - a[2]; d[3] += a[4] - a[3]; d[4] += a[5] - a[4];
      d[5] += n[0] - a[5]; d[6] += a[9] - a[8]; d[7] += a[10] - a[9];
    }
    const double k = ncol - 2;
    for (int sd = 0; sd < (twist ? 2 : 1); ++sd) {
      const long long *ph = &h[(size_t)16 * 4096 + sd * 8];
      fprintf(stderr, "[BA_TRACE] side %d: factor+forward %lld  back-substitution %lld  tail %lld cycles\n", sd,
              ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2]);
    }
    fprintf(stderr, "[BA_TRACE] tiles %d | tile warp 0: extract %.0f waitW %.0f P %.0f waitP %.0f U %.0f refill %.0f | "
            "factor warp: waitD %.0f A %.0f | step %.0f cycles\n", nt, d[0] / k, d[1] / k, d[2] / k, d[3] / k, d[4] / k,
            d[5] / k, d[6] / k, d[7] / k, (double)(h[(size_t)(ncol - 1) * 16] - h[16]) / k);
  }
  return BA_OK;
}

}  // namespace ba
