// ba_solve_diag.cu — reduced camera solve on the FP64 tensor cores, diagonal tile ownership
// (ba.py:60-70 block_solve, :5-19 CholeskySolver, :323-325 NaN retry).
//
//   A = S + (ep + lm diag S) I ;  A = L L^T ;  dX = A^-1 y          band half-width bw <= 120
//
// Blocked right-looking band Cholesky with 8x8 tiles. The active window is the 16x16-tile square [J, J+15]^2;
// every lower tile of it lives in REGISTERS as the C fragment of an m8n8k4 DMMA. Ownership is by WINDOW-RELATIVE
// DIAGONAL: tile warp w holds the diagonals d = w and d = 15 - w (tiles (j + d, j), j = 0 .. 15 - d: 17 tiles per
// warp, 32 or 34 DMMAs per column). When the window slides by one tile column every tile moves one step along its
// own diagonal, i.e. stays in the same warp, and because a DMMA has separate C and D operands the move is free:
//       tile[j-1] <- L_{j+d} L_j^T + tile[j]        (tiles hold -A, so the update is an addition)
// All register indices are compile-time constants (one instantiation of the column body per warp): no step
// tables, no predicated merges, no per-tile shared slots. Per tile column J:
//   [A] the factor warp factors the 8x8 diagonal block in registers (every lane redundantly, branch-free; lanes
//       0-7 also solve for the columns of W = L_JJ^-1, lane 8 forward-substitutes the right-hand side);
//   [P] each warp turns the first tile of its diagonals into L_dJ = A_dJ W^T (2 DMMAs), stores it to shared memory
//       in operand layout (one 16-byte load per lane and tile later), to global L, and updates the right-hand side;
//   [U] each warp fetches the 15 panel tiles as DMMA operands (15 LDS.128) and updates its 15-16 tiles, two passes
//       of independent DMMAs; the entering tile row (loaded from global at the top of the step) fills the free end
//       of each diagonal.
// Look-ahead: the warp of diagonal 1 finishes L_{J+1,J} first and signals the warp of diagonal 0, which updates
// tile (J+1,J+1), adds the damping and hands it to the factor warp: [A] of column J+1 overlaps [P]/[U] of column J.
// Twist (two CTAs eliminating from both ends), streaming hand-over from the Schur kernel, stage-format factor and
// the bulk-copy back substitution are those of DESIGN.md §4 K3.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "ba_internal.h"

namespace ba {

namespace {

constexpr int kNW = 8;                        // tile warps
constexpr int kThreadsDg = 32 * (kNW + 1);    // + the factor warp
constexpr int kBackStages = 4;                // tile rows of L in flight during the back substitution
constexpr int kPs = 12;                       // row stride (doubles) of the shared diagonal tile

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  // not volatile: pure function of its operands, so the compiler may interleave the DMMAs of independent tiles
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
      : "=d"(d0), "=d"(d1)
      : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// named barriers (id 0 is __syncthreads): 1 = panel tiles complete (tile warps), 2 = diagonal tile published
// (warp of diagonal 0 -> factor warp), 3 = W_J / zJ ready (factor warp -> tile warps; also closes the previous [U]),
// 4 = panel tile L_{J+1,J} stored (warp of diagonal 1 -> warp of diagonal 0), 5 = x_J of the back substitution
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

constexpr int tri8(int a, int b) { return a * (a + 1) / 2 + b; }
// operand layout of an 8x8 tile in shared memory: element (g, c) at g*8 + (c&3)*2 + (c>>2), so that lane (g, q) of a
// DMMA reads its two fragment elements (g, q) and (g, 4+q) with ONE 16-byte load at 2*lane — a warp reads 512
// contiguous bytes
__device__ __forceinline__ int op_idx(int g, int c) { return g * 8 + (c & 3) * 2 + (c >> 2); }

#define BA_TR(slot) do { if (trace && lane == 0) trace[(size_t)J * 16 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ long long clk_after(double dep) { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep)); return t; }
__device__ __forceinline__ long long clk_after(int dep) { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep)); return t; }
#define BA_TRD(slot, dep) do { if (trace && lane == 0) trace[(size_t)J * 16 + (slot)] = clk_after(dep); } while (0)

__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

// twist != 0: launched as a cluster of two CTAs. CTA 0 eliminates tile columns [0, Jm0) of the matrix, CTA 1 the
// last Jm1 tile columns, working on the index-reversed matrix (same code, reversed coordinates); CTA 1 then hands
// CTA 0 what its eliminations contributed to the 16 middle tile columns, CTA 0 finishes the middle, solves it, and
// both back-substitute their side in parallel. Exchange through global scratch XD + cluster barriers.
__global__ void __launch_bounds__(kThreadsDg, 1) k_solve_band_diag(CallView cv, int allow_retry, double *__restrict__ L_all,
                                                                   double *__restrict__ XD, int *__restrict__ gfl, int twist,
                                                                   long long *__restrict__ trace, SolveFeed feed) {
  extern __shared__ __align__(16) double dsm[];
  // mode 2: stand-by launch behind a streaming one — runs only if that one gave up (its producer was not running
  // concurrently: kernels serialised by a profiler / sanitizer), as a plain solve of the by now complete system
  if (feed.mode == 2 && *reinterpret_cast<const volatile int *>(feed.redo) == 0) return;
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
  const bool is_factor = warp == kNW, is_tile = warp < kNW;
  const int g = lane >> 2, q = lane & 3;
  const int M = cv.M, bw = cv.bw, ld = cv.ld, off = cv.off;
  const int NT8 = (M + 7) >> 3, Mp = NT8 * 8;
  const int side = twist ? (int)blockIdx.x : 0;
  const int Jm0 = (NT8 - 16) / 2, Jm1 = NT8 - 16 - Jm0;
  const int c1 = twist ? (side ? Jm1 : Jm0) : NT8;                  // end of this side's first segment
  const int NTloc = twist ? c1 + 16 : NT8;                          // tiles this side ever sees (local coordinates)
  // This side's factor in "stage format": 128 doubles per row; row r of tile row T = r / 8 holds L(r, c) for the 120
  // columns c in [8 (T - 15), 8 T) at offset c - 8 T + 120, and W_T (the inverted diagonal tile, row r - 8 T) in the
  // last 8 slots. A tile row is one contiguous 8 KB block: the back substitution fetches it with ONE bulk copy.
  double *__restrict__ L = L_all + (size_t)side * Mp * 128;
  auto Lg = [&](int rl, int cl) { return (size_t)rl * 128 + (cl - 8 * (rl >> 3) + 120); };
  double *z = dsm;                         // [Mp]   right-hand side -> forward solution -> solution
  double *dd = z + Mp;                     // [Mp]   damping ep + lm * S_rr, added when a diagonal tile is factored
  double *Psm = dd + Mp;                   // [16][64] panel tiles L_{J+i,J}, i = 1..15, operand layout
  double *Dsm = Psm + 16 * 64;             // [8][kPs] diagonal tile handed to the factor warp
  double *Wsm = Dsm + 8 * kPs;             // [64]     -W = -L_JJ^-1, operand layout
  double *zJ = Wsm + 64;                   // [8]
  double *Lst = zJ + 8;                    // [kBackStages][8][128] back-substitution stages
  double *xsol = dd;                       // solution of the back substitution (dd is dead then)
  __shared__ int s_fail, s_nan, s_abort;
  __shared__ __align__(8) unsigned long long s_mbar[kBackStages];
  constexpr int kIssueThread = 160;        // lane 0 of tile warp 5: issues the stage copies
  int bs_it = 0;
  // ---- streaming mode (feed.flags != nullptr), see ba_internal.h SolveFeed ----
  __shared__ volatile int s_cursor, s_giveup;
  int wcur = 0, fprobe = 0;
  bool probe_on = false;
  auto need_of = [&](int R) -> int {
    if (!feed.flags) return 0;
    if (side == 0) return feed.top_need[min(8 * R + 7, M - 1) / 6 + feed.fixedp];
    const int rlo = max(Mp - 8 * R - 8, 0);
    return rlo >= M ? 0 : feed.bot_need[rlo / 6 + feed.fixedp];
  };
  auto poll = [&]() {                                               // factor warp only (all 32 lanes)
    if (wcur < feed.n_units) {
      const int idx = wcur + lane;
      int f = 0;
      if (idx < feed.n_units) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(feed.flags + idx) : "memory");
      const unsigned mk = __ballot_sync(0xffffffffu, f == feed.epoch);
      wcur += mk == 0xffffffffu ? 32 : __ffs(~mk) - 1;
      if (lane == 0) s_cursor = wcur;
    }
  };
  const int spin_cap = feed.spin_cap > 0 ? feed.spin_cap : (1 << 16);
  auto wait_cursor = [&](int need) {
    if (need > 0 && s_cursor < need) {
      int spins = 0;
      while (s_cursor < need && !s_giveup && ++spins < spin_cap) __nanosleep(40);
      if (s_cursor < need) s_giveup = 1;
      __threadfence();
    }
  };
  const double *S = cv.S;
  const double ep = (double)cv.ep;
  auto Sg = [&](int r, int c) { return (size_t)r * ld + c + off; };
  int status = 0;
  long long *phase = trace ? trace + 16 * 4096 + side * 8 : nullptr;
  if (side) trace = nullptr;
  if (phase && tau == 0) phase[0] = clock64();

  if (tau == 0) {
    for (int k = 0; k < kBackStages; ++k)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[k])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    auto Aval = [&](int rl, int cl) -> double {                  // local coordinates (reversed on side 1)
      if (cl > rl) return 0.0;
      const int r = side ? Mp - 1 - cl : rl, c = side ? Mp - 1 - rl : cl;   // global, r >= c
      if (r >= M) return r == c ? 1.0 : 0.0;
      if (r - c > bw) return 0.0;
      return __ldcg(S + Sg(r, c));
    };
    // tile (a, b), a >= b, C-fragment layout, branch-free (32-bit index arithmetic, predicated loads)
    auto load_frag = [&](int a, int b, double &c0, double &c1) {
      const int rl = 8 * a + g, cl = 8 * b + 2 * q;
      const int r0 = side ? Mp - 1 - cl : rl, c0g = side ? Mp - 1 - rl : cl;
      const int r1 = side ? r0 - 1 : r0, c1g = side ? c0g : c0g + 1;
      const int i0 = r0 * ld + c0g + off, i1 = side ? i0 - ld : i0 + 1;
      const bool in0 = cl <= rl && r0 < M && r0 - c0g <= bw, in1 = cl + 1 <= rl && r1 < M && r1 - c1g <= bw;
      const double p0 = (cl <= rl && r0 >= M && r0 == c0g) ? 1.0 : 0.0, p1 = (cl + 1 <= rl && r1 >= M && r1 == c1g) ? 1.0 : 0.0;
      c0 = in0 ? __ldcg(S + i0) : p0;
      c1 = in1 ? __ldcg(S + i1) : p1;
    };
    auto load_row = [&](int rl) {
      const int r = side ? Mp - 1 - rl : rl;
      z[rl] = r < M ? __ldcg(cv.y + r) : 0.0;
      dd[rl] = r < M ? ep + lm * __ldcg(S + Sg(r, r)) : 0.0;       // A = S + (ep + lm * S) .* I, ba.py:67
    };
    if (feed.flags) {                                              // the first window (16 tile rows) must be complete
      if (tau == 0) { s_cursor = 0; s_giveup = 0; }
      __syncthreads();
      const int need0 = need_of(min(15, NTloc - 1));
      if (is_factor) {
        wcur = 0;
        int spins = 0;
        while (wcur < need0 && ++spins < (spin_cap >> 4) + 8) { poll(); if (wcur < need0) __nanosleep(500); }
        if (wcur < need0 && lane == 0) s_giveup = 1;
      }
      else wait_cursor(need0);
      __syncthreads();
      __threadfence();
    }
    const int rows_now = feed.flags ? min(Mp, 128) : Mp;
    for (int rl = tau; rl < rows_now; rl += kThreadsDg) load_row(rl);
    if (tau == 0) { s_fail = (feed.flags && s_giveup) ? 1 : 0; s_nan = 0; s_abort = 0; }
    int *gf = gfl + 4 * attempt;                                   // [0] side 0 failed, [1] side 1 failed, [2] NaN
    __syncthreads();
    const int nsegs = twist ? 2 : 1;
    if (is_factor) {
      // =================== factor warp: [A] for column J while the tile warps still update column J-1 ==========
      for (int seg = 0; seg < nsegs; ++seg) {
        if (seg == 1) {                                            // twist hand-over (see the tile-warp branch)
          __syncthreads();
          cluster_sync();
          __syncthreads();
        }
        const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
        bool stop = false;
        for (int J = jb; J < je; ++J) {
          BA_TR(8);
          if (feed.flags && wcur < feed.n_units) {
            if (probe_on) {
              const unsigned mk = __ballot_sync(0xffffffffu, fprobe == feed.epoch);
              const int adv = mk == 0xffffffffu ? 32 : __ffs(~mk) - 1;
              if (adv) { wcur += adv; __threadfence(); if (lane == 0) s_cursor = wcur; }
            }
            const int nd = J + 16 < NTloc ? need_of(J + 16) : 0;
            int spins = 0;
            while (wcur < nd && !s_giveup && ++spins < (spin_cap >> 4) + 8) poll();
            if (wcur < nd && lane == 0) s_giveup = 1;
            fprobe = 0;
            if (wcur + lane < feed.n_units) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(fprobe) : "l"(feed.flags + wcur + lane) : "memory");
            probe_on = true;
          }
          if (feed.flags && (s_giveup || s_fail)) {                // the producer is not there: leave like a failed pivot
            __syncwarp();
            if (lane == 0) s_fail = 1;
            __syncwarp();
            bar_arrive(3, 32 * (kNW + 1));
            stop = true;
            break;
          }
          bar_sync(2, 64);                                         // tile (J,J) (+ damping) is in Dsm, z_J is final
          double a[36];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) a[tri8(i, j)] = Dsm[i * kPs + j];
          double zr[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) zr[k] = z[8 * J + k];
          BA_TRD(9, a[35]);
          // Right-looking 8x8 Cholesky fused with the forward substitutions: one right-hand side per lane, same
          // instruction stream, no branches — lanes 0..7 solve L_JJ w = e_lane (column `lane` of W = L_JJ^-1), lane 8
          // solves L_JJ zJ = z_J. A pivot that is not positive (or NaN) poisons the block with NaN / inf, which
          // nobody reads: `ok` turns into the failure flag (potrf info != 0, ba.py:11).
          bool ok = true;
          double wv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double piv = a[tri8(k, k)];
            ok = ok && ((float)piv > 0.0f);
            double y;                                              // seed from the high word (~2^-20) + one Newton step
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
            const double inv = fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
            double sv = lane == 8 ? zr[k] : (lane == k ? 1.0 : 0.0);
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
              for (int i = 0; i < 8; ++i) Wsm[op_idx(i, lane)] = -wv[i];
            } else if (lane == 8) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { z[8 * J + i] = wv[i]; zJ[i] = wv[i]; }
            }
          } else if (lane == 0) {
            s_fail = 1;
          }
          BA_TR(10);
          bar_arrive(3, 32 * (kNW + 1));                           // W_J, zJ (or the failure flag) published
          if (!ok) { stop = true; break; }
        }
        if (stop) break;
      }
    } else if (is_tile) {
      // =================== tile warps: one instantiation of the column loop per warp ===========================
      auto run = [&](auto d1c) {
        constexpr int D1 = decltype(d1c)::value, D2 = 15 - D1;
        constexpr int N1 = 16 - D1, NTL = 17;                       // tiles of diagonal D1; N1 + (16 - D2) = 17
        // slot t holds window-relative tile (TJ + TD, TJ)
#define TD(t) ((t) < N1 ? D1 : D2)
#define TJ(t) ((t) < N1 ? (t) : (t) - N1)
        double ct[NTL][2];                                          // -A of the tile, C-fragment layout
#pragma unroll
        for (int t = 0; t < NTL; ++t) {
          double v0 = 0.0, v1 = 0.0;
          if (TJ(t) + TD(t) < NTloc) load_frag(TJ(t) + TD(t), TJ(t), v0, v1);
          ct[t][0] = -v0; ct[t][1] = -v1;
        }
        auto publish_diag = [&](int Jn) {                           // D1 == 0: tile (Jn,Jn) = ct[0] -> factor warp
          const double dmp = dd[8 * Jn + g];
          Dsm[g * kPs + 2 * q] = (2 * q == g ? dmp : 0.0) - ct[0][0];
          Dsm[g * kPs + 2 * q + 1] = (2 * q + 1 == g ? dmp : 0.0) - ct[0][1];
          bar_arrive(2, 64);
        };
        double pend_z = 0.0, pend_d = 0.0;                          // streaming mode: y / damping of the rows fetched last column
        int pend_row = -1;
        if (D1 == 0) publish_diag(0);
        __syncwarp();
        const int oc0 = ((2 * q) & 3) * 2 + ((2 * q) >> 2), oc1 = ((2 * q + 1) & 3) * 2 + ((2 * q + 1) >> 2);
        const int src0 = (lane & ~3) | (q >> 1), src1 = src0 + 2;
        bool failed = false;
        for (int seg = 0; seg < nsegs && !failed; ++seg) {
          if (seg == 1) {
            // ---- twist hand-over. Side 1: what its eliminations did to the 16 middle tile columns (window minus
            //      the untouched matrix) and to the right-hand side goes to XD in its local coordinates. Side 0 adds
            //      it to its window and publishes the diagonal tile of column c1. ----
            __syncthreads();
            if (side == 1 && !s_fail) {
#pragma unroll
              for (int t = 0; t < NTL; ++t) {
                const int rl = 8 * (c1 + TJ(t) + TD(t)) + g, cl = 8 * (c1 + TJ(t)) + 2 * q;
                double *d = XD + (size_t)(rl - 8 * c1) * 128 + (cl - 8 * c1);
                d[0] = -ct[t][0] - Aval(rl, cl);
                d[1] = -ct[t][1] - Aval(rl, cl + 1);
              }
              const int i = warp * 32 + lane;
              if (i < 128) { const int r = Mp - 1 - (8 * c1 + i); XD[16384 + i] = z[8 * c1 + i] - (r < M ? __ldcg(cv.y + r) : 0.0); }
            }
            if (tau == 0) gf[side] = s_fail;
            cluster_sync();
            if (tau == 0) s_abort = side == 0 ? (gf[1] | s_fail) : s_fail;
            __syncthreads();
            if (side == 0 && !s_abort) {
#pragma unroll
              for (int t = 0; t < NTL; ++t) {
                const int r = 8 * (c1 + TJ(t) + TD(t)) + g, c = 8 * (c1 + TJ(t)) + 2 * q;
                const double *d = XD + (size_t)(Mp - 1 - c - 8 * Jm1) * 128 + (Mp - 1 - r - 8 * Jm1);
                ct[t][0] -= d[0];
                ct[t][1] -= d[-128];
              }
              const int i = warp * 32 + lane;
              if (i < 128) z[8 * c1 + i] += XD[16384 + 127 - i];
              bar_sync(1, 32 * kNW);                               // z of the middle complete before the factor warp reads it
              if (D1 == 0) publish_diag(c1);
              __syncwarp();
            }
          }
          const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
          for (int J = jb; J < je; ++J) {
            if (warp == 0) BA_TR(0);
            const int an = J + 16;                                  // tile row entering the window for column J + 1
            if (feed.flags) {
              // rows of the entering tile row: complete? Then fetch their y and diagonal — into registers now, into
              // z / dd one column later (first needed by that column's [P])
              if (pend_row >= 0) {
                if (warp == 0 && lane < 8) { z[pend_row + lane] = pend_z; dd[pend_row + lane] = pend_d; }
                pend_row = -1;
              }
              if (an < NTloc) {
                wait_cursor(need_of(an));
                if (8 * an >= 128) {
                  if (warp == 0 && lane < 8) {
                    const int rl = 8 * an + lane, r = side ? Mp - 1 - rl : rl;
                    pend_z = r < M ? __ldcg(cv.y + r) : 0.0;
                    pend_d = r < M ? ep + lm * __ldcg(S + Sg(r, r)) : 0.0;
                  }
                  pend_row = 8 * an;
                }
              }
            }
            double rf[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            if (an < NTloc) {                                       // tiles (an, an - d): the free ends of the two diagonals
              load_frag(an, an - D1, rf[0][0], rf[0][1]);
              load_frag(an, an - D2, rf[1][0], rf[1][1]);
            }
            if (warp == 0) BA_TR(1);
            bar_sync(3, 32 * (kNW + 1));                           // W_J, zJ ready; every tile warp is past [U](J-1)
            const int sf = s_fail;
            if (warp == 0) BA_TRD(2, sf);
            if (sf) { failed = true; break; }
            if (warp == kNW - 1) {                                  // W_J -> global for the back substitution
              L[(size_t)(8 * J + (lane >> 3)) * 128 + 120 + (lane & 7)] = -Wsm[op_idx(lane >> 3, lane & 7)];
              L[(size_t)(8 * J + 4 + (lane >> 3)) * 128 + 120 + (lane & 7)] = -Wsm[op_idx(4 + (lane >> 3), lane & 7)];
            }
            // ---- [P] panel tiles: L_dJ = A_dJ W^T = (-A_dJ) (-W)^T ----
            const double2 wb = *reinterpret_cast<const double2 *>(Wsm + 2 * lane);   // B[k][n] = -W[n][k], n = g, k = 4h + q
            const double zq0 = zJ[2 * q], zq1 = zJ[2 * q + 1];
            const int cJ = 8 * J + 2 * q;
            auto panel = [&](const double c0, const double c1v, const int d) {
              // C fragment (cols 2q, 2q+1 of row g) -> A fragments (col 4h + q of row g)
              const double v00 = __shfl_sync(0xffffffffu, c0, src0), v01 = __shfl_sync(0xffffffffu, c1v, src0);
              const double v10 = __shfl_sync(0xffffffffu, c0, src1), v11 = __shfl_sync(0xffffffffu, c1v, src1);
              const double a0 = (q & 1) ? v01 : v00, a1 = (q & 1) ? v11 : v10;
              double p0, p1;
              dmma884(p0, p1, a0, wb.x, 0.0, 0.0);
              dmma884(p0, p1, a1, wb.y, p0, p1);
              double part = p0 * zq0 + p1 * zq1;                   // right-hand side: z_a -= L_aJ zJ
              part += __shfl_xor_sync(0xffffffffu, part, 1);
              part += __shfl_xor_sync(0xffffffffu, part, 2);
              double *ps = Psm + d * 64 + g * 8;
              ps[oc0] = p0; ps[oc1] = p1;
              const int r = 8 * (J + d) + g;
              if (r < 8 * NTloc) {                                 // tiles below this side's matrix are all zero
                if (q == 0) z[r] -= part;
                double *lp = L + Lg(r, cJ);                          // even offset in a 1 KB-aligned row: one 16-byte store
                if (r - cJ <= bw) *reinterpret_cast<double2 *>(lp) = make_double2(p0, p1);
                else if (r - cJ - 1 <= bw) lp[1] = p1;
              }
            };
            const bool have_next = J + 1 < je;
            if (D1 >= 1) panel(ct[0][0], ct[0][1], D1);
            if (D1 == 1 && have_next) bar_arrive(4, 64);           // L_{J+1,J} stored: the look-ahead may start
            panel(ct[N1][0], ct[N1][1], D2);
            // ---- look-ahead: tile (J+1,J+1) only needs L_{J+1,J} ----
            if (D1 == 0 && have_next) {
              bar_sync(4, 64);
              const double2 f1 = *reinterpret_cast<const double2 *>(Psm + 64 + 2 * lane);
              double c0 = ct[1][0], c1v = ct[1][1];
              dmma884(c0, c1v, f1.x, f1.x, c0, c1v);
              dmma884(c0, c1v, f1.y, f1.y, c0, c1v);
              ct[0][0] = c0; ct[0][1] = c1v;
              publish_diag(J + 1);
            }
            if (warp == 0) BA_TR(3);
            bar_sync(1, 32 * kNW);                                 // all panel tiles (and z updates) of column J done
            if (warp == 0) BA_TRD(4, Psm[64]);
            // ---- [U] trailing update + slide: tile[j-1] <- L_{j+d} L_j^T + tile[j], two passes of independent DMMAs ----
            {
              double2 pf[16];
#pragma unroll
              for (int i = 1; i < 16; ++i) pf[i] = *reinterpret_cast<const double2 *>(Psm + i * 64 + 2 * lane);
#pragma unroll
              for (int t = 0; t < NTL; ++t) {
                if (TJ(t) >= 1) {
                  if (D1 == 0 && t == 1) { if (!have_next) dmma884(ct[t][0], ct[t][1], pf[TJ(t) + TD(t)].x, pf[TJ(t)].x, ct[t][0], ct[t][1]); }
                  else dmma884(ct[t][0], ct[t][1], pf[TJ(t) + TD(t)].x, pf[TJ(t)].x, ct[t][0], ct[t][1]);
                }
              }
#pragma unroll
              for (int t = 0; t < NTL; ++t) {
                if (TJ(t) >= 1) {
                  if (D1 == 0 && t == 1) { if (!have_next) dmma884(ct[t - 1][0], ct[t - 1][1], pf[TJ(t) + TD(t)].y, pf[TJ(t)].y, ct[t][0], ct[t][1]); }
                  else dmma884(ct[t - 1][0], ct[t - 1][1], pf[TJ(t) + TD(t)].y, pf[TJ(t)].y, ct[t][0], ct[t][1]);
                }
              }
              ct[N1 - 1][0] = -rf[0][0]; ct[N1 - 1][1] = -rf[0][1];
              ct[NTL - 1][0] = -rf[1][0]; ct[NTL - 1][1] = -rf[1][1];
            }
            if (warp == 0) BA_TR(5);
          }
          if (pend_row >= 0) {
            if (warp == 0 && lane < 8) { z[pend_row + lane] = pend_z; dd[pend_row + lane] = pend_d; }
            pend_row = -1;
          }
        }
#undef TD
#undef TJ
      };
      switch (warp) {
        case 0: run(std::integral_constant<int, 0>{}); break;
        case 1: run(std::integral_constant<int, 1>{}); break;
        case 2: run(std::integral_constant<int, 2>{}); break;
        case 3: run(std::integral_constant<int, 3>{}); break;
        case 4: run(std::integral_constant<int, 4>{}); break;
        case 5: run(std::integral_constant<int, 5>{}); break;
        case 6: run(std::integral_constant<int, 6>{}); break;
        default: run(std::integral_constant<int, 7>{}); break;
      }
    }
    __syncthreads();
    if (phase && tau == 0) phase[1] = clock64();
    const bool bad = s_fail || s_abort;
    const int ncols = (twist && side == 1) ? c1 : NTloc;            // tile columns of L (and W_J) this side owns

    // ---- backward substitution L^T x = z by tile rows, descending, in local coordinates ----
    auto stage_issue = [&](int J, int it) {
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&s_mbar[it % kBackStages]);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(Lst + (size_t)(it % kBackStages) * (8 * 128));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(8 * 128 * 8) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(L + (size_t)J * (8 * 128)), "r"(8 * 128 * 8), "r"(mb) : "memory");
    };
    auto stage_wait = [&](int it) {
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&s_mbar[it % kBackStages]);
      const unsigned par = (unsigned)(it / kBackStages) & 1u;
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(par) : "memory");
      }
    };
    asm volatile("fence.proxy.async;" ::: "memory");                // this CTA's stores to L -> visible to the bulk copies
    __threadfence_block();
    bool anybad = bad;
    if (twist && side == 1) {                                      // wait for x_mid
      cluster_sync();
      anybad = bad || gf[0] != 0 || gf[1] != 0;
      if (!anybad) {
        for (int i = tau; i < 128; i += kThreadsDg) xsol[8 * c1 + i] = XD[16384 + 128 + (Mp - 1 - (8 * c1 + i) - 8 * Jm0)];
      }
      __syncthreads();
    }
    bool synced2 = !(twist && side == 0);                          // side 0 owes the cluster one barrier (x_mid hand-over)
    if (!anybad) {
      __syncthreads();
      const int it0 = bs_it;
      int issued = 0, consumed = 0;
      if (tau == kIssueThread) {
        for (int k = 0; k < kBackStages - 1 && NTloc - 1 - k >= 0; ++k) stage_issue(NTloc - 1 - k, it0 + k);
      }
      issued = min(kBackStages - 1, NTloc);
      for (int J = NTloc - 1; J >= 0; --J) {
        if (!synced2 && J == c1 - 1) {                             // middle solved: publish it, then carry on downwards
          __syncthreads();
          for (int i = tau; i < 128; i += kThreadsDg) XD[16384 + 128 + i] = xsol[8 * c1 + i];
          if (tau == 0) gf[0] = 0;
          cluster_sync();
          synced2 = true;
          if (gf[1] != 0) { anybad = true; break; }
        }
        const int it = it0 + (NTloc - 1 - J);
        __syncthreads();
        if (J - (kBackStages - 1) >= 0) {
          if (tau == kIssueThread) stage_issue(J - (kBackStages - 1), it + kBackStages - 1);
          ++issued;
        }
        if (tau < 160) stage_wait(it);
        ++consumed;
        const double *st = Lst + (size_t)(it % kBackStages) * (8 * 128);
        if (tau < 8 && J < ncols) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            if (k >= tau) s0 = fma(st[k * 128 + 120 + tau], z[8 * J + k], s0);
            if (k + 1 >= tau) s1 = fma(st[(k + 1) * 128 + 120 + tau], z[8 * J + k + 1], s1);
          }
          xsol[8 * J + tau] = s0 + s1;
        }
        if (tau < 160) bar_sync(5, 160);
        if (tau >= 32 && tau < 32 + 120) {
          const int xcol = tau - 32, c = 8 * (J - 15) + xcol;
          if (c >= 0) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int gg = 0; gg < 8; gg += 2) { s0 += st[gg * 128 + xcol] * xsol[8 * J + gg]; s1 += st[(gg + 1) * 128 + xcol] * xsol[8 * J + gg + 1]; }
            z[c] -= s0 + s1;
          }
        }
      }
      if (tau < 160) for (; consumed < issued; ++consumed) stage_wait(it0 + consumed);
      bs_it = it0 + issued;
    }
    if (!synced2) {
      if (tau == 0) gf[0] = 1;
      cluster_sync();
      synced2 = true;
      anybad = true;
    }
    __syncthreads();
    if (phase && tau == 0) phase[2] = clock64();
    const int nrows = 8 * ((twist && side == 1) ? c1 : NTloc);
    int nan_local = 0;
    for (int rl = tau; rl < nrows; rl += kThreadsDg) {
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
    if (anybad) { status |= (attempt == 0) ? 1 : 4; break; }
    if (any_nan && allow_retry && attempt == 0) { status |= 2; __syncthreads(); continue; }   // ba.py:324-325
    break;
  }
  if (feed.mode == 1 && tau == 0 && s_giveup) atomicOr(feed.redo, 1);
  if (tau == 0 && side == 0) cv.status[0] = status | ((feed.mode == 1 && s_giveup) ? 8 : 0);
  if (phase && tau == 0) phase[3] = clock64();
}

size_t solve_diag_smem_bytes(int M) {
  const int Mp = ((M + 7) / 8) * 8;
  return ((size_t)2 * Mp + 16 * 64 + 8 * kPs + 64 + 8 + kBackStages * (8 * 128)) * sizeof(double);
}

int launch_solve_band_diag(const CallView &cv, int allow_retry, double *scratch, const SolveFeed &feed, cudaStream_t s) {
  const size_t Mp = ((size_t)(cv.M + 7) / 8) * 8;
  const int nt = (int)(Mp / 8);
  double *L_all = scratch, *XD = L_all + 2 * Mp * 128;
  {                                     // never-written slots of L must read as zero: clear when the shape changes
    const long long key = ((long long)cv.M << 20) | cv.bw;
    if (feed.shape_key && *feed.shape_key != key) {
      BA_CUDA(cudaMemsetAsync(L_all, 0, 2 * Mp * 128 * sizeof(double), s));
      *feed.shape_key = key;
    }
  }
  int *gfl = reinterpret_cast<int *>(XD + 128 * 128 + 256);
  const int twist = nt >= (feed.twist_min > 0 ? feed.twist_min : 64) ? 1 : 0;
  BA_CUDA(cudaMemsetAsync(gfl, 0, 8 * sizeof(int), s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(twist ? 2 : 1);
  cfg.blockDim = dim3(kThreadsDg);
  cfg.dynamicSmemBytes = solve_diag_smem_bytes(cv.M);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = twist ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BA_CUDA(cudaLaunchKernelEx(&cfg, k_solve_band_diag, cv, allow_retry, L_all, XD, gfl, twist, feed.trace, feed));
  BA_LAUNCH_CHECK();
  return BA_OK;
}

int solve_diag_prepare_device() {
  return cudaFuncSetAttribute(k_solve_band_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64) == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}

}  // namespace ba
