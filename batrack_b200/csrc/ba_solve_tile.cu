// ba_solve_tile.cu — reduced camera solve for SMALL and MEDIUM systems: one CTA, the band window as 8x8 tiles in
// shared memory, FP64 tensor-core (DMMA) updates  (ba.py:60-70 block_solve, :5-19 CholeskySolver, :323-325 NaN retry).
//
//   A = S + (ep + lm diag S) I ;  A = L L^T ;  dX = A^-1 y
//
// The systems BA-Track really builds are small: 15 free poses in the DAVIS window (90 unknowns, dense), up to 48 in a
// full-sequence Sintel window (288 unknowns, half bandwidth 131). The register-window solver (ba_solve_diag.cu) is built
// for long bands <= 120 wide and pays ~3.4 k cycles per tile column whatever the size; the scalar window solver took
// 640 us on the Sintel window. This kernel covers every system whose band fits (bw/8 + 2)^2 tiles of shared memory
// (half bandwidth <= 145: dense systems up to 152 unknowns).
//
// Window: tile rows J .. J + nbt, tile (a, a - d) at win[(a % R) * R + d], R = nbt + 1. Per tile column J, two barriers:
//   [P] every warp turns its panel tiles into L_aJ = A_aJ W_J^T (2 DMMAs), in place in OPERAND layout (one 16-byte load
//       per lane feeds a DMMA later), to the right-hand side (z_a -= L_aJ zJ) and to the row-major factor in global
//       memory (the back substitution streams it back);
//   [U] trailing update C_ab -= L_aJ L_bJ^T of the window, tile pairs spread over the warps — while warp 0 updates tile
//       (J+1, J+1) first and factors it (8x8 Cholesky + W = L^-1 + forward substitution in registers, every lane
//       redundantly, as in ba_solve_diag.cu): the factorisation of the next diagonal tile hides behind the update. The
//       warps that share warp 0's scheduler take no DMMA work (an FP64-busy neighbour slows the pivot chain 2-3x,
//       tools/microbench_factor.cu); they load the tile row that enters the window.
// Back substitution by tile rows, descending; the rows of L come back through a cp.async double buffer.
#include <algorithm>

#include "ba_internal.h"

namespace ba {

namespace {

constexpr int kTsWarps = 16;
constexpr int kTsThreads = 32 * kTsWarps;
constexpr int kTsMaxR = 20;                   // window tile rows (nbt + 1): 20^2 tiles = 200 KB

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
      : "=d"(d0), "=d"(d1)
      : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
constexpr int tri8(int a, int b) { return a * (a + 1) / 2 + b; }
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kTsThreads, 1) k_solve_tiles(CallView cv, int allow_retry, double *__restrict__ Lg, int nbt, long long *__restrict__ trace) {
  extern __shared__ __align__(16) double dsm[];
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5, g = lane >> 2, q = lane & 3;
  const int M = cv.M, bw = cv.bw, ld = cv.ld, off = cv.off;
  const int NT = (M + 7) >> 3, Mp = NT * 8, R = nbt + 1;
  double *win = dsm;                          // [R][R][64]
  double *z = win + (size_t)max(R * R, 2 * (R + 1)) * 64;   // [Mp] right-hand side -> forward solution -> solution
  double *Wsm = z + Mp;                       // [2][64] W = L_JJ^-1, operand layout, by column parity
  double *zJ = Wsm + 128;                     // [2][8]
  double *Dt = zJ + 16;                       // [64]   diagonal tile on its way into warp 0's registers
  __shared__ int s_colfail[2], s_nan;
  double *Lt = Lg;                            // [NT][R][64] row-major tiles of L: tile (a, a - d) at (a * R + d) * 64
  double *Wg = Lg + (size_t)NT * R * 64;      // [NT][64]    W_J, row-major
  const double *__restrict__ S = cv.S;
  const double ep = (double)cv.ep;
  // operand layout of an 8x8 tile: element (r, c) at r*8 + (c&3)*2 + (c>>2): lane (g, q) reads (g, q) and (g, 4+q) with
  // one 16-byte load at 2*lane; oc0 / oc1: where the lane's C-fragment elements (g, 2q), (g, 2q+1) go
  const int oc0 = ((2 * q) & 3) * 2 + ((2 * q) >> 2), oc1 = ((2 * q + 1) & 3) * 2 + ((2 * q + 1) >> 2);
  int status = 0;
  // optional phase stamps (BA_OPT_SOLVER_TRACE bit 2, tools/solver_ab.py): 16 slots per tile column
#define TS_TR(w, slot) do { if (trace && warp == (w) && lane == 0) trace[(size_t)J * 16 + (slot)] = clock64(); } while (0)

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    // tile row a of A = S + (ep + lm diag S) I  -> window slot a % R, row-major tiles; rows >= M are identity padding
    auto load_row = [&](int a, int t, int nt) {
      double *dst = win + (size_t)(a % R) * R * 64;
      constexpr int kB = 14;                                         // loads in flight per thread (a store right behind its
      for (int e0 = t; e0 < R * 64; e0 += kB * nt) {                 // load would serialise the L2 round trips)
        double v[kB];
#pragma unroll
        for (int u = 0; u < kB; ++u) {
          const int e = e0 + u * nt;
          const int d = e >> 6, i = (e >> 3) & 7, j = e & 7;
          const int r = 8 * a + i, c = 8 * (a - d) + j;
          v[u] = 0.0;
          if (e < R * 64 && c >= 0 && c <= r) {
            if (r >= M) v[u] = r == c ? 1.0 : 0.0;
            else if (r - c <= bw) v[u] = S[(size_t)r * ld + c + off];
          }
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
          const int e = e0 + u * nt;
          if (e < R * 64) {
            const int d = e >> 6, i = (e >> 3) & 7, j = e & 7;
            const bool dg = d == 0 && i == j && 8 * a + i < M;
            dst[e] = dg ? v[u] + (ep + lm * v[u]) : v[u];            // ba.py:67
          }
        }
      }
    };
    // [A] warp 0: factor the diagonal tile of column Jc (row-major in Dt), W and the forward substitution of the block
    auto factor = [&](int Jc) {
      const int pc = Jc & 1;
      double a[36];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; j += 2) {
          const double2 v = *reinterpret_cast<const double2 *>(Dt + i * 8 + j);
          a[tri8(i, j)] = v.x;
          if (j + 1 <= i) a[tri8(i, j + 1)] = v.y;
        }
      double zr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) zr[k] = lane == 8 ? z[8 * Jc + k] : 0.0;     // (lane 8 alone reads and later rewrites the block of z)
      bool ok = true;
      double wv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double piv = a[tri8(k, k)];
        ok = ok && ((float)piv > 0.0f);                             // potrf info != 0 (incl. NaN), ba.py:11
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
        const double inv = fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
        double sv = lane == 8 ? zr[k] : (lane == k ? 1.0 : 0.0);    // lanes 0-7: column `lane` of W; lane 8: zJ
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
      if (lane <= 8) {
        double *dst = lane < 8 ? Wsm + 64 * pc + (lane & 3) * 2 + (lane >> 2) : zJ + 8 * pc;
        const int st = lane < 8 ? 8 : 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i * st] = wv[i];
        if (lane == 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) z[8 * Jc + i] = wv[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) Wg[(size_t)Jc * 64 + i * 8 + lane] = wv[i];   // row-major W for the back substitution
        }
      }
      if (lane == 0) s_colfail[pc] = ok ? 0 : 1;
    };

    for (int a = warp; a < min(R, NT); a += kTsWarps) load_row(a, lane, 32);
    for (int r = tau; r < Mp; r += kTsThreads) z[r] = r < M ? cv.y[r] : 0.0;
    if (tau == 0) { s_colfail[0] = s_colfail[1] = 0; s_nan = 0; }
    __syncthreads();
    if (warp == 0) {
      Dt[lane * 2] = win[lane * 2]; Dt[lane * 2 + 1] = win[lane * 2 + 1];     // tile (0, 0): slot 0, d = 0
      __syncwarp();
      factor(0);
    }
    __syncthreads();
    bool failed = false;
    for (int J = 0; J < NT; ++J) {
      const int p = J & 1;
      if (s_colfail[p]) { failed = true; break; }
      const int na = min(nbt, NT - 1 - J);                          // tile rows below J
      const int jm = J % R;                                         // window slot of tile row J; row J + 1 + x sits in slot_of(x)
      auto slot_of = [&](int x) { const int s2 = jm + 1 + x; return s2 >= R ? s2 - R : s2; };
      TS_TR(0, 0);
      // ---- [P] panel tiles ----
      {
        const double2 wb = *reinterpret_cast<const double2 *>(Wsm + 64 * p + 2 * lane);      // B[k][n] = W[n][k]
        const double zq0 = zJ[8 * p + 2 * q], zq1 = zJ[8 * p + 2 * q + 1];
        for (int k = warp; k < na; k += kTsWarps) {
          const int a = J + 1 + k, d = k + 1;
          double *t = win + (slot_of(k) * R + d) * 64;
          const double ax = t[g * 8 + q], ay = t[g * 8 + 4 + q];    // row-major A tile -> A fragments
          double l0, l1;
          dmma884(l0, l1, ax, wb.x, 0.0, 0.0);
          dmma884(l0, l1, ay, wb.y, l0, l1);                        // L_aJ = A_aJ W^T
          __syncwarp();                                             // everybody has read the tile
          t[g * 8 + oc0] = l0; t[g * 8 + oc1] = l1;                 // operand layout, in place
          double part = l0 * zq0 + l1 * zq1;                        // z_a -= L_aJ zJ
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          if (q == 0) z[8 * a + g] -= part;
          *reinterpret_cast<double2 *>(Lt + ((size_t)a * R + d) * 64 + g * 8 + 2 * q) = make_double2(l0, l1);
        }
      }
      TS_TR(0, 1);
      __syncthreads();
      TS_TR(0, 2); TS_TR(5, 8); TS_TR(4, 12);
      // ---- [U] trailing update + look-ahead factorisation + the entering tile row ----
      if (warp == 0) {
        if (na >= 1) {
          const int a = J + 1;
          const double2 la = *reinterpret_cast<const double2 *>(win + (slot_of(0) * R + 1) * 64 + 2 * lane);
          double *c = win + (slot_of(0) * R) * 64;
          double2 cc = *reinterpret_cast<const double2 *>(c + g * 8 + 2 * q);
          dmma884(cc.x, cc.y, -la.x, la.x, cc.x, cc.y);
          dmma884(cc.x, cc.y, -la.y, la.y, cc.x, cc.y);
          *reinterpret_cast<double2 *>(Dt + g * 8 + 2 * q) = cc;
          __syncwarp();
          TS_TR(0, 3);
          factor(a);
          TS_TR(0, 4);
        }
      } else if ((warp & 3) == 0) {
        // warps 4, 8, 12 (warp 0's scheduler): no FP64 work — the tile row that enters the window (slot of row J)
        const int a = J + R;
        if (a < NT) load_row(a, (warp / 4 - 1) * 32 + lane, 96);
        TS_TR(4, 13);
      } else {
        // pairs (a, b), J + 1 <= b <= a <= J + na, without (J+1, J+1): 12 warps, two pairs in flight per warp.
        // pair index pi = ra (ra + 1) / 2 + rb, 0 <= rb <= ra < na, walked incrementally
        const int w12 = warp - 1 - (warp >> 2);                      // 0 .. 11
        const int npair = na * (na + 1) / 2;
        int ra = 0, rb = 1 + w12;
        auto norm = [&]() { while (rb > ra) { rb -= ra + 1; ++ra; } };
        norm();
        auto tile_ptrs = [&](int xa, int xb, const double *&pa, const double *&pb, double *&pc) {
          const int sa = slot_of(xa) * R, sb = slot_of(xb) * R;
          pa = win + (sa + (xa + 1)) * 64 + 2 * lane;
          pb = win + (sb + (xb + 1)) * 64 + 2 * lane;
          pc = win + (sa + (xa - xb)) * 64 + g * 8 + 2 * q;
        };
        for (int pi = 1 + w12; pi < npair; pi += 24) {
          const double *pa0, *pb0, *pa1, *pb1;
          double *pc0, *pc1;
          tile_ptrs(ra, rb, pa0, pb0, pc0);
          rb += 12; norm();
          const bool two = pi + 12 < npair;
          tile_ptrs(two ? ra : 0, two ? rb : 0, pa1, pb1, pc1);
          rb += 12; norm();
          const double2 la0 = *reinterpret_cast<const double2 *>(pa0), lb0 = *reinterpret_cast<const double2 *>(pb0);
          double2 c0 = *reinterpret_cast<const double2 *>(pc0);
          double2 la1 = la0, lb1 = lb0, c1 = c0;
          if (two) { la1 = *reinterpret_cast<const double2 *>(pa1); lb1 = *reinterpret_cast<const double2 *>(pb1); c1 = *reinterpret_cast<const double2 *>(pc1); }
          dmma884(c0.x, c0.y, -la0.x, lb0.x, c0.x, c0.y);
          dmma884(c1.x, c1.y, -la1.x, lb1.x, c1.x, c1.y);
          dmma884(c0.x, c0.y, -la0.y, lb0.y, c0.x, c0.y);
          dmma884(c1.x, c1.y, -la1.y, lb1.y, c1.x, c1.y);
          *reinterpret_cast<double2 *>(pc0) = c0;
          if (two) *reinterpret_cast<double2 *>(pc1) = c1;
        }
        TS_TR(5, 9);
      }
      __syncthreads();
      TS_TR(0, 5);
    }
    if (failed) {                                                   // dX = 0 (ba.py:12-13); no NaN -> no retry
      for (int r = tau; r < M; r += kTsThreads) cv.dX[r] = 0.0;
      status |= (attempt == 0) ? 1 : 4;
      break;
    }
    // ---- backward substitution L^T x = z by tile rows, descending. Row J of L (tiles (J, J - d), row-major) and W_J
    //      come back from global memory through a double buffer in the (now dead) window ----
    __threadfence_block();
    __syncthreads();
    double *stage = win;                                            // [2][(R + 1) * 64]
    const int srow = (R + 1) * 64;
    auto prefetch = [&](int J) {
      if (J >= 0) {
        double *dst = stage + (size_t)(J & 1) * srow;
        const double *src = Lt + (size_t)J * R * 64;
        for (int e = tau * 2; e < R * 64; e += kTsThreads * 2) cp_async16(dst + e, src + e);
        if (tau < 32) cp_async16(dst + R * 64 + 2 * tau, Wg + (size_t)J * 64 + 2 * tau);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(NT - 1);
    for (int J = NT - 1; J >= 0; --J) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                              // row J staged; z_J final (updates of row J + 1 done)
      prefetch(J - 1);                                              // into the stage row J + 1 was read from: everybody is past it
      const double *st = stage + (size_t)(J & 1) * srow;
      if (tau < 8) {                                                // x_J = W_J^T z_J
        const double *w = st + R * 64;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i += 2) { s0 = fma(w[i * 8 + tau], z[8 * J + i], s0); s1 = fma(w[(i + 1) * 8 + tau], z[8 * J + i + 1], s1); }
        Dt[tau] = s0 + s1;
      }
      __syncthreads();
      if (tau < 8) z[8 * J + tau] = Dt[tau];
      {                                                             // z_b -= L_{J,b}^T x_J, b = J - d
        const int d = 1 + (tau >> 3), j = tau & 7;
        if (d <= nbt && J - d >= 0) {
          const double *t = st + (size_t)d * 64;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int i = 0; i < 8; i += 2) { s0 = fma(t[i * 8 + j], Dt[i], s0); s1 = fma(t[(i + 1) * 8 + j], Dt[i + 1], s1); }
          z[8 * (J - d) + j] -= s0 + s1;
        }
      }
    }
    __syncthreads();
    int nan_local = 0;
    for (int r = tau; r < M; r += kTsThreads) { const double v = z[r]; cv.dX[r] = v; nan_local |= (v != v); }
    if (nan_local) s_nan = 1;
    __syncthreads();
    if (s_nan && allow_retry && attempt == 0) { status |= 2; __syncthreads(); continue; }   // ba.py:324-325
    break;
  }
  if (tau == 0) cv.status[0] = status;
}

static int win_tiles(int R) { return std::max(R * R, 2 * (R + 1)); }   // the back substitution's two stages reuse the window
static size_t solve_tiles_smem(int M, int nbt) {
  const int NT = (M + 7) / 8, R = nbt + 1;
  return ((size_t)win_tiles(R) * 64 + (size_t)NT * 8 + 128 + 16 + 64) * sizeof(double);
}

// Does the tile solver cover this system? nbt = tile sub-diagonals of the band.
bool solve_tiles_applies(int M, int bw, int *nbt_out) {
  if (M <= 0) return false;
  const int NT = (M + 7) / 8;
  const int nbt = std::min((bw + 7) >> 3, NT - 1);
  if (nbt_out) *nbt_out = nbt;
  if (nbt + 1 > kTsMaxR) return false;
  return solve_tiles_smem(M, nbt) <= 227 * 1024 - 64;
}

size_t solve_tiles_scratch_doubles(int M, int bw) {
  int nbt = 0;
  if (!solve_tiles_applies(M, bw, &nbt)) return 0;
  const size_t NT = (size_t)(M + 7) / 8;
  return NT * (nbt + 1) * 64 + NT * 64 + 8;
}

int launch_solve_tiles(const CallView &cv, int allow_retry, double *scratch, long long *trace, cudaStream_t s) {
  int nbt = 0;
  if (!solve_tiles_applies(cv.M, cv.bw, &nbt)) return BA_ERR_ARG;
  k_solve_tiles<<<1, kTsThreads, solve_tiles_smem(cv.M, nbt), s>>>(cv, allow_retry, scratch, nbt, trace);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

int solve_tiles_prepare_device() {
  return cudaFuncSetAttribute(k_solve_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64) == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}

}  // namespace ba
