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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "ba_internal.h"

namespace ba {

namespace {

constexpr int kNW = 9;                        // tile warps
constexpr int kNP = 2;                        // panel warps. Twelve warps in all (hardware warp -> role: see the kernel): a
                                              // thirteenth would cap the kernel at 128 registers per thread (warps are
                                              // allocated in fours), and the tile warps need ~170
constexpr int kPanelTiles = 8;                // panel tiles per panel warp (8 + 7)
constexpr int kThreadsDg = 32 * (kNW + kNP + 1);   // all 12 warps; hardware warp 0 is the factor warp
constexpr int kHwThreadsDg = kThreadsDg;
// named barriers (id 0 is __syncthreads); [p] = parity of the tile column
constexpr int kBarP = 1;                      // 1, 2: panel tiles of column J stored      (panel warps arrive, tile warps wait)
constexpr int kBarA = 3;                      // 3, 4: next column's A tiles shipped        (tile warps arrive, panel + factor warps wait)
constexpr int kBarW = 5;                      // 5, 6: W_J / zJ (or the failure flag) ready (factor warp arrives, panel warps wait)
constexpr int kBarD = 7;                      // diagonal tile published                    (warp of diagonal 0 -> factor warp)
constexpr int kBarX = 8;                      // x_J of the back substitution
constexpr int kBarZ = 9;                      // twist hand-over: z of the middle complete (tile warps)
constexpr int kBarN = 10;                     // 10, 11: tile (J+1,J+1) without column J's update shipped (warp of diagonal 0 -> factor warp)
constexpr int kCntP = 32 * (kNW + kNP), kCntA = 32 * (kNW + kNP + 1), kCntW = 32 * (kNP + 1);
constexpr int kTwistShift = 60;               // side 0 eliminates 1/2 + 1/60 of the tile columns outside the middle
constexpr int kLs = 132;                       // doubles per row of the stage-format factor in global memory (128 + 4 of padding)
constexpr int kStRow = kLs * 8;                   // bytes between the rows of a staged tile row: rows 0,2,4,6 (1,3,5,7) of a tile
                                              // start 64 bytes apart modulo 128: a B-fragment load takes the minimum two wavefronts
constexpr int kStSlot = 8 * kStRow;           // one staged tile row
// BA_VERIFY_SYNC (tools/racecheck_verify.sh): every thread that reads or writes behind one of the back substitution's
// mbarriers arrives on it itself (the product build lets one lane arrive after a __syncwarp), and the helper warps wait
// for the row's `full` barrier themselves — the form in which compute-sanitizer's racecheck can follow the protocol
#ifdef BA_VERIFY_SYNC
constexpr int kXrdyCount = 32, kFdoneCount = 96;
#else
constexpr int kXrdyCount = 1, kFdoneCount = 3;
#endif
constexpr int kNear = 3;                      // tile distances the chain warp of the back substitution keeps to itself
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

// ---- shared-memory / mbarrier accessors on 32-bit shared addresses with immediate offsets: the back substitution's
//      chain warp is bound by its instruction count (a lone warp issues every ~6 cycles), so no address is recomputed
template <int OFF> __device__ __forceinline__ double lds64o(unsigned a) {
  double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ double2 lds128o(unsigned a) {
  double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ void sts128o(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(x), "d"(y) : "memory");
}
template <int OFF> __device__ __forceinline__ unsigned mbar_test_o(unsigned a, unsigned par) {
  unsigned done;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1+%2], %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
               : "=r"(done) : "r"(a), "n"(OFF), "r"(par) : "memory");
  return done;
}
template <int OFF> __device__ __forceinline__ void mbar_wait_o(unsigned a, unsigned par) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1+%2], %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done) : "r"(a), "n"(OFF), "r"(par) : "memory");
  }
}
template <int OFF> __device__ __forceinline__ void mbar_arrive_o(unsigned a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1];" ::"r"(a), "n"(OFF) : "memory");
}

}  // namespace

// twist != 0: launched as a cluster of two CTAs. CTA 0 eliminates tile columns [0, Jm0) of the matrix, CTA 1 the
// last Jm1 tile columns, working on the index-reversed matrix (same code, reversed coordinates); CTA 1 then hands
// CTA 0 what its eliminations contributed to the 16 middle tile columns, CTA 0 finishes the middle, solves it, and
// both back-substitute their side in parallel. Exchange through global scratch XD + cluster barriers.
__global__ void __launch_bounds__(kHwThreadsDg, 1) k_solve_band_diag(CallView cv, int allow_retry, double *__restrict__ L_all,
                                                                   double *__restrict__ XD, int *__restrict__ gfl, int twist,
                                                                   long long *__restrict__ trace, SolveFeed feed, int rlog) {
  extern __shared__ __align__(16) double dsm[];
  // mode 2: stand-by launch behind a streaming one — runs only if that one gave up (its producer was not running
  // concurrently: kernels serialised by a profiler / sanitizer), as a plain solve of the by now complete system
  if (feed.mode == 2 && *reinterpret_cast<const volatile int *>(feed.redo) == 0) return;
  // The 8x8 factorisation is a dependent FP64 chain: behind two DMMA warps on its scheduler it runs 3.3x slower than
  // alone (tools/microbench_factor.cu: 674 -> 2216 cycles). Scheduler 0 therefore holds the factor warp and the two
  // panel warps, which only work while the factor warp waits for the next diagonal tile; the nine tile warps (the
  // DMMA-bound trailing update) sit three per scheduler on schedulers 1-3.
  const int lane = threadIdx.x & 31, hw = threadIdx.x >> 5;
  // hardware warp -> logical warp (tile warps 0..8, panel warps 9, 10, factor warp 11), one nibble per hardware warp.
  // Hardware warp h issues on scheduler h & 3: scheduler 0 holds the factor warp, the lighter panel warp (seven tiles)
  // and one tile warp; the other panel warp sits on scheduler 1, whose tile warps wait for the panel while it is formed.
  // Measured alternatives: both panel warps next to the factor warp 3.47k cycles per column, none 3.72k, this 3.28k.
  const int warp = (int)((0x7658432A109BULL >> (4 * hw)) & 15ULL);
  const int tau = 32 * warp + lane;                                 // logical thread id: tile warps first
  const bool is_factor = warp == kNW + kNP, is_tile = warp < kNW, is_panel = !is_factor && !is_tile;
  const int g = lane >> 2, q = lane & 3;
  const int M = cv.M, bw = cv.bw, ld = cv.ld, off = cv.off;
  const int NT8 = (M + 7) >> 3, Mp = NT8 * 8;
  const int side = twist ? (int)blockIdx.x : 0;
  // side 1 (reversed coordinates: its two elements of a fragment sit in different rows of S) runs ~6 % slower per tile
  // column than side 0; side 0 takes a few columns more so that both reach the hand-over together
  const int Jm0 = min(NT8 - 16, (NT8 - 16) / 2 + (NT8 - 16) / kTwistShift), Jm1 = NT8 - 16 - Jm0;
  const int c1 = twist ? (side ? Jm1 : Jm0) : NT8;                  // end of this side's first segment
  const int NTloc = twist ? c1 + 16 : NT8;                          // tiles this side ever sees (local coordinates)
  // This side's factor in "stage format": 128 doubles per row; row r of tile row T = r / 8 holds L(r, c) for the 120
  // columns c in [8 (T - 15), 8 T) at offset c - 8 T + 120, and W_T (the inverted diagonal tile, row r - 8 T) in the
  // last 8 slots. A tile row is one contiguous 8 KB block: the back substitution fetches it with ONE bulk copy.
  double *__restrict__ L = L_all + (size_t)side * Mp * kLs;
  double *z = dsm;                         // [Mp]   right-hand side -> forward solution -> solution
  double *dd = z + Mp;                     // [Mp]   damping ep + lm * S_rr, added when a diagonal tile is factored
  double *Psm = dd + Mp;                   // [2][16][64] panel tiles L_{J+i,J}, i = 1..15, operand layout, by column parity
  double *Asm = Psm + 2 * 16 * 64;         // [2][16][64] -A_{J+i,J} (what the panel warps turn into L), operand layout
  double *Dsm = Asm + 2 * 16 * 64;         // [8][kPs] diagonal tile handed to the factor warp
  double *Wsm = Dsm + 8 * kPs;             // [2][64]  -W = -L_JJ^-1, operand layout
  double *zJ = Wsm + 2 * 64;               // [2][8]
  double *Dnsm = zJ + 2 * 8;               // [2][64]  -D_{J+1} without column J's update, C layout [g][c] (diagonal 0 -> factor warp)
  double *Fsm = Dnsm + 2 * 64;             // [64]     factor warp: L_{J+1,J} in operand layout
  double *Hsm = Fsm + 64;                  // [64]     twist hand-over: -D_{c1} back from the factor warp, C layout
  double *Lst = Psm;                       // [R][8][128] back-substitution stages: OVERLAY everything from Psm on (dead by then)
  double *xsol = dd;                       // solution of the back substitution (dd is dead then)
  __shared__ int s_fail, s_nan, s_abort;
  __shared__ volatile int s_colfail[2];      // the failure flag of the column of each parity (the factor warp runs one column ahead)
  __shared__ __align__(8) unsigned long long s_mb[48];          // back substitution: full [0,16), xrdy [16,32), fdone [32,48)
  unsigned long long *const s_full = s_mb, *const s_xrdy = s_mb + 16, *const s_fdone = s_mb + 32;
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

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    if (tau == 0) {                                                // the back substitution's step numbers start at 0 in every attempt
      for (int k = 0; k < 48; ++k) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&s_mb[k]);
        if (attempt) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(k >= 32 ? kFdoneCount : (k >= 16 ? kXrdyCount : 1)) : "memory");
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
      const int need0 = need_of(min(17, NTloc - 1));
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
    const int rows_now = feed.flags ? min(Mp, 144) : Mp;             // tile rows 0..17; later rows are fetched as they enter
    for (int rl = tau; rl < rows_now; rl += kThreadsDg) load_row(rl);
    if (tau == 0) { s_fail = (feed.flags && s_giveup) ? 1 : 0; s_nan = 0; s_abort = 0; s_colfail[0] = s_colfail[1] = 0; }
    int *gf = gfl + 4 * attempt;                                   // [0] side 0 failed, [1] side 1 failed, [2] NaN
    __syncthreads();
    const int nsegs = twist ? 2 : 1;
    // a hand-over segment (segment 0 of a twisted solve) keeps the pipeline one column past its end: the exchange
    // needs the tiles the factor warp and the warp of diagonal 0 hold for column c1
    const int oc0 = ((2 * q) & 3) * 2 + ((2 * q) >> 2), oc1 = ((2 * q + 1) & 3) * 2 + ((2 * q + 1) >> 2);
    if (is_factor) {
      // =================== factor warp: the whole critical chain of the factorisation =============================
      //   [A]  8x8 Cholesky of the diagonal tile, W_J = L_JJ^-1, zJ                       (every lane redundantly)
      //   [L]  L_{J+1,J} = A_{J+1,J} W_J^T  (2 DMMAs; -A_{J+1,J} was shipped by the warp of diagonal 1 a column ago)
      //   [D]  D_{J+1} = D_{J+1} - L L^T + damping  (2 DMMAs; D_{J+1} with the updates of columns <= J-1 was shipped by
      //        the warp of diagonal 0 a column ago), z_{J+1} -= L zJ
      // so the next [A] never waits for the tile warps' DMMA queues: they only have to stay less than a column behind.
      for (int seg = 0; seg < nsegs; ++seg) {
        if (seg == 1) {                                            // twist hand-over (see the tile-warp branch)
          __syncthreads();
          cluster_sync();
          __syncthreads();
        }
        const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
        const bool handover = twist && seg == 0;
        bool stop = false;
        if (jb < je) bar_sync(kBarD, 64);                          // tile (jb,jb) (+ damping) is in Dsm
        for (int J = jb; J < je; ++J) {
          const int p = J & 1;
          BA_TR(8);
          if (feed.flags && wcur < feed.n_units) {
            if (probe_on) {
              const unsigned mk = __ballot_sync(0xffffffffu, fprobe == feed.epoch);
              const int adv = mk == 0xffffffffu ? 32 : __ffs(~mk) - 1;
              if (adv) { wcur += adv; __threadfence(); if (lane == 0) s_cursor = wcur; }
            }
            const int nd = J + 18 < NTloc ? need_of(J + 18) : 0;
            int spins = 0;
            while (wcur < nd && !s_giveup && ++spins < (spin_cap >> 4) + 8) poll();
            if (wcur < nd && lane == 0) s_giveup = 1;
            fprobe = 0;
            if (wcur + lane < feed.n_units) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(fprobe) : "l"(feed.flags + wcur + lane) : "memory");
            probe_on = true;
          }
          if (feed.flags && (s_giveup || s_fail)) {                // the producer is not there: leave like a failed pivot
            __syncwarp();
            if (lane == 0) { s_fail = 1; s_colfail[p] = 1; }
            __syncwarp();
            bar_arrive(kBarW + p, kCntW);
            bar_sync(kBarA + p, kCntA);                            // keep the count of this column's barrier whole
            stop = true;
            break;
          }
          double a[36];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; j += 2) {                      // 16-byte broadcast loads (rows are 96 bytes apart)
              const double2 v = *reinterpret_cast<const double2 *>(Dsm + i * kPs + j);
              a[tri8(i, j)] = v.x;
              if (j + 1 <= i) a[tri8(i, j + 1)] = v.y;
            }
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
          double *wdst = lane < 8 ? Wsm + 64 * p + (lane & 3) * 2 + (lane >> 2) : zJ + 8 * p;
          const int wstep = lane < 8 ? 8 : 1;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double piv = a[tri8(k, k)];
            ok = ok && ((float)piv > 0.0f);
            double y;                                              // seed from the high word (~2^-20) + one Newton step
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
            const double inv = fma(fma(-piv * y, 0.5 * y, 0.5), y, y);
            double sv = lane == 8 ? zr[k] : (lane == k ? -1.0 : 0.0);   // lanes 0-7 solve for the columns of -W
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
          // lanes 0-7: column `lane` of -W (operand layout); lane 8: zJ. One predicated store per row, no divergence; a
          // failed column publishes garbage nobody reads (the flag goes with it)
          if (lane <= 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) wdst[i * wstep] = wv[i];
          }
          if (lane == 0) { s_colfail[p] = ok ? 0 : 1; if (!ok) s_fail = 1; }
          __syncwarp();
          BA_TR(10);
          bar_arrive(kBarW + p, kCntW);                            // W_J, zJ (or the failure flag) published
          if (ok && lane == 8) {                                   // the forward solution of the block (the back substitution reads it)
#pragma unroll
            for (int i = 0; i < 8; ++i) z[8 * J + i] = wv[i];
          }
          bar_sync(kBarA + p, kCntA);                              // -A_{J+1,J} is in Asm[p] (shipped a column ago)
          if (!ok) { stop = true; break; }
          if (J + 1 < je || handover) {
            // ---- [L], [D]: the chain to the next diagonal tile stays inside this warp ----
            const int pn = (J + 1) & 1;
            bar_sync(kBarN + pn, 64);                              // -D_{J+1} (updates of columns <= J-1) is in Dnsm[pn]
            const double2 wb = *reinterpret_cast<const double2 *>(Wsm + 64 * p + 2 * lane);
            const double2 af = *reinterpret_cast<const double2 *>(Asm + (16 * p + 1) * 64 + 2 * lane);
            double l0, l1;
            dmma884(l0, l1, af.x, wb.x, 0.0, 0.0);
            dmma884(l0, l1, af.y, wb.y, l0, l1);                   // L_{J+1,J}, C layout
            // L L^T sums over the columns of L in any order: with columns 0,2,4,6 in the first k-step and 1,3,5,7 in the
            // second, the C fragment of [L] (elements 2q, 2q+1 of row g) is both the A and the B operand — no trip
            // through shared memory
            double n0 = Dnsm[64 * pn + g * 8 + 2 * q], n1 = Dnsm[64 * pn + g * 8 + 2 * q + 1];
            dmma884(n0, n1, l0, l0, n0, n1);
            dmma884(n0, n1, l1, l1, n0, n1);                       // -D_{J+1} with the update of column J
            double part = l0 * zJ[8 * p + 2 * q] + l1 * zJ[8 * p + 2 * q + 1];
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0 && J + 1 < NTloc) z[8 * (J + 1) + g] -= part;   // z_{J+1} -= L zJ (the panel warps skip d = 1)
            if (J + 1 < je) {
              const double dmp = dd[8 * (J + 1) + g];
              Dsm[g * kPs + 2 * q] = (2 * q == g ? dmp : 0.0) - n0;
              Dsm[g * kPs + 2 * q + 1] = (2 * q + 1 == g ? dmp : 0.0) - n1;
            } else {                                               // hand-over: back to the warp of diagonal 0
              Hsm[g * 8 + 2 * q] = n0; Hsm[g * 8 + 2 * q + 1] = n1;
            }
            __syncwarp();
          }
        }
        (void)stop;                                                // a failed side still meets the hand-over barriers of segment 1
      }
    } else if (is_panel) {
      // =================== panel warps: L_dJ = A_dJ W_J^T for d = 1..15 (eight and seven tiles), the right-hand side
      //                     update z_{J+d} -= L_dJ zJ, and the factor's way to global memory — while the tile warps are
      //                     still busy with the trailing update of column J-1 ==========================================
      const int pw = warp - kNW;
      for (int seg = 0; seg < nsegs; ++seg) {
        if (seg == 1) {                                            // twist hand-over (see the tile-warp branch)
          __syncthreads();
          cluster_sync();
          __syncthreads();
        }
        const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
        for (int J = jb; J < je; ++J) {
          const int p = J & 1;
          if (pw == 0) BA_TR(4);
          bar_sync(kBarA + p, kCntA);                              // -A_{dJ} of every diagonal is in Asm[p]
          if (pw == 0) BA_TR(5);
          bar_sync(kBarW + p, kCntW);                              // W_J, zJ ready (or the column failed)
          const int sf = s_colfail[p];
          if (pw == 0) BA_TRD(6, sf);
          if (sf) { bar_arrive(kBarP + p, kCntP); break; }         // wake the tile warps: they leave through the flag too
          const double2 wb = *reinterpret_cast<const double2 *>(Wsm + 64 * p + 2 * lane);   // B[k][n] = -W[n][k], n = g, k = 4h + q
          const double zq0 = zJ[8 * p + 2 * q], zq1 = zJ[8 * p + 2 * q + 1];
          // loads, first k-step, second k-step, stores: the tiles are independent, keep them in flight together
          double2 af[kPanelTiles];
          double p0[kPanelTiles], p1[kPanelTiles];
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) {
            const int d = kPanelTiles * pw + 1 + k;
            af[k] = *reinterpret_cast<const double2 *>(Asm + (16 * p + (d > 15 ? 15 : d)) * 64 + 2 * lane);
          }
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) dmma884(p0[k], p1[k], af[k].x, wb.x, 0.0, 0.0);
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) dmma884(p0[k], p1[k], af[k].y, wb.y, p0[k], p1[k]);   // (-A) (-W)^T
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) {
            const int d = kPanelTiles * pw + 1 + k;
            if (d > 15) continue;
            double *ps = Psm + (16 * p + d) * 64 + g * 8;
            ps[oc0] = p0[k]; ps[oc1] = p1[k];
          }
          double part[kPanelTiles];
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) part[k] = p0[k] * zq0 + p1[k] * zq1;   // z_a -= L_aJ zJ
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 1);
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], 2);
          // z_{J+2} is read by the factor warp one column from now, with nothing but this barrier in between: d = 2 goes
          // before the arrive (d = 1 is the factor warp's own)
          if (pw == 0 && q == 0 && J + 2 < NTloc) z[8 * (J + 2) + g] -= part[1];
          bar_arrive(kBarP + p, kCntP);                            // the trailing update of column J may start
          if (pw == 0) BA_TR(7);
          // off the critical path: the rest of the right-hand side, factor -> global (stage format), W_J -> global
#pragma unroll
          for (int k = 0; k < kPanelTiles; ++k) {
            const int d = kPanelTiles * pw + 1 + k;
            if (d > 15) continue;
            if (J + d < NTloc) {                                   // tiles below this side's matrix are all zero
              const int r = 8 * (J + d) + g;
              if (q == 0 && d > 2) z[r] -= part[k];
              double *lp = L + (size_t)r * kLs + (2 * q - 8 * d + 120);   // L(r, 8J + 2q): even offset, one 16-byte store
              const int dl = 8 * d + g - 2 * q;
              if (dl <= bw) *reinterpret_cast<double2 *>(lp) = make_double2(p0[k], p1[k]);
              else if (dl - 1 <= bw) lp[1] = p1[k];
            }
          }
          if (pw == kNP - 1) {
            L[(size_t)(8 * J + (lane >> 3)) * kLs + 120 + (lane & 7)] = -Wsm[64 * p + op_idx(lane >> 3, lane & 7)];
            L[(size_t)(8 * J + 4 + (lane >> 3)) * kLs + 120 + (lane & 7)] = -Wsm[64 * p + op_idx(4 + (lane >> 3), lane & 7)];
          }
        }
      }
    } else if (is_tile) {
      // =================== tile warps ===========================================================================
      // Refill of diagonal d for tile row R = tile (R, R - d): the two elements of a lane sit at a fixed offset inside
      // the tile, so their band-storage indices advance by a constant per tile row and the in-band tests are loop
      // invariants; only "row < M" (side 0, last tile rows) is checked per row. Rows are fetched ONE COLUMN AHEAD of
      // their use (the tile that closes diagonal 15 is next column's panel tile right away).
      const int rstep = side ? -(8 * ld + 8) : (8 * ld + 8);
      const int rsec = side ? -ld : 1;                              // second element: next column (side 0) / previous row (side 1)
      auto refill_setup = [&](int d, int &ridx, bool &rin0, bool &rin1) {
        const int dl = 8 * d + g - 2 * q;                           // row - column of the lane's first element
        const int rl = g, cl = -8 * d + 2 * q;                      // tile row 0 (extrapolated)
        const int r0 = side ? Mp - 1 - cl : rl, c0g = side ? Mp - 1 - rl : cl;
        ridx = r0 * ld + c0g + off;
        rin0 = dl >= 0 && dl <= bw;
        rin1 = dl - 1 >= 0 && dl - 1 <= bw;
      };
      auto refill_fetch = [&](int R, int ridx, bool rin0, bool rin1, bool diag, double &v0, double &v1) {
        v0 = v1 = 0.0;
        if (R < NTloc) {
          const bool live = side || 8 * R + g < M;                  // side 0: rows >= M are identity padding
          const int i0 = ridx + R * rstep;
          if (rin0 && live) v0 = __ldcg(S + i0);
          if (rin1 && live) v1 = __ldcg(S + i0 + rsec);
          if (diag && !live) { v0 = (2 * q == g) ? 1.0 : 0.0; v1 = (2 * q + 1 == g) ? 1.0 : 0.0; }
        }
      };
      // ---- warp 0: diagonal 0. Its tiles (J+j, J+j), j = 2..15, live in registers; the tile that reaches j = 2 gets
      //      column J's update first and goes to the factor warp (Dnsm), which owns it from there on. ----
      auto run_diag0 = [&]() {
        constexpr int NS = 14;                                      // slot s <-> window-relative (s + 2, s + 2)
        double ct[NS][2], dx[2][2];                                 // dx: tiles (jb, jb), (jb+1, jb+1) at a segment start
        {
          double v0, v1;
#pragma unroll
          for (int s2 = 0; s2 < NS; ++s2) {
            v0 = v1 = 0.0;
            if (s2 + 2 < NTloc) load_frag(s2 + 2, s2 + 2, v0, v1);
            ct[s2][0] = -v0; ct[s2][1] = -v1;
          }
          load_frag(0, 0, v0, v1); dx[0][0] = -v0; dx[0][1] = -v1;
          v0 = v1 = 0.0;
          if (1 < NTloc) load_frag(1, 1, v0, v1);
          dx[1][0] = -v0; dx[1][1] = -v1;
        }
        int ridx; bool rin0, rin1;
        refill_setup(0, ridx, rin0, rin1);
        double rf[2], rn[2], pend_z = 0.0, pend_d = 0.0, pnxt_z = 0.0, pnxt_d = 0.0;
        for (int seg = 0; seg < nsegs; ++seg) {
          if (seg == 1) {
            __syncthreads();
            if (!s_fail) {                                          // the two tiles the chain holds for column c1, c1 + 1
              dx[0][0] = Hsm[g * 8 + 2 * q]; dx[0][1] = Hsm[g * 8 + 2 * q + 1];
              dx[1][0] = Dnsm[64 * ((c1 + 1) & 1) + g * 8 + 2 * q]; dx[1][1] = Dnsm[64 * ((c1 + 1) & 1) + g * 8 + 2 * q + 1];
            }
            if (side == 1 && !s_fail) {
#pragma unroll
              for (int t = 0; t < NS + 2; ++t) {
                const int rl = 8 * (c1 + t) + g, cl = 8 * (c1 + t) + 2 * q;
                double *d = XD + (size_t)(rl - 8 * c1) * 128 + (cl - 8 * c1);
                const double t0 = t < 2 ? dx[t < 2 ? t : 0][0] : ct[t >= 2 ? t - 2 : 0][0], t1 = t < 2 ? dx[t < 2 ? t : 0][1] : ct[t >= 2 ? t - 2 : 0][1];
                d[0] = -t0 - Aval(rl, cl);
                d[1] = -t1 - Aval(rl, cl + 1);
              }
              const int i = lane;                                   // warp 0: rows 0..31 of the middle's right-hand side
              { const int r = Mp - 1 - (8 * c1 + i); XD[16384 + i] = z[8 * c1 + i] - (r < M ? __ldcg(cv.y + r) : 0.0); }
            }
            if (tau == 0) gf[side] = s_fail;
            cluster_sync();
            if (tau == 0) s_abort = side == 0 ? (gf[1] | s_fail) : s_fail;
            __syncthreads();
            if (side == 0 && !s_abort) {
#pragma unroll
              for (int t = 0; t < NS + 2; ++t) {
                const int r = 8 * (c1 + t) + g, c = 8 * (c1 + t) + 2 * q;
                const double *d = XD + (size_t)(Mp - 1 - c - 8 * Jm1) * 128 + (Mp - 1 - r - 8 * Jm1);
                if (t < 2) { dx[t < 2 ? t : 0][0] -= d[0]; dx[t < 2 ? t : 0][1] -= d[-128]; }
                else { ct[t >= 2 ? t - 2 : 0][0] -= d[0]; ct[t >= 2 ? t - 2 : 0][1] -= d[-128]; }
              }
              z[8 * c1 + lane] += XD[16384 + 127 - lane];
              bar_sync(kBarZ, 32 * kNW);                           // z of the middle complete before anybody reads it
            }
          }
          const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
          const bool handover = twist && seg == 0;
          if (jb < je) {                                            // pipeline prologue
            refill_fetch(jb + 16, ridx, rin0, rin1, true, rf[0], rf[1]);
            const double dmp = dd[8 * jb + g];                      // tile (jb,jb) + damping -> factor warp
            Dsm[g * kPs + 2 * q] = (2 * q == g ? dmp : 0.0) - dx[0][0];
            Dsm[g * kPs + 2 * q + 1] = (2 * q + 1 == g ? dmp : 0.0) - dx[0][1];
            bar_arrive(kBarD, 64);
            if (jb + 1 < je || handover) {                          // tile (jb+1,jb+1), no update yet
              Dnsm[64 * ((jb + 1) & 1) + g * 8 + 2 * q] = dx[1][0]; Dnsm[64 * ((jb + 1) & 1) + g * 8 + 2 * q + 1] = dx[1][1];
              bar_arrive(kBarN + ((jb + 1) & 1), 64);
            }
            bar_arrive(kBarA + (jb & 1), kCntA);                   // (nothing to ship: keeps the barrier's count)
          }
          for (int J = jb; J < je; ++J) {
            const int p = J & 1;
            BA_TR(0);
            if (feed.flags && J + 18 < NTloc) wait_cursor(need_of(J + 18));
            refill_fetch(J + 17, ridx, rin0, rin1, true, rn[0], rn[1]);
            if (feed.flags && lane < 8 && J + 18 < NTloc && 8 * (J + 18) >= 144) {
              const int rl = 8 * (J + 18) + lane, r = side ? Mp - 1 - rl : rl;
              pnxt_z = r < M ? __ldcg(cv.y + r) : 0.0;
              pnxt_d = r < M ? ep + lm * __ldcg(S + Sg(r, r)) : 0.0;
            }
            bar_sync(kBarP + p, kCntP);                            // all panel tiles L_{dJ} are in Psm[p]
            const int sf = s_colfail[p];
            BA_TRD(1, sf);
            if (sf) break;
            const bool have_next = J + 1 < je;
            double2 pf[16];
#pragma unroll
            for (int i = 2; i < 16; ++i) pf[i] = *reinterpret_cast<const double2 *>(Psm + (16 * p + i) * 64 + 2 * lane);
            // ---- tile (J+2,J+2) first: with column J's update it goes to the factor warp ----
            {
              double c0 = ct[0][0], c1v = ct[0][1];
              dmma884(c0, c1v, pf[2].x, pf[2].x, c0, c1v);
              dmma884(c0, c1v, pf[2].y, pf[2].y, c0, c1v);
              const bool wanted = J + 2 < je || (handover && J + 2 == je);       // the factor warp will wait for it
              if (wanted || (handover && J + 2 == je + 1)) {
                Dnsm[64 * (J & 1) + g * 8 + 2 * q] = c0; Dnsm[64 * (J & 1) + g * 8 + 2 * q + 1] = c1v;
              }
              if (wanted) bar_arrive(kBarN + (J & 1), 64);
            }
            if (have_next) {
              if (feed.flags && lane < 8 && J + 17 < NTloc && 8 * (J + 17) >= 144) {
                z[8 * (J + 17) + lane] = pend_z; dd[8 * (J + 17) + lane] = pend_d;    // rows first touched by column J + 2
              }
              bar_arrive(kBarA + ((J + 1) & 1), kCntA);
            }
            BA_TR(2);
            // ---- the rest of diagonal 0 + slide ----
#pragma unroll
            for (int s2 = 1; s2 < NS; ++s2) dmma884(ct[s2][0], ct[s2][1], pf[s2 + 2].x, pf[s2 + 2].x, ct[s2][0], ct[s2][1]);
#pragma unroll
            for (int s2 = 1; s2 < NS; ++s2) dmma884(ct[s2 - 1][0], ct[s2 - 1][1], pf[s2 + 2].y, pf[s2 + 2].y, ct[s2][0], ct[s2][1]);
            ct[NS - 1][0] = -rf[0]; ct[NS - 1][1] = -rf[1];
            rf[0] = rn[0]; rf[1] = rn[1];
            pend_z = pnxt_z; pend_d = pnxt_d;
            BA_TR(3);
          }
        }
      };
      // ---- warps 1..8: diagonal D1 = w and, for w >= 2, diagonal D2 = 17 - w: 15 tiles, 30 DMMAs per column ----
      auto run = [&](auto d1c) {
        constexpr int D1 = decltype(d1c)::value, D2 = D1 >= 2 ? 17 - D1 : -1;
        constexpr int N1 = 16 - D1, N2 = D2 >= 0 ? 16 - D2 : 0, NTL = N1 + N2;
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
        // the first tile of each diagonal is the next column's panel tile: ship -A to the panel warps (operand layout)
        auto ship = [&](int Jn) {
          double *as = Asm + 16 * (Jn & 1) * 64 + g * 8;
          as[D1 * 64 + oc0] = ct[0][0]; as[D1 * 64 + oc1] = ct[0][1];
          if (D2 >= 0) { as[(D2 >= 0 ? D2 : 0) * 64 + oc0] = ct[N1 < NTL ? N1 : 0][0]; as[(D2 >= 0 ? D2 : 0) * 64 + oc1] = ct[N1 < NTL ? N1 : 0][1]; }
          bar_arrive(kBarA + (Jn & 1), kCntA);
        };
        int ridx[2];
        bool rin0[2], rin1[2];
        refill_setup(D1, ridx[0], rin0[0], rin1[0]);
        refill_setup(D2 >= 0 ? D2 : 0, ridx[1], rin0[1], rin1[1]);
        auto fetch_row = [&](int R, double (&rfv)[2][2]) {
          refill_fetch(R, ridx[0], rin0[0], rin1[0], false, rfv[0][0], rfv[0][1]);
          rfv[1][0] = rfv[1][1] = 0.0;
          if (D2 >= 0) refill_fetch(R, ridx[1], rin0[1], rin1[1], false, rfv[1][0], rfv[1][1]);
        };
        double rf[2][2], rn[2][2];                                  // tile row J + 16 (in hand), J + 17 (in flight)
        for (int seg = 0; seg < nsegs; ++seg) {
          if (seg == 1) {
            // ---- twist hand-over. Side 1: what its eliminations did to the 16 middle tile columns (window minus
            //      the untouched matrix) and to the right-hand side goes to XD in its local coordinates. Side 0 adds
            //      it to its window and restarts the pipeline at column c1. ----
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
            cluster_sync();
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
              bar_sync(kBarZ, 32 * kNW);                           // z of the middle complete before anybody reads it
            }
          }
          const int jb = seg ? c1 : 0, je = seg ? ((side == 0 && !s_abort) ? NTloc : c1) : c1;
          if (jb < je) {                                            // pipeline prologue: column jb's A tiles
            fetch_row(jb + 16, rf);
            ship(jb);
          }
          for (int J = jb; J < je; ++J) {
            const int p = J & 1;
            // tile row J + 17 (one column ahead): in streaming mode its Schur units must be complete first
            if (feed.flags && J + 18 < NTloc) wait_cursor(need_of(J + 18));
            fetch_row(J + 17, rn);
            bar_sync(kBarP + p, kCntP);                            // all panel tiles L_{dJ} are in Psm[p]
            if (s_colfail[p]) break;
            const bool have_next = J + 1 < je;
            double2 pf[16];
#pragma unroll
            for (int i = 1; i < 16; ++i) pf[i] = *reinterpret_cast<const double2 *>(Psm + (16 * p + i) * 64 + 2 * lane);
            // ---- the tiles that become next column's panel tiles first: tile[0] <- L_{1+d} L_1^T + tile[1] ----
#pragma unroll
            for (int t = 0; t < NTL; ++t) {
              if (TJ(t) == 1) {
                dmma884(ct[t][0], ct[t][1], pf[1 + TD(t)].x, pf[1].x, ct[t][0], ct[t][1]);
                dmma884(ct[t - 1][0], ct[t - 1][1], pf[1 + TD(t)].y, pf[1].y, ct[t][0], ct[t][1]);
              }
            }
            if (N2 == 1) { ct[NTL - 1][0] = -rf[1][0]; ct[NTL - 1][1] = -rf[1][1]; }   // diagonal 15: the entering tile IS the next panel tile
            if (have_next) ship(J + 1);
            // ---- the rest of the trailing update + slide: tile[j-1] <- L_{j+d} L_j^T + tile[j], two passes ----
#pragma unroll
            for (int t = 0; t < NTL; ++t)
              if (TJ(t) >= 2) dmma884(ct[t][0], ct[t][1], pf[TJ(t) + TD(t)].x, pf[TJ(t)].x, ct[t][0], ct[t][1]);
#pragma unroll
            for (int t = 0; t < NTL; ++t)
              if (TJ(t) >= 2) dmma884(ct[t - 1][0], ct[t - 1][1], pf[TJ(t) + TD(t)].y, pf[TJ(t)].y, ct[t][0], ct[t][1]);
            ct[N1 - 1][0] = -rf[0][0]; ct[N1 - 1][1] = -rf[0][1];
            if (N2 >= 2) { ct[NTL - 1][0] = -rf[1][0]; ct[NTL - 1][1] = -rf[1][1]; }
#pragma unroll
            for (int k = 0; k < 2; ++k) { rf[k][0] = rn[k][0]; rf[k][1] = rn[k][1]; }
          }
        }
#undef TD
#undef TJ
      };
      switch (warp) {
        case 0: run_diag0(); break;
        case 1: run(std::integral_constant<int, 1>{}); break;
        case 2: run(std::integral_constant<int, 2>{}); break;
        case 3: run(std::integral_constant<int, 3>{}); break;
        case 4: run(std::integral_constant<int, 4>{}); break;
        case 5: run(std::integral_constant<int, 5>{}); break;
        case 6: run(std::integral_constant<int, 6>{}); break;
        case 7: run(std::integral_constant<int, 7>{}); break;
        default: run(std::integral_constant<int, 8>{}); break;
      }
    }
    __syncthreads();
    if (phase && tau == 0) phase[1] = clock64();
    const bool bad = s_fail || s_abort;
    const int ncols = (twist && side == 1) ? c1 : NTloc;            // tile columns of L (and W_J) this side owns

    // ---- backward substitution L^T x = z by tile rows, descending, in local coordinates ----
    // One CHAIN warp (the factor warp: alone on scheduler 0 now) carries the only dependent sequence,
    //       x_J = W_J^T (z_J - sum_{d=1..15} L_{J+d,J}^T x_{J+d}),
    // in registers — no CTA barrier, no shuffle per tile row. It keeps the NEAR field d = 1..4 itself (x_{J+1} is added
    // to the sums of rows J .. J-3 as soon as it exists); three HELPER warps subtract the FAR field d = 5..15 from z
    // five or more steps before the chain reads the row (a thread owns a row while d runs 15 -> 5: one store, no
    // read-modify-write races); one LOADER thread keeps R tile rows of the stage-format factor in flight (bulk copies).
    // Hand-shakes are mbarrier rings indexed by the step number n = Jhi - J: full (row landed), xrdy (x_J published),
    // fdone (far field of x_J applied).
    const int R = 1 << rlog, Rm = R - 1, Jhi = NTloc - 1;
    auto mba = [&](unsigned long long *b) { return (unsigned)__cvta_generic_to_shared(b); };
    auto mb_wait = [&](unsigned mb, unsigned par) {
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(par) : "memory");
      }
    };
    auto full_mb = [&](int n) { return mba(&s_full[n & Rm]); };
    auto ring_par = [&](int n) { return (unsigned)(n >> 4) & 1u; };
    // one bulk copy per tile row (8 rows x kLs doubles, contiguous in global memory: the padding travels with the rows;
    // eight copies of one row each cost the issuing thread ~60 cycles apiece and throttle the chain, tools/microbench_chain2.cu)
    auto stage_issue = [&](int J, int n) {
      const unsigned mb = full_mb(n);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(Lst) + (unsigned)(n & Rm) * (unsigned)kStSlot;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(kStSlot) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(L + (size_t)J * (8 * kLs)), "r"(kStSlot), "r"(mb) : "memory");
    };
    asm volatile("fence.proxy.async;" ::: "memory");                // this CTA's stores to L (and to the shared memory the
    __threadfence_block();                                          // stages overlay) -> visible to the bulk copies
    bool anybad = bad;
    // x_mid travels through global memory behind a release / acquire flag (gf[3]: 1 = there, 2 = side 0 failed): side 0
    // publishes it from inside its chain, and a split cluster barrier (arrive early, wait late) there turned out to hang
    // under compute-sanitizer, which appears to block at the arrive
    if (twist && side == 1) {                                      // wait for x_mid
      if (tau == 0) {
        int f = 0;
        do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(gf + 3) : "memory"); if (!f) __nanosleep(64); } while (!f);
      }
      __syncthreads();
      anybad = bad || __ldcg(gf + 3) != 1 || __ldcg(gf + 0) != 0 || __ldcg(gf + 1) != 0;
      if (!anybad) {
        for (int i = tau; i < 128; i += kThreadsDg) xsol[8 * c1 + i] = __ldcg(XD + 16384 + 128 + (Mp - 1 - (8 * c1 + i) - 8 * Jm0));
      }
      __syncthreads();
    }
    bool synced2 = !(twist && side == 0);                          // side 0 owes side 1 the flag (x_mid hand-over)
    if (!anybad) {
      __syncthreads();
      constexpr int it0 = 0;
      const bool mid_out = !synced2;                               // side 0 of a twisted solve publishes the middle
      if (is_factor) {
        // =================== chain warp ===================
        // A lone warp issues an instruction every ~6 cycles here, so the step is written for instruction count. Every
        // 8x8 product is a DMMA pair with the vector as the A operand (all eight rows equal) and the tile as B; with
        // the tile's rows fetched in the order 0,2,4,6 | 1,3,5,7 the C fragment a lane gets back (elements 2q, 2q+1)
        // IS the A operand of the next product: no shuffle, no layout change anywhere on the chain. The step is
        // unrolled over the ring position k = n & 15, so every shared address is a base register + an immediate.
        // Near field as a register pipeline: a0 / a1 / a2 hold the sums of rows J-1 / J-2 / J-3.
        auto chain = [&](auto rlc) {
          constexpr int RL = decltype(rlc)::value, RM = (1 << RL) - 1;
          const unsigned mbb = (unsigned)__cvta_generic_to_shared(s_mb);
          const unsigned stg = (unsigned)__cvta_generic_to_shared(Lst) + (unsigned)(2 * q * kStRow + g * 8);
          unsigned zq = (unsigned)__cvta_generic_to_shared(z) + (unsigned)(8 * Jhi + 2 * q) * 8u;   // z pair of the round's first row
          unsigned xq = (unsigned)__cvta_generic_to_shared(xsol) + (unsigned)(8 * Jhi + 2 * q) * 8u;
          unsigned parR = 0;                                       // parity of the round: (n >> 4) & 1
          bool later = false;                                      // round > 0
          int Jr = Jhi;
          double x0 = 0.0, x1 = 0.0, a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
          double b00 = 0.0, b01 = 0.0, b10 = 0.0, b11 = 0.0, b20 = 0.0, b21 = 0.0, w0, w1;
          long long w_far = 0, w_row = 0;
          mbar_wait_o<0>(mbb, 0);
          w0 = lds64o<120 * 8>(stg); w1 = lds64o<120 * 8 + kStRow>(stg);
          // CK: the checked form (first round: rows without a far field yet, side 1's given rows, the hand-over of the
          // middle; last round: the end of the matrix). Rounds in between run the same step without any of the tests.
          auto step = [&](auto kc, auto ckc) -> bool {
            constexpr int k = decltype(kc)::value, k1 = (k + 1) & 15, kf = (k - kNear - 1) & 15;
            constexpr bool CK = decltype(ckc)::value;
            constexpr int so = (k & RM) * kStSlot, so1 = (k1 & RM) * kStSlot;   // stage of row J, of row J - 1
            const int J = Jr - k;
            const bool nxt = !CK || J > 0, far = !CK || k >= kNear + 1 || later;
            // parity of the row ring: sixteen stages turn once per round, eight or four several times per round
            const unsigned pf1 = RL == 4 ? (k == 15 ? parR ^ 1u : parR) : (unsigned)((k1 >> RL) & 1);
            const unsigned pfd = k >= kNear + 1 ? parR : parR ^ 1u;
            unsigned fl = 1, fd = 1;                               // probes first: their answers are back when they are needed
            if (nxt) fl = mbar_test_o<8 * (k1 & RM)>(mbb, pf1);
            if (far) fd = mbar_test_o<8 * (32 + kf)>(mbb, pfd);
            double t0, t1, e0, e1;
            dmma884(e0, e1, x0, b00, a00, a01); dmma884(t0, t1, x1, b01, e0, e1);       // row J: near field complete
            dmma884(e0, e1, x0, b10, a10, a11); dmma884(a00, a01, x1, b11, e0, e1);     // row J - 1
            if (!(fd & fl)) {
              const long long w0c = phase ? clock64() : 0;
              if (far) mbar_wait_o<8 * (32 + kf)>(mbb, pfd);
              const long long w1c = phase ? clock64() : 0;
              if (nxt) mbar_wait_o<8 * (k1 & RM)>(mbb, pf1);
              if (phase) { w_far += w1c - w0c; w_row += clock64() - w1c; }
            }
            const double2 zz = lds128o<-64 * k>(zq);
            t0 = zz.x - t0; t1 = zz.y - t1;
            double xn0, xn1;
            dmma884(e0, e1, t0, w0, 0.0, 0.0); dmma884(xn0, xn1, t1, w1, e0, e1);       // x_J = W_J^T t
            dmma884(e0, e1, x0, b20, 0.0, 0.0); dmma884(a10, a11, x1, b21, e0, e1);     // row J - 2
            if (nxt) {                               // next step's operands: W of row J - 1, near tiles of row J
              w0 = lds64o<so1 + 120 * 8>(stg); w1 = lds64o<so1 + 120 * 8 + kStRow>(stg);
              b00 = lds64o<so + 112 * 8>(stg); b01 = lds64o<so + 112 * 8 + kStRow>(stg);
              b10 = lds64o<so + 104 * 8>(stg); b11 = lds64o<so + 104 * 8 + kStRow>(stg);
              b20 = lds64o<so + 96 * 8>(stg); b21 = lds64o<so + 96 * 8 + kStRow>(stg);
            }
            if (CK && J >= ncols) { const double2 xg = lds128o<-64 * k>(xq); xn0 = xg.x; xn1 = xg.y; }   // given (side 1: the middle)
            else if (g == 0) sts128o<-64 * k>(xq, xn0, xn1);
#ifdef BA_VERIFY_SYNC
            mbar_arrive_o<8 * (16 + k)>(mbb);
#else
            __syncwarp();
            if (lane == 0) mbar_arrive_o<8 * (16 + k)>(mbb);
#endif
            x0 = xn0; x1 = xn1;
            if (CK && mid_out && J == c1) {                        // middle solved: publish it, then carry on downwards
              __syncwarp();                                        // (lanes 0-3 wrote x, every lane reads it)
              for (int kk = lane; kk < 128; kk += 32) XD[16384 + 128 + kk] = xsol[8 * c1 + kk];
              __threadfence();
              __syncwarp();
              if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gf + 3), "r"(1) : "memory");
            }
            return CK && J == 0;
          };
#define BA_ROUND(CKV)                                                                                  \
          if (step(std::integral_constant<int, 0>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 1>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 2>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 3>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 4>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 5>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 6>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 7>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 8>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 9>{}, std::integral_constant<bool, CKV>{})) break;      \
          if (step(std::integral_constant<int, 10>{}, std::integral_constant<bool, CKV>{})) break;     \
          if (step(std::integral_constant<int, 11>{}, std::integral_constant<bool, CKV>{})) break;     \
          if (step(std::integral_constant<int, 12>{}, std::integral_constant<bool, CKV>{})) break;     \
          if (step(std::integral_constant<int, 13>{}, std::integral_constant<bool, CKV>{})) break;     \
          if (step(std::integral_constant<int, 14>{}, std::integral_constant<bool, CKV>{})) break;     \
          if (step(std::integral_constant<int, 15>{}, std::integral_constant<bool, CKV>{})) break;
          for (;;) {
            if (later && Jr >= 16) { BA_ROUND(false) }             // rows Jr .. Jr - 15 > 0, all with a far field
            else { BA_ROUND(true) }
            Jr -= 16; zq -= 1024u; xq -= 1024u; parR ^= 1u; later = true;
          }
#undef BA_ROUND
          if (phase && lane == 0) { phase[4] = w_far; phase[5] = w_row; }
        };
        if (rlog == 4) chain(std::integral_constant<int, 4>{});
        else if (rlog == 3) chain(std::integral_constant<int, 3>{});
        else chain(std::integral_constant<int, 2>{});
      } else if (tau < 96) {
        // =================== helper warps: far field d = kNear + 1 .. 15 (twelve tiles x eight rows = 96 threads) ========
        // x_J published implies row J of the factor has landed (the chain warp waited for it first). All addresses
        // advance incrementally: the warps must keep up with the chain warp.
        constexpr int kFar = 15 - kNear;                           // tiles of the far field = rows in flight per thread cycle
        const int sI = tau >> 3, i = tau & 7;
        int m = ((sI - Jhi + 15) % kFar + kFar) % kFar;            // row J - 15 + m is this thread's, d = 15 - m
        const unsigned mbx = (unsigned)__cvta_generic_to_shared(s_mb) + 128u;   // xrdy ring (fdone: + 128)
        const unsigned st0 = (unsigned)__cvta_generic_to_shared(Lst) + (unsigned)i * 8u;
        unsigned xa = (unsigned)__cvta_generic_to_shared(xsol) + (unsigned)(8 * Jhi) * 8u;
        unsigned slot = 0, ring = 0, par = 0;
        double facc = 0.0;
        long long w_x = 0;
        for (int J = Jhi; J >= 0; --J) {
          const int d = 15 - m;
          const unsigned st = st0 + slot * (unsigned)kStSlot + (unsigned)(120 - 8 * d) * 8u;
          { const long long t0 = phase ? clock64() : 0; mbar_wait_o<0>(mbx + ring * 8u, par); if (phase) w_x += clock64() - t0; }
#ifdef BA_VERIFY_SYNC
          mb_wait(full_mb(Jhi - J), (unsigned)((Jhi - J) >> rlog) & 1u);
#endif
          const double2 xa0 = lds128o<0>(xa), xb0 = lds128o<16>(xa), xc0 = lds128o<32>(xa), xd0 = lds128o<48>(xa);
          const double h0 = lds64o<0>(st), h1 = lds64o<kStRow>(st), h2 = lds64o<2 * kStRow>(st), h3 = lds64o<3 * kStRow>(st);
          const double h4 = lds64o<4 * kStRow>(st), h5 = lds64o<5 * kStRow>(st), h6 = lds64o<6 * kStRow>(st), h7 = lds64o<7 * kStRow>(st);
          const double c0 = fma(h1, xa0.y, h0 * xa0.x), c1s = fma(h3, xb0.y, h2 * xb0.x);
          const double c2 = fma(h5, xc0.y, h4 * xc0.x), c3 = fma(h7, xd0.y, h6 * xd0.x);
          facc += (c0 + c1s) + (c2 + c3);
          if (d == kNear + 1) {
            const int r = J - (kNear + 1);
            if (r >= 0) z[8 * r + i] -= facc;
            facc = 0.0;
          }
          m = m == kFar - 1 ? 0 : m + 1;
#ifdef BA_VERIFY_SYNC
          mbar_arrive_o<128>(mbx + ring * 8u);
#else
          __syncwarp();
          if (lane == 0) mbar_arrive_o<128>(mbx + ring * 8u);
#endif
          xa -= 64u;
          slot = (slot + 1) & (unsigned)Rm;
          ring = (ring + 1) & 15u;
          par ^= (ring == 0);
        }
        if (phase && tau == 0) phase[7] = w_x;
      } else if (tau == 96) {
        // =================== loader ===================
        for (int k = 0; k < R && Jhi - k >= 0; ++k) stage_issue(Jhi - k, it0 + k);
        for (int J = Jhi - R; J >= 0; --J) {
          const int n = it0 + (Jhi - J);                           // the slot held row J + R: the helpers are done with it
          mb_wait(mba(&s_fdone[(n - R) & 15]), ring_par(n - R));   // after x_{J+R}, the chain after step J + R - 1
          mb_wait(mba(&s_xrdy[(n - R + 1) & 15]), ring_par(n - R + 1));
          stage_issue(J, n);
        }
      }
      __syncthreads();
      if (mid_out) synced2 = true;
    }
    if (!synced2) {                                                // side 0 failed: tell side 1 (which waits for the flag)
      if (tau == 0) {
        gf[0] = 1;
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gf + 3), "r"(2) : "memory");
      }
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
  // bit 3: the streamed solve gave up waiting (mode 1) / this stand-by launch redid it (mode 2 only gets here then)
  if (tau == 0 && side == 0) cv.status[0] = status | (((feed.mode == 1 && s_giveup) || feed.mode == 2) ? 8 : 0);
  if (phase && tau == 0) phase[3] = clock64();
}

// shared memory: z, dd (alive to the end) + the larger of the factorisation's buffers and the R back-substitution stages
// that overlay them; R = 16, 8 or 4 tile rows, whatever fits (cfg3: 16, 1024 key frames: 16)
static size_t diag_fixed_bytes() { return ((size_t)4 * 16 * 64 + 8 * kPs + 2 * 64 + 2 * 8 + 4 * 64) * sizeof(double); }
static int diag_stage_log2(int M) {
  const size_t Mp = ((size_t)(M + 7) / 8) * 8;
  for (int lg = 4; lg > 2; --lg)
    if (2 * Mp * sizeof(double) + ((size_t)kStSlot << lg) <= 227 * 1024 - 1024) return lg;
  return 2;
}
size_t solve_diag_smem_bytes(int M) {
  const size_t Mp = ((size_t)(M + 7) / 8) * 8;
  return 2 * Mp * sizeof(double) + std::max(diag_fixed_bytes(), (size_t)kStSlot << diag_stage_log2(M));
}

int launch_solve_band_diag(const CallView &cv, int allow_retry, double *scratch, const SolveFeed &feed, cudaStream_t s) {
  const size_t Mp = ((size_t)(cv.M + 7) / 8) * 8;
  const int nt = (int)(Mp / 8);
  double *L_all = scratch, *XD = L_all + 2 * Mp * kLs;
  {                                     // never-written slots of L must read as zero: clear when the shape changes
    const long long key = ((long long)cv.M << 20) | cv.bw;
    if (feed.shape_key && *feed.shape_key != key) {
      BA_CUDA(cudaMemsetAsync(L_all, 0, 2 * Mp * kLs * sizeof(double), s));
      *feed.shape_key = key;
    }
  }
  int *gfl = reinterpret_cast<int *>(XD + 128 * 128 + 256);
  const int twist = nt >= (feed.twist_min > 0 ? feed.twist_min : 64) ? 1 : 0;
  BA_CUDA(cudaMemsetAsync(gfl, 0, 8 * sizeof(int), s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(twist ? 2 : 1);
  cfg.blockDim = dim3(kHwThreadsDg);
  cfg.dynamicSmemBytes = solve_diag_smem_bytes(cv.M);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = twist ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BA_CUDA(cudaLaunchKernelEx(&cfg, k_solve_band_diag, cv, allow_retry, L_all, XD, gfl, twist, feed.trace, feed, diag_stage_log2(cv.M)));
  BA_LAUNCH_CHECK();
  return BA_OK;
}

int solve_diag_prepare_device() {
  return cudaFuncSetAttribute(k_solve_band_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024) == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}

}  // namespace ba
