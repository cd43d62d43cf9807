// ba_kernels.cu — the BA iteration: edge pass, per-track Schur complement, reduced solve,
// back-substitution and retractions (reference: main/backend/ba.py:217-339 and the projective_ops /
// lietorch code it calls). See DESIGN.md for the data layout and the roofline of each kernel.
#include <cstdlib>
#include <cstring>

#include "ba_internal.h"
#include "ba_math.cuh"

namespace ba {

// Reduced-system accumulators are fp64: the per-edge math is fp32 like the reference's, but every sum
// that feeds the (ill-conditioned) reduced solve is carried in double — see DESIGN.md §Precision.
// fire-and-forget reduction (REDG): atomicAdd with an unused result sometimes compiles to ATOMG, whose response
// travels back through the crossbar and keeps the warp from retiring
__device__ __forceinline__ void red_add(double *addr, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// Ad(X)^T applied in double (R, t are the fp32 pair constants)
__device__ __forceinline__ void adjT_apply_d(const float *R, Vec3 t, const double *a, double *b) {
  const double tx = t.x, ty = t.y, tz = t.z;
  const double cx = ty * a[2] - tz * a[1], cy = tz * a[0] - tx * a[2], cz = tx * a[1] - ty * a[0];
  const double u0 = a[3] - cx, u1 = a[4] - cy, u2 = a[5] - cz;
  b[0] = (double)R[0] * a[0] + (double)R[3] * a[1] + (double)R[6] * a[2];
  b[1] = (double)R[1] * a[0] + (double)R[4] * a[1] + (double)R[7] * a[2];
  b[2] = (double)R[2] * a[0] + (double)R[5] * a[1] + (double)R[8] * a[2];
  b[3] = (double)R[0] * u0 + (double)R[3] * u1 + (double)R[6] * u2;
  b[4] = (double)R[1] * u0 + (double)R[4] * u1 + (double)R[7] * u2;
  b[5] = (double)R[2] * u0 + (double)R[5] * u1 + (double)R[8] * u2;
}

__device__ __forceinline__ bool pose_free(int pose, const CallView &c) {
  int a = pose - c.fixedp;          // ba.py:272-274 index shift; :33-39 range mask
  return a >= 0 && a < c.n;
}
__device__ __forceinline__ double *S_at(const CallView &c, int r, int col) {
  return c.S + (size_t)r * c.ld + col + c.off;
}

// =================================================================================================
// K1  edge pass.  One CTA per chunk (<= tc consecutive tracks of one pattern group, degree d <= 256).
//   thread <-> (track slot kappa, pattern position p); the (i,j) pair of a position is fixed, so
//   Gij / adjoint / intrinsics are per-position constants (computed once per CTA, kept in registers),
//   and Bjj, vj are accumulated in registers over the chunk's tracks. Ji = -Ad(Gij)^T Jj
//   (projective_ops.py:96) makes every i-side quantity a fixed linear image of the j-side one:
//       Bii = A Bjj A^T,  Bij = -A Bjj,  vi = -A vj,  Eik = -A Ejk,   A = Ad(Gij)^T,
//   applied once per position (B, v; in fp64 at the flush) or per edge (E).
//   E 6-vectors whose slot is fed by this edge alone are stored straight into the group's dense E rows
//   [track][slot*6+c]; the others (e.g. the source frame's slot, fed by every edge of the track) and
//   the C, w scalars are reduced per track through shared memory in a fixed order.
//   No per-edge index is read: ii/jj/kk are implied by (group pattern, position, track).
// =================================================================================================
constexpr int kAccComps = 27;     // per-thread accumulators: Bjj lower[21] vj[6]
constexpr int kPosFloats = 20;    // per-position constants: R[9] t[3] 1/fxi 1/fyi cxi cyi fxj fyj cxj cyj
constexpr int kFlushPos = 24;     // positions flushed per round (27+36+90 doubles each in the dead prefetch buffers)
constexpr int kFlushOuts = 90;    // Bjj[21] Bii[21] Bij[36] vj[6] vi[6]
constexpr int kStagePasses = 4;   // passes whose targets / weights are prefetched together (2 stages in flight)
constexpr size_t kEdgeSmemBytes = (size_t)(kAccComps + kPosFloats + 4) * kEdgeThreads * sizeof(float) +
                                  (size_t)2 * kStagePasses * 2 * kEdgeThreads * sizeof(float2);

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__constant__ unsigned char c_tri_a[21] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5};
__constant__ unsigned char c_tri_b[21] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5};

template <bool STRUCT_ONLY>
__global__ void __launch_bounds__(kEdgeThreads, 2) k_edge_pass(PlanView pv, CallView cv) {
  constexpr int NT = kEdgeThreads;
  extern __shared__ __align__(16) float dyn_smem[];
  float *sh = dyn_smem;                                 // [27 NT] staging during passes; accumulators / outputs at flush
  float *shc = sh + kAccComps * NT;                     // [20 NT] per-position constants
  float *spatch = shc + kPosFloats * NT;                // [4 NT]  (x, y, inverse depth, mono disparity) of the chunk's tracks
  float2 *sin = reinterpret_cast<float2 *>(spatch + 4 * NT);   // [2 stages][kStagePasses][2][NT] target / weight
  const int tau = threadIdx.x;
  const ChunkDesc cd = pv.cdesc[blockIdx.x];
  const int g = cd.g, pat0 = cd.pat0, d = cd.d;
  if (d > NT || cd.reg) return;                         // long tracks: k_edge_pass_long; regular groups: k_edge_pass_v2
  const int t0 = cd.t0, t1 = cd.t1, gt0 = cd.gt0, ebase = cd.ebase;
  const int sbase = 2 * pat0;
  const int *slot_pose = pv.slot_pose + sbase;
  float *Erows = cv.Est + cd.eoff;                      // entry-major: [6 W][Ts]
  const int Ts = cd.Ts;
  const int nm = cd.nm;
  const int *ms_ptr = pv.ms_ptr + sbase + g;
  const int *ms_slot = pv.ms_slot + sbase;
  const int R = cd.R;                                   // staged (multi-slot) items per track
  const int Tp = NT / d;                                // tracks per pass
  const bool active = tau < Tp * d;
  const int kappa = tau / d;
  const int p = tau - kappa * d;
  const int estride = (Tp * R) | 1;                     // floats between staged E components
  float *stE = sh;                                      // [6][estride]   (6*estride <= 12*NT + 6)
  float *stC = sh + 13 * NT;                            // [2][NT]

  // ---- the chunk's patches (gathered through kx once), and per-position constants, once per CTA ----
  for (int x = tau; x < t1 - t0; x += NT) {
    const size_t kp = (size_t)__ldg(pv.kx + t0 + x);
    const float *pp = cv.patches + 3 * kp;
    spatch[x] = __ldg(pp); spatch[NT + x] = __ldg(pp + 1); spatch[2 * NT + x] = __ldg(pp + 2);
    spatch[3 * NT + x] = cv.monodisp ? __ldg(cv.monodisp + kp) : 0.0f;
  }
  if (tau < d) {
    const int i = pv.pat_i[pat0 + tau], j = pv.pat_j[pat0 + tau];
    PairConst c = pair_const(cv.poses + 7 * i, cv.poses + 7 * j, cv.intr + 4 * i, cv.intr + 4 * j);
#pragma unroll
    for (int k = 0; k < 9; ++k) shc[k * NT + tau] = c.R[k];
    shc[9 * NT + tau] = c.t.x; shc[10 * NT + tau] = c.t.y; shc[11 * NT + tau] = c.t.z;
    shc[12 * NT + tau] = 1.0f / c.fxi; shc[13 * NT + tau] = 1.0f / c.fyi;
    shc[14 * NT + tau] = c.cxi; shc[15 * NT + tau] = c.cyi;
    shc[16 * NT + tau] = c.fxj; shc[17 * NT + tau] = c.fyj; shc[18 * NT + tau] = c.cxj; shc[19 * NT + tau] = c.cyj;
  }
  __syncthreads();
  PairConst pc;
  float ifxi = 0.f, ifyi = 0.f;
  int ri_rank = -1, rj_rank = -1, li = 0, lj = 0;
  bool fi = false, fj = false;
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) pc.R[k] = shc[k * NT + p];
    pc.t = {shc[9 * NT + p], shc[10 * NT + p], shc[11 * NT + p]};
    ifxi = shc[12 * NT + p]; ifyi = shc[13 * NT + p];
    pc.cxi = shc[14 * NT + p]; pc.cyi = shc[15 * NT + p];
    pc.fxj = shc[16 * NT + p]; pc.fyj = shc[17 * NT + p]; pc.cxj = shc[18 * NT + p]; pc.cyj = shc[19 * NT + p];
    if (!STRUCT_ONLY) {
      ri_rank = pv.pat_ri[pat0 + p]; rj_rank = pv.pat_rj[pat0 + p];
      li = pv.pat_li[pat0 + p]; lj = pv.pat_lj[pat0 + p];
      fi = pose_free(slot_pose[li], cv); fj = pose_free(slot_pose[lj], cv);   // ba.py:33-39 masks
    }
  }
  float acc[kAccComps];
#pragma unroll
  for (int k = 0; k < kAccComps; ++k) acc[k] = 0.0f;
  const int outs_per_track = STRUCT_ONLY ? 1 : 6 * nm + 1;

  // ---- targets / weights: each thread prefetches its own edges, kStagePasses passes per cp.async group,
  //      two groups in flight; the loads of a whole stage overlap instead of one round trip per pass ----
  const int npass = (t1 - t0 + Tp - 1) / Tp;
  const int nstage = (npass + kStagePasses - 1) / kStagePasses;
  auto issue_stage = [&](int st) {
    if (active) {
      float2 *buf = sin + (size_t)(st & 1) * kStagePasses * 2 * NT;
      for (int k = 0; k < kStagePasses; ++k) {
        const int t = t0 + (st * kStagePasses + k) * Tp + kappa;
        if (t < t1) {
          const int q = ebase + (t - gt0) * d + p;
          const int e = pv.perm_identity ? q : __ldg(pv.eperm + q);
          if (cv.tstride == 2) cp_async8(buf + (2 * k) * NT + tau, cv.targets + 2 * (size_t)e);
          else {
            const float *tp = cv.targets + (size_t)e * cv.tstride;
            float *dstf = reinterpret_cast<float *>(buf + (2 * k) * NT + tau);
            cp_async4(dstf, tp); cp_async4(dstf + 1, tp + 1);
          }
          cp_async8(buf + (2 * k + 1) * NT + tau, cv.weights + 2 * (size_t)e);
        }
      }
    }
    cp_async_commit();
  };
  issue_stage(0);

  for (int ps = 0; ps < npass; ++ps) {
    const int ts = t0 + ps * Tp;
    const int st = ps / kStagePasses, kk = ps - st * kStagePasses;
    if (kk == 0) {
      if (st + 1 < nstage) { issue_stage(st + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    }
    const int t = ts + kappa;
    if (active && t < t1) {
      const float2 *buf = sin + (size_t)(st & 1) * kStagePasses * 2 * NT;
      const float2 tg = buf[(2 * kk) * NT + tau], wg = buf[(2 * kk + 1) * NT + tau];
      const int tl = t - t0;
      const float ppx = spatch[tl], ppy = spatch[NT + tl], ppd = spatch[2 * NT + tl];
      EdgeTerms et;
      edge_terms(pc, ifxi, ifyi, ppx, ppy, ppd, tg.x, tg.y, wg.x, wg.y, cv.bounds, cv.loss, et);
      const float wz0 = et.w0 * et.Jz0, wz1 = et.w1 * et.Jz1;      // (w Jz)^T, ba.py:255
      stC[tau] = wz0 * et.Jz0 + wz1 * et.Jz1;                      // C term, ba.py:287
      stC[NT + tau] = wz0 * et.r0 + wz1 * et.r1;                   // w term, ba.py:292
      if (!STRUCT_ONLY) {
        float Ej[6], Ei[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          const float wa0 = et.w0 * et.Jj0[a], wa1 = et.w1 * et.Jj1[a];   // (w Jj)^T, ba.py:254
          Ej[a] = wa0 * et.Jz0 + wa1 * et.Jz1;                            // Ejk, ba.py:263
          acc[21 + a] += wa0 * et.r0 + wa1 * et.r1;                       // vj, ba.py:266
#pragma unroll
          for (int b = 0; b <= a; ++b) acc[tri(a, b)] += wa0 * et.Jj0[b] + wa1 * et.Jj1[b];  // Bjj, :260
        }
        adjT_apply(pc.R, pc.t, Ej, Ei);                                   // Eik = -A Ejk, ba.py:262
        float *col = Erows + (t - gt0);
        if (rj_rank >= 0) {
#pragma unroll
          for (int a = 0; a < 6; ++a) stE[a * estride + kappa * R + rj_rank] = Ej[a];
        } else {
#pragma unroll
          for (int a = 0; a < 6; ++a) col[(size_t)(6 * lj + a) * Ts] = fj ? Ej[a] : 0.0f;
        }
        if (ri_rank >= 0) {
#pragma unroll
          for (int a = 0; a < 6; ++a) stE[a * estride + kappa * R + ri_rank] = -Ei[a];
        } else {
#pragma unroll
          for (int a = 0; a < 6; ++a) col[(size_t)(6 * li + a) * Ts] = fi ? -Ei[a] : 0.0f;
        }
      }
    }
    __syncthreads();
    // ---- per-track reduction of the staged terms, fixed order ----
    const int ntr = min(Tp, t1 - ts);
    for (int o = tau; o < ntr * outs_per_track; o += NT) {
      const int k2 = o / outs_per_track;
      const int r = o - k2 * outs_per_track;
      const int t = ts + k2;
      if (!STRUCT_ONLY && r < 6 * nm) {
        const int ms = r / 6, comp = r - 6 * ms;
        const int s = ms_slot[ms];
        float sum = 0.0f;
        if (pose_free(slot_pose[s], cv)) {
          const float *src = stE + comp * estride + k2 * R;
          for (int x = ms_ptr[ms]; x < ms_ptr[ms + 1]; ++x) sum += src[x];
        }
        Erows[(size_t)(6 * s + comp) * Ts + (t - gt0)] = sum;
      } else {
        // C and w of the track (ba.py:287,292), then the damped inverse Q and the prior-adjusted w
        // (ba.py:296-311; BA: :184) — fused here so that no separate per-track kernel is needed
        const float *src = stC + k2 * d;
        float C = 0.0f, w = 0.0f;
        for (int x = 0; x < d; ++x) { C += src[x]; w += src[NT + x]; }
        const float lam = cv.lmbda_vec ? cv.lmbda_vec[t] : cv.lmbda;
        if (cv.monodisp) {
          const float md = spatch[3 * NT + (t - t0)];
          const float mk = md > 1e-2f ? 1.0f : 0.0f;
          C = C + mk * cv.alpha;
          C = C + lam;
          w = w - mk * cv.alpha * (spatch[2 * NT + (t - t0)] - md);
        } else {
          C = C + lam;
        }
        cv.Qw[t] = make_float2(1.0f / C, w);
      }
    }
    __syncthreads();
  }

  if (!STRUCT_ONLY) {
    // ---- flush: reduce Bjj / vj over kappa (fp64), map to the i side, scatter with fp64 atomics. Every phase is
    //      spread over the CTA: (position, component) sums, (position, column) products A Bjj, (position, row)
    //      products (A Bjj) A^T, then one atomic per thread and output. ----
#pragma unroll
    for (int k = 0; k < kAccComps; ++k) sh[k * NT + tau] = active ? acc[k] : 0.0f;
    cp_async_wait<0>();
    __syncthreads();
    double *sumD = reinterpret_cast<double *>(spatch);          // [kFlushPos][27]  (patches / prefetch buffers are dead)
    double *ABs = sumD + kFlushPos * kAccComps;                  // [kFlushPos][6][6]  A Bjj
    double *outD = ABs + kFlushPos * 36;                         // [kFlushPos][90]
    for (int pb = 0; pb < d; pb += kFlushPos) {
      const int np = min(kFlushPos, d - pb);
      for (int x = tau; x < np * kAccComps; x += NT) {
        const int pl = x / kAccComps, k = x - pl * kAccComps;
        double sv = 0.0;
        for (int kp = 0; kp < Tp; ++kp) sv += (double)sh[k * NT + kp * d + pb + pl];
        sumD[x] = sv;
      }
      __syncthreads();
      for (int x = tau; x < np * 7; x += NT) {                   // (position, column c of Bjj) and (position, vj)
        const int pl = x / 7, c = x - pl * 7, pos = pb + pl;
        float Rm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = shc[k * NT + pos];
        const Vec3 tt{shc[9 * NT + pos], shc[10 * NT + pos], shc[11 * NT + pos]};
        const double *Sm = sumD + pl * kAccComps;
        double *o = outD + pl * kFlushOuts;
        double col[6], res[6];
        if (c < 6) {
#pragma unroll
          for (int a = 0; a < 6; ++a) col[a] = Sm[a >= c ? tri(a, c) : tri(c, a)];   // column c of Bjj (= its row c)
          adjT_apply_d(Rm, tt, col, res);
#pragma unroll
          for (int a = 0; a < 6; ++a) { ABs[pl * 36 + a * 6 + c] = res[a]; o[42 + 6 * a + c] = -res[a]; }   // Bij = -A Bjj, ba.py:280
        } else {
#pragma unroll
          for (int a = 0; a < 6; ++a) col[a] = Sm[21 + a];
          adjT_apply_d(Rm, tt, col, res);
#pragma unroll
          for (int a = 0; a < 6; ++a) { o[78 + a] = col[a]; o[84 + a] = -res[a]; }   // vj ba.py:290; vi = -A vj ba.py:289
#pragma unroll
          for (int k = 0; k < 21; ++k) o[k] = Sm[k];                                  // Bjj, ba.py:282
        }
      }
      __syncthreads();
      for (int x = tau; x < np * 6; x += NT) {                   // (position, row a): Bii = (A Bjj) A^T, ba.py:279
        const int pl = x / 6, a = x - pl * 6, pos = pb + pl;
        float Rm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = shc[k * NT + pos];
        const Vec3 tt{shc[9 * NT + pos], shc[10 * NT + pos], shc[11 * NT + pos]};
        double rowv[6], res[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) rowv[c] = ABs[pl * 36 + a * 6 + c];
        adjT_apply_d(Rm, tt, rowv, res);
        double *o = outD + pl * kFlushOuts + 21;
        for (int b2 = 0; b2 <= a; ++b2) o[tri(a, b2)] = res[b2];
      }
      __syncthreads();
      for (int x = tau; x < np * kFlushOuts; x += NT) {
        const int pl = x / kFlushOuts, o = x - pl * kFlushOuts;
        const int pi = pv.pat_i[pat0 + pb + pl], pj = pv.pat_j[pat0 + pb + pl];
        const bool f_i = pose_free(pi, cv), f_j = pose_free(pj, cv);
        const int ri = 6 * (pi - cv.fixedp), rj = 6 * (pj - cv.fixedp);
        const double val = outD[x];
        if (o < 21) {
          if (f_j) red_add(S_at(cv, rj + c_tri_a[o], rj + c_tri_b[o]), val);
        } else if (o < 42) {
          if (f_i) red_add(S_at(cv, ri + c_tri_a[o - 21], ri + c_tri_b[o - 21]), val);
        } else if (o < 78) {
          if (f_i && f_j) {                                  // Bij (rows i, cols j) and Bji = Bij^T, ba.py:280-281
            const int a = (o - 42) / 6, b = (o - 42) - 6 * a;
            if (pi > pj) red_add(S_at(cv, ri + a, rj + b), val);
            else if (pi < pj) red_add(S_at(cv, rj + b, ri + a), val);
            else if (a >= b) red_add(S_at(cv, ri + a, ri + b), val + outD[pl * kFlushOuts + 42 + 6 * b + a]);
          }
        } else if (o < 84) {
          if (f_j) red_add(cv.y + rj + (o - 78), val);
        } else {
          if (f_i) red_add(cv.y + ri + (o - 84), val);
        }
      }
      __syncthreads();
    }
  }
}

// =================================================================================================
// K1v2  edge pass for REGULAR groups (every SLAM graph: all edges of a track leave its source frame and go to
//   distinct target frames).  lane <-> track, a warp walks pattern positions.  One CTA = 8 warps arranged as
//   KT track slices (32 consecutive tracks each) x KP position splits (KT * KP = 8, KP chosen by the plan from the
//   track length so that a warp walks ~8-10 positions and the machine sees enough warps on small graphs too).
//   * per track quantities (C, w, the source slot's E 6-vector, patch, prior) live in the lane's registers for
//     the whole walk: no per-track reduction and no CTA barrier inside the walk; the KP partial sums of a track
//     meet once, in shared memory, in fixed order;
//   * the (i, j) pair of a position is warp-uniform: Gij / intrinsics come from shared memory as broadcast loads;
//   * E is entry-major ([6W][Ts]): the 6 stores of a position are 128-byte coalesced rows;
//   * targets / weights of 2 positions x 32 tracks are staged per warp with cp.async, two slices in flight,
//     private to the warp (__syncwarp only);
//   * Bjj / vj of a position are summed over the 32 tracks through a padded shared-memory transpose (27 stores,
//     8 vector loads per lane) and kept per (track slice, position); after the walk the CTA adds the slices' sums
//     in fp64, maps them to the i side (Bii = A Bjj A^T, Bij = -A Bjj, vi = -A vj, A = Ad(Gij)^T) and issues the
//     fp64 atomics, with Bii / vi pre-summed over the positions (they all hit the same pose block).
// =================================================================================================
constexpr int kE2PosBlock = 32;                 // positions per block (constants / per-position sums staged per block) ...
constexpr int kE2PosBlockWide = 96;             // ... and on small graphs (>= 4 position splits, <= 2 track slices per CTA), where the
                                                // per-block set-up and flush (8 k cycles) would otherwise be paid three times per track
__host__ __device__ constexpr int e2_pos_block(int kp) { return kp >= 4 ? kE2PosBlockWide : kE2PosBlock; }
// TPL = tracks per lane. 1: lane <-> track. 2 (large graphs): a lane walks tracks t and t + 32 together — the per-position
// constants, the control flow and the 27-value transpose reduction are shared by two edges, and the two independent edge
// computations give the in-order warp instruction-level parallelism (the kernel is bound by instruction issue, not HBM).
constexpr int e2_slice(int tpl) { return tpl == 2 ? 2 : 4; }       // positions per cp.async slice
constexpr int kE2RedStride = 36;                // floats per row of the transpose buffer (conflict-free both ways)
constexpr int kE2AccStride = 28;
constexpr int e2_arr_f2(int tpl) { return 32 * tpl * (e2_slice(tpl) + 1); }            // float2 per staged array: [track][slice + 1]
constexpr int e2_warp_scratch(int tpl) { return kAccComps * kE2RedStride + 2 * 2 * 2 * e2_arr_f2(tpl); }   // floats: transpose buffer + 2 stages x (targets, weights)
constexpr int kE2WarpScratch = e2_warp_scratch(1);                  // the smaller of the two: what the flush scratch may count on
constexpr int kE2Threads = 32 * kEdge2Warps;
constexpr int kE2FlushPos = (kEdge2Warps * kE2WarpScratch * 4) / ((kAccComps + 36 + kFlushOuts) * 8) < kE2PosBlock
                                ? (kEdge2Warps * kE2WarpScratch * 4) / ((kAccComps + 36 + kFlushOuts) * 8) : kE2PosBlock;
// dynamic shared memory: per warp scratch, per track slice (KT = kEdge2Warps / KP) the per-position sums, constants
static size_t edge2_smem_bytes(int kp, int tpl) {
  const int PB = e2_pos_block(kp);
  return (size_t)(kEdge2Warps * e2_warp_scratch(tpl) + (kEdge2Warps / kp) * PB * kE2AccStride + PB * kPosFloats + 5 * PB + 8) * sizeof(float);
}
static_assert(kE2FlushPos >= 1, "flush scratch too small");
static_assert(kEdge2Warps * 8 * 64 <= kEdge2Warps * kE2WarpScratch, "per-track partials alias the scratch");

template <bool STRUCT_ONLY, int TPL>
__global__ void __launch_bounds__(kE2Threads, TPL == 2 ? 2 : 3) k_edge_pass_v2(PlanView pv, CallView cv) {
  constexpr int kE2Slice = e2_slice(TPL), kE2ArrF2 = e2_arr_f2(TPL), kWarpScratch = e2_warp_scratch(TPL);
  extern __shared__ __align__(16) float dyn_smem[];
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
  const int xc = blockIdx.x;
  const int g = pv.x_grp[xc];
  if (!pv.g_reg[g]) return;                                        // irregular group: generic kernels
  const int KP = pv.e2_kp;                                         // position splits; KT = kEdge2Warps / KP track slices
  const int ks = warp / KP, sp = warp - ks * KP;
  const int t0 = pv.x_t0[xc], t1 = pv.x_t0[xc + 1];
  const int gt0 = pv.g_t0[g];
  const int Ts = (pv.g_t0[g + 1] - gt0 + 3) & ~3;
  const int pat0 = pv.g_pat[g], d = pv.g_pat[g + 1] - pat0;
  const int ebase = pv.tptr[gt0];
  const int *slot_pose = pv.slot_pose + 2 * pat0;
  float *Eb = cv.Est + pv.g_eoff[g];
  const int li = pv.pat_li[pat0];                                  // the one source slot of the group
  const bool fi = pose_free(slot_pose[li], cv);                    // ba.py:33-39 masks

  float *scratch = dyn_smem;                                       // [warps][kE2WarpScratch]; fp64 flush scratch afterwards
  float *red = scratch + warp * kWarpScratch;                      // [27][36]
  float2 *stage = reinterpret_cast<float2 *>(red + kAccComps * kE2RedStride);   // [2][2][32][slice + 1]
  float *sacc_all = scratch + kEdge2Warps * kWarpScratch;          // [track slice][32][28]
  const int PB = e2_pos_block(KP);                                 // positions per block
  float *sacc = sacc_all + ks * (PB * kE2AccStride);
  float *sconst = sacc_all + (kEdge2Warps / KP) * (PB * kE2AccStride);          // [PB][20]
  int *slj = reinterpret_cast<int *>(sconst + PB * kPosFloats);                 // [PB + 1] target slot of the position (+ look-ahead)
  int *sfj = slj + PB + 1;                                                      // [PB] target pose free?
  int *spp = sfj + PB;                                                          // [PB] pattern position (edge offset in the track)
  int *spj = spp + PB;                                                          // [PB] target pose
  int *hlist = spj + PB;                                                        // [PB] block positions that start a slot run
  int *s_ctl = hlist + PB;                                                      // [0] positions in this block, [1] run heads

  const int tw0 = t0 + 32 * TPL * ks;                              // first track of this warp's slice (lane: tracks tw0 + lane [+ 32])
  const bool warp_has = tw0 < t1;
  int t[TPL];
  bool have[TPL];
  float ppx[TPL], ppy[TPL], ppd[TPL], C[TPL], w[TPL], Eis[TPL][6], Ejs[TPL][6];
#pragma unroll
  for (int u = 0; u < TPL; ++u) {
    t[u] = tw0 + 32 * u + lane;
    have[u] = t[u] < t1;
    ppx[u] = 0.0f; ppy[u] = 0.0f; ppd[u] = 1.0f;
    if (have[u]) {
      const float *pp = cv.patches + 3 * (size_t)__ldg(pv.kx + t[u]);
      ppx[u] = __ldg(pp); ppy[u] = __ldg(pp + 1); ppd[u] = __ldg(pp + 2);
    }
    C[u] = 0.0f; w[u] = 0.0f;
#pragma unroll
    for (int a = 0; a < 6; ++a) { Eis[u][a] = 0.0f; Ejs[u][a] = 0.0f; }
  }
  int hd = 0;
  const int nS = min(kEdge2Warps / KP, (t1 - t0 + 32 * TPL - 1) / (32 * TPL));   // track slices of this CTA that own tracks

  // The walk follows pat_ps: positions ordered by target slot, so that the positions feeding one E slot (a SLAM
  // graph observes a (patch, frame) pair several times) are consecutive; blocks of <= 32 positions and the KP
  // splits are cut at slot boundaries, and a lane adds the 6-vectors of a run in registers before the one store.
  const int *ps = pv.pat_ps + pat0;
  for (int pb = 0; pb < d;) {
    __syncthreads();                                               // previous block's flush is done with the scratch
    if (tau <= PB) {
      int lj = -1;
      if (pb + tau < d) { const int p = ps[pb + tau]; lj = pv.pat_lj[pat0 + p]; if (tau < PB) spp[tau] = p; }
      slj[tau] = lj;
    }
    __syncthreads();
    if (warp == 0) {
      int e = min(PB, d - pb);
      if (pb + e < d) { while (e > 1 && slj[e] == slj[e - 1]) --e; }           // runs are <= kMaxSlotRun < 32 long
      int nhd = 0;
      for (int x0 = 0; x0 < e; x0 += 32) {                                      // run heads, in order
        const int x = x0 + lane;
        const bool head = x < e && (x == 0 || slj[x] != slj[x - 1]);
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        if (head) hlist[nhd + __popc(hm & ((1u << lane) - 1u))] = x;
        nhd += __popc(hm);
      }
      if (lane == 0) { s_ctl[0] = e; s_ctl[1] = nhd; }
    }
    __syncthreads();
    const int np = s_ctl[0], nh = s_ctl[1];
    if (tau < np) {
      const int p = spp[tau];
      const int i = pv.pat_i[pat0 + p], j = pv.pat_j[pat0 + p];
      const PairConst c = pair_const(cv.poses + 7 * i, cv.poses + 7 * j, cv.intr + 4 * i, cv.intr + 4 * j);
      float *o = sconst + tau * kPosFloats;
#pragma unroll
      for (int k = 0; k < 9; ++k) o[k] = c.R[k];
      o[9] = c.t.x; o[10] = c.t.y; o[11] = c.t.z;
      o[12] = 1.0f / c.fxi; o[13] = 1.0f / c.fyi; o[14] = c.cxi; o[15] = c.cyi;
      o[16] = c.fxj; o[17] = c.fyj; o[18] = c.cxj; o[19] = c.cyj;
      spj[tau] = j;
      sfj[tau] = pose_free(j, cv) ? 1 : 0;
    }
    __syncthreads();
    // this warp's positions of the block: [pw0, pw1), cut at slot boundaries
    const int per = (np + KP - 1) / KP;
    int pw0 = min(np, sp * per), pw1 = min(np, (sp + 1) * per);
    while (pw0 > 0 && pw0 < np && slj[pw0] == slj[pw0 - 1]) ++pw0;
    while (pw1 > 0 && pw1 < np && slj[pw1] == slj[pw1 - 1]) ++pw1;
    if (warp_has && pw0 < pw1) {
      const int nslice = (pw1 - pw0 + kE2Slice - 1) / kE2Slice;
      auto issue = [&](int sl) {
        float2 *bt = stage + (sl & 1) * (2 * kE2ArrF2);
        const int pp = lane & (kE2Slice - 1);
        const int pl = pw0 + kE2Slice * sl + pp;
        if (pl < pw1) {
#pragma unroll
          for (int k = 0; k < TPL * kE2Slice; ++k) {
            const int tk = lane / kE2Slice + (32 / kE2Slice) * k, tt = tw0 + tk;
            if (tt < t1) {
              const int q = ebase + (tt - gt0) * d + spp[pl];
              const int e = pv.perm_identity ? q : __ldg(pv.eperm + q);
              float2 *dt = bt + tk * (kE2Slice + 1) + pp;
              if (cv.tstride == 2) cp_async8(dt, cv.targets + 2 * (size_t)e);
              else {
                const float *tp = cv.targets + (size_t)e * cv.tstride;
                cp_async4(reinterpret_cast<float *>(dt), tp); cp_async4(reinterpret_cast<float *>(dt) + 1, tp + 1);
              }
              cp_async8(dt + kE2ArrF2, cv.weights + 2 * (size_t)e);
            }
          }
        }
        cp_async_commit();
      };
      issue(0);
      for (int sl = 0; sl < nslice; ++sl) {
        if (sl + 1 < nslice) { issue(sl + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncwarp();
        const float2 *bt = stage + (sl & 1) * (2 * kE2ArrF2) + lane * (kE2Slice + 1);
        const int npp = min(kE2Slice, pw1 - pw0 - kE2Slice * sl);
        for (int pp = 0; pp < npp; ++pp) {
          const int pl = pw0 + kE2Slice * sl + pp;                 // position within the block (warp-uniform)
          const float4 *cs = reinterpret_cast<const float4 *>(sconst + pl * kPosFloats);
          const float4 c0 = cs[0], c1 = cs[1], c2 = cs[2], c3 = cs[3], c4 = cs[4];
          PairConst pc;
          pc.R[0] = c0.x; pc.R[1] = c0.y; pc.R[2] = c0.z; pc.R[3] = c0.w;
          pc.R[4] = c1.x; pc.R[5] = c1.y; pc.R[6] = c1.z; pc.R[7] = c1.w;
          pc.R[8] = c2.x; pc.t = {c2.y, c2.z, c2.w};
          pc.cxi = c3.z; pc.cyi = c3.w; pc.fxj = c4.x; pc.fyj = c4.y; pc.cxj = c4.z; pc.cyj = c4.w;
          pc.fxi = 0.f; pc.fyi = 0.f;                              // unused: the inverses are passed separately
          EdgeTerms et[TPL];
#pragma unroll
          for (int u = 0; u < TPL; ++u) {
            float2 tg = bt[32 * u * (kE2Slice + 1) + pp], wg = bt[kE2ArrF2 + 32 * u * (kE2Slice + 1) + pp];
            if (!have[u]) { tg = make_float2(0.f, 0.f); wg = make_float2(0.f, 0.f); }
            edge_terms(pc, c3.x, c3.y, ppx[u], ppy[u], ppd[u], tg.x, tg.y, wg.x, wg.y, cv.bounds, cv.loss, et[u]);
            const float wz0 = et[u].w0 * et[u].Jz0, wz1 = et[u].w1 * et[u].Jz1;  // (w Jz)^T, ba.py:255
            C[u] += wz0 * et[u].Jz0 + wz1 * et[u].Jz1;             // ba.py:287
            w[u] += wz0 * et[u].r0 + wz1 * et[u].r1;               // ba.py:292
          }
          if (!STRUCT_ONLY) {
            // Jj0[1] and Jj1[0] are structural zeros (projective_ops.py:83-95): entries that involve component 0 /
            // component 1 have one term only, Bjj(1,0) vanishes (adding the exact zero the reference adds). The transpose
            // buffer gets the lane's sum over its TPL tracks.
            float Ej[TPL][6];
            {
              float s21 = 0.f, s22 = 0.f, s00 = 0.f, s11 = 0.f;
#pragma unroll
              for (int u = 0; u < TPL; ++u) {
                const float wa00 = et[u].w0 * et[u].Jj0[0], wa11 = et[u].w1 * et[u].Jj1[1];   // (w Jj)^T, ba.py:254
                Ej[u][0] = wa00 * et[u].Jz0; Ej[u][1] = wa11 * et[u].Jz1;                     // Ejk, ba.py:263
                s21 += wa00 * et[u].r0; s22 += wa11 * et[u].r1;                               // vj, ba.py:266
                s00 += wa00 * et[u].Jj0[0]; s11 += wa11 * et[u].Jj1[1];                       // Bjj, ba.py:260
              }
              red[21 * kE2RedStride + lane] = s21;
              red[22 * kE2RedStride + lane] = s22;
              red[tri(0, 0) * kE2RedStride + lane] = s00;
              red[tri(1, 0) * kE2RedStride + lane] = 0.0f;
              red[tri(1, 1) * kE2RedStride + lane] = s11;
#pragma unroll
              for (int a = 2; a < 6; ++a) {
                float wa0[TPL], wa1[TPL];
                float sv = 0.f, sa0 = 0.f, sa1 = 0.f;
#pragma unroll
                for (int u = 0; u < TPL; ++u) {
                  wa0[u] = et[u].w0 * et[u].Jj0[a]; wa1[u] = et[u].w1 * et[u].Jj1[a];
                  Ej[u][a] = fmaf(wa1[u], et[u].Jz1, wa0[u] * et[u].Jz0);
                  sv += fmaf(wa1[u], et[u].r1, wa0[u] * et[u].r0);
                  sa0 += wa0[u] * et[u].Jj0[0];
                  sa1 += wa1[u] * et[u].Jj1[1];
                }
                red[(21 + a) * kE2RedStride + lane] = sv;
                red[tri(a, 0) * kE2RedStride + lane] = sa0;
                red[tri(a, 1) * kE2RedStride + lane] = sa1;
#pragma unroll
                for (int b = 2; b <= a; ++b) {
                  float sb = 0.f;
#pragma unroll
                  for (int u = 0; u < TPL; ++u) sb += fmaf(wa1[u], et[u].Jj1[b], wa0[u] * et[u].Jj0[b]);
                  red[tri(a, b) * kE2RedStride + lane] = sb;
                }
              }
            }
            const int lj = slj[pl];
            const bool head = pl == pw0 || lj != slj[pl - 1];                   // first position of a slot run
            const bool run_end = pl + 1 == pw1 || slj[pl + 1] != lj;            // last position of the run: the one store
#pragma unroll
            for (int u = 0; u < TPL; ++u) {
              float Ei[6];
              adjT_apply(pc.R, pc.t, Ej[u], Ei);                                // Eik = -A Ejk, ba.py:262
#pragma unroll
              for (int a = 0; a < 6; ++a) Eis[u][a] -= Ei[a];
              if (lj == li) {                                                   // self edge: its j side feeds the source slot too
#pragma unroll
                for (int a = 0; a < 6; ++a) Eis[u][a] += Ej[u][a];
              } else {
                if (head) {
#pragma unroll
                  for (int a = 0; a < 6; ++a) Ejs[u][a] = Ej[u][a];
                } else {
#pragma unroll
                  for (int a = 0; a < 6; ++a) Ejs[u][a] += Ej[u][a];
                }
                if (have[u] && run_end) {
                  const bool fj = sfj[pl] != 0;
                  float *col = Eb + (size_t)(6 * lj) * Ts + (t[u] - gt0);
#pragma unroll
                  for (int a = 0; a < 6; ++a) col[(size_t)a * Ts] = fj ? Ejs[u][a] : 0.0f;
                }
              }
            }
            if (head) hd = pl;
            __syncwarp();
            if (lane < kAccComps) {                                             // sum of component `lane` over the 32 tracks
              const float4 *row = reinterpret_cast<const float4 *>(red + lane * kE2RedStride);
              float4 s4 = row[0];
#pragma unroll
              for (int k = 1; k < 8; ++k) { const float4 v = row[k]; s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w; }
              const float sv = (s4.x + s4.y) + (s4.z + s4.w);                   // positions of one (i, j) pair share a row
              sacc[hd * kE2AccStride + lane] = head ? sv : sacc[hd * kE2AccStride + lane] + sv;
            }
            __syncwarp();
          }
        }
        __syncwarp();                                              // slice buffer free before it is refilled
      }
    }
    if (!STRUCT_ONLY) {
      // ---- flush of this block of positions: add the slices' sums (fp64), map to the i side, fp64 atomics ----
      __syncthreads();
      constexpr int NT = kE2Threads;
      double *sumD = reinterpret_cast<double *>(scratch);          // [kE2FlushPos][27]
      double *ABs = sumD + kE2FlushPos * kAccComps;                // [kE2FlushPos][6][6]  A Bjj
      double *outD = ABs + kE2FlushPos * 36;                       // [kE2FlushPos][90]
      const int pi = slot_pose[li];
      const int ri = 6 * (pi - cv.fixedp);
      for (int fb = 0; fb < nh; fb += kE2FlushPos) {              // over the run heads: one (i, j) pair each
        const int nq = min(kE2FlushPos, nh - fb);
        for (int x = tau; x < nq * kAccComps; x += NT) {
          const int pl = x / kAccComps, k = x - pl * kAccComps;
          double sv = 0.0;
          for (int ww = 0; ww < nS; ++ww) sv += (double)sacc_all[(ww * PB + hlist[fb + pl]) * kE2AccStride + k];
          sumD[x] = sv;
        }
        __syncthreads();
        for (int x = tau; x < nq * 7; x += NT) {                   // (position, column c of Bjj) and (position, vj)
          const int pl = x / 7, c = x - pl * 7;
          const float *cs = sconst + hlist[fb + pl] * kPosFloats;
          const Vec3 tt{cs[9], cs[10], cs[11]};
          const double *Sm = sumD + pl * kAccComps;
          double *o = outD + pl * kFlushOuts;
          double col[6], res[6];
          if (c < 6) {
#pragma unroll
            for (int a = 0; a < 6; ++a) col[a] = Sm[a >= c ? tri(a, c) : tri(c, a)];   // column c of Bjj (= its row c)
            adjT_apply_d(cs, tt, col, res);
#pragma unroll
            for (int a = 0; a < 6; ++a) { ABs[pl * 36 + a * 6 + c] = res[a]; o[42 + 6 * a + c] = -res[a]; }   // Bij = -A Bjj, ba.py:280
          } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) col[a] = Sm[21 + a];
            adjT_apply_d(cs, tt, col, res);
#pragma unroll
            for (int a = 0; a < 6; ++a) { o[78 + a] = col[a]; o[84 + a] = -res[a]; }   // vj ba.py:290; vi = -A vj ba.py:289
#pragma unroll
            for (int k = 0; k < 21; ++k) o[k] = Sm[k];                                  // Bjj, ba.py:282
          }
        }
        __syncthreads();
        for (int x = tau; x < nq * 6; x += NT) {                   // (position, row a): Bii = (A Bjj) A^T, ba.py:279
          const int pl = x / 6, a = x - pl * 6;
          const float *cs = sconst + hlist[fb + pl] * kPosFloats;
          const Vec3 tt{cs[9], cs[10], cs[11]};
          double rowv[6], res[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) rowv[c] = ABs[pl * 36 + a * 6 + c];
          adjT_apply_d(cs, tt, rowv, res);
          double *o = outD + pl * kFlushOuts + 21;
#pragma unroll
          for (int b2 = 0; b2 < 6; ++b2) if (b2 <= a) o[tri(a, b2)] = res[b2];
        }
        __syncthreads();
        // Bjj, Bij (+ Bji), vj: one atomic per (position, entry); Bii, vi: summed over the positions first
        for (int x = tau; x < nq * kFlushOuts; x += NT) {
          const int pl = x / kFlushOuts, o = x - pl * kFlushOuts;
          if ((o >= 21 && o < 42) || o >= 84) continue;
          const int pj = spj[hlist[fb + pl]];
          const bool f_j = sfj[hlist[fb + pl]] != 0;
          const int rj = 6 * (pj - cv.fixedp);
          const double val = outD[x];
          if (o < 21) {
            if (f_j) red_add(S_at(cv, rj + c_tri_a[o], rj + c_tri_b[o]), val);
          } else if (o < 78) {
            if (fi && f_j) {                                 // Bij (rows i, cols j) and Bji = Bij^T, ba.py:280-281
              const int a = (o - 42) / 6, b = (o - 42) - 6 * a;
              if (pi > pj) red_add(S_at(cv, ri + a, rj + b), val);
              else if (pi < pj) red_add(S_at(cv, rj + b, ri + a), val);
              else if (a >= b) red_add(S_at(cv, ri + a, ri + b), val + outD[pl * kFlushOuts + 42 + 6 * b + a]);
            }
          } else {
            if (f_j) red_add(cv.y + rj + (o - 78), val);
          }
        }
        if (fi && tau < 27) {
          const int o = tau < 21 ? 21 + tau : 84 + (tau - 21);
          double sv = 0.0;
          for (int pl = 0; pl < nq; ++pl) sv += outD[pl * kFlushOuts + o];
          if (tau < 21) red_add(S_at(cv, ri + c_tri_a[tau], ri + c_tri_b[tau]), sv);
          else red_add(cv.y + ri + (tau - 21), sv);
        }
        __syncthreads();
      }
    }
    pb += np;
  }

  // ---- per track: the KP partial sums meet in shared memory (fixed order), then Q, w and the source slot's E ----
  if (KP > 1) {
    __syncthreads();
    float *tp = scratch + warp * (8 * 32 * TPL);                   // [TPL][8][32] per warp
#pragma unroll
    for (int u = 0; u < TPL; ++u) {
      float *tu = tp + u * (8 * 32);
      tu[lane] = C[u]; tu[32 + lane] = w[u];
#pragma unroll
      for (int a = 0; a < 6; ++a) tu[(2 + a) * 32 + lane] = Eis[u][a];
    }
    __syncthreads();
    if (sp == 0) {
      for (int k = 1; k < KP; ++k) {
#pragma unroll
        for (int u = 0; u < TPL; ++u) {
          const float *o = scratch + (warp + k) * (8 * 32 * TPL) + u * (8 * 32);
          C[u] += o[lane]; w[u] += o[32 + lane];
#pragma unroll
          for (int a = 0; a < 6; ++a) Eis[u][a] += o[(2 + a) * 32 + lane];
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < TPL; ++u) {
    if (have[u] && sp == 0) {
      if (!STRUCT_ONLY) {
        float *col = Eb + (size_t)(6 * li) * Ts + (t[u] - gt0);
#pragma unroll
        for (int a = 0; a < 6; ++a) col[(size_t)a * Ts] = fi ? Eis[u][a] : 0.0f;
      }
      // damped inverse Q and prior-adjusted w (ba.py:296-311; BA: :184)
      const float lam = cv.lmbda_vec ? cv.lmbda_vec[t[u]] : cv.lmbda;
      float Cc = C[u], ww = w[u];
      if (cv.monodisp) {
        const float md = __ldg(cv.monodisp + (size_t)__ldg(pv.kx + t[u]));
        const float mk = md > 1e-2f ? 1.0f : 0.0f;
        Cc = Cc + mk * cv.alpha;
        Cc = Cc + lam;
        ww = ww - mk * cv.alpha * (ppd[u] - md);
      } else {
        Cc = Cc + lam;
      }
      cv.Qw[t[u]] = make_float2(1.0f / Cc, ww);
    }
  }
}

// K1-long: tracks with more than 256 edges (do not occur in BA-Track's graphs, S_slam * steps <= 72;
// kept so that the operator is total). One thread per edge, float atomics into pre-zeroed E rows / Cw,
// fp64 atomics into S / y. Slow path.
template <bool STRUCT_ONLY>
__global__ void k_edge_pass_long(PlanView pv, CallView cv) {
  const int chunk = blockIdx.x;
  const int g = pv.c_grp[chunk];
  const int pat0 = pv.g_pat[g];
  const int d = pv.g_pat[g + 1] - pat0;
  if (d <= kEdgeThreads || pv.g_reg[g]) return;
  const int t0 = pv.c_t0[chunk], t1 = pv.c_t0[chunk + 1];
  const int gt0 = pv.g_t0[g];
  const int Ts = (pv.g_t0[g + 1] - gt0 + 3) & ~3;
  const int ebase = pv.tptr[gt0];
  const int *slot_pose = pv.slot_pose + 2 * pat0;
  float *Erows = cv.Est + pv.g_eoff[g];
  const long long total = (long long)(t1 - t0) * d;
  for (long long x = threadIdx.x; x < total; x += blockDim.x) {
    const int t = t0 + (int)(x / d), p = (int)(x % d);
    const int i = pv.pat_i[pat0 + p], j = pv.pat_j[pat0 + p];
    const PairConst pc = pair_const(cv.poses + 7 * i, cv.poses + 7 * j, cv.intr + 4 * i, cv.intr + 4 * j);
    const int q = ebase + (t - gt0) * d + p;
    const int e = pv.perm_identity ? q : pv.eperm[q];
    const float *tp = cv.targets + (size_t)e * cv.tstride, *wp = cv.weights + 2 * (size_t)e;
    const float *pp = cv.patches + 3 * (size_t)pv.kx[t];
    EdgeTerms et;
    edge_terms(pc, 1.0f / pc.fxi, 1.0f / pc.fyi, pp[0], pp[1], pp[2], tp[0], tp[1], wp[0], wp[1], cv.bounds, cv.loss, et);
    const float wz0 = et.w0 * et.Jz0, wz1 = et.w1 * et.Jz1;
    atomicAdd(reinterpret_cast<float *>(cv.Cw + t), wz0 * et.Jz0 + wz1 * et.Jz1);
    atomicAdd(reinterpret_cast<float *>(cv.Cw + t) + 1, wz0 * et.r0 + wz1 * et.r1);
    if (STRUCT_ONLY) continue;
    const bool fi = pose_free(i, cv), fj = pose_free(j, cv);
    const int li = pv.pat_li[pat0 + p], lj = pv.pat_lj[pat0 + p];
    (void)slot_pose;
    float Ej[6], Ei[6], vjf[6];
    double Bl[21], vj[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const float wa0 = et.w0 * et.Jj0[a], wa1 = et.w1 * et.Jj1[a];
      Ej[a] = wa0 * et.Jz0 + wa1 * et.Jz1;
      vjf[a] = wa0 * et.r0 + wa1 * et.r1;
      vj[a] = (double)vjf[a];
#pragma unroll
      for (int b = 0; b <= a; ++b) Bl[tri(a, b)] = (double)(wa0 * et.Jj0[b] + wa1 * et.Jj1[b]);
    }
    adjT_apply(pc.R, pc.t, Ej, Ei);
    float *col = Erows + (t - gt0);
    if (fj) for (int a = 0; a < 6; ++a) atomicAdd(col + (size_t)(6 * lj + a) * Ts, Ej[a]);
    if (fi) for (int a = 0; a < 6; ++a) atomicAdd(col + (size_t)(6 * li + a) * Ts, -Ei[a]);
    const int ri = 6 * (i - cv.fixedp), rj = 6 * (j - cv.fixedp);
    if (fj) for (int a = 0; a < 6; ++a) {
      red_add(cv.y + rj + a, vj[a]);
      for (int b = 0; b <= a; ++b) red_add(S_at(cv, rj + a, rj + b), Bl[tri(a, b)]);
    }
    if (fi) {
      double AB[6][6], vi[6];
      for (int c = 0; c < 6; ++c) {
        double col[6], out[6];
        for (int a = 0; a < 6; ++a) col[a] = a >= c ? Bl[tri(a, c)] : Bl[tri(c, a)];
        adjT_apply_d(pc.R, pc.t, col, out);
        for (int a = 0; a < 6; ++a) AB[a][c] = out[a];
      }
      adjT_apply_d(pc.R, pc.t, vj, vi);
      for (int a = 0; a < 6; ++a) {
        red_add(cv.y + ri + a, -vi[a]);
        double rowd[6];
        adjT_apply_d(pc.R, pc.t, AB[a], rowd);
        for (int b = 0; b <= a; ++b) red_add(S_at(cv, ri + a, ri + b), rowd[b]);
        if (fj) for (int b = 0; b < 6; ++b) {
          if (i > j) red_add(S_at(cv, ri + a, rj + b), -AB[a][b]);
          else if (i < j) red_add(S_at(cv, rj + b, ri + a), -AB[a][b]);
          else if (a >= b) red_add(S_at(cv, ri + a, ri + b), -(AB[a][b] + AB[b][a]));
        }
      }
    }
  }
}

// =================================================================================================
// K1b  per track: damped inverse Q and prior-adjusted w (ba.py:296-311; BA: :184)
// =================================================================================================
__global__ void k_track_q(PlanView pv, CallView cv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pv.m) return;
  const int g = pv.t_grp[t];
  if (pv.g_pat[g + 1] - pv.g_pat[g] <= kEdgeThreads || pv.g_reg[g]) return;   // the edge pass already wrote (Q, w) for this track
  const float2 cw = cv.Cw[t];
  const float lam = cv.lmbda_vec ? cv.lmbda_vec[t] : cv.lmbda;
  float C = cw.x, w = cw.y;
  if (cv.monodisp) {
    const int k = pv.kx[t];
    const float md = cv.monodisp[k];
    const float mk = md > 1e-2f ? 1.0f : 0.0f;
    C = C + mk * cv.alpha;
    C = C + lam;
    w = w - mk * cv.alpha * (cv.patches[3 * (size_t)k + 2] - md);
  } else {
    C = C + lam;
  }
  cv.Qw[t] = make_float2(1.0f / C, w);
}

// =================================================================================================
// K2  per-track Schur complement (ba.py:311-322):  S -= sum_k Q_k E_k E_k^T,  y -= sum_k Q_k w_k E_k
//   One CTA per unit (<= tu consecutive tracks of one group). thread <-> one 6x6 slot-pair block
//   (a >= b) of the group's local (6W)^2 matrix; fp32 products are summed over runs of kSchurRun
//   tracks in fp32 registers and the runs are added into fp64 registers (the subtraction B - E Q E^T
//   cancels heavily, so the long sums must not round at fp32); one flush of fp64 atomics per unit.
// =================================================================================================
constexpr int kSchurRun = 16;     // tracks per fp32 run (even)
constexpr int kSchurStages = 3;   // sub-tiles of E in flight (cp.async ring)
constexpr int kSchurPad = 2;      // floats of padding per staged row: 8-byte loads of neighbouring slots hit distinct banks

// Stage layout: E entries [rowlen][tile + kSchurPad] (entry-major like global E: the copy is 8-byte cp.async along
// the tracks, coalesced), then (Q, w) [tile] float2. A thread reads the two tracks (k, k+1) of one entry with one
// 8-byte load.
// Streaming mode (flags != nullptr): CTA k runs unit order[k] of the small-unit list and, when all its atomics are
// visible, stores `epoch` (1; the flags are cleared with S at the start of the call) into flags[k] — the band solver runs next to this kernel and reads rows of S / y as soon as
// the units that feed them are complete (SolveFeed, ba_solve_mma.cu).
__global__ void __launch_bounds__(kSchurThreads, 2) k_schur(PlanView pv, CallView cv, int tile_tracks, const int *__restrict__ ut0,
                                                            const int *__restrict__ ugrp, const int *__restrict__ order,
                                                            int *__restrict__ flags, int epoch, int tc_on) {   // tc_on: min tracks of a tensor-core unit, 0 = off
  constexpr int NT = kSchurThreads;
  extern __shared__ __align__(16) float smem[];
  const int tau = threadIdx.x;
  const int u = order ? order[blockIdx.x] : blockIdx.x;
  const int g = ugrp[u];
  const int t0 = ut0[u], t1 = ut0[u + 1];
  const int gt0 = pv.g_t0[g];
  const int W = pv.g_W[g];
  const int rowlen = 6 * W;
  const int *slot_pose = pv.slot_pose + 2 * pv.g_pat[g];
  if (tc_on) {                                     // units whose free slots fit 128 operand rows belong to k_schur_tc
    int nf = 0;
    for (int s = 0; s < W; ++s) nf += pose_free(slot_pose[s], cv) ? 1 : 0;
    if (nf <= kSchurTcMaxFree && t1 - t0 >= tc_on) return;
  }
  const float *Erows = cv.Est + pv.g_eoff[g];
  const int Ts = (pv.g_t0[g + 1] - gt0 + 3) & ~3;
  const int ts = tile_tracks + kSchurPad;
  const int stage_floats = rowlen * ts + 2 * tile_tracks;
  const int npairs = W * (W + 1) / 2;
  const int nst = (t1 - t0 + tile_tracks - 1) / tile_tracks;
  const int half = tile_tracks >> 1;

  auto issue = [&](int st) {
    if (st < nst) {
      const int tt = t0 + st * tile_tracks, nt = min(tile_tracks, t1 - tt);
      float *dst = smem + (size_t)(st % kSchurStages) * stage_floats;
      const float *src = Erows + (tt - gt0);                       // (tt - gt0) is a multiple of 4 (plan units)
      const int nh = (nt + 1) >> 1;                                // track pairs to copy (the odd tail reads row padding)
      for (int o = tau; o < rowlen * half; o += NT) {
        const int r = o / half, k2 = o - r * half;
        if (k2 < nh) cp_async8(dst + r * ts + 2 * k2, src + (size_t)r * Ts + 2 * k2);
      }
      float2 *dq = reinterpret_cast<float2 *>(dst + (size_t)rowlen * ts);
      for (int o = tau; o < tile_tracks; o += NT) {
        if (o < nt) cp_async8(dq + o, cv.Qw + tt + o);
        else dq[o] = make_float2(0.0f, 0.0f);                      // tracks past the unit contribute nothing
      }
    }
    cp_async_commit();
  };
  // E entries past the unit's last track are never copied: start from finite values (they meet Q = 0)
  for (int o = tau; o < kSchurStages * stage_floats; o += NT) smem[o] = 0.0f;

  // threads of the warps that hold no pair (first pair batch) take y -= E Q w (ba.py:322) while the others
  // multiply; without such warps every thread does it before its pairs
  const int pair_warps = (min(npairs, NT) + 31) >> 5;
  const bool y_idle = pair_warps < NT / 32;
  const int y_t0 = y_idle ? 32 * pair_warps : 0, y_nt = NT - y_t0;

  for (int pb = 0; pb < npairs; pb += NT) {
    const int x = pb + tau;
    int a = 0, b = 0;
    const bool have = x < npairs;
    if (have) {
      a = (int)((sqrtf(8.0f * (float)x + 1.0f) - 1.0f) * 0.5f);
      while (a * (a + 1) / 2 > x) --a;
      while ((a + 1) * (a + 2) / 2 <= x) ++a;
      b = x - a * (a + 1) / 2;
    }
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;

    __syncthreads();                               // ring free (previous pair batch done) / zero fill visible
    issue(0);
    issue(1);
    for (int st = 0; st < nst; ++st) {
      cp_async_wait<1>();
      __syncthreads();                             // stage st landed for everyone; stage st-1 fully consumed
      issue(st + 2);
      const int nt = min(tile_tracks, t1 - (t0 + st * tile_tracks));
      const int nt2 = (nt + 1) & ~1;
      const float *Es = smem + (size_t)(st % kSchurStages) * stage_floats;
      const float2 *qw = reinterpret_cast<const float2 *>(Es + (size_t)rowlen * ts);
      if (pb == 0 && tau >= y_t0) {                // y -= E Q w   (ba.py:322)
        for (int r = tau - y_t0; r < rowlen; r += y_nt) {
          const int pose = slot_pose[r / 6];
          if (pose_free(pose, cv)) {
            const float *er = Es + r * ts;
            double s2 = 0.0;
            for (int k0 = 0; k0 < nt; k0 += 8) {
              float ps = 0.0f;
              const int k1 = min(k0 + 8, nt);
              for (int k = k0; k < k1; ++k) ps += qw[k].x * qw[k].y * er[k];
              s2 += (double)ps;
            }
            red_add(cv.y + 6 * (pose - cv.fixedp) + (r % 6), -s2);
          }
        }
      }
      if (have) {
        const float *ea = Es + 6 * a * ts, *eb = Es + 6 * b * ts;
        for (int k0 = 0; k0 < nt2; k0 += kSchurRun) {
          float part[36];
#pragma unroll
          for (int k = 0; k < 36; ++k) part[k] = 0.0f;
          const int k1 = min(k0 + kSchurRun, nt2);
          for (int k = k0; k < k1; k += 2) {
            const float q0 = qw[k].x, q1 = qw[k + 1].x;
            float2 vb[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) vb[c] = *reinterpret_cast<const float2 *>(eb + c * ts + k);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              const float2 v = *reinterpret_cast<const float2 *>(ea + c * ts + k);
              const float vx = q0 * v.x, vy = q1 * v.y;
#pragma unroll
              for (int e2 = 0; e2 < 6; ++e2) part[c * 6 + e2] = fmaf(vy, vb[e2].y, fmaf(vx, vb[e2].x, part[c * 6 + e2]));
            }
          }
#pragma unroll
          for (int k = 0; k < 36; ++k) acc[k] += (double)part[k];
        }
      }
    }
    if (have) {                                   // S -= (E Q) E^T   (ba.py:321), lower storage only
      const int pa = slot_pose[a], pb2 = slot_pose[b];          // pa >= pb2 (slots ascend by pose)
      if (pose_free(pa, cv) && pose_free(pb2, cv)) {
        const int ra = 6 * (pa - cv.fixedp), rb = 6 * (pb2 - cv.fixedp);
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
          for (int e2 = 0; e2 < 6; ++e2)
            if (a != b || e2 <= c) red_add(S_at(cv, ra + c, rb + e2), -acc[c * 6 + e2]);
      }
    }
  }
  if (flags) {                                    // release: every thread's atomics first, then the flag
    __threadfence();
    __syncthreads();
    if (tau == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
  }
}

// =================================================================================================
// K3  reduced solve (ba.py:60-70 block_solve, :5-19 CholeskySolver, :323-325 NaN retry), in fp64.
//   A = S + (ep + lm * diag S) I;  A = L L^T;  dX = A^-1 y.  One CTA; the band window
//   [j, j+bw] x [j, j+bw] lives in shared memory as a circular buffer, columns are eliminated
//   right-looking, the forward substitution rides along, L goes to global (band) storage and the
//   backward substitution streams it back in blocks of rows. A dense system with 6n <= kMaxWindow
//   is the special case bw = 6n - 1.
// =================================================================================================
// 1/sqrt(x) in fp64 from the fp32 SFU seed + two Newton steps (rel. error ~1e-15); avoids the slow
// DSQRT + DDIV sequences on the critical path of every pivot
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y = (double)rsqrtf((float)x);
  y = y * (1.5 - 0.5 * x * y * y);
  y = y * (1.5 - 0.5 * x * y * y);
  y = y * (1.5 - 0.5 * x * y * y);
  return y;
}

__global__ void __launch_bounds__(kSolveThreads) k_solve_window(CallView cv, int allow_retry) {
  extern __shared__ double dsm[];
  constexpr int NT = kSolveThreads, NW = NT / 32;
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
  const int M = cv.M, bw = cv.bw, WS = bw + 1, WSP = WS | 1;
  double *win = dsm;                       // [WS][WSP]
  double *z = win + WS * WSP;              // [M]
  double *ls = z + M;                      // [WS]
  __shared__ int s_flag;
  const double *__restrict__ S = cv.S;
  double *__restrict__ L = cv.L;
  const int ld = cv.ld, off = cv.off;
  auto Sg = [&](int r, int c) { return (size_t)r * ld + c + off; };
  const double ep = (double)cv.ep;
  int status = 0;

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    // ---- load rows 0..bw of the window, z = y ----
    for (int r = warp; r < min(WS, M); r += NW)
      for (int c = lane; c <= r; c += 32) {
        double v = S[Sg(r, c)];
        if (c == r) v = v + (ep + lm * v);                       // ba.py:67
        win[(r % WS) * WSP + (c % WS)] = v;
      }
    for (int r = tau; r < M; r += NT) z[r] = cv.y[r];
    if (tau == 0) s_flag = 0;
    bool failed = false;
    // the row that enters the window after column j is prefetched one column ahead into registers
    constexpr int PF = (kMaxWindow + NT - 1) / NT;
    double pf[PF];
    auto prefetch_row = [&](int rn) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int x = tau + u * NT;
        pf[u] = 0.0;
        if (rn < M && x <= bw) {
          const int c = rn - bw + x;
          double v = S[Sg(rn, c)];
          if (c == rn) v = v + (ep + lm * v);
          pf[u] = v;
        }
      }
    };
    prefetch_row(WS);
    for (int j = 0; j < M; ++j) {
      __syncthreads();
      const int jm = j % WS;
      const double piv = win[jm * WSP + jm];
      if (!(piv > 0.0)) { failed = true; break; }                 // potrf info != 0 (incl. NaN), ba.py:11
      double inv, dg;
      if (piv > 1e-30 && piv < 1e30) { inv = rsqrt_fast(piv); dg = piv * inv; }
      else { dg = sqrt(piv); inv = 1.0 / dg; }
      const double zj = z[j] * inv;
      const int nb = min(bw, M - 1 - j);
      if (tau < nb) {
        const int r = j + 1 + tau;
        const double l = win[(r % WS) * WSP + jm] * inv;
        ls[tau] = l;
        L[Sg(r, j)] = l;
      }
      if (tau == 0) L[Sg(j, j)] = dg;
      __syncthreads();
      if (tau == 0) z[j] = zj;
      if (tau < nb) z[j + 1 + tau] -= ls[tau] * zj;
      // store the prefetched row j + WS (reuses the shared-memory row of the retired row j) and
      // start fetching the next one
      const int rn = j + WS;
      if (rn < M) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          const int x = tau + u * NT;
          if (x <= bw) win[jm * WSP + ((rn - bw + x) % WS)] = pf[u];
        }
      }
      prefetch_row(rn + 1);
      // rank-1 update of the trailing window, lower triangle: warp <-> rows, lanes <-> columns. Fixed
      // trip counts + full unrolling so that every load of a thread is in flight before the first FMA.
      const int j1 = (j + 1) % WS;
      constexpr int RU = (kMaxWindow + NW - 1) / NW, CU = (kMaxWindow + 31) / 32;
#pragma unroll
      for (int k = 0; k < RU; ++k) {
        const int rr = warp + k * NW;
        if (rr < nb) {
          int rs = j1 + rr; if (rs >= WS) rs -= WS;
          const double lr = ls[rr];
          double *wrow = win + rs * WSP;
          double lc[CU], wv[CU];
          int cs[CU];
#pragma unroll
          for (int u = 0; u < CU; ++u) {
            const int cc = lane + 32 * u;
            int c2 = j1 + cc; if (c2 >= WS) c2 -= WS;
            cs[u] = c2;
            const bool on = cc <= rr;
            lc[u] = on ? ls[cc] : 0.0;
            wv[u] = on ? wrow[c2] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < CU; ++u)
            if (lane + 32 * u <= rr) wrow[cs[u]] = wv[u] - lr * lc[u];
        }
      }
    }
    __syncthreads();
    if (failed) {                                                   // dX = 0 (ba.py:12-13); no NaN -> no retry
      for (int r = tau; r < M; r += NT) cv.dX[r] = 0.0;
      status |= (attempt == 0) ? 1 : 4;
      break;
    }
    // ---- backward substitution L^T x = z, rows of L streamed back in blocks of WS rows ----
    for (int jb = M - 1; jb >= 0; jb -= WS) {
      const int lo = max(jb - WS + 1, 0);
      __syncthreads();
      for (int r = lo + warp; r <= jb; r += NW)
        for (int x = lane; x <= bw; x += 32) {
          const int c = r - bw + x;
          win[(r - lo) * WSP + x] = c >= 0 ? L[Sg(r, c)] : 0.0;
        }
      __syncthreads();
      if (warp == 0) {
        for (int j = jb; j >= lo; --j) {
          const double *row = win + (j - lo) * WSP;                 // row[x] = L(j, j - bw + x)
          const double xj = z[j] / row[bw];
          __syncwarp();
          if (lane == 0) z[j] = xj;
          for (int x = lane; x < bw; x += 32) {
            const int c = j - bw + x;
            if (c >= 0) z[c] -= row[x] * xj;
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
    int nan_local = 0;
    for (int r = tau; r < M; r += NT) { const double v = z[r]; cv.dX[r] = v; nan_local |= (v != v); }
    if (nan_local) s_flag = 1;
    __syncthreads();
    if (s_flag && allow_retry && attempt == 0) { status |= 2; __syncthreads(); continue; }   // ba.py:324-325
    break;
  }
  if (tau == 0) cv.status[0] = status;
}

// K3b  dense fallback for reduced systems whose band does not fit the shared-memory window
//   (6n > kMaxWindow with no usable band structure: unstructured graphs, loop closures). Same
//   algorithm on the dense lower matrix held in global memory (L2-resident), one CTA, fp64.
//   Correct for any size; a multi-CTA blocked version is future work.
__global__ void __launch_bounds__(kSolveThreads) k_solve_dense(CallView cv, int allow_retry) {
  extern __shared__ double dsm[];
  constexpr int NT = kSolveThreads;
  const int tau = threadIdx.x;
  const int M = cv.M;
  double *z = dsm;            // [M]
  double *ls = dsm + M;       // [M]
  __shared__ int s_flag;
  const double *S = cv.S;
  double *L = cv.L;
  const double ep = (double)cv.ep;
  int status = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    for (size_t idx = tau; idx < (size_t)M * M; idx += NT) {
      const int r = (int)(idx / M), c = (int)(idx % M);
      if (c > r) continue;
      double v = S[idx];
      if (c == r) v = v + (ep + lm * v);
      L[idx] = v;
    }
    for (int r = tau; r < M; r += NT) z[r] = cv.y[r];
    if (tau == 0) s_flag = 0;
    bool failed = false;
    for (int j = 0; j < M; ++j) {
      __syncthreads();
      const double piv = L[(size_t)j * M + j];
      if (!(piv > 0.0)) { failed = true; break; }
      const double dg = sqrt(piv), inv = 1.0 / dg;
      const double zj = z[j] * inv;
      for (int r = j + 1 + tau; r < M; r += NT) ls[r] = L[(size_t)r * M + j] * inv;
      __syncthreads();
      if (tau == 0) { L[(size_t)j * M + j] = dg; z[j] = zj; }
      for (int r = j + 1 + tau; r < M; r += NT) { L[(size_t)r * M + j] = ls[r]; z[r] -= ls[r] * zj; }
      const int nb = M - 1 - j;
      // trailing update, lower triangle: one warp per row, lanes over columns
      for (int rr = tau >> 5; rr < nb; rr += NT / 32) {
        const int r = j + 1 + rr;
        const double lr = ls[r];
        double *row = L + (size_t)r * M;
        for (int c = j + 1 + (tau & 31); c <= r; c += 32) row[c] -= lr * ls[c];
      }
    }
    __syncthreads();
    if (failed) {
      for (int r = tau; r < M; r += NT) cv.dX[r] = 0.0;
      status |= (attempt == 0) ? 1 : 4;
      break;
    }
    for (int j = M - 1; j >= 0; --j) {
      __syncthreads();
      const double xj = z[j] / L[(size_t)j * M + j];
      __syncthreads();
      if (tau == 0) z[j] = xj;
      const double *row = L + (size_t)j * M;
      for (int c = tau; c < j; c += NT) z[c] -= row[c] * xj;
    }
    __syncthreads();
    int nan_local = 0;
    for (int r = tau; r < M; r += NT) { const double v = z[r]; cv.dX[r] = v; nan_local |= (v != v); }
    if (nan_local) s_flag = 1;
    __syncthreads();
    if (s_flag && allow_retry && attempt == 0) { status |= 2; __syncthreads(); continue; }
    break;
  }
  if (tau == 0) cv.status[0] = status;
}

// =================================================================================================
// K4  back-substitution dZ = Q (w - E^T dX) (ba.py:328 / :317), disparity retraction + clamp of EVERY patch
//     (ba.py:42-44,332-334) and the fresh patches tensor, in one pass: one thread per patch. E is entry-major, so
//     consecutive threads (consecutive tracks of a group) read consecutive floats; the pose of a slot and its dX
//     are uniform across the threads of a group.
// =================================================================================================
// torch.clamp propagates NaN (ba.py:333); fminf / fmaxf would return the bound instead
__device__ __forceinline__ float clamp_disp(float v) { return v != v ? v : fminf(fmaxf(v, 1e-3f), 10.0f); }

__device__ __forceinline__ void pose_retr_one(const CallView &cv, int i) {
  float a[6] = {0, 0, 0, 0, 0, 0};
  if (pose_free(i, cv)) {
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = (float)cv.dX[6 * (i - cv.fixedp) + c];
  }
  Pose dXp = pose_exp(a);
  float tmp[7];
  pose_store(dXp, tmp);
  Pose r = pose_mul(pose_load(tmp), pose_load(cv.poses + 7 * (size_t)i));   // mul re-loads both operands
  pose_store(r, cv.poses_out + 7 * (size_t)i);
}

// Three kinds of blocks in one launch:
//   [0, nb_trk)            32 consecutive tracks per CTA, lane <-> track (coalesced entry-major E reads), the 8 warps
//                          split the slots of the track's group; partial dots meet in shared memory in fixed order;
//                          warp 0 writes dZ and the three floats of the track's patch;
//   [nb_trk, nb_trk+nb_cp) one thread per patch: patches without a track are copied with the clamp;
//   beyond                 pose retraction T <- Exp(dx) T for every pose of the buffer, dx = 0 outside the window
//                          (ba.py:47-49,336-337; lietorch/groups.py:153-156), when retr_n > 0
__global__ void __launch_bounds__(256) k_backsub(PlanView pv, CallView cv, int use_dx, int nb_trk, int nb_cp, int nb_retr, int retr_n,
                                                 double2 *__restrict__ zero_ptr, long long zero_n16, int nb_zero) {
  __shared__ double part[8][32];
  const int bid = blockIdx.x;
  if (bid >= nb_trk + nb_cp + nb_retr) {                           // clear the next call's reduced system (the other buffer)
    const double2 z2 = make_double2(0.0, 0.0);
    for (long long i = (long long)(bid - nb_trk - nb_cp - nb_retr) * blockDim.x + threadIdx.x; i < zero_n16; i += (long long)nb_zero * blockDim.x)
      zero_ptr[i] = z2;
    return;
  }
  if (bid >= nb_trk + nb_cp) {
    const int i = (bid - nb_trk - nb_cp) * blockDim.x + threadIdx.x;
    if (i < retr_n) pose_retr_one(cv, i);
    return;
  }
  if (bid >= nb_trk) {
    const int k = (bid - nb_trk) * blockDim.x + threadIdx.x;
    if (k >= pv.NM || pv.patch_track[k] >= 0) return;
    cv.patches_out[3 * (size_t)k] = cv.patches[3 * (size_t)k];
    cv.patches_out[3 * (size_t)k + 1] = cv.patches[3 * (size_t)k + 1];
    cv.patches_out[3 * (size_t)k + 2] = clamp_disp(cv.patches[3 * (size_t)k + 2]);   // clamp hits every patch, ba.py:333
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = bid * 32 + lane;
  const bool have = t < pv.m;
  if (use_dx) {
    double dot = 0.0;
    if (have) {
      const int g = pv.t_grp[t];
      const int W = pv.g_W[g], gt0 = pv.g_t0[g];
      const int Ts = (pv.g_t0[g + 1] - gt0 + 3) & ~3;
      const int *slot_pose = pv.slot_pose + 2 * pv.g_pat[g];
      const float *col = cv.Est + pv.g_eoff[g] + (t - gt0);
      for (int sl = warp; sl < W; sl += 8) {
        const int pose = slot_pose[sl];
        if (!pose_free(pose, cv)) continue;
        const double *dx = cv.dX + 6 * (pose - cv.fixedp);
        const float *e = col + (size_t)(6 * sl) * Ts;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c += 2) { d0 += (double)e[(size_t)c * Ts] * dx[c]; d1 += (double)e[(size_t)(c + 1) * Ts] * dx[c + 1]; }
        dot += d0 + d1;
      }
    }
    part[warp][lane] = dot;
    __syncthreads();
  }
  if (warp == 0 && have) {
    const float2 qw = cv.Qw[t];
    double dot = 0.0;
    if (use_dx) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dot += part[k][lane];
    }
    const float dz = (float)((double)qw.x * ((double)qw.y - dot));
    cv.dZ[t] = dz;
    const size_t k = (size_t)pv.kx[t];
    cv.patches_out[3 * k] = cv.patches[3 * k];
    cv.patches_out[3 * k + 1] = cv.patches[3 * k + 1];
    cv.patches_out[3 * k + 2] = clamp_disp(cv.patches[3 * k + 2] + dz);
  }
}

// ---- debug: expand the lower (band) storage to a dense symmetric matrix, cast fp64 -> fp32 --------
__global__ void k_debug_dense(const double *S, int M, int ld, int off, int bw, float *out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * M) return;
  int r = idx / M, c = idx % M;
  if (c > r) { int t = r; r = c; c = t; }
  out[idx] = (r - c <= bw) ? (float)S[(size_t)r * ld + c + off] : 0.0f;
}
__global__ void k_debug_cast(const double *in, float *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

static int make_call(BaPlan *pl, const BaProblem *pb, CallView *cv) {
  if (!pl || !pb || !pb->poses || !pb->patches || !pb->intrinsics || !pb->targets || !pb->weights) return BA_ERR_ARG;
  if (pb->fixedp < 0 || pb->loss < 0 || pb->loss > 2) return BA_ERR_ARG;
  if (int rc = plan_finalize(pl)) return rc;                       // a pending ba_plan_update: wait for its shape block
  std::memset(cv, 0, sizeof(*cv));
  cv->poses = pb->poses; cv->patches = pb->patches; cv->monodisp = pb->monodisp; cv->intr = pb->intrinsics;
  cv->targets = pb->targets; cv->weights = pb->weights; cv->lmbda_vec = pb->lmbda_vec;
  cv->lmbda = pb->lmbda; cv->ep = pb->ep; cv->alpha = pb->alpha;
  for (int k = 0; k < 4; ++k) cv->bounds[k] = pb->bounds[k];
  cv->fixedp = pb->fixedp; cv->loss = pb->loss; cv->structure_only = pb->structure_only;
  cv->tstride = pb->targets_stride == 0 ? 2 : pb->targets_stride;
  if (cv->tstride < 2) return BA_ERR_ARG;
  int n, bw, ld, off; int64_t sf;
  layout_for(pl, pb->fixedp, &n, &bw, &ld, &off, &sf);
  cv->n = n; cv->M = 6 * n; cv->ld = ld; cv->off = off; cv->bw = bw;
  double *sy = (pl->sy_cur && pl->SY2) ? pl->SY2 : pl->SY;        // the buffer of this call (toggled when a call has solved)
  cv->S = sy; cv->y = sy + sf;
  cv->Est = pl->Est; cv->Cw = pl->Cw; cv->Qw = pl->Qw; cv->dX = pl->dX; cv->dZ = pl->dZ; cv->L = pl->L;
  cv->status = pl->status;
  cv->poses_out = pb->poses_out; cv->patches_out = pb->patches_out;
  return BA_OK;
}

}  // namespace ba

using namespace ba;

// Can the reduced system of this call go to the DMMA band solver (and therefore be streamed to it)?
// The shared-memory tile solver takes the short systems (every window BA-Track really builds) and the bands the register
// window of the band solver does not cover (half bandwidth 121 .. 145: the full-sequence Sintel window). Measured
// (tools/solver_ab.py): 90 unknowns dense 28 us vs 39, 288 unknowns / bw 131 114 us vs 637 (scalar window solver);
// from ~33 tile columns on a band <= 120 the band solver wins (378 unknowns: 113 us vs 134).
constexpr int kTileSolverMaxCols = 32;
static bool band_solver_covers(const CallView &cv) {
  return cv.ld != cv.M && cv.bw <= kMmaMaxBw && std::max(solve_mma_smem_bytes(cv.M), solve_diag_smem_bytes(cv.M)) <= 227 * 1024 - 1024;
}
static bool tile_solver_applies(const BaPlan *pl, const CallView &cv) {
  if (pl->opt.solver != 0 && pl->opt.solver != 4) return false;
  if (!solve_tiles_applies(cv.M, cv.bw, nullptr)) return false;
  return pl->opt.solver == 4 || (cv.M + 7) / 8 <= kTileSolverMaxCols || !band_solver_covers(cv);
}
static bool mma_solver_applies(const BaPlan *pl, const CallView &cv) {
  const bool want_mma = pl->opt.solver == 0 || pl->opt.solver == 1 || pl->opt.solver == 5;   // BA_OPT_SOLVER 2 / 3 / 4 force another one (tests, A/B timing)
  if (tile_solver_applies(pl, cv)) return false;
  return want_mma && band_solver_covers(cv);
}
static int launch_band_solver(const BaPlan *pl, const CallView &cv, int allow_retry, const SolveFeed &feed, cudaStream_t s) {
  return pl->opt.solver == 1 ? launch_solve_band_mma(cv, allow_retry, pl->Wg, feed, s) : launch_solve_band_diag(cv, allow_retry, pl->Wg, feed, s);
}

// Function attributes are per device: set for every kernel of this translation unit when a plan is created on a
// device for the first time (ba_plan.cu), never on the hot path.
namespace ba {
int kernels_prepare_device() {
  BA_CUDA(cudaFuncSetAttribute(k_edge_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEdgeSmemBytes));
  BA_CUDA(cudaFuncSetAttribute(k_edge_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEdgeSmemBytes));
  size_t e2max[3] = {0, 0, 0};
  for (int tpl = 1; tpl <= 2; ++tpl)
    for (int kp = 1; kp <= kEdge2Warps; kp *= 2) e2max[tpl] = std::max(e2max[tpl], edge2_smem_bytes(kp, tpl));
  BA_CUDA(cudaFuncSetAttribute(k_edge_pass_v2<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e2max[1]));
  BA_CUDA(cudaFuncSetAttribute(k_edge_pass_v2<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e2max[1]));
  BA_CUDA(cudaFuncSetAttribute(k_edge_pass_v2<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e2max[2]));
  BA_CUDA(cudaFuncSetAttribute(k_edge_pass_v2<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e2max[2]));
  BA_CUDA(cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  BA_CUDA(cudaFuncSetAttribute(k_solve_window, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
  BA_CUDA(cudaFuncSetAttribute(k_solve_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
  return BA_OK;
}
}  // namespace ba

// streaming != 0 (single-device ba_step, band solver): after the edge pass the solver is launched on the plan's own
// stream, then the Schur kernel runs its small units in "both ends first" order and publishes completion flags;
// the solver eliminates columns while the Schur kernel is still working on the middle of the pose range.
static int assemble_impl(BaPlan *pl, const BaProblem *pb, void *stream_, int streaming) {
  CallView cv;
  int rc = make_call(pl, pb, &cv);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream_;
  const PlanView &pv = pl->v;
  const bool so = pb->structure_only || cv.n == 0;                 // ba.py:316
  pl->last_n = cv.n; pl->last_fixedp = pb->fixedp;
  pl->ev_mask = 0;
  // regular groups (every SLAM graph): lane-per-track kernel; irregular groups: the generic kernels (each kernel
  // returns at once on the groups of the other kind)
  const bool any_regular = pv.n_irregular < pv.G, any_irregular = pv.n_irregular > 0;
  const bool has_long = any_irregular && pv.dmax_irregular > kEdgeThreads;
  if (has_long) {   // the slow path accumulates with atomics: its targets start from zero
    BA_CUDA(cudaMemsetAsync(cv.Cw, 0, (size_t)pv.m * sizeof(float2), s));
    if (!so) BA_CUDA(cudaMemsetAsync(cv.Est, 0, (size_t)pl->est_floats * sizeof(float), s));
  }
  if (so) {
    BA_MARK(pl, BA_STAGE_EDGE, s);
    if (any_regular) {
      if (pv.e2_tpl == 2) k_edge_pass_v2<true, 2><<<pv.n_xchunks, kE2Threads, edge2_smem_bytes(pv.e2_kp, 2), s>>>(pv, cv);
      else k_edge_pass_v2<true, 1><<<pv.n_xchunks, kE2Threads, edge2_smem_bytes(pv.e2_kp, 1), s>>>(pv, cv);
      BA_LAUNCH_CHECK();
    }
    if (any_irregular) { k_edge_pass<true><<<pv.n_chunks, kEdgeThreads, kEdgeSmemBytes, s>>>(pv, cv); BA_LAUNCH_CHECK(); }
    if (has_long) { k_edge_pass_long<true><<<pv.n_chunks, kEdgeThreads, 0, s>>>(pv, cv); BA_LAUNCH_CHECK(); }
  } else {
    BA_MARK(pl, BA_STAGE_ZERO, s);
    // streaming: the completion flags of the Schur units sit right behind y and are cleared by the same memset (no
    // per-call state on the host: a captured CUDA graph of this call can be replayed)
    {
      // skipped when the previous call's back-substitution kernel has already cleared this buffer (solve_update_impl)
      const size_t zbytes = (size_t)((cv.y - cv.S) + cv.M) * sizeof(double) + (streaming ? (size_t)pv.n_ounits * sizeof(int) : 0);
      const int b = (pl->sy_cur && pl->SY2) ? 1 : 0;
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      BA_CUDA(cudaStreamIsCapturing(s, &cs));
      if (cs != cudaStreamCaptureStatusNone) pl->sy_untracked = 1;
      if (pl->sy_untracked || (size_t)pl->sy_clean[b] < zbytes) BA_CUDA(cudaMemsetAsync(cv.S, 0, zbytes, s));
      pl->sy_clean[b] = 0;
      pl->sy_last = cv.S;
    }
    if (streaming) {
      BA_CUDA(cudaEventRecord(pl->ev_step_begin, s));               // flags cleared; the previous call's back-substitution (reads dX) is behind
      BA_CUDA(cudaStreamWaitEvent(pl->solve_stream, pl->ev_step_begin, 0));
    }
    BA_MARK(pl, BA_STAGE_EDGE, s);
    if (any_regular) {
      if (pv.e2_tpl == 2) k_edge_pass_v2<false, 2><<<pv.n_xchunks, kE2Threads, edge2_smem_bytes(pv.e2_kp, 2), s>>>(pv, cv);
      else k_edge_pass_v2<false, 1><<<pv.n_xchunks, kE2Threads, edge2_smem_bytes(pv.e2_kp, 1), s>>>(pv, cv);
      BA_LAUNCH_CHECK();
    }
    if (any_irregular) { k_edge_pass<false><<<pv.n_chunks, kEdgeThreads, kEdgeSmemBytes, s>>>(pv, cv); BA_LAUNCH_CHECK(); }
    if (has_long) { k_edge_pass_long<false><<<pv.n_chunks, kEdgeThreads, 0, s>>>(pv, cv); BA_LAUNCH_CHECK(); }
  }
  BA_MARK(pl, BA_STAGE_TRACKQ, s);
  if (has_long) {                    // the atomics path leaves raw (C, w) sums; everything else writes (Q, w) directly
    k_track_q<<<(pv.m + 255) / 256, 256, 0, s>>>(pv, cv); BA_LAUNCH_CHECK();
  }
  BA_MARK(pl, BA_STAGE_SCHUR, s);
  if (!so) {
    const int rowmax = 6 * pl->info.max_slots;
    int tile = std::max(4, pl->opt.schur_tile & ~3);                     // per stage; kSchurStages stages in flight; 2 CTAs per SM
    while (tile > 4 && (size_t)kSchurStages * (rowmax * (tile + kSchurPad) + 2 * tile) * sizeof(float) > 100 * 1024) tile -= 4;
    const size_t smem = (size_t)kSchurStages * (rowmax * (tile + kSchurPad) + 2 * tile) * sizeof(float);
    if (smem > 200 * 1024) return BA_ERR_ARG;
    // tensor-core kernel (tcgen05, 3xTF32) for every unit whose free pose slots fit 128 operand rows; the SIMT kernel
    // takes the rest (and everything with BA_OPT_SCHUR = 1)
    // (it exits at once on the units of the other kernel). Either launch is skipped when the plan knows it has no work.
    const int umin = streaming ? pl->min_ounit : pl->min_unit, umax = streaming ? pl->max_ounit : pl->max_unit;
    const bool tc = pl->opt.schur == 0 && umax >= kSchurTcMinTracks;
    const bool simt = !tc || umin < kSchurTcMinTracks || pl->info.max_slots > kSchurTcMaxFree;
    if (streaming) {
      int *flags = reinterpret_cast<int *>(cv.y + cv.M);
      const SolveFeed feed = make_feed(pl, flags, pv.top_need, pv.bot_need, 1, pv.n_ounits, pb->fixedp, 1, pl->status + 2);
      BA_CUDA(cudaMemsetAsync(pl->status + 2, 0, sizeof(int), pl->solve_stream));
      rc = launch_band_solver(pl, cv, pb->monodisp ? 1 : 0, feed, pl->solve_stream);
      if (rc) return rc;
      BA_CUDA(cudaEventRecord(pl->ev_solved, pl->solve_stream));
      const size_t stream_smem = (size_t)pl->opt.stream_smem_kb * 1024;   // BA_OPT_STREAM_SMEM_KB: throttle occupancy (tests force the give-up path with it)
      if (tc) { rc = launch_schur_tc(pv, cv, pv.n_ounits, pv.o_t0, pv.o_grp, pv.o_order, flags, 1, pl->opt.schur_acc, kSchurTcMinTracks, (pl->opt.trace & 2) ? pl->trace_buf : nullptr, s); if (rc) return rc; }
      if (simt) {
        k_schur<<<pv.n_ounits, kSchurThreads, std::max(smem, stream_smem), s>>>(pv, cv, tile, pv.o_t0, pv.o_grp, pv.o_order, flags, 1, tc ? kSchurTcMinTracks : 0);
        BA_LAUNCH_CHECK();
      }
    } else {
      if (tc) { rc = launch_schur_tc(pv, cv, pv.n_units, pv.u_t0, pv.u_grp, nullptr, nullptr, 0, pl->opt.schur_acc, kSchurTcMinTracks, (pl->opt.trace & 2) ? pl->trace_buf : nullptr, s); if (rc) return rc; }
      if (simt) { k_schur<<<pv.n_units, kSchurThreads, smem, s>>>(pv, cv, tile, pv.u_t0, pv.u_grp, nullptr, nullptr, 0, tc ? kSchurTcMinTracks : 0); BA_LAUNCH_CHECK(); }
    }
  }
  BA_MARK(pl, BA_STAGE_SOLVE, s);      // closes SCHUR; a sharded caller's all-reduce lands in SOLVE's interval
  return BA_OK;
}

extern "C" int ba_assemble(BaPlan *pl, const BaProblem *pb, void *stream_) { return assemble_impl(pl, pb, stream_, 0); }

extern "C" int ba_plan_reduced_system(const BaPlan *pl, double **ptr, int64_t *n_values) {
  int64_t *n_floats = n_values;
  if (!pl || !ptr || !n_floats || pl->last_fixedp < 0) return BA_ERR_ARG;
  int n, bw, ld, off; int64_t sf;
  layout_for(pl, pl->last_fixedp, &n, &bw, &ld, &off, &sf);
  *ptr = (pl->sy_cur && pl->SY2) ? pl->SY2 : pl->SY;             // the buffer ba_assemble has just filled
  *n_floats = sf + 6 * (int64_t)n;
  return BA_OK;
}

// solved != 0: the solve was launched by the streaming assemble; wait for it on the caller's stream
static int solve_update_impl(BaPlan *pl, const BaProblem *pb, void *stream_, int solved) {
  CallView cv;
  int rc = make_call(pl, pb, &cv);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream_;
  const PlanView &pv = pl->v;
  const bool so = pb->structure_only || cv.n == 0;
  // a structure-only call leaves the poses alone (ba.py:336-339 returns the caller's object): poses_out may be NULL then
  if ((!pb->poses_out && !so) || !pb->patches_out) return BA_ERR_ARG;
  if (!so && solved) {
    // join the streaming solve; the stand-by launch behind it does the solve only if that one gave up (kernels
    // serialised by a profiler / sanitizer: its producer never ran next to it)
    BA_CUDA(cudaStreamWaitEvent(s, pl->ev_solved, 0));
    rc = launch_band_solver(pl, cv, pb->monodisp ? 1 : 0, make_feed(pl, nullptr, nullptr, nullptr, 0, 0, 0, 2, pl->status + 2), s);
    if (rc) return rc;
  } else if (!so) {
    if (!(pl->ev_mask & (1u << BA_STAGE_SOLVE))) BA_MARK(pl, BA_STAGE_SOLVE, s);
    const int WS = cv.bw + 1, WSP = WS | 1;
    const size_t smem = ((size_t)WS * WSP + cv.M + WS) * sizeof(double);
    if (tile_solver_applies(pl, cv)) {
      rc = launch_solve_tiles(cv, pb->monodisp ? 1 : 0, pl->Wg, (pl->opt.trace & 4) ? pl->trace_buf : nullptr, s);
      if (rc) return rc;
    } else if (mma_solver_applies(pl, cv)) {
      rc = launch_band_solver(pl, cv, pb->monodisp ? 1 : 0, make_feed(pl, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr), s);
      if (rc) return rc;
    } else if (cv.ld != cv.M && smem <= 227 * 1024 - 64 && pl->opt.solver != 3) {
      k_solve_window<<<1, kSolveThreads, smem, s>>>(cv, pb->monodisp ? 1 : 0); BA_LAUNCH_CHECK();
    } else {
      const size_t smem_d = 2 * (size_t)cv.M * sizeof(double);
      if (cv.ld != cv.M || smem_d > 227 * 1024 - 64) return BA_ERR_ARG;    // > 14k unknowns without band structure
      k_solve_dense<<<1, kSolveThreads, smem_d, s>>>(cv, pb->monodisp ? 1 : 0); BA_LAUNCH_CHECK();
    }
  }
  BA_MARK(pl, BA_STAGE_BACKSUB, s);
  {
    const int nb_trk = (pv.m + 31) / 32, nb_cp = (pv.NM + 255) / 256, nb_retr = so ? 0 : (pv.N + 255) / 256;
    // a call that solved hands the next one a cleared reduced system: the other buffer, zeroed by spare blocks here
    double *zero_ptr = nullptr;
    long long zero_n16 = 0;
    if (!so && pl->SY2 && !pl->sy_untracked) {
      const int other = pl->sy_cur ? 0 : 1;
      const size_t zbytes = ((size_t)((cv.y - cv.S) + cv.M) * sizeof(double) + (size_t)pv.n_ounits * sizeof(int) + 15) & ~(size_t)15;
      zero_ptr = other ? pl->SY2 : pl->SY;
      zero_n16 = (long long)(zbytes / 16);
      pl->sy_clean[other] = (int64_t)zbytes;
      pl->sy_cur = other;
    }
    const int nb_zero = zero_ptr ? (int)std::min<long long>((zero_n16 + 2047) / 2048, 148) : 0;
    k_backsub<<<nb_trk + nb_cp + nb_retr + nb_zero, 256, 0, s>>>(pv, cv, so ? 0 : 1, nb_trk, nb_cp, nb_retr, so ? 0 : pv.N,
                                                                  reinterpret_cast<double2 *>(zero_ptr), zero_n16, nb_zero);
    BA_LAUNCH_CHECK();
  }
  BA_MARK(pl, BA_STAGE_RETR, s);
  if (so && pb->poses_out && pb->poses_out != pb->poses)
    BA_CUDA(cudaMemcpyAsync(pb->poses_out, pb->poses, (size_t)pv.N * 7 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  BA_MARK(pl, BA_N_STAGES, s);
  return BA_OK;
}

extern "C" int ba_solve_update(BaPlan *pl, const BaProblem *pb, void *stream_) { return solve_update_impl(pl, pb, stream_, 0); }

extern "C" int ba_step(BaPlan *pl, const BaProblem *pb, void *stream) {
  // Streaming hand-over Schur -> solve when the band solver applies (BA_STREAM=0 switches it off)
  const int stream_on = pl ? pl->opt.stream : 0;
  CallView cv;
  int rc = make_call(pl, pb, &cv);
  if (rc) return rc;
  const bool so = pb->structure_only || cv.n == 0;
  if (stream_on && !so && pl->v.n_ounits > 0 && mma_solver_applies(pl, cv) && pb->poses_out && pb->patches_out) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!pl->solve_stream) {
      int lo = 0, hi = 0;
      BA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      BA_CUDA(cudaStreamCreateWithPriority(&pl->solve_stream, cudaStreamNonBlocking, hi));
      BA_CUDA(cudaEventCreateWithFlags(&pl->ev_step_begin, cudaEventDisableTiming));
      BA_CUDA(cudaEventCreateWithFlags(&pl->ev_solved, cudaEventDisableTiming));
    }
    rc = assemble_impl(pl, pb, stream, 1);
    if (rc) return rc;
    return solve_update_impl(pl, pb, stream, 1);
  }
  rc = ba_assemble(pl, pb, stream);
  if (rc) return rc;
  return ba_solve_update(pl, pb, stream);
}

extern "C" int ba_update(BaPlan *pl, const BaProblem *pb, const float *weights_all, int32_t iters, void *stream) {
  if (!pl || !pb || !weights_all || iters <= 0 || !pb->poses_out || !pb->patches_out) return BA_ERR_ARG;
  const size_t np = (size_t)pl->v.N * 7, nq = (size_t)pl->v.NM * 3;
  if (!pl->pp_buf[0]) {
    for (int k = 0; k < 2; ++k) {
      void *q = nullptr;
      BA_CUDA(cudaMallocAsync(&q, (np + nq + 8) * sizeof(float), (cudaStream_t)stream));
      pl->owned.push_back(q);
      pl->pp_buf[k] = (float *)q;
    }
  }
  BaProblem cur = *pb;
  const int total = 2 * iters;
  for (int c = 0; c < total; ++c) {
    const bool so = (c & 1) != 0;                       // main/batrack.py:871-872 then :874-875
    const bool last_pose_call = c == total - 2, last = c == total - 1;
    // the structure-only calls leave the poses where they are (no copy); the last pose call writes the caller's buffer
    float *po = so ? nullptr : (last_pose_call ? pb->poses_out : pl->pp_buf[(c >> 1) & 1]);
    float *qo = last ? pb->patches_out : pl->pp_buf[c & 1] + np;
    cur.poses_out = po;
    cur.patches_out = qo;
    cur.structure_only = so ? 1 : 0;
    cur.weights = so ? weights_all : pb->weights;
    int rc = ba_step(pl, &cur, stream);
    if (rc) return rc;
    if (!so) cur.poses = po;
    cur.patches = qo;
  }
  return BA_OK;
}

extern "C" int ba_plan_debug_dense(const BaPlan *pl, int32_t n, float *S, float *y, float *dX, float *Q,
                                   float *w, float *dZ, void *stream_) {
  if (!pl || pl->last_fixedp < 0) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  int nn, bw, ld, off; int64_t sf;
  layout_for(pl, pl->last_fixedp, &nn, &bw, &ld, &off, &sf);
  if (n != nn) return BA_ERR_ARG;
  const int M = 6 * nn;
  const double *sy = pl->sy_last ? pl->sy_last : pl->SY;           // [S | y] of the last call that assembled one
  if (S && M > 0) { k_debug_dense<<<(M * M + 255) / 256, 256, 0, s>>>(sy, M, ld, off, bw, S); BA_LAUNCH_CHECK(); }
  if (y && M > 0) { k_debug_cast<<<(M + 255) / 256, 256, 0, s>>>(sy + sf, y, M); BA_LAUNCH_CHECK(); }
  if (dX && M > 0) { k_debug_cast<<<(M + 255) / 256, 256, 0, s>>>(pl->dX, dX, M); BA_LAUNCH_CHECK(); }
  const int m = pl->v.m;
  if (Q) BA_CUDA(cudaMemcpy2DAsync(Q, sizeof(float), pl->Qw, sizeof(float2), sizeof(float), m, cudaMemcpyDeviceToDevice, s));
  if (w) BA_CUDA(cudaMemcpy2DAsync(w, sizeof(float), reinterpret_cast<float *>(pl->Qw) + 1, sizeof(float2), sizeof(float), m, cudaMemcpyDeviceToDevice, s));
  if (dZ) BA_CUDA(cudaMemcpyAsync(dZ, pl->dZ, m * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return BA_OK;
}

extern "C" int ba_plan_status_ptr(const BaPlan *pl, int32_t **dev_status) {
  if (!pl || !dev_status) return BA_ERR_ARG;
  *dev_status = pl->status;
  return BA_OK;
}
