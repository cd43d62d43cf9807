// ba_kernels.cu — the BA iteration: edge pass, per-track Schur complement, reduced solve,
// back-substitution and retractions (reference: main/backend/ba.py:217-339 and the projective_ops /
// lietorch code it calls). See DESIGN.md for the data layout and the roofline of each kernel.
#include <cstring>

#include "ba_internal.h"
#include "ba_math.cuh"

namespace ba {

// Reduced-system accumulators are fp64: the per-edge math is fp32 like the reference's, but every sum
// that feeds the (ill-conditioned) reduced solve is carried in double — see DESIGN.md §Precision.
__device__ __forceinline__ void red_add(double *addr, double v) { atomicAdd(addr, v); }

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
// K1  edge pass.  One CTA per chunk (<= tc consecutive tracks of one pattern group).
//   thread <-> (track slot kappa, pattern position p); the (i,j) pair of a position is fixed, so
//   Gij / adjoint / intrinsics are registers, and Bjj, vj are accumulated in registers over the
//   chunk's tracks. Ji = -Ad(Gij)^T Jj (projective_ops.py:96) makes every i-side quantity a fixed
//   linear image of the j-side one:  Bii = A Bjj A^T, Bij = -A Bjj, vi = -A vj, Eik = -A Ejk,
//   with A = Ad(Gij)^T, applied once per position (B, v) or per edge (E).
//   Per-edge E 6-vectors and the C, w scalars go through shared memory and are reduced per track in
//   a fixed order into the group's dense E rows [track][slot*6+c]  (never a dense [n, m] E).
// =================================================================================================
// staging components per edge: Eik[6] Ejk[6] c w (14); accumulators per thread: Bjj lower[21] vj[6]
constexpr int kAccComps = 27;     // Bjj lower[21] vj[6]

template <bool STRUCT_ONLY>
__global__ void __launch_bounds__(kEdgeThreads, 2) k_edge_pass(PlanView pv, CallView cv) {
  constexpr int NT = kEdgeThreads;
  __shared__ float sh[kAccComps * NT];      // staging (14*NT) during passes, accumulators (27*NT) at flush
  const int tau = threadIdx.x;
  const int chunk = blockIdx.x;
  const int g = pv.c_grp[chunk];
  const int t0 = pv.c_t0[chunk], t1 = pv.c_t0[chunk + 1];
  const int gt0 = pv.g_t0[g];
  const int pat0 = pv.g_pat[g];
  const int d = pv.g_pat[g + 1] - pat0;
  const int W = pv.g_W[g];
  const int ebase = pv.tptr[gt0];
  const int sbase = 2 * pat0;
  const int *slot_pose = pv.slot_pose + sbase;
  const int *slot_ptr = pv.slot_ptr + sbase + g;
  const int *slot_items = pv.slot_items + sbase;
  float *Erows = cv.Est + pv.g_eoff[g];
  const int rowlen = 6 * W;
  const int outs_per_track = STRUCT_ONLY ? 2 : rowlen + 2;

  for (int p0 = 0; p0 < d; p0 += NT) {
    const int dc = min(d - p0, NT);
    const int Tp = NT / dc;
    const bool active = tau < Tp * dc;
    const int kappa = tau / dc;
    const int p = p0 + (tau - kappa * dc);

    PairConst pc;
    if (active) {
      const int i = pv.pat_i[pat0 + p], j = pv.pat_j[pat0 + p];
      pc = pair_const(cv.poses + 7 * i, cv.poses + 7 * j, cv.intr + 4 * i, cv.intr + 4 * j);
    }
    float acc[kAccComps];
#pragma unroll
    for (int k = 0; k < kAccComps; ++k) acc[k] = 0.0f;

    for (int ts = t0; ts < t1; ts += Tp) {
      const int t = ts + kappa;
      if (active && t < t1) {
        const int q = ebase + (t - gt0) * d + p;
        const int e = pv.perm_identity ? q : __ldg(pv.eperm + q);
        float2 tg;
        if (cv.tstride == 2) tg = __ldg(reinterpret_cast<const float2 *>(cv.targets) + e);
        else { const float *tp = cv.targets + (size_t)e * cv.tstride; tg = make_float2(__ldg(tp), __ldg(tp + 1)); }
        const float2 wg = __ldg(reinterpret_cast<const float2 *>(cv.weights) + e);
        const float *pp = cv.patches + 3 * (size_t)__ldg(pv.kx + t);
        EdgeTerms et;
        edge_terms(pc, __ldg(pp), __ldg(pp + 1), __ldg(pp + 2), tg.x, tg.y, wg.x, wg.y, cv.bounds, cv.loss, et);
        const float wz0 = et.w0 * et.Jz0, wz1 = et.w1 * et.Jz1;      // (w Jz)^T, ba.py:255
        sh[12 * NT + tau] = wz0 * et.Jz0 + wz1 * et.Jz1;             // C term, ba.py:287
        sh[13 * NT + tau] = wz0 * et.r0 + wz1 * et.r1;               // w term, ba.py:292
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
#pragma unroll
          for (int a = 0; a < 6; ++a) { sh[a * NT + tau] = -Ei[a]; sh[(6 + a) * NT + tau] = Ej[a]; }
        }
      }
      __syncthreads();
      // ---- per-track reduction over pattern positions, fixed order ----
      const int ntr = min(Tp, t1 - ts);
      for (int o = tau; o < ntr * outs_per_track; o += NT) {
        const int k2 = o / outs_per_track;
        const int r = o - k2 * outs_per_track;
        const int t = ts + k2;
        if (!STRUCT_ONLY && r < rowlen) {
          const int s = r / 6, comp = r - 6 * s;
          float sum = 0.0f;
          if (pose_free(slot_pose[s], cv)) {                                // ba.py:33-36 mask
            for (int it = slot_ptr[s]; it < slot_ptr[s + 1]; ++it) {
              const int item = slot_items[it];
              const int pp2 = (item >> 1) - p0;
              if (pp2 >= 0 && pp2 < dc) sum += sh[((item & 1) * 6 + comp) * NT + k2 * dc + pp2];
            }
          }
          float *dst = Erows + (size_t)(t - gt0) * rowlen + r;
          *dst = (p0 == 0) ? sum : *dst + sum;
        } else {
          const int comp = STRUCT_ONLY ? r : r - rowlen;                    // 0: C, 1: w
          float sum = 0.0f;
          for (int pp2 = 0; pp2 < dc; ++pp2) sum += sh[(12 + comp) * NT + k2 * dc + pp2];
          float *dst = reinterpret_cast<float *>(cv.Cw + t) + comp;
          *dst = (p0 == 0) ? sum : *dst + sum;
        }
      }
      __syncthreads();
    }

    if (!STRUCT_ONLY) {
      // ---- flush Bjj / vj of this position batch: reduce over kappa, map to the i side, scatter ----
#pragma unroll
      for (int k = 0; k < kAccComps; ++k) sh[k * NT + tau] = active ? acc[k] : 0.0f;
      __syncthreads();
      if (tau < dc) {
        double Bl[21], vj[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) { double s = 0.0; for (int kp = 0; kp < Tp; ++kp) s += (double)sh[k * NT + kp * dc + tau]; Bl[k] = s; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { double s = 0.0; for (int kp = 0; kp < Tp; ++kp) s += (double)sh[(21 + k) * NT + kp * dc + tau]; vj[k] = s; }
        const int pi = pv.pat_i[pat0 + p0 + tau], pj = pv.pat_j[pat0 + p0 + tau];
        const bool fi = pose_free(pi, cv), fj = pose_free(pj, cv);
        const int ri = 6 * (pi - cv.fixedp), rj = 6 * (pj - cv.fixedp);
        // tau < dc means kappa == 0, so this thread's pair constants `pc` are those of position p0 + tau
        if (fj) {
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            red_add(cv.y + rj + a, vj[a]);                                   // ba.py:290
#pragma unroll
            for (int b = 0; b <= a; ++b) red_add(S_at(cv, rj + a, rj + b), Bl[tri(a, b)]);   // Bjj, ba.py:282
          }
        }
        if (fi) {
          double AB[6][6];                     // A * Bjj   (column c of Bjj is its row c)
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double col[6], out[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) col[a] = a >= c ? Bl[tri(a, c)] : Bl[tri(c, a)];
            adjT_apply_d(pc.R, pc.t, col, out);
#pragma unroll
            for (int a = 0; a < 6; ++a) AB[a][c] = out[a];
          }
          double vi[6];
          adjT_apply_d(pc.R, pc.t, vj, vi);
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            red_add(cv.y + ri + a, -vi[a]);                                  // vi = -A vj, ba.py:289
            double row[6];
            adjT_apply_d(pc.R, pc.t, AB[a], row);                            // Bii = (A Bjj) A^T, ba.py:279
#pragma unroll
            for (int b = 0; b <= a; ++b) red_add(S_at(cv, ri + a, ri + b), row[b]);
          }
          if (fj) {                            // Bij = -A Bjj (rows i, cols j); Bji = Bij^T, ba.py:280-281
            if (pi > pj) {
#pragma unroll
              for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) red_add(S_at(cv, ri + a, rj + b), -AB[a][b]);
            } else if (pi < pj) {
#pragma unroll
              for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) red_add(S_at(cv, rj + b, ri + a), -AB[a][b]);
            } else {
#pragma unroll
              for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) red_add(S_at(cv, ri + a, ri + b), -(AB[a][b] + AB[b][a]));
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

// =================================================================================================
// K1b  per track: damped inverse Q and prior-adjusted w (ba.py:296-311; BA: :184)
// =================================================================================================
__global__ void k_track_q(PlanView pv, CallView cv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pv.m) return;
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
constexpr int kSchurRun = 8;

__global__ void __launch_bounds__(kSchurThreads) k_schur(PlanView pv, CallView cv, int tile_tracks) {
  constexpr int NT = kSchurThreads;
  extern __shared__ float smem[];
  const int tau = threadIdx.x;
  const int u = blockIdx.x;
  const int g = pv.u_grp[u];
  const int t0 = pv.u_t0[u], t1 = pv.u_t0[u + 1];
  const int gt0 = pv.g_t0[g];
  const int W = pv.g_W[g];
  const int rowlen = 6 * W;
  const int *slot_pose = pv.slot_pose + 2 * pv.g_pat[g];
  const float *Erows = cv.Est + pv.g_eoff[g];
  float *Es = smem;                               // [tile_tracks][rowlen]
  float *qs = smem + (size_t)tile_tracks * rowlen; // [tile_tracks] Q_k
  float *qws = qs + tile_tracks;                  // [tile_tracks] Q_k w_k
  const int npairs = W * (W + 1) / 2;

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

    for (int tt = t0; tt < t1; tt += tile_tracks) {
      const int nt = min(tile_tracks, t1 - tt);
      const float *src = Erows + (size_t)(tt - gt0) * rowlen;
      for (int o = tau; o < nt * rowlen; o += NT) Es[o] = src[o];
      for (int o = tau; o < nt; o += NT) { const float2 qw = cv.Qw[tt + o]; qs[o] = qw.x; qws[o] = qw.x * qw.y; }
      __syncthreads();
      if (pb == 0) {                              // y -= E Q w   (ba.py:322)
        for (int r = tau; r < rowlen; r += NT) {
          const int pose = slot_pose[r / 6];
          if (pose_free(pose, cv)) {
            double s = 0.0;
            for (int k = 0; k < nt; ++k) s += (double)(qws[k] * Es[k * rowlen + r]);
            red_add(cv.y + 6 * (pose - cv.fixedp) + (r % 6), -s);
          }
        }
      }
      if (have) {
        for (int k0 = 0; k0 < nt; k0 += kSchurRun) {
          float part[36];
#pragma unroll
          for (int k = 0; k < 36; ++k) part[k] = 0.0f;
          const int k1 = min(k0 + kSchurRun, nt);
          for (int k = k0; k < k1; ++k) {
            const float q = qs[k];
            const float *ea = Es + k * rowlen + 6 * a, *eb = Es + k * rowlen + 6 * b;
            float va[6], vb[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) { va[c] = q * ea[c]; vb[c] = eb[c]; }
#pragma unroll
            for (int c = 0; c < 6; ++c)
#pragma unroll
              for (int e2 = 0; e2 < 6; ++e2) part[c * 6 + e2] += va[c] * vb[e2];
          }
#pragma unroll
          for (int k = 0; k < 36; ++k) acc[k] += (double)part[k];
        }
      }
      __syncthreads();
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
}

// =================================================================================================
// K3  reduced solve (ba.py:60-70 block_solve, :5-19 CholeskySolver, :323-325 NaN retry), in fp64.
//   A = S + (ep + lm * diag S) I;  A = L L^T;  dX = A^-1 y.  One CTA; the band window
//   [j, j+bw] x [j, j+bw] lives in shared memory as a circular buffer, columns are eliminated
//   right-looking, the forward substitution rides along, L goes to global (band) storage and the
//   backward substitution streams it back in blocks of rows. A dense system with 6n <= kMaxWindow
//   is the special case bw = 6n - 1.
// =================================================================================================
__global__ void __launch_bounds__(kSolveThreads) k_solve_window(CallView cv, int allow_retry) {
  extern __shared__ double dsm[];
  constexpr int NT = kSolveThreads;
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
  const int M = cv.M, bw = cv.bw, WS = bw + 1, WSP = WS | 1;
  double *win = dsm;                       // [WS][WSP]
  double *z = win + WS * WSP;              // [M]
  double *ls = z + M;                      // [WS]
  __shared__ int s_flag;
  const double *S = cv.S;
  double *L = cv.L;
  const int ld = cv.ld, off = cv.off;
  auto Sg = [&](int r, int c) { return (size_t)r * ld + c + off; };
  const double ep = (double)cv.ep;
  int status = 0;

  for (int attempt = 0; attempt < 2; ++attempt) {
    const double lm = attempt == 0 ? 1e-4 : 1e-3;
    // ---- load rows 0..bw of the window, z = y ----
    for (int r = warp; r < min(WS, M); r += NT / 32)
      for (int c = lane; c <= r; c += 32) {
        double v = S[Sg(r, c)];
        if (c == r) v = v + (ep + lm * v);                       // ba.py:67
        win[(r % WS) * WSP + (c % WS)] = v;
      }
    for (int r = tau; r < M; r += NT) z[r] = cv.y[r];
    if (tau == 0) s_flag = 0;
    bool failed = false;
    for (int j = 0; j < M; ++j) {
      __syncthreads();
      const int jm = j % WS;
      const double piv = win[jm * WSP + jm];
      if (!(piv > 0.0)) { failed = true; break; }                 // potrf info != 0 (incl. NaN), ba.py:11
      const double dg = sqrt(piv), inv = 1.0 / dg;
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
      // rank-1 update of the trailing window, lower triangle
      const int j1 = (j + 1) % WS;
      for (int rr = warp; rr < nb; rr += NT / 32) {
        int rs = j1 + rr; if (rs >= WS) rs -= WS;
        const double lr = ls[rr];
        for (int cc = lane; cc <= rr; cc += 32) {
          int cs = j1 + cc; if (cs >= WS) cs -= WS;
          win[rs * WSP + cs] -= lr * ls[cc];
        }
      }
      // bring in row j + WS (reuses the shared-memory row of the retired row j)
      const int rn = j + WS;
      if (rn < M) {
        for (int x = tau; x <= bw; x += NT) {
          const int c = rn - bw + x;
          double v = S[Sg(rn, c)];
          if (c == rn) v = v + (ep + lm * v);
          win[jm * WSP + (c % WS)] = v;
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
      for (int r = lo + warp; r <= jb; r += NT / 32)
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
// K4  back-substitution dZ = Q (w - E^T dX) (ba.py:328 / :317), disparity retraction + clamp
//     (ba.py:42-44,332-334). One warp per track.
// =================================================================================================
__global__ void k_patches_copy_clamp(const float *__restrict__ in, float *__restrict__ out, int NM) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= NM) return;
  out[3 * (size_t)k] = in[3 * (size_t)k];
  out[3 * (size_t)k + 1] = in[3 * (size_t)k + 1];
  out[3 * (size_t)k + 2] = fminf(fmaxf(in[3 * (size_t)k + 2], 1e-3f), 10.0f);   // clamp hits every patch, ba.py:333
}

__global__ void k_backsub(PlanView pv, CallView cv, int use_dx) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= pv.m) return;
  const float2 qw = cv.Qw[t];
  double dot = 0.0;
  if (use_dx) {
    const int g = pv.t_grp[t];
    const int W = pv.g_W[g], rowlen = 6 * W;
    const int *slot_pose = pv.slot_pose + 2 * pv.g_pat[g];
    const float *row = cv.Est + pv.g_eoff[g] + (size_t)(t - pv.g_t0[g]) * rowlen;
    for (int r = lane; r < rowlen; r += 32) {
      const int s = r / 6;
      const int pose = slot_pose[s];
      if (pose_free(pose, cv)) dot += (double)row[r] * cv.dX[6 * (pose - cv.fixedp) + (r - 6 * s)];
    }
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  if (lane == 0) {
    const float dz = (float)((double)qw.x * ((double)qw.y - dot));
    cv.dZ[t] = dz;
    const size_t k = (size_t)pv.kx[t];
    cv.patches_out[3 * k + 2] = fminf(fmaxf(cv.patches[3 * k + 2] + dz, 1e-3f), 10.0f);
  }
}

// pose retraction T <- Exp(dx) T for every pose of the buffer, dx = 0 outside the window
// (ba.py:47-49,336-337; lietorch/groups.py:153-156)
__global__ void k_pose_retr(CallView cv, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
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
  cv->S = pl->SY; cv->y = pl->SY + sf;
  cv->Est = pl->Est; cv->Cw = pl->Cw; cv->Qw = pl->Qw; cv->dX = pl->dX; cv->dZ = pl->dZ; cv->L = pl->L;
  cv->status = pl->status;
  cv->poses_out = pb->poses_out; cv->patches_out = pb->patches_out;
  return BA_OK;
}

}  // namespace ba

using namespace ba;

extern "C" int ba_assemble(BaPlan *pl, const BaProblem *pb, void *stream_) {
  CallView cv;
  int rc = make_call(pl, pb, &cv);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream_;
  const PlanView &pv = pl->v;
  const bool so = pb->structure_only || cv.n == 0;                 // ba.py:316
  pl->last_n = cv.n; pl->last_fixedp = pb->fixedp;
  pl->ev_mask = 0;
  if (so) {
    BA_MARK(pl, BA_STAGE_EDGE, s);
    k_edge_pass<true><<<pv.n_chunks, kEdgeThreads, 0, s>>>(pv, cv); BA_LAUNCH_CHECK();
  } else {
    BA_MARK(pl, BA_STAGE_ZERO, s);
    BA_CUDA(cudaMemsetAsync(cv.S, 0, (size_t)((cv.y - cv.S) + cv.M) * sizeof(double), s));
    BA_MARK(pl, BA_STAGE_EDGE, s);
    k_edge_pass<false><<<pv.n_chunks, kEdgeThreads, 0, s>>>(pv, cv); BA_LAUNCH_CHECK();
  }
  BA_MARK(pl, BA_STAGE_TRACKQ, s);
  k_track_q<<<(pv.m + 255) / 256, 256, 0, s>>>(pv, cv); BA_LAUNCH_CHECK();
  BA_MARK(pl, BA_STAGE_SCHUR, s);
  if (!so) {
    const int rowmax = 6 * pl->info.max_slots;
    int tile = (int)((64 * 1024 / sizeof(float)) / (rowmax + 2));
    tile = tile < 1 ? 1 : (tile > 128 ? 128 : tile);
    const size_t smem = (size_t)tile * (rowmax + 2) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      BA_CUDA(cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    if (smem > 200 * 1024) return BA_ERR_ARG;
    k_schur<<<pv.n_units, kSchurThreads, smem, s>>>(pv, cv, tile); BA_LAUNCH_CHECK();
  }
  BA_MARK(pl, BA_STAGE_SOLVE, s);      // closes SCHUR; a sharded caller's all-reduce lands in SOLVE's interval
  return BA_OK;
}

extern "C" int ba_plan_reduced_system(const BaPlan *pl, double **ptr, int64_t *n_values) {
  int64_t *n_floats = n_values;
  if (!pl || !ptr || !n_floats || pl->last_fixedp < 0) return BA_ERR_ARG;
  int n, bw, ld, off; int64_t sf;
  layout_for(pl, pl->last_fixedp, &n, &bw, &ld, &off, &sf);
  *ptr = pl->SY;
  *n_floats = sf + 6 * (int64_t)n;
  return BA_OK;
}

extern "C" int ba_solve_update(BaPlan *pl, const BaProblem *pb, void *stream_) {
  CallView cv;
  int rc = make_call(pl, pb, &cv);
  if (rc) return rc;
  if (!pb->poses_out || !pb->patches_out) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  const PlanView &pv = pl->v;
  const bool so = pb->structure_only || cv.n == 0;
  if (!so) {
    if (!(pl->ev_mask & (1u << BA_STAGE_SOLVE))) BA_MARK(pl, BA_STAGE_SOLVE, s);
    static bool attr_set = false;
    if (!attr_set) {
      BA_CUDA(cudaFuncSetAttribute(k_solve_window, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
      BA_CUDA(cudaFuncSetAttribute(k_solve_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
      attr_set = true;
    }
    const int WS = cv.bw + 1, WSP = WS | 1;
    const size_t smem = ((size_t)WS * WSP + cv.M + WS) * sizeof(double);
    if (cv.ld != cv.M && smem <= 227 * 1024 - 64) {
      k_solve_window<<<1, kSolveThreads, smem, s>>>(cv, pb->monodisp ? 1 : 0); BA_LAUNCH_CHECK();
    } else {
      const size_t smem_d = 2 * (size_t)cv.M * sizeof(double);
      if (cv.ld != cv.M || smem_d > 227 * 1024 - 64) return BA_ERR_ARG;    // > 14k unknowns without band structure
      k_solve_dense<<<1, kSolveThreads, smem_d, s>>>(cv, pb->monodisp ? 1 : 0); BA_LAUNCH_CHECK();
    }
  }
  BA_MARK(pl, BA_STAGE_BACKSUB, s);
  k_patches_copy_clamp<<<(pv.NM + 255) / 256, 256, 0, s>>>(pb->patches, pb->patches_out, pv.NM); BA_LAUNCH_CHECK();
  k_backsub<<<(int)(((int64_t)pv.m * 32 + 255) / 256), 256, 0, s>>>(pv, cv, so ? 0 : 1); BA_LAUNCH_CHECK();
  BA_MARK(pl, BA_STAGE_RETR, s);
  if (so) {
    BA_CUDA(cudaMemcpyAsync(pb->poses_out, pb->poses, (size_t)pv.N * 7 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    k_pose_retr<<<(pv.N + 127) / 128, 128, 0, s>>>(cv, pv.N); BA_LAUNCH_CHECK();
  }
  BA_MARK(pl, BA_N_STAGES, s);
  return BA_OK;
}

extern "C" int ba_step(BaPlan *pl, const BaProblem *pb, void *stream) {
  int rc = ba_assemble(pl, pb, stream);
  if (rc) return rc;
  return ba_solve_update(pl, pb, stream);
}

extern "C" int ba_plan_debug_dense(const BaPlan *pl, int32_t n, float *S, float *y, float *dX, float *Q,
                                   float *w, float *dZ, void *stream_) {
  if (!pl || pl->last_fixedp < 0) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  int nn, bw, ld, off; int64_t sf;
  layout_for(pl, pl->last_fixedp, &nn, &bw, &ld, &off, &sf);
  if (n != nn) return BA_ERR_ARG;
  const int M = 6 * nn;
  if (S && M > 0) { k_debug_dense<<<(M * M + 255) / 256, 256, 0, s>>>(pl->SY, M, ld, off, bw, S); BA_LAUNCH_CHECK(); }
  if (y && M > 0) { k_debug_cast<<<(M + 255) / 256, 256, 0, s>>>(pl->SY + sf, y, M); BA_LAUNCH_CHECK(); }
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
