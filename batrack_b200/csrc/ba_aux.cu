// ba_aux.cu — standalone SE3 forward ops (the lietorch_backends surface the BA caller touches,
// main/backend/lietorch/src/lietorch.cpp:286-316, kernels lietorch_gpu.cu:21-283), reprojection
// without Jacobians (projective_ops.py:54-70), and the host-buffer wrapper used for end-to-end timing.
#include <cstdint>
#include <cstring>

#include "ba_internal.h"
#include "ba_math.cuh"
#include "se3_t.cuh"

namespace ba {

// One thread per element. Elements are 7/6/4-float AoS rows; a warp's rows are contiguous, so the
// loads of a warp cover whole 128-byte lines even though each thread's row is 28 bytes.
enum Se3Op { OP_EXP, OP_LOG, OP_INV, OP_MUL, OP_ADJ, OP_ADJT, OP_ACT, OP_ACT4, OP_MAT };

template <int OP, typename T>
__global__ void k_se3(const T *__restrict__ X, const T *__restrict__ Y, T *__restrict__ out, int64_t B) {
  namespace g = se3t;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  if (OP == OP_EXP) {
    T a[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = X[6 * i + c];
    g::store(g::exp<T>(a), out + 7 * i);
  } else if (OP == OP_LOG) {
    T a[6];
    g::log(g::load(X + 7 * i), a);
#pragma unroll
    for (int c = 0; c < 6; ++c) out[6 * i + c] = a[c];
  } else if (OP == OP_INV) {
    g::store(g::inv(g::load(X + 7 * i)), out + 7 * i);
  } else if (OP == OP_MUL) {
    g::store(g::mul(g::load(X + 7 * i), g::load(Y + 7 * i)), out + 7 * i);
  } else if (OP == OP_ADJ || OP == OP_ADJT) {
    const g::P<T> Pz = g::load(X + 7 * i);
    T R[9], a[6], b[6];
    g::qmatrix(Pz.q, R);
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = Y[6 * i + c];
    if (OP == OP_ADJ) g::adj(R, Pz.t, a, b); else g::adjT(R, Pz.t, a, b);
#pragma unroll
    for (int c = 0; c < 6; ++c) out[6 * i + c] = b[c];
  } else if (OP == OP_ACT) {
    const g::P<T> Pz = g::load(X + 7 * i);
    const g::V<T> r = g::qrotate(Pz.q, g::V<T>{Y[3 * i], Y[3 * i + 1], Y[3 * i + 2]});
    out[3 * i] = r.x + Pz.t.x; out[3 * i + 1] = r.y + Pz.t.y; out[3 * i + 2] = r.z + Pz.t.z;
  } else if (OP == OP_ACT4) {
    const g::P<T> Pz = g::load(X + 7 * i);
    const T h = Y[4 * i + 3];
    const g::V<T> r = g::qrotate(Pz.q, g::V<T>{Y[4 * i], Y[4 * i + 1], Y[4 * i + 2]});
    out[4 * i] = r.x + Pz.t.x * h; out[4 * i + 1] = r.y + Pz.t.y * h; out[4 * i + 2] = r.z + Pz.t.z * h; out[4 * i + 3] = h;
  } else if (OP == OP_MAT) {
    const g::P<T> Pz = g::load(X + 7 * i);
    T R[9];
    g::qmatrix(Pz.q, R);
    T *M = out + 16 * i;
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = Pz.t.x;
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = Pz.t.y;
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = Pz.t.z;
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
  }
}

template <int OP, typename T>
static int launch_se3(const T *X, const T *Y, T *out, int64_t B, void *stream) {
  if (B < 0 || (B > 0 && (!X || !out))) return BA_ERR_ARG;
  if (B == 0) return BA_OK;
  k_se3<OP, T><<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, Y, out, B);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

__global__ void k_reproject(const float *__restrict__ poses, const float *__restrict__ patches,
                            const float *__restrict__ intr, const int64_t *__restrict__ ii,
                            const int64_t *__restrict__ jj, const int64_t *__restrict__ kk, int64_t E,
                            int N, int NM, int tonly, float *__restrict__ coords, float *__restrict__ valid) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t i = ii[e], j = jj[e], k = kk[e];
  if (i < 0 || i >= N || j < 0 || j >= N || k < 0 || k >= NM) {
    coords[2 * e] = coords[2 * e + 1] = __int_as_float(0x7fc00000);
    if (valid) valid[e] = 0.0f;
    return;
  }
  Pose G = pose_mul(pose_load(poses + 7 * j), pose_inv(pose_load(poses + 7 * i)));
  if (tonly) G.q = {0.0f, 0.0f, 0.0f, 1.0f};                   // projective_ops.py:63-64
  G.q = qnormalize(G.q);
  const float *p = patches + 3 * k, *Ki = intr + 4 * i, *Kj = intr + 4 * j;
  const float x0 = (p[0] - Ki[2]) / Ki[0], y0 = (p[1] - Ki[3]) / Ki[1], d = p[2];
  Vec3 r = qrotate(G.q, {x0, y0, 1.0f});
  const float X = r.x + G.t.x * d, Y = r.y + G.t.y * d, Z = r.z + G.t.z * d;
  const float dc = 1.0f / fmaxf(Z, 1e-2f);
  coords[2 * e] = Kj[0] * (dc * X) + Kj[2];
  coords[2 * e + 1] = Kj[1] * (dc * Y) + Kj[3];
  if (valid) valid[e] = Z > kMinDepth ? 1.0f : 0.0f;
}


// transform() with every optional output of projective_ops.py:54-105: the reprojection (with or without the
// inverse-depth channel of proj(depth=True), :47-50), the validity mask, and the analytic Jacobians Ji, Jj (2x6) and
// Jz (2x1) of :72-100 — the same edge_terms() the fused BA edge pass uses, materialised for callers that want them.
__global__ void k_transform_full(const float *__restrict__ poses, const float *__restrict__ patches,
                                 const float *__restrict__ intr, const int64_t *__restrict__ ii,
                                 const int64_t *__restrict__ jj, const int64_t *__restrict__ kk, int64_t E,
                                 int N, int NM, int tonly, int depth, float *__restrict__ coords, float *__restrict__ valid,
                                 float *__restrict__ Ji, float *__restrict__ Jj, float *__restrict__ Jz) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t i = ii[e], j = jj[e], k = kk[e];
  const int cs = depth ? 3 : 2;
  const float nanv = __int_as_float(0x7fc00000);
  if (i < 0 || i >= N || j < 0 || j >= N || k < 0 || k >= NM) {
    for (int c = 0; c < cs; ++c) coords[cs * e + c] = nanv;
    if (valid) valid[e] = 0.0f;
    if (Ji) for (int c = 0; c < 12; ++c) Ji[12 * e + c] = nanv;
    if (Jj) for (int c = 0; c < 12; ++c) Jj[12 * e + c] = nanv;
    if (Jz) { Jz[2 * e] = nanv; Jz[2 * e + 1] = nanv; }
    return;
  }
  Pose G = pose_mul(pose_load(poses + 7 * j), pose_inv(pose_load(poses + 7 * i)));   // :61
  if (tonly) G.q = {0.0f, 0.0f, 0.0f, 1.0f};                                         // :63-64
  G.q = qnormalize(G.q);
  PairConst pc;
  qmatrix(G.q, pc.R);
  pc.t = G.t;
  const float *Ki = intr + 4 * i, *Kj = intr + 4 * j, *p = patches + 3 * k;
  pc.fxi = Ki[0]; pc.fyi = Ki[1]; pc.cxi = Ki[2]; pc.cyi = Ki[3];
  pc.fxj = Kj[0]; pc.fyj = Kj[1]; pc.cxj = Kj[2]; pc.cyj = Kj[3];
  // exact divisions here (this is the materialising path; the fused edge pass uses the SFU forms)
  const float x0 = (p[0] - pc.cxi) / pc.fxi, y0 = (p[1] - pc.cyi) / pc.fyi, d0 = p[2];
  const float X = pc.R[0] * x0 + pc.R[1] * y0 + pc.R[2] + pc.t.x * d0;
  const float Y = pc.R[3] * x0 + pc.R[4] * y0 + pc.R[5] + pc.t.y * d0;
  const float Z = pc.R[6] * x0 + pc.R[7] * y0 + pc.R[8] + pc.t.z * d0;
  const float H = d0;
  const float dc = 1.0f / fmaxf(Z, 1e-2f);                                           // :43
  coords[cs * e] = pc.fxj * (dc * X) + pc.cxj;
  coords[cs * e + 1] = pc.fyj * (dc * Y) + pc.cyj;
  if (depth) coords[cs * e + 2] = dc * H;                                            // :47
  if (valid) valid[e] = Z > kMinDepth ? 1.0f : 0.0f;                                 // :100 / :103
  if (Ji || Jj || Jz) {
    const float dj = fabsf(Z) > kMinDepth ? 1.0f / Z : 0.0f;                         // :80-81
    const float a = pc.fxj * dj, b = -pc.fxj * X * dj * dj, c = pc.fyj * dj, f = -pc.fyj * Y * dj * dj;
    float J0[6] = {a * H, 0.0f, b * H, b * Y, a * Z - b * X, -a * Y};                // Jp Ja, :83-95
    float J1[6] = {0.0f, c * H, f * H, -c * Z + f * Y, -f * X, c * X};
    if (Jj) {
#pragma unroll
      for (int q = 0; q < 6; ++q) { Jj[12 * e + q] = J0[q]; Jj[12 * e + 6 + q] = J1[q]; }
    }
    if (Ji) {                                                                        // Ji = -Ad(Gij)^T Jj, :96
      float o0[6], o1[6];
      adjT_apply(pc.R, pc.t, J0, o0);
      adjT_apply(pc.R, pc.t, J1, o1);
#pragma unroll
      for (int q = 0; q < 6; ++q) { Ji[12 * e + q] = -o0[q]; Ji[12 * e + 6 + q] = -o1[q]; }
    }
    if (Jz) {                                                                        // Jp * Gij.matrix()[:, 3], :98
      Jz[2 * e] = a * pc.t.x + b * pc.t.z;
      Jz[2 * e + 1] = c * pc.t.y + f * pc.t.z;
    }
  }
}

// point_cloud (projective_ops.py:107-109): T_ix^-1 * iproj(patch), one thread per patch
__global__ void k_point_cloud(const float *__restrict__ poses, const float *__restrict__ patches,
                              const float *__restrict__ intr, const int64_t *__restrict__ ix, int64_t n, int N,
                              float *__restrict__ out) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t f = ix[k];
  if (f < 0 || f >= N) { for (int c = 0; c < 4; ++c) out[4 * k + c] = __int_as_float(0x7fc00000); return; }
  const Pose Ti = pose_inv(pose_load(poses + 7 * f));
  const float *K = intr + 4 * f, *p = patches + 3 * k;
  const float x0 = (p[0] - K[2]) / K[0], y0 = (p[1] - K[3]) / K[1], d = p[2];
  Pose Tl = Ti;
  Tl.q = qnormalize(Ti.q);                                                  // act4 re-loads (normalises) its operand
  const Vec3 r = qrotate(Tl.q, {x0, y0, 1.0f});
  out[4 * k] = r.x + Tl.t.x * d; out[4 * k + 1] = r.y + Tl.t.y * d; out[4 * k + 2] = r.z + Tl.t.z * d; out[4 * k + 3] = d;
}

// back_proj (projective_ops.py:129-152): pixel + depth -> homogeneous camera (or world, with c2w) point
__global__ void k_back_proj(const float *__restrict__ xy, const float *__restrict__ depth, const float *__restrict__ intr,
                            const float *__restrict__ c2w, int B, int64_t n, float *__restrict__ P) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * n) return;
  const int b = (int)(idx / n);
  const float *K = intr + 4 * b;
  const float D = depth[idx];
  const float X = (xy[2 * idx] - K[2]) / K[0], Y = (xy[2 * idx + 1] - K[3]) / K[1];
  float v[4] = {X * D, Y * D, D, 1.0f};
  if (c2w) {
    const float *T = c2w + 16 * b;
    float o[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) o[r] = T[4 * r] * v[0] + T[4 * r + 1] * v[1] + T[4 * r + 2] * v[2] + T[4 * r + 3] * v[3];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = o[r];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) P[4 * idx + r] = v[r];
}

// proj_to_frames (projective_ops.py:154-176): world points into S cameras, plain 1/Z like the reference
__global__ void k_proj_to_frames(const float *__restrict__ P, const float *__restrict__ intr, const float *__restrict__ w2c,
                                 int B, int S, int64_t n, float *__restrict__ xy) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * S * n) return;
  const int64_t bs = idx / n, k = idx - bs * n;
  const int b = (int)(bs / S);
  const float *T = w2c + 16 * bs, *K = intr + 4 * bs, *p = P + 4 * ((int64_t)b * n + k);
  const float Xc = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3] * p[3];
  const float Yc = T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7] * p[3];
  const float Dc = T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11] * p[3];
  const float dc = 1.0f / Dc;
  xy[2 * idx] = K[0] * (Xc * dc) + K[2];
  xy[2 * idx + 1] = K[1] * (Yc * dc) + K[3];
}

}  // namespace ba

using namespace ba;

#define SE3_ENTRY(PFX, T)                                                                                                   \
  extern "C" int PFX##_expm(const T *a, T *X, int64_t B, void *s) { return launch_se3<OP_EXP, T>(a, nullptr, X, B, s); }      \
  extern "C" int PFX##_logm(const T *X, T *a, int64_t B, void *s) { return launch_se3<OP_LOG, T>(X, nullptr, a, B, s); }      \
  extern "C" int PFX##_inv(const T *X, T *Y, int64_t B, void *s) { return launch_se3<OP_INV, T>(X, nullptr, Y, B, s); }       \
  extern "C" int PFX##_mul(const T *X, const T *Y, T *Z, int64_t B, void *s) { return (B > 0 && !Y) ? BA_ERR_ARG : launch_se3<OP_MUL, T>(X, Y, Z, B, s); }   \
  extern "C" int PFX##_adj(const T *X, const T *a, T *b, int64_t B, void *s) { return (B > 0 && !a) ? BA_ERR_ARG : launch_se3<OP_ADJ, T>(X, a, b, B, s); }   \
  extern "C" int PFX##_adjT(const T *X, const T *a, T *b, int64_t B, void *s) { return (B > 0 && !a) ? BA_ERR_ARG : launch_se3<OP_ADJT, T>(X, a, b, B, s); } \
  extern "C" int PFX##_act(const T *X, const T *p, T *q, int64_t B, void *s) { return (B > 0 && !p) ? BA_ERR_ARG : launch_se3<OP_ACT, T>(X, p, q, B, s); }   \
  extern "C" int PFX##_act4(const T *X, const T *p, T *q, int64_t B, void *s) { return (B > 0 && !p) ? BA_ERR_ARG : launch_se3<OP_ACT4, T>(X, p, q, B, s); } \
  extern "C" int PFX##_as_matrix(const T *X, T *M, int64_t B, void *s) { return launch_se3<OP_MAT, T>(X, nullptr, M, B, s); }
SE3_ENTRY(se3, float)
SE3_ENTRY(se3d, double)
#undef SE3_ENTRY

extern "C" int ba_reproject(const float *poses, const float *patches, const float *intrinsics, const int64_t *ii,
                            const int64_t *jj, const int64_t *kk, int64_t E, int32_t N, int32_t NM, int32_t tonly,
                            float *coords, float *valid, void *stream) {
  if (E < 0 || N <= 0 || NM <= 0) return BA_ERR_ARG;
  if (E == 0) return BA_OK;
  if (!poses || !patches || !intrinsics || !ii || !jj || !kk || !coords) return BA_ERR_ARG;
  k_reproject<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(poses, patches, intrinsics, ii, jj, kk, E,
                                                                            N, NM, tonly, coords, valid);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

extern "C" int ba_transform(const float *poses, const float *patches, const float *intrinsics, const int64_t *ii,
                            const int64_t *jj, const int64_t *kk, int64_t E, int32_t N, int32_t NM, int32_t tonly,
                            int32_t depth, float *coords, float *valid, float *Ji, float *Jj, float *Jz, void *stream) {
  if (E < 0 || N <= 0 || NM <= 0) return BA_ERR_ARG;
  if (E == 0) return BA_OK;
  if (!poses || !patches || !intrinsics || !ii || !jj || !kk || !coords) return BA_ERR_ARG;
  k_transform_full<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(poses, patches, intrinsics, ii, jj, kk, E, N, NM,
                                                                                  tonly, depth, coords, valid, Ji, Jj, Jz);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

extern "C" int ba_point_cloud(const float *poses, const float *patches, const float *intrinsics, const int64_t *ix,
                              int64_t n, int32_t N, float *out, void *stream) {
  if (n < 0 || N <= 0) return BA_ERR_ARG;
  if (n == 0) return BA_OK;
  if (!poses || !patches || !intrinsics || !ix || !out) return BA_ERR_ARG;
  k_point_cloud<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(poses, patches, intrinsics, ix, n, N, out);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

extern "C" int ba_back_proj(const float *xy, const float *depth, const float *intrinsics, const float *c2w, int32_t B,
                            int64_t n, float *P, void *stream) {
  if (B < 0 || n < 0) return BA_ERR_ARG;
  if ((int64_t)B * n == 0) return BA_OK;
  if (!xy || !depth || !intrinsics || !P) return BA_ERR_ARG;
  k_back_proj<<<(unsigned)(((int64_t)B * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xy, depth, intrinsics, c2w, B, n, P);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

extern "C" int ba_proj_to_frames(const float *P, const float *intrinsics, const float *w2c, int32_t B, int32_t S, int64_t n,
                                 float *xy, void *stream) {
  if (B < 0 || S < 0 || n < 0) return BA_ERR_ARG;
  if ((int64_t)B * S * n == 0) return BA_OK;
  if (!P || !intrinsics || !w2c || !xy) return BA_ERR_ARG;
  k_proj_to_frames<<<(unsigned)(((int64_t)B * S * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, intrinsics, w2c, B, S, n, xy);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// ---- host-buffer entry points -------------------------------------------------------------------
// The per-call inputs (poses, patches, monodisp, intrinsics, targets, weights, lmbda_vec) are HOST arrays;
// they are copied into one of two plan-owned device staging slots, ba_step runs on the caller's stream, and the
// two results are copied back. Indices live in the plan already. The copies run on the plan's own H2D / D2H
// streams, so that the upload of call k+1 and the download of call k-1 overlap the kernels of call k (with
// pinned host memory); consecutive calls serialise on the caller's stream because they share the workspace.
namespace {
struct HostPipe {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t in_ready[2] = {nullptr, nullptr}, computed[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
  char *stage[2] = {nullptr, nullptr};
  size_t bytes = 0;
  unsigned long long seq = 0;
  unsigned pre_mask = 0;         // inputs of call `seq` that ba_prefetch_host_async has already put on the upload stream
  bool pre_waited = false;       // ... which has waited for the slot's previous user then
  const void *out_ptr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // host arrays the call of each slot downloads into (poses, patches)
  size_t out_len[2][2] = {{0, 0}, {0, 0}};
};
enum { PRE_MONO = 1, PRE_INTR = 2, PRE_TG = 4, PRE_W = 8, PRE_LAM = 16 };

void host_pipe_destroy(void *p_) {
  HostPipe *p = static_cast<HostPipe *>(p_);
  if (!p) return;
  if (p->h2d) cudaStreamSynchronize(p->h2d);
  if (p->d2h) cudaStreamSynchronize(p->d2h);
  for (int k = 0; k < 2; ++k) {
    if (p->stage[k]) cudaFree(p->stage[k]);
    if (p->in_ready[k]) cudaEventDestroy(p->in_ready[k]);
    if (p->computed[k]) cudaEventDestroy(p->computed[k]);
    if (p->out_done[k]) cudaEventDestroy(p->out_done[k]);
  }
  if (p->h2d) cudaStreamDestroy(p->h2d);
  if (p->d2h) cudaStreamDestroy(p->d2h);
  delete p;
}
}  // namespace

// Upload half: the host arrays of `ph` go to the next staging slot on the plan's upload stream, `stream` is made to wait
// for them (and for the previous user of the slot), and `pd` receives the device-side problem (outputs pointing into
// the slot). Download half (ba_unstage_host_async): once `stream` has computed, the two results return to the host
// arrays of `ph` on the download stream.
namespace {
struct StageLayout { size_t o_pose, o_pat, o_mono, o_intr, o_tg, o_w, o_lam, o_pout, o_qout, total; };
StageLayout stage_layout(const BaPlan *pl) {
  const size_t N = pl->v.N, NM = pl->v.NM, E = (size_t)pl->v.E, m = pl->v.m, f = sizeof(float);
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  StageLayout L;
  L.o_pose = 0; L.o_pat = L.o_pose + up(7 * N * f); L.o_mono = L.o_pat + up(3 * NM * f); L.o_intr = L.o_mono + up(NM * f);
  L.o_tg = L.o_intr + up(4 * N * f); L.o_w = L.o_tg + up(2 * E * f); L.o_lam = L.o_w + up(2 * E * f); L.o_pout = L.o_lam + up(m * f);
  L.o_qout = L.o_pout + up(7 * N * f); L.total = L.o_qout + up(3 * NM * f);
  return L;
}
int host_pipe_get(BaPlan *pl, size_t total, HostPipe **out) {
  HostPipe *hp = static_cast<HostPipe *>(pl->host_pipe);
  if (!hp) {
    hp = new HostPipe();
    pl->host_pipe = hp;
    pl->host_pipe_destroy = host_pipe_destroy;
    BA_CUDA(cudaStreamCreateWithFlags(&hp->h2d, cudaStreamNonBlocking));
    BA_CUDA(cudaStreamCreateWithFlags(&hp->d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      BA_CUDA(cudaEventCreateWithFlags(&hp->in_ready[k], cudaEventDisableTiming));
      BA_CUDA(cudaEventCreateWithFlags(&hp->computed[k], cudaEventDisableTiming));
      BA_CUDA(cudaEventCreateWithFlags(&hp->out_done[k], cudaEventDisableTiming));
    }
    BA_CUDA(cudaMalloc((void **)&hp->stage[0], total));
    BA_CUDA(cudaMalloc((void **)&hp->stage[1], total));
    hp->bytes = total;
  }
  if (hp->bytes < total) {                             // a capacity plan re-derived for a larger graph: grow the slots
    BA_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < 2; ++k) { BA_CUDA(cudaFree(hp->stage[k])); hp->stage[k] = nullptr; }
    BA_CUDA(cudaMalloc((void **)&hp->stage[0], total));
    BA_CUDA(cudaMalloc((void **)&hp->stage[1], total));
    hp->bytes = total;
    hp->seq = 0; hp->pre_mask = 0; hp->pre_waited = false;
  }
  *out = hp;
  return BA_OK;
}
}  // namespace

// Early upload for the NEXT ba_stage_host_async / ba_step_host_async call: the inputs that do not depend on the previous
// call's results (targets, weights, intrinsics, monodisp, lmbda_vec — whichever pointers of `ph` are non-NULL; poses /
// patches are ignored) go to that call's staging slot now, on the upload stream, while the previous call still computes.
// The next call then uploads only what is missing. This is what lets DEPENDENT steps (iteration k+1 starts from the host
// results of iteration k, main/batrack.py:869-884) hide the 20 MB of per-step observations behind the kernels.
extern "C" int ba_prefetch_host_async(BaPlan *pl, const BaProblem *ph, void *stream_) {
  (void)stream_;
  if (!pl || !ph) return BA_ERR_ARG;
  if (ph->targets && ph->targets_stride != 0 && ph->targets_stride != 2) return BA_ERR_ARG;
  if (int rc = ba::plan_finalize(pl)) return rc;
  const StageLayout L = stage_layout(pl);
  HostPipe *hp = nullptr;
  if (int rc = host_pipe_get(pl, L.total, &hp)) return rc;
  const size_t N = pl->v.N, NM = pl->v.NM, E = (size_t)pl->v.E, m = pl->v.m, f = sizeof(float);
  const int slot = (int)(hp->seq & 1);
  char *d = hp->stage[slot];
  if (hp->seq >= 2 && !hp->pre_waited) BA_CUDA(cudaStreamWaitEvent(hp->h2d, hp->computed[slot], 0));   // inputs of call seq-2 consumed
  hp->pre_waited = true;
  if (ph->monodisp) { BA_CUDA(cudaMemcpyAsync(d + L.o_mono, ph->monodisp, NM * f, cudaMemcpyHostToDevice, hp->h2d)); hp->pre_mask |= PRE_MONO; }
  if (ph->intrinsics) { BA_CUDA(cudaMemcpyAsync(d + L.o_intr, ph->intrinsics, 4 * N * f, cudaMemcpyHostToDevice, hp->h2d)); hp->pre_mask |= PRE_INTR; }
  if (ph->targets) { BA_CUDA(cudaMemcpyAsync(d + L.o_tg, ph->targets, 2 * E * f, cudaMemcpyHostToDevice, hp->h2d)); hp->pre_mask |= PRE_TG; }
  if (ph->weights) { BA_CUDA(cudaMemcpyAsync(d + L.o_w, ph->weights, 2 * E * f, cudaMemcpyHostToDevice, hp->h2d)); hp->pre_mask |= PRE_W; }
  if (ph->lmbda_vec) { BA_CUDA(cudaMemcpyAsync(d + L.o_lam, ph->lmbda_vec, m * f, cudaMemcpyHostToDevice, hp->h2d)); hp->pre_mask |= PRE_LAM; }
  return BA_OK;
}

extern "C" int ba_stage_host_async(BaPlan *pl, const BaProblem *ph, BaProblem *pd_out, void *stream_) {
  if (!pl || !ph || !pd_out || !ph->poses || !ph->patches || !ph->intrinsics || !ph->targets || !ph->weights ||
      !ph->poses_out || !ph->patches_out)
    return BA_ERR_ARG;
  if (ph->targets_stride != 0 && ph->targets_stride != 2) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  if (int rc = ba::plan_finalize(pl)) return rc;
  const size_t N = pl->v.N, NM = pl->v.NM, E = (size_t)pl->v.E, m = pl->v.m;
  const size_t f = sizeof(float);
  const StageLayout L = stage_layout(pl);
  const size_t o_pose = L.o_pose, o_pat = L.o_pat, o_mono = L.o_mono, o_intr = L.o_intr, o_tg = L.o_tg, o_w = L.o_w, o_lam = L.o_lam,
               o_pout = L.o_pout, o_qout = L.o_qout;
  HostPipe *hp = nullptr;
  if (int rc = host_pipe_get(pl, L.total, &hp)) return rc;
  const int slot = (int)(hp->seq & 1);
  char *d = hp->stage[slot];
  if (hp->seq >= 2 && !hp->pre_waited) BA_CUDA(cudaStreamWaitEvent(hp->h2d, hp->computed[slot], 0));   // inputs of call seq-2 consumed
  const unsigned pre = hp->pre_mask;                   // already on the upload stream (ba_prefetch_host_async)
  hp->pre_mask = 0; hp->pre_waited = false;
  BaProblem pd = *ph;
#define H2D(field, off, bytes, bit)                                                                     \
  do { if (!(pre & (bit))) BA_CUDA(cudaMemcpyAsync(d + (off), ph->field, (bytes), cudaMemcpyHostToDevice, hp->h2d)); \
       pd.field = (const float *)(d + (off)); } while (0)
  // Host-buffer hazard: an input array that a call still in flight downloads INTO (iteration k+1 starting from the host
  // results of iteration k, main/batrack.py:869-884) is uploaded only after that download — ordered on the device, so a
  // caller may enqueue dependent steps back to back without waiting on the host in between.
  for (int k = 0; k < 2; ++k)
    for (int a = 0; a < 2; ++a) {
      const uintptr_t o = reinterpret_cast<uintptr_t>(hp->out_ptr[k][a]);
      if (!o) continue;
      auto overlaps = [&](const void *in, size_t len) {
        const uintptr_t i0 = reinterpret_cast<uintptr_t>(in);
        return in && i0 < o + hp->out_len[k][a] && o < i0 + len;
      };
      if (overlaps(ph->poses, 7 * N * f) || overlaps(ph->patches, 3 * NM * f) || overlaps(ph->monodisp, NM * f) ||
          overlaps(ph->intrinsics, 4 * N * f) || overlaps(ph->targets, 2 * E * f) || overlaps(ph->weights, 2 * E * f)) {
        BA_CUDA(cudaStreamWaitEvent(hp->h2d, hp->out_done[k], 0));
        break;
      }
    }
  H2D(poses, o_pose, 7 * N * f, 0);
  H2D(patches, o_pat, 3 * NM * f, 0);
  if (ph->monodisp) H2D(monodisp, o_mono, NM * f, PRE_MONO);
  H2D(intrinsics, o_intr, 4 * N * f, PRE_INTR);
  H2D(targets, o_tg, 2 * E * f, PRE_TG);
  H2D(weights, o_w, 2 * E * f, PRE_W);
  if (ph->lmbda_vec) H2D(lmbda_vec, o_lam, m * f, PRE_LAM);
#undef H2D
  pd.targets_stride = 2;
  BA_CUDA(cudaEventRecord(hp->in_ready[slot], hp->h2d));
  BA_CUDA(cudaStreamWaitEvent(s, hp->in_ready[slot], 0));
  if (hp->seq >= 2) BA_CUDA(cudaStreamWaitEvent(s, hp->out_done[slot], 0));          // results of call seq-2 downloaded
  pd.poses_out = (float *)(d + o_pout);
  pd.patches_out = (float *)(d + o_qout);
  *pd_out = pd;
  return BA_OK;
}

extern "C" int ba_unstage_host_async(BaPlan *pl, const BaProblem *ph, const BaProblem *pd, void *stream_) {
  if (!pl || !ph || !pd || !pl->host_pipe || !ph->poses_out || !ph->patches_out) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  HostPipe *hp = static_cast<HostPipe *>(pl->host_pipe);
  const size_t N = pl->v.N, NM = pl->v.NM, f = sizeof(float);
  const int slot = (int)(hp->seq & 1);
  BA_CUDA(cudaEventRecord(hp->computed[slot], s));
  BA_CUDA(cudaStreamWaitEvent(hp->d2h, hp->computed[slot], 0));
  BA_CUDA(cudaMemcpyAsync(ph->poses_out, pd->poses_out, 7 * N * f, cudaMemcpyDeviceToHost, hp->d2h));
  BA_CUDA(cudaMemcpyAsync(ph->patches_out, pd->patches_out, 3 * NM * f, cudaMemcpyDeviceToHost, hp->d2h));
  BA_CUDA(cudaEventRecord(hp->out_done[slot], hp->d2h));
  hp->out_ptr[slot][0] = ph->poses_out; hp->out_len[slot][0] = 7 * N * f;
  hp->out_ptr[slot][1] = ph->patches_out; hp->out_len[slot][1] = 3 * NM * f;
  hp->seq++;
  return BA_OK;
}

extern "C" int ba_step_host_async(BaPlan *pl, const BaProblem *ph, void *stream_) {
  BaProblem pd;
  int rc = ba_stage_host_async(pl, ph, &pd, stream_);
  if (rc) return rc;
  rc = ba_step(pl, &pd, stream_);
  if (rc) return rc;
  return ba_unstage_host_async(pl, ph, &pd, stream_);
}

// Orders `stream` after the download of every call submitted so far (so an event recorded on it next sees the
// results on the host); block != 0 also waits for it on the host.
extern "C" int ba_host_sync(BaPlan *pl, void *stream_, int block) {
  if (!pl) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  HostPipe *hp = static_cast<HostPipe *>(pl->host_pipe);
  if (hp && hp->seq > 0) {
    BA_CUDA(cudaStreamWaitEvent(s, hp->out_done[(hp->seq - 1) & 1], 0));
    if (hp->seq > 1) BA_CUDA(cudaStreamWaitEvent(s, hp->out_done[hp->seq & 1], 0));
  }
  if (block) BA_CUDA(cudaStreamSynchronize(s));
  return BA_OK;
}

// One call, synchronous: ba_step_host_async + wait.
extern "C" int ba_step_host(BaPlan *pl, const BaProblem *ph, void *stream_) {
  int rc = ba_step_host_async(pl, ph, stream_);
  if (rc) return rc;
  return ba_host_sync(pl, stream_, 1);
}

// ---- trajectory hand-off: main/batrack.py:223-228 get_pose, :898-915 terminate, :1080-1088 get_results.
// Frame t is either a keyframe (slot[t] >= 0: its pose sits in the pose buffer) or was dropped by keyframe() and carries
// delta[t] = (t0, dP): pose(t) = dP * pose(t0), recursively. One thread per frame walks its chain, composes it,
// inverts (world-from-camera) and writes the 7-vector in terminate()'s order [tx ty tz qw qx qy qz] and / or the 4x4
// matrix get_results() stores as cams_T_world.
namespace ba {
__global__ void k_trajectory(const float *__restrict__ poses, const int *__restrict__ slot, const int *__restrict__ t0,
                             const float *__restrict__ dP, int T, int max_chain, float *__restrict__ out7, float *__restrict__ out44,
                             int *__restrict__ err) {
  namespace g = se3t;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  // chain length first, then compose from the keyframe outwards — the order of the reference's recursion
  // (get_pose(t) = dP_t * get_pose(t0)), so the roundings agree
  int cur = t, L = 0;
  while (slot[cur] < 0) {
    cur = t0[cur];
    if (cur < 0 || cur >= T || ++L > max_chain) { atomicOr(err, 1); return; }   // a frame with neither pose nor delta
  }
  g::P<float> X = g::load(poses + 7 * (size_t)slot[cur]);
  for (int k = L - 1; k >= 0; --k) {
    int c = t;
    for (int s2 = 0; s2 < k; ++s2) c = t0[c];
    X = g::mul(g::load(dP + 7 * (size_t)c), X);
  }
  X = g::inv(X);
  if (out7) {
    float *o = out7 + 7 * (size_t)t;
    o[0] = X.t.x; o[1] = X.t.y; o[2] = X.t.z; o[3] = X.q.w; o[4] = X.q.x; o[5] = X.q.y; o[6] = X.q.z;   // batrack.py:908
  }
  if (out44) {
    float R[9];
    g::qmatrix(X.q, R);
    float *m = out44 + 16 * (size_t)t;
    m[0] = R[0]; m[1] = R[1]; m[2] = R[2]; m[3] = X.t.x;
    m[4] = R[3]; m[5] = R[4]; m[6] = R[5]; m[7] = X.t.y;
    m[8] = R[6]; m[9] = R[7]; m[10] = R[8]; m[11] = X.t.z;
    m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
  }
}
}  // namespace ba

extern "C" int ba_trajectory(const float *poses, const int32_t *slot, const int32_t *t0, const float *dP, int32_t n_frames,
                             float *out7, float *out44, int32_t *err, void *stream) {
  if (n_frames < 0) return BA_ERR_ARG;
  if (n_frames == 0) return BA_OK;
  if (!poses || !slot || !t0 || !dP || !err || (!out7 && !out44)) return BA_ERR_ARG;
  ba::k_trajectory<<<(n_frames + 127) / 128, 128, 0, (cudaStream_t)stream>>>(poses, slot, t0, dP, n_frames, n_frames, out7, out44, err);
  BA_LAUNCH_CHECK();
  return BA_OK;
}
