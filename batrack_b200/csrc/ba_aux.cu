// ba_aux.cu — standalone SE3 forward ops (the lietorch_backends surface the BA caller touches,
// main/backend/lietorch/src/lietorch.cpp:286-316, kernels lietorch_gpu.cu:21-283), reprojection
// without Jacobians (projective_ops.py:54-70), and the host-buffer wrapper used for end-to-end timing.
#include <cstring>

#include "ba_internal.h"
#include "ba_math.cuh"

namespace ba {

// One thread per element. Elements are 7/6/4-float AoS rows; a warp's rows are contiguous, so the
// loads of a warp cover whole 128-byte lines even though each thread's row is 28 bytes.
enum Se3Op { OP_EXP, OP_LOG, OP_INV, OP_MUL, OP_ADJ, OP_ADJT, OP_ACT, OP_ACT4, OP_MAT };

template <int OP>
__global__ void k_se3(const float *__restrict__ X, const float *__restrict__ Y, float *__restrict__ out, int64_t B) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  if (OP == OP_EXP) {
    float a[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = X[6 * i + c];
    pose_store(pose_exp(a), out + 7 * i);
  } else if (OP == OP_LOG) {
    float a[6];
    pose_log(pose_load(X + 7 * i), a);
#pragma unroll
    for (int c = 0; c < 6; ++c) out[6 * i + c] = a[c];
  } else if (OP == OP_INV) {
    pose_store(pose_inv(pose_load(X + 7 * i)), out + 7 * i);
  } else if (OP == OP_MUL) {
    pose_store(pose_mul(pose_load(X + 7 * i), pose_load(Y + 7 * i)), out + 7 * i);
  } else if (OP == OP_ADJ || OP == OP_ADJT) {
    Pose P = pose_load(X + 7 * i);
    float R[9], a[6], b[6];
    qmatrix(P.q, R);
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = Y[6 * i + c];
    if (OP == OP_ADJ) adj_apply(R, P.t, a, b); else adjT_apply(R, P.t, a, b);
#pragma unroll
    for (int c = 0; c < 6; ++c) out[6 * i + c] = b[c];
  } else if (OP == OP_ACT) {
    Pose P = pose_load(X + 7 * i);
    Vec3 r = qrotate(P.q, {Y[3 * i], Y[3 * i + 1], Y[3 * i + 2]});
    out[3 * i] = r.x + P.t.x; out[3 * i + 1] = r.y + P.t.y; out[3 * i + 2] = r.z + P.t.z;
  } else if (OP == OP_ACT4) {
    Pose P = pose_load(X + 7 * i);
    const float h = Y[4 * i + 3];
    Vec3 r = qrotate(P.q, {Y[4 * i], Y[4 * i + 1], Y[4 * i + 2]});
    out[4 * i] = r.x + P.t.x * h; out[4 * i + 1] = r.y + P.t.y * h; out[4 * i + 2] = r.z + P.t.z * h; out[4 * i + 3] = h;
  } else if (OP == OP_MAT) {
    Pose P = pose_load(X + 7 * i);
    float R[9];
    qmatrix(P.q, R);
    float *T = out + 16 * i;
    T[0] = R[0]; T[1] = R[1]; T[2] = R[2]; T[3] = P.t.x;
    T[4] = R[3]; T[5] = R[4]; T[6] = R[5]; T[7] = P.t.y;
    T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = P.t.z;
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
  }
}

template <int OP>
static int launch_se3(const float *X, const float *Y, float *out, int64_t B, void *stream) {
  if (B < 0 || (B > 0 && (!X || !out))) return BA_ERR_ARG;
  if (B == 0) return BA_OK;
  k_se3<OP><<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, Y, out, B);
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

}  // namespace ba

using namespace ba;

extern "C" int se3_expm(const float *a, float *X, int64_t B, void *s) { return launch_se3<OP_EXP>(a, nullptr, X, B, s); }
extern "C" int se3_logm(const float *X, float *a, int64_t B, void *s) { return launch_se3<OP_LOG>(X, nullptr, a, B, s); }
extern "C" int se3_inv(const float *X, float *Y, int64_t B, void *s) { return launch_se3<OP_INV>(X, nullptr, Y, B, s); }
extern "C" int se3_mul(const float *X, const float *Y, float *Z, int64_t B, void *s) { return (B > 0 && !Y) ? BA_ERR_ARG : launch_se3<OP_MUL>(X, Y, Z, B, s); }
extern "C" int se3_adj(const float *X, const float *a, float *b, int64_t B, void *s) { return (B > 0 && !a) ? BA_ERR_ARG : launch_se3<OP_ADJ>(X, a, b, B, s); }
extern "C" int se3_adjT(const float *X, const float *a, float *b, int64_t B, void *s) { return (B > 0 && !a) ? BA_ERR_ARG : launch_se3<OP_ADJT>(X, a, b, B, s); }
extern "C" int se3_act(const float *X, const float *p, float *q, int64_t B, void *s) { return (B > 0 && !p) ? BA_ERR_ARG : launch_se3<OP_ACT>(X, p, q, B, s); }
extern "C" int se3_act4(const float *X, const float *p, float *q, int64_t B, void *s) { return (B > 0 && !p) ? BA_ERR_ARG : launch_se3<OP_ACT4>(X, p, q, B, s); }
extern "C" int se3_as_matrix(const float *X, float *T, int64_t B, void *s) { return launch_se3<OP_MAT>(X, nullptr, T, B, s); }

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
};

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

extern "C" int ba_step_host_async(BaPlan *pl, const BaProblem *ph, void *stream_) {
  if (!pl || !ph || !ph->poses || !ph->patches || !ph->intrinsics || !ph->targets || !ph->weights ||
      !ph->poses_out || !ph->patches_out)
    return BA_ERR_ARG;
  if (ph->targets_stride != 0 && ph->targets_stride != 2) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  const size_t N = pl->v.N, NM = pl->v.NM, E = (size_t)pl->v.E, m = pl->v.m;
  const size_t f = sizeof(float);
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t o_pose = 0, o_pat = o_pose + up(7 * N * f), o_mono = o_pat + up(3 * NM * f),
               o_intr = o_mono + up(NM * f), o_tg = o_intr + up(4 * N * f), o_w = o_tg + up(2 * E * f),
               o_lam = o_w + up(2 * E * f), o_pout = o_lam + up(m * f), o_qout = o_pout + up(7 * N * f),
               total = o_qout + up(3 * NM * f);
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
  if (hp->bytes < total) return BA_ERR_ARG;            // the plan fixes N, NM, E, m: cannot happen
  const int slot = (int)(hp->seq & 1);
  char *d = hp->stage[slot];
  if (hp->seq >= 2) BA_CUDA(cudaStreamWaitEvent(hp->h2d, hp->computed[slot], 0));   // inputs of call seq-2 consumed
  BaProblem pd = *ph;
#define H2D(field, off, bytes)                                                                          \
  do { BA_CUDA(cudaMemcpyAsync(d + (off), ph->field, (bytes), cudaMemcpyHostToDevice, hp->h2d));        \
       pd.field = (const float *)(d + (off)); } while (0)
  H2D(poses, o_pose, 7 * N * f);
  H2D(patches, o_pat, 3 * NM * f);
  if (ph->monodisp) H2D(monodisp, o_mono, NM * f);
  H2D(intrinsics, o_intr, 4 * N * f);
  H2D(targets, o_tg, 2 * E * f);
  H2D(weights, o_w, 2 * E * f);
  if (ph->lmbda_vec) H2D(lmbda_vec, o_lam, m * f);
#undef H2D
  pd.targets_stride = 2;
  BA_CUDA(cudaEventRecord(hp->in_ready[slot], hp->h2d));
  BA_CUDA(cudaStreamWaitEvent(s, hp->in_ready[slot], 0));
  if (hp->seq >= 2) BA_CUDA(cudaStreamWaitEvent(s, hp->out_done[slot], 0));          // results of call seq-2 downloaded
  pd.poses_out = (float *)(d + o_pout);
  pd.patches_out = (float *)(d + o_qout);
  int rc = ba_step(pl, &pd, stream_);
  if (rc) return rc;
  BA_CUDA(cudaEventRecord(hp->computed[slot], s));
  BA_CUDA(cudaStreamWaitEvent(hp->d2h, hp->computed[slot], 0));
  BA_CUDA(cudaMemcpyAsync(ph->poses_out, pd.poses_out, 7 * N * f, cudaMemcpyDeviceToHost, hp->d2h));
  BA_CUDA(cudaMemcpyAsync(ph->patches_out, pd.patches_out, 3 * NM * f, cudaMemcpyDeviceToHost, hp->d2h));
  BA_CUDA(cudaEventRecord(hp->out_done[slot], hp->d2h));
  hp->seq++;
  return BA_OK;
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
