// ba_graph.cu — the factor graph itself, kept on the device: edge list (ii, jj, kk) and the per-edge
// payload the caller carries next to it (targets_3d [E,3], weights [E,2], weights_pose [E,2]) in fixed capacity
// buffers, the live edge count in device memory. Replaces the torch.cat / boolean-mask bookkeeping of
//   main/batrack.py:189-204  append_factors      (ii = ix[patch], jj = frame, kk = patch appended)
//   main/batrack.py:206-212  remove_factors(m)   (stable removal of edges and their payload)
//   main/batrack.py:1023-1026, 1072-1073         removal window:  ix[kk] < first kept frame
//   main/batrack.py:1042-1051  keyframe()        drop edges of frame k, shift the indices behind it
// whose dynamic shapes force a host synchronisation per operation. Here every operation is a few launches on the
// caller's stream; the buffers never move (so a plan update on them replays one captured graph, ba_plan.cu) and the host
// only tracks an upper bound of the edge count until it asks for the real one.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>

#include "ba_internal.h"

struct BaGraph {
  int64_t cap;                 // edges
  int64_t n_upper;             // host-side upper bound of the live count
  int device;
  int64_t *ii, *jj, *kk, *t_ii, *t_jj, *t_kk;
  float *tgt, *w, *wp, *t_f;   // [cap,3], [cap,2], [cap,2], temp [cap,7]
  int *flag, *pos, *count;     // keep flags, their exclusive scan, live count
  char *cub_tmp;
  size_t cub_bytes;
};

namespace ba {

namespace {

__global__ void k_graph_append(int64_t *__restrict__ ii, int64_t *__restrict__ jj, int64_t *__restrict__ kk, float *__restrict__ tgt,
                               float *__restrict__ w, float *__restrict__ wp, int *__restrict__ count, int64_t cap,
                               const int64_t *__restrict__ patch, const int64_t *__restrict__ frame, const int64_t *__restrict__ ix,
                               const float *__restrict__ tgt_new, const float *__restrict__ w_new, const float *__restrict__ wp_new, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int base = *count;
  if (e < n && base + e < cap) {
    const int64_t k = patch[e];
    const int64_t d = base + e;
    ii[d] = ix ? ix[k] : 0; jj[d] = frame[e]; kk[d] = k;                    // batrack.py:196-198
    for (int c = 0; c < 3; ++c) tgt[3 * d + c] = tgt_new ? tgt_new[3 * e + c] : 0.f;
    for (int c = 0; c < 2; ++c) { w[2 * d + c] = w_new ? w_new[2 * e + c] : 0.f; wp[2 * d + c] = wp_new ? wp_new[2 * e + c] : 0.f; }
  }
}
__global__ void k_graph_bump(int *count, int n, int64_t cap) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *count = (int)min((int64_t)*count + n, cap);
}

// mode 0: mask[e] != 0 removes; 1: ix[kk[e]] < a removes (removal window); 2: ii[e] == a || jj[e] == a removes (keyframe)
__global__ void k_graph_flags(const int64_t *__restrict__ ii, const int64_t *__restrict__ jj, const int64_t *__restrict__ kk,
                              const int *__restrict__ count, int n_upper, int mode, int64_t a, const unsigned char *__restrict__ mask,
                              const int64_t *__restrict__ ix, int *__restrict__ flag) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_upper) return;
  int keep = 0;
  if (e < *count) {
    if (mode == 0) keep = mask[e] ? 0 : 1;
    else if (mode == 1) keep = ix[kk[e]] < a ? 0 : 1;
    else keep = (ii[e] == a || jj[e] == a) ? 0 : 1;
  }
  flag[e] = keep;
}
// survivors -> temp at their compacted position; mode 2 also shifts the indices behind the removed frame
// (batrack.py:1049-1051: kk[ii > k] -= M; ii[ii > k] -= 1; jj[jj > k] -= 1)
__global__ void k_graph_gather(const int64_t *__restrict__ ii, const int64_t *__restrict__ jj, const int64_t *__restrict__ kk,
                               const float *__restrict__ tgt, const float *__restrict__ w, const float *__restrict__ wp,
                               const int *__restrict__ flag, const int *__restrict__ pos, int n_upper, int mode, int64_t a, int64_t M,
                               int64_t *__restrict__ t_ii, int64_t *__restrict__ t_jj, int64_t *__restrict__ t_kk, float *__restrict__ t_f) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_upper || !flag[e]) return;
  const int d = pos[e];
  int64_t i = ii[e], j = jj[e], k = kk[e];
  if (mode == 2) {
    if (i > a) { k -= M; i -= 1; }
    if (j > a) j -= 1;
  }
  t_ii[d] = i; t_jj[d] = j; t_kk[d] = k;
  float *f = t_f + 7 * (size_t)d;
  f[0] = tgt[3 * (size_t)e]; f[1] = tgt[3 * (size_t)e + 1]; f[2] = tgt[3 * (size_t)e + 2];
  f[3] = w[2 * (size_t)e]; f[4] = w[2 * (size_t)e + 1]; f[5] = wp[2 * (size_t)e]; f[6] = wp[2 * (size_t)e + 1];
}
__global__ void k_graph_scatter(int64_t *__restrict__ ii, int64_t *__restrict__ jj, int64_t *__restrict__ kk, float *__restrict__ tgt,
                                float *__restrict__ w, float *__restrict__ wp, const int *__restrict__ flag, const int *__restrict__ pos,
                                int n_upper, int *__restrict__ count, const int64_t *__restrict__ t_ii, const int64_t *__restrict__ t_jj,
                                const int64_t *__restrict__ t_kk, const float *__restrict__ t_f) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_upper <= 0) return;
  const int live = pos[n_upper - 1] + flag[n_upper - 1];
  if (d < live) {
    ii[d] = t_ii[d]; jj[d] = t_jj[d]; kk[d] = t_kk[d];
    const float *f = t_f + 7 * (size_t)d;
    tgt[3 * (size_t)d] = f[0]; tgt[3 * (size_t)d + 1] = f[1]; tgt[3 * (size_t)d + 2] = f[2];
    w[2 * (size_t)d] = f[3]; w[2 * (size_t)d + 1] = f[4]; wp[2 * (size_t)d] = f[5]; wp[2 * (size_t)d + 1] = f[6];
  }
}
__global__ void k_graph_set_count(const int *__restrict__ flag, const int *__restrict__ pos, int n_upper, int *__restrict__ count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *count = n_upper > 0 ? pos[n_upper - 1] + flag[n_upper - 1] : 0;
}

template <typename T> cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

}  // namespace

}  // namespace ba

using namespace ba;

extern "C" int ba_graph_create(int64_t cap_edges, void *stream_, BaGraph **out) {
  if (!out || cap_edges <= 0 || cap_edges >= (int64_t)1 << 31) return BA_ERR_ARG;
  *out = nullptr;
  BaGraph *g = new BaGraph();
  std::memset(g, 0, sizeof(*g));
  g->cap = cap_edges;
  BA_CUDA(cudaGetDevice(&g->device));
  const size_t c = (size_t)cap_edges;
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A(dalloc(&g->ii, c)); A(dalloc(&g->jj, c)); A(dalloc(&g->kk, c)); A(dalloc(&g->t_ii, c)); A(dalloc(&g->t_jj, c)); A(dalloc(&g->t_kk, c));
  A(dalloc(&g->tgt, 3 * c)); A(dalloc(&g->w, 2 * c)); A(dalloc(&g->wp, 2 * c)); A(dalloc(&g->t_f, 7 * c));
  A(dalloc(&g->flag, c)); A(dalloc(&g->pos, c)); A(dalloc(&g->count, 1));
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, g->cub_bytes, g->flag, g->pos, (int)cap_edges, (cudaStream_t)0);
  A(dalloc(&g->cub_tmp, g->cub_bytes));
  if (e == cudaSuccess) e = cudaMemsetAsync(g->count, 0, sizeof(int), (cudaStream_t)stream_);
  if (e != cudaSuccess) { set_cuda_error(e, "ba_graph_create"); ba_graph_destroy(g); return BA_ERR_CUDA; }
  *out = g;
  return BA_OK;
}

extern "C" void ba_graph_destroy(BaGraph *g) {
  if (!g) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != g->device) cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  void *ps[] = {g->ii, g->jj, g->kk, g->t_ii, g->t_jj, g->t_kk, g->tgt, g->w, g->wp, g->t_f, g->flag, g->pos, g->count, g->cub_tmp};
  for (void *p : ps) if (p) cudaFree(p);
  if (cur != g->device) cudaSetDevice(cur);
  delete g;
}

extern "C" int ba_graph_arrays(const BaGraph *g, int64_t **ii, int64_t **jj, int64_t **kk, float **targets_3d, float **weights,
                               float **weights_pose, int32_t **n_edges_dev, int64_t *n_edges_upper, int64_t *capacity) {
  if (!g) return BA_ERR_ARG;
  if (ii) *ii = g->ii;
  if (jj) *jj = g->jj;
  if (kk) *kk = g->kk;
  if (targets_3d) *targets_3d = g->tgt;
  if (weights) *weights = g->w;
  if (weights_pose) *weights_pose = g->wp;
  if (n_edges_dev) *n_edges_dev = g->count;
  if (n_edges_upper) *n_edges_upper = g->n_upper;
  if (capacity) *capacity = g->cap;
  return BA_OK;
}

extern "C" int ba_graph_append(BaGraph *g, const int64_t *patch, const int64_t *frame, int64_t n, const int64_t *ix,
                               const float *targets_3d, const float *weights, const float *weights_pose, void *stream_) {
  if (!g || n < 0 || (n > 0 && (!patch || !frame))) return BA_ERR_ARG;
  if (g->n_upper + n > g->cap) return BA_ERR_CAPACITY;
  if (n == 0) return BA_OK;
  cudaStream_t s = (cudaStream_t)stream_;
  k_graph_append<<<(int)((n + 255) / 256), 256, 0, s>>>(g->ii, g->jj, g->kk, g->tgt, g->w, g->wp, g->count, g->cap, patch, frame, ix,
                                                         targets_3d, weights, weights_pose, (int)n);
  BA_LAUNCH_CHECK();
  k_graph_bump<<<1, 1, 0, s>>>(g->count, (int)n, g->cap);
  BA_LAUNCH_CHECK();
  g->n_upper += n;
  return BA_OK;
}

extern "C" int ba_graph_remove(BaGraph *g, int32_t mode, int64_t a, int64_t patches_per_frame, const uint8_t *mask, const int64_t *ix,
                               void *stream_) {
  if (!g || mode < 0 || mode > 2 || (mode == 0 && !mask) || (mode == 1 && !ix)) return BA_ERR_ARG;
  const int nu = (int)g->n_upper;
  if (nu == 0) return BA_OK;
  cudaStream_t s = (cudaStream_t)stream_;
  const int nb = (nu + 255) / 256;
  k_graph_flags<<<nb, 256, 0, s>>>(g->ii, g->jj, g->kk, g->count, nu, mode, a, mask, ix, g->flag);
  BA_LAUNCH_CHECK();
  size_t bytes = g->cub_bytes;
  BA_CUDA(cub::DeviceScan::ExclusiveSum(g->cub_tmp, bytes, g->flag, g->pos, nu, s));
  g_launches.fetch_add(2);
  k_graph_gather<<<nb, 256, 0, s>>>(g->ii, g->jj, g->kk, g->tgt, g->w, g->wp, g->flag, g->pos, nu, mode, a, patches_per_frame,
                                    g->t_ii, g->t_jj, g->t_kk, g->t_f);
  BA_LAUNCH_CHECK();
  k_graph_scatter<<<nb, 256, 0, s>>>(g->ii, g->jj, g->kk, g->tgt, g->w, g->wp, g->flag, g->pos, nu, g->count, g->t_ii, g->t_jj, g->t_kk, g->t_f);
  BA_LAUNCH_CHECK();
  k_graph_set_count<<<1, 1, 0, s>>>(g->flag, g->pos, nu, g->count);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// The one place the host learns the live count: synchronises `stream` and tightens the upper bound.
extern "C" int ba_graph_count(BaGraph *g, int64_t *n_edges, void *stream_) {
  if (!g || !n_edges) return BA_ERR_ARG;
  int c = 0;
  BA_CUDA(cudaMemcpyAsync(&c, g->count, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  BA_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  g->n_upper = c;
  *n_edges = c;
  return BA_OK;
}

// The host learned the live count some other way (the shape block of a finalized plan update carries it): tighten the
// upper bound without a synchronisation of its own.
extern "C" int ba_graph_tighten(BaGraph *g, int64_t n_edges) {
  if (!g || n_edges < 0 || n_edges > g->n_upper) return BA_ERR_ARG;
  g->n_upper = n_edges;
  return BA_OK;
}
