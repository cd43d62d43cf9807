// ba_plan.cu — topology plan: everything about a factor graph that depends only on (ii, jj, kk).
//
// Replaces the per-call index work of the reference (main/backend/ba.py:219 ii.max()/jj.max(),
// :269-277 index shift + torch.unique(kk, return_inverse=True, sorted=True)) by a cached,
// device-built structure:
//   * edges grouped by track, stable            -> eperm, tptr, kx   (kx == unique(kk))
//   * "pattern groups": runs of consecutive tracks whose edge lists have identical (ii, jj)
//     sequences. In SLAM graphs every patch of a keyframe is connected to the same frames
//     (main/batrack.py:399-410 flatmeshgrid; removals are per frame, :1023-1026,1042-1047), so a
//     group is "all patches of one keyframe". Within a group the relative pose Gij, its adjoint and
//     the intrinsics are per-position constants, and index traffic disappears from the edge pass.
//   * per group: the sorted list of distinct poses it touches ("slots"), so the group's E rows and
//     its Schur contribution are small dense objects.
// The caller's ii/jj/kk are only read (edge indices stay bit-exact).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "ba_internal.h"

namespace ba {

std::atomic<long long> g_launches{0};
static std::mutex g_err_mu;
static std::string g_last_cuda_error;

int set_cuda_error(cudaError_t e, const char *what) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_cuda_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  (void)cudaGetLastError();
  return BA_ERR_CUDA;
}

// meta block written by the prep kernels
enum { META_ERR = 0, META_MAXPOSE, META_NOTIDENT, META_DMAX, META_WMAX, META_SPAN, META_NIRREG, META_DMAX_IRREG, META_MINUNIT, META_MINOUNIT, META_COUNT = 10 };

__global__ void k_prep_keys(const int64_t *__restrict__ ii, const int64_t *__restrict__ jj,
                            const int64_t *__restrict__ kk, int64_t E, int N, int NM,
                            unsigned *__restrict__ key, int *__restrict__ val, unsigned *__restrict__ eij,
                            int *__restrict__ meta) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t i = ii[e], j = jj[e], k = kk[e];
  bool bad = i < 0 || i >= N || j < 0 || j >= N || k < 0 || k >= NM;
  if (bad) { atomicOr(&meta[META_ERR], 1); i = j = k = 0; }
  key[e] = (unsigned)k;
  val[e] = (int)e;
  eij[e] = (unsigned)i | ((unsigned)j << 16);
  int mx = (int)(i > j ? i : j);
  // one atomic per warp
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&meta[META_MAXPOSE], mx);
}

__global__ void k_track_flags(const unsigned *__restrict__ skey, const int *__restrict__ eperm,
                              const unsigned *__restrict__ eij, int E, int *__restrict__ tflag,
                              unsigned *__restrict__ sij, int *__restrict__ meta) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= E) return;
  tflag[q] = (q == 0 || skey[q] != skey[q - 1]) ? 1 : 0;
  int e = eperm[q];
  sij[q] = eij[e];
  if (e != q) meta[META_NOTIDENT] = 1;
}

__global__ void k_fill_tracks(const int *__restrict__ tflag, const int *__restrict__ tinc,
                              const unsigned *__restrict__ skey, int E, int *__restrict__ kx,
                              int *__restrict__ tptr) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= E) return;
  if (tflag[q]) {
    int t = tinc[q] - 1;
    kx[t] = (int)skey[q];
    tptr[t] = q;
  }
  if (q == E - 1) tptr[tinc[q]] = E;
}

// A track starts a new pattern group unless its edge list repeats the previous track's (ii,jj) list.
__global__ void k_group_flags(const int *__restrict__ tinc, const int *__restrict__ tptr,
                              const unsigned *__restrict__ sij, int E, int *__restrict__ gflag,
                              int *__restrict__ meta) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= E) return;
  int t = tinc[q] - 1;
  int pos = q - tptr[t];
  int dcur = tptr[t + 1] - tptr[t];
  if (pos == 0) atomicMax(&meta[META_DMAX], dcur);
  if (t == 0) { if (pos == 0) gflag[0] = 1; return; }
  int dprev = tptr[t] - tptr[t - 1];
  if (dcur != dprev) { if (pos == 0) gflag[t] = 1; return; }
  if (sij[q] != sij[tptr[t - 1] + pos]) gflag[t] = 1;
}

__global__ void k_fill_groups(const int *__restrict__ gflag, const int *__restrict__ ginc, int m,
                              int *__restrict__ g_t0, int *__restrict__ t_grp) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  int g = ginc[t] - 1;
  t_grp[t] = g;
  if (gflag[t]) g_t0[g] = t;
  if (t == m - 1) g_t0[g + 1] = m;
}

// Work units = a group's tracks split into the fewest equal pieces of at most `tc` consecutive tracks.
// Groups g_lo <= g < g_hi use units of at most tc_mid tracks instead (streaming Schur units: small at both ends of
// the pose range, where the solver starts, large in the middle).
__global__ void k_chunk_flags(const int *__restrict__ t_grp, const int *__restrict__ g_t0, int m, int tc,
                              int *__restrict__ cflag, int *__restrict__ min_len = nullptr, int g_lo = 0, int g_hi = 0, int tc_mid = 0) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int g = t_grp[t];
  if (g >= g_lo && g < g_hi) tc = tc_mid;
  const int T = g_t0[g + 1] - g_t0[g];
  const int pieces = (T + tc - 1) / tc;
  const int len = ((T + pieces - 1) / pieces + 3) & ~3;          // multiple of 4: 16-byte aligned starts in the E rows
  const int rel = t - g_t0[g];
  cflag[t] = (rel % len == 0) ? 1 : 0;
  if (min_len && rel % len == 0) atomicMin(min_len, min(len, T - rel));     // shortest unit (the last piece of a group)
}
__global__ void k_fill_chunks(const int *__restrict__ cflag, const int *__restrict__ cinc,
                              const int *__restrict__ t_grp, int m, int *__restrict__ c_t0,
                              int *__restrict__ c_grp) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  if (cflag[t]) { int c = cinc[t] - 1; c_t0[c] = t; c_grp[c] = t_grp[t]; }
  if (t == m - 1) c_t0[cinc[t]] = m;
}

__global__ void k_group_degree(const int *__restrict__ g_t0, const int *__restrict__ tptr, int G,
                               int *__restrict__ g_d) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  g_d[g] = g < G ? tptr[g_t0[g] + 1] - tptr[g_t0[g]] : 0;
}

// One CTA per group: slots (distinct poses, ascending), local slot of every pattern position, and
// the per-slot item lists used by the edge pass to reduce per-edge 6-vectors into E rows.
__global__ void k_group_slots(const int *__restrict__ g_t0, const int *__restrict__ g_pat,
                              const int *__restrict__ tptr, const unsigned *__restrict__ sij, int N,
                              int *__restrict__ pat_i, int *__restrict__ pat_j, int *__restrict__ pat_li,
                              int *__restrict__ pat_lj, int *__restrict__ slot_pose,
                              int *__restrict__ slot_ptr, int *__restrict__ slot_items,
                              int *__restrict__ g_nm, int *__restrict__ ms_ptr, int *__restrict__ ms_slot,
                              int *__restrict__ pat_ri, int *__restrict__ pat_rj,
                              int *__restrict__ g_W, long long *__restrict__ g_esz, int *__restrict__ g_reg,
                              int *__restrict__ pat_ps, int *__restrict__ meta) {
  extern __shared__ unsigned sm[];
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nwords = (N + 31) >> 5;
  unsigned *bits = sm;                    // [nwords]
  int *pre = (int *)(sm + nwords);        // [nwords + 1] exclusive popcount prefix
  const int pat0 = g_pat[g], d = g_pat[g + 1] - pat0;
  const int q0 = tptr[g_t0[g]];
  const int T = g_t0[g + 1] - g_t0[g];

  for (int w = tid; w < nwords; w += nt) bits[w] = 0u;
  __syncthreads();
  for (int p = tid; p < d; p += nt) {
    unsigned ij = sij[q0 + p];
    int i = ij & 0xffff, j = ij >> 16;
    pat_i[pat0 + p] = i;
    pat_j[pat0 + p] = j;
    atomicOr(&bits[i >> 5], 1u << (i & 31));
    atomicOr(&bits[j >> 5], 1u << (j & 31));
  }
  __syncthreads();
  if (tid == 0) {                         // nwords <= 2048: a serial prefix is fine for a cached plan
    int s = 0;
    for (int w = 0; w < nwords; ++w) { pre[w] = s; s += __popc(bits[w]); }
    pre[nwords] = s;
  }
  __syncthreads();
  const int W = pre[nwords];
  const int sbase = 2 * pat0;
  for (int w = tid; w < nwords; w += nt) {
    unsigned b = bits[w];
    int s = pre[w];
    while (b) { int bit = __ffs(b) - 1; b &= b - 1; slot_pose[sbase + s++] = (w << 5) + bit; }
  }
  auto slot_of = [&](int pose) { return pre[pose >> 5] + __popc(bits[pose >> 5] & ((1u << (pose & 31)) - 1u)); };
  for (int p = tid; p < d; p += nt) {
    unsigned ij = sij[q0 + p];
    pat_li[pat0 + p] = slot_of(ij & 0xffff);
    pat_lj[pat0 + p] = slot_of(ij >> 16);
  }
  __syncthreads();
  if (tid == 0) {
    // counting sort of the 2d items by slot, ascending item id inside a slot (deterministic sums)
    int *sp = slot_ptr + sbase + g;       // [W+1]
    for (int s = 0; s <= W; ++s) sp[s] = 0;
    for (int p = 0; p < d; ++p) { sp[pat_li[pat0 + p] + 1]++; sp[pat_lj[pat0 + p] + 1]++; }
    for (int s = 0; s < W; ++s) sp[s + 1] += sp[s];
    int *it = slot_items + sbase;
    // fill using a moving cursor kept in `it` tail-free fashion: second pass with per-slot counters
    // stored temporarily in pre[] (no longer needed by other threads after the barrier above)
    for (int s = 0; s < W && s <= nwords; ++s) pre[s] = 0;
    if (W <= nwords + 1) {
      for (int p = 0; p < d; ++p) {
        int a = pat_li[pat0 + p]; it[sp[a] + pre[a]++] = 2 * p;
        int b = pat_lj[pat0 + p]; it[sp[b] + pre[b]++] = 2 * p + 1;
      }
    } else {                              // more slots than bitmap words: quadratic but tiny
      for (int s = 0; s < W; ++s) {
        int c = sp[s];
        for (int p = 0; p < d; ++p) {
          if (pat_li[pat0 + p] == s) it[c++] = 2 * p;
          if (pat_lj[pat0 + p] == s) it[c++] = 2 * p + 1;
        }
      }
    }
    // multi slots (>= 2 items) and the rank of their items in slot order
    int *mp = ms_ptr + sbase + g, *msl = ms_slot + sbase;
    int nm = 0, R = 0;
    for (int p = 0; p < d; ++p) { pat_ri[pat0 + p] = -1; pat_rj[pat0 + p] = -1; }
    for (int s = 0; s < W; ++s) {
      if (sp[s + 1] - sp[s] < 2) continue;
      mp[nm] = R;
      msl[nm++] = s;
      for (int x = sp[s]; x < sp[s + 1]; ++x) {
        const int item = it[x];
        if (item & 1) pat_rj[pat0 + (item >> 1)] = R++; else pat_ri[pat0 + (item >> 1)] = R++;
      }
    }
    mp[nm] = R;
    g_nm[g] = nm;
    g_W[g] = W;
    // E rows are stored entry-major ("SoA"): [6 W][Ts] floats, Ts = T rounded up to 4 (16-byte rows)
    g_esz[g] = (long long)((T + 3) & ~3) * 6 * W;
    // positions ordered by (target slot, position): the walk order of the lane-per-track edge pass, in which the
    // positions that feed one E slot are consecutive (a SLAM graph observes a (patch, frame) pair several times)
    int maxrun = 0;
    {
      int xo = 0;
      for (int s = 0; s < W; ++s) {
        int run = 0;
        for (int x = sp[s]; x < sp[s + 1]; ++x)
          if (it[x] & 1) { pat_ps[pat0 + xo++] = it[x] >> 1; ++run; }
        maxrun = max(maxrun, run);
      }
    }
    // "regular" group (every SLAM graph: one source frame per track): all positions share the source slot, and no
    // target slot is fed more than kMaxSlotRun times -> the lane-per-track edge pass (k_edge_pass_v2) applies
    bool reg = maxrun <= kMaxSlotRun;
    const int s0 = pat_li[pat0];
    for (int p = 0; p < d; ++p) reg = reg && pat_li[pat0 + p] == s0;
    g_reg[g] = reg ? 1 : 0;
    if (!reg) { atomicAdd(&meta[META_NIRREG], 1); atomicMax(&meta[META_DMAX_IRREG], d); }
    atomicMax(&meta[META_WMAX], W);
    int lo = slot_pose[sbase], hi = slot_pose[sbase + W - 1];
    atomicMax(&meta[META_SPAN], hi - lo);
  }
}

__global__ void k_zero_last(long long *p, int idx) { p[idx] = 0; }

// Streaming Schur -> solve hand-over. CTA k of the Schur kernel runs unit o_order[k]: units from both ends of the
// track range alternate, so the rows the two elimination fronts of the solver need first are final first.
__global__ void k_unit_order(int n, int *__restrict__ order) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) order[k] = (k & 1) ? n - 1 - (k >> 1) : (k >> 1);
}
// maxorder[q] = 1 + the last CTA (in launch order) whose unit touches pose q
__global__ void k_unit_reach(const int *__restrict__ order, const int *__restrict__ o_grp, const int *__restrict__ g_pat,
                             const int *__restrict__ g_W, const int *__restrict__ slot_pose, int n, int *__restrict__ maxorder) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int g = o_grp[order[k]];
  const int *sp = slot_pose + 2 * g_pat[g];
  for (int s = 0; s < g_W[g]; ++s) atomicMax(&maxorder[sp[s]], k + 1);
}
// top_need[p] = max over q <= p, bot_need[p] = max over q >= p (one thread: N <= 65535, once per plan)
__global__ void k_need_prefix(const int *__restrict__ maxorder, int N, int *__restrict__ top_need, int *__restrict__ bot_need) {
  if (blockIdx.x || threadIdx.x) return;
  int m = 0;
  for (int p = 0; p < N; ++p) { m = max(m, maxorder[p]); top_need[p] = m; }
  m = 0;
  for (int p = N - 1; p >= 0; --p) { m = max(m, maxorder[p]); bot_need[p] = m; }
}

// inverse of kx: compact track of a patch, -1 for patches without edges (back-substitution runs per patch)
__global__ void k_patch_track(const int *__restrict__ kx, int m, int *__restrict__ patch_track) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) patch_track[kx[t]] = t;
}

__global__ void k_chunk_desc(PlanView v, ChunkDesc *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.n_chunks) return;
  ChunkDesc d;
  d.g = v.c_grp[c]; d.t0 = v.c_t0[c]; d.t1 = v.c_t0[c + 1]; d.gt0 = v.g_t0[d.g];
  d.pat0 = v.g_pat[d.g]; d.d = v.g_pat[d.g + 1] - d.pat0; d.W = v.g_W[d.g]; d.ebase = v.tptr[d.gt0];
  d.nm = v.g_nm[d.g]; d.R = v.ms_ptr[2 * d.pat0 + d.g + d.nm]; d.Ts = (v.g_t0[d.g + 1] - d.gt0 + 3) & ~3; d.reg = v.g_reg[d.g];
  d.eoff = v.g_eoff[d.g]; d.pad2 = 0;
  out[c] = d;
}

// ---- small RAII helpers (host) ---------------------------------------------------------------
struct Scratch {
  cudaStream_t s;
  std::vector<void *> blocks;
  explicit Scratch(cudaStream_t st) : s(st) {}
  ~Scratch() { for (void *b : blocks) cudaFreeAsync(b, s); }
  template <typename T> cudaError_t get(T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(n, 1) * sizeof(T), s);
    if (e == cudaSuccess) blocks.push_back(q);
    *p = (T *)q;
    return e;
  }
};

// Plan-owned memory comes from the device's stream-ordered pool with the release threshold lifted: a SLAM
// front end rebuilds the plan whenever the graph changes (every frame), and after the first few plans every
// allocation is a pool hit instead of a cudaMalloc (~40 of them per plan; measured in tools/plan_build_time.py).
// All process-wide state is per device: the stream plan memory is released on, and the "attributes set / tables
// built" flag (function attributes and allocations belong to one device).
constexpr int kMaxDevices = 64;
static cudaStream_t g_mem_stream[kMaxDevices] = {nullptr};
static bool g_dev_ready[kMaxDevices] = {false};
static std::mutex g_dev_mu;
static cudaError_t mem_pool_init(int dev) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (g_mem_stream[dev]) return cudaSuccess;
  cudaError_t e;
  cudaMemPool_t pool;
  if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
  unsigned long long thr = ~0ull;
  if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr)) != cudaSuccess) return e;
  return cudaStreamCreateWithFlags(&g_mem_stream[dev], cudaStreamNonBlocking);
}
// First plan on a device: opt every kernel in to its dynamic shared memory, build the legacy solver's step tables
// (synchronously, so that no later stream races with the build and nothing is allocated under graph capture).
static int prepare_device(int dev, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (dev < 0 || dev >= kMaxDevices) return BA_ERR_ARG;
  if (g_dev_ready[dev]) return BA_OK;
  int rc = kernels_prepare_device();
  if (!rc) rc = solve_diag_prepare_device();
  if (!rc) rc = solve_mma_prepare_device(dev, s);
  if (!rc) rc = solve_tiles_prepare_device();
  if (!rc) rc = schur_tc_prepare_device();
  if (!rc) g_dev_ready[dev] = true;
  return rc;
}
static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
static void options_from_env(BaOptions *o) {
  o->solver = 0;
  if (const char *e = getenv("BA_SOLVER")) {               // "diag" / "mma" / "window" / "dense" or 0..3
    if (e[0] == 'm' || e[0] == '1') o->solver = 1;
    else if (e[0] == 'w' || e[0] == '2') o->solver = 2;
    else if ((e[0] == 'd' && e[1] == 'e') || e[0] == '3') o->solver = 3;
    else if (e[0] == 't' || e[0] == '4') o->solver = 4;
    else if ((e[0] == 'd' && e[1] == 'i') || e[0] == '5') o->solver = 5;
  }
  o->stream = env_int("BA_STREAM", 0) ? 1 : 0;   // off: with the tensor-core Schur kernel the plain sequence is faster
  o->stream_smem_kb = std::max(0, env_int("BA_STREAM_SMEM_KB", 0));
  o->schur_tile = std::max(4, env_int("BA_SCHUR_TILE", 64));
  o->twist_min = std::max(17, env_int("BA_TWIST_MIN", 64));
  o->spin_cap = std::max(0, env_int("BA_SPIN_CAP", 0));
  o->trace = env_int("BA_SOLVER_TRACE", 0) ? 1 : 0;
  o->schur = env_int("BA_SCHUR", 0) ? 1 : 0;
  o->schur_acc = std::min(8, std::max(1, env_int("BA_SCHUR_ACC", 4)));
}
template <typename T> static cudaError_t own(BaPlan *pl, T **p, size_t n) {
  void *q = nullptr;
  cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(n, 1) * sizeof(T), pl->mem_stream);
  if (e == cudaSuccess) pl->owned.push_back(q);
  *p = (T *)q;
  return e;
}

static cudaError_t inclusive_sum(Scratch &sc, const int *in, int *out, int n, cudaStream_t s) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, s);
  if (e != cudaSuccess) return e;
  char *tmp;
  if ((e = sc.get(&tmp, bytes)) != cudaSuccess) return e;
  return cub::DeviceScan::InclusiveSum(tmp, bytes, in, out, n, s);
}
template <typename T> static cudaError_t exclusive_sum(Scratch &sc, const T *in, T *out, int n, cudaStream_t s) {
  size_t bytes = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s);
  if (e != cudaSuccess) return e;
  char *tmp;
  if ((e = sc.get(&tmp, bytes)) != cudaSuccess) return e;
  return cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, s);
}

static inline int cdiv(int64_t a, int b) { return (int)((a + b - 1) / b); }

void layout_for(const BaPlan *p, int fixedp, int *n, int *bw, int *ld, int *off, int64_t *s_floats) {
  int nn = std::max(p->n_total_layout - fixedp, 0);
  int M = 6 * nn;
  int b = std::min(6 * p->bwb_layout + 5, std::max(M - 1, 0));
  const size_t win_bytes = ((size_t)(b + 1) * ((b + 1) | 1) + M + b + 1) * sizeof(double);
  if (M > 0 && b + 1 <= kMaxWindow && win_bytes <= 227 * 1024 - 64) {   // lower band storage: S(r,c) at r*bw + c + bw
    *ld = b; *off = b; *s_floats = (int64_t)M * (b + 1);
  } else {                                // dense lower storage
    b = std::max(M - 1, 0);
    *ld = M; *off = 0; *s_floats = (int64_t)M * M;
  }
  *n = nn; *bw = b;
}

static int alloc_workspace(BaPlan *pl) {
  // capacity for the smallest fixedp (0): the reduced system only shrinks as fixedp grows
  int n, bw, ld, off; int64_t sf;
  layout_for(pl, 0, &n, &bw, &ld, &off, &sf);
  int64_t need = sf + 6 * (int64_t)n + 8 + (pl->v.n_ounits + 1) / 2 + 2;     // [S | y | completion flags of the streaming Schur units]
  if (need > pl->sy_floats) {
    BA_CUDA(own(pl, &pl->SY, need));
    BA_CUDA(own(pl, &pl->L, sf + 8));
    BA_CUDA(own(pl, &pl->dX, 6 * (size_t)n + 8));
    BA_CUDA(own(pl, &pl->Wg, std::max(solve_mma_scratch_doubles(6 * n, std::min(bw, kMmaMaxBw)), solve_tiles_scratch_doubles(6 * n, bw)) + 64));   // solver scratch
    pl->sy_floats = need;
  }
  pl->info.banded = (ld != 6 * n) ? 1 : 0;
  return BA_OK;
}

}  // namespace ba

using namespace ba;

extern "C" int ba_plan_create(const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t E,
                              int32_t N, int32_t NM, void *stream_, BaPlan **out) {
  if (!out) return BA_ERR_ARG;
  *out = nullptr;
  if (!ii || !jj || !kk || E <= 0 || E >= (int64_t)1 << 31 || N <= 0 || NM <= 0) return BA_ERR_ARG;
  if (N > 65535) return BA_ERR_TOO_MANY_POSES;
  cudaStream_t s = (cudaStream_t)stream_;
  int dev = 0;
  BA_CUDA(cudaGetDevice(&dev));

  if (mem_pool_init(dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "memory pool");
  {
    const int rc = prepare_device(dev, s);
    if (rc) return rc;
  }
  BaPlan *pl = new BaPlan();
  options_from_env(&pl->opt);
  pl->trace_buf = nullptr;
  const int want_trace = pl->opt.trace;
  pl->opt.trace = 0;
  pl->mem_stream = s;                      // allocations are ordered on the creation stream (used on it right away)
  pl->solve_stream = nullptr; pl->ev_step_begin = pl->ev_solved = nullptr; pl->epoch = 0; pl->solve_shape_key = -1;
  std::memset(&pl->info, 0, sizeof(pl->info));
  std::memset(&pl->v, 0, sizeof(pl->v));
  pl->device = dev;
  pl->SY = pl->L = pl->dX = pl->Wg = nullptr;
  pl->Est = pl->dZ = nullptr;
  pl->Cw = pl->Qw = nullptr;
  pl->status = nullptr;
  pl->sy_floats = 0;
  pl->last_n = pl->last_fixedp = -1;
  pl->host_pipe = nullptr;
  pl->host_pipe_destroy = nullptr;
  pl->pp_buf[0] = pl->pp_buf[1] = nullptr;
  pl->timing = 0;
  pl->ev_mask = 0;
  for (auto &e : pl->ev) e = nullptr;
  if (want_trace) ba_plan_set_option(pl, BA_OPT_SOLVER_TRACE, 1);

  auto fail = [&](int code) { ba_plan_destroy(pl); return code; };
#define PL_CUDA(call)                                                                      \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_cuda_error(e_, #call); return fail(BA_ERR_CUDA); } } while (0)
#define PL_LAUNCH() do { g_launches.fetch_add(1); PL_CUDA(cudaGetLastError()); } while (0)

  const int TB = 256;
  const int nE = (int)E;
  int hmeta[META_COUNT];
  {
    Scratch sc(s);
    unsigned *key, *skey, *eij, *sij;
    int *val, *meta, *tflag, *tinc;
    PL_CUDA(sc.get(&key, E)); PL_CUDA(sc.get(&skey, E)); PL_CUDA(sc.get(&eij, E)); PL_CUDA(sc.get(&sij, E));
    PL_CUDA(sc.get(&val, E)); PL_CUDA(sc.get(&meta, META_COUNT)); PL_CUDA(sc.get(&tflag, E)); PL_CUDA(sc.get(&tinc, E));
    int *eperm;
    PL_CUDA(own(pl, &eperm, E));
    PL_CUDA(cudaMemsetAsync(meta, 0, META_COUNT * sizeof(int), s));
    PL_CUDA(cudaMemsetAsync(meta + META_MINUNIT, 0x7f, 2 * sizeof(int), s));

    k_prep_keys<<<cdiv(E, TB), TB, 0, s>>>(ii, jj, kk, E, N, NM, key, val, eij, meta); PL_LAUNCH();
    {
      int end_bit = 1;
      while (end_bit < 32 && ((int64_t)1 << end_bit) < (int64_t)NM) ++end_bit;
      size_t bytes = 0;
      PL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key, skey, val, eperm, nE, 0, end_bit, s));
      char *tmp;
      PL_CUDA(sc.get(&tmp, bytes));
      PL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, key, skey, val, eperm, nE, 0, end_bit, s));
      g_launches.fetch_add(3);
    }
    k_track_flags<<<cdiv(E, TB), TB, 0, s>>>(skey, eperm, eij, nE, tflag, sij, meta); PL_LAUNCH();
    PL_CUDA(inclusive_sum(sc, tflag, tinc, nE, s));
    int m = 0;
    PL_CUDA(cudaMemcpyAsync(&m, tinc + (nE - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaMemcpyAsync(hmeta, meta, sizeof(hmeta), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaStreamSynchronize(s));
    if (hmeta[META_ERR]) return fail(BA_ERR_INDEX_RANGE);

    int *kx, *tptr, *t_grp, *gflag, *ginc;
    PL_CUDA(own(pl, &kx, m)); PL_CUDA(own(pl, &tptr, m + 1)); PL_CUDA(own(pl, &t_grp, m));
    PL_CUDA(sc.get(&gflag, m)); PL_CUDA(sc.get(&ginc, m));
    PL_CUDA(cudaMemsetAsync(gflag, 0, (size_t)m * sizeof(int), s));
    k_fill_tracks<<<cdiv(E, TB), TB, 0, s>>>(tflag, tinc, skey, nE, kx, tptr); PL_LAUNCH();
    k_group_flags<<<cdiv(E, TB), TB, 0, s>>>(tinc, tptr, sij, nE, gflag, meta); PL_LAUNCH();
    PL_CUDA(inclusive_sum(sc, gflag, ginc, m, s));
    int G = 0, hmeta_dmax = 0;
    PL_CUDA(cudaMemcpyAsync(&G, ginc + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaMemcpyAsync(&hmeta_dmax, meta + META_DMAX, sizeof(int), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaStreamSynchronize(s));

    int *g_t0, *g_pat, *g_W, *g_d, *g_reg;
    long long *g_eoff, *g_esz;
    PL_CUDA(own(pl, &g_t0, G + 1)); PL_CUDA(own(pl, &g_pat, G + 1)); PL_CUDA(own(pl, &g_W, G));
    PL_CUDA(own(pl, &g_eoff, G + 1)); PL_CUDA(own(pl, &g_reg, G));
    PL_CUDA(sc.get(&g_d, G + 1)); PL_CUDA(sc.get(&g_esz, G + 1));
    k_fill_groups<<<cdiv(m, TB), TB, 0, s>>>(gflag, ginc, m, g_t0, t_grp); PL_LAUNCH();
    k_group_degree<<<cdiv(G + 1, TB), TB, 0, s>>>(g_t0, tptr, G, g_d); PL_LAUNCH();
    PL_CUDA(exclusive_sum(sc, g_d, g_pat, G + 1, s));

    // work units: edge-pass chunks and Schur units (fewer, larger: their flush is (6W)^2 atomics)
    const int sms = 148;
    // edge pass: about one wave of CTAs (2 resident per SM) — per-CTA set-up and flush are amortised over
    // more tracks; Schur: units of <= 128 tracks. BA_EDGE_TC / BA_SCHUR_TU override for experiments.
    int tc = std::min(256, std::max(8, cdiv(m, sms)));
    int tu = std::min(256, std::max(16, 16 * cdiv(cdiv(m, 2 * sms), 16)));
    if (tu > 128) tu = 256;               // whole 256-track groups: measured 4 us faster than two 128-track units at cfg3
    if (const char *e = getenv("BA_EDGE_TC")) tc = std::max(1, atoi(e));
    if (const char *e = getenv("BA_SCHUR_TU")) tu = std::max(1, atoi(e));
    int *cflag, *cinc;
    PL_CUDA(sc.get(&cflag, m)); PL_CUDA(sc.get(&cinc, m));
    // lane-per-track edge pass: a CTA of kEdge2Warps warps = KT track slices x KP position splits. One split is the
    // cheapest per edge (fewest flushes; measured at 256 KF / 64k tracks: 44 us vs 54 / 60 us with 2 / 4 splits);
    // graphs with few tracks split the positions until the machine sees ~12 warps per SM (25-frame window, 5200
    // tracks x <= 72 edges: 70 / 53 / 39 us with 2 / 4 / 8 splits).
    int kp = 1;
    while (kp < kEdge2Warps && (int64_t)cdiv(m, 32) * kp < 12 * sms && cdiv(hmeta_dmax, kp) > 2) kp *= 2;
    if (const char *e = getenv("BA_EDGE2_KP")) { int v2 = atoi(e); if (v2 == 1 || v2 == 2 || v2 == 4 || v2 == 8) kp = v2; }
    const int tx = 32 * (kEdge2Warps / kp);
    int counts[4] = {0, 0, 0, 0};
    int *unit_t0[4], *unit_grp[4];
    int to = 64;                                       // Schur units of the streaming hand-over to the solver
    if (const char *e = getenv("BA_STREAM_TU")) to = std::max(16, atoi(e) & ~3);
    const int unit_len[4] = {tc, tu, tx, to};
    int to_max = to;
    for (int pass = 0; pass < 4; ++pass) {
      if (pass == 3) {                                 // optionally large units in the middle of the pose range (measured at
        int gend = 1 << 30;                            // 256 KF: 24 / 40 / 64 end groups small: 525 / 517 / 492 us, all small 494)
        if (const char *e = getenv("BA_STREAM_GEND")) gend = std::max(0, atoi(e));
        if (gend < G - gend) to_max = std::max(to, 256);
        k_chunk_flags<<<cdiv(m, TB), TB, 0, s>>>(t_grp, g_t0, m, unit_len[pass], cflag, meta + META_MINOUNIT, std::min(gend, G), G - std::min(gend, G), 256);
      } else {
        k_chunk_flags<<<cdiv(m, TB), TB, 0, s>>>(t_grp, g_t0, m, unit_len[pass], cflag, pass == 1 ? meta + META_MINUNIT : nullptr);
      }
      PL_LAUNCH();
      PL_CUDA(inclusive_sum(sc, cflag, cinc, m, s));
      PL_CUDA(cudaMemcpyAsync(&counts[pass], cinc + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
      PL_CUDA(cudaStreamSynchronize(s));
      PL_CUDA(own(pl, &unit_t0[pass], counts[pass] + 1));
      PL_CUDA(own(pl, &unit_grp[pass], counts[pass]));
      k_fill_chunks<<<cdiv(m, TB), TB, 0, s>>>(cflag, cinc, t_grp, m, unit_t0[pass], unit_grp[pass]); PL_LAUNCH();
    }

    int pat_total = 0;
    PL_CUDA(cudaMemcpyAsync(&pat_total, g_pat + G, sizeof(int), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaStreamSynchronize(s));
    int *pat_i, *pat_j, *pat_li, *pat_lj, *slot_pose, *slot_ptr, *slot_items, *g_nm, *ms_ptr, *ms_slot, *pat_ri, *pat_rj, *pat_ps;
    PL_CUDA(own(pl, &g_nm, G)); PL_CUDA(own(pl, &pat_ri, pat_total)); PL_CUDA(own(pl, &pat_rj, pat_total));
    PL_CUDA(own(pl, &ms_slot, 2 * (size_t)pat_total)); PL_CUDA(own(pl, &ms_ptr, 2 * (size_t)pat_total + G + 1));
    PL_CUDA(own(pl, &pat_i, pat_total)); PL_CUDA(own(pl, &pat_j, pat_total)); PL_CUDA(own(pl, &pat_ps, pat_total));
    PL_CUDA(own(pl, &pat_li, pat_total)); PL_CUDA(own(pl, &pat_lj, pat_total));
    PL_CUDA(own(pl, &slot_pose, 2 * (size_t)pat_total)); PL_CUDA(own(pl, &slot_items, 2 * (size_t)pat_total));
    PL_CUDA(own(pl, &slot_ptr, 2 * (size_t)pat_total + G + 1));
    {
      int nwords = (N + 31) / 32;
      size_t smem = (size_t)(2 * nwords + 1) * sizeof(int);
      k_group_slots<<<G, 128, smem, s>>>(g_t0, g_pat, tptr, sij, N, pat_i, pat_j, pat_li, pat_lj, slot_pose,
                                         slot_ptr, slot_items, g_nm, ms_ptr, ms_slot, pat_ri, pat_rj, g_W, g_esz, g_reg, pat_ps, meta); PL_LAUNCH();
    }
    k_zero_last<<<1, 1, 0, s>>>(g_esz, G); PL_LAUNCH();
    PL_CUDA(exclusive_sum(sc, g_esz, g_eoff, G + 1, s));
    long long esize = 0;
    PL_CUDA(cudaMemcpyAsync(&esize, g_eoff + G, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaMemcpyAsync(hmeta, meta, sizeof(hmeta), cudaMemcpyDeviceToHost, s));
    PL_CUDA(cudaStreamSynchronize(s));

    PlanView &v = pl->v;
    v.E = E; v.N = N; v.NM = NM; v.m = m; v.G = G; v.n_chunks = counts[0]; v.n_units = counts[1];
    v.perm_identity = hmeta[META_NOTIDENT] ? 0 : 1;
    v.eperm = eperm; v.kx = kx; v.tptr = tptr; v.t_grp = t_grp; v.g_t0 = g_t0; v.g_pat = g_pat; v.g_W = g_W;
    v.g_eoff = g_eoff; v.pat_i = pat_i; v.pat_j = pat_j; v.pat_li = pat_li; v.pat_lj = pat_lj;
    v.slot_pose = slot_pose; v.slot_ptr = slot_ptr; v.slot_items = slot_items;
    v.g_nm = g_nm; v.ms_ptr = ms_ptr; v.ms_slot = ms_slot; v.pat_ri = pat_ri; v.pat_rj = pat_rj; v.dmax = hmeta[META_DMAX];
    v.c_t0 = unit_t0[0]; v.c_grp = unit_grp[0]; v.u_t0 = unit_t0[1]; v.u_grp = unit_grp[1];
    v.x_t0 = unit_t0[2]; v.x_grp = unit_grp[2]; v.n_xchunks = counts[2]; v.e2_kp = kp;
    {
      int *order, *flag, *maxo, *tneed, *bneed;
      const int no = counts[3];
      PL_CUDA(own(pl, &order, no)); PL_CUDA(own(pl, &flag, no)); PL_CUDA(own(pl, &tneed, N)); PL_CUDA(own(pl, &bneed, N));
      PL_CUDA(sc.get(&maxo, N));
      PL_CUDA(cudaMemsetAsync(flag, 0, (size_t)no * sizeof(int), s));
      PL_CUDA(cudaMemsetAsync(maxo, 0, (size_t)N * sizeof(int), s));
      k_unit_order<<<cdiv(no, TB), TB, 0, s>>>(no, order); PL_LAUNCH();
      k_unit_reach<<<cdiv(no, TB), TB, 0, s>>>(order, unit_grp[3], g_pat, g_W, slot_pose, no, maxo); PL_LAUNCH();
      k_need_prefix<<<1, 32, 0, s>>>(maxo, N, tneed, bneed); PL_LAUNCH();
      v.o_t0 = unit_t0[3]; v.o_grp = unit_grp[3]; v.o_order = order; v.n_ounits = no; v.o_flag = flag;
      v.top_need = tneed; v.bot_need = bneed;
    }
    v.pat_ps = pat_ps; v.g_reg = g_reg; v.n_irregular = hmeta[META_NIRREG]; v.dmax_irregular = hmeta[META_DMAX_IRREG];
    {
      int *ptk;
      PL_CUDA(own(pl, &ptk, NM));
      PL_CUDA(cudaMemsetAsync(ptk, 0xff, (size_t)NM * sizeof(int), s));
      k_patch_track<<<cdiv(m, TB), TB, 0, s>>>(kx, m, ptk); PL_LAUNCH();
      v.patch_track = ptk;
    }

    {
      ChunkDesc *cd;
      PL_CUDA(own(pl, &cd, v.n_chunks));
      k_chunk_desc<<<cdiv(v.n_chunks, TB), TB, 0, s>>>(v, cd); PL_LAUNCH();
      v.cdesc = cd;
    }

    BaPlanInfo &in = pl->info;
    in.n_edges = E; in.n_poses = N; in.n_patches = NM; in.n_total = hmeta[META_MAXPOSE] + 1;
    in.n_tracks = m; in.n_groups = G; in.n_chunks = counts[0]; in.max_degree = hmeta[META_DMAX];
    in.max_slots = hmeta[META_WMAX]; in.block_bandwidth = hmeta[META_SPAN]; in.perm_identity = v.perm_identity;
    pl->n_total_layout = in.n_total;
    pl->bwb_layout = in.block_bandwidth;
    pl->min_unit = hmeta[META_MINUNIT]; pl->min_ounit = hmeta[META_MINOUNIT]; pl->max_unit = tu; pl->max_ounit = to_max;

    PL_CUDA(own(pl, &pl->Est, (size_t)esize + 8));
    PL_CUDA(cudaMemsetAsync(pl->Est, 0, ((size_t)esize + 8) * sizeof(float), s));   // row padding stays zero (finite) for ever
    pl->est_floats = esize;
    PL_CUDA(own(pl, &pl->Cw, m)); PL_CUDA(own(pl, &pl->Qw, m)); PL_CUDA(own(pl, &pl->dZ, m));
    PL_CUDA(own(pl, &pl->status, 4));
    PL_CUDA(cudaMemsetAsync(pl->status, 0, 4 * sizeof(int), s));
    if (alloc_workspace(pl) != BA_OK) return fail(BA_ERR_CUDA);
    in.workspace_bytes = (int64_t)(esize + 8) * 4 + (int64_t)m * 20 + pl->sy_floats * 16;
  }  // Scratch frees (stream-ordered)
  PL_CUDA(cudaStreamSynchronize(s));
#undef PL_CUDA
#undef PL_LAUNCH
  *out = pl;
  return BA_OK;
}

extern "C" void ba_plan_destroy(BaPlan *pl) {
  if (!pl) return;
  // kernels of the last calls may still be running on the caller's streams: wait (what cudaFree did implicitly),
  // then hand the blocks back to the pool
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != pl->device) cudaSetDevice(pl->device);        // the plan's memory, streams and events live on its device
  cudaDeviceSynchronize();
  for (void *p : pl->owned) cudaFreeAsync(p, g_mem_stream[pl->device]);
  if (pl->trace_buf) cudaFree(pl->trace_buf);
  if (pl->solve_stream) cudaStreamDestroy(pl->solve_stream);
  if (pl->ev_step_begin) cudaEventDestroy(pl->ev_step_begin);
  if (pl->ev_solved) cudaEventDestroy(pl->ev_solved);
  if (pl->host_pipe && pl->host_pipe_destroy) pl->host_pipe_destroy(pl->host_pipe);
  for (auto &e : pl->ev) if (e) cudaEventDestroy(e);
  if (cur != pl->device) cudaSetDevice(cur);
  delete pl;
}

constexpr size_t kTraceValues = 16 * 4096 + 32;

extern "C" int ba_plan_set_option(BaPlan *pl, int32_t key, int32_t value) {
  if (!pl) return BA_ERR_ARG;
  switch (key) {
    case BA_OPT_SOLVER: if (value < 0 || value > 5) return BA_ERR_ARG; pl->opt.solver = value; break;
    case BA_OPT_STREAM: pl->opt.stream = value ? 1 : 0; break;
    case BA_OPT_STREAM_SMEM_KB: if (value < 0 || value > 200) return BA_ERR_ARG; pl->opt.stream_smem_kb = value; break;
    case BA_OPT_SCHUR_TILE: if (value < 4) return BA_ERR_ARG; pl->opt.schur_tile = value & ~3; break;
    case BA_OPT_TWIST_MIN: if (value < 17) return BA_ERR_ARG; pl->opt.twist_min = value; break;
    case BA_OPT_SPIN_CAP: if (value < 0) return BA_ERR_ARG; pl->opt.spin_cap = value; break;
    case BA_OPT_SOLVER_TRACE:
      if (value && !pl->trace_buf) {
        BA_CUDA(cudaMalloc(&pl->trace_buf, kTraceValues * sizeof(long long)));
        BA_CUDA(cudaMemset(pl->trace_buf, 0, kTraceValues * sizeof(long long)));
      }
      pl->opt.trace = value;              // bit 0: band solver, bit 1: tensor-core Schur kernel
      break;
    case BA_OPT_SCHUR: if (value < 0 || value > 1) return BA_ERR_ARG; pl->opt.schur = value; break;
    case BA_OPT_SCHUR_ACC: if (value < 1 || value > 8) return BA_ERR_ARG; pl->opt.schur_acc = value; break;
    default: return BA_ERR_ARG;
  }
  return BA_OK;
}

extern "C" int ba_plan_get_option(const BaPlan *pl, int32_t key, int32_t *value) {
  if (!pl || !value) return BA_ERR_ARG;
  switch (key) {
    case BA_OPT_SOLVER: *value = pl->opt.solver; break;
    case BA_OPT_STREAM: *value = pl->opt.stream; break;
    case BA_OPT_STREAM_SMEM_KB: *value = pl->opt.stream_smem_kb; break;
    case BA_OPT_SCHUR_TILE: *value = pl->opt.schur_tile; break;
    case BA_OPT_TWIST_MIN: *value = pl->opt.twist_min; break;
    case BA_OPT_SPIN_CAP: *value = pl->opt.spin_cap; break;
    case BA_OPT_SOLVER_TRACE: *value = pl->opt.trace; break;
    case BA_OPT_SCHUR: *value = pl->opt.schur; break;
    case BA_OPT_SCHUR_ACC: *value = pl->opt.schur_acc; break;
    default: return BA_ERR_ARG;
  }
  return BA_OK;
}

extern "C" int ba_plan_read_trace(const BaPlan *pl, int64_t *out, int64_t n, void *stream) {
  if (!pl || !out || n <= 0 || !pl->trace_buf) return BA_ERR_ARG;
  BA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  BA_CUDA(cudaDeviceSynchronize());
  BA_CUDA(cudaMemcpy(out, pl->trace_buf, (size_t)std::min<int64_t>(n, (int64_t)kTraceValues) * sizeof(long long), cudaMemcpyDeviceToHost));
  return BA_OK;
}

extern "C" int ba_plan_info(const BaPlan *pl, BaPlanInfo *out) {
  if (!pl || !out) return BA_ERR_ARG;
  *out = pl->info;
  return BA_OK;
}

extern "C" int ba_plan_set_layout(BaPlan *pl, int32_t n_total, int32_t bwb) {
  if (!pl || n_total < pl->info.n_total || bwb < pl->info.block_bandwidth || n_total > pl->info.n_poses)
    return BA_ERR_ARG;
  pl->n_total_layout = n_total;
  pl->bwb_layout = bwb;
  return ba::alloc_workspace(pl);
}

extern "C" int ba_plan_enable_timing(BaPlan *pl, int enable) {
  if (!pl) return BA_ERR_ARG;
  if (enable) for (auto &e : pl->ev) if (!e) BA_CUDA(cudaEventCreate(&e));
  pl->timing = enable ? 1 : 0;
  pl->ev_mask = 0;
  return BA_OK;
}

// Stage k ran between boundary events k and k+1; stages that did not run report 0.
extern "C" int ba_plan_last_timing(BaPlan *pl, float *ms) {
  if (!pl || !ms || !pl->timing) return BA_ERR_ARG;
  for (int k = BA_N_STAGES; k >= 0; --k)
    if (pl->ev_mask & (1u << k)) { BA_CUDA(cudaEventSynchronize(pl->ev[k])); break; }
  for (int k = 0; k < BA_N_STAGES; ++k) {
    ms[k] = 0.0f;
    if (!(pl->ev_mask & (1u << k))) continue;
    int nxt = k + 1;
    while (nxt <= BA_N_STAGES && !(pl->ev_mask & (1u << nxt))) ++nxt;
    if (nxt > BA_N_STAGES) continue;
    // a boundary only opens a stage if that stage actually launched something: see ba_kernels.cu
    BA_CUDA(cudaEventElapsedTime(&ms[k], pl->ev[k], pl->ev[nxt]));
  }
  return BA_OK;
}

extern "C" int ba_plan_tracks(const BaPlan *pl, int32_t *kx_out, void *stream) {
  if (!pl || !kx_out) return BA_ERR_ARG;
  BA_CUDA(cudaMemcpyAsync(kx_out, pl->v.kx, (size_t)pl->v.m * sizeof(int), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  return BA_OK;
}

extern "C" const char *ba_error_string(int code) {
  switch (code) {
    case BA_OK: return "ok";
    case BA_ERR_CUDA: return "CUDA runtime error (see ba_last_cuda_error)";
    case BA_ERR_ARG: return "invalid argument";
    case BA_ERR_INDEX_RANGE: return "edge index out of range";
    case BA_ERR_TOO_MANY_POSES: return "pose buffer longer than 65535";
    case BA_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    default: return "unknown error";
  }
}
extern "C" const char *ba_last_cuda_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(ba::g_err_mu);
  copy = ba::g_last_cuda_error;
  return copy.c_str();
}
extern "C" int ba_version(void) { return 100; }
extern "C" int64_t ba_launch_count(void) { return (int64_t)ba::g_launches.load(); }
