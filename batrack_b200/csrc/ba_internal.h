// ba_internal.h — plan layout shared by the translation units of libbatrack_ba.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <atomic>
#include <vector>

#include "../../include/batrack_ba.h"

namespace ba {

constexpr int kEdgeThreads = 256;     // CTA size of the generic edge pass (irregular groups)
constexpr int kMaxSlotRun = 24;       // most positions of a track that may feed one E slot in the lane-per-track edge pass
constexpr int kEdge2Warps = 8;        // warps per CTA of the lane-per-track edge pass: KT track slices x KP position splits
constexpr int kSchurThreads = 256;    // CTA size of the per-track Schur kernel
constexpr int kSolveThreads = 1024;   // CTA size of the window Cholesky
constexpr int kMaxWindow = 150;       // largest band window (bw + 1) the shared-memory solver holds (fp64)

// Everything the edge pass needs to know about one chunk, gathered at plan time so that a CTA starts with
// one 48-byte load instead of a chain of dependent index loads.
struct ChunkDesc {
  int g, t0, t1, gt0;      // group, track range, first track of the group
  int pat0, d, W, ebase;   // pattern offset, degree, slots, first sorted edge of the group
  int nm, R, Ts, reg;      // multi slots, staged items per track, E row stride (tracks, padded to 4), regular group
  long long eoff;          // offset of the group's E rows
  long long pad2;
};

// Device-side view of the cached topology (all pointers device memory owned by the plan).
struct PlanView {
  int64_t E;
  int N, NM, m, G, n_chunks, n_units;
  int perm_identity;
  const int *eperm;        // [E]   original edge id of the q-th edge in track-major (stable) order
  const int *kx;           // [m]   patch index of compact track t (== torch.unique(kk)[t])
  const int *tptr;         // [m+1] first sorted edge of track t
  const int *t_grp;        // [m]   group of track t
  const int *g_t0;         // [G+1] first track of group g
  const int *g_pat;        // [G+1] offset of group g's pattern (its degree d = g_pat[g+1]-g_pat[g])
  const int *g_W;          // [G]   distinct poses ("slots") touched by group g
  const long long *g_eoff; // [G+1] offset (floats) of group g's E block, entry-major: [6 W_g][Ts_g], Ts = T_g rounded up to 4
  const int *pat_i, *pat_j;    // [sum d] raw source / target pose per pattern position
  const int *pat_li, *pat_lj;  // [sum d] the same as group-local slots
  const int *slot_pose;    // group g's slots (ascending pose ids) at 2*g_pat[g] .. + W_g
  const int *slot_ptr;     // CSR over slots of the items feeding them, at 2*g_pat[g] + g .. + W_g + 1
  const int *slot_items;   // items (position*2 + {0: via source i, 1: via target j}) at 2*g_pat[g] ..
  // slots fed by >= 2 items ("multi" slots) are reduced through shared memory; the others are stored
  // directly by the one thread that produces them
  const int *g_nm;         // [G]   number of multi slots of group g
  const int *ms_ptr;       // ranks of the items of the k-th multi slot: [ms_ptr[k], ms_ptr[k+1]), at 2*g_pat[g] + g
  const int *ms_slot;      // slot id of the k-th multi slot, at 2*g_pat[g]
  const int *pat_ri, *pat_rj;  // [sum d] rank of the position's i-side / j-side item among multi items, or -1
  int dmax;                // longest track
  const int *c_t0, *c_grp; // [n_chunks+1], [n_chunks]  edge-pass work units (track ranges)
  const int *u_t0, *u_grp; // [n_units+1],  [n_units]   Schur work units
  const ChunkDesc *cdesc;  // [n_chunks]
  // lane-per-track edge pass (regular groups): CTA units of <= 32 * (kEdge2Warps / e2_kp) tracks
  int e2_kp;               // position splits per track slice (1, 2, 4 or 8)
  int e2_tpl;              // tracks per lane (1, or 2 on large graphs): CTA units of <= 32 * e2_tpl * (kEdge2Warps / e2_kp) tracks
  const int *x_t0, *x_grp; // [n_xchunks+1], [n_xchunks]
  int n_xchunks;
  const int *g_reg;        // [G] 1 = regular group (one source slot; no target slot fed more than kMaxSlotRun times)
  const int *pat_ps;       // [sum d] positions of a group ordered by (target slot, position)
  int n_irregular;         // groups that need the generic edge pass
  int dmax_irregular;      // longest track among them
  const int *patch_track;  // [NM] compact track of a patch or -1
  // streaming Schur -> solve hand-over: small units published in an order that serves both ends of the pose range first
  const int *o_t0, *o_grp; // [n_ounits+1], [n_ounits]  Schur units of <= 64 tracks
  const int *o_order;      // [n_ounits] unit run by CTA k
  int n_ounits;
  int *o_flag;             // unused (the flags live behind y in the reduced-system buffer and are cleared with it)
  const int *top_need, *bot_need;   // [N] see SolveFeed
};

// Per-call view: problem pointers + the layout of the reduced system for this fixedp.
struct CallView {
  const float *poses, *patches, *monodisp, *intr, *targets, *weights, *lmbda_vec;
  float lmbda, ep, alpha;
  float bounds[4];
  int fixedp, n, loss, structure_only;   // n = free poses
  int tstride;                           // floats per targets row (2 or 3)
  int M;                                 // 6 n
  int ld, off;                           // S(r,c), r >= c, lives at S[r*ld + c + off]
  int bw;                                // scalar half bandwidth (max r - c)
  double *S, *y;                         // reduced system (this rank's partial sums), fp64 accumulators
  float *Est;                            // E rows
  float2 *Cw;                            // per track (C, w) sums            (ba.py:287,292)
  float2 *Qw;                            // per track (Q, w adjusted)         (ba.py:303-311)
  double *dX, *L;                        // pose update and Cholesky factor, fp64
  float *dZ;
  int *status;
  float *poses_out, *patches_out;
};

}  // namespace ba

// Per-plan switches (ba_plan_set_option); defaults come from the environment once, at plan creation.
struct BaOptions {
  int solver;          // BA_OPT_SOLVER: 0 automatic (tile solver for short systems, diagonal-ownership DMMA band solver for long bands), 1 legacy
                       // circular-ownership DMMA solver, 2 scalar window, 3 dense, 4 shared-memory tile solver, 5 diagonal-ownership band solver
  int stream;          // BA_OPT_STREAM: Schur -> solve streaming hand-over in ba_step
  int stream_smem_kb;  // BA_OPT_STREAM_SMEM_KB: dynamic shared memory forced on the streamed Schur kernel (occupancy throttle, tests)
  int schur_tile;      // BA_OPT_SCHUR_TILE: tracks per cp.async stage of the SIMT Schur kernel
  int twist_min;       // BA_OPT_TWIST_MIN
  int spin_cap;        // BA_OPT_SPIN_CAP
  int trace;           // BA_OPT_SOLVER_TRACE
  int schur;           // BA_OPT_SCHUR: 0 tcgen05 (tensor-core) Schur kernel where it applies, 1 SIMT kernel
  int schur_acc;       // BA_OPT_SCHUR_ACC: chunks of 32 tracks accumulated in TMEM (fp32) before the fp64 read-back
};

// Capacities of a plan's arrays. ba_plan_create sizes them exactly (one read-back between the two halves of the build); a
// capacity plan (ba_plan_create_capacity) is sized once for the largest graph the caller will ever hand to ba_plan_update.
struct BaCaps { int64_t E; int m, G, pat, units[4]; int64_t est; };
// host-chosen overrides of the derived unit lengths (environment, experiments): -1 = derive on device
struct BaTuning { int tc, tu, kp, to, gend, tpl; };
// Everything the device-side build of a plan writes (ba_plan.cu): plan arrays (the PlanView points into them), build
// scratch, the shape block and its pinned host copy.
struct PlanBuild {
  BaCaps caps;
  BaTuning tun;
  int front_G;                         // capacity of the front half's group arrays (>= caps.G)
  unsigned *key, *skey, *eij, *sij;
  int *val, *tflag, *tinc, *gflag, *ginc, *g_d, *maxo;
  int4 *cflag, *cinc;
  long long *g_esz, *g_eoff;
  char *cub_tmp;
  size_t cub_bytes;
  int *eperm, *kx, *tptr, *t_grp, *g_t0, *g_pat, *g_W, *g_reg, *g_nm;
  int *pat_i, *pat_j, *pat_li, *pat_lj, *pat_ri, *pat_rj, *pat_ps, *slot_pose, *slot_ptr, *slot_items, *ms_ptr, *ms_slot;
  int *unit_t0[4], *unit_grp[4], *order, *top_need, *bot_need, *patch_track;
  ba::ChunkDesc *cdesc;
  int *shape_dev, *shape_host;
  cudaEvent_t ev_shape;
  const void *g_key[4];                // (ii, jj, kk, n_edges_dev) of the last ba_plan_update and, when they repeat,
  int64_t g_n;                         // the captured graph of the whole derivation
  cudaGraphExec_t g_exec;
  cudaStream_t g_stream;
  cudaEvent_t g_ev_in;
  int g_failed;
  int pending;                         // a build is in flight: the host has not read its shape block yet
  int valid;                           // the plan describes a graph
};

struct BaPlan {
  BaPlanInfo info;
  PlanBuild b;
  BaOptions opt;
  long long *trace_buf;
  ba::PlanView v;
  int n_total_layout, bwb_layout;      // what the reduced-system layout uses (>= the local values)
  int min_unit, max_unit;              // shortest / longest Schur unit (tracks): which of the two Schur kernels have work
  int min_ounit, max_ounit;            // the same for the units of the streaming hand-over
  int device;
  // workspace
  double *SY, *L, *dX, *Wg;            // Wg: inverted diagonal tiles of the tensor-core solver
  float *Est, *dZ;
  float2 *Cw, *Qw;
  int *status;
  int64_t sy_floats;                   // capacity of SY (elements)
  // The reduced system is double-buffered when it is small: the back-substitution kernel of a call clears the OTHER buffer
  // for the next call (no memset node in steady state: ~6 us per call), while [S | y] of the call itself stay readable
  // (ba_plan_debug_dense, sharded exchange). sy_clean[b] = leading bytes of buffer b known to be zero.
  double *SY2;
  int sy_cur;
  int64_t sy_clean[2];
  int sy_untracked;                    // a call on this plan was captured into a CUDA graph: replays dirty buffers unseen
  double *sy_last;                     // buffer of the last call that assembled a system
  int64_t est_floats;                  // size of Est
  int last_n, last_fixedp;             // layout of the last ba_assemble
  float *pp_buf[2];                    // ping-pong (poses | patches) buffers of ba_update
  // double-buffered staging / copy streams of ba_step_host[_async] (ba_aux.cu)
  void *host_pipe;
  void (*host_pipe_destroy)(void *);
  std::vector<void *> owned;           // every pool block of the plan, for ba_plan_destroy
  cudaStream_t mem_stream;             // stream the plan's allocations are ordered on
  // streaming Schur -> solve (single-device ba_step): the solve runs on its own stream next to the Schur kernel
  cudaStream_t solve_stream;
  cudaEvent_t ev_step_begin, ev_solved;
  int epoch;
  long long solve_shape_key;
  // optional per-stage timing (ba_plan_enable_timing)
  int timing;
  cudaEvent_t ev[BA_N_STAGES + 1];
  unsigned ev_mask;                    // which stage boundaries were recorded by the last step
};

namespace ba {
extern std::atomic<long long> g_launches;
int set_cuda_error(cudaError_t e, const char *what);
// host side of a finished plan build: waits for the shape block of a pending ba_plan_update (no-op otherwise)
int plan_finalize(BaPlan *pl);
void layout_for(const BaPlan *p, int fixedp, int *n, int *bw, int *ld, int *off, int64_t *s_floats);
constexpr int kMmaMaxBw = 120;        // widest band the 16x16-tile register window of the DMMA solver covers
size_t solve_mma_smem_bytes(int M);
size_t solve_mma_scratch_doubles(int M, int bw);
// Streaming mode of the band solver: the Schur kernel runs concurrently and publishes, per unit and in a fixed
// order, a completion flag (= epoch, 1: the flags are cleared with S at the start of the call); top_need[p] / bot_need[p] = number of leading units of that order that must be
// complete before the rows of every pose <= p / >= p are final. flags == nullptr: S and y are final at launch.
// mode 0: plain solve; 1: streaming (writes redo[0] = 1 if it gave up waiting: the producer did not run concurrently,
// e.g. kernels serialised by a profiler); 2: stand-by launch, does the plain solve only if redo[0] != 0.
struct SolveFeed {
  const int *flags, *top_need, *bot_need;
  int epoch, n_units, fixedp;
  int mode;
  int *redo;
  long long *shape_key;     // host: (M, bw) the solver scratch was last cleared for (plan-owned)
  int spin_cap;             // give-up bound of the streaming waits (iterations of a 40 ns sleep); 0 = default (~3 ms)
  int twist_min;            // tile columns from which the two-CTA twisted factorisation is used; 0 = default (64)
  long long *trace;         // optional per-column clock stamps (BA_OPT_SOLVER_TRACE), device memory owned by the plan
};
inline SolveFeed make_feed(const BaPlan *pl, const int *flags, const int *top_need, const int *bot_need, int epoch, int n_units,
                           int fixedp, int mode, int *redo);
int launch_solve_band_mma(const CallView &cv, int allow_retry, double *scratch, const SolveFeed &feed, cudaStream_t s);
// diagonal-ownership band solver (ba_solve_diag.cu): same contract, the default
int launch_solve_band_diag(const CallView &cv, int allow_retry, double *scratch, const SolveFeed &feed, cudaStream_t s);
size_t solve_diag_smem_bytes(int M);
// shared-memory tile solver for small / medium systems (ba_solve_tile.cu): one CTA, band window as 8x8 tiles
bool solve_tiles_applies(int M, int bw, int *nbt_out);
size_t solve_tiles_scratch_doubles(int M, int bw);
int launch_solve_tiles(const CallView &cv, int allow_retry, double *scratch, long long *trace, cudaStream_t s);
int solve_tiles_prepare_device();
// per-device one-time setup (function attributes, static tables); called from ba_plan_create under a lock
int solve_diag_prepare_device();
int solve_mma_prepare_device(int dev, cudaStream_t s);
int kernels_prepare_device();
int schur_tc_prepare_device();
constexpr int kSchurTcMaxFree = 20;     // free pose slots of a group the tensor-core Schur kernel covers (6 * 20 + 1 <= 128 operand rows)
int launch_schur_tc(const PlanView &pv, const CallView &cv, int n_units, const int *ut0, const int *ugrp, const int *order,
                    int *flags, int epoch, int acc_chunks, int min_tracks, long long *trace, cudaStream_t s);
constexpr int kSchurTcMinTracks = 96;   // shorter units stay on the SIMT kernel (faster there, and the 8-keyframe class of
                                        // ill-conditioned small windows keeps plain fp32 products)
}  // namespace ba

namespace ba {
inline SolveFeed make_feed(const BaPlan *pl, const int *flags, const int *top_need, const int *bot_need, int epoch, int n_units,
                           int fixedp, int mode, int *redo) {
  SolveFeed f;
  f.flags = flags; f.top_need = top_need; f.bot_need = bot_need; f.epoch = epoch; f.n_units = n_units; f.fixedp = fixedp;
  f.mode = mode; f.redo = redo; f.shape_key = const_cast<long long *>(&pl->solve_shape_key);
  f.spin_cap = pl->opt.spin_cap; f.twist_min = pl->opt.twist_min; f.trace = (pl->opt.trace & 1) ? pl->trace_buf : nullptr;
  return f;
}
}  // namespace ba

#define BA_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) return ba::set_cuda_error(e_, #call);              \
  } while (0)

// NVTX ranges (SURVEY.md §5 tracing): one host-side range per stage of a BA call around its launches ("ba:edge_pass",
// "ba:schur", ...), and one around a plan build / update; free when no tool is attached.
namespace ba {
inline void nvtx_stage(int k) {
  static const char *const names[BA_N_STAGES] = {"ba:zero", "ba:edge_pass", "ba:track_q", "ba:schur", "ba:solve", "ba:backsub", "ba:pose_retr"};
  static thread_local bool open = false;
  if (open) { nvtxRangePop(); open = false; }
  if (k >= 0 && k < BA_N_STAGES) { nvtxRangePushA(names[k]); open = true; }
}
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace ba

#define BA_MARK(pl, k, s)                                                     \
  do {                                                                        \
    ba::nvtx_stage(k);                                                        \
    if ((pl)->timing) { BA_CUDA(cudaEventRecord((pl)->ev[(k)], (s))); (pl)->ev_mask |= 1u << (k); } \
  } while (0)

#define BA_LAUNCH_CHECK()                                                     \
  do {                                                                        \
    ba::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t e_ = cudaGetLastError();                                      \
    if (e_ != cudaSuccess) return ba::set_cuda_error(e_, "kernel launch");    \
  } while (0)
