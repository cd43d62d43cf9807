/* batrack_ba.h — C ABI of the B200-native bundle-adjustment backend (libbatrack_ba.so).
 *
 * Drop-in boundary for BA-Track's sparse-SLAM hot path. Every entry point names the reference
 * interface it replaces (paths relative to the reference repo root):
 *
 *   ba_plan_*        the per-call index preparation of main/backend/ba.py:219,269-277
 *                    (ii.max()/jj.max(), index shift, torch.unique(kk)) — hoisted out of the
 *                    iteration and cached per graph topology
 *   ba_step          one call of BA_rgbd_droid (main/backend/ba.py:217-339) or BA (:103-213)
 *   ba_assemble /    the same call split at the point where a keyframe-sharded graph exchanges the
 *   ba_solve_update  reduced camera system: ba.py:223-322 | ba.py:323-337
 *   ba_reproject     pops.transform(..., jacobian=False) (main/backend/projective_ops.py:54-70,102-105)
 *   se3_*            the SE3 forward entry points of the lietorch_backends extension
 *                    (main/backend/lietorch/src/lietorch.cpp:18,69,97,155,214,286-316)
 *
 * Conventions: all array arguments are DEVICE pointers unless a name ends in _host; float32 values,
 * int64 indices exactly as the reference's caller holds them (main/batrack.py:864-875); every launch
 * goes to the caller's stream (`stream` is a cudaStream_t passed as void*); nothing synchronises the
 * host except ba_plan_create (once per topology) and the *_host helpers. Inputs are never modified.
 * Return value: 0 on success, a negative BA_ERR_* code otherwise (ba_error_string() explains it).
 * Numerical failure is NOT an error, exactly like the reference: a failed Cholesky leaves the poses
 * unchanged (ba.py:9-13), NaNs in the pose update trigger one retry with lm = 1e-3 (ba.py:324-325).
 */
#ifndef BATRACK_BA_H
#define BATRACK_BA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BA_OK 0
#define BA_ERR_CUDA (-1)          /* a CUDA runtime call failed (see ba_last_cuda_error) */
#define BA_ERR_ARG (-2)           /* null pointer / negative size / unsupported value */
#define BA_ERR_INDEX_RANGE (-3)   /* an edge index is outside [0,N) / [0,NM) */
#define BA_ERR_TOO_MANY_POSES (-4)/* N > 65535 (pose indices are packed to 16 bits) */
#define BA_ERR_NO_DEVICE (-5)     /* no sm_100 device / kernel image not loadable */
#define BA_ERR_CAPACITY (-6)      /* the graph handed to ba_plan_update exceeds a capacity of the plan */

#define BA_LOSS_TRIVIAL 0         /* compute_kernel_weight, main/backend/ba.py:81-100 */
#define BA_LOSS_HUBER 1
#define BA_LOSS_CAUCHY 2

typedef struct BaPlan BaPlan;     /* opaque: cached topology + per-iteration workspace */

/* What ba_plan_create learned about the graph (host-side copy). */
typedef struct {
  int64_t n_edges;       /* E */
  int32_t n_poses;       /* N: length of the pose buffer */
  int32_t n_patches;     /* NM: length of the patch buffer */
  int32_t n_total;       /* max(ii.max(), jj.max()) + 1          (ba.py:219) */
  int32_t n_tracks;      /* m = len(unique(kk))                  (ba.py:276-277) */
  int32_t n_groups;      /* runs of consecutive tracks sharing one (ii,jj) edge pattern */
  int32_t n_chunks;      /* CTA work units of the edge pass */
  int32_t max_degree;    /* longest track (edges) */
  int32_t max_slots;     /* most distinct poses touched by one group */
  int32_t block_bandwidth; /* max |pose_a - pose_b| over poses coupled by one group */
  int32_t perm_identity; /* 1 if the caller's edge order is already track-major */
  int32_t banded;        /* 1 if the reduced system is stored/solved in band form */
  int64_t workspace_bytes;
} BaPlanInfo;

/* One BA call. Field comments give the reference argument (ba.py:217) and its torch shape. */
typedef struct {
  const float *poses;        /* poses.data        [1,N,7]  tx ty tz qx qy qz qw */
  const float *patches;      /* patches           [1,NM,3,1,1]  x y inverse-depth (P = 1) */
  const float *monodisp;     /* patches_monodisp  [1,NM,1]; NULL selects BA (ba.py:103) */
  const float *intrinsics;   /* intrinsics        [1,N,4]  fx fy cx cy */
  const float *targets;      /* targets_2d        [1,E,2] */
  const float *weights;      /* weights           [1,E,2] */
  const float *lmbda_vec;    /* lmbda as a tensor of m values (ba.py:299-300) or NULL */
  float lmbda;               /* lmbda as a python float (used when lmbda_vec == NULL) */
  float ep;                  /* ep */
  float alpha;               /* alpha */
  float bounds[4];           /* bounds [x0,y0,x1,y1] */
  int32_t fixedp;            /* fixedp */
  int32_t structure_only;    /* structure_only */
  int32_t loss;              /* BA_LOSS_* */
  int32_t targets_stride;    /* floats between consecutive targets rows: 2, or 3 when the caller passes the
                                view targets_3d[...,:2] (main/batrack.py:871); 0 means 2 */
  float *poses_out;          /* returned poses.data [1,N,7] (fresh buffer; may NOT alias poses). A structure-only call leaves the
                                poses alone (ba.py:336-339): NULL is allowed then and nothing is copied */
  float *patches_out;        /* returned patches    [1,NM,3,1,1] (fresh buffer) */
} BaProblem;

/* ---- topology plan ------------------------------------------------------------------------- */

/* Builds the cached topology for (ii,jj,kk) [E] int64 device arrays: validates ranges, groups the
 * edges by track (stable), compacts kk -> [0,m) like torch.unique(sorted=True), finds pattern
 * groups / chunks, and allocates the workspace. Synchronises `stream`. */
int ba_plan_create(const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t n_edges,
                   int32_t n_poses, int32_t n_patches, void *stream, BaPlan **out);
void ba_plan_destroy(BaPlan *plan);
int ba_plan_info(const BaPlan *plan, BaPlanInfo *out);

/* Factor-graph maintenance without host synchronisation (replaces the per-frame plan rebuild behind
 * main/batrack.py:189-212 append_factors / remove_factors, :399-410 __edges, :1023-1073 keyframe).
 * ba_plan_create_capacity allocates a plan ONCE for the largest graph the caller will hand over: cap_edges edges,
 * cap_tracks patches with edges, cap_groups pattern groups (tracks with identical (ii,jj) lists: one per source keyframe in a
 * SLAM graph), cap_pattern pattern positions (sum over groups of the edges of one track), cap_est floats of per-track E
 * storage (0: 6 * (cap_edges + 4 * cap_tracks)). It describes no graph yet.
 * ba_plan_update re-derives the whole plan for a new (ii,jj,kk) ON THE DEVICE: ~35 kernel launches and one asynchronous
 * copy of the shape block (counts) to pinned host memory on `stream` — no synchronisation, no allocation. The counts are
 * read by the first call that uses the plan (or by ba_plan_finalize), which waits for that copy only. n_edges is an upper
 * bound of the edge count; n_edges_dev (device memory, may be NULL) holds the live count when the caller keeps it on the
 * device. The index arrays must stay valid until then. Returns BA_ERR_CAPACITY (from the call that finalizes) when the
 * graph does not fit: build an exact plan with ba_plan_create instead. */
int ba_plan_create_capacity(int64_t cap_edges, int32_t cap_tracks, int32_t cap_groups, int32_t cap_pattern, int64_t cap_est,
                            int32_t n_poses, int32_t n_patches, void *stream, BaPlan **out);
int ba_plan_update(BaPlan *plan, const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t n_edges,
                   const int32_t *n_edges_dev, void *stream);
int ba_plan_finalize(BaPlan *plan);

/* The factor graph itself on the device: edge list (ii, jj, kk) + the per-edge payload the caller keeps next to it
 * (targets_3d [E,3], weights [E,2], weights_pose [E,2]) in fixed capacity buffers, the live edge count in device memory.
 * Every operation is a few launches on `stream`, none synchronises; the buffers never move.
 *   ba_graph_append  main/batrack.py:189-204 append_factors: n edges (kk = patch[e], jj = frame[e], ii = ix[patch[e]]) with
 *                    their payload rows (NULL: zeros), appended behind the live edges
 *   ba_graph_remove  stable removal, payload included (main/batrack.py:206-212 remove_factors):
 *                    mode 0  mask[e] != 0                                  (any caller-computed mask, uint8 [n_upper])
 *                    mode 1  ix[kk[e]] < a                                 (:1023-1026, :1072-1073 removal window)
 *                    mode 2  ii[e] == a || jj[e] == a, then kk[ii > a] -= patches_per_frame, ii[ii > a] -= 1, jj[jj > a] -= 1
 *                                                                          (:1042-1051 keyframe())
 *   ba_graph_arrays  device pointers, the device-side count and the host's upper bound of it (what ba_plan_update takes)
 *   ba_graph_count   the live count (synchronises `stream`);  ba_graph_tighten: tell the graph a count learned elsewhere */
typedef struct BaGraph BaGraph;
int ba_graph_create(int64_t cap_edges, void *stream, BaGraph **out);
void ba_graph_destroy(BaGraph *graph);
int ba_graph_arrays(const BaGraph *graph, int64_t **ii, int64_t **jj, int64_t **kk, float **targets_3d, float **weights,
                    float **weights_pose, int32_t **n_edges_dev, int64_t *n_edges_upper, int64_t *capacity);
int ba_graph_append(BaGraph *graph, const int64_t *patch, const int64_t *frame, int64_t n, const int64_t *ix,
                    const float *targets_3d, const float *weights, const float *weights_pose, void *stream);
int ba_graph_remove(BaGraph *graph, int32_t mode, int64_t a, int64_t patches_per_frame, const uint8_t *mask, const int64_t *ix,
                    void *stream);
int ba_graph_count(BaGraph *graph, int64_t *n_edges, void *stream);
int ba_graph_tighten(BaGraph *graph, int64_t n_edges);

/* Sharded graphs: make n_total / block_bandwidth agree across ranks so that every
 * rank lays the reduced camera system out identically. Pass the max over ranks. */
int ba_plan_set_layout(BaPlan *plan, int32_t n_total, int32_t block_bandwidth);

/* Per-plan switches (defaults: the values below, or the environment variable of the same name without the
 * BA_OPT_ prefix — e.g. BA_SOLVER, BA_STREAM — read once when the plan is created). Returns BA_ERR_ARG for an
 * unknown key or an unsupported value. None of them changes results beyond summation order. */
#define BA_OPT_SOLVER 1          /* 0 automatic (default): shared-memory tile solver for short systems (< 64 tile columns, half
                                    bandwidth <= 145), DMMA band solver with diagonal tile ownership for long bands, scalar
                                    window / dense single-CTA Cholesky for the rest; 1 DMMA band solver, circular ownership
                                    (round-1 kernel); 2 scalar window Cholesky; 3 dense single-CTA Cholesky; 4 tile solver
                                    wherever it applies; 5 diagonal-ownership band solver wherever it applies */
#define BA_OPT_STREAM 2          /* 1: ba_step starts the band solver next to the Schur kernel (default 0: the plain sequence is faster) */
#define BA_OPT_STREAM_SMEM_KB 3  /* dynamic shared memory forced on the streamed Schur kernel (occupancy throttle; tests) */
#define BA_OPT_SCHUR_TILE 4      /* tracks per cp.async stage of the SIMT Schur kernel (default 64) */
#define BA_OPT_TWIST_MIN 5       /* tile columns (8 unknowns each) from which two CTAs eliminate from both ends (default 64) */
#define BA_OPT_SPIN_CAP 6        /* give-up bound of the streaming solver's waits, in 40 ns sleeps (default 65536 ~ 3 ms) */
#define BA_OPT_SOLVER_TRACE 7    /* 1: the band solver records per-column clock stamps (ba_plan_read_trace) */
#define BA_OPT_SCHUR 8           /* 0 (default): tcgen05 tensor-core Schur kernel where it applies; 1: SIMT kernel */
#define BA_OPT_SCHUR_ACC 9       /* chunks of 32 tracks the tensor-core Schur kernel accumulates in TMEM (fp32) before the fp64
                                    read-back: 1..8, default 2 */
int ba_plan_set_option(BaPlan *plan, int32_t key, int32_t value);
int ba_plan_get_option(const BaPlan *plan, int32_t key, int32_t *value);
/* Copies the solver trace of the last traced solve to the host: n_values int64 clock stamps
 * ([column][16 slots] for up to 4096 columns, then 4 phase stamps per side). Synchronises `stream`. */
int ba_plan_read_trace(const BaPlan *plan, int64_t *out_host, int64_t n_values, void *stream);

/* Device copy-outs for tests and callers that need the compacted indices (ba.py:276): kx[m] int32
 * patch index of every compact track, sorted ascending. */
int ba_plan_tracks(const BaPlan *plan, int32_t *kx_out /* device, m ints */, void *stream);

/* ---- the iteration ------------------------------------------------------------------------- */

/* One full BA call on one device: assemble -> Schur -> solve -> back-substitute -> retract.
 * Everything is enqueued on `stream` and ordered with it; no host synchronisation. When the reduced system goes to
 * the band solver, the solver runs on a plan-owned high-priority stream next to the Schur kernel (event-ordered with
 * `stream` on both sides; DESIGN.md §4 "Streaming hand-over"); the call keeps no per-call state on the host, so it
 * may be captured into a CUDA graph (after one warm-up call on the capture stream) and replayed. One plan = one call
 * in flight: calls on the same plan must be issued on one stream (they share the workspace). */
int ba_step(BaPlan *plan, const BaProblem *prob, void *stream);

/* The BA driver loop of BATRACK.update (main/batrack.py:869-875) in one call: `iters` times
 *   { pose + depth step on prob->weights (weights_pose), depth-only step on weights_all (weights) },
 * each step consuming the previous step's poses / patches. prob->structure_only is ignored; results of the
 * last step land in prob->poses_out / patches_out. Intermediate states live in plan-owned buffers; the
 * topology plan, the workspace and every launch are shared by the 2*iters steps. */
int ba_update(BaPlan *plan, const BaProblem *prob, const float *weights_all, int32_t iters, void *stream);

/* First half: residuals, Jacobians, per-track Schur complement. Leaves the (partial) reduced camera
 * system in the plan's exchange buffer [S | y] (see ba_plan_reduced_system). No-op for
 * structure-only calls apart from the per-track C, w. */
int ba_assemble(BaPlan *plan, const BaProblem *prob, void *stream);

/* The exchange buffer: `n_values` contiguous DOUBLES holding the lower (band) storage of S followed
 * by y (the reduced system is accumulated in fp64). A sharded caller all-reduces
 * (sum) exactly this range between ba_assemble and ba_solve_update. Valid for the fixedp of the last
 * ba_assemble. */
int ba_plan_reduced_system(const BaPlan *plan, double **ptr, int64_t *n_values);

/* Second half: damped Cholesky solve of the reduced system, depth back-substitution, retractions. */
int ba_solve_update(BaPlan *plan, const BaProblem *prob, void *stream);

/* Debug/test view of the last reduced system: writes dense row-major S [6n,6n] (symmetrised),
 * y [6n], dX [6n], per-track C_damped^-1 = Q [m], w [m], dZ [m] (any pointer may be NULL). */
int ba_plan_debug_dense(const BaPlan *plan, int32_t n, float *S, float *y, float *dX, float *Q,
                        float *w, float *dZ, void *stream);

/* Solver status of the last ba_solve_update: bit0 = Cholesky failed (dX = 0), bit1 = NaN retry
 * with lm = 1e-3 taken, bit2 = retry failed as well. Device int; copy when needed. */
int ba_plan_status_ptr(const BaPlan *plan, int32_t **dev_status);

/* Per-kernel device timing of the last ba_step / ba_assemble + ba_solve_update (bench.py's roofline):
 * when enabled, CUDA events are recorded on the launch stream between the stages; ba_plan_last_timing
 * waits for the last event and writes BA_N_STAGES durations in milliseconds. */
#define BA_STAGE_ZERO 0      /* memset of the reduced system */
#define BA_STAGE_EDGE 1      /* edge pass: residuals + Jacobians + per-track reduction (the HBM-bound kernel) */
#define BA_STAGE_TRACKQ 2    /* per-track damping / prior */
#define BA_STAGE_SCHUR 3     /* per-track Schur complement */
#define BA_STAGE_SOLVE 4     /* damped Cholesky solve of the reduced camera system */
#define BA_STAGE_BACKSUB 5   /* depth back-substitution + disparity retraction */
#define BA_STAGE_RETR 6      /* pose retraction */
#define BA_N_STAGES 7
int ba_plan_enable_timing(BaPlan *plan, int enable);
int ba_plan_last_timing(BaPlan *plan, float *ms_out /* host, BA_N_STAGES floats */);

/* Host-buffer entry points (what a host-side caller holding CPU arrays binds; bench.py's e2e leg).
 * Every pointer of `prob_host` is a HOST array (pinned memory for copy/compute overlap; pageable works but
 * serialises); targets_stride must be 2. Inputs go to one of two plan-owned device staging slots on the plan's
 * own upload stream, ba_step runs on `stream`, results return on the plan's download stream: the upload of
 * call k+1 and the download of call k-1 overlap the kernels of call k.
 *   ba_step_host_async  enqueue one step and return; host inputs must stay untouched until the step's upload has
 *                       run and the host outputs are valid only after ba_host_sync. Host-buffer hazards are tracked: an
 *                       input array that one of the two calls still in flight downloads INTO (iteration k+1 starting
 *                       from the poses / patches iteration k returns, main/batrack.py:869-884) is uploaded after that
 *                       download, ordered on the device — dependent steps may be enqueued back to back, the host need
 *                       not wait in between.
 *   ba_host_sync        order `stream` after the download of every step submitted so far; block != 0 also waits
 *                       on the host.
 *   ba_step_host        one synchronous step (= async + blocking sync).
 *   ba_stage_host_async / ba_unstage_host_async   the two halves of ba_step_host_async for callers that put their own
 *                       launches between them — a sharded caller runs ba_assemble, its all-reduce of the reduced
 *                       system and ba_solve_update on the device-side problem `prob_dev` that staging returns. */
int ba_step_host_async(BaPlan *plan, const BaProblem *prob_host, void *stream);
int ba_stage_host_async(BaPlan *plan, const BaProblem *prob_host, BaProblem *prob_dev, void *stream);
int ba_unstage_host_async(BaPlan *plan, const BaProblem *prob_host, const BaProblem *prob_dev, void *stream);
/* Early upload for the NEXT ba_step_host_async / ba_stage_host_async call: whichever of targets, weights, intrinsics,
 * monodisp, lmbda_vec are non-NULL in `problem_host` (inputs that do not depend on the previous call's results; poses /
 * patches are ignored) go to that call's staging slot on the upload stream now, while the previous call still computes;
 * the next call uploads only the rest. Lets DEPENDENT steps (BATRACK.update feeds iteration k+1 with iteration k's poses
 * and patches, main/batrack.py:869-884) hide the per-step observations behind the kernels. */
int ba_prefetch_host_async(BaPlan *plan, const BaProblem *problem_host, void *stream);
int ba_host_sync(BaPlan *plan, void *stream, int block);
int ba_step_host(BaPlan *plan, const BaProblem *prob_host, void *stream);

/* ---- reprojection without Jacobians --------------------------------------------------------- */
/* coords[e] = pixel of patch kk[e] (seen in ii[e]) in frame jj[e]; valid[e] = Z > 0.2 (may be NULL).
 * tonly != 0 drops the rotation of Gij (projective_ops.py:63-64). */
int ba_reproject(const float *poses, const float *patches, const float *intrinsics,
                 const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t n_edges,
                 int32_t n_poses, int32_t n_patches, int32_t tonly, float *coords, float *valid,
                 void *stream);

/* transform() with every optional output of projective_ops.py:54-105: coords [E, depth ? 3 : 2] (the third channel is
 * the inverse depth of proj(depth=True), :47-50), valid [E] or NULL, and the analytic Jacobians Ji, Jj [E,2,6] and
 * Jz [E,2] of :72-100 (any of them may be NULL). */
int ba_transform(const float *poses, const float *patches, const float *intrinsics,
                 const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t n_edges,
                 int32_t n_poses, int32_t n_patches, int32_t tonly, int32_t depth, float *coords, float *valid,
                 float *Ji, float *Jj, float *Jz, void *stream);
/* point_cloud (projective_ops.py:107-109): out[k] = T_ix[k]^-1 * iproj(patch k), [n,4] homogeneous (x, y, z, inverse depth) */
int ba_point_cloud(const float *poses, const float *patches, const float *intrinsics, const int64_t *ix,
                   int64_t n_points, int32_t n_poses, float *out, void *stream);
/* back_proj (projective_ops.py:129-152): xy [B,n,2], depth [B,n], intrinsics [B,4], c2w [B,4,4] or NULL -> P [B,n,4] */
int ba_back_proj(const float *xy, const float *depth, const float *intrinsics, const float *c2w, int32_t B,
                 int64_t n, float *P, void *stream);
/* proj_to_frames (projective_ops.py:154-176): P [B,n,4], intrinsics [B,S,4], w2c [B,S,4,4] -> xy [B,S,n,2] */
int ba_proj_to_frames(const float *P, const float *intrinsics, const float *w2c, int32_t B, int32_t S, int64_t n,
                      float *xy, void *stream);

/* ---- SE3 forward ops (lietorch_backends, group id 3) ---------------------------------------- */
/* All arrays contiguous [B,7] / [B,6] / [B,4] float32 device buffers. */
int se3_expm(const float *a, float *X, int64_t B, void *stream);                  /* lietorch.cpp:18 */
int se3_logm(const float *X, float *a, int64_t B, void *stream);                  /* lietorch.cpp:43 */
int se3_inv(const float *X, float *Y, int64_t B, void *stream);                   /* lietorch.cpp:69 */
int se3_mul(const float *X, const float *Y, float *Z, int64_t B, void *stream);   /* lietorch.cpp:97 */
int se3_adj(const float *X, const float *a, float *b, int64_t B, void *stream);   /* lietorch.cpp:126 */
int se3_adjT(const float *X, const float *a, float *b, int64_t B, void *stream);  /* lietorch.cpp:155 */
int se3_act(const float *X, const float *p, float *q, int64_t B, void *stream);   /* lietorch.cpp:185 */
int se3_act4(const float *X, const float *p, float *q, int64_t B, void *stream);  /* lietorch.cpp:214 */
int se3_as_matrix(const float *X, float *T, int64_t B, void *stream);             /* lietorch.cpp:258 */

/* The same nine ops in double precision (the reference dispatches float and double, lietorch/include/dispatch.h:37-45;
 * its identity tests run in fp64, lietorch/run_tests.py:16-52). */
int se3d_expm(const double *a, double *X, int64_t B, void *stream);
int se3d_logm(const double *X, double *a, int64_t B, void *stream);
int se3d_inv(const double *X, double *Y, int64_t B, void *stream);
int se3d_mul(const double *X, const double *Y, double *Z, int64_t B, void *stream);
int se3d_adj(const double *X, const double *a, double *b, int64_t B, void *stream);
int se3d_adjT(const double *X, const double *a, double *b, int64_t B, void *stream);
int se3d_act(const double *X, const double *p, double *q, int64_t B, void *stream);
int se3d_act4(const double *X, const double *p, double *q, int64_t B, void *stream);
int se3d_as_matrix(const double *X, double *T, int64_t B, void *stream);

/* ---- misc ----------------------------------------------------------------------------------- */
const char *ba_error_string(int code);
const char *ba_last_cuda_error(void);
int ba_version(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t ba_launch_count(void);

/* Trajectory hand-off (main/batrack.py:223-228 get_pose, :898-915 terminate, :1080-1088 get_results): frame t of the
 * n_frames processed so far is a keyframe (slot[t] >= 0: row of `poses` [N,7]) or was dropped by keyframe() and carries
 * delta[t] = (t0[t], dP[t] [7]) with pose(t) = dP[t] * pose(t0[t]). Writes the inverted (world-from-camera) poses as
 * [tx ty tz qw qx qy qz] rows (out7, terminate()) and / or as 4x4 matrices (out44, cams_T_world); either may be NULL.
 * err (device int32) gets bit 0 when a frame has neither a pose nor a resolvable delta chain. */
int ba_trajectory(const float *poses, const int32_t *slot, const int32_t *t0, const float *dP, int32_t n_frames,
                  float *out7, float *out44, int32_t *err, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BATRACK_BA_H */
