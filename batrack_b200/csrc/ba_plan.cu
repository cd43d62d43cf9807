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

// Shape block: everything the build learns about the graph, kept in DEVICE memory while the plan is being derived (no
// kernel needs the host to read a count back) and copied to the host once, at the end.
enum { META_ERR = 0, META_MAXPOSE, META_NOTIDENT, META_DMAX, META_WMAX, META_SPAN, META_NIRREG, META_DMAX_IRREG, META_MINUNIT, META_MINOUNIT,
       SH_E, SH_M, SH_G, SH_CNT0, SH_CNT1, SH_CNT2, SH_CNT3, SH_PAT, SH_OVERFLOW, SH_LEN0, SH_LEN1, SH_LEN2, SH_LEN3, SH_KP, SH_OMAX,
       SH_ESIZE_LO, SH_ESIZE_HI, SH_GEND, SH_TPL, META_COUNT = 32 };

__global__ void k_shape_init(int *__restrict__ sh, int E_host, const int *__restrict__ E_dev) {
  const int t = threadIdx.x;
  if (t < META_COUNT) sh[t] = (t == META_MINUNIT || t == META_MINOUNIT) ? 0x7f7f7f7f : 0;
  __syncthreads();
  if (t == 0) sh[SH_E] = E_dev ? min(*E_dev, E_host) : E_host;
}

// keys beyond the live edge count sort to the end (key = NM)
__global__ void k_prep_keys(const int64_t *__restrict__ ii, const int64_t *__restrict__ jj,
                            const int64_t *__restrict__ kk, int E_up, int N, int NM,
                            unsigned *__restrict__ key, int *__restrict__ val, unsigned *__restrict__ eij,
                            int *__restrict__ meta) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E_up) return;
  val[e] = e;
  if (e >= meta[SH_E]) { key[e] = (unsigned)NM; eij[e] = 0; return; }
  int64_t i = ii[e], j = jj[e], k = kk[e];
  bool bad = i < 0 || i >= N || j < 0 || j >= N || k < 0 || k >= NM;
  if (bad) { atomicOr(&meta[META_ERR], 1); i = j = k = 0; }
  key[e] = (unsigned)k;
  eij[e] = (unsigned)i | ((unsigned)j << 16);
  int mx = (int)(i > j ? i : j);
  // one atomic per warp
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&meta[META_MAXPOSE], mx);
}

__global__ void k_track_flags(const unsigned *__restrict__ skey, const int *__restrict__ eperm,
                              const unsigned *__restrict__ eij, int E_up, int *__restrict__ tflag,
                              unsigned *__restrict__ sij, int *__restrict__ meta) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= E_up) return;
  if (q >= meta[SH_E]) { tflag[q] = 0; return; }
  tflag[q] = (q == 0 || skey[q] != skey[q - 1]) ? 1 : 0;
  int e = eperm[q];
  sij[q] = eij[e];
  if (e != q) meta[META_NOTIDENT] = 1;
}

__global__ void k_fill_tracks(const int *__restrict__ tflag, const int *__restrict__ tinc,
                              const unsigned *__restrict__ skey, int E_up, int cap_m, int *__restrict__ kx,
                              int *__restrict__ tptr, int *__restrict__ meta) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = meta[SH_E];
  if (q >= E) return;
  if (tflag[q]) {
    int t = tinc[q] - 1;
    if (t < cap_m) { kx[t] = (int)skey[q]; tptr[t] = q; }
  }
  if (q == E - 1) {
    const int m = tinc[q];
    meta[SH_M] = m;
    if (m > cap_m) atomicOr(&meta[SH_OVERFLOW], 1); else tptr[m] = E;
  }
}

// A track starts a new pattern group unless its edge list repeats the previous track's (ii,jj) list.
__global__ void k_group_flags(const int *__restrict__ tinc, const int *__restrict__ tptr,
                              const unsigned *__restrict__ sij, int E_up, int *__restrict__ gflag,
                              int *__restrict__ meta) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= meta[SH_E] || meta[SH_OVERFLOW]) return;
  int t = tinc[q] - 1;
  int pos = q - tptr[t];
  int dcur = tptr[t + 1] - tptr[t];
  if (pos == 0) atomicMax(&meta[META_DMAX], dcur);
  if (t == 0) { if (pos == 0) gflag[0] = 1; return; }
  int dprev = tptr[t] - tptr[t - 1];
  if (dcur != dprev) { if (pos == 0) gflag[t] = 1; return; }
  if (sij[q] != sij[tptr[t - 1] + pos]) gflag[t] = 1;
}

__global__ void k_fill_groups(const int *__restrict__ gflag, const int *__restrict__ ginc, int cap_m, int cap_G,
                              int *__restrict__ g_t0, int *__restrict__ t_grp, int *__restrict__ meta) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = meta[SH_M];
  if (t >= m || t >= cap_m) return;
  int g = ginc[t] - 1;
  t_grp[t] = g;
  if (gflag[t] && g < cap_G) g_t0[g] = t;
  if (t == m - 1) {
    meta[SH_G] = g + 1;
    if (g + 1 > cap_G) atomicOr(&meta[SH_OVERFLOW], 2); else g_t0[g + 1] = m;
  }
}

__global__ void k_group_degree(const int *__restrict__ g_t0, const int *__restrict__ tptr, int cap_G,
                               int *__restrict__ g_d, const int *__restrict__ meta) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > cap_G) return;
  const int G = meta[SH_OVERFLOW] ? 0 : meta[SH_G];
  g_d[g] = g < G ? tptr[g_t0[g] + 1] - tptr[g_t0[g]] : 0;
}

// Unit lengths and the position split of the lane-per-track edge pass, from the graph's size (what the host used to
// choose between two synchronisations). tun: host overrides (-1 = derive).
__global__ void k_derive(const int *__restrict__ g_pat, int cap_G, int cap_pat, BaTuning tun, int *__restrict__ meta) {
  if (threadIdx.x || blockIdx.x) return;
  const int m = meta[SH_M], G = min(meta[SH_G], cap_G), dmax = meta[META_DMAX];
  const int sms = 148;
  auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
  // edge pass: about one wave of CTAs (2 resident per SM) — per-CTA set-up and flush are amortised over more tracks;
  // Schur: units of <= 128 tracks, or whole 256-track groups (measured 4 us faster than two 128-track units at cfg3)
  int tc = min(256, max(8, cdiv(m, sms)));
  // Schur units: ~128 of them when the graph allows units of >= 96 tracks (the tensor-core kernel's territory: its flush
  // costs the same whatever the unit's length, so long units and about one per SM), shorter units for small graphs (SIMT
  // kernel). Measured: 64 KF x 256 tracks 31 us with 56-track units, 19 us with 128; a 2-GPU shard of cfg3 63 -> ~20 us
  int tu = min(256, max(16, 16 * cdiv(cdiv(m, 128), 16)));
  if (tu > 128) tu = 256;
  if (tun.tc > 0) tc = tun.tc;
  if (tun.tu > 0) tu = tun.tu;
  // lane-per-track edge pass: a CTA of kEdge2Warps warps = KT track slices x KP position splits. One split is the cheapest
  // per edge (fewest flushes; measured at 256 KF / 64k tracks: 44 us vs 54 / 60 us with 2 / 4 splits); graphs with few
  // tracks split the positions until the machine sees ~12 warps per SM (25-frame window, 5200 tracks x <= 72 edges:
  // 70 / 53 / 39 us with 2 / 4 / 8 splits)
  // Two tracks per lane when there are enough tracks to keep every scheduler busy that way (the large graphs, where the
  // kernel is bound by instruction issue): constants, control flow and the transpose reduction are shared by two edges.
  int tpl = cdiv(m, 64) >= 4 * sms ? 2 : 1;
  if (tun.tpl > 0) tpl = tun.tpl;
  int kp = 1;
  while (kp < kEdge2Warps && (long long)cdiv(m, 32 * tpl) * kp < 12 * sms && cdiv(dmax, kp) > 2) kp *= 2;
  if (tun.kp > 0) kp = tun.kp;
  const int to = tun.to > 0 ? tun.to : 64;                   // Schur units of the streaming hand-over to the solver
  meta[SH_LEN0] = tc; meta[SH_LEN1] = tu; meta[SH_LEN2] = 32 * tpl * (kEdge2Warps / kp); meta[SH_LEN3] = to;
  meta[SH_KP] = kp; meta[SH_TPL] = tpl;
  const int gend = tun.gend >= 0 ? min(tun.gend, G) : G;     // optionally large units in the middle of the pose range
  meta[SH_GEND] = gend;
  meta[SH_OMAX] = gend < G - gend ? max(to, 256) : to;
  const int pat = meta[SH_OVERFLOW] ? 0 : g_pat[G];
  meta[SH_PAT] = pat;
  if (pat > cap_pat) atomicOr(&meta[SH_OVERFLOW], 4);
}

// Work units = a group's tracks split into the fewest equal pieces of at most len[k] consecutive tracks, for the four
// kinds of units at once (edge-pass chunks, Schur units, lane-per-track chunks, streaming Schur units: small at both
// ends of the pose range, where the solver starts, optionally large in the middle).
__global__ void k_chunk_flags(const int *__restrict__ t_grp, const int *__restrict__ g_t0, int cap_m,
                              int4 *__restrict__ cflag, int *__restrict__ meta) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cap_m) return;
  int4 f = make_int4(0, 0, 0, 0);
  if (t < meta[SH_M] && !meta[SH_OVERFLOW]) {
    const int g = t_grp[t];
    const int T = g_t0[g + 1] - g_t0[g], rel = t - g_t0[g];
    int *fp = &f.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int tc = meta[SH_LEN0 + k];
      if (k == 3 && g >= meta[SH_GEND] && g < meta[SH_G] - meta[SH_GEND]) tc = 256;
      const int pieces = (T + tc - 1) / tc;
      const int len = ((T + pieces - 1) / pieces + 3) & ~3;      // multiple of 4: 16-byte aligned starts in the E rows
      fp[k] = (rel % len == 0) ? 1 : 0;
      if (fp[k] && k == 1) atomicMin(&meta[META_MINUNIT], min(len, T - rel));   // shortest unit (the last piece of a group)
      if (fp[k] && k == 3) atomicMin(&meta[META_MINOUNIT], min(len, T - rel));
    }
  }
  cflag[t] = f;
}
struct UnitArrays { int *t0[4]; int *grp[4]; int cap[4]; };
__global__ void k_fill_chunks(const int4 *__restrict__ cflag, const int4 *__restrict__ cinc,
                              const int *__restrict__ t_grp, UnitArrays ua, int *__restrict__ meta) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = meta[SH_M];
  if (t >= m || meta[SH_OVERFLOW]) return;
  const int4 f = cflag[t], c = cinc[t];
  const int *fp = &f.x, *cp = &c.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (fp[k] && cp[k] - 1 < ua.cap[k]) { ua.t0[k][cp[k] - 1] = t; ua.grp[k][cp[k] - 1] = t_grp[t]; }
    if (t == m - 1) {
      meta[SH_CNT0 + k] = cp[k];
      if (cp[k] > ua.cap[k]) atomicOr(&meta[SH_OVERFLOW], 8); else ua.t0[k][cp[k]] = m;
    }
  }
}

// One CTA per group: slots (distinct poses, ascending), local slot of every pattern position, and
// the per-slot item lists used by the edge pass to reduce per-edge 6-vectors into E rows.
constexpr int kSlotsFast = 512;             // longest track whose item tables are built in shared memory
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
  if (g >= meta[SH_G] || meta[SH_OVERFLOW]) return;             // grid = capacity; the live count sits in the shape block
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
  // ---- item tables. Fast path (d <= kSlotsFast: every graph BA-Track builds): everything in shared memory, all
  //      threads; the serial path below keeps arbitrary degrees correct ----
  if (d <= kSlotsFast && W <= 2 * kSlotsFast) {
    int *sh = (int *)(sm + 2 * nwords + 1);
    int *s_slot = sh;                       // [2d] slot of item x (x = 2p: via source i, 2p+1: via target j)
    int *s_cnt = s_slot + 2 * kSlotsFast;   // [W+1] items per slot -> start of the slot's items
    int *s_cntj = s_cnt + 2 * kSlotsFast + 1;   // [W+1] j-side items per slot -> start among the j-side items
    int *s_mrank = s_cntj + 2 * kSlotsFast + 1; // [W+1] first rank of a multi slot's items, -1: not a multi slot
    int *s_misc = s_mrank + 2 * kSlotsFast + 1; // [0] nm, [1] R, [2] maxrun, [3] all positions share the source slot
    for (int x = tid; x <= W; x += nt) { s_cnt[x] = 0; s_cntj[x] = 0; }
    if (tid == 0) { s_misc[2] = 0; s_misc[3] = 1; }
    __syncthreads();
    for (int p = tid; p < d; p += nt) {
      const int a = pat_li[pat0 + p], b = pat_lj[pat0 + p];
      s_slot[2 * p] = a; s_slot[2 * p + 1] = b;
      atomicAdd(&s_cnt[a + 1], 1); atomicAdd(&s_cnt[b + 1], 1); atomicAdd(&s_cntj[b + 1], 1);
    }
    __syncthreads();
    const int s0 = s_slot[0];
    for (int p = tid; p < d; p += nt) if (s_slot[2 * p] != s0) s_misc[3] = 0;
    if (tid == 0) {                         // W <= 2d short prefixes, in shared memory
      int nm = 0, Rr = 0, mr = 0;
      int *mp = ms_ptr + sbase + g, *msl = ms_slot + sbase;
      for (int sl = 0; sl < W; ++sl) {
        const int c = s_cnt[sl + 1], cj = s_cntj[sl + 1];
        mr = max(mr, cj);
        if (c >= 2) { s_mrank[sl] = Rr; mp[nm] = Rr; msl[nm++] = sl; Rr += c; } else s_mrank[sl] = -1;
        s_cnt[sl + 1] = s_cnt[sl] + c;
        s_cntj[sl + 1] = s_cntj[sl] + cj;
      }
      mp[nm] = Rr;
      s_misc[0] = nm; s_misc[1] = Rr; s_misc[2] = mr;
    }
    __syncthreads();
    int *sp = slot_ptr + sbase + g, *it = slot_items + sbase;
    for (int x = tid; x <= W; x += nt) sp[x] = s_cnt[x];
    for (int x = tid; x < 2 * d; x += nt) {   // ascending item id inside a slot (deterministic sums)
      const int sl = s_slot[x];
      int before = 0, beforej = 0;
      for (int y = 0; y < x; ++y) { const bool same = s_slot[y] == sl; before += same; beforej += same && (y & 1); }
      it[s_cnt[sl] + before] = x;
      const int rk = s_mrank[sl] >= 0 ? s_mrank[sl] + before : -1;
      if (x & 1) {
        pat_rj[pat0 + (x >> 1)] = rk;
        pat_ps[pat0 + s_cntj[sl] + beforej] = x >> 1;       // positions ordered by (target slot, position)
      } else {
        pat_ri[pat0 + (x >> 1)] = rk;
      }
    }
    if (tid == 0) {
      g_nm[g] = s_misc[0];
      g_W[g] = W;
      g_esz[g] = (long long)((T + 3) & ~3) * 6 * W;          // E rows, entry-major: [6 W][Ts] floats, Ts = T rounded up to 4
      const bool reg = s_misc[2] <= kMaxSlotRun && s_misc[3];
      g_reg[g] = reg ? 1 : 0;
      if (!reg) { atomicAdd(&meta[META_NIRREG], 1); atomicMax(&meta[META_DMAX_IRREG], d); }
      atomicMax(&meta[META_WMAX], W);
      atomicMax(&meta[META_SPAN], slot_pose[sbase + W - 1] - slot_pose[sbase]);
    }
    return;
  }
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


// Streaming Schur -> solve hand-over. CTA k of the Schur kernel runs unit o_order[k]: units from both ends of the
// track range alternate, so the rows the two elimination fronts of the solver need first are final first.
__global__ void k_unit_order(const int *__restrict__ meta, int *__restrict__ order) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = meta[SH_OVERFLOW] ? 0 : meta[SH_CNT3];
  if (k < n) order[k] = (k & 1) ? n - 1 - (k >> 1) : (k >> 1);
}
// maxorder[q] = 1 + the last CTA (in launch order) whose unit touches pose q
__global__ void k_unit_reach(const int *__restrict__ order, const int *__restrict__ o_grp, const int *__restrict__ g_pat,
                             const int *__restrict__ g_W, const int *__restrict__ slot_pose, const int *__restrict__ meta,
                             int *__restrict__ maxorder) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (meta[SH_OVERFLOW] || k >= meta[SH_CNT3]) return;
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
__global__ void k_patch_track(const int *__restrict__ kx, const int *__restrict__ meta, int *__restrict__ patch_track) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (!meta[SH_OVERFLOW] && t < meta[SH_M]) patch_track[kx[t]] = t;
}

__global__ void k_chunk_desc(PlanView v, const int *__restrict__ meta, long long *__restrict__ g_eoff_w, int64_t cap_est, ChunkDesc *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) {                                                  // the E storage this graph needs (the scan's last entry)
    const long long es = meta[SH_OVERFLOW] ? 0 : g_eoff_w[meta[SH_G]];
    const_cast<int *>(meta)[SH_ESIZE_LO] = (int)(es & 0xffffffffll);
    const_cast<int *>(meta)[SH_ESIZE_HI] = (int)(es >> 32);
    if (es > cap_est) atomicOr(const_cast<int *>(meta) + SH_OVERFLOW, 16);
  }
  if (meta[SH_OVERFLOW] || c >= meta[SH_CNT0]) return;
  ChunkDesc d;
  d.g = v.c_grp[c]; d.t0 = v.c_t0[c]; d.t1 = v.c_t0[c + 1]; d.gt0 = v.g_t0[d.g];
  d.pat0 = v.g_pat[d.g]; d.d = v.g_pat[d.g + 1] - d.pat0; d.W = v.g_W[d.g]; d.ebase = v.tptr[d.gt0];
  d.nm = v.g_nm[d.g]; d.R = v.ms_ptr[2 * d.pat0 + d.g + d.nm]; d.Ts = (v.g_t0[d.g + 1] - d.gt0 + 3) & ~3; d.reg = v.g_reg[d.g];
  d.eoff = v.g_eoff[d.g]; d.pad2 = 0;
  out[c] = d;
}

// single-CTA inclusive / exclusive prefix sums for the plan's short arrays (tracks, groups): one launch instead of
// CUB's two, and no temporary storage. T = int, int4 or long long.
__device__ __forceinline__ int4 operator+(const int4 &a, const int4 &b) { return make_int4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ int zero_of<int>() { return 0; }
template <> __device__ __forceinline__ long long zero_of<long long>() { return 0; }
template <> __device__ __forceinline__ int4 zero_of<int4>() { return make_int4(0, 0, 0, 0); }
__device__ __forceinline__ int shfl_up_t(int v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
__device__ __forceinline__ long long shfl_up_t(long long v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
__device__ __forceinline__ int4 shfl_up_t(int4 v, int o) {
  return make_int4(__shfl_up_sync(0xffffffffu, v.x, o), __shfl_up_sync(0xffffffffu, v.y, o), __shfl_up_sync(0xffffffffu, v.z, o),
                   __shfl_up_sync(0xffffffffu, v.w, o));
}
// tiles of 1024 elements: coalesced load, warp-shuffle scan, warp totals through shared memory, running carry
template <typename T, bool EXCL> __global__ void __launch_bounds__(1024) k_scan1(const T *__restrict__ in, T *__restrict__ out, int n) {
  __shared__ T wsum[32];
  __shared__ T s_carry;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) s_carry = zero_of<T>();
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const T x = i < n ? in[i] : zero_of<T>();
    T v = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const T u = shfl_up_t(v, o); if (lane >= o) v = v + u; }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
      T ws = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const T u = shfl_up_t(ws, o); if (lane >= o) ws = ws + u; }
      wsum[lane] = ws;
    }
    __syncthreads();
    const T carry = s_carry;
    T incl = carry + v;
    if (w > 0) incl = incl + wsum[w - 1];
    const T up = shfl_up_t(v, 1);                          // (all lanes: the last tile is ragged)
    if (i < n) {
      if (EXCL) {                                          // incl - x without needing operator-
        T ex = carry;
        if (w > 0) ex = ex + wsum[w - 1];
        if (lane > 0) ex = ex + up;
        out[i] = ex;
      } else out[i] = incl;
    }
    __syncthreads();
    if (t == 1023) s_carry = incl;
    __syncthreads();
  }
}

// ---- small RAII helpers (host) ---------------------------------------------------------------
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
    pl->SY2 = nullptr;
    if ((size_t)need * sizeof(double) <= ((size_t)64 << 20)) BA_CUDA(own(pl, &pl->SY2, need));
    pl->sy_cur = 0; pl->sy_clean[0] = pl->sy_clean[1] = 0; pl->sy_last = nullptr;
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

// ---------------------------------------------------------------------------------------------------------------------
// Building a plan. The derivation runs entirely on the device: counts (tracks, groups, units, pattern length, E storage)
// stay in the shape block and every kernel is launched over the CAPACITY of its array and returns on the live count.
//   front half: sort by track, track / group boundaries, group degrees, derived unit lengths
//   back half:  work units, group slots and item tables, E offsets, unit order of the streaming hand-over, descriptors
// ba_plan_create sizes the arrays exactly: one synchronisation between the halves (groups, pattern length) and one at
// the end (E storage). A capacity plan (ba_plan_create_capacity + ba_plan_update) owns everything up front: an update
// enqueues ~35 launches and one asynchronous copy of the shape block to pinned host memory — no synchronisation, no
// allocation; the first call that uses the plan waits for that copy's event (finalize), which in a SLAM front end lies
// a whole tracker inference later (main/batrack.py:189-212, 399-410, 1023-1073 change the graph, :856-895 runs BA).
// ---------------------------------------------------------------------------------------------------------------------
namespace ba {

static BaTuning tuning_from_env() {
  BaTuning t = {-1, -1, -1, -1, -1, -1};
  if (const char *e = getenv("BA_EDGE2_TPL")) { int v2 = atoi(e); if (v2 == 1 || v2 == 2) t.tpl = v2; }
  if (const char *e = getenv("BA_EDGE_TC")) t.tc = std::max(1, atoi(e));
  if (const char *e = getenv("BA_SCHUR_TU")) t.tu = std::max(1, atoi(e));
  if (const char *e = getenv("BA_EDGE2_KP")) { int v2 = atoi(e); if (v2 == 1 || v2 == 2 || v2 == 4 || v2 == 8) t.kp = v2; }
  if (const char *e = getenv("BA_STREAM_TU")) t.to = std::max(16, atoi(e) & ~3);
  if (const char *e = getenv("BA_STREAM_GEND")) t.gend = std::max(0, atoi(e));
  return t;
}

#define PB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_cuda_error(e_, #call); } while (0)
#define PB_LAUNCH() do { g_launches.fetch_add(1); PB_CUDA(cudaGetLastError()); } while (0)

// arrays of the front half + build scratch, for up to cap_E edges / cap_m tracks / cap_G groups
static int alloc_front(BaPlan *pl, int64_t cap_E, int cap_m, int cap_G, int N, int NM) {
  PlanBuild &b = pl->b;
  b.caps.E = cap_E; b.caps.m = cap_m;
  b.front_G = cap_G;
  PB_CUDA(own(pl, &b.key, cap_E)); PB_CUDA(own(pl, &b.skey, cap_E)); PB_CUDA(own(pl, &b.eij, cap_E)); PB_CUDA(own(pl, &b.sij, cap_E));
  PB_CUDA(own(pl, &b.val, cap_E)); PB_CUDA(own(pl, &b.tflag, cap_E)); PB_CUDA(own(pl, &b.tinc, cap_E)); PB_CUDA(own(pl, &b.eperm, cap_E));
  PB_CUDA(own(pl, &b.kx, cap_m)); PB_CUDA(own(pl, &b.tptr, (size_t)cap_m + 1)); PB_CUDA(own(pl, &b.t_grp, cap_m));
  PB_CUDA(own(pl, &b.gflag, cap_m)); PB_CUDA(own(pl, &b.ginc, cap_m));
  PB_CUDA(own(pl, &b.cflag, cap_m)); PB_CUDA(own(pl, &b.cinc, cap_m));
  PB_CUDA(own(pl, &b.g_t0, (size_t)cap_G + 1)); PB_CUDA(own(pl, &b.g_pat, (size_t)cap_G + 1)); PB_CUDA(own(pl, &b.g_d, (size_t)cap_G + 1));
  PB_CUDA(own(pl, &b.shape_dev, META_COUNT));
  PB_CUDA(own(pl, &b.maxo, N)); PB_CUDA(own(pl, &b.top_need, N)); PB_CUDA(own(pl, &b.bot_need, N));
  PB_CUDA(own(pl, &b.patch_track, NM));
  if (!b.shape_host) PB_CUDA(cudaMallocHost(&b.shape_host, META_COUNT * sizeof(int)));
  if (!b.ev_shape) PB_CUDA(cudaEventCreateWithFlags(&b.ev_shape, cudaEventDisableTiming));
  {                                      // CUB temporary storage for the two edge-sized primitives
    size_t b1 = 0, b2 = 0;
    PB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, b.key, b.skey, b.val, b.eperm, (int)cap_E, 0, 32, (cudaStream_t)0));
    PB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b2, b.tflag, b.tinc, (int)cap_E, (cudaStream_t)0));
    b.cub_bytes = std::max(b1, b2);
    PB_CUDA(own(pl, &b.cub_tmp, b.cub_bytes));
  }
  return BA_OK;
}

// arrays of the back half for up to G groups, pat pattern positions, units[k] work units
static int alloc_back(BaPlan *pl, int cap_G, int cap_pat, const int units[4]) {
  PlanBuild &b = pl->b;
  b.caps.G = cap_G; b.caps.pat = cap_pat;
  for (int k = 0; k < 4; ++k) {
    b.caps.units[k] = units[k];
    PB_CUDA(own(pl, &b.unit_t0[k], (size_t)units[k] + 1)); PB_CUDA(own(pl, &b.unit_grp[k], units[k]));
  }
  PB_CUDA(own(pl, &b.g_W, cap_G)); PB_CUDA(own(pl, &b.g_eoff, (size_t)cap_G + 1)); PB_CUDA(own(pl, &b.g_esz, (size_t)cap_G + 1));
  PB_CUDA(own(pl, &b.g_reg, cap_G)); PB_CUDA(own(pl, &b.g_nm, cap_G));
  const size_t P = (size_t)cap_pat;
  PB_CUDA(own(pl, &b.pat_ri, P)); PB_CUDA(own(pl, &b.pat_rj, P)); PB_CUDA(own(pl, &b.ms_slot, 2 * P)); PB_CUDA(own(pl, &b.ms_ptr, 2 * P + cap_G + 1));
  PB_CUDA(own(pl, &b.pat_i, P)); PB_CUDA(own(pl, &b.pat_j, P)); PB_CUDA(own(pl, &b.pat_ps, P)); PB_CUDA(own(pl, &b.pat_li, P)); PB_CUDA(own(pl, &b.pat_lj, P));
  PB_CUDA(own(pl, &b.slot_pose, 2 * P)); PB_CUDA(own(pl, &b.slot_items, 2 * P)); PB_CUDA(own(pl, &b.slot_ptr, 2 * P + cap_G + 1));
  PB_CUDA(own(pl, &b.order, units[3])); PB_CUDA(own(pl, &b.cdesc, units[0]));
  return BA_OK;
}

static int build_front(BaPlan *pl, const int64_t *ii, const int64_t *jj, const int64_t *kk, int E_up, const int *E_dev, cudaStream_t s) {
  PlanBuild &b = pl->b;
  const int TB = 256, N = pl->v.N, NM = pl->v.NM;
  int *meta = b.shape_dev;
  k_shape_init<<<1, 32, 0, s>>>(meta, E_up, E_dev); PB_LAUNCH();
  k_prep_keys<<<cdiv(E_up, TB), TB, 0, s>>>(ii, jj, kk, E_up, N, NM, b.key, b.val, b.eij, meta); PB_LAUNCH();
  {
    int end_bit = 1;                                             // keys 0 .. NM (NM = "beyond the live edges")
    while (end_bit < 32 && ((int64_t)1 << end_bit) <= (int64_t)NM) ++end_bit;
    size_t bytes = b.cub_bytes;
    PB_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, bytes, b.key, b.skey, b.val, b.eperm, E_up, 0, end_bit, s));
    g_launches.fetch_add(3);
  }
  k_track_flags<<<cdiv(E_up, TB), TB, 0, s>>>(b.skey, b.eperm, b.eij, E_up, b.tflag, b.sij, meta); PB_LAUNCH();
  {
    size_t bytes = b.cub_bytes;
    PB_CUDA(cub::DeviceScan::InclusiveSum(b.cub_tmp, bytes, b.tflag, b.tinc, E_up, s));
    g_launches.fetch_add(2);
  }
  k_fill_tracks<<<cdiv(E_up, TB), TB, 0, s>>>(b.tflag, b.tinc, b.skey, E_up, b.caps.m, b.kx, b.tptr, meta); PB_LAUNCH();
  PB_CUDA(cudaMemsetAsync(b.gflag, 0, (size_t)b.caps.m * sizeof(int), s));
  k_group_flags<<<cdiv(E_up, TB), TB, 0, s>>>(b.tinc, b.tptr, b.sij, E_up, b.gflag, meta); PB_LAUNCH();
  k_scan1<int, false><<<1, 1024, 0, s>>>(b.gflag, b.ginc, b.caps.m); PB_LAUNCH();
  k_fill_groups<<<cdiv(b.caps.m, TB), TB, 0, s>>>(b.gflag, b.ginc, b.caps.m, b.front_G, b.g_t0, b.t_grp, meta); PB_LAUNCH();
  k_group_degree<<<cdiv(b.front_G + 1, TB), TB, 0, s>>>(b.g_t0, b.tptr, b.front_G, b.g_d, meta); PB_LAUNCH();
  k_scan1<int, true><<<1, 1024, 0, s>>>(b.g_d, b.g_pat, b.front_G + 1); PB_LAUNCH();
  k_derive<<<1, 32, 0, s>>>(b.g_pat, b.front_G, b.caps.pat > 0 ? b.caps.pat : 0x7fffffff, b.tun, meta); PB_LAUNCH();
  return BA_OK;
}

static void view_pointers(BaPlan *pl) {
  PlanBuild &b = pl->b;
  PlanView &v = pl->v;
  v.eperm = b.eperm; v.kx = b.kx; v.tptr = b.tptr; v.t_grp = b.t_grp; v.g_t0 = b.g_t0; v.g_pat = b.g_pat; v.g_W = b.g_W;
  v.g_eoff = b.g_eoff; v.pat_i = b.pat_i; v.pat_j = b.pat_j; v.pat_li = b.pat_li; v.pat_lj = b.pat_lj;
  v.slot_pose = b.slot_pose; v.slot_ptr = b.slot_ptr; v.slot_items = b.slot_items;
  v.g_nm = b.g_nm; v.ms_ptr = b.ms_ptr; v.ms_slot = b.ms_slot; v.pat_ri = b.pat_ri; v.pat_rj = b.pat_rj;
  v.c_t0 = b.unit_t0[0]; v.c_grp = b.unit_grp[0]; v.u_t0 = b.unit_t0[1]; v.u_grp = b.unit_grp[1];
  v.x_t0 = b.unit_t0[2]; v.x_grp = b.unit_grp[2];
  v.o_t0 = b.unit_t0[3]; v.o_grp = b.unit_grp[3]; v.o_order = b.order; v.o_flag = nullptr;
  v.top_need = b.top_need; v.bot_need = b.bot_need;
  v.pat_ps = b.pat_ps; v.g_reg = b.g_reg; v.patch_track = b.patch_track; v.cdesc = b.cdesc;
}

static int build_back(BaPlan *pl, cudaStream_t s, bool record) {
  PlanBuild &b = pl->b;
  const int TB = 256, N = pl->v.N, NM = pl->v.NM;
  int *meta = b.shape_dev;
  const int cm = b.caps.m, cG = b.caps.G;
  k_chunk_flags<<<cdiv(cm, TB), TB, 0, s>>>(b.t_grp, b.g_t0, cm, b.cflag, meta); PB_LAUNCH();
  k_scan1<int4, false><<<1, 1024, 0, s>>>(b.cflag, b.cinc, cm); PB_LAUNCH();
  {
    UnitArrays ua;
    for (int k = 0; k < 4; ++k) { ua.t0[k] = b.unit_t0[k]; ua.grp[k] = b.unit_grp[k]; ua.cap[k] = b.caps.units[k]; }
    k_fill_chunks<<<cdiv(cm, TB), TB, 0, s>>>(b.cflag, b.cinc, b.t_grp, ua, meta); PB_LAUNCH();
  }
  PB_CUDA(cudaMemsetAsync(b.g_esz, 0, ((size_t)cG + 1) * sizeof(long long), s));
  {
    const int nwords = (N + 31) / 32;
    const size_t smem = (size_t)(2 * nwords + 1 + 8 * kSlotsFast + 3 + 8) * sizeof(int);
    k_group_slots<<<cG, 128, smem, s>>>(b.g_t0, b.g_pat, b.tptr, b.sij, N, b.pat_i, b.pat_j, b.pat_li, b.pat_lj, b.slot_pose,
                                        b.slot_ptr, b.slot_items, b.g_nm, b.ms_ptr, b.ms_slot, b.pat_ri, b.pat_rj, b.g_W, b.g_esz, b.g_reg,
                                        b.pat_ps, meta); PB_LAUNCH();
  }
  k_scan1<long long, true><<<1, 1024, 0, s>>>(b.g_esz, b.g_eoff, cG + 1); PB_LAUNCH();
  PB_CUDA(cudaMemsetAsync(b.maxo, 0, (size_t)N * sizeof(int), s));
  {
    const int no = b.caps.units[3];
    k_unit_order<<<cdiv(no, TB), TB, 0, s>>>(meta, b.order); PB_LAUNCH();
    k_unit_reach<<<cdiv(no, TB), TB, 0, s>>>(b.order, b.unit_grp[3], b.g_pat, b.g_W, b.slot_pose, meta, b.maxo); PB_LAUNCH();
    k_need_prefix<<<1, 32, 0, s>>>(b.maxo, N, b.top_need, b.bot_need); PB_LAUNCH();
  }
  PB_CUDA(cudaMemsetAsync(b.patch_track, 0xff, (size_t)NM * sizeof(int), s));
  k_patch_track<<<cdiv(cm, TB), TB, 0, s>>>(b.kx, meta, b.patch_track); PB_LAUNCH();
  view_pointers(pl);
  k_chunk_desc<<<cdiv(std::max(b.caps.units[0], 1), TB), TB, 0, s>>>(pl->v, meta, b.g_eoff, b.caps.est > 0 ? b.caps.est : ((int64_t)1 << 62), b.cdesc); PB_LAUNCH();
  // a capacity plan's E storage exists already: its row padding must read as zero (finite) whatever graph was there before
  if (pl->Est) PB_CUDA(cudaMemsetAsync(pl->Est, 0, ((size_t)pl->est_floats + 8) * sizeof(float), s));
  PB_CUDA(cudaMemcpyAsync(b.shape_host, meta, META_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (record) {                                                  // (a captured build records after the graph launch)
    PB_CUDA(cudaEventRecord(b.ev_shape, s));
    b.pending = 1;
  }
  return BA_OK;
}

// Host side of a finished build: wait for the shape block (the only wait of an update), fill in the counts the launch
// code needs, make sure the workspace covers the reduced system of this graph.
int plan_finalize(BaPlan *pl) {
  PlanBuild &b = pl->b;
  if (!b.pending) return b.valid ? BA_OK : BA_ERR_ARG;
  BA_CUDA(cudaEventSynchronize(b.ev_shape));
  b.pending = 0;
  b.valid = 0;
  const int *h = b.shape_host;
  if (h[META_ERR]) return BA_ERR_INDEX_RANGE;
  if (h[SH_OVERFLOW]) return BA_ERR_CAPACITY;
  PlanView &v = pl->v;
  v.E = h[SH_E]; v.m = h[SH_M]; v.G = h[SH_G];
  v.n_chunks = h[SH_CNT0]; v.n_units = h[SH_CNT1]; v.n_xchunks = h[SH_CNT2]; v.n_ounits = h[SH_CNT3];
  v.perm_identity = h[META_NOTIDENT] ? 0 : 1;
  v.dmax = h[META_DMAX]; v.e2_kp = h[SH_KP]; v.e2_tpl = h[SH_TPL];
  v.n_irregular = h[META_NIRREG]; v.dmax_irregular = h[META_DMAX_IRREG];
  BaPlanInfo &in = pl->info;
  in.n_edges = v.E; in.n_poses = v.N; in.n_patches = v.NM; in.n_total = h[META_MAXPOSE] + 1;
  in.n_tracks = v.m; in.n_groups = v.G; in.n_chunks = v.n_chunks; in.max_degree = h[META_DMAX];
  in.max_slots = h[META_WMAX]; in.block_bandwidth = h[META_SPAN]; in.perm_identity = v.perm_identity;
  pl->n_total_layout = in.n_total;
  pl->bwb_layout = in.block_bandwidth;
  pl->min_unit = h[META_MINUNIT]; pl->min_ounit = h[META_MINOUNIT]; pl->max_unit = h[SH_LEN1]; pl->max_ounit = h[SH_OMAX];
  const long long esize = ((long long)(unsigned)h[SH_ESIZE_LO]) | ((long long)h[SH_ESIZE_HI] << 32);
  cudaStream_t s = pl->mem_stream;
  if (!pl->Est) {                                                // exact plan: the E storage is sized now
    BA_CUDA(own(pl, &pl->Est, (size_t)esize + 8));
    BA_CUDA(cudaMemsetAsync(pl->Est, 0, ((size_t)esize + 8) * sizeof(float), s));   // row padding stays zero (finite) for ever
    pl->est_floats = esize;
    BA_CUDA(own(pl, &pl->Cw, b.caps.m)); BA_CUDA(own(pl, &pl->Qw, b.caps.m)); BA_CUDA(own(pl, &pl->dZ, b.caps.m));
    BA_CUDA(own(pl, &pl->status, 4));
    BA_CUDA(cudaMemsetAsync(pl->status, 0, 4 * sizeof(int), s));
  }
  pl->last_n = pl->last_fixedp = -1;
  pl->solve_shape_key = -1;
  if (alloc_workspace(pl) != BA_OK) return BA_ERR_CUDA;
  in.workspace_bytes = (int64_t)(pl->est_floats + 8) * 4 + (int64_t)b.caps.m * 20 + pl->sy_floats * 16;
  BA_CUDA(cudaStreamSynchronize(s));                             // (allocation-ordering stream; nothing is queued on it in steady state)
  b.valid = 1;
  return BA_OK;
}

static BaPlan *new_plan(int32_t N, int32_t NM, cudaStream_t s, int dev) {
  BaPlan *pl = new BaPlan();
  options_from_env(&pl->opt);
  pl->trace_buf = nullptr;
  pl->mem_stream = s;                      // allocations are ordered on the creation stream (used on it right away)
  pl->solve_stream = nullptr; pl->ev_step_begin = pl->ev_solved = nullptr; pl->epoch = 0; pl->solve_shape_key = -1;
  std::memset(&pl->info, 0, sizeof(pl->info));
  std::memset(&pl->v, 0, sizeof(pl->v));
  std::memset(&pl->b, 0, sizeof(pl->b));
  pl->b.tun = tuning_from_env();
  pl->v.N = N; pl->v.NM = NM;
  pl->device = dev;
  pl->SY = pl->L = pl->dX = pl->Wg = nullptr;
  pl->SY2 = nullptr; pl->sy_cur = 0; pl->sy_clean[0] = pl->sy_clean[1] = 0; pl->sy_untracked = 0; pl->sy_last = nullptr;
  pl->Est = pl->dZ = nullptr;
  pl->Cw = pl->Qw = nullptr;
  pl->status = nullptr;
  pl->sy_floats = 0; pl->est_floats = 0;
  pl->last_n = pl->last_fixedp = -1;
  pl->host_pipe = nullptr;
  pl->host_pipe_destroy = nullptr;
  pl->pp_buf[0] = pl->pp_buf[1] = nullptr;
  pl->timing = 0;
  pl->ev_mask = 0;
  for (auto &e : pl->ev) e = nullptr;
  return pl;
}

}  // namespace ba

using namespace ba;

static int plan_prologue(int32_t N, int32_t NM, cudaStream_t s, int *dev) {
  if (N <= 0 || NM <= 0) return BA_ERR_ARG;
  if (N > 65535) return BA_ERR_TOO_MANY_POSES;
  BA_CUDA(cudaGetDevice(dev));
  if (mem_pool_init(*dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "memory pool");
  return prepare_device(*dev, s);
}

extern "C" int ba_plan_create(const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t E,
                              int32_t N, int32_t NM, void *stream_, BaPlan **out) {
  if (!out) return BA_ERR_ARG;
  *out = nullptr;
  if (!ii || !jj || !kk || E <= 0 || E >= (int64_t)1 << 31) return BA_ERR_ARG;
  NvtxRange nvtx_range("ba:plan_create");
  cudaStream_t s = (cudaStream_t)stream_;
  int dev = 0;
  int rc = plan_prologue(N, NM, s, &dev);
  if (rc) return rc;
  BaPlan *pl = new_plan(N, NM, s, dev);
  const int want_trace = pl->opt.trace;
  pl->opt.trace = 0;
  if (want_trace) ba_plan_set_option(pl, BA_OPT_SOLVER_TRACE, want_trace);
  auto fail = [&](int code) { ba_plan_destroy(pl); return code; };
  const int m_up = (int)std::min<int64_t>(E, NM);
  if ((rc = alloc_front(pl, E, m_up, m_up, N, NM)) != BA_OK) return fail(rc);
  if ((rc = build_front(pl, ii, jj, kk, (int)E, nullptr, s)) != BA_OK) return fail(rc);
  int h[META_COUNT];
  if (cudaMemcpyAsync(h, pl->b.shape_dev, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
    return fail(set_cuda_error(cudaGetLastError(), "plan shape"));
  if (h[META_ERR]) return fail(BA_ERR_INDEX_RANGE);
  {
    // exact sizes for the back half; a group's tracks split into ceil(T / len) pieces of equal length: at most
    // m / len + G units of a kind
    int units[4];
    for (int k = 0; k < 4; ++k) units[k] = h[SH_M] / std::max(4, (h[SH_LEN0 + k] & ~3)) + 2 * h[SH_G] + 1;
    if ((rc = alloc_back(pl, h[SH_G], h[SH_PAT], units)) != BA_OK) return fail(rc);
  }
  if ((rc = build_back(pl, s, true)) != BA_OK) return fail(rc);
  if ((rc = plan_finalize(pl)) != BA_OK) return fail(rc);
  *out = pl;
  return BA_OK;
}

// A plan with room for any graph of up to cap_edges edges, cap_tracks distinct patches with edges, cap_groups pattern
// groups, cap_pattern pattern positions in total (sum over groups of the edges of one track) and cap_est floats of E
// storage (0: 6 * (cap_edges + 4 * cap_tracks)); it describes no graph until ba_plan_update has run.
extern "C" int ba_plan_create_capacity(int64_t cap_edges, int32_t cap_tracks, int32_t cap_groups, int32_t cap_pattern, int64_t cap_est,
                                       int32_t N, int32_t NM, void *stream_, BaPlan **out) {
  if (!out) return BA_ERR_ARG;
  *out = nullptr;
  if (cap_edges <= 0 || cap_edges >= (int64_t)1 << 31 || cap_tracks <= 0 || cap_groups <= 0 || cap_pattern <= 0) return BA_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream_;
  int dev = 0;
  int rc = plan_prologue(N, NM, s, &dev);
  if (rc) return rc;
  BaPlan *pl = new_plan(N, NM, s, dev);
  const int want_trace = pl->opt.trace;
  pl->opt.trace = 0;
  if (want_trace) ba_plan_set_option(pl, BA_OPT_SOLVER_TRACE, want_trace);
  auto fail = [&](int code) { ba_plan_destroy(pl); return code; };
  cap_tracks = (int32_t)std::min<int64_t>(cap_tracks, std::min<int64_t>(cap_edges, NM));
  cap_groups = std::min(cap_groups, cap_tracks);
  if ((rc = alloc_front(pl, cap_edges, cap_tracks, cap_groups, N, NM)) != BA_OK) return fail(rc);
  int units[4];
  for (int k = 0; k < 4; ++k) units[k] = cap_tracks / (k == 2 ? 32 : (k == 0 ? 8 : 16)) + 2 * cap_groups + 1;
  if ((rc = alloc_back(pl, cap_groups, cap_pattern, units)) != BA_OK) return fail(rc);
  const int64_t est = cap_est > 0 ? cap_est : 6 * (cap_edges + 4 * (int64_t)cap_tracks);
  pl->b.caps.est = est;
  if (cudaSuccess != own(pl, &pl->Est, (size_t)est + 8) || cudaSuccess != own(pl, &pl->Cw, cap_tracks) || cudaSuccess != own(pl, &pl->Qw, cap_tracks) ||
      cudaSuccess != own(pl, &pl->dZ, cap_tracks) || cudaSuccess != own(pl, &pl->status, 4) ||
      cudaSuccess != cudaMemsetAsync(pl->status, 0, 4 * sizeof(int), s) || cudaSuccess != cudaStreamSynchronize(s))
    return fail(set_cuda_error(cudaGetLastError(), "capacity plan"));
  pl->est_floats = est;
  *out = pl;
  return BA_OK;
}

// Re-derive a capacity plan for a new graph, on the device, without synchronising or allocating. n_edges bounds the
// edge count from above; n_edges_dev (optional, device memory) holds the live count when the graph itself is
// maintained on the device (ba_graph_*). The index arrays must stay valid until the plan has been used or finalized.
extern "C" int ba_plan_update(BaPlan *pl, const int64_t *ii, const int64_t *jj, const int64_t *kk, int64_t n_edges,
                              const int32_t *n_edges_dev, void *stream_) {
  if (!pl || !ii || !jj || !kk || n_edges <= 0 || !pl->b.caps.est) return BA_ERR_ARG;
  if (n_edges > pl->b.caps.E) return BA_ERR_CAPACITY;
  NvtxRange nvtx_range("ba:plan_update");
  cudaStream_t s = (cudaStream_t)stream_;
  PlanBuild &b = pl->b;
  b.valid = 0;
  // A caller that keeps its edge list in fixed device buffers (the live count in n_edges_dev) hands over the same
  // pointers every time: from the second such call on the whole derivation is ONE graph launch (captured once).
  const bool same = b.g_key[0] == ii && b.g_key[1] == jj && b.g_key[2] == kk && b.g_key[3] == n_edges_dev && b.g_n == n_edges;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  BA_CUDA(cudaStreamIsCapturing(s, &cs));
  if (same && cs == cudaStreamCaptureStatusNone) {
    // captured and replayed on a stream of the plan (the caller's may be the legacy default stream, which cannot be
    // captured), ordered behind / before the caller's stream by events
    if (!b.g_stream) {
      BA_CUDA(cudaStreamCreateWithFlags(&b.g_stream, cudaStreamNonBlocking));
      BA_CUDA(cudaEventCreateWithFlags(&b.g_ev_in, cudaEventDisableTiming));
    }
    if (!b.g_exec && !b.g_failed) {
      cudaGraph_t g = nullptr;
      int rc = BA_OK;
      if (cudaStreamBeginCapture(b.g_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = build_front(pl, ii, jj, kk, (int)n_edges, n_edges_dev, b.g_stream);
        if (!rc) rc = build_back(pl, b.g_stream, false);
        const cudaError_t ce = cudaStreamEndCapture(b.g_stream, &g);
        if (rc || ce != cudaSuccess || cudaGraphInstantiate(&b.g_exec, g, 0) != cudaSuccess) { b.g_exec = nullptr; b.g_failed = 1; (void)cudaGetLastError(); }
        if (g) cudaGraphDestroy(g);
      } else { b.g_failed = 1; (void)cudaGetLastError(); }
    }
    if (b.g_exec) {
      BA_CUDA(cudaEventRecord(b.g_ev_in, s));
      BA_CUDA(cudaStreamWaitEvent(b.g_stream, b.g_ev_in, 0));
      BA_CUDA(cudaGraphLaunch(b.g_exec, b.g_stream));
      g_launches.fetch_add(1);
      BA_CUDA(cudaEventRecord(b.ev_shape, b.g_stream));
      BA_CUDA(cudaStreamWaitEvent(s, b.ev_shape, 0));
      b.pending = 1;
      return BA_OK;
    }
  }
  if (!same) {
    if (b.g_exec) { cudaGraphExecDestroy(b.g_exec); b.g_exec = nullptr; }
    b.g_key[0] = ii; b.g_key[1] = jj; b.g_key[2] = kk; b.g_key[3] = n_edges_dev; b.g_n = n_edges; b.g_failed = 0;
  }
  int rc = build_front(pl, ii, jj, kk, (int)n_edges, n_edges_dev, s);
  if (rc) return rc;
  return build_back(pl, s, true);
}

extern "C" int ba_plan_finalize(BaPlan *pl) {
  if (!pl) return BA_ERR_ARG;
  return plan_finalize(pl);
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
  if (pl->b.shape_host) cudaFreeHost(pl->b.shape_host);
  if (pl->b.g_exec) cudaGraphExecDestroy(pl->b.g_exec);
  if (pl->b.g_stream) cudaStreamDestroy(pl->b.g_stream);
  if (pl->b.g_ev_in) cudaEventDestroy(pl->b.g_ev_in);
  if (pl->b.ev_shape) cudaEventDestroy(pl->b.ev_shape);
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
  if (int rc = plan_finalize(const_cast<BaPlan *>(pl))) return rc;
  *out = pl->info;
  return BA_OK;
}

extern "C" int ba_plan_set_layout(BaPlan *pl, int32_t n_total, int32_t bwb) {
  if (!pl) return BA_ERR_ARG;
  if (int rc = plan_finalize(pl)) return rc;
  if (n_total < pl->info.n_total || bwb < pl->info.block_bandwidth || n_total > pl->info.n_poses)
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
  if (int rc = plan_finalize(const_cast<BaPlan *>(pl))) return rc;
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
    case BA_ERR_CAPACITY: return "graph larger than the capacity of the plan";
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
