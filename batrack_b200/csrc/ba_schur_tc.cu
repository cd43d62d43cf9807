// ba_schur_tc.cu — per-track Schur complement (ba.py:311-322, block_matmul :52-58) on the 5th-generation tensor
// cores:  S -= sum_k Q_k E_k E_k^T,  y -= sum_k Q_k w_k E_k  for one unit (<= 256 consecutive tracks of one pattern
// group) per CTA, as ONE symmetric rank-K update in tcgen05 / TMEM:
//
//     X = [ E sqrt(Q) ; w sqrt(Q) ]   ((6 Wf + 1) x T, rows of FREE pose slots only, K-major: tracks contiguous)
//     D = X X^T                        D[r][n] = (E Q E^T)[r][n],   D[6 Wf][n] = (E Q w)[n]
//
// tcgen05 has no fp32 MMA kind, so X is split X = hi + lo with both parts rounded to TF32 (cvt.rna) and
//     D ~= hi hi^T + hi lo^T + lo hi^T                          (3xTF32: the dropped lo lo^T term is ~2^-24 relative)
// is issued as three tcgen05.mma.kind::tf32 per 8 tracks, A and B descriptors pointing at the SAME shared tiles.
//
// Warp-specialised, mbarrier-pipelined, no CTA barrier in the loop, two CTAs resident per SM (all 256 units of the
// headline graph are in flight at once):
//   * 8 converter warps: every thread owns four (operand row, 4-track granule) cells of a 32-track chunk. It loads
//     them straight from the entry-major E block with 16-byte loads (the next chunk's loads are in flight while this
//     one is converted), scales by sqrt(Q), splits hi / lo and stores both into the 128-byte-swizzled K-major UMMA
//     layout of stage c % 3, then fence.proxy.async + one mbarrier arrive per warp (`full`).
//   * 1 MMA warp (one thread): waits for `full`, issues the 12 MMAs of the chunk into the TMEM accumulator of the
//     chunk's span and commits to `empty` (stage reusable) — after the last chunk to `done`.
//   * fp32 accumulation in TMEM runs over one span (half of the unit, <= 128 tracks; 256 TMEM columns per CTA = two
//     128-column accumulators); the converter warps read both spans back with tcgen05.ld at the end, add them in fp64
//     and flush with fp64 atomics into the band (the subtraction B - E Q E^T cancels).
#include "ba_internal.h"

namespace ba {

namespace {

constexpr int kTcConvWarps = 8;                 // converter warps; also the epilogue: 4 TMEM lane quadrants x 2 column halves
constexpr int kTcThreads = 32 * (kTcConvWarps + 1);   // + the MMA warp
constexpr int kTcChunk = 32;                    // tracks per chunk: one 128-byte swizzle row of tf32
constexpr int kTcStages = 2;                    // operand stages (hi | lo tiles of one chunk)
constexpr int kTcRawStages = 3;                 // raw E chunks in flight (cp.async ring)
constexpr int kTcRawRows = 6 * kSchurTcMaxFree;                // 6 * 20: the w row is not staged (it comes with Q)
constexpr int kTcMaxFree = kSchurTcMaxFree;     // free slots per group: 6 * 20 + 1 = 121 operand rows
constexpr int kTcTile = 128 * 128;              // bytes of one operand tile: 128 rows x 128 B
constexpr int kTcOpBytes = kTcStages * 2 * kTcTile;     // [stage][hi | lo]
constexpr int kTcRawBytes = kTcRawStages * kTcRawRows * 128;
constexpr int kTcSmemBytes = kTcOpBytes + kTcRawBytes + 1024;   // 109 KB + alignment slack: two CTAs per SM
constexpr int kTcTmemCols = 256;                // two 128-column fp32 accumulators (one per span); 2 CTAs x 256 = the SM's 512

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// fire-and-forget reduction (REDG; atomicAdd here compiled to ATOMG with a discarded result: responses over the crossbar)
__device__ __forceinline__ void red_add_f64(double *addr, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// 16-byte cp.async with zero fill when `valid` is false (src-size 0)
__device__ __forceinline__ void cp_async16_zfill(unsigned dst, const void *src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte offset = 1024 B between
// 8-row groups in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
  return (unsigned long long)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float *v) {
  unsigned r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

__device__ __forceinline__ bool pose_free_tc(int pose, const CallView &c) {
  const int a = pose - c.fixedp;
  return a >= 0 && a < c.n;
}

}  // namespace

// CTA k runs unit order[k] (or k). A group whose free slots do not fit 128 rows is left to the SIMT kernel
// (k_schur skips the others when `tc_on`), so the two kernels partition the units between them.
__global__ void __launch_bounds__(kTcThreads, 2) k_schur_tc(PlanView pv, CallView cv, const int *__restrict__ ut0,
                                                            const int *__restrict__ ugrp, const int *__restrict__ order,
                                                            int *__restrict__ flags, int epoch, int acc_chunks, int min_tracks, long long *__restrict__ trace) {
  extern __shared__ unsigned char tc_smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[kTcStages], s_empty[kTcStages], s_done;
  __shared__ unsigned s_tmem;
  __shared__ int s_rowsrc[128];                  // E row (6 * slot + comp) of operand row r, -1: the w row, -2: unused
  __shared__ int s_off[128];                     // S / y offset 6 * (pose - fixedp) + comp of operand row r
  __shared__ int s_nfree;
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
  // optional phase stamps (BA_OPT_SOLVER_TRACE bit 1, tools/schur_trace.py): conversion warp 0 -> slots 0..5, MMA warp -> 8..13
#define TC_TR(k) do { if (trace && lane == 0 && (warp == 0 || warp == kTcConvWarps)) trace[(size_t)blockIdx.x * 16 + (warp ? 8 : 0) + (k)] = clock64(); } while (0)
  TC_TR(0);
  const int u = order ? order[blockIdx.x] : blockIdx.x;
  const int g = ugrp[u];
  const int t0 = ut0[u], t1 = ut0[u + 1];
  const int gt0 = pv.g_t0[g];
  const int W = pv.g_W[g];
  const int *slot_pose = pv.slot_pose + 2 * pv.g_pat[g];
  const float *Erows = cv.Est + pv.g_eoff[g] + (t0 - gt0);
  const int Ts = (pv.g_t0[g + 1] - gt0 + 3) & ~3;
  auto publish = [&]() {                          // streaming hand-over: this unit's atomics are visible
    if (flags) {
      __threadfence();
      __syncthreads();
      if (tau == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    }
  };
  if (t1 - t0 < min_tracks) return;               // the SIMT kernel takes this unit (and publishes its flag)
  // ---- set-up: TMEM (MMA warp), mbarriers (warp 1), slot scan (warp 0), one barrier. The operand tiles are not cleared:
  //      rows >= nrows are never written and their products land in accumulator rows / columns nobody reads ----
  if (warp == kTcConvWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTcTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (tau == 32) {
    for (int k = 0; k < kTcStages; ++k) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_full[k])), "r"(kTcConvWarps) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_empty[k])) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_done)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                                // free slots of the group -> operand rows (lane <-> slot)
    int nf = 0;
    for (int sb = 0; sb < W; sb += 32) {
      const int sl = sb + lane;
      const int pose = sl < W ? slot_pose[sl] : -1;
      const bool fr = sl < W && pose_free_tc(pose, cv);
      const unsigned mk = __ballot_sync(0xffffffffu, fr);
      const int rank = nf + __popc(mk & ((1u << lane) - 1u));
      if (fr && rank < kTcMaxFree) {
#pragma unroll
        for (int c = 0; c < 6; ++c) { s_rowsrc[6 * rank + c] = 6 * sl + c; s_off[6 * rank + c] = 6 * (pose - cv.fixedp) + c; }
      }
      nf += __popc(mk);
    }
    if (lane == 0) s_nfree = nf;
    if (nf <= kTcMaxFree)
      for (int r = 6 * nf + lane; r < 128; r += 32) { s_rowsrc[r] = r == 6 * nf ? -1 : -2; s_off[r] = 0; }
  }
  __syncthreads();
  const int nfree = s_nfree;
  if (nfree > kTcMaxFree || nfree == 0) {         // too many slots: the SIMT kernel takes the unit; none: nothing to add
    if (nfree == 0) publish();
    if (warp == kTcConvWarps) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(kTcTmemCols) : "memory");
    }
    return;
  }
  const int Rw = 6 * nfree, nrows = Rw + 1;       // the w row rides along as operand row Rw
  const int Nmma = (nrows + 15) & ~15;            // UMMA N (multiple of 16 for M = 128)

  unsigned char *op = reinterpret_cast<unsigned char *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);   // [stage][hi, lo][128 rows x 128 B]

  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;
  TC_TR(1);

  const int ntr = t1 - t0;
  const int nch = (ntr + kTcChunk - 1) / kTcChunk;
  // spans: chunks [0, half) accumulate in TMEM columns [0, 128), chunks [half, nch) in [128, 256)
  const int half = (nch >= 2 && acc_chunks < nch) ? (nch + 1) / 2 : nch;

  if (warp == kTcConvWarps) {
    // =============================== MMA warp (one thread issues) ===============================================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A / B tf32 (2 << 7, 2 << 10), both
      // K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(Nmma >> 3) << 17) | ((128u >> 4) << 24);
      for (int c = 0; c < nch; ++c) {
        const int st = c % kTcStages;
        mbar_wait(&s_full[st], (unsigned)(c / kTcStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned hi_a = smem_u32(op + (size_t)st * (2 * kTcTile)), lo_a = hi_a + kTcTile;
        const unsigned td = tmem + (c >= half ? 128u : 0u);
        const bool span_first = c == 0 || c == half;
#pragma unroll
        for (int ks = 0; ks < kTcChunk / 8; ++ks) {                    // UMMA K = 8 tf32 = 32 bytes inside the swizzle row
          const unsigned long long dh = umma_desc(hi_a + 32 * ks), dl = umma_desc(lo_a + 32 * ks);
          umma_tf32(td, dh, dh, idesc, (ks > 0 || !span_first) ? 1u : 0u);
          umma_tf32(td, dh, dl, idesc, 1u);
          umma_tf32(td, dl, dh, idesc, 1u);
        }
        umma_commit(&s_empty[st]);                                     // stage st is free again when these MMAs are done
      }
      umma_commit(&s_done);
      TC_TR(2);
    }
  } else {
    // =============================== converter warps ===========================================================
    const int j = tau & 7;                         // this thread's 16-byte granule (4 tracks) of a chunk row
    const int rg = tau >> 3;                       // rows rg, rg + 32, rg + 64, rg + 96
    unsigned char *rawb = op + kTcOpBytes;         // [kTcRawStages][kTcRawRows][128 B] raw E chunks, as in HBM
    const float *src[4];
    int kind[4];                                   // 0: E row, 1: the w row, 2: nothing
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = rg + 32 * k;
      const int rs = s_rowsrc[r];
      kind[k] = r < Rw ? 0 : (r == Rw ? 1 : 2);
      src[k] = Erows + (size_t)(rs >= 0 ? rs : 0) * Ts + 4 * j;
    }
    // Every thread stages exactly the cells it converts later, so cp.async.wait_group alone orders the data: no barrier.
    auto issue = [&](int c) {
      if (c < nch) {
        const bool valid = kTcChunk * c + 4 * j < ntr;     // E rows are padded with zeros up to a multiple of 4 tracks
        const unsigned dst = smem_u32(rawb + (size_t)(c % kTcRawStages) * (kTcRawRows * 128) + 16 * j);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (kind[k] == 0) cp_async16_zfill(dst + (rg + 32 * k) * 128, src[k] + kTcChunk * c, valid);
      }
      cp_commit();
    };
    const float2 *qsrc = cv.Qw + t0 + 4 * j;
    auto fetch_q = [&](int c, float4 &qa, float4 &qb) {    // (Q, w) of the thread's 4 tracks of chunk c
      float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
      if (c < nch) {
        const int tk = kTcChunk * c + 4 * j;
        const float2 *qs = qsrc + kTcChunk * c;
        if (tk + 3 < ntr) {
          const float4 a = __ldg(reinterpret_cast<const float4 *>(qs)), b = __ldg(reinterpret_cast<const float4 *>(qs + 2));
          q0 = make_float2(a.x, a.y); q1 = make_float2(a.z, a.w); q2 = make_float2(b.x, b.y); q3 = make_float2(b.z, b.w);
        } else {
          if (tk < ntr) q0 = __ldg(qs);
          if (tk + 1 < ntr) q1 = __ldg(qs + 1);
          if (tk + 2 < ntr) q2 = __ldg(qs + 2);
        }
      }
      qa = make_float4(q0.x, q0.y, q1.x, q1.y);
      qb = make_float4(q2.x, q2.y, q3.x, q3.y);
    };
    // One chunk: convert the thread's cells of chunk c (its (Q, w) in qa / qb, fetched two chunks ago), then fetch the
    // (Q, w) of chunk c + 3 into the same registers. Three static register sets rotate through the unrolled loop below,
    // so no register move ever waits for a load in flight.
    auto do_chunk = [&](int c, float4 &qa, float4 &qb) {
      const int st = c % kTcStages;
      issue(c + 2);
      const float s0 = sqrtf(qa.x), s1 = sqrtf(qa.z), s2 = sqrtf(qb.x), s3 = sqrtf(qb.z);
      const float4 wrow = make_float4(qa.y, qa.w, qb.y, qb.w);
      fetch_q(c + 3, qa, qb);
      cp_wait<2>();                                                    // this thread's cells of chunk c have landed
      const unsigned char *rs = rawb + (size_t)(c % kTcRawStages) * (kTcRawRows * 128) + 16 * j;
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = kind[k] == 0 ? *reinterpret_cast<const float4 *>(rs + (rg + 32 * k) * 128) : wrow;
      if (c >= kTcStages) mbar_wait(&s_empty[st], (unsigned)(c / kTcStages - 1) & 1u);   // the MMAs of chunk c - 2 are done
      unsigned char *hi = op + (size_t)st * (2 * kTcTile), *lo = hi + kTcTile;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kind[k] == 2) continue;
        const int r = rg + 32 * k;
        const float x0 = v[k].x * s0, x1 = v[k].y * s1, x2 = v[k].z * s2, x3 = v[k].w * s3;
        const float h0 = tf32_rna(x0), h1 = tf32_rna(x1), h2 = tf32_rna(x2), h3 = tf32_rna(x3);
        const int o = (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4);
        *reinterpret_cast<float4 *>(hi + o) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4 *>(lo + o) = make_float4(tf32_rna(x0 - h0), tf32_rna(x1 - h1), tf32_rna(x2 - h2), tf32_rna(x3 - h3));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[st])) : "memory");
    };
    float4 q0a, q0b, q1a, q1b, q2a, q2b;
    issue(0);
    issue(1);
    fetch_q(0, q0a, q0b);
    fetch_q(1, q1a, q1b);
    fetch_q(2, q2a, q2b);
    for (int c = 0; c < nch; c += 3) {
      do_chunk(c, q0a, q0b);
      if (c + 1 < nch) do_chunk(c + 1, q1a, q1b);
      if (c + 2 < nch) do_chunk(c + 2, q2a, q2b);
    }

    // ---- epilogue: both spans back from TMEM, added in fp64, flushed: S -= D (lower storage), y -= D[Rw][.]
    //      (ba.py:321-322). Warp (lq, h): TMEM lane quadrant lq (operand rows 32 lq ..), column blocks 2h, 2h + 1; only
    //      blocks of the lower triangle are read (the w row lies in the last occupied quadrant). tcgen05.ld hands a lane
    //      one accumulator ROW; the 32 x 32 block is transposed through shared memory (the operand stages are dead by
    //      now) so that a warp's 32 atomics of one instruction hit 32 CONSECUTIVE doubles of one row of S — 8 sectors
    //      per instruction instead of 32: the flush is bound by the L2's atomic sector operations ----
    TC_TR(2);
    mbar_wait(&s_done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every converter warp arrived on `full` for the last chunk before `done` could fire, so the operand / raw stages are
    // dead here; the named barrier states it in a form tools understand, too (racecheck does not follow mbarriers)
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kTcConvWarps) : "memory");
    TC_TR(3);
    const int lq = warp & 3, hcol = warp >> 2;
    const bool two = half < nch;
    double *tb = reinterpret_cast<double *>(op) + (size_t)warp * (32 * 33);   // this warp's 32 x 33 transposition buffer
#pragma unroll 1
    for (int cb = 2 * hcol; cb < 2 * hcol + 2; ++cb) {
      if (cb > lq || 32 * cb >= nrows || 32 * lq >= nrows) continue;   // warp-uniform
      const unsigned ta = tmem + ((unsigned)(32 * lq) << 16) + (unsigned)(32 * cb);
      {
        float v[32], v2[32];
        tmem_ld32(ta, v);
        if (two) tmem_ld32(ta + 128u, v2);
#pragma unroll
        for (int k = 0; k < 32; ++k) tb[lane * 33 + k] = two ? (double)v[k] + (double)v2[k] : (double)v[k];
      }
      __syncwarp();
      const int n = 32 * cb + lane;                                     // this lane's column of the block
      const int offn = s_off[n];
      const int rend = min(32, nrows - 32 * lq);
      for (int rr = 0; rr < rend; ++rr) {
        const int r = 32 * lq + rr;
        const double d = tb[rr * 33 + lane];
        if (r < Rw) { if (n <= r) red_add_f64(cv.S + (size_t)s_off[r] * cv.ld + cv.off + offn, -d); }
        else if (n < Rw) red_add_f64(cv.y + offn, -d);                    // r == Rw: the w row
      }
      __syncwarp();
    }
  }
  TC_TR(4);
  publish();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  TC_TR(5);
  if (warp == kTcConvWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
}

int schur_tc_prepare_device() {
  return cudaFuncSetAttribute(k_schur_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}

int launch_schur_tc(const PlanView &pv, const CallView &cv, int n_units, const int *ut0, const int *ugrp, const int *order,
                    int *flags, int epoch, int acc_chunks, int min_tracks, long long *trace, cudaStream_t s) {
  k_schur_tc<<<n_units, kTcThreads, kTcSmemBytes, s>>>(pv, cv, ut0, ugrp, order, flags, epoch, acc_chunks < 1 ? 1 : acc_chunks, min_tracks, trace);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
