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
// Accumulation in TMEM is fp32 and only runs over one chunk of 32 tracks (4 k-steps x 3 MMAs); every chunk's
// 128 x N accumulator is read back with tcgen05.ld and added into fp64 registers (the subtraction B - E Q E^T
// cancels), and the unit ends with one flush of fp64 atomics into the band, like the SIMT kernel.
// Pipeline per chunk c (all 256 threads): cp.async raw E rows of chunk c+2  ->  scale by sqrt(Q), split hi / lo,
// store in the 128-byte-swizzled K-major UMMA layout (stage c&1)  ->  one thread issues the 12 MMAs into TMEM buffer
// c&1 and commits to an mbarrier  ->  everyone reads buffer (c-1)&1 back while the tensor core works on chunk c.
#include "ba_internal.h"

namespace ba {

namespace {

constexpr int kTcThreads = 512;                 // 16 warps: 4 TMEM lane quadrants x 4 column blocks of 32 in the epilogue
constexpr int kTcChunk = 32;                    // tracks per chunk: one 128-byte swizzle row of tf32
constexpr int kTcRawStages = 3;
constexpr int kTcMaxFree = kSchurTcMaxFree;  // free slots per group: 6 * 21 + 1 = 127 rows <= 128
constexpr int kTcTile = 128 * 128;              // bytes of one operand tile: 128 rows x 128 B
constexpr int kTcOpBytes = 2 * 2 * kTcTile;     // [stage][hi | lo]
constexpr int kTcRawBytes = kTcRawStages * (128 * 128 + 32 * 8);   // raw E rows [128][32] floats + (Q, w) [32]
constexpr int kTcSmemBytes = kTcOpBytes + kTcRawBytes + 1024;      // + alignment slack

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// 16-byte cp.async with zero fill when `valid` is false (src-size 0)
__device__ __forceinline__ void cp_async16_zfill(void *dst, const void *src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0));
}
__device__ __forceinline__ void cp_async8_zfill(void *dst, const void *src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 8 : 0));
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
__global__ void __launch_bounds__(kTcThreads, 1) k_schur_tc(PlanView pv, CallView cv, const int *__restrict__ ut0,
                                                            const int *__restrict__ ugrp, const int *__restrict__ order,
                                                            int *__restrict__ flags, int epoch, int acc_chunks, int min_tracks) {
  extern __shared__ unsigned char tc_smem_raw[];
  __shared__ unsigned long long s_bar[2];
  __shared__ unsigned s_tmem;
  __shared__ int s_rowsrc[128];                  // E row (6 * slot + comp) of operand row r, -1: the w row, -2: unused
  __shared__ int s_off[128];                     // S / y offset 6 * (pose - fixedp) + comp of operand row r
  __shared__ int s_nfree;
  const int tau = threadIdx.x, lane = tau & 31, warp = tau >> 5;
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
  if (tau == 0) {                                 // free slots of the group -> operand rows
    int nf = 0;
    for (int s = 0; s < W; ++s) {
      const int pose = slot_pose[s];
      if (!pose_free_tc(pose, cv)) continue;
      if (nf < kTcMaxFree) {
        for (int c = 0; c < 6; ++c) { s_rowsrc[6 * nf + c] = 6 * s + c; s_off[6 * nf + c] = 6 * (pose - cv.fixedp) + c; }
      }
      ++nf;
    }
    s_nfree = nf;
    if (nf <= kTcMaxFree) {
      s_rowsrc[6 * nf] = -1; s_off[6 * nf] = 0;
      for (int r = 6 * nf + 1; r < 128; ++r) { s_rowsrc[r] = -2; s_off[r] = 0; }
    }
  }
  __syncthreads();
  const int nfree = s_nfree;
  if (nfree > kTcMaxFree || t1 - t0 < min_tracks) return;    // the SIMT kernel takes this unit (and publishes its flag)
  if (nfree == 0) { publish(); return; }
  const int Rw = 6 * nfree, nrows = Rw + 1;       // the w row rides along as operand row Rw
  const int Nmma = (nrows + 15) & ~15;            // UMMA N (multiple of 16 for M = 128)

  unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char *op = base;                                            // [2 stages][hi, lo][128 rows x 128 B]
  float *raw = reinterpret_cast<float *>(base + kTcOpBytes);           // [3][128][32]
  float2 *rawq = reinterpret_cast<float2 *>(raw + kTcRawStages * 128 * 32);   // [3][32]

  // ---- set-up: TMEM (2 accumulator buffers of 128 columns), mbarriers, zero rows >= nrows of the operand tiles ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tau == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int o = tau; o < kTcOpBytes / 16; o += kTcThreads) reinterpret_cast<float4 *>(op)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;

  const int ntr = t1 - t0;
  const int nch = (ntr + kTcChunk - 1) / kTcChunk;
  const int j = tau & 7;                           // this thread's 16-byte granule (4 tracks) within a chunk row
  auto issue_raw = [&](int c) {
    if (c < nch) {
      float *rs = raw + (size_t)(c % kTcRawStages) * (128 * 32);
      const int tk = kTcChunk * c + 4 * j;         // first track of the granule, relative to t0
      const bool valid = tk < ntr;
      for (int r = tau >> 3; r < Rw; r += kTcThreads / 8)
        cp_async16_zfill(rs + r * 32 + 4 * j, valid ? Erows + (size_t)s_rowsrc[r] * Ts + tk : Erows, valid);
      if (tau < 32) {
        const bool qv = kTcChunk * c + tau < ntr;
        cp_async8_zfill(rawq + (c % kTcRawStages) * 32 + tau, qv ? cv.Qw + t0 + kTcChunk * c + tau : cv.Qw, qv);
      }
    }
    cp_commit();
  };
  // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A / B tf32 (2 << 7, 2 << 10), both K-major,
  // N >> 3 at bit 17, M >> 4 at bit 24
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(Nmma >> 3) << 17) | ((128u >> 4) << 24);

  // epilogue role: TMEM lane quadrant (32 operand rows) x block of 32 columns; only blocks of the lower triangle
  // (and, through the w row, nothing else: the w row lies in the last occupied quadrant) are read back
  const int lq = warp & 3, cb = warp >> 2;
  const int row = 32 * lq + lane;
  const bool epi_active = cb <= lq && 32 * cb < nrows && 32 * lq < nrows;
  double acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.0;
  // TMEM accumulates `acc_chunks` chunks (fp32) before it is read back and added into the fp64 registers. MMAs complete
  // in issue order and every chunk commits to s_bar[c & 1]: waiting for chunk c's phase means chunks <= c are done.
  auto chunk_wait = [&](int c) { mbar_wait(&s_bar[c & 1], (unsigned)(c >> 1) & 1u); };
  auto read_back = [&](int c_last) {              // c_last: last chunk of an accumulation span (already waited for)
    const int b = (c_last / acc_chunks) & 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (epi_active) {
      const unsigned ta = tmem + ((unsigned)(32 * lq) << 16) + (unsigned)(b * 128 + 32 * cb);
      float v[32];
      tmem_ld32(ta, v);
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] += (double)v[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  issue_raw(0);
  issue_raw(1);
  for (int c = 0; c < nch; ++c) {
    issue_raw(c + 2);
    cp_wait<2>();
    __syncthreads();                               // raw chunk c landed for everyone
    {
      // ---- scale by sqrt(Q), split into TF32 hi / lo, store in the swizzled K-major operand layout ----
      const float *rs = raw + (size_t)(c % kTcRawStages) * (128 * 32);
      const float4 qa = *reinterpret_cast<const float4 *>(rawq + (c % kTcRawStages) * 32 + 4 * j);       // (Q, w) of 2 tracks
      const float4 qb = *reinterpret_cast<const float4 *>(rawq + (c % kTcRawStages) * 32 + 4 * j + 2);
      const float s0 = sqrtf(qa.x), s1 = sqrtf(qa.z), s2 = sqrtf(qb.x), s3 = sqrtf(qb.z);
      unsigned char *hi = op + (size_t)(c & 1) * (2 * kTcTile), *lo = hi + kTcTile;
      for (int r = tau >> 3; r < nrows; r += kTcThreads / 8) {
        float4 v;
        if (r < Rw) v = *reinterpret_cast<const float4 *>(rs + r * 32 + 4 * j);
        else v = make_float4(qa.y, qa.w, qb.y, qb.w);                  // the w row
        const float x0 = v.x * s0, x1 = v.y * s1, x2 = v.z * s2, x3 = v.w * s3;
        const float h0 = tf32_rna(x0), h1 = tf32_rna(x1), h2 = tf32_rna(x2), h3 = tf32_rna(x3);
        const int o = (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4);
        *reinterpret_cast<float4 *>(hi + o) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4 *>(lo + o) = make_float4(tf32_rna(x0 - h0), tf32_rna(x1 - h1), tf32_rna(x2 - h2), tf32_rna(x3 - h3));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tau == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned hi_a = smem_u32(op + (size_t)(c & 1) * (2 * kTcTile)), lo_a = hi_a + kTcTile;
      const unsigned td = tmem + (unsigned)(((c / acc_chunks) & 1) * 128);
      const bool span_first = c % acc_chunks == 0;
#pragma unroll
      for (int ks = 0; ks < kTcChunk / 8; ++ks) {                      // UMMA K = 8 tf32 = 32 bytes inside the swizzle row
        const unsigned long long dh = umma_desc(hi_a + 32 * ks), dl = umma_desc(lo_a + 32 * ks);
        umma_tf32(td, dh, dh, idesc, (ks > 0 || !span_first) ? 1u : 0u);
        umma_tf32(td, dh, dl, idesc, 1u);
        umma_tf32(td, dl, dh, idesc, 1u);
      }
      umma_commit(&s_bar[c & 1]);
    }
    if (c >= 1) {                                                      // MMA(c-1) done: its operand stage is free again,
      chunk_wait(c - 1);                                               // and if it closed a span, the span is read back
      if (c % acc_chunks == 0) read_back(c - 1);                       // while the tensor core works on chunk c
    }
  }
  chunk_wait(nch - 1);
  read_back(nch - 1);

  // ---- flush: S -= D (lower storage), y -= D[Rw][.]  (ba.py:321-322) ----
  if (epi_active) {
    if (row < Rw) {
      double *Srow = cv.S + (size_t)s_off[row] * cv.ld + cv.off;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int n = 32 * cb + k;
        if (n <= row) atomicAdd(Srow + s_off[n], -acc[k]);
      }
    } else if (row == Rw) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int n = 32 * cb + k;
        if (n < Rw) atomicAdd(cv.y + s_off[n], -acc[k]);
      }
    }
  }
  publish();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int schur_tc_prepare_device() {
  return cudaFuncSetAttribute(k_schur_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}

int launch_schur_tc(const PlanView &pv, const CallView &cv, int n_units, const int *ut0, const int *ugrp, const int *order,
                    int *flags, int epoch, int acc_chunks, int min_tracks, cudaStream_t s) {
  k_schur_tc<<<n_units, kTcThreads, kTcSmemBytes, s>>>(pv, cv, ut0, ugrp, order, flags, epoch, acc_chunks < 1 ? 1 : acc_chunks, min_tracks);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

}  // namespace ba
