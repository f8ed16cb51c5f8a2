// tc_gemm.cuh -- tcgen05 (5th-gen tensor core) 3xTF32 GEMMs for the hidden Linear layers (sm_100a).
//
// fp32 accuracy on the TF32 tensor pipe: every fp32 operand x is split once into
//     x_hi = tf32(x)            (cvt.rna, low 13 mantissa bits zero)
//     x_lo = tf32(x - x_hi)     (exact difference, then rounded)
// and  A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  is accumulated in fp32 in tensor memory (TMEM)
// by three tcgen05.mma.kind::tf32 per K-step (the dropped lo*lo term is ~2^-22 relative).
//
// Orientation ("features on lanes"): D[feature, row] = W[feature, :] . X[row, :]
//     A operand = W   [128 out-features x K]   K-major, resident in shared memory for the whole kernel
//     B operand = X   [TN rows          x K]   K-major, streamed tile by tile (rows = point*jet columns)
//     D         in TMEM: lane = out-feature, column = row of the tile
// so an epilogue thread owns one feature and sees all jet columns of a point in consecutive TMEM
// columns, and a warp's global/shared accesses for one row are 128 contiguous bytes.
//
// Shared-memory operand layout = the canonical UMMA K-major SWIZZLE_128B layout: the K axis is cut into
// 32-element (128 B) blocks; inside a block rows are 128 B apart, 8-row groups 1024 B apart, and the
// 16-byte chunk index of a row is XORed with (row % 8).
//
// Warp roles (1 CTA per SM, persistent over row tiles):
//     NLW warps  loaders : global fp32 -> hi/lo split in registers -> swizzled smem stage (register prefetch ring)
//     4 warps    epilogue: TMEM -> registers (tcgen05.ld) -> (+bias) -> global
//     1 warp     MMA     : one elected thread issues tcgen05.mma / tcgen05.commit; owns TMEM alloc
// Pipelines: smem full/empty mbarriers (loaders <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "jet_math.cuh"
#include "tc_api.h"

namespace pinnk {

namespace tc {

constexpr int kEpiThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive on a "slot may be refilled" barrier AFTER the shared-memory loads that read the slot have completed.  `dep` must be
// computed from one destination register of EVERY such load; storing it to a scratch word makes the store -- and the arrive
// behind it -- wait for the loads' data (an unused asm operand does not: ptxas drops it).  A plain arrive only follows the
// loads' ISSUE: they can still sit in the load/store queue when the TMA warp sees the slot free and refills it, and the late
// ones then read the NEXT tile's rows (profiles/r02_stale_tile_rows.md).
__device__ __forceinline__ void mbar_arrive_after_loads(uint64_t* bar, uint32_t dep, uint32_t* scratch) {
  asm volatile("st.shared.u32 [%1], %2;\n\tmbarrier.arrive.shared::cta.b64 _, [%0];"
               ::"r"(smem_u32(bar)), "r"(smem_u32(scratch)), "r"(dep) : "memory");
}
// Bounded wait: a protocol bug must trap (error returned to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  // the suspend-time hint lets the hardware park the warp until the phase completes instead of re-polling every few
  // hundred cycles (polling warps took ~25 % of the issue slots of the rows kernels)
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) return;
  }
  __trap();
}
// Wait used by the (many) epilogue / flush warps: they have slack (double-buffered accumulators), so after a failed
// probe they sleep instead of re-polling -- polling warps were stealing ~30 % of the issue slots from the convert warps
// that share their scheduler.
#ifndef PINNK_EPI_SLEEP
#define PINNK_EPI_SLEEP 64      // ns between probes (A/B: PINNK_NVCC_EXTRA=-DPINNK_EPI_SLEEP=...; 0 = hinted try_wait instead)
#endif
// Wait of the wgrad flush warps: a segment (SEG tiles, ~3 us) separates two flushes and the accumulators are double-buffered, so
// waking up to half a microsecond late costs nothing, while probing every ~100 ns made these four mostly idle warps issue a
// fifth of the kernel's instructions (152 probes per tile and SM; profiles/r02_final.md) on schedulers the convert warps need.
__device__ __forceinline__ void mbar_wait_lazy(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(512);
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
#if PINNK_EPI_SLEEP == 0
  mbar_wait(bar, parity);
  return;
#endif
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(PINNK_EPI_SLEEP);
  }
  __trap();
}
// One lane of a converged warp (elect.sync).  The single-thread issuers (tcgen05.mma / commit, TMA) branch on this and
// not on `lane == 0`: under a lane-id branch the compiler treats the region as divergent and wraps EVERY uniform-datapath
// instruction (UTCHMMA, UBLKCP ...) in an ELECT / BRA.U.ANY serialisation loop, ~50 cycles per MMA -- the MMA issue, not
// the tensor pipe, was what bounded the rows kernels (2600 issue cycles per 64-row tile against 1536 of tensor time).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {                        // one thread
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive TMEM columns of this thread's lane, no wait (caller issues tmem_wait_ld once for a batch of loads)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.wait::ld that NAMES the registers the pending loads write.  The destination registers of a tcgen05.ld are not valid
// until the wait, but to the compiler they are ordinary asm outputs, ready as soon as the load statement has been issued: under
// register pressure it spilled them to local memory between the load and the wait (the sin epilogue of the 256-wide layers)
// and read back whatever the registers held before the data arrived -- a handful of wrong elements per launch, different
// from run to run (profiles/r02_stale_tile_rows.md).  As "+r" operands of the wait they cannot be touched before it.
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[32], uint32_t (&b)[32]) {
  tmem_wait_ld(a);
  tmem_wait_ld(b);
}

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_128B: SBO = 1024 B (one 8-row group), LBO unused (1)
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);     // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
  lo.x = to_tf32(v.x - hi.x); lo.y = to_tf32(v.y - hi.y); lo.z = to_tf32(v.z - hi.z); lo.w = to_tf32(v.w - hi.w);
}
// byte offset of 16-byte chunk `chunk` (along K) of row `row` inside a K-major SW128 operand with `rows` rows
__device__ __forceinline__ uint32_t sw128_offset(int rows, int row, int chunk) {
  const int kb = chunk >> 3, cj = chunk & 7;
  return (uint32_t)(kb * rows * 128 + (row >> 3) * 1024 + (row & 7) * 128 + ((cj ^ (row & 7)) << 4));
}

// UMMA shared-memory matrix descriptor, MN-major tf32.  32-bit MN-major operands must use the SWIZZLE_128B_BASE32B
// layout (type 1): an atom is 4 K-rows x 128 B (32 floats along M/N), rows 128 B apart, the 32-byte chunk index of
// a row XORed with (row % 4).  Further 32-float groups along M/N are LBO bytes apart, further 4-row groups along K are
// SBO bytes apart; one kind::tf32 MMA (K = 8) consumes two 4-row groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                          // layout type: SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {   // both operands MN-major
  return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// "features on lanes" GEMM over row tiles (forward and dgrad share it):
//   TRANS_W = false:  Y[M, ldy] (cols n0..n0+127) = X[M, K] * W[n0.., 0..K)^T  (+ bias on rows with row % jet_cols == 0)
//                     W is [n_out, K] row-major (ldw = K)                      -- nn.Linear forward
//   TRANS_W = true :  Y[M, ldy] (cols n0..n0+127) = X[M, K] * W[0..K, n0..)    -- dgrad: X = dL/dZ [M, out=K], W is
//                     [out=K, in] row-major (ldw = in), Y = dL/dX [M, in]
// NLW loader warps (global fp32 -> hi/lo -> swizzled smem, PF tiles of register prefetch), 4 epilogue warps, 1 MMA warp.
template <int K, int TN, int STAGES, int ACC, int NLW, int PF, bool TRANS_W>
__global__ void __launch_bounds__((NLW + 4 * (TN / 16) + 1) * 32, 1)
linear_rows_kernel(const float* __restrict__ X, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                   float* __restrict__ Y, int64_t M, int ldy, int jet_cols) {
  static_assert(K % 32 == 0 && TN % 16 == 0 && TN <= 256 && (TN % NLW) == 0, "tile shape");
  constexpr int KB = K / 32;                      // 128-byte K blocks
  constexpr int CHUNKS = K / 4;                   // 16-byte chunks per row
  constexpr uint32_t W_BYTES = 128 * K * 4;       // one of W_hi / W_lo
  constexpr uint32_t X_BYTES = TN * K * 4;        // one of X_hi / X_lo per stage
  constexpr int EH = TN / 16;                     // epilogue warps per TMEM lane quarter: 16 tile rows each
  constexpr int NEW = 4 * EH;
  constexpr int NTHREADS = (NLW + NEW + 1) * 32;
  constexpr int EPI0 = NLW, MMAW = NLW + NEW;
  static_assert(EPI0 % 4 == 0, "epilogue warps must start at a multiple of 4 (TMEM lane quarters)");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_hi = smem;
  uint8_t* w_lo = smem + W_BYTES;
  uint8_t* x_st = smem + 2 * W_BYTES;             // [STAGES][hi|lo][X_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_st + (size_t)STAGES * 2 * X_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * 128;
  const int64_t ntiles = (M + TN - 1) / TN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NLW); mbar_init(&empty[s], 1); }     // one arrive per warp
    for (int b = 0; b < ACC; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], NEW); }
    fence_mbar_init();
  }
  // TMEM: per accumulator buffer KB "main" partials (one per 32-wide K block, 4 accumulate steps each) plus one
  // "correction" partial (the lo*hi + hi*lo terms).  The tensor core rounds every accumulate step toward zero, so
  // short accumulation chains summed afterwards in registers (round-to-nearest) keep the result at fp32 quality.
  constexpr int BUF_COLS = TN * (KB + 1);
  constexpr uint32_t TMEM_COLS = (ACC * BUF_COLS <= 32) ? 32 : (ACC * BUF_COLS <= 64) ? 64 : (ACC * BUF_COLS <= 128) ? 128
                               : (ACC * BUF_COLS <= 256) ? 256 : 512;
  static_assert(ACC * BUF_COLS <= 512, "TMEM budget");
  if (warp == MMAW) tmem_alloc(tmem_slot, TMEM_COLS);
  // resident weights: split to hi/lo and store in the swizzled A-operand layout (row = output feature of this kernel)
  if (!TRANS_W) {
    for (int idx = threadIdx.x; idx < 128 * CHUNKS; idx += NTHREADS) {
      const int row = idx / CHUNKS, chunk = idx - row * CHUNKS;
      const float4 v = *reinterpret_cast<const float4*>(W + (int64_t)(n0 + row) * ldw + chunk * 4);
      float4 hi, lo;
      split4(v, hi, lo);
      const uint32_t off = sw128_offset(128, row, chunk);
      *reinterpret_cast<float4*>(w_hi + off) = hi;
      *reinterpret_cast<float4*>(w_lo + off) = lo;
    }
  } else {
    for (int idx = threadIdx.x; idx < K * 32; idx += NTHREADS) {        // 32 float4 per W row segment [k, n0..n0+127]
      const int k = idx >> 5, c4 = idx & 31;
      const float4 v = *reinterpret_cast<const float4*>(W + (int64_t)k * ldw + n0 + c4 * 4);
      float4 hi, lo;
      split4(v, hi, lo);
      const float h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t off = sw128_offset(128, c4 * 4 + e, k >> 2) + (uint32_t)(k & 3) * 4;
        *reinterpret_cast<float*>(w_hi + off) = h[e];
        *reinterpret_cast<float*>(w_lo + off) = l[e];
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NLW) {
    // ===================== loaders =====================
    constexpr int PER_ROW = (CHUNKS + 31) / 32;          // chunks per lane per row
    constexpr int RPW = TN / NLW;                        // rows per warp per tile
    float4 v[PF][RPW][PER_ROW];
    auto issue = [&](int64_t tile, float4 (&dst)[RPW][PER_ROW]) {
      const int64_t r0 = tile * TN;
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int64_t row = r0 + warp + NLW * i;
#pragma unroll
        for (int c = 0; c < PER_ROW; ++c) {
          const int chunk = lane + 32 * c;
          dst[i][c] = (tile < ntiles && row < M && chunk < CHUNKS )
                          ? __ldg(reinterpret_cast<const float4*>(X + row * K + chunk * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    const int64_t stride = gridDim.x;
    int64_t tile = blockIdx.x;
#pragma unroll
    for (int d = 0; d < PF; ++d) issue(tile + d * stride, v[d]);
    int it = 0;
    while (tile < ntiles) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        if (tile < ntiles) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* xh = x_st + (size_t)s * 2 * X_BYTES;
          uint8_t* xl = xh + X_BYTES;
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            const int r = warp + NLW * i;
#pragma unroll
            for (int c = 0; c < PER_ROW; ++c) {
              const int chunk = lane + 32 * c;
              if (chunk < CHUNKS) {
                float4 hi, lo;
                split4(v[d][i][c], hi, lo);
                const uint32_t off = sw128_offset(TN, r, chunk);
                *reinterpret_cast<float4*>(xh + off) = hi;
                *reinterpret_cast<float4*>(xl + off) = lo;
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
          issue(tile + PF * stride, v[d]);
          tile += stride;
          ++it;
        }
      }
    }
  } else if (warp < MMAW) {
    // ===================== epilogue: warp (q, h) owns lanes 32q.. and tile rows 16h..16h+15 =====================
    const int e = warp - EPI0;
    const int q = e & 3, h = e >> 2;
    const int f = q * 32 + lane;
    const float bf = (!TRANS_W && bias) ? bias[n0 + f] : 0.f;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it % ACC;
      const uint32_t ph = (uint32_t)(it / ACC) & 1u;
      const int64_t r0 = tile * TN + h * 16;
      // which of my 16 rows are value-column rows (bias applies): bit j set <=> (r0 + j) % jet_cols == 0
      uint32_t vmask = 0;
      if (!TRANS_W && bias) {
        int cj = (int)((uint32_t)r0 % (uint32_t)jet_cols);     // row counts of a chunk fit 32 bits
#pragma unroll
        for (int j = 0; j < 16; ++j) { vmask |= (cj == 0 ? 1u : 0u) << j; cj = (cj + 1 == jet_cols) ? 0 : cj + 1; }
      }
      float* yp = Y + r0 * ldy + n0 + f;
      mbar_wait(&tfull[b], ph);
      tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BUF_COLS + h * 16);
      uint32_t part[KB + 1][16];
#pragma unroll
      for (int kb = 0; kb <= KB; ++kb) tmem_ld16_nowait(tb + kb * TN, part[kb]);
#pragma unroll
      for (int kb = 0; kb <= KB; ++kb) tmem_wait_ld(part[kb]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
      const int nrows = (M - r0 >= 16) ? 16 : (int)(M - r0 > 0 ? M - r0 : 0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v = __uint_as_float(part[KB][j]);               // correction partial first (smallest magnitude)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) v += __uint_as_float(part[kb][j]);
        if ((vmask >> j) & 1u) v += bf;
        if (j < nrows) yp[(int64_t)j * ldy] = v;
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(128, TN);
      // descriptors differ only in the 14-bit start-address field (16-byte units): build the constant part once
      const uint64_t dconst = make_desc_k_sw128(0);
      const uint32_t wh = smem_u32(w_hi) >> 4, wl = smem_u32(w_lo) >> 4;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES, b = it % ACC;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u, bph = (uint32_t)(it / ACC) & 1u;
        mbar_wait(&tempty[b], bph ^ 1u);
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t xh = smem_u32(x_st + (size_t)s * 2 * X_BYTES) >> 4, xl = xh + (X_BYTES >> 4);
        const uint32_t d = tmem_base + (uint32_t)(b * BUF_COLS);
        const uint32_t d_corr = d + (uint32_t)(KB * TN);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ao = (uint32_t)((kb * 128 * 128 + ks * 32) >> 4), bo = (uint32_t)((kb * TN * 128 + ks * 32) >> 4);
            const uint64_t a_hi = dconst | (uint64_t)(wh + ao), a_lo = dconst | (uint64_t)(wl + ao);
            const uint64_t b_hi = dconst | (uint64_t)(xh + bo), b_lo = dconst | (uint64_t)(xl + bo);
            umma_tf32(d_corr, a_lo, b_hi, idesc, (kb | ks) ? 1u : 0u);
            umma_tf32(d_corr, a_hi, b_lo, idesc, 1);
            umma_tf32(d + (uint32_t)(kb * TN), a_hi, b_hi, idesc, ks ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);     // smem stage may be refilled once these MMAs have read it
        umma_commit(&tfull[b]);     // accumulator complete
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMAW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int K, int TN, int STAGES, int ACC, int NLW, int PF, bool TRANS_W>
static int launch_linear_rows(const float* X, const float* W, int ldw, const float* bias, float* Y, int64_t M, int n_cols,
                              int jet_cols, int sm_count, cudaStream_t st) {
  constexpr size_t smem = 1024 + 2 * (size_t)128 * K * 4 + (size_t)STAGES * 2 * TN * K * 4 + (2 * STAGES + 2 * ACC) * 8 + 16;
  static_assert(smem <= 232448, "shared memory budget (227 KB per CTA)");
  auto kern = linear_rows_kernel<K, TN, STAGES, ACC, NLW, PF, TRANS_W>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  const int64_t ntiles = (M + TN - 1) / TN;
  const int per_y = n_cols / 128;
  int gx = sm_count / per_y;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > ntiles) gx = (int)ntiles;
  dim3 grid((unsigned)gx, (unsigned)per_y, 1);
  kern<<<grid, (NLW + 4 * (TN / 16) + 1) * 32, smem, st>>>(X, W, ldw, bias, Y, M, n_cols, jet_cols);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// TS variant of the rows kernel for K = 128: the resident weight operand lives in TENSOR MEMORY (A-from-TMEM,
// tcgen05.mma [d], [a], bdesc), which (1) removes the 4 KB shared-memory read of A that every SS-mode MMA pays and
// (2) frees 128 KB of shared memory, so row tiles are 64 wide (half the MMA instructions per row) and 3 deep.
//   TMEM columns: [0,128) W_hi | [128,256) W_lo | 2 accumulator buffers x {main 64, correction 64}
// fp32 -> tf32 hi/lo splitting uses integer round-to-nearest (2 ALU ops per conversion).
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// round-to-nearest (ties away) fp32 -> tf32 on the bit pattern: what cvt.rna.tf32.f32 computes for finite inputs
__device__ __forceinline__ uint32_t rn_tf32_bits(uint32_t b) { return (b + 0x1000u) & 0xFFFFE000u; }
// (leaving the low 13 bits of lo in place -- the tensor core drops them -- would save one AND per element; measured: same
// GEMM error, no measurable speed-up, so the explicit mask stays)
__device__ __forceinline__ void split_bits(float v, uint32_t& hi, uint32_t& lo) {
  hi = rn_tf32_bits(__float_as_uint(v));
  lo = rn_tf32_bits(__float_as_uint(v - __uint_as_float(hi)));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// TMA bulk copy (cp.async.bulk, SASS UBLKCP): contiguous global bytes -> shared memory, completion signalled on an
// mbarrier by transaction bytes.  Issued by one thread; no register staging, no generic-proxy fence on the load side.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// TMA tiled copy (cp.async.bulk.tensor.2d, SASS UTMALDG): one instruction moves a [box rows x 128 floats] tile of a row-major
// matrix whose rows are wider than the tile (one K / N half of a 256-wide layer: 512-byte segments at a 1 KB stride) into
// dense shared memory.  Rows past the end of the matrix are zero-filled and still counted in the transaction bytes.
// (Before: one 512-byte cp.async.bulk per row from a single thread, ~50 cycles of issue each -- 3200 cycles per 64-row
// tile, more than the tile's HBM time, bounded every 256-wide kernel.)
__device__ __forceinline__ void tma_tile_2d(uint32_t dst_smem, const CUtensorMap* tm, int col0, int row0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tm)), "r"(col0), "r"(row0), "r"(smem_u32(bar))
               : "memory");
}
// Host: tensor map of a row-major fp32 matrix [rows, cols] with row stride ld floats, box = [box_rows x 128].
// The driver entry point is looked up through the runtime (no link against libcuda).
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline bool make_tmap_rows(CUtensorMap* tm, const float* base, int64_t rows, int cols, int ld, int box_rows) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || p == nullptr) return false;
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- thread-block cluster helpers (the paired reverse kernel: one CTA pair per tile stream)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// mbarrier traffic between the two CTAs of a cluster (the K-split rows kernel): arrive on a barrier in the PARTNER's shared
// memory with release semantics at cluster scope (everything this thread -- and, through bar.warp.sync, its warp -- wrote
// before is visible to a thread that then observes the phase with an acquire at cluster scope), and the matching wait.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
// "slot consumed" signal: orders nothing but itself (the caller has made sure its loads of the slot have delivered); a
// release here would put a memory barrier -- behind the warp's own output stores -- on every epilogue warp's path per tile
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(64);
  }
  __trap();
}
// Lock-step throttle between the two CTAs of a pair that walk the SAME tile stream (one does dgrad + adjoint, the other
// wgrad): each publishes the number of macro tiles whose loads it has issued into the partner's shared memory (DSMEM
// store) and never runs more than kPairLead macro tiles ahead of the partner, so that the second reader of a tile finds
// it in L2 -- the tile pair (dZ, Y) comes out of HBM once instead of twice.
struct PairSync {
  uint32_t my_word = 0;        // local shared address of the word the PARTNER writes (its issued macro tiles)
  uint32_t peer_word = 0;      // shared::cluster address of the partner's word (where this CTA publishes)
  uint32_t lead = 4;           // macro tiles a CTA may run ahead of its partner (PINNK_PAIR_LEAD)
  __device__ __forceinline__ void wait_turn(uint32_t n) const {      // before issuing the loads of macro tile n
    const uint32_t kPairLead = lead;
    if (n < kPairLead) return;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
      uint32_t v;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(my_word) : "memory");
      if (v + kPairLead > n) return;
      __nanosleep(40);
    }
    __trap();
  }
  __device__ __forceinline__ void publish(uint32_t issued) const {
    asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(peer_word), "r"(issued) : "memory");
  }
};

// Optional stage timers (build with -DPINNK_STAGE_TIMERS): block 0 accumulates, per role, the cycles spent waiting on
// each pipeline barrier; read back with pinnk_debug_stage_timers().  Slots: 0 tma:raw_empty  1 cvt:raw_full  2 cvt:empty
// 3 cvt:work  4 mma:tempty  5 mma:full  6 mma:issue  7 epi:tfull  8 epi:work  9 tiles  10 total cycles of block 0
#ifdef PINNK_STAGE_TIMERS
static __device__ unsigned long long g_stage_timers[16];
#define PK_T0() const long long _t0 = clock64()
#define PK_TACC(var) var += clock64() - _t0
#else
#define PK_T0()
#define PK_TACC(var)
#endif

// EPI selects the epilogue:
//   EPI_PLAIN   Y = acc (+ bias on value rows)
//   EPI_ACT     forward Linear + activation jets: Y = Z = acc + bias (the stash the reverse pass needs) and
//               Yact = act(Z) jets (the next layer's operand)              -- replaces gemm + act_fwd_kernel
//   EPI_ACTBWD  dgrad + activation adjoint: acc = dL/d(act output); Zs = stashed pre-activation jets of that activation;
//               Y = dL/d(pre-activation)                                   -- replaces gemm + act_bwd_kernel
// For the fused epilogues the jet layout is a compile-time (K0, K1): value column, K0 columns of direction 0, K1 of
// direction 1, with C = 1 + K0 + K1 dividing 32 so that every epilogue warp owns whole points.
//   EPI_ACTBWD_Y  as EPI_ACTBWD for tanh, but Zs holds the activation OUTPUT jets (the forward pass then never writes the
//               pre-activations: 512 B per row less traffic); z jets are recovered with tanh_dir_recover
//   EPI_PARTIAL producer half of the K-split kernel (below): the epilogue only moves its partial product to the pair's ring
enum { EPI_PLAIN = 0, EPI_ACT = 1, EPI_ACTBWD = 2, EPI_ACTBWD_Y = 3, EPI_FIRSTBWD = 4, EPI_PARTIAL = 5 };
constexpr int kKsRing = 8;                 // ring slots ([64 rows x 128 features] fp32 partial products) per CTA pair

// rows (tile columns) one epilogue warp owns for a jet column count JC: the largest multiple of JC that is <= 16 and a
// quarter of a 48-, 60- or 64-row tile
__host__ __device__ constexpr int jets_rows_per_warp(int jc) {
  return (jc >= 1 && 16 % jc == 0) ? 16 : (jc == 3 || jc == 5) ? 15 : (jc == 6) ? 12 : 0;
}

// EPI_FIRSTBWD: dgrad of the first HIDDEN layer fused with the whole reverse of the network's input layer
// (nn.Linear(in_dim <= 4, width) + activation): the epilogue holds dL/dY0 jets of feature f, recomputes the input layer's
// pre-activation jets from (x, t) -- z0 = b0[f] + W0[f,:] . xt, first-order coefficient W0[f,:] . vec_d, higher orders 0 --
// runs the activation adjoint and accumulates dW0[f,:], db0[f] in registers over all tiles of the CTA (atomics at the end).
// dL/dY0 is never written and the separate first-layer reverse kernel disappears.
struct FirstLayer {
  const float* x = nullptr;     // [n, in_dim - 1] with t given, else [n, in_dim]
  const float* t = nullptr;     // [n] or null
  const float* W0 = nullptr;    // [width, in_dim]
  const float* b0 = nullptr;    // [width] or null
  float* gW0 = nullptr;         // [width, in_dim] accumulated, or null
  float* gb0 = nullptr;         // [width] accumulated, or null
  int in_dim = 0;
  float vec[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // direction vectors of the (at most two) jet directions
};
__device__ __forceinline__ float load_xt(const float* __restrict__ x, const float* __restrict__ t, int64_t p, int i, int in_dim) {
  if (t == nullptr) return __ldg(x + p * in_dim + i);
  return (i < in_dim - 1) ? __ldg(x + p * (in_dim - 1) + i) : __ldg(t + p);
}

// EPI_ACT only: the network's output layer (nn.Linear(width, 1)) folded into the epilogue of the last hidden layer.
// Every epilogue warp reduces w_out[f] * y[row, f] over its 32 features (butterfly of 16 shuffles for 16 rows) and writes
// the partial output jets to u_part[(4 * n_block + lane quarter) * M + row]; a tiny kernel adds the partials in a fixed
// order (deterministic) and the bias.  The activation output Yact then needs no store at all when nothing else reads it
// (forward-only scoring): 512 B per row less traffic, and no separate pass over Y for the output layer.
struct OutFuse {
  const float* w_out = nullptr;
  float* u_part = nullptr;
};

// (body of linear_rows_ts_kernel; cta_x / ncta_x / cta_y are the block index / grid size the kernel passes in, so that the
// paired reverse kernel can run it on one CTA of each cluster with the pair index; PAIR adds the lock-step throttle to the
// TMA warp)
// KS: K-split of a 256-wide contraction over the two CTAs of a cluster.  KS = 1 (EPI_PARTIAL): this CTA contracts its K half
// and hands the partial product of every tile to its partner through `ring` (kKsRing slots in global memory -- they stay in
// L2 -- guarded by mbarriers in the two CTAs' shared memory); KS = 2 (ACCUM): this CTA contracts the other K half, adds the
// partner's partial product and runs the epilogue.  Both walk the same tiles in the same order.
template <bool TRANS_W, int EPI, int ACT, int K0, int K1, int NLW, int ECOLS, bool ACCUM, int LDYC, bool LOSSF = false, bool PAIR = false,
          int KS = 0>
__device__ __forceinline__ void
linear_rows_ts_body(const float* __restrict__ X, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                    float* __restrict__ Y, int64_t M, int ldy_rt, int jet_cols, const float* __restrict__ Zs,
                    float* __restrict__ Yact, float omega, int ldx, const OutFuse& of, const FirstLayer& fl,
                    const TcLossFuse& lfv, const CUtensorMap* tmxp, const int cta_x, const int ncta_x, const int cta_y,
                    const PairSync ps, float* __restrict__ ring = nullptr) {
  static_assert(!LOSSF || (EPI == EPI_ACT && ACT == 1 && ECOLS == 16 && !ACCUM), "loss fusion: tanh forward epilogue only");
  static_assert((KS == 1) == (EPI == EPI_PARTIAL) && (KS != 2 || ACCUM) && (KS == 0 || (!LOSSF && !PAIR)), "K-split roles");
  // LDYC: compile-time row stride of Y / Zs / Yact (0 = use the runtime value): with it every row address of the
  // epilogue is base + immediate instead of a 64-bit multiply-add per access
  const int ldy = LDYC ? LDYC : ldy_rt;
  constexpr bool IS_BWD = (EPI == EPI_ACTBWD || EPI == EPI_ACTBWD_Y);
  static_assert(EPI != EPI_ACTBWD_Y || ACT == 1, "the output-jet reverse epilogue exists for tanh only");
  // ldx: row stride of X in floats (K = 128 columns of it are contracted: one half of a 256-wide layer);
  // ACCUM: the epilogue adds the partial result already stored in Y (second K half of a 256-wide layer)
  constexpr int K = 128, TN = 64, STAGES = 2, RS = 3, ACC = 2, NEW = 4 * (TN / ECOLS);
  static_assert(ECOLS == 16 || ECOLS == 32, "epilogue warps own 16 or 32 tile rows");
  constexpr int JC = 1 + K0 + K1;                       // jet columns of the fused epilogues
  constexpr int MAXK = (K0 > K1 ? K0 : K1) > 0 ? (K0 > K1 ? K0 : K1) : 1;
  // Rows an epilogue warp really owns (ECE <= ECOLS) and rows per tile (TNE = 4 * ECE <= TN): jet column counts that do
  // not divide 16 (3, 5, 6: Heat / convection, KdV / wave, 1-D Cahn-Hilliard) use 60- or 48-row tiles inside the same
  // 64-row MMA (the spare operand rows are zero), so that every epilogue warp still owns whole points.
  // (EPI_PARTIAL takes the tile geometry of its partner: 32-row warps when that runs the plain epilogue, else the jet layout)
  constexpr int ECE = (EPI == EPI_PLAIN || (EPI == EPI_PARTIAL && ECOLS == 32)) ? ECOLS : jets_rows_per_warp(JC);
  constexpr int TNE = (TN / ECOLS) * ECE;
  static_assert(ECE > 0 && ECE <= ECOLS && (EPI == EPI_PLAIN || EPI == EPI_PARTIAL || (ECE % JC) == 0), "unsupported jet column count for the fused epilogues");
  constexpr int CHUNKS = K / 4;
  constexpr uint32_t X_BYTES = TN * K * 4;        // one of X_hi / X_lo per stage (32 KB)
  constexpr uint32_t RAW_BYTES = TN * K * 4;      // raw fp32 row tile (32 KB)
  constexpr int EPI0 = NLW, MMAW = NLW + NEW, TMAW = NLW + NEW + 1;   // NLW convert warps | NEW epilogue | MMA | TMA
  constexpr uint32_t COL_WHI = 0, COL_WLO = 128, COL_ACC = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_st = smem;                           // [STAGES][hi|lo][X_BYTES]  UMMA operand tiles
  uint8_t* raw_st = x_st + (size_t)STAGES * 2 * X_BYTES;     // [RS][RAW_BYTES]   raw tiles landed by TMA
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_st + (size_t)RS * RAW_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + ACC;
  uint64_t* raw_full = bars + 2 * STAGES + 2 * ACC;
  uint64_t* raw_empty = raw_full + RS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + RS);
  float* u_smem = reinterpret_cast<float*>(tmem_slot + 4);      // [4 row groups][4 lane quarters][16]  (LOSSF)
  uint64_t* pfull = reinterpret_cast<uint64_t*>(u_smem);        // KS (never LOSSF): [kKsRing] partial product landed in slot
  uint64_t* pempty = pfull + kKsRing;                           //                   [kKsRing] slot read by the consumer
  uint32_t* ks_cnt = reinterpret_cast<uint32_t*>(pempty + kKsRing);   //             [kKsRing] producer warps done with the slot

#ifdef PINNK_STAGE_TIMERS
  unsigned long long g_entry; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = cta_y * 128;
  const int64_t ntiles = (M + TNE - 1) / TNE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NLW); mbar_init(&empty[s], 1); }
    for (int b = 0; b < ACC; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], NEW); }
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], NLW); }
    if constexpr (KS != 0) {
      // pfull: ONE arrive per tile (the producer warp that finishes last publishes for all); pempty: one per consumer warp
      for (int s = 0; s < kKsRing; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], NEW); ks_cnt[s] = 0u; }
    }
    fence_mbar_init();
  }
  if (warp == MMAW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // K-split: the partner's barriers (same offsets in its shared memory); both CTAs' barriers exist before either can arrive
  uint32_t peer_pfull = 0, peer_pempty = 0;
  if constexpr (KS != 0) {
    const uint32_t peer = cluster_ctarank() ^ 1u;
    peer_pfull = mapa_u32(smem_u32(pfull), peer);
    peer_pempty = mapa_u32(smem_u32(pempty), peer);
    cluster_sync_all();
  }
  (void)peer_pfull; (void)peer_pempty;

  // resident weights -> TMEM: an epilogue warp (q, h) owns TMEM lanes 32q.. (rows f of the A operand) and stages the
  // 32-column blocks h, h + NEW/4, ...: all epilogue warps load at once, so the prologue is one round of global latency
  if (warp >= EPI0 && warp < MMAW) {
    const int q = (warp - EPI0) & 3, hq = (warp - EPI0) >> 2, f = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 32 * hq; c0 < K; c0 += 32 * (NEW / 4)) {
      uint32_t hi[32], lo[32];
      if constexpr (TRANS_W) {
#pragma unroll
        for (int j = 0; j < 32; ++j) split_bits(W[(int64_t)(c0 + j) * ldw + n0 + f], hi[j], lo[j]);   // lanes contiguous
      } else {
        // a thread's 32 values are one 128-byte line of its weight row: 8 x LDG.128 (ldw % 4 == 0).  kv: contraction columns
        // that exist -- a layer with fewer than 128 inputs (the 64 Fourier features of fourier.py) runs with the missing
        // columns zero here and zero-filled by the tiled copy of its input rows
        const int kv = (ldx < K) ? ldx : K;
        const float4* wr = reinterpret_cast<const float4*>(W + (int64_t)(n0 + f) * ldw + c0);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 v = (c0 + 4 * j4 < kv) ? __ldg(wr + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
          split_bits(v.x, hi[4 * j4], lo[4 * j4]); split_bits(v.y, hi[4 * j4 + 1], lo[4 * j4 + 1]);
          split_bits(v.z, hi[4 * j4 + 2], lo[4 * j4 + 2]); split_bits(v.w, hi[4 * j4 + 3], lo[4 * j4 + 3]);
        }
      }
      tmem_st32(lane_base + COL_WHI + c0, hi);
      tmem_st32(lane_base + COL_WLO + c0, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == TMAW) {
    // ===================== TMA producer: raw 64-row tiles (contiguous 32 KB), one elected thread =====================
    if (elect_one()) {
      const uint32_t rb = smem_u32(raw_st);
      long long t_a = 0; (void)t_a;
      int it = 0;
      for (int64_t tile = cta_x; tile < ntiles; tile += ncta_x, ++it) {
        const int s = it % RS;
        const uint32_t ph = (uint32_t)(it / RS) & 1u;
        { PK_T0(); mbar_wait(&raw_empty[s], ph ^ 1u); PK_TACC(t_a); }
        if constexpr (PAIR) ps.wait_turn((uint32_t)it);
        const int64_t r0 = tile * TNE;
        const uint32_t nrows = (M - r0 >= TNE) ? TNE : (uint32_t)(M - r0);
        const uint32_t dst = rb + (uint32_t)s * RAW_BYTES;
        if (ldx == K) {
          mbar_arrive_expect_tx(&raw_full[s], nrows * (uint32_t)(K * 4));
          tma_bulk_g2s(dst, X + r0 * K, nrows * (uint32_t)(K * 4), &raw_full[s]);
        } else {          // one K half of a wider matrix: tiled copy, box = [TNE rows x 128] (rows past M arrive as zeros)
          mbar_arrive_expect_tx(&raw_full[s], (uint32_t)TNE * (uint32_t)(K * 4));
          tma_tile_2d(dst, tmxp, 0, (int)r0, &raw_full[s]);
        }
        if constexpr (PAIR) ps.publish((uint32_t)it + 1u);
      }
      if constexpr (PAIR) ps.publish(0x7fffffffu);        // done: never hold the partner back
#ifdef PINNK_STAGE_TIMERS
      if (cta_x == 0 && cta_y == 0) atomicAdd(&g_stage_timers[0], (unsigned long long)t_a);
#endif
    }
    __syncwarp();
  } else if (warp < NLW) {
    // ===================== convert warps: raw tile -> hi/lo operand tile (lane = 16-byte chunk of a row) ==============
    constexpr int RPW = TN / NLW;
    const uint32_t x_base = smem_u32(x_st), rb = smem_u32(raw_st);
    long long t_a = 0, t_b = 0, t_c = 0; (void)t_a; (void)t_b; (void)t_c;
    int it = 0;
    for (int64_t tile = cta_x; tile < ntiles; tile += ncta_x, ++it) {
      const int rs = it % RS, s = it % STAGES;
      const uint32_t rph = (uint32_t)(it / RS) & 1u, ph = (uint32_t)(it / STAGES) & 1u;
      const int64_t r0 = tile * TNE;
      const int nrows = (M - r0 >= TNE) ? TNE : (int)(M - r0);
      { PK_T0(); mbar_wait(&raw_full[rs], rph); PK_TACC(t_a); }
      PK_T0();
      const uint32_t raw = rb + (uint32_t)rs * RAW_BYTES;
      float4 v[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = warp + NLW * i;
        v[i] = (r < nrows) ? lds128(raw + r * (K * 4) + lane * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t dep = 0;                                           // one register of every load above
#pragma unroll
      for (int i = 0; i < RPW; ++i) dep |= __float_as_uint(v[i].w);
      __syncwarp();
      if (lane == 0) mbar_arrive_after_loads(&raw_empty[rs], dep, tmem_slot + 1);    // raw data IS in registers: slot may be refilled
      PK_TACC(t_c);
      { PK_T0(); mbar_wait(&empty[s], ph ^ 1u); PK_TACC(t_b); }
      const long long _t1 = clock64(); (void)_t1;
      const uint32_t xh = x_base + (uint32_t)s * 2 * X_BYTES, xl = xh + X_BYTES;
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = warp + NLW * i;
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_bits(v[i].x, h0, l0); split_bits(v[i].y, h1, l1); split_bits(v[i].z, h2, l2); split_bits(v[i].w, h3, l3);
        const uint32_t off = sw128_offset(TN, r, lane);
        sts128(xh + off, h0, h1, h2, h3);
        sts128(xl + off, l0, l1, l2, l3);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
#ifdef PINNK_STAGE_TIMERS
      t_c += clock64() - _t1;
#endif
    }
#ifdef PINNK_STAGE_TIMERS
    if (cta_x == 0 && cta_y == 0 && warp == 0 && lane == 0) {
      atomicAdd(&g_stage_timers[1], (unsigned long long)t_a); atomicAdd(&g_stage_timers[2], (unsigned long long)t_b);
      atomicAdd(&g_stage_timers[3], (unsigned long long)t_c);
    }
#endif
  } else if (warp < MMAW) {
    // ===================== epilogue: warp (q, h) owns lanes 32q.. and tile rows ECOLS*h .. ECOLS*h + ECOLS-1 ==========
    const int e = warp - EPI0;
    const int q = e & 3, h = e >> 2;
    const int f = q * 32 + lane;
    const float bf = (!TRANS_W && bias) ? bias[n0 + f] : 0.f;
    const bool store_z = (EPI != EPI_ACT) || (Y != nullptr);     // forward-only callers (scoring) pass no stash buffer
    const uint32_t lane_acc = tmem_base + ((uint32_t)(q * 32) << 16) + COL_ACC + (uint32_t)(h * ECE);
    // EPI_FIRSTBWD: this thread's row of the input layer and its gradient accumulators
    float w0r[4] = {0.f, 0.f, 0.f, 0.f}, z1d[2] = {0.f, 0.f}, aw0[4] = {0.f, 0.f, 0.f, 0.f}, ab0 = 0.f, b0f = 0.f;
    if constexpr (EPI == EPI_FIRSTBWD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) w0r[i] = (i < fl.in_dim) ? fl.W0[(n0 + f) * fl.in_dim + i] : 0.f;
      b0f = fl.b0 ? fl.b0[n0 + f] : 0.f;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fmaf(w0r[i], fl.vec[d][i], a);
        z1d[d] = (ACT == 2) ? a * omega : a;
      }
    }
    // LOSSF: per-thread accumulators of the output layer's gradient and of the loss
    float gw_acc = 0.f, gb_acc = 0.f;
    double loss_acc = 0.0;
    (void)gw_acc; (void)gb_acc; (void)loss_acc;
    // One tile of this warp.  FULL: all ECOLS rows exist, so no access is predicated and (with LDYC) every row address
    // is the tile base plus an immediate.
    long long t_ea = 0, t_eb = 0; (void)t_ea; (void)t_eb;
    // K-split consumer: the partner's partial product of a tile is fetched one tile AHEAD (while the previous tile's
    // epilogue runs): an epilogue warp walks every tile of the CTA in turn, and an L2 round trip per tile on that serial
    // path (~1 us against 0.8 us of tensor time per tile) halved the throughput.  The MMA warp stays two tiles behind the
    // partner, so the slot is always complete when it is asked for here.  Rows h * ECE .. of ring slot it % kKsRing (every
    // row of a slot is written; L2 only -- the slot is rewritten every kKsRing tiles and L1 is not coherent).
    float part_nx[(KS == 2) ? ECOLS : 1], part_n2[(KS == 2) ? ECOLS : 1];     // tiles it + 1 and it + 2
    auto ks_prefetch = [&](const int it_, float (&dst)[(KS == 2) ? ECOLS : 1]) {
      if constexpr (KS == 2) {
        const int s_ = it_ % kKsRing;
        // (CTA-scope wait: an acquire at cluster scope invalidates the SM's whole L1 -- CCTL.IVALL -- and 17 of those per tile
        // were the longest item of the consumer's tile time.  The phase flips only after the partner's release has made its
        // stores visible in L2, the loads below are issued after the flip is seen and bypass L1 (ld.global.cg).)
        mbar_wait(&pfull[s_], (uint32_t)(it_ / kKsRing) & 1u);
        const float* const rp = ring + (size_t)s_ * (64 * 128) + (size_t)(h * ECE) * 128 + f;
#pragma unroll
        for (int j = 0; j < ECOLS; ++j) dst[j] = (j < ECE) ? __ldcg(rp + j * 128) : 0.f;
      }
    };
    auto run_tile = [&](auto full_tag, const int b, const uint32_t ph, const int64_t r0, const int nrows, const int it_,
                        const bool has_next2) {
      constexpr bool FULL = decltype(full_tag)::value;
      const int ks_s = it_ % kKsRing;                               // K-split: ring slot and barrier phase of this tile
      const uint32_t ks_ph = (uint32_t)(it_ / kKsRing) & 1u;
      (void)ks_s; (void)ks_ph; (void)has_next2;
      uint32_t vmask = 0;
      if constexpr (EPI == EPI_PLAIN) {
        if (!TRANS_W && bias) {
          int cj = (int)((uint32_t)r0 % (uint32_t)jet_cols);
#pragma unroll
          for (int j = 0; j < ECOLS; ++j) { vmask |= (cj == 0 ? 1u : 0u) << j; cj = (cj + 1 == jet_cols) ? 0 : cj + 1; }
        }
      }
      float* const yp = Y + r0 * ldy + n0 + f;
      // the stashed pre-activations do not depend on the MMA: fetch them while the accumulator is still being produced
      float zsr[IS_BWD ? ECOLS : 1];
      if constexpr (IS_BWD) {
        const float* const zs0 = Zs + r0 * ldy + n0 + f;
#pragma unroll
        for (int j = 0; j < ECOLS; ++j) zsr[j] = (j < ECE && (FULL || j < nrows)) ? __ldg(zs0 + j * ldy) : 0.f;
      }
      float xin[(EPI == EPI_FIRSTBWD) ? ECE / JC : 1][4];
      if constexpr (EPI == EPI_FIRSTBWD) {
        const int64_t p0 = r0 / JC;                    // r0 is a multiple of JC (whole points per warp)
#pragma unroll
        for (int pp = 0; pp < ECE / JC; ++pp)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            xin[pp][i] = (i < fl.in_dim && (FULL || pp * JC < nrows)) ? load_xt(fl.x, fl.t, p0 + pp, i, fl.in_dim) : 0.f;
      }
      // second K half of a 256-wide layer: the first half's partial result is in Y
      float part[ACCUM ? ECOLS : 1];
      uint32_t pdep = 0;
      (void)pdep;
      if constexpr (ACCUM && KS == 2) {
        // (fetched two tiles ago -- ks_prefetch; read below, after the accumulator has arrived)
      } else if constexpr (ACCUM) {
#pragma unroll
        for (int j = 0; j < ECOLS; ++j) part[j] = (j < ECE && (FULL || j < nrows)) ? yp[j * ldy] : 0.f;
      }
      { PK_T0(); mbar_wait_relaxed(&tfull[b], ph); PK_TACC(t_ea); }
      PK_T0();
      tc_fence_after();
      const uint32_t tb = lane_acc + (uint32_t)(b * 2 * TN);
      uint32_t pm[ECOLS], pc[ECOLS];
      if constexpr (ECOLS == 32) { tmem_ld32_nowait(tb, pm); tmem_ld32_nowait(tb + TN, pc); }
      else { tmem_ld16_nowait(tb, pm); tmem_ld16_nowait(tb + TN, pc); }
      tmem_wait_ld(pm, pc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
      float acc[ECOLS];                                            // main + correction (+ first K half)
#pragma unroll
      for (int j = 0; j < ECOLS; ++j) {
        acc[j] = __uint_as_float(pc[j]) + __uint_as_float(pm[j]);
        if constexpr (ACCUM && KS == 2) { acc[j] += part_nx[j]; pdep |= __float_as_uint(part_nx[j]); }
        else if constexpr (ACCUM) acc[j] += part[j];
      }
      if constexpr (KS == 2) {
        // this tile's slot has been read (its values are in acc): the partner may refill it; the next tile's partial product
        // is already on its way (or here), the one after that is requested now
        __syncwarp();
        if (lane == 0) {
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(tmem_slot + 2)), "r"(pdep) : "memory");
          mbar_arrive_remote_relaxed(peer_pempty + (uint32_t)ks_s * 8u);
        }
#pragma unroll
        for (int j = 0; j < ECOLS; ++j) part_nx[j] = part_n2[j];
        if (has_next2) ks_prefetch(it_ + 2, part_n2);
      }
      if constexpr (EPI == EPI_PLAIN) {
#pragma unroll
        for (int j = 0; j < ECOLS; ++j) {
          float val = acc[j];
          if ((vmask >> j) & 1u) val += bf;
          if (FULL || j < nrows) yp[j * ldy] = val;
        }
      } else if constexpr (EPI == EPI_PARTIAL) {
        // hand the partial product to the partner: wait until it has read this slot's previous tile, store, publish
        mbar_wait(&pempty[ks_s], ks_ph ^ 1u);          // (write-after-read only: nothing of the partner's is read here)
        float* const rp = ring + (size_t)ks_s * (64 * 128) + (size_t)(h * ECE) * 128 + f;
#pragma unroll
        for (int j = 0; j < ECOLS; ++j)
          if (j < ECE) __stcg(rp + j * 128, acc[j]);
        // Publish once per tile: every warp counts itself done (acq_rel at CTA scope, after a warp barrier that orders all its
        // lanes' stores before the count), and the warp that arrives last releases the slot to the partner at cluster scope.
        // A release is cumulative over what happens-before it, so one memory barrier per tile covers all 16 warps' stores
        // (a barrier per warp and tile -- let alone per lane -- was the longest item on these warps' serial path).
        __syncwarp();
        if (lane == 0) {
          uint32_t old;
          asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&ks_cnt[ks_s])) : "memory");
          if (old == (uint32_t)(NEW - 1)) {
            asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&ks_cnt[ks_s])), "r"(0u) : "memory");
            mbar_arrive_remote(peer_pfull + (uint32_t)ks_s * 8u);
          }
        }
      } else {
        // whole points: columns [pp*JC, pp*JC + JC) of this warp's ECOLS are the jet of point pp at feature f
        float* const ya = (EPI == EPI_ACT && Yact != nullptr) ? Yact + r0 * ldy + n0 + f : nullptr;
        float yk[LOSSF ? ECOLS : 1];                              // output jets kept for the in-kernel adjoint
        (void)yk;
        const float wo = (EPI == EPI_ACT && of.w_out != nullptr) ? of.w_out[n0 + f] : 0.f;
#pragma unroll
        for (int pp = 0; pp < ECE / JC; ++pp) {
          const int jb = pp * JC;
          if (FULL || jb < nrows) {
            float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1];
            if constexpr (EPI == EPI_ACT) {
              z[0] = acc[jb] + bf;
              if (store_z) yp[jb * ldy] = z[0];
              if (ACT == 1) { y[0] = tanhf(z[0]); w[0] = 1.f - y[0] * y[0]; }
              else { z[0] *= omega; sincosf(z[0], &y[0], &w[0]); }
              if (ya) ya[jb * ldy] = y[0];
              acc[jb] = y[0] * wo;                                   // (the accumulator value is consumed: reuse the slot)
              if constexpr (LOSSF) yk[jb] = y[0];
#pragma unroll
              for (int d = 0; d < 2; ++d) {
                const int KD = d ? K1 : K0, cb = jb + (d ? K0 : 0);
                if (KD > 0) {
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k)
                    if (k <= KD) {
                      z[k] = acc[cb + k];
                      if (store_z) yp[(cb + k) * ldy] = z[k];
                      if (ACT == 2) z[k] *= omega;
                    }
                  if (ACT == 1) tanh_dir_fwd<MAXK, float>(KD, z, y, w);
                  else sincos_dir_fwd<MAXK, float>(KD, z, y, w);
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k)
                    if (k <= KD) {
                      if (ya) ya[(cb + k) * ldy] = y[k];
                      acc[cb + k] = y[k] * wo;
                      if constexpr (LOSSF) yk[cb + k] = y[k];
                    }
                }
              }
            } else if constexpr (EPI == EPI_FIRSTBWD) {   // reverse of the input layer: nothing is written per row
              float yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
              z[0] = b0f;
#pragma unroll
              for (int i = 0; i < 4; ++i) z[0] = fmaf(w0r[i], xin[pp][i], z[0]);
              if (ACT == 1) { y[0] = tanhf(z[0]); w[0] = 1.f - y[0] * y[0]; }
              else { z[0] *= omega; sincosf(z[0], &y[0], &w[0]); }
              yb[0] = acc[jb];
              float wb0 = 0.f;
#pragma unroll
              for (int d = 0; d < 2; ++d) {
                const int KD = d ? K1 : K0, cb = jb + (d ? K0 : 0);
                if (KD > 0) {
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k) {
                    z[k] = (k == 1) ? z1d[d] : 0.f;
                    yb[k] = (k <= KD) ? acc[cb + k] : 0.f;
                  }
                  if (ACT == 1) {
                    tanh_dir_fwd<MAXK, float>(KD, z, y, w);
                    tanh_dir_bwd<MAXK, float>(KD, z, y, w, yb, zb, wb0);
                  } else {
#pragma unroll
                    for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
                    wb[0] = wb0;
                    sincos_dir_fwd<MAXK, float>(KD, z, y, w);
                    sincos_dir_bwd<MAXK, float>(KD, z, y, w, yb, wb, zb);
                    wb0 = wb[0];
                  }
                  const float g1 = (ACT == 2) ? zb[1] * omega : zb[1];   // only the first-order pre-activation depends on W0
#pragma unroll
                  for (int i = 0; i < 4; ++i) aw0[i] = fmaf(g1, fl.vec[d][i], aw0[i]);
                }
              }
              const float g0 = (ACT == 1) ? tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0) : (yb[0] * w[0] - wb0 * y[0]) * omega;
              ab0 += g0;
#pragma unroll
              for (int i = 0; i < 4; ++i) aw0[i] = fmaf(g0, xin[pp][i], aw0[i]);
            } else if constexpr (EPI == EPI_ACTBWD_Y) {   // tanh adjoint from the stashed OUTPUT jets
              float yb[MAXK + 1], zb[MAXK + 1];
              y[0] = zsr[jb];
              w[0] = 1.f - y[0] * y[0];
              const float inv_w0 = (w[0] > 1e-30f) ? __fdividef(1.f, w[0]) : 0.f;
              yb[0] = acc[jb];
              float wb0 = 0.f;
#pragma unroll
              for (int d = 0; d < 2; ++d) {
                const int KD = d ? K1 : K0, cb = jb + (d ? K0 : 0);
                if (KD > 0) {
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k) {
                    y[k] = (k <= KD) ? zsr[cb + k] : 0.f;
                    yb[k] = (k <= KD) ? acc[cb + k] : 0.f;
                  }
                  tanh_dir_recover<MAXK, float>(KD, y, w, z, inv_w0);
                  tanh_dir_bwd<MAXK, float>(KD, z, y, w, yb, zb, wb0);
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k)
                    if (k <= KD) yp[(cb + k) * ldy] = zb[k];
                }
              }
              yp[jb * ldy] = tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0);
            } else {   // EPI_ACTBWD
              float yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
              z[0] = zsr[jb];
              if (ACT == 1) { y[0] = tanhf(z[0]); w[0] = 1.f - y[0] * y[0]; }
              else { z[0] *= omega; sincosf(z[0], &y[0], &w[0]); }
              yb[0] = acc[jb];
              float wb0 = 0.f;
#pragma unroll
              for (int d = 0; d < 2; ++d) {
                const int KD = d ? K1 : K0, cb = jb + (d ? K0 : 0);
                if (KD > 0) {
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k) {
                    if (k <= KD) {
                      z[k] = zsr[cb + k];
                      if (ACT == 2) z[k] *= omega;
                      yb[k] = acc[cb + k];
                    } else {
                      yb[k] = 0.f;
                    }
                  }
                  if (ACT == 1) {
                    tanh_dir_fwd<MAXK, float>(KD, z, y, w);
                    tanh_dir_bwd<MAXK, float>(KD, z, y, w, yb, zb, wb0);
                  } else {
#pragma unroll
                    for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
                    wb[0] = wb0;
                    sincos_dir_fwd<MAXK, float>(KD, z, y, w);
                    sincos_dir_bwd<MAXK, float>(KD, z, y, w, yb, wb, zb);
                    wb0 = wb[0];
                  }
#pragma unroll
                  for (int k = 1; k <= MAXK; ++k)
                    if (k <= KD) yp[(cb + k) * ldy] = (ACT == 2) ? zb[k] * omega : zb[k];
                }
              }
              yp[jb * ldy] = (ACT == 1) ? tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0)
                                        : (yb[0] * w[0] - wb0 * y[0]) * omega;
            }
          }
        }
        if constexpr (EPI == EPI_ACT) {
          if (of.w_out != nullptr) {
            // acc[j] = w_out[f] * y[row j, f]: sum over the 32 features of this warp.  Level with stride S halves the
            // number of rows a lane carries; after the last level lanes 2r, 2r+1 both hold the sum of row r.
            if (!FULL) {
#pragma unroll
              for (int j = 0; j < ECOLS; ++j) if (j >= nrows) acc[j] = 0.f;
            }
#pragma unroll
            for (int j = ECE; j < ECOLS; ++j) acc[j] = 0.f;          // columns of the next warp's rows
            constexpr int NLVL = (ECOLS == 16) ? 4 : 5;
#pragma unroll
            for (int lvl = 0; lvl < NLVL; ++lvl) {
              const int S = 16 >> lvl, n = ECOLS >> lvl;
              const bool upper = (lane & S) != 0;
#pragma unroll
              for (int i = 0; i < ECOLS / 2; ++i) {
                if (i < n / 2) {
                  const float a = acc[i], bq = acc[i + n / 2];
                  const float send = upper ? a : bq, keep = upper ? bq : a;
                  acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
                }
              }
            }
#pragma unroll
            for (int S = 32 / ECOLS / 2; S >= 1; S >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], S);
            const int rj = lane / (32 / ECOLS);
            if constexpr (!LOSSF) {
              if ((lane % (32 / ECOLS)) == 0 && rj < ECE && (FULL || rj < nrows))
                of.u_part[(int64_t)(cta_y * 4 + q) * M + r0 + rj] = acc[0];
            } else {
              // ---- output jets of the tile through shared memory (fixed summation order over the four lane quarters)
              const TcLossFuse& lf = lfv;
              float* const ug = u_smem + h * 64;
              if ((lane & 1) == 0 && rj < ECE) ug[q * 16 + rj] = (FULL || rj < nrows) ? acc[0] : 0.f;
              named_bar_sync(1 + h, 128);
              float Uv = 0.f;
              if (lane < ECE) {
                Uv = ((ug[lane] + ug[16 + lane]) + ug[32 + lane]) + ug[48 + lane];
                if (lf.b_out != nullptr && (lane % JC) == 0) Uv += __ldg(lf.b_out);
              }
              named_bar_sync(1 + h, 128);                          // the group's slots may be rewritten (next tile)
              // ---- residual, loss and seeds: lane l < ECE works on point l / JC (the JC lanes of a point redundantly),
              // so ONE residual evaluation per tile and warp covers all its points; lane l then owns dL/dU of row l
              float sd = 0.f;
              {
                const int pl_ = (lane < ECE) ? lane / JC : 0, cl_ = (lane < ECE) ? lane - pl_ * JC : 0;
                float u[kMaxCols], du[kMaxCols];
#pragma unroll
                for (int c = 0; c < JC; ++c) u[c] = __shfl_sync(0xffffffffu, Uv, pl_ * JC + c);
                // column layout of this instantiation, [u | K0 columns of direction 0 | K1 columns of direction 1], as compile-time
                // constants (the launcher checks lf.js against it): with the run-time JetSpec the residual's U[js.col0[..]] /
                // dU[..] accesses were dynamically indexed, which put u[] and du[] in LOCAL memory -- 15 local loads / stores per
                // warp and tile, 2.7 GB of L2 traffic per 4 Mi-row launch next to the kernel's 4.3 GB (ncu, r02_full_lossf_raw.csv)
                JetSpec jsl;
                jsl.ncols = JC;
                jsl.ndirs = (K1 > 0) ? 2 : 1;
                jsl.col0[0] = 1;
                jsl.col0[1] = 1 + K0;
#pragma unroll
                for (int d = 2; d < kMaxDirs; ++d) jsl.col0[d] = 1;     // (five-direction operators never come here: the launcher declines)
                const float e = pde_residual<float>(lf.pde, jsl, u, du);
                float drho;
                const float rho = loss_rho<float>(lf.loss_kind, lf.huber_delta, e, &drho);
                const bool live = lane < ECE && (FULL || lane < nrows);
                float dsel = 0.f;
#pragma unroll
                for (int c = 0; c < JC; ++c) dsel = (c == cl_) ? du[c] : dsel;
                sd = live ? drho * lf.grad_weight * dsel : 0.f;
                if (q == 0 && live && cl_ == 0) { loss_acc += (double)rho * (double)lf.weight; gb_acc += sd; }
              }
              // ---- adjoint of output layer + tanh, per point (every lane: its own feature f):
              // dL/dy[c] = dL/dU[c] * w_out[f];  dL/dw_out[f] += dL/dU[c] * y[c]
#pragma unroll
              for (int pp = 0; pp < ECE / JC; ++pp) {
                const int jb = pp * JC;
                float s_[JC];
#pragma unroll
                for (int c = 0; c < JC; ++c) s_[c] = __shfl_sync(0xffffffffu, sd, jb + c);
                if (FULL || jb < nrows) {
                  float y[MAXK + 1], w[MAXK + 1], z[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1];
                  y[0] = yk[jb];
                  w[0] = 1.f - y[0] * y[0];
                  const float inv_w0 = (w[0] > 1e-30f) ? __fdividef(1.f, w[0]) : 0.f;
                  yb[0] = s_[0] * wo;
                  gw_acc = fmaf(s_[0], y[0], gw_acc);
                  float wb0 = 0.f;
                  float* const gp = lf.dz_out + (r0 + jb) * ldy + n0 + f;
#pragma unroll
                  for (int d = 0; d < 2; ++d) {
                    const int KD = d ? K1 : K0, cb = d ? K0 : 0;
                    if (KD > 0) {
#pragma unroll
                      for (int k = 1; k <= MAXK; ++k) {
                        y[k] = (k <= KD) ? yk[jb + cb + k] : 0.f;
                        yb[k] = (k <= KD) ? s_[(k <= KD) ? cb + k : 0] * wo : 0.f;
                        if (k <= KD) gw_acc = fmaf(s_[cb + k], y[k], gw_acc);
                      }
                      tanh_dir_recover<MAXK, float>(KD, y, w, z, inv_w0);
                      tanh_dir_bwd<MAXK, float>(KD, z, y, w, yb, zb, wb0);
#pragma unroll
                      for (int k = 1; k <= MAXK; ++k)
                        if (k <= KD) gp[(cb + k) * ldy] = zb[k];
                    }
                  }
                  gp[0] = tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0);
                }
              }
            }
          }
        }
      }
      PK_TACC(t_eb);
    };
    int it = 0;
    if constexpr (KS == 2) {
      if ((int64_t)cta_x < ntiles) ks_prefetch(0, part_nx);
      if ((int64_t)cta_x + ncta_x < ntiles) ks_prefetch(1, part_n2);
    }
    for (int64_t tile = cta_x; tile < ntiles; tile += ncta_x, ++it) {
      const int b = it % ACC;
      const uint32_t ph = (uint32_t)(it / ACC) & 1u;
      const int64_t r0 = tile * TNE + h * ECE;
      const bool has_next2 = tile + 2 * (int64_t)ncta_x < ntiles;
      if (M - r0 >= ECE) run_tile(std::true_type(), b, ph, r0, ECE, it, has_next2);
      else run_tile(std::false_type(), b, ph, r0, (int)(M - r0 > 0 ? M - r0 : 0), it, has_next2);
    }
    if constexpr (LOSSF) {
      if (lfv.gw_out) atomicAdd(lfv.gw_out + n0 + f, gw_acc);
      if (q == 0) {                                    // the value-column lanes of this warp hold the partial sums
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gb_acc += __shfl_xor_sync(0xffffffffu, gb_acc, o);
          loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        }
        if (lane == 0) {
          if (lfv.gb_out) atomicAdd(lfv.gb_out, gb_acc);
          if (lfv.loss_slot) atomicAdd(lfv.loss_slot, loss_acc);
        }
      }
    }
    if constexpr (EPI == EPI_FIRSTBWD) {
      if (fl.gW0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < fl.in_dim) atomicAdd(fl.gW0 + (int64_t)(n0 + f) * fl.in_dim + i, aw0[i]);
      }
      if (fl.gb0) atomicAdd(fl.gb0 + n0 + f, ab0);
    }
#ifdef PINNK_STAGE_TIMERS
    if (cta_x == 0 && cta_y == 0 && e == 0 && lane == 0) {
      atomicAdd(&g_stage_timers[7], (unsigned long long)t_ea); atomicAdd(&g_stage_timers[8], (unsigned long long)t_eb);
    }
#endif
  } else if (warp == MMAW) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(128, TN);
      const uint64_t dconst = make_desc_k_sw128(0);
      const uint32_t x_base = smem_u32(x_st) >> 4;
      long long t_a = 0, t_b = 0, t_c = 0, n_t = 0; const long long t_start = clock64();
      (void)t_a; (void)t_b; (void)t_c; (void)n_t; (void)t_start;
#ifdef PINNK_STAGE_TIMERS
      unsigned long long g_start; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
#endif
      int it = 0;
      for (int64_t tile = cta_x; tile < ntiles; tile += ncta_x, ++it) {
        const int s = it % STAGES, b = it % ACC;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u, bph = (uint32_t)(it / ACC) & 1u;
        { PK_T0(); mbar_wait(&tempty[b], bph ^ 1u); PK_TACC(t_a); }
        { PK_T0(); mbar_wait(&full[s], ph); PK_TACC(t_b); }
        if constexpr (KS == 2) {
          // stay kKsLag tiles behind the partner: when this tile's accumulator completes, its partial product has been in
          // the ring long enough for the epilogue warps' loads (issued before they wait for the accumulator) to have landed;
          // in lock step every tile would expose an L2 round trip in the epilogue
          constexpr int kKsLag = 3;
          if (tile + (int64_t)kKsLag * ncta_x < ntiles) {
            const int it2 = it + kKsLag;
            mbar_wait(&pfull[it2 % kKsRing], (uint32_t)(it2 / kKsRing) & 1u);
          }
        }
        PK_T0();
        tc_fence_after();
        const uint32_t xh = x_base + (uint32_t)s * (2 * X_BYTES >> 4), xl = xh + (X_BYTES >> 4);
        const uint32_t d_main = tmem_base + COL_ACC + (uint32_t)(b * 2 * TN), d_corr = d_main + TN;
#pragma unroll
        for (int k8 = 0; k8 < K / 8; ++k8) {
          const uint32_t bo = (uint32_t)(((k8 >> 2) * TN * 128 + (k8 & 3) * 32) >> 4);
          const uint64_t b_hi = dconst | (uint64_t)(xh + bo), b_lo = dconst | (uint64_t)(xl + bo);
          const uint32_t a_hi = tmem_base + COL_WHI + (uint32_t)(k8 * 8), a_lo = tmem_base + COL_WLO + (uint32_t)(k8 * 8);
          umma_tf32_ts(d_corr, a_lo, b_hi, idesc, k8 ? 1u : 0u);
          umma_tf32_ts(d_corr, a_hi, b_lo, idesc, 1u);
          umma_tf32_ts(d_main, a_hi, b_hi, idesc, k8 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        umma_commit(&tfull[b]);
        PK_TACC(t_c);
        ++n_t;
      }
#ifdef PINNK_STAGE_TIMERS
      {   // slowest and fastest CTA of the launch (loop time): static tile assignment vs uneven memory paths
        unsigned long long g_e2; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_e2));
        atomicMax(&g_stage_timers[15], g_e2 - g_start);
      }
      if (cta_x == 0 && cta_y == 0) {
        atomicAdd(&g_stage_timers[4], (unsigned long long)t_a); atomicAdd(&g_stage_timers[5], (unsigned long long)t_b);
        atomicAdd(&g_stage_timers[6], (unsigned long long)t_c); atomicAdd(&g_stage_timers[9], (unsigned long long)n_t);
        atomicAdd(&g_stage_timers[10], (unsigned long long)(clock64() - t_start));
        unsigned long long g_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
        atomicAdd(&g_stage_timers[11], g_end - g_start);
        atomicAdd(&g_stage_timers[12], g_start - g_entry);     // kernel entry -> first MMA wait (prologue)
      }
#endif
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  // K-split: no CTA may exit while its partner can still arrive on its barriers
  if constexpr (KS != 0) cluster_sync_all();
  if (warp == MMAW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
#ifdef PINNK_STAGE_TIMERS
    if (cta_x == 0 && cta_y == 0 && lane == 0) {
      unsigned long long g_exit; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_exit));
      atomicAdd(&g_stage_timers[13], g_exit - g_entry);         // whole kernel, block 0
      atomicAdd(&g_stage_timers[14], 1ull);
    }
#endif
  }
}

template <bool TRANS_W, int EPI, int ACT, int K0, int K1, int NLW, int ECOLS, bool ACCUM, int LDYC, bool LOSSF = false>
__global__ void __launch_bounds__((NLW + 4 * (64 / ECOLS) + 2) * 32, 1)
linear_rows_ts_kernel(const float* __restrict__ X, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                      float* __restrict__ Y, int64_t M, int ldy_rt, int jet_cols, const float* __restrict__ Zs,
                      float* __restrict__ Yact, float omega, int ldx, OutFuse of, FirstLayer fl,
                      TcLossFuse lfv, const __grid_constant__ CUtensorMap tmx) {
  linear_rows_ts_body<TRANS_W, EPI, ACT, K0, K1, NLW, ECOLS, ACCUM, LDYC, LOSSF, false>(
      X, W, ldw, bias, Y, M, ldy_rt, jet_cols, Zs, Yact, omega, ldx, of, fl, lfv, &tmx, (int)blockIdx.x, (int)gridDim.x,
      (int)blockIdx.y, PairSync{});
}

template <bool TRANS_W, int EPI, int ACT, int K0, int K1, bool ACCUM, int LDYC, bool LOSSF = false>
static int launch_linear_rows_ts_inst(const float* X, const float* W, int ldw, const float* bias, float* Y, int64_t M, int n_cols,
                                      int jet_cols, const float* Zs, float* Yact, float omega, int sm_count, cudaStream_t st,
                                      int ldx, OutFuse of = OutFuse{}, FirstLayer fl = FirstLayer{}, const TcLossFuse* loss = nullptr) {
  constexpr size_t smem = 1024 + (size_t)2 * 2 * 64 * 128 * 4 + (size_t)3 * 64 * 128 * 4 + (2 * 2 + 2 * 2 + 2 * 3) * 8 + 16 + 1024;
  static_assert(smem <= 232448, "shared memory budget (227 KB per CTA)");
  constexpr int NLW = 8, ECOLS = (EPI == EPI_PLAIN) ? 32 : 16;
  constexpr int TNE = (EPI == EPI_PLAIN) ? 64 : 4 * jets_rows_per_warp(1 + K0 + K1);      // rows per tile
  const int64_t ntiles = (M + TNE - 1) / TNE;
  const int per_y = n_cols / 128;
  int gx = sm_count / per_y;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > ntiles) gx = (int)ntiles;
  dim3 grid((unsigned)gx, (unsigned)per_y, 1);
  constexpr int threads = (NLW + 4 * (64 / ECOLS) + 2) * 32;
  auto kern = linear_rows_ts_kernel<TRANS_W, EPI, ACT, K0, K1, NLW, ECOLS, ACCUM, LDYC, LOSSF>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  TcLossFuse lf;
  memset(&lf, 0, sizeof(lf));
  if (loss) {
    lf = *loss;
    // the kernel's residual uses the compile-time column layout [u | K0 | K1]
    const bool two = K1 > 0;
    if (lf.js.ncols != 1 + K0 + K1 || lf.js.ndirs != (two ? 2 : 1) || lf.js.col0[0] != 1 || (two && lf.js.col0[1] != 1 + K0))
      return TC_UNSUPPORTED;
  }
  alignas(64) CUtensorMap tmx;
  memset(&tmx, 0, sizeof(tmx));
  if (ldx != 128 && !make_tmap_rows(&tmx, X, M, ldx < 128 ? ldx : 128, ldx, TNE)) return -1;
  kern<<<grid, threads, smem, st>>>(X, W, ldw, bias, Y, M, n_cols, jet_cols, Zs, Yact, omega, ldx, of, fl, lf, tmx);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// 128-wide outputs (every BASELINE feed-forward / Fourier layer) get the compile-time row stride; 256-wide layers run as
// two K halves (ACCUM) with the runtime stride
template <bool TRANS_W, int EPI, int ACT, int K0, int K1>
static int launch_linear_rows_ts(const float* X, const float* W, int ldw, const float* bias, float* Y, int64_t M, int n_cols,
                                 int jet_cols, const float* Zs, float* Yact, float omega, int sm_count, cudaStream_t st,
                                 int ldx = 128, int accum = 0, OutFuse of = OutFuse{}, FirstLayer fl = FirstLayer{},
                                 const TcLossFuse* loss = nullptr) {
  if constexpr (EPI == EPI_ACT && ACT == 1) {
    if (loss != nullptr) {         // loss-fused last hidden layer (tanh, one K pass, one 128-wide output block)
      if (accum || n_cols != 128 || of.w_out == nullptr) return TC_UNSUPPORTED;
      return launch_linear_rows_ts_inst<TRANS_W, EPI, ACT, K0, K1, false, 128, true>(X, W, ldw, bias, Y, M, n_cols, jet_cols, Zs, Yact, omega, sm_count, st, ldx, of, fl, loss);
    }
  } else {
    if (loss != nullptr) return TC_UNSUPPORTED;
  }
  if (accum)
    return launch_linear_rows_ts_inst<TRANS_W, EPI, ACT, K0, K1, true, 0>(X, W, ldw, bias, Y, M, n_cols, jet_cols, Zs, Yact, omega, sm_count, st, ldx, of, fl);
  if (n_cols == 128)
    return launch_linear_rows_ts_inst<TRANS_W, EPI, ACT, K0, K1, false, 128>(X, W, ldw, bias, Y, M, n_cols, jet_cols, Zs, Yact, omega, sm_count, st, ldx, of, fl);
  return launch_linear_rows_ts_inst<TRANS_W, EPI, ACT, K0, K1, false, 0>(X, W, ldw, bias, Y, M, n_cols, jet_cols, Zs, Yact, omega, sm_count, st, ldx, of, fl);
}

// ------------------------------------------------------------------------------------------------
// K-split rows kernel for 256-wide contractions (ResNet 6x256, SIREN 5x256): ONE launch of CTA pairs (clusters of two along
// grid z) instead of two K-half launches with the partial product written to and re-read from HBM.  A 128 x 256 weight block
// in hi/lo form is 256 KB -- all of one SM's tensor memory -- so the contraction has to be split over two SMs either way;
// here both halves run at the same time on the same tiles: rank 0 (EPI_PARTIAL) contracts K half 0 and hands each 64 x 128
// partial product to rank 1 through a 4-slot ring in global memory (128 KB per pair, 9.5 MB in all: it lives in L2), rank 1
// contracts K half 1, adds the partial product and runs the epilogue (bias / activation jets / activation adjoint).  HBM
// traffic per row drops from 4.5 - 5.5 KB (two passes) to what the layer needs (read X, (stash,) write Y), and both halves are
// tensor-bound like the first pass alone was (82 % tensor pipe active, profiles/r02s_c3_rows.md).
template <bool TRANS_W, int EPI, int ACT, int K0, int K1, int ECOLS, int LDYC>
__global__ void __cluster_dims__(1, 1, 2) __launch_bounds__((8 + 4 * (64 / ECOLS) + 2) * 32, 1)
linear_rows_ts_ksplit_kernel(const float* __restrict__ X, const float* __restrict__ W, int ldw, int w_half_stride,
                             const float* __restrict__ bias, float* __restrict__ Y, int64_t M, int ldy_rt, int jet_cols,
                             const float* __restrict__ Zs, float* __restrict__ Yact, float omega, int ldx, OutFuse of,
                             float* __restrict__ ring_base, const __grid_constant__ CUtensorMap tm0,
                             const __grid_constant__ CUtensorMap tm1) {
  float* const ring = ring_base + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * (kKsRing * 64 * 128);
  if (cluster_ctarank() == 0) {
    linear_rows_ts_body<TRANS_W, EPI_PARTIAL, 1, K0, K1, 8, ECOLS, false, 128, false, false, 1>(
        X, W, ldw, nullptr, nullptr, M, ldy_rt, jet_cols, nullptr, nullptr, 1.f, ldx, OutFuse{}, FirstLayer{}, TcLossFuse{}, &tm0,
        (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, PairSync{}, ring);
  } else {
    linear_rows_ts_body<TRANS_W, EPI, ACT, K0, K1, 8, ECOLS, true, LDYC, false, false, 2>(
        X + 128, W + w_half_stride, ldw, bias, Y, M, ldy_rt, jet_cols, Zs, Yact, omega, ldx, of, FirstLayer{}, TcLossFuse{}, &tm1,
        (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, PairSync{}, ring);
  }
}

// rows below which the two-pass route is used (a K-split launch keeps every SM busy only with >= 2 tiles per CTA pair)
constexpr int64_t kKsMinRows = 16384;

template <bool TRANS_W, int EPI, int ACT, int K0, int K1>
static int launch_linear_rows_ts_ksplit(const float* X, const float* W, int ldw, const float* bias, float* Y, int64_t M, int n_cols,
                                        int jet_cols, const float* Zs, float* Yact, float omega, int sm_count, cudaStream_t st,
                                        OutFuse of, float* ring, int64_t ring_floats) {
  constexpr size_t smem = 1024 + (size_t)2 * 2 * 64 * 128 * 4 + (size_t)3 * 64 * 128 * 4 + (2 * 2 + 2 * 2 + 2 * 3) * 8 + 16 + 1024;
  static_assert(smem <= 232448, "shared memory budget (227 KB per CTA)");
  constexpr int ECOLS = 16;        // 16 epilogue warps in every variant (the plain epilogue too: twice the loads in flight)
  constexpr int TNE = (EPI == EPI_PLAIN) ? 64 : 4 * jets_rows_per_warp(1 + K0 + K1);
  constexpr int ldx = 256;
  const int64_t ntiles = (M + TNE - 1) / TNE;
  const int per_y = n_cols / 128;
  int gx = sm_count / (2 * per_y);
  if (gx < 1 || ring == nullptr || M < kKsMinRows) return TC_UNSUPPORTED;
  if ((int64_t)gx > ntiles) gx = (int)ntiles;
  if ((int64_t)gx * per_y * kKsRing * 64 * 128 > ring_floats) return TC_UNSUPPORTED;
  constexpr int threads = (8 + 4 * (64 / ECOLS) + 2) * 32;
  if (n_cols != 256) return TC_UNSUPPORTED;                    // (compile-time output stride: every 256-wide net of the path)
  auto kern = linear_rows_ts_ksplit_kernel<TRANS_W, EPI, ACT, K0, K1, ECOLS, 256>;
  static bool configured = false;
  static int max_clusters = 0;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    // How many CTA pairs can be resident at once?  Not necessarily sm_count / 2: a pair needs both SMs of one TPC, and parts
    // ship with TPCs that have a single SM enabled.  A persistent grid with even one pair too many runs that pair as a
    // second wave -- twice the kernel time (measured: 74 pairs on a 148-SM B200 took 2x the time of the pairs that fit).
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (unsigned)(sm_count / 2), 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = sm_count / 2; }
    max_clusters = n;
    configured = true;
  }
  if (const char* e = getenv("PINNK_KS_PAIRS")) { const int v = atoi(e); if (v > 0) max_clusters = v; }     // (experiments)
  if (gx * per_y > max_clusters) gx = max_clusters / per_y;
  if (gx < 1) return TC_UNSUPPORTED;
  dim3 grid((unsigned)gx, (unsigned)per_y, 2);
  alignas(64) CUtensorMap tm0, tm1;
  if (!make_tmap_rows(&tm0, X, M, 128, ldx, TNE) || !make_tmap_rows(&tm1, X + 128, M, 128, ldx, TNE)) return -1;
  // K half 1 of the weights: columns 128.. of W[n_out, 256] (forward) or rows 128.. of W[256, n_in] (dgrad)
  const int w_half_stride = TRANS_W ? 128 * ldw : 128;
  kern<<<grid, threads, smem, st>>>(X, W, ldw, w_half_stride, bias, Y, M, n_cols, jet_cols, Zs, Yact, omega, ldx, of, ring, tm0, tm1);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: dW[o0.., i0..] (128 x 128 block) += sum_rows G[row, o0..]^T X[row, i0..]   (+ db[o] += G rows with
// row % jet_cols == 0).  The contraction runs over the rows, so both operands are MN-major (see make_desc_mn_sw128).
//
// Data path: one TMA thread bulk-copies raw fp32 row tiles (global -> smem ring, mbarrier transaction bytes), NCW convert
// warps turn a raw tile into the hi/lo UMMA operand tile (smem -> registers -> swizzled smem), one thread issues the
// MMAs.  No warp that executes the generic->async proxy fence has global loads in flight (the fence is a full memory
// barrier for the executing thread and would otherwise expose the HBM latency every tile).
//
// Accuracy: the tensor core rounds every accumulate step toward zero, which over the thousands of K-steps of a
// 1M-point batch would shrink the gradient by ~1e-4.  So each CTA accumulates SEG row tiles at a time in one of two
// "main" TMEM accumulators (hi*hi products), while four flush warps fold the finished segment into a running fp32 sum
// (kept in TMEM, added in registers with round-to-nearest).  The tiny lo*hi + hi*lo corrections accumulate in their own
// TMEM region for the whole kernel.   TMEM: [0,128) main0 | [128,256) main1 | [256,384) corr | [384,512) sum
// (body of wgrad_kernel; cta_x / ncta_x / cta_y as in linear_rows_ts_body.  PAIR: the CTA is the wgrad half of a reverse-pass
// pair and walks the 64-row macro tiles of its partner -- 32-row tiles 2T, 2T+1 for T = cta_x, cta_x + ncta_x, ... -- with the
// lock-step throttle in its TMA warp)
// GLD (TS mode): the G warps read their rows of dZ straight from global memory into registers, two tiles ahead, instead of
// through the TMA ring -- G goes to tensor memory anyway, so its trip through shared memory (a TMA write and a shared-memory
// read of 16 KB per tile, 256 of the ~1 150 wavefronts the shared-memory port carries per tile) bought nothing.  These warps
// execute no generic->async proxy fence (their operand path is tcgen05.st), so loads in flight do not stall on one.
template <int TK, int RS, int OS, int NCW, int SEG, bool TSA, bool PAIR = false, bool GLD = false>
__device__ __forceinline__ void
wgrad_body(const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx, float* __restrict__ dW, int lddw,
           float* __restrict__ db, int64_t M, int jet_cols, int in_blocks, const CUtensorMap* tmgp, const CUtensorMap* tmxmp,
           float* __restrict__ det_part, float* __restrict__ det_bpart, const int cta_x, const int ncta_x, const int cta_y,
           const PairSync ps) {
  // det_part != null (PINNK_DETERMINISTIC=1): every CTA stores its 128 x 128 partial sum to its own slab
  // det_part[(blockIdx.y * gridDim.x + blockIdx.x) * 16384] (and its bias partials to det_bpart[cta][2][128]) instead of
  // reducing into dW / db with atomics in arrival order; wgrad_det_reduce_kernel then adds the slabs in CTA order, so the
  // gradient is bit-identical from run to run.
  static_assert(TK == 32 && (NCW == 8 || NCW == 16), "tile shape");
  // TSA: the G^T operand (A of the MMA: 128 out-features x 32 rows) lives in TENSOR MEMORY instead of shared memory.  An
  // SS-mode 128x128x8 MMA reads 8 KB of operands from shared memory in its 64 cycles -- all of the SM's 128 B/cycle -- and
  // the staging traffic (TMA 32 KB + LDS 32 KB + STS 64 KB per 32-row tile) has to share that port: 224 KB per tile =
  // 1750 cycles, which is what the kernel measured (1600 cycles per tile at any clock; two convert groups changed nothing).
  // With A in TMEM the G warps write tcgen05.st instead of STS and the MMA fetches only B: 144 KB per tile.
  // TMEM (TSA): [0,256) main0|main1 | [256,384) sum | [384,512) 2 stages x {G_hi 32, G_lo 32}; the lo*hi + hi*lo corrections
  // go to the segment's main accumulator (no room for their own; a segment is 48 accumulate steps instead of 16).
  static_assert(!TSA || (OS == 2 && NCW == 16), "TS-mode wgrad: two operand stages, 8 G + 8 X convert warps");
  static_assert(!GLD || TSA, "direct G loads need the tensor-memory operand path");
  constexpr uint32_t OP_BYTES = TK * 512;          // one of G_hi / G_lo / X_hi / X_lo per operand stage (TK rows x 128 floats)
  constexpr uint32_t RAW_BYTES = TK * 512;         // raw G (or X) tile
  // warp roles: [0, NCW) convert | NCW..NCW+3 flush (TMEM lane quarters; NCW % 4 == 0) | NCW+4 MMA | NCW+5 TMA | 2 idle
  constexpr int EPI0 = NCW, MMAW = NCW + 4, TMAW = NCW + 5;
  constexpr uint32_t COL_MAIN = 0, COL_CORR = TSA ? 0 : 256, COL_SUM = TSA ? 256 : 384, COL_A = 384;   // (no COL_CORR with TSA)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int OPS = TSA ? 2 : 4;                          // operand tiles per stage: TS mode keeps only X_hi | X_lo here
  uint8_t* op_base = smem;                                  // [OS][G_hi|G_lo|X_hi|X_lo][OP_BYTES]   (TSA: [OS][X_hi|X_lo])
  uint8_t* raw_base = smem + (size_t)OS * OPS * OP_BYTES;   // [RS][G|X][RAW_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (size_t)RS * 2 * RAW_BYTES);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = bars + RS;
  uint64_t* full = bars + 2 * RS;
  uint64_t* empty = bars + 2 * RS + OS;
  uint64_t* tfull = bars + 2 * RS + 2 * OS;          // [2] segment accumulated
  uint64_t* tempty = tfull + 2;                      // [2] segment flushed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

#ifdef PINNK_STAGE_TIMERS
  unsigned long long g_entry; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = (cta_y / in_blocks) * 128, i0 = (cta_y % in_blocks) * 128;
  const int64_t ntiles = (M + TK - 1) / TK;
  // tile handled in iteration `it` of this CTA, and how many it handles
  auto tile_at = [&](int64_t it) -> int64_t {
    if constexpr (PAIR) return 2 * ((int64_t)cta_x + (it >> 1) * ncta_x) + (it & 1);
    else return (int64_t)cta_x + it * ncta_x;
  };
  int64_t my_tiles;
  if constexpr (PAIR) {
    const int64_t nmac = (ntiles + 1) / 2;
    const int64_t my_mac = (nmac > (int64_t)cta_x) ? (nmac - cta_x + ncta_x - 1) / ncta_x : 0;
    my_tiles = 2 * my_mac;
    if (my_mac > 0 && 2 * ((int64_t)cta_x + (my_mac - 1) * ncta_x) + 1 >= ntiles) --my_tiles;     // odd tail
  } else {
    my_tiles = (ntiles > (int64_t)cta_x) ? (ntiles - cta_x + ncta_x - 1) / ncta_x : 0;
  }
  const int64_t my_segs = (my_tiles + SEG - 1) / SEG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], GLD ? NCW / 2 : NCW); }   // (GLD: X warps only)
    for (int s = 0; s < OS; ++s) { mbar_init(&full[s], NCW); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    fence_mbar_init();
  }
  if (warp == MMAW) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TMAW) {
    // ===================== TMA producer: raw row tiles, one elected thread =====================
    if (elect_one()) {
      const uint32_t rb = smem_u32(raw_base);
      long long t_a = 0; (void)t_a;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t tile = tile_at(it);
        const int s = (int)(it % RS);
        const uint32_t ph = (uint32_t)(it / RS) & 1u;
        { PK_T0(); mbar_wait(&raw_empty[s], ph ^ 1u); PK_TACC(t_a); }
        if constexpr (PAIR) { if ((it & 1) == 0) ps.wait_turn((uint32_t)(it >> 1)); }
        const int64_t r0 = tile * TK;
        const int nrows = (M - r0 >= TK) ? TK : (int)(M - r0);
        const uint32_t dg = rb + (uint32_t)s * 2 * RAW_BYTES, dx = dg + RAW_BYTES;
        // contiguous 128-wide tensors: one bulk copy; 128-column blocks of wider tensors: one tiled copy (box [TK x 128],
        // rows past M arrive as zeros and count in the transaction bytes)
        mbar_arrive_expect_tx(&raw_full[s], (GLD ? 0u : (uint32_t)(ldg == 128 ? nrows : TK) * 512u) + (uint32_t)(ldx == 128 ? nrows : TK) * 512u);
        if constexpr (!GLD) {
          if (ldg == 128) tma_bulk_g2s(dg, G + r0 * 128, (uint32_t)nrows * 512u, &raw_full[s]);
          else tma_tile_2d(dg, tmgp, o0, (int)r0, &raw_full[s]);
        }
        if (ldx == 128) tma_bulk_g2s(dx, X + r0 * 128, (uint32_t)nrows * 512u, &raw_full[s]);
        else tma_tile_2d(dx, tmxmp, i0, (int)r0, &raw_full[s]);
        if constexpr (PAIR) { if (it & 1) ps.publish((uint32_t)(it >> 1) + 1u); }
      }
      if constexpr (PAIR) ps.publish(0x7fffffffu);        // done: never hold the partner back
#ifdef PINNK_STAGE_TIMERS
      if (cta_x == 0 && cta_y == 0) atomicAdd(&g_stage_timers[0], (unsigned long long)t_a);
#endif
    }
    __syncwarp();
  } else if (warp < NCW) {
    // ===================== convert warps: raw [rows x 128] tile -> TRANSPOSED hi/lo operand tile [128 features x TK] ====
    // Both UMMA operands are K-major (K = rows of the tile): feature f owns one 128-byte swizzle row holding its TK = 32
    // row values.  Warp w converts row quad (w % 8) of tensor (w / 8: 0 = G, 1 = X); lane l owns features l, l+32, l+64,
    // l+96, so the four values of a feature for the quad are one 16-byte chunk and a quarter-warp's STS.128 hit eight
    // distinct 16-byte slots (conflict-free), while the raw reads are lane-contiguous LDS.32.
    static_assert(NCW == 16 && TK == 32, "convert mapping");
    if (TSA && GLD && warp < 8) {
      // ---- G warps, TS mode, direct loads: warp (q, h) owns out-features f = 32q + lane and rows 16h .. 16h + 15 of a tile; its
      // 16 values per lane come from global memory (each warp instruction reads 128 contiguous bytes of one row) into one of
      // two register tiles, requested two tiles before they are converted
      const int q = warp & 3, h = warp >> 2, f = q * 32 + lane;
      float bsum1 = 0.f;
      const bool want_b1 = (db != nullptr) && (i0 == 0);
      const bool b_fixed1 = (jet_cols == 1 || jet_cols == 2 || jet_cols == 4);
      const uint32_t b_mask1 = (jet_cols == 1) ? 0xFu : (jet_cols == 2) ? 0x5u : 0x1u;
      const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16) + COL_A + (uint32_t)(h * 16);
      const float* const gcol = G + o0 + f;
      auto gload = [&](const int64_t it, float (&dst)[16]) {
        if (it < my_tiles) {
          const int64_t r0 = tile_at(it) * TK + h * 16;
          const float* gp = gcol + r0 * ldg;
          if (M - r0 >= 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = __ldg(gp + (int64_t)j * ldg);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j] = (r0 + j < M) ? __ldg(gp + (int64_t)j * ldg) : 0.f;
          }
        }
      };
      int os = 0;
      uint32_t oph = 0;
      auto gconvert = [&](const int64_t it, const float (&v)[16]) {
        mbar_wait(&empty[os], oph ^ 1u);
        tc_fence_after();
        if (want_b1) {
          if (b_fixed1) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((b_mask1 >> (j & 3)) & 1u) bsum1 += v[j];
          } else {
            uint32_t cj = (uint32_t)((uint32_t)(tile_at(it) * TK + h * 16) % (uint32_t)jet_cols);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (cj == 0) bsum1 += v[j];
              cj = (cj + 1 == (uint32_t)jet_cols) ? 0u : cj + 1;
            }
          }
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_bits(v[j], hi[j], lo[j]);
        tmem_st16(a_lane + (uint32_t)os * 64u, hi);
        tmem_st16(a_lane + (uint32_t)os * 64u + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[os]);
        if (++os == OS) { os = 0; oph ^= 1u; }
      };
      float va[16], vb[16];
      gload(0, va);
      gload(1, vb);
      for (int64_t it = 0; it < my_tiles; it += 2) {
        gconvert(it, va);
        gload(it + 2, va);
        if (it + 1 < my_tiles) {
          gconvert(it + 1, vb);
          gload(it + 3, vb);
        }
      }
      if (want_b1) {
        if (det_bpart) det_bpart[((int64_t)(cta_y * ncta_x + cta_x) * 2 + h) * 128 + f] = bsum1;
        else atomicAdd(db + o0 + f, bsum1);
      }
    } else if (TSA && warp < 8) {
      // ---- G warps, TS mode: warp (q = warp & 3, h = warp >> 2) owns TMEM lanes 32q.. (out-features f = 32q + lane) and
      // rows 16h .. 16h + 15 of the tile: 16 lane-contiguous LDS.32, hi/lo split, two tcgen05.st of 16 columns
      const int q = warp & 3, h = warp >> 2, f = q * 32 + lane;
      float bsum1 = 0.f;
      const bool want_b1 = (db != nullptr) && (i0 == 0);
      const bool b_fixed1 = (jet_cols == 1 || jet_cols == 2 || jet_cols == 4);
      const uint32_t b_mask1 = (jet_cols == 1) ? 0xFu : (jet_cols == 2) ? 0x5u : 0x1u;
      const uint32_t lane_raw1 = smem_u32(raw_base) + (uint32_t)(h * 16) * 512u + (uint32_t)f * 4u;
      const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16) + COL_A + (uint32_t)(h * 16);
      int rs = 0, os = 0;
      uint32_t rph = 0, oph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t tile = tile_at(it);
        const int64_t r0 = tile * TK;
        const int nrows = (M - r0 >= TK) ? TK : (int)(M - r0);
        mbar_wait(&raw_full[rs], rph);
        const uint32_t raw = lane_raw1 + (uint32_t)rs * 2 * RAW_BYTES;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = 0.f;
          if (nrows == TK || h * 16 + j < nrows) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(raw + (uint32_t)(j * 512)));
          v[j] = x;
        }
        uint32_t dep = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) dep |= __float_as_uint(v[j]);
        __syncwarp();
        if (lane == 0) mbar_arrive_after_loads(&raw_empty[rs], dep, tmem_slot + 1);
        mbar_wait(&empty[os], oph ^ 1u);
        tc_fence_after();
        if (want_b1) {
          if (b_fixed1) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((b_mask1 >> (j & 3)) & 1u) bsum1 += v[j];
          } else {
            uint32_t cj = (uint32_t)((uint32_t)(r0 + h * 16) % (uint32_t)jet_cols);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (cj == 0) bsum1 += v[j];
              cj = (cj + 1 == (uint32_t)jet_cols) ? 0u : cj + 1;
            }
          }
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_bits(v[j], hi[j], lo[j]);
        tmem_st16(a_lane + (uint32_t)os * 64u, hi);
        tmem_st16(a_lane + (uint32_t)os * 64u + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[os]);
        if (++rs == RS) { rs = 0; rph ^= 1u; }
        if (++os == OS) { os = 0; oph ^= 1u; }
      }
      if (want_b1) {
        if (det_bpart) det_bpart[((int64_t)(cta_y * ncta_x + cta_x) * 2 + h) * 128 + f] = bsum1;
        else atomicAdd(db + o0 + f, bsum1);
      }
    } else {
    const int rq = warp & 7, sel = warp >> 3;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool want_b = (db != nullptr) && (i0 == 0) && sel == 0;
    // value rows of this warp's quad: tiles start at multiples of 32 rows, so for jet_cols in {1, 2, 4} the pattern is fixed
    const bool b_fixed = (jet_cols == 1 || jet_cols == 2 || jet_cols == 4);
    const uint32_t b_mask = (jet_cols == 1) ? 0xFu : (jet_cols == 2) ? 0x5u : 0x1u;
    const uint32_t ob = smem_u32(op_base), rb = smem_u32(raw_base);
    // everything that does not depend on the tile is hoisted: raw read offset of this lane, swizzled store offsets
    const uint32_t lane_raw = (uint32_t)sel * RAW_BYTES + (uint32_t)(rq * 4) * 512u + (uint32_t)lane * 4u;
    uint32_t off_e[4];                                        // K-major SW128: row = feature, 16-byte chunk = row quad
#pragma unroll
    for (int e = 0; e < 4; ++e) off_e[e] = (TSA ? 0u : (uint32_t)sel * 2 * OP_BYTES) + sw128_offset(128, lane + 32 * e, rq);
    long long t_a = 0, t_b = 0, t_c = 0; (void)t_a; (void)t_b; (void)t_c;
    int rs = 0, os = 0;
    uint32_t rph = 0, oph = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = tile_at(it);
      const int64_t r0 = tile * TK;
      const int nrows = (M - r0 >= TK) ? TK : (int)(M - r0);
      { PK_T0(); mbar_wait(&raw_full[rs], rph); PK_TACC(t_a); }
      PK_T0();
      const uint32_t raw = rb + (uint32_t)rs * 2 * RAW_BYTES + lane_raw;
      float v[4][4];                                       // [row of the quad][feature e]
      if (nrows == TK) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[r][e]) : "r"(raw + (uint32_t)(r * 512 + e * 128)));
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = 0.f;
            if (rq * 4 + r < nrows) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(raw + (uint32_t)(r * 512 + e * 128)));
            v[r][e] = x;
          }
      }
      uint32_t dep = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) dep |= __float_as_uint(v[r][e]);
      __syncwarp();
      if (lane == 0) mbar_arrive_after_loads(&raw_empty[rs], dep, tmem_slot + 1);    // raw data IS in registers: slot may be refilled
      PK_TACC(t_c);
      { PK_T0(); mbar_wait(&empty[os], oph ^ 1u); PK_TACC(t_b); }
      const long long _t1 = clock64(); (void)_t1;
      const uint32_t hi_base = ob + (uint32_t)os * OPS * OP_BYTES, lo_base = hi_base + OP_BYTES;
      if (want_b) {
        if (b_fixed) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if ((b_mask >> r) & 1u) { bsum[0] += v[r][0]; bsum[1] += v[r][1]; bsum[2] += v[r][2]; bsum[3] += v[r][3]; }
        } else {
          uint32_t cj = (uint32_t)((uint32_t)(r0 + rq * 4) % (uint32_t)jet_cols);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (cj == 0) { bsum[0] += v[r][0]; bsum[1] += v[r][1]; bsum[2] += v[r][2]; bsum[3] += v[r][3]; }
            cj = (cj + 1 == (uint32_t)jet_cols) ? 0u : cj + 1;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_bits(v[0][e], h0, l0); split_bits(v[1][e], h1, l1); split_bits(v[2][e], h2, l2); split_bits(v[3][e], h3, l3);
        sts128(hi_base + off_e[e], h0, h1, h2, h3);
        sts128(lo_base + off_e[e], l0, l1, l2, l3);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[os]);
#ifdef PINNK_STAGE_TIMERS
      t_c += clock64() - _t1;
#endif
      if (++rs == RS) { rs = 0; rph ^= 1u; }
      if (++os == OS) { os = 0; oph ^= 1u; }
    }
#ifdef PINNK_STAGE_TIMERS
    if (cta_x == 0 && cta_y == 0 && warp == 0 && lane == 0) {
      atomicAdd(&g_stage_timers[1], (unsigned long long)t_a); atomicAdd(&g_stage_timers[2], (unsigned long long)t_b);
      atomicAdd(&g_stage_timers[3], (unsigned long long)t_c);
    }
#endif
    if (want_b) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(db + o0 + lane + 32 * e, bsum[e]);
    }
    }   // (X warps / SS-mode G warps)
  } else if (warp < MMAW) {
    // ===================== flush warps: fold finished segments into the fp32 running sum, write out at the end ======
    if (my_segs > 0) {
      const int q = warp - EPI0;
      const int f = q * 32 + lane;                               // out-feature (TMEM lane)
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int64_t seg = 0; seg < my_segs; ++seg) {
        const int b = (int)(seg & 1);
        mbar_wait_lazy(&tfull[b], (uint32_t)(seg >> 1) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t p[32], a[32];
          tmem_ld32_nowait(lane_base + COL_MAIN + b * 128 + c0, p);
          if (seg > 0) tmem_ld32_nowait(lane_base + COL_SUM + c0, a);
          tmem_wait_ld(p);
          if (seg > 0) tmem_wait_ld(a);
          if (seg > 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) p[j] = __float_as_uint(__uint_as_float(p[j]) + __uint_as_float(a[j]));
          }
          tmem_st32(lane_base + COL_SUM + c0, p);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[b]);
      }
      // all MMAs (including the corrections) are complete: the last tfull commit covered them
      float* tr = reinterpret_cast<float*>(raw_base);            // the raw ring is idle now: [128][132] transpose buffer
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t p[32], a[32];
        tmem_ld32_nowait(lane_base + COL_SUM + c0, p);
        if constexpr (!TSA) tmem_ld32_nowait(lane_base + COL_CORR + c0, a);
        tmem_wait_ld(p);
        if constexpr (!TSA) tmem_wait_ld(a);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[f * 132 + c0 + j] = TSA ? __uint_as_float(p[j]) : __uint_as_float(p[j]) + __uint_as_float(a[j]);
      }
      named_bar_sync(1, kEpiThreads);
      const int t = threadIdx.x - EPI0 * 32;
      float* const dw0 = dW + (int64_t)o0 * lddw + i0;
      // columns of this block that exist (lddw = in_dim): a layer narrower than 128 inputs (the 64 Fourier features of fourier.py)
      // runs as one block whose missing input columns arrive as zeros from the tiled copy and are not written back
      const int ncol = (lddw - i0 < 128) ? lddw - i0 : 128;
      if (det_part != nullptr) {
        float4* const slab = reinterpret_cast<float4*>(det_part + (int64_t)(cta_y * ncta_x + cta_x) * 16384);
        for (int idx = t; idx < 128 * 32; idx += kEpiThreads)
          slab[idx] = *reinterpret_cast<const float4*>(tr + (idx >> 5) * 132 + (idx & 31) * 4);
      } else if (((reinterpret_cast<uintptr_t>(dw0) & 15) == 0) && (lddw & 3) == 0) {
        // 16-byte vector reductions: a quarter of the atomic operations (all CTAs add into the same 128 x 128 block)
        for (int idx = t; idx < 128 * 32; idx += kEpiThreads) {
          const int row = idx >> 5, c4 = (idx & 31) * 4;
          if (c4 >= ncol) continue;
          const float4 v = *reinterpret_cast<const float4*>(tr + row * 132 + c4);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw0 + (int64_t)row * lddw + c4), "f"(v.x), "f"(v.y),
                       "f"(v.z), "f"(v.w) : "memory");
        }
      } else {
        for (int idx = t; idx < 128 * 128; idx += kEpiThreads) {
          const int row = idx >> 7, col = idx & 127;
          if (col < ncol) atomicAdd(dw0 + (int64_t)row * lddw + col, tr[row * 132 + col]);
        }
      }
    }
  } else if (warp == MMAW) {
    // ===================== MMA issuer =====================
    if (my_segs > 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 128);
      const uint64_t dconst = make_desc_k_sw128(0);
      const uint32_t sbase = smem_u32(op_base) >> 4;
      long long t_a = 0, t_b = 0, t_c = 0, n_t = 0; const long long t_start = clock64();
      (void)t_a; (void)t_b; (void)t_c; (void)n_t; (void)t_start;
#ifdef PINNK_STAGE_TIMERS
      unsigned long long g_start; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
#endif
      int64_t seg = 0;
      int in_seg = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % OS);
        const uint32_t ph = (uint32_t)(it / OS) & 1u;
        const int b = (int)(seg & 1);
        if (in_seg == 0) { PK_T0(); mbar_wait(&tempty[b], ((uint32_t)(seg >> 1) & 1u) ^ 1u); PK_TACC(t_a); }    // previous use of this buffer flushed
        { PK_T0(); mbar_wait(&full[s], ph); PK_TACC(t_b); }
        PK_T0();
        tc_fence_after();
        const uint32_t gh = sbase + (uint32_t)s * (OPS * OP_BYTES >> 4), gl = gh + (OP_BYTES >> 4),
                       xh = TSA ? gh : gl + (OP_BYTES >> 4), xl = xh + (OP_BYTES >> 4);
        const uint32_t d_main = tmem_base + COL_MAIN + (uint32_t)b * 128, d_corr = tmem_base + COL_CORR;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint32_t o = (uint32_t)ks * (32 >> 4);                 // 8 K values = 32 bytes inside the swizzle row
          const uint64_t a_hi = dconst | (uint64_t)(gh + o), a_lo = dconst | (uint64_t)(gl + o);
          const uint64_t b_hi = dconst | (uint64_t)(xh + o), b_lo = dconst | (uint64_t)(xl + o);
          if constexpr (TSA) {
            const uint32_t ta_hi = tmem_base + COL_A + (uint32_t)s * 64u + (uint32_t)ks * 8u, ta_lo = ta_hi + 32u;
            umma_tf32_ts(d_main, ta_lo, b_hi, idesc, (in_seg | ks) ? 1u : 0u);
            umma_tf32_ts(d_main, ta_hi, b_lo, idesc, 1u);
            umma_tf32_ts(d_main, ta_hi, b_hi, idesc, 1u);
          } else {
            umma_tf32(d_corr, a_lo, b_hi, idesc, (it | ks) ? 1u : 0u);
            umma_tf32(d_corr, a_hi, b_lo, idesc, 1u);
            umma_tf32(d_main, a_hi, b_hi, idesc, (in_seg | ks) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        if (++in_seg == SEG || it + 1 == my_tiles) {
          umma_commit(&tfull[b]);
          in_seg = 0;
          ++seg;
        }
        PK_TACC(t_c);
        ++n_t;
      }
#ifdef PINNK_STAGE_TIMERS
      {   // slowest and fastest CTA of the launch (loop time): static tile assignment vs uneven memory paths
        unsigned long long g_e2; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_e2));
        atomicMax(&g_stage_timers[15], g_e2 - g_start);
      }
      if (cta_x == 0 && cta_y == 0) {
        atomicAdd(&g_stage_timers[4], (unsigned long long)t_a); atomicAdd(&g_stage_timers[5], (unsigned long long)t_b);
        atomicAdd(&g_stage_timers[6], (unsigned long long)t_c); atomicAdd(&g_stage_timers[9], (unsigned long long)n_t);
        atomicAdd(&g_stage_timers[10], (unsigned long long)(clock64() - t_start));
        unsigned long long g_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
        atomicAdd(&g_stage_timers[11], g_end - g_start);
        atomicAdd(&g_stage_timers[12], g_start - g_entry);     // kernel entry -> first MMA wait (prologue)
      }
#endif
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMAW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
#ifdef PINNK_STAGE_TIMERS
    if (cta_x == 0 && cta_y == 0 && lane == 0) {
      unsigned long long g_exit; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_exit));
      atomicAdd(&g_stage_timers[13], g_exit - g_entry);         // whole kernel, block 0
      atomicAdd(&g_stage_timers[14], 1ull);
    }
#endif
  }
}

template <int TK, int RS, int OS, int NCW, int SEG, bool TSA, bool GLD = false>
__global__ void __launch_bounds__((NCW + 8) * 32, 1)
wgrad_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx, float* __restrict__ dW, int lddw,
             float* __restrict__ db, int64_t M, int jet_cols, int in_blocks, const __grid_constant__ CUtensorMap tmg,
             const __grid_constant__ CUtensorMap tmxm, float* __restrict__ det_part, float* __restrict__ det_bpart) {
  wgrad_body<TK, RS, OS, NCW, SEG, TSA, false, GLD>(G, ldg, X, ldx, dW, lddw, db, M, jet_cols, in_blocks, &tmg, &tmxm, det_part, det_bpart,
                                                    (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, PairSync{});
}

// fixed-order reduction of the per-CTA partial slabs of a deterministic wgrad launch: dW[block] += sum_x part[block][x]
// (x = 0 .. gx-1 in order), db[o] += sum_x (bpart[x][0][o] + bpart[x][1][o]).  grid (16, blocks), 256 threads.
static __global__ void wgrad_det_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bpart, int gx,
                                               int in_blocks, float* __restrict__ dW, int lddw, float* __restrict__ db) {
  const int blk = blockIdx.y, o0 = (blk / in_blocks) * 128, i0 = (blk % in_blocks) * 128;
  const int idx = blockIdx.x * 256 + threadIdx.x;                    // float4 index inside the 128 x 128 block
  const int row = idx >> 5, c4 = (idx & 31) * 4;
  const float4* p = reinterpret_cast<const float4*>(part + (int64_t)blk * gx * 16384) + idx;
  // the partial sums are combined in double (exactly, for all practical purposes) and rounded once: fixed order AND closer to
  // the true sum than any fp32 ordering of the 148 partials
  double ax = 0.0, ay = 0.0, az = 0.0, aw = 0.0;
  for (int x = 0; x < gx; ++x) {
    const float4 v = p[(int64_t)x * 4096];
    ax += (double)v.x; ay += (double)v.y; az += (double)v.z; aw += (double)v.w;
  }
  if (i0 + c4 < lddw) {                                              // (in_dim is a multiple of 4: whole float4 groups exist or not)
    float* d = dW + (int64_t)(o0 + row) * lddw + i0 + c4;
    d[0] += (float)ax; d[1] += (float)ay; d[2] += (float)az; d[3] += (float)aw;
  }
  if (db != nullptr && i0 == 0 && idx < 128) {
    double b = 0.0;
    for (int x = 0; x < gx; ++x) {
      const float* q = bpart + ((int64_t)(blk * gx + x) * 2) * 128 + idx;
      b += (double)q[0];
      b += (double)q[128];
    }
    db[o0 + idx] += (float)b;
  }
}

template <int TK, int RS, int OS, int NCW, int SEG, bool TSA, bool GLD = false>
static int launch_wgrad(const float* G, const float* X, float* dW, float* db, int64_t M, int in_dim, int out_dim,
                        int jet_cols, int sm_count, cudaStream_t st, float* det_scratch = nullptr, int64_t det_floats = 0) {
  constexpr size_t smem = 1024 + (size_t)OS * (TSA ? 2 : 4) * TK * 512 + (size_t)RS * 2 * TK * 512 + (2 * RS + 2 * OS + 4) * 8 + 16;
  static_assert(smem <= 232448 && (size_t)RS * 2 * TK * 512 >= 128 * 132 * 4, "shared memory budget / transpose buffer");
  auto kern = wgrad_kernel<TK, RS, OS, NCW, SEG, TSA, GLD>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  const int64_t ntiles = (M + TK - 1) / TK;
  const int in_blocks = (in_dim + 127) / 128, blocks = in_blocks * (out_dim / 128);
  int gx = sm_count / blocks;
  if (gx < 1) gx = 1;
  if ((int64_t)gx > ntiles) gx = (int)ntiles;
  float* det_part = nullptr;
  float* det_bpart = nullptr;
  if (det_scratch != nullptr) {
    if (!TSA) return TC_UNSUPPORTED;
    // per CTA: a 128 x 128 slab + 2 x 128 bias partials; fewer CTAs when the scratch buffer is small (small chunks)
    const int64_t per_cta = 16384 + 256;
    const int64_t cap = det_floats / (per_cta * blocks);
    if (cap < 1) return TC_UNSUPPORTED;
    if ((int64_t)gx > cap) gx = (int)cap;
    det_part = det_scratch;
    det_bpart = det_scratch + (int64_t)gx * blocks * 16384;
  }
  dim3 grid((unsigned)gx, (unsigned)blocks, 1);
  alignas(64) CUtensorMap tmg, tmx;
  memset(&tmg, 0, sizeof(tmg));
  memset(&tmx, 0, sizeof(tmx));
  if (out_dim != 128 && !make_tmap_rows(&tmg, G, M, out_dim, out_dim, TK)) return -1;
  if (in_dim != 128 && !make_tmap_rows(&tmx, X, M, in_dim, in_dim, TK)) return -1;
  kern<<<grid, (NCW + 8) * 32, smem, st>>>(G, out_dim, X, in_dim, dW, in_dim, db, M, jet_cols, in_blocks, tmg, tmx, det_part,
                                           det_bpart);
  if (cudaGetLastError() != cudaSuccess) return -1;
  if (det_part != nullptr)
    wgrad_det_reduce_kernel<<<dim3(16, (unsigned)blocks, 1), 256, 0, st>>>(det_part, det_bpart, gx, in_blocks, dW, in_dim, db);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// Paired reverse kernel of one hidden Linear(128, 128) + tanh: dgrad + tanh adjoint AND the weight gradient in ONE launch.
// Both products read the same two tensors (dZ of this layer, the output jets Y of the previous one).  As separate kernels
// they pull them out of HBM twice (2560 B per row with the dZ_prev store).  Here the grid is 74 clusters of two CTAs: rank 0
// runs the rows kernel (EPI_ACTBWD_Y: dZ_prev = tanh'(.)^T (dZ W)), rank 1 the TS-mode wgrad kernel, both over the SAME
// 64-row macro tiles T = pair, pair + 74, ..., held in lock step by the DSMEM throttle (PairSync): the second reader of
// a tile hits L2, so a row costs 1536 B of HBM traffic.  Each role keeps its whole SM (512 TMEM columns, 225 KB of shared
// memory, its warp roles) -- nothing of the two pipelines had to shrink, which a single-CTA fusion would have required.
template <int K0, int K1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(832, 1)
bwd_pair_kernel(const float* __restrict__ dZ, const float* __restrict__ W, const float* __restrict__ Yprev,
                float* __restrict__ dZprev, float* __restrict__ dW, float* __restrict__ db, int64_t M, int jet_cols,
                const __grid_constant__ CUtensorMap tm_unused, uint32_t dyn_bytes, uint32_t lead) {
  // the throttle word lives in the last 16 bytes of the dynamic allocation: the same offset in both CTAs, past either role's
  // layout (no static shared memory: it would push the 1024-byte aligned dynamic region over the 227 KB budget)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t rank = cluster_ctarank();
  PairSync ps;
  ps.lead = lead;
  ps.my_word = smem_u32(smem_raw) + dyn_bytes - 16u;
  if (threadIdx.x == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(ps.my_word), "r"(0u) : "memory");
  __syncthreads();
  cluster_sync_all();                         // both words are zero before either CTA can publish into its partner
  ps.peer_word = mapa_u32(ps.my_word, rank ^ 1u);
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  if (rank == 0) {
    linear_rows_ts_body<true, EPI_ACTBWD_Y, 1, K0, K1, 8, 16, false, 128, false, true>(
        dZ, W, 128, nullptr, dZprev, M, 128, jet_cols, Yprev, nullptr, 1.f, 128, OutFuse{}, FirstLayer{}, TcLossFuse{},
        &tm_unused, pair, npairs, 0, ps);
  } else {
    wgrad_body<32, 5, 2, 16, 4, true, true>(dZ, 128, Yprev, 128, dW, 128, db, M, jet_cols, 1, &tm_unused, &tm_unused, nullptr,
                                            nullptr, pair, npairs, 0, ps);
  }
  cluster_sync_all();                         // no CTA may exit while its partner can still store into its shared memory
}

template <int K0, int K1>
static int launch_bwd_pair(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db, int64_t M,
                           int sm_count, cudaStream_t st) {
  constexpr size_t smem_rows = 1024 + (size_t)2 * 2 * 64 * 128 * 4 + (size_t)3 * 64 * 128 * 4 + (2 * 2 + 2 * 2 + 2 * 3) * 8 + 16 + 1024;
  constexpr size_t smem_wg = 1024 + (size_t)2 * 2 * 32 * 512 + (size_t)5 * 2 * 32 * 512 + (2 * 5 + 2 * 2 + 4) * 8 + 16;
  constexpr size_t smem = (smem_rows > smem_wg ? smem_rows : smem_wg) + 16;      // + the throttle word
  static_assert(smem <= 232448, "shared memory budget (227 KB per CTA)");
  static_assert(jets_rows_per_warp(1 + K0 + K1) == 16, "paired kernel: 64-row tiles only (1, 2 or 4 jet columns)");
  auto kern = bwd_pair_kernel<K0, K1>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = true;
  }
  const int64_t nmac = (M + 63) / 64;
  int pairs = sm_count / 2;
  if (pairs < 1) pairs = 1;
  if ((int64_t)pairs > nmac) pairs = (int)nmac;
  alignas(64) CUtensorMap tmu;
  memset(&tmu, 0, sizeof(tmu));
  uint32_t lead = 4;
  if (const char* e = getenv("PINNK_PAIR_LEAD")) { const int v = atoi(e); if (v > 0) lead = (uint32_t)v; }
  kern<<<dim3((unsigned)(2 * pairs), 1, 1), 832, smem, st>>>(dZ, W, Yprev, dZprev, dW, db, M, 1 + K0 + K1, tmu, (uint32_t)smem, lead);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace tc

#if defined(PINNK_TC_TU_FWD) || defined(PINNK_TC_TU_BWD) || defined(PINNK_TC_TU_BWD_Y) || defined(PINNK_TC_TU_BWD_FIRST)
// jet layouts the fused epilogues are instantiated for: (K0, K1) = orders of the (at most two) directions
template <bool TRANS_W, int EPI, int ACT>
static inline int tc_dispatch_jets(int k0, int k1, const float* X, const float* W, int ldw, const float* bias, float* Y,
                                   int64_t M, int n_cols, const float* Zs, float* Yact, float omega, int sm_count,
                                   cudaStream_t st, int ldx = 128, int accum = 0, tc::OutFuse of = tc::OutFuse{},
                                   tc::FirstLayer fl = tc::FirstLayer{}, const TcLossFuse* loss = nullptr) {
#define PK_TC_CASE(A, B)                                                                                         \
  if (k0 == A && k1 == B)                                                                                        \
    return tc::launch_linear_rows_ts<TRANS_W, EPI, ACT, A, B>(X, W, ldw, bias, Y, M, n_cols, 1 + A + B, Zs, Yact, omega, \
                                                              sm_count, st, ldx, accum, of, fl, loss);
  PK_TC_CASE(0, 0) PK_TC_CASE(1, 0) PK_TC_CASE(2, 1) PK_TC_CASE(3, 0)
  PK_TC_CASE(1, 1) PK_TC_CASE(3, 1) PK_TC_CASE(2, 2) PK_TC_CASE(4, 1)
#undef PK_TC_CASE
  return TC_UNSUPPORTED;
}
// A 256-wide contraction runs as two K halves of the K = 128 kernel: the first writes the partial result (plain
// epilogue, no bias), the second adds it in its epilogue (bias, activation ... as requested).
#endif

#if defined(PINNK_TC_TU_KS_FWD) || defined(PINNK_TC_TU_KS_BWD)
// K-split launches of the jet layouts the fused epilogues are instantiated for
template <bool TRANS_W, int EPI, int ACT>
static inline int tc_ks_dispatch_jets(int k0, int k1, const float* X, const float* W, int ldw, const float* bias, float* Y,
                                      int64_t M, int n_cols, const float* Zs, float* Yact, float omega, int sm_count,
                                      cudaStream_t st, tc::OutFuse of, float* ring, int64_t ring_floats) {
#define PK_KS_CASE(A, B)                                                                                          \
  if (k0 == A && k1 == B)                                                                                         \
    return tc::launch_linear_rows_ts_ksplit<TRANS_W, EPI, ACT, A, B>(X, W, ldw, bias, Y, M, n_cols, 1 + A + B, Zs, Yact, omega, \
                                                                     sm_count, st, of, ring, ring_floats);
  PK_KS_CASE(0, 0) PK_KS_CASE(1, 0) PK_KS_CASE(2, 1) PK_KS_CASE(3, 0)
  PK_KS_CASE(1, 1) PK_KS_CASE(3, 1) PK_KS_CASE(2, 2) PK_KS_CASE(4, 1)
#undef PK_KS_CASE
  return TC_UNSUPPORTED;
}
#endif

#ifdef PINNK_TC_TU_KS_FWD
// K = 256 forward Linear in one K-split launch (see linear_rows_ts_ksplit_kernel); TC_UNSUPPORTED -> caller runs two passes
int tc_ks_linear_fwd(const float* X, const float* W, const float* bias, float* Z, int64_t M, int N, int jet_cols, int sm_count,
                     cudaStream_t st, float* ring, int64_t ring_floats) {
  if (M < 1 || (N % 128) != 0) return TC_UNSUPPORTED;
  return tc::launch_linear_rows_ts_ksplit<false, tc::EPI_PLAIN, 1, 0, 0>(X, W, 256, bias, Z, M, N, jet_cols, nullptr, nullptr, 1.f,
                                                                         sm_count, st, tc::OutFuse{}, ring, ring_floats);
}
int tc_ks_linear_act_fwd(const float* X, const float* W, const float* bias, float* Z, float* Yact, int64_t M, int N, int k0,
                         int k1, int act, float omega, int sm_count, cudaStream_t st, const float* w_out, float* u_part,
                         float* ring, int64_t ring_floats) {
  if (M < 1 || (N % 128) != 0 || !tc_jets_supported(k0, k1) || (act != 1 && act != 2)) return TC_UNSUPPORTED;
  tc::OutFuse of;
  of.w_out = w_out;
  of.u_part = u_part;
  if (act == 1) return tc_ks_dispatch_jets<false, tc::EPI_ACT, 1>(k0, k1, X, W, 256, bias, Z, M, N, nullptr, Yact, 1.f, sm_count, st, of, ring, ring_floats);
  return tc_ks_dispatch_jets<false, tc::EPI_ACT, 2>(k0, k1, X, W, 256, bias, Z, M, N, nullptr, Yact, omega, sm_count, st, of, ring, ring_floats);
}
#endif

#ifdef PINNK_TC_TU_KS_BWD
// out_dim = 256 dgrad (+ activation adjoint from the stashed pre-activations) in one K-split launch
int tc_ks_linear_dgrad(const float* dZ, const float* W, float* dX, int64_t M, int in_dim, int sm_count, cudaStream_t st,
                       float* ring, int64_t ring_floats) {
  if (M < 1 || (in_dim % 128) != 0) return TC_UNSUPPORTED;
  return tc::launch_linear_rows_ts_ksplit<true, tc::EPI_PLAIN, 1, 0, 0>(dZ, W, in_dim, nullptr, dX, M, in_dim, 1, nullptr, nullptr, 1.f,
                                                                        sm_count, st, tc::OutFuse{}, ring, ring_floats);
}
int tc_ks_linear_dgrad_actbwd(const float* dZ, const float* W, const float* Zprev, float* dZprev, int64_t M, int in_dim, int k0,
                              int k1, int act, float omega, int sm_count, cudaStream_t st, float* ring, int64_t ring_floats) {
  if (M < 1 || (in_dim % 128) != 0 || !tc_jets_supported(k0, k1) || (act != 1 && act != 2)) return TC_UNSUPPORTED;
  if (act == 1) return tc_ks_dispatch_jets<true, tc::EPI_ACTBWD, 1>(k0, k1, dZ, W, in_dim, nullptr, dZprev, M, in_dim, Zprev, nullptr, 1.f, sm_count, st, tc::OutFuse{}, ring, ring_floats);
  return tc_ks_dispatch_jets<true, tc::EPI_ACTBWD, 2>(k0, k1, dZ, W, in_dim, nullptr, dZprev, M, in_dim, Zprev, nullptr, omega, sm_count, st, tc::OutFuse{}, ring, ring_floats);
}
#endif

#ifdef PINNK_TC_TU_FWD
int tc_stage_timers_fwd(unsigned long long* out16, int reset) {
#ifdef PINNK_STAGE_TIMERS
  if (cudaMemcpyFromSymbol(out16, tc::g_stage_timers, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc::g_stage_timers, z, sizeof(z)); }
  return 0;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  (void)reset;
  return 1;
#endif
}
// Z[M,N] = X[M,K] W[N,K]^T (+ bias on value-column rows).  Returns 0 when launched,
// TC_UNSUPPORTED when the shape is not covered (caller uses the exact-fp32 CUDA-core GEMM), <0 on error.
// PINNK_ENABLE_KSPLIT=1: 256-wide contractions as ONE launch of CTA pairs (linear_rows_ts_ksplit_kernel) instead of two K-half
// launches with the partial product through HBM.  Read per call so that tests can switch it.  Bit-identical to the two passes
// and NOT the default: measured slower (one 256 x 256 layer over 1.31 M rows: forward 1.65 ms vs 1.29, dgrad 1.30 vs 1.14;
// profiles/r02y_ksplit.md has the experiments that ruled out the hand-off, the cluster launch and the ring depth).
static inline bool ksplit_enabled(const float* ring) {
  if (ring == nullptr) return false;
  const char* e = getenv("PINNK_ENABLE_KSPLIT");
  return e && e[0] == '1';
}
int tc_linear_fwd(const float* X, const float* W, const float* bias, float* Z, int64_t M, int K, int N,
                                int jet_cols, int sm_count, cudaStream_t st, float* ring, int64_t ring_floats) {
  if (M < 1 || (N % 128) != 0) return TC_UNSUPPORTED;
  static int use_ss = -1;
  if (use_ss < 0) { const char* e = getenv("PINNK_TC_SS"); use_ss = (e && e[0] == '1') ? 1 : 0; }
  if (K == 128 && !use_ss) return tc::launch_linear_rows_ts<false, tc::EPI_PLAIN, 1, 0, 0>(X, W, K, bias, Z, M, N, jet_cols, nullptr, nullptr, 1.f, sm_count, st);
  if (K == 256 && ksplit_enabled(ring)) {
    const int rc = tc_ks_linear_fwd(X, W, bias, Z, M, N, jet_cols, sm_count, st, ring, ring_floats);
    if (rc != TC_UNSUPPORTED) return rc;
  }
  if (K == 256) {      // two K halves, the second accumulates onto the first and adds the bias
    int rc = tc::launch_linear_rows_ts<false, tc::EPI_PLAIN, 1, 0, 0>(X, W, K, nullptr, Z, M, N, jet_cols, nullptr, nullptr, 1.f, sm_count, st, K, 0);
    if (rc) return rc;
    return tc::launch_linear_rows_ts<false, tc::EPI_PLAIN, 1, 0, 0>(X + 128, W + 128, K, bias, Z, M, N, jet_cols, nullptr, nullptr, 1.f, sm_count, st, K, 1);
  }
  if (K == 128) return tc::launch_linear_rows<128, 32, 3, 3, 8, 3, false>(X, W, K, bias, Z, M, N, jet_cols, sm_count, st);
  if (K == 64) return tc::launch_linear_rows<64, 64, 4, 2, 8, 2, false>(X, W, K, bias, Z, M, N, jet_cols, sm_count, st);
  if (K > 256 && (K % 128) == 0 && K <= 1024 && !use_ss) {
    // wider contractions (the RL sampler's 512-wide Q-network): K / 128 passes of the K = 128 kernel, each adding its partial
    // product to Z; the last one adds the bias
    for (int k0 = 0; k0 < K; k0 += 128) {
      const bool last = k0 + 128 == K;
      int rc = tc::launch_linear_rows_ts<false, tc::EPI_PLAIN, 1, 0, 0>(X + k0, W + k0, K, last ? bias : nullptr, Z, M, N, jet_cols, nullptr,
                                                                       nullptr, 1.f, sm_count, st, K, k0 ? 1 : 0);
      if (rc) return rc;
    }
    return 0;
  }
  return TC_UNSUPPORTED;
}
// Forward Linear + activation jets in one kernel: Z = X W^T + b (stash), Yact = act(Z).  act: 1 tanh, 2 sin(omega z).
int tc_linear_act_fwd(const float* X, const float* W, const float* bias, float* Z, float* Yact, int64_t M, int K,
                                    int N, int k0, int k1, int act, float omega, int sm_count, cudaStream_t st,
                                    const float* w_out, float* u_part, const TcLossFuse* loss, float* ring, int64_t ring_floats) {
  // K: 128, 256 (two K halves / K-split), or one narrower block (multiple of 4: the Fourier network's first hidden layer, K = 64)
  const bool narrow = K < 128 && K >= 16 && (K % 4) == 0;
  if (M < 1 || (K != 128 && K != 256 && !narrow) || (N % 128) != 0 || !tc_jets_supported(k0, k1) || (act != 1 && act != 2)) return TC_UNSUPPORTED;
  if (Yact == nullptr && w_out == nullptr) return TC_UNSUPPORTED;      // nothing would be produced
  if (loss != nullptr && w_out == nullptr) return TC_UNSUPPORTED;
  if (narrow && (loss != nullptr)) return TC_UNSUPPORTED;
  if (K == 256 && loss == nullptr && ksplit_enabled(ring)) {
    const int rc = tc_ks_linear_act_fwd(X, W, bias, Z, Yact, M, N, k0, k1, act, omega, sm_count, st, w_out, u_part, ring, ring_floats);
    if (rc != TC_UNSUPPORTED) return rc;
  }
  tc::OutFuse of;
  of.w_out = w_out;
  of.u_part = u_part;
  int accum = 0;
  if (K == 256) {
    if (Z == nullptr) return TC_UNSUPPORTED;       // the partial result needs a buffer
    int rc = tc::launch_linear_rows_ts<false, tc::EPI_PLAIN, 1, 0, 0>(X, W, K, nullptr, Z, M, N, 1, nullptr, nullptr, 1.f, sm_count, st, K, 0);
    if (rc) return rc;
    X += 128; W += 128; accum = 1;
  }
  if (loss != nullptr && (act != 1 || K != 128)) return TC_UNSUPPORTED;
  if (act == 1) return tc_dispatch_jets<false, tc::EPI_ACT, 1>(k0, k1, X, W, K, bias, Z, M, N, nullptr, Yact, 1.f, sm_count, st, K, accum, of, tc::FirstLayer{}, loss);
  return tc_dispatch_jets<false, tc::EPI_ACT, 2>(k0, k1, X, W, K, bias, Z, M, N, nullptr, Yact, omega, sm_count, st, K, accum, of);
}
#endif

#ifdef PINNK_TC_TU_BWD
int tc_stage_timers_bwd(unsigned long long* out16, int reset) {
#ifdef PINNK_STAGE_TIMERS
  if (cudaMemcpyFromSymbol(out16, tc::g_stage_timers, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc::g_stage_timers, z, sizeof(z)); }
  return 0;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  (void)reset;
  return 1;
#endif
}
// dX[M,in] = dZ[M,out] W[out,in]   (W row-major [out,in])
static inline bool ksplit_enabled_bwd(const float* ring) {          // (see ksplit_enabled in the forward unit)
  if (ring == nullptr) return false;
  const char* e = getenv("PINNK_ENABLE_KSPLIT");
  return e && e[0] == '1';
}
int tc_linear_dgrad(const float* dZ, const float* W, float* dX, int64_t M, int in_dim, int out_dim,
                                  int sm_count, cudaStream_t st, float* ring, int64_t ring_floats) {
  if (M < 1 || (in_dim % 128) != 0) return TC_UNSUPPORTED;
  static int use_ss = -1;
  if (use_ss < 0) { const char* e = getenv("PINNK_TC_SS"); use_ss = (e && e[0] == '1') ? 1 : 0; }
  if (out_dim == 128 && !use_ss) return tc::launch_linear_rows_ts<true, tc::EPI_PLAIN, 1, 0, 0>(dZ, W, in_dim, nullptr, dX, M, in_dim, 1, nullptr, nullptr, 1.f, sm_count, st);
  if (out_dim == 256 && ksplit_enabled_bwd(ring)) {
    const int rc = tc_ks_linear_dgrad(dZ, W, dX, M, in_dim, sm_count, st, ring, ring_floats);
    if (rc != TC_UNSUPPORTED) return rc;
  }
  if (out_dim == 256) {
    int rc = tc::launch_linear_rows_ts<true, tc::EPI_PLAIN, 1, 0, 0>(dZ, W, in_dim, nullptr, dX, M, in_dim, 1, nullptr, nullptr, 1.f, sm_count, st, out_dim, 0);
    if (rc) return rc;
    return tc::launch_linear_rows_ts<true, tc::EPI_PLAIN, 1, 0, 0>(dZ + 128, W + (int64_t)128 * in_dim, in_dim, nullptr, dX, M, in_dim, 1, nullptr, nullptr, 1.f, sm_count, st, out_dim, 1);
  }
  if (out_dim == 128) return tc::launch_linear_rows<128, 32, 3, 3, 8, 3, true>(dZ, W, in_dim, nullptr, dX, M, in_dim, 1, sm_count, st);
  if (out_dim > 256 && (out_dim % 128) == 0 && out_dim <= 1024 && !use_ss) {
    // wider layers (the 512-wide residual network of the reference's YAML): out_dim / 128 passes, each adding its partial product
    for (int k0 = 0; k0 < out_dim; k0 += 128) {
      int rc = tc::launch_linear_rows_ts<true, tc::EPI_PLAIN, 1, 0, 0>(dZ + k0, W + (int64_t)k0 * in_dim, in_dim, nullptr, dX, M, in_dim, 1,
                                                                      nullptr, nullptr, 1.f, sm_count, st, out_dim, k0 ? 1 : 0);
      if (rc) return rc;
    }
    return 0;
  }
  return TC_UNSUPPORTED;
}
// dgrad + activation adjoint in one kernel: dZprev = act'(Zprev)^T (dZ W)
int tc_linear_dgrad_actbwd(const float* dZ, const float* W, const float* Zprev, float* dZprev, int64_t M,
                                         int in_dim, int out_dim, int k0, int k1, int act, float omega, int sm_count,
                                         cudaStream_t st, int from_y, float* ring, int64_t ring_floats) {
  if (from_y && act != 1) return TC_UNSUPPORTED;
  if (M < 1 || (out_dim != 128 && out_dim != 256) || (in_dim % 128) != 0 || !tc_jets_supported(k0, k1) || (act != 1 && act != 2))
    return TC_UNSUPPORTED;
  if (out_dim == 256 && !from_y && ksplit_enabled_bwd(ring)) {
    const int rc = tc_ks_linear_dgrad_actbwd(dZ, W, Zprev, dZprev, M, in_dim, k0, k1, act, omega, sm_count, st, ring, ring_floats);
    if (rc != TC_UNSUPPORTED) return rc;
  }
  int accum = 0;
  if (out_dim == 256) {
    int rc = tc::launch_linear_rows_ts<true, tc::EPI_PLAIN, 1, 0, 0>(dZ, W, in_dim, nullptr, dZprev, M, in_dim, 1, nullptr, nullptr, 1.f, sm_count, st, out_dim, 0);
    if (rc) return rc;
    dZ += 128; W += (int64_t)128 * in_dim; accum = 1;
  }
  if (act == 1 && from_y) return tc_dgrad_actbwd_y(k0, k1, dZ, W, in_dim, dZprev, M, Zprev, sm_count, st, out_dim, accum);
  if (act == 1) return tc_dispatch_jets<true, tc::EPI_ACTBWD, 1>(k0, k1, dZ, W, in_dim, nullptr, dZprev, M, in_dim, Zprev, nullptr, 1.f, sm_count, st, out_dim, accum);
  return tc_dispatch_jets<true, tc::EPI_ACTBWD, 2>(k0, k1, dZ, W, in_dim, nullptr, dZprev, M, in_dim, Zprev, nullptr, omega, sm_count, st, out_dim, accum);
}
#endif

#ifdef PINNK_TC_TU_BWD_Y
int tc_stage_timers_bwd_y(unsigned long long* out16, int reset) {
#ifdef PINNK_STAGE_TIMERS
  if (cudaMemcpyFromSymbol(out16, tc::g_stage_timers, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc::g_stage_timers, z, sizeof(z)); }
  return 0;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  (void)reset;
  return 1;
#endif
}
// the output-jet variant of the fused dgrad + tanh adjoint (own translation unit: compile time)
int tc_dgrad_actbwd_y(int k0, int k1, const float* dZ, const float* W, int in_dim, float* dZprev, int64_t M,
                      const float* Yprev, int sm_count, cudaStream_t st, int out_dim, int accum) {
  return tc_dispatch_jets<true, tc::EPI_ACTBWD_Y, 1>(k0, k1, dZ, W, in_dim, nullptr, dZprev, M, in_dim, Yprev, nullptr, 1.f, sm_count, st, out_dim, accum);
}
#endif

#ifdef PINNK_TC_TU_BWD_FIRST
// dgrad of the first hidden layer + the whole reverse of the input layer (EPI_FIRSTBWD)
int tc_linear_dgrad_firstbwd(const float* dZ, const float* W, int64_t M, int in_dim, int out_dim, int k0, int k1, int act,
                             float omega, const float* x, const float* t, int net_in_dim, const float* vec0,
                             const float* vec1, const float* W0, const float* b0, float* gW0, float* gb0, int sm_count,
                             cudaStream_t st) {
  if (M < 1 || (out_dim != 128 && out_dim != 256) || (in_dim % 128) != 0 || !tc_jets_supported(k0, k1) || (act != 1 && act != 2) ||
      net_in_dim < 1 || net_in_dim > 4)
    return TC_UNSUPPORTED;
  tc::FirstLayer fl;
  fl.x = x; fl.t = t; fl.W0 = W0; fl.b0 = b0; fl.gW0 = gW0; fl.gb0 = gb0; fl.in_dim = net_in_dim;
  for (int i = 0; i < 4; ++i) { fl.vec[0][i] = vec0 ? vec0[i] : 0.f; fl.vec[1][i] = vec1 ? vec1[i] : 0.f; }
  if (out_dim == 256) return TC_UNSUPPORTED;      // (two K halves would need a partial-sum buffer: not wired up)
  if (act == 1) return tc_dispatch_jets<true, tc::EPI_FIRSTBWD, 1>(k0, k1, dZ, W, in_dim, nullptr, nullptr, M, in_dim, nullptr, nullptr, 1.f, sm_count, st, out_dim, 0, tc::OutFuse{}, fl);
  return tc_dispatch_jets<true, tc::EPI_FIRSTBWD, 2>(k0, k1, dZ, W, in_dim, nullptr, nullptr, M, in_dim, nullptr, nullptr, omega, sm_count, st, out_dim, 0, tc::OutFuse{}, fl);
}
#endif

#ifdef PINNK_TC_TU_WGRAD
int tc_stage_timers_wgrad(unsigned long long* out16, int reset) {
#ifdef PINNK_STAGE_TIMERS
  if (cudaMemcpyFromSymbol(out16, tc::g_stage_timers, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc::g_stage_timers, z, sizeof(z)); }
  return 0;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  (void)reset;
  return 1;
#endif
}
// dW[out,in] += dZ[M,out]^T X[M,in] ;  db[out] += sum over value-column rows of dZ
int tc_linear_wgrad(const float* dZ, const float* X, float* dW, float* db, int64_t M, int in_dim, int out_dim,
                                  int jet_cols, int sm_count, cudaStream_t st, float* det_scratch, int64_t det_floats) {
  // in_dim: multiples of 128, or one narrower block (multiple of 4, e.g. the 64 Fourier features feeding the first hidden layer)
  if (M < 1 || !((in_dim % 128) == 0 || (in_dim < 128 && (in_dim % 4) == 0 && in_dim >= 16)) || (out_dim % 128) != 0 || dW == nullptr)
    return TC_UNSUPPORTED;
  // PINNK_WGRAD_GLD=1: the G operand straight from global memory into registers instead of through the TMA ring (A/B; read per call)
  { const char* e = getenv("PINNK_WGRAD_GLD");
    if (e && e[0] == '1')
      return tc::launch_wgrad<32, 5, 2, 16, 4, true, true>(dZ, X, dW, db, M, in_dim, out_dim, jet_cols, sm_count, st, det_scratch, det_floats); }
  if (det_scratch != nullptr)
    return tc::launch_wgrad<32, 5, 2, 16, 4, true>(dZ, X, dW, db, M, in_dim, out_dim, jet_cols, sm_count, st, det_scratch, det_floats);
  static int ss = -1;       // PINNK_WGRAD_SS=1: both operands in shared memory (the earlier kernel, kept for A/B runs)
  if (ss < 0) { const char* e = getenv("PINNK_WGRAD_SS"); ss = (e && e[0] == '1') ? 1 : 0; }
  if (ss) return tc::launch_wgrad<32, 3, 2, 16, 4, false>(dZ, X, dW, db, M, in_dim, out_dim, jet_cols, sm_count, st);
  // TS mode frees the G operand tiles in shared memory: the raw ring is 5 deep (160 KB of loads in flight per SM)
  static int rs3 = -1;
  if (rs3 < 0) { const char* e = getenv("PINNK_WGRAD_RS3"); rs3 = (e && e[0] == '1') ? 1 : 0; }
  if (rs3) return tc::launch_wgrad<32, 3, 2, 16, 4, true>(dZ, X, dW, db, M, in_dim, out_dim, jet_cols, sm_count, st);
  return tc::launch_wgrad<32, 5, 2, 16, 4, true>(dZ, X, dW, db, M, in_dim, out_dim, jet_cols, sm_count, st);
}
#endif
#ifdef PINNK_TC_TU_PAIR
// dZprev[M,128] = tanh'(from the output jets Yprev)^T (dZ[M,128] W[128,128]);  dW += dZ^T Yprev;  db += value rows of dZ
int tc_bwd_pair(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db, int64_t M, int in_dim,
                int out_dim, int k0, int k1, int sm_count, cudaStream_t st) {
  if (M < 64 || in_dim != 128 || out_dim != 128 || dW == nullptr || sm_count < 2) return TC_UNSUPPORTED;
  if (k0 == 0 && k1 == 0) return tc::launch_bwd_pair<0, 0>(dZ, W, Yprev, dZprev, dW, db, M, sm_count, st);
  if (k0 == 1 && k1 == 0) return tc::launch_bwd_pair<1, 0>(dZ, W, Yprev, dZprev, dW, db, M, sm_count, st);
  if (k0 == 2 && k1 == 1) return tc::launch_bwd_pair<2, 1>(dZ, W, Yprev, dZprev, dW, db, M, sm_count, st);
  if (k0 == 3 && k1 == 0) return tc::launch_bwd_pair<3, 0>(dZ, W, Yprev, dZprev, dW, db, M, sm_count, st);
  return TC_UNSUPPORTED;
}
#endif
}  // namespace pinnk
