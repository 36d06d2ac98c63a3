// Blackwell (sm_100a) tensor-core plumbing used by the FC kernels: tcgen05.mma
// with shared-memory operands and TMEM accumulators, TMEM allocation and loads,
// mbarrier completion tracking.  Hand-written PTX wrappers; no library code.
//
// Numerics: the reference computes its FCs in float32.  bf16 tensor-core
// operands carry 8 significant bits, so every fp32 operand x is split into
// hi = bf16(x) and lo = bf16(x - hi) (|x - hi - lo| <= 2^-17 |x|) and a product
// is evaluated as  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  with fp32 accumulation in
// TMEM ("bf16x3").  The dropped a_lo*b_lo term and the split residuals are
// ~2^-16 relative per product, which keeps the logits inside the 1e-4 relative
// budget of BASELINE.json (measured: tests/test_gpu_forward.py).
//
// Shared-memory operand layout (both A[M,K] and B[N,K], K-major, no swizzle =
// the "interleaved" canonical layout): 16-byte chunks of 8 bf16 along K; a core
// matrix is 8 rows x 16 B stored contiguously (128 B); chunk j of row r lives at
//     base + j * LBO + (r / 8) * SBO + (r % 8) * 16
// One tcgen05.mma consumes K = 16 (two chunks); advancing K by 16 moves the
// descriptor start address by 2 * LBO.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef GN_MBAR_SLEEP_NS
#define GN_MBAR_SLEEP_NS 40
#endif

namespace gn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- programmatic dependent launch ------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream is still running (its CTAs are placed as the predecessor's
// retire); griddep_wait() returns once the predecessor has completed and its writes are
// visible.  Everything before it (barrier init, TMEM allocation, operand images that were
// written long before) overlaps the predecessor's tail.  Both are no-ops in a normal launch.
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- mbarrier -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
// Bounded wait: a malformed descriptor must end as a reported launch failure
// (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}

// The same wait for warps that are NOT on the tensor-core issue path: back off between
// polls, so that 20-odd waiting warps do not take the issue slots the single UMMA-issuing
// thread of a warp-specialised kernel needs (measured: 75-100 cycles per tcgen05.mma issue
// with hot spin loops around it, profiles/r1_tc_kernels.md).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (uint32_t spin = 0;; ++spin) {
    __nanosleep(GN_MBAR_SLEEP_NS);
    if (mbar_try_wait(bar, parity)) return;
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bulk global -> shared copy (TMA, non-tensor) ----------------------------------
// One thread: arm `bar` with the byte count, then issue the copy; the barrier
// completes (phase flips) when all bytes have landed.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// shared -> global bulk copy (bulk async-group completion): issue, commit, then wait until
// the source shared memory has been read (…_read) or the copies are complete.
__device__ __forceinline__ void bulk_copy_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_group_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_group0() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- tensor-map TMA (cp.async.bulk.tensor) ------------------------------------------
// The tensor map (CUtensorMap, built on the host by gn::encode_tmap_2d, gn_tma.cuh) is a
// __grid_constant__ kernel parameter; `tmap` is its generic address.  Completion is
// reported to `bar` as box-bytes of complete_tx (out-of-bounds rows are zero filled and
// still counted).  L2 cache-policy words: createpolicy encodings used by CUTLASS.
constexpr uint64_t TMA_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t TMA_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t TMA_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2D tile: box {box0 (inner), box1 (rows)} at coordinates {c0 (inner element), c1 (row)}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int32_t c0,
                                            int32_t c1, uint64_t* bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1),
        "r"(smem_u32(bar)), "l"(hint) : "memory");
}
// shared -> global 2D tile store (bulk async-group completion; rows / columns outside the
// tensor are clipped).  The source box must have been made visible with fence_smem_to_async.
__device__ __forceinline__ void tma_store_2d(const void* tmap, int32_t c0, int32_t c1,
                                             uint32_t src_smem) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(src_smem) : "memory");
}
// gather4 (sm_100): four rows r0..r3 of a 2D tensor (box {box0, 1}) land as four consecutive
// box0-wide rows at dst_smem, swizzled like a tile load
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const void* tmap, int32_t c0,
                                            int32_t r0, int32_t r1, int32_t r2, int32_t r3,
                                            uint64_t* bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
      ".cta_group::1.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(r0), "r"(r1), "r"(r2),
        "r"(r3), "r"(smem_u32(bar)), "l"(hint) : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (layout_type 2): rows of 128 bytes
// (64 bf16 of K), 8-row groups 1024 bytes apart (SBO), 16-byte chunk j of row r stored at
// chunk position j ^ (r % 8) - the layout a SWIZZLE_128B tensor map writes.  The tile base
// must be 1024-byte aligned; a K = 16 step advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

// ---- TMEM ---------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) columns; the base address
// is written to *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy smem writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_smem_to_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// TMEM address: lane in bits [16,32), column in bits [0,16).  A warp may only
// touch lanes 32*(warp%4) .. +31.  32x32b.x16: thread i of the warp receives 16
// consecutive fp32 columns of lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 consecutive fp32 columns per thread (one instruction, one wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors ----------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle (layout_type 0), version 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int m, int n) {
  return (1u << 4)                      // D format: F32
         | (1u << 7)                    // A format: BF16
         | (1u << 10)                   // B format: BF16
         | ((uint32_t)(n >> 3) << 17)   // N / 8
         | ((uint32_t)(m >> 4) << 24);  // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]^T.  Issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T ("TS" form): A rows live in TMEM lanes, element k of
// row m in 32-bit column k / 2 (even k in the low half), so a K = 16 step is 8 columns.
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-converged forms: the WHOLE warp executes the statement with warp-uniform operands and
// elect.sync picks the lane that issues.  With `if (lane == 0)` around the plain forms the
// operands live in per-thread registers and the compiler wraps every tcgen05 instruction in a
// waterfall loop (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~9 instructions per UMMA); uniform
// operands stay in uniform registers.  elect.sync with a full mask always picks the same lane,
// so mma_commit_elect tracks the UMMAs issued through these forms.
__device__ __forceinline__ void mma_bf16_ss_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_bf16x3_elect(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo,
                                                 uint64_t b_hi, uint64_t b_lo, uint32_t a_off16,
                                                 uint32_t b_off16, uint32_t idesc, uint32_t accumulate) {
  mma_bf16_ss_elect(tmem_d, a_lo + a_off16, b_hi + b_off16, idesc, accumulate);
  mma_bf16_ss_elect(tmem_d, a_hi + a_off16, b_lo + b_off16, idesc, 1);
  mma_bf16_ss_elect(tmem_d, a_hi + a_off16, b_hi + b_off16, idesc, 1);
}

// 8 consecutive 32-bit columns of this thread's TMEM lane (32x32b.x8)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Arrive on `bar` when every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}

// The three UMMAs of one bf16x3 k-step.  Descriptors are kernel-lifetime constants
// plus a small start-address offset: `off16` values are byte offsets >> 4 added to
// the 14-bit start-address field (shared memory is < 256 KB, so the sum never
// carries into the neighbouring field).  Keeping the issue path this short matters:
// ONE thread issues every UMMA of a CTA, and ~100 scalar instructions per k-step
// (building descriptors from scratch) made that thread, not the tensor pipe, the
// bottleneck of the first version of these kernels (profiles/r1_tc_kernels.md).
__device__ __forceinline__ void mma_bf16x3(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo,
                                           uint64_t b_hi, uint64_t b_lo, uint32_t a_off16,
                                           uint32_t b_off16, uint32_t idesc, uint32_t accumulate) {
  mma_bf16_ss(tmem_d, a_lo + a_off16, b_hi + b_off16, idesc, accumulate);
  mma_bf16_ss(tmem_d, a_hi + a_off16, b_lo + b_off16, idesc, 1);
  mma_bf16_ss(tmem_d, a_hi + a_off16, b_hi + b_off16, idesc, 1);
}

// ---- fp32 -> (hi, lo) bf16 split --------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// two values -> packed hi pair and lo pair (element 0 in the low half)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace umma
}  // namespace gn
