// Diagnostic entry point: one 128 x 64 x K product on the tensor cores with the
// exact building blocks the FC kernels use (bf16x3 split operands in the
// no-swizzle K-major shared-memory layout, tcgen05.mma into TMEM, tcgen05.ld
// epilogue).  tests/test_gpu_umma.py checks it against a float64 product, so a
// wrong descriptor / layout assumption shows up here and not inside a fused kernel.
#include "gn_common.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int ST_M = 128, ST_N = 64, ST_KMAX = 256;

// c[128,64] = a[128,k] @ w[k,64]
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ a, const float* __restrict__ w,
                     float* __restrict__ c, int k) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  const int chunks = k / 8;
  // [chunk][row][8 bf16]
  uint4* a_hi = reinterpret_cast<uint4*>(smem);
  uint4* a_lo = a_hi + chunks * ST_M;
  uint4* b_hi = a_lo + chunks * ST_M;
  uint4* b_lo = b_hi + chunks * ST_N;

  if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_barrier_init();
  }
  // A: thread t owns row t
  for (int j = 0; j < chunks; ++j) {
    const float4 v0 = ldg4(a + (size_t)t * k + j * 8), v1 = ldg4(a + (size_t)t * k + j * 8 + 4);
    uint4 h, l;
    umma::split_bf16x2(v0.x, v0.y, h.x, l.x);
    umma::split_bf16x2(v0.z, v0.w, h.y, l.y);
    umma::split_bf16x2(v1.x, v1.y, h.z, l.z);
    umma::split_bf16x2(v1.z, v1.w, h.w, l.w);
    a_hi[j * ST_M + t] = h;
    a_lo[j * ST_M + t] = l;
  }
  // B[n][kk] = w[kk][n]: thread handles n = t % 64, chunks of parity t / 64
  for (int j = t >> 6; j < chunks; j += 2) {
    const int n = t & 63;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __ldg(w + (size_t)(j * 8 + e) * ST_N + n);
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    b_hi[j * ST_N + n] = h;
    b_lo[j * ST_N + n] = l;
  }
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base;

  if (t == 0) {
    const uint32_t idesc = umma::idesc_bf16_f32(ST_M, ST_N);
    const uint32_t lbo_a = ST_M * 16, lbo_b = ST_N * 16, sbo = 128;
    uint32_t acc = 0;
    for (int ks = 0; ks < k / 16; ++ks) {
      const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi) + ks * 2 * lbo_a, lbo_a, sbo);
      const uint64_t dal = umma::smem_desc(umma::smem_u32(a_lo) + ks * 2 * lbo_a, lbo_a, sbo);
      const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b, sbo);
      const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b, sbo);
      umma::mma_bf16_ss(tmem, dal, dbh, idesc, acc);
      umma::mma_bf16_ss(tmem, dah, dbl, idesc, 1);
      umma::mma_bf16_ss(tmem, dah, dbh, idesc, 1);
      acc = 1;
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();

#pragma unroll
  for (int c0 = 0; c0 < ST_N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) c[(size_t)t * ST_N + c0 + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace gn

extern "C" int gn_selftest_umma(const float* a, const float* w, float* c, int k,
                                gn_stream_t stream) {
  GN_REQUIRE(a && w && c, "gn_selftest_umma: null pointer");
  GN_REQUIRE(k >= 16 && k % 16 == 0 && k <= gn::ST_KMAX, "gn_selftest_umma: k=%d must be a "
             "multiple of 16 in [16, %d]", k, gn::ST_KMAX);
  GN_REQUIRE((((uintptr_t)a | (uintptr_t)w) & 15) == 0, "gn_selftest_umma: unaligned pointer");
  const int smem = (k / 8) * (gn::ST_M + gn::ST_N) * 16 * 2;
  cudaError_t e = cudaFuncSetAttribute(gn::umma_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_selftest_umma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  gn::umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, w, c, k);
  GN_CHECK_LAUNCH("gn_selftest_umma");
  return GN_OK;
}

// ---------------------------------------------------------------------------------
// Same product with the A operand in TENSOR MEMORY (tcgen05.mma "TS" form): each
// thread stores its row's bf16 hi / lo pairs with tcgen05.st; pins the A-in-TMEM layout.
// ---------------------------------------------------------------------------------
namespace gn {
__global__ void __launch_bounds__(128, 1)
umma_selftest_ts_kernel(const float* __restrict__ a, const float* __restrict__ w,
                        float* __restrict__ c, int k) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  const int chunks = k / 8;
  uint4* b_hi = reinterpret_cast<uint4*>(smem);
  uint4* b_lo = b_hi + chunks * ST_N;
  if (warp == 0) umma::tmem_alloc(&tmem_base, 512);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_barrier_init();
  }
  for (int j = t >> 6; j < chunks; j += 2) {
    const int n = t & 63;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __ldg(w + (size_t)(j * 8 + e) * ST_N + n);
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    b_hi[j * ST_N + n] = h;
    b_lo[j * ST_N + n] = l;
  }
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base;
  // D: columns [0,64); A hi: [64, 64 + k/2); A lo: [64 + k/2, 64 + k)
  const uint32_t tlane = (uint32_t)(warp * 32) << 16;
  const uint32_t ta_hi = tmem + 64, ta_lo = tmem + 64 + k / 2;
  for (int j = 0; j < chunks; j += 2) {   // 16 elements = 8 columns per store
    uint32_t h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x0 = __ldg(a + (size_t)t * k + j * 8 + 2 * e);
      const float x1 = __ldg(a + (size_t)t * k + j * 8 + 2 * e + 1);
      umma::split_bf16x2(x0, x1, h[e], l[e]);
    }
    umma::tmem_st8(ta_hi + tlane + j * 4, h);
    umma::tmem_st8(ta_lo + tlane + j * 4, l);
  }
  umma::tmem_st_wait();
  umma::tc_fence_before();
  __syncthreads();
  if (t == 0) {
    umma::tc_fence_after();
    const uint32_t idesc = umma::idesc_bf16_f32(ST_M, ST_N);
    const uint32_t lbo_b = ST_N * 16, sbo = 128;
    for (int ks = 0; ks < k / 16; ++ks) {
      const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b, sbo);
      const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b, sbo);
      umma::mma_bf16_ts(tmem, ta_lo + ks * 8, dbh, idesc, ks > 0);
      umma::mma_bf16_ts(tmem, ta_hi + ks * 8, dbl, idesc, 1);
      umma::mma_bf16_ts(tmem, ta_hi + ks * 8, dbh, idesc, 1);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
#pragma unroll
  for (int c0 = 0; c0 < ST_N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + tlane + c0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) c[(size_t)t * ST_N + c0 + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}
}  // namespace gn

extern "C" int gn_selftest_umma_ts(const float* a, const float* w, float* c, int k,
                                   gn_stream_t stream) {
  GN_REQUIRE(a && w && c, "gn_selftest_umma_ts: null pointer");
  GN_REQUIRE(k >= 16 && k % 16 == 0 && k <= gn::ST_KMAX, "gn_selftest_umma_ts: bad k=%d", k);
  const int smem = (k / 8) * gn::ST_N * 16 * 2;
  cudaError_t e = cudaFuncSetAttribute(gn::umma_selftest_ts_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_selftest_umma_ts: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  gn::umma_selftest_ts_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, w, c, k);
  GN_CHECK_LAUNCH("gn_selftest_umma_ts");
  return GN_OK;
}

// ---------------------------------------------------------------------------------
// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode, no-swizzle
// K-major operands) for a given N, issued back to back by one thread.  out[0] =
// cycles for `reps` UMMAs (clock64 around issue .. commit wait), out[1] = reps.
// ---------------------------------------------------------------------------------
namespace gn {
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int n, int reps, int distinct_b, long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5;
  // A: 2 chunks x 128 rows; B: distinct_b copies of 2 chunks x n rows
  for (int i = t; i < (2 * 128 * 16 + distinct_b * 2 * n * 16) / 4; i += 128)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;   // arbitrary finite bf16 pairs
  if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_barrier_init();
  }
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (t == 0) {
    const uint32_t idesc = umma::idesc_bf16_f32(128, n);
    const uint32_t sa = umma::smem_u32(smem), sb = sa + 2 * 128 * 16;
    const uint64_t da = umma::smem_desc(sa, 128 * 16, 128);
    const uint64_t db = umma::smem_desc(sb, n * 16, 128);
    const uint32_t bstep = (2 * n * 16) >> 4;
    const long long t0 = clock64();
    // tight issue loop: 4 UMMAs per iteration over (up to) 4 distinct B tiles
    const uint32_t m = (uint32_t)distinct_b - 1u;      // distinct_b in {1, 2, 4}
    const uint64_t db0 = db, db1 = db + (1u & m) * bstep, db2 = db + (2u & m) * bstep,
                   db3 = db + (3u & m) * bstep;
    umma::mma_bf16_ss(tmem, da, db0, idesc, 0);
    for (int r = 1; r + 4 <= reps; r += 4) {
      umma::mma_bf16_ss(tmem, da, db1, idesc, 1);
      umma::mma_bf16_ss(tmem, da, db2, idesc, 1);
      umma::mma_bf16_ss(tmem, da, db3, idesc, 1);
      umma::mma_bf16_ss(tmem, da, db0, idesc, 1);
    }
    umma::mma_commit(&bar);
    umma::mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
    out[1] = reps;
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}
}  // namespace gn

extern "C" int gn_selftest_umma_rate(int n, int reps, int distinct_b, int ctas, int64_t* out_dev,
                                     gn_stream_t stream) {
  GN_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && reps > 4 && (distinct_b == 1 || distinct_b == 2 || distinct_b == 4) && ctas >= 1 &&
                 out_dev, "gn_selftest_umma_rate: bad arguments");
  const int smem = 2 * 128 * 16 + distinct_b * 2 * n * 16;
  GN_REQUIRE(smem <= 200 * 1024, "gn_selftest_umma_rate: too many distinct B tiles");
  cudaError_t e = cudaFuncSetAttribute(gn::umma_rate_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_selftest_umma_rate: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  gn::umma_rate_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(
      n, reps, distinct_b, reinterpret_cast<long long*>(out_dev));
  GN_CHECK_LAUNCH("gn_selftest_umma_rate");
  return GN_OK;
}

// ---------------------------------------------------------------------------------
// Tensor-map TMA self-test: pins the conventions gn_block_tma.cu relies on.
//   1. tile load of 128 rows x 64 bf16 (SWIZZLE_128B) from mat[row0 ...]        -> A0
//   2. 32 x gather4 of the rows idx[0..127] of the same matrix                  -> A1
//   3. tile load of wmat[64 x 64 bf16]                                          -> B
//   4. raw dump of A0 | A1 | B (40 KB), then D[128 x 64] = (A0 + A1) . B^T with eight
//      SS-form UMMAs on SWIZZLE_128B K-major descriptors (K advance = 32 bytes)
// tests/test_gpu_umma.py checks the dump against the expected swizzle (16-byte chunk j of
// row r at r * 128 + ((j ^ (r % 8)) * 16)) and D against a float64 product.
// ---------------------------------------------------------------------------------
#include "gn_tma.cuh"

namespace gn {
constexpr uint32_t TS_A_BYTES = 128 * 128, TS_B_BYTES = 64 * 128;

__global__ void __launch_bounds__(128, 1)
tma_selftest_kernel(const __grid_constant__ CUtensorMap tm_tile,
                    const __grid_constant__ CUtensorMap tm_row,
                    const __grid_constant__ CUtensorMap tm_w, const int32_t* __restrict__ idx,
                    int row0, unsigned char* __restrict__ dump, float* __restrict__ d_out) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t sbase = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (sbase - umma::smem_u32(smem_raw));
  const uint32_t s_a0 = sbase, s_a1 = sbase + TS_A_BYTES, s_b = sbase + 2 * TS_A_BYTES;

  if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::fence_barrier_init();
  }
  __syncthreads();
  if (warp == 1) {
    const int4 r = __ldg(reinterpret_cast<const int4*>(idx) + lane);
    if (lane == 0) {
      umma::mbar_expect_tx(&bars[0], 2 * TS_A_BYTES + TS_B_BYTES);
      umma::tma_load_2d(s_a0, &tm_tile, 0, row0, &bars[0], umma::TMA_EVICT_FIRST);
      umma::tma_load_2d(s_b, &tm_w, 0, 0, &bars[0], umma::TMA_EVICT_LAST);
    }
    __syncwarp();
    umma::tma_gather4(s_a1 + lane * 512, &tm_row, 0, r.x, r.y, r.z, r.w, &bars[0],
                      umma::TMA_EVICT_LAST);
  }
  umma::mbar_wait(&bars[0], 0);
  for (int i = t; i < (int)(2 * TS_A_BYTES + TS_B_BYTES) / 16; i += 128)
    reinterpret_cast<uint4*>(dump)[i] = reinterpret_cast<const uint4*>(smem)[i];
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (t == 0) {
    const uint32_t idesc = umma::idesc_bf16_f32(128, 64);
    const uint64_t da0 = umma::smem_desc_sw128(s_a0), da1 = umma::smem_desc_sw128(s_a1);
    const uint64_t db = umma::smem_desc_sw128(s_b);
    for (int ks = 0; ks < 4; ++ks) {
      umma::mma_bf16_ss(tmem, da0 + ks * 2, db + ks * 2, idesc, ks > 0);
      umma::mma_bf16_ss(tmem, da1 + ks * 2, db + ks * 2, idesc, 1);
    }
    umma::mma_commit(&bars[1]);
  }
  umma::mbar_wait(&bars[1], 0);
  umma::tc_fence_after();
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) d_out[(size_t)t * 64 + c0 + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}
}  // namespace gn

extern "C" int gn_selftest_tma(const void* mat_bf16, int rows, const void* wmat_bf16,
                               const int32_t* idx, int row0, void* dump, float* d_out,
                               gn_stream_t stream) {
  GN_REQUIRE(mat_bf16 && wmat_bf16 && idx && dump && d_out, "gn_selftest_tma: null pointer");
  GN_REQUIRE(rows >= 1 && row0 >= 0, "gn_selftest_tma: bad sizes");
  GN_REQUIRE((((uintptr_t)mat_bf16 | (uintptr_t)wmat_bf16 | (uintptr_t)idx | (uintptr_t)dump) & 15) == 0,
             "gn_selftest_tma: unaligned pointer");
  CUtensorMap tm_tile, tm_row, tm_w;
  int r = gn::encode_tmap_2d_bf16(&tm_tile, mat_bf16, rows, 64, 128, 128, 64);
  if (r == 0) r = gn::encode_tmap_2d_bf16(&tm_row, mat_bf16, rows, 64, 128, 1, 64);
  if (r == 0) r = gn::encode_tmap_2d_bf16(&tm_w, wmat_bf16, 64, 64, 128, 64, 64);
  if (r != 0) {
    gn::set_error("gn_selftest_tma: cuTensorMapEncodeTiled failed (%d)", r);
    return GN_ERR_CUDA;
  }
  const int smem = 2 * gn::TS_A_BYTES + gn::TS_B_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(gn::tma_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_selftest_tma: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  gn::tma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
      tm_tile, tm_row, tm_w, idx, row0, static_cast<unsigned char*>(dump), d_out);
  GN_CHECK_LAUNCH("gn_selftest_tma");
  return GN_OK;
}

// ---------------------------------------------------------------------------------
// Store-bandwidth micro-benchmark (no reference counterpart): the ceiling the write-only
// dense IoU kernel (gn_iou.cu) is measured against.  Every mode writes `bytes` bytes of
// constants, persistent CTAs, each warp instruction covering contiguous memory:
//   0  st.global.v4.f32 (128-bit, default policy)     1  st.global.cs.v4.f32 (streaming)
//   2  st.global.v8.f32 (256-bit, sm_100)             3  st.global.wt.v4.f32 (write-through)
//   4  cp.async.bulk shared -> global, 16 KB per copy (TMA store path)
//   5  st.global.v8.f32 with an L2 evict-first policy
// ---------------------------------------------------------------------------------
namespace gn {
template <int MODE>
__global__ void __launch_bounds__(256)
store_bw_kernel(float* __restrict__ dst, size_t n16) {   // n16 = number of 16-byte units
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  if constexpr (MODE == 4) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = v;
    umma::fence_smem_to_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      const size_t chunks = n16 / 1024;   // 16 KB chunks
      int inflight = 0;
      for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        umma::bulk_copy_s2g(reinterpret_cast<unsigned char*>(dst) + c * 16384, umma::smem_u32(sm), 16384);
        umma::bulk_commit_group();
        if (++inflight >= 8) {
          asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
          inflight = 4;
        }
      }
      umma::bulk_wait_group0();
    }
  } else if constexpr (MODE == 2 || MODE == 5) {
    uint64_t pol = 0;
    if (MODE == 5) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t n32 = n16 / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32;
         i += (size_t)gridDim.x * blockDim.x) {
      float* p = dst + i * 8;
      if (MODE == 2)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};"
                     ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      else
        asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4}, %5;"
                     ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (size_t)gridDim.x * blockDim.x) {
      float4* p = reinterpret_cast<float4*>(dst) + i;
      if (MODE == 0) *p = v;
      else if (MODE == 1) __stcs(p, v);
      else __stwt(p, v);
    }
  }
}
}  // namespace gn

extern "C" int gn_selftest_store_bw(void* dst, int64_t bytes, int mode, int ctas_per_sm,
                                    gn_stream_t stream) {
  GN_REQUIRE(dst && bytes >= 16384 && bytes % 16384 == 0, "gn_selftest_store_bw: bytes must be a "
             "positive multiple of 16384");
  GN_REQUIRE(((uintptr_t)dst & 127) == 0, "gn_selftest_store_bw: dst must be 128-byte aligned");
  GN_REQUIRE(mode >= 0 && mode <= 5 && ctas_per_sm >= 1 && ctas_per_sm <= 32,
             "gn_selftest_store_bw: bad mode / ctas_per_sm");
  const int grid = gn::sm_count() * ctas_per_sm;
  const size_t n16 = (size_t)bytes / 16;
  float* d = static_cast<float*>(dst);
  cudaStream_t s = (cudaStream_t)stream;
  switch (mode) {
    case 0: gn::store_bw_kernel<0><<<grid, 256, 0, s>>>(d, n16); break;
    case 1: gn::store_bw_kernel<1><<<grid, 256, 0, s>>>(d, n16); break;
    case 2: gn::store_bw_kernel<2><<<grid, 256, 0, s>>>(d, n16); break;
    case 3: gn::store_bw_kernel<3><<<grid, 256, 0, s>>>(d, n16); break;
    case 4: gn::store_bw_kernel<4><<<grid, 256, 16384, s>>>(d, n16); break;
    default: gn::store_bw_kernel<5><<<grid, 256, 0, s>>>(d, n16); break;
  }
  GN_CHECK_LAUNCH("gn_selftest_store_bw");
  return GN_OK;
}
