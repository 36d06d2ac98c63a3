// Generic fully connected layer and its weight gradient on the tensor cores (tcgen05,
// bf16x3 = fp32 semantics, gn_umma.cuh) - the shape-agnostic counterparts of gn_fc.cu /
// gn_train.cu that the TRAINING step uses for every tf.contrib.layers.fully_connected of
// nms_net/network.py (:229-272, 328-341, 348-405) and its MatMul gradients:
//
//   gn_fc_fwd_tc         y = act(res + (x . mask) @ W + b)        rows x k  ->  rows x n
//                        (also dx = (dy . relu') @ W^T with the transposed operand image)
//   gn_fc_bwd_weight_tc  dW += x^T @ (dy . relu'),  db += colsum(dy . relu')
//
// `mask` (optional, same shape as the row operand it masks) fuses tf.nn.relu's gradient:
// the value is used where mask > 0 and 0 elsewhere, so no separate masking pass runs.
// Both kernels read fp32 activations, split them into bf16 hi / lo on the way into shared
// memory (K-major, no-swizzle "interleaved" layout of gn_umma.cuh) and accumulate in TMEM.
#include "gn_common.cuh"
#include "gn_umma.cuh"

namespace gn {

// ==================================================================================
// forward-type GEMM: tile = 128 rows, K in chunks of 64, N <= 256 in one accumulator
// ==================================================================================
constexpr int FT_TILE = 128, FT_THREADS = 256, FT_KC = 64;
constexpr uint32_t FT_SBO = 128;
constexpr uint32_t FT_LBO_A = FT_TILE * 16 + 32;          // skewed chunk pitch (bank spread)
constexpr uint32_t FT_A_HALF = (FT_KC / 8) * FT_LBO_A;    // hi (or lo) part of an A chunk
constexpr uint32_t FT_OFF_B = 2 * FT_A_HALF;
constexpr int FT_STG_PITCH = 64 + 4;                      // staging row pitch (floats)

__host__ __device__ constexpr uint32_t ft_b_half(int n) { return (FT_KC / 8) * (uint32_t)n * 16; }
__host__ __device__ constexpr uint32_t ft_smem(int n) {
  return FT_OFF_B + 2 * ft_b_half(n) + (uint32_t)n * 4 + 64;
}
static_assert(FT_TILE * FT_STG_PITCH * 4 <= FT_OFF_B + 2 * 8 * 32 * 16, "staging tile must fit");

// units of 8 consecutive k of one row, fetched into registers one chunk ahead
struct FtPrefetch {
  float4 v[4][2];
};

template <bool VEC>
__device__ __forceinline__ void ft_fetch(FtPrefetch& pf, const float* __restrict__ x,
                                         const float* __restrict__ mask, int ldx, int row0,
                                         int rows, int k, int kc, int t) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = i * FT_THREADS + t;
    const int r = u >> 3, col = kc * FT_KC + (u & 7) * 8;
    const int gr = row0 + r;
    float vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) vals[e] = 0.f;
    if (gr < rows && col < k) {
      const float* p = x + (size_t)gr * ldx + col;
      if (VEC) {     // ldx % 4 == 0, k % 8 == 0, 16-byte aligned base
        const float4 a = ldg4(p), b = ldg4(p + 4);
        vals[0] = a.x; vals[1] = a.y; vals[2] = a.z; vals[3] = a.w;
        vals[4] = b.x; vals[5] = b.y; vals[6] = b.z; vals[7] = b.w;
        if (mask != nullptr) {
          const float* q = mask + (size_t)gr * ldx + col;
          const float4 ma = ldg4(q), mb = ldg4(q + 4);
          const float m[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) vals[e] = m[e] > 0.f ? vals[e] : 0.f;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (col + e < k) {
            float a = __ldg(p + e);
            if (mask != nullptr && !(__ldg(mask + (size_t)gr * ldx + col + e) > 0.f)) a = 0.f;
            vals[e] = a;
          }
      }
    }
    pf.v[i][0] = make_float4(vals[0], vals[1], vals[2], vals[3]);
    pf.v[i][1] = make_float4(vals[4], vals[5], vals[6], vals[7]);
  }
}

__device__ __forceinline__ void ft_store(const FtPrefetch& pf, unsigned char* a_hi, int t) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = i * FT_THREADS + t;
    const int r = u >> 3, piece = u & 7;
    uint4 h, l;
    umma::split_bf16x2(pf.v[i][0].x, pf.v[i][0].y, h.x, l.x);
    umma::split_bf16x2(pf.v[i][0].z, pf.v[i][0].w, h.y, l.y);
    umma::split_bf16x2(pf.v[i][1].x, pf.v[i][1].y, h.z, l.z);
    umma::split_bf16x2(pf.v[i][1].z, pf.v[i][1].w, h.w, l.w);
    const uint32_t off = (uint32_t)piece * FT_LBO_A + (uint32_t)r * 16;
    *reinterpret_cast<uint4*>(a_hi + off) = h;
    *reinterpret_cast<uint4*>(a_hi + FT_A_HALF + off) = l;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(FT_THREADS, 2)
fc_tc_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mask,
             const unsigned char* __restrict__ wimg, const float* __restrict__ bias,
             const float* __restrict__ res, int ld_res, int relu, float* __restrict__ y, int ldy,
             int rows_host, const int32_t* __restrict__ rows_dev, int k, int kpad, int n,
             uint32_t tmem_cols) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  int rows = rows_host;
  if (rows_dev != nullptr) rows = min(rows, __ldg(rows_dev));
  const int num_tiles = (rows + FT_TILE - 1) / FT_TILE;
  if ((int)blockIdx.x >= num_tiles) return;

  const uint32_t bhalf = ft_b_half(n);
  unsigned char* a_hi = smem;
  float* bias_s = reinterpret_cast<float*>(smem + FT_OFF_B + 2 * bhalf);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + FT_OFF_B + 2 * bhalf + (uint32_t)n * 4);
  uint64_t* wbar = bar + 1;
  float* stage = reinterpret_cast<float*>(smem);       // aliases A / B after the last UMMA

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, tmem_cols);
  if (t == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_init(wbar, 1);
    umma::fence_barrier_init();
  }
  for (int c = t; c < n; c += FT_THREADS) bias_s[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sa = umma::smem_u32(smem), sb = sa + FT_OFF_B;
  const uint64_t d_ah = umma::smem_desc(sa, FT_LBO_A, FT_SBO);
  const uint64_t d_al = umma::smem_desc(sa + FT_A_HALF, FT_LBO_A, FT_SBO);
  const uint64_t d_bh = umma::smem_desc(sb, (uint32_t)n * 16, FT_SBO);
  const uint64_t d_bl = umma::smem_desc(sb + bhalf, (uint32_t)n * 16, FT_SBO);
  const uint32_t idesc = umma::idesc_bf16_f32(FT_TILE, n);
  const int kchunks = (kpad + FT_KC - 1) / FT_KC;
  const uint32_t img_half = (uint32_t)(kpad / 8) * (uint32_t)n * 16;   // hi image bytes
  const int quad = warp & 3, half = warp >> 2;
  const uint32_t tlane = (uint32_t)(quad * 32) << 16;
  uint32_t par = 0, wpar = 0;

  FtPrefetch pf;
  ft_fetch<VEC>(pf, x, mask, ldx, blockIdx.x * FT_TILE, rows, k, 0, t);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row0 = tile * FT_TILE;
    for (int kc = 0; kc < kchunks; ++kc) {
      // operands of chunk kc: A from the prefetch registers, B by bulk copy from the image
      const int ksteps = min(FT_KC, kpad - kc * FT_KC) / 16;
      const uint32_t bbytes = (uint32_t)ksteps * 2 * (uint32_t)n * 16;
      if (t == 0) {
        umma::mbar_expect_tx(wbar, 2 * bbytes);
        const unsigned char* src = wimg + (size_t)kc * (FT_KC / 8) * n * 16;
        umma::bulk_copy_g2s(sb, src, bbytes, wbar);
        umma::bulk_copy_g2s(sb + bhalf, src + img_half, bbytes, wbar);
      }
      ft_store(pf, a_hi, t);
      // next chunk (or the next tile's first chunk) into registers while this one computes
      {
        int ntile = tile, nkc = kc + 1;
        if (nkc == kchunks) { nkc = 0; ntile = tile + gridDim.x; }
        if (ntile < num_tiles) ft_fetch<VEC>(pf, x, mask, ldx, ntile * FT_TILE, rows, k, nkc, t);
      }
      umma::fence_smem_to_async();
      umma::tc_fence_before();
      __syncthreads();
      if (t < 32) {      // warp 0, converged: elect.sync picks the issuing lane (gn_umma.cuh)
        umma::mbar_wait(wbar, wpar);
        umma::tc_fence_after();
        for (int ks = 0; ks < ksteps; ++ks)
          umma::mma_bf16x3_elect(tmem, d_ah, d_al, d_bh, d_bl, ks * (2 * FT_LBO_A >> 4),
                                 ks * (2 * (uint32_t)n * 16 >> 4), idesc, (kc | ks) != 0);
        umma::mma_commit_elect(bar);
      }
      wpar ^= 1;
      umma::mbar_wait(bar, par);        // the UMMAs have consumed the chunk: buffers reusable
      par ^= 1;
      umma::tc_fence_after();
    }
    // ---- epilogue: 64 (or 32) columns per pass through the staging tile -----------------
    for (int c0 = 0; c0 < n; c0 += 64) {
      const int width = min(64, n - c0);
      if (half * 32 < width) {
        float v[32];
        umma::tmem_ld32(tmem + tlane + (uint32_t)(c0 + half * 32), v);
        umma::tmem_ld_wait();
        float* dst = stage + (quad * 32 + lane) * FT_STG_PITCH + half * 32;
        const float* bb = bias_s + c0 + half * 32;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(dst + g * 4) =
              make_float4(v[g * 4] + bb[g * 4], v[g * 4 + 1] + bb[g * 4 + 1],
                          v[g * 4 + 2] + bb[g * 4 + 2], v[g * 4 + 3] + bb[g * 4 + 3]);
      }
      __syncthreads();
      const int per_row = width / 4;                    // float4 per row: 16 or 8
      for (int idx = t; idx < FT_TILE * per_row; idx += FT_THREADS) {
        const int r = idx / per_row, c4 = (idx - r * per_row) * 4;
        const int gr = row0 + r;
        if (gr >= rows) continue;
        float4 o = *reinterpret_cast<const float4*>(stage + r * FT_STG_PITCH + c4);
        if (res != nullptr) {
          const float4 rv = ldg4(res + (size_t)gr * ld_res + c0 + c4);
          o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(y + (size_t)gr * ldy + c0 + c4) = o;
      }
      __syncthreads();
    }
    umma::tc_fence_before();
    __syncthreads();       // staging (= operand buffers) and the accumulator are reused
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, tmem_cols);
}

// ==================================================================================
// weight gradient: dW[k, n] += x^T dy.  M = k (tiles of 128, zero padded), N = n, the
// reduction runs over the rows, R = 32 (or 16) rows per stage.
//
// The operands stream from HBM exactly once, so the kernel lives on memory-level
// parallelism.  (A first version loaded through registers: 8 KB in flight per SM, 0.4 TB/s,
// 435 us per call whatever the shape.)  Row blocks of x, dy and the relu mask are contiguous
// in memory, so ONE thread moves them with cp.async.bulk into a ring of raw fp32 stages
// (up to 4 in flight); eight converter warps read a raw stage, apply the mask, split to bf16
// hi / lo and write both operands TRANSPOSED (an 8-row group of one column = one 16-byte
// K-major chunk) into a two-stage operand ring; one thread issues the UMMAs.
// Measured (experiments/wgrad_time.py, P = 311 072 rows): 101 us for 64 x 64 (2.4 TB/s of
// operand traffic), 243 us for 256 x 256 (3.9 TB/s).  The converters set the pace: one pass
// over a warp's units costs 900-1 900 cycles almost independently of the unit width, so
// sixteen converter warps with at most one unit of x and one of dy per thread and stage are
// used (8 warps, 4-column units: 136 us; dealing units out as quarter-warps so that every warp
// ran both passes: 172 us).
//   warps 0-15 converters (0-7 also the epilogue)   warp 16 UMMA issuer   warp 17 bulk-copy producer
// ==================================================================================
constexpr int WT_CONV_WARPS = 16, WT_CONV = WT_CONV_WARPS * 32;
constexpr int WT_THREADS = (WT_CONV_WARPS + 2) * 32;
constexpr int WT_MAX_RAW = 4, WT_OPS = 2;

struct WtPlan {
  int rows_per_stage, raw_stages, bulk;
  uint32_t raw_stage_bytes, op_stage_bytes, smem;
};

static WtPlan wt_plan(int k, int n, bool has_mask, bool contiguous) {
  WtPlan p;
  const int mrows = (k + 127) / 128 * 128;
  p.bulk = contiguous ? 1 : 0;
  for (int R = 32; R >= 16; R -= 16) {
    const uint32_t raw = ((uint32_t)R * 4u * (uint32_t)(k + n * (has_mask ? 2 : 1)) + 127u) & ~127u;
    const uint32_t op = 2u * (uint32_t)(R / 8) * (uint32_t)(mrows + n + 2) * 16u;
    for (int rs = WT_MAX_RAW; rs >= 2; --rs) {
      const uint32_t total = (contiguous ? rs * raw : 0) + WT_OPS * op;
      if (total <= 200u * 1024u) {
        p.rows_per_stage = R;
        p.raw_stages = rs;
        p.raw_stage_bytes = raw;
        p.op_stage_bytes = op;
        p.smem = total < 120u * 1024u ? 120u * 1024u : total;
        return p;
      }
    }
  }
  p.rows_per_stage = 0;
  return p;
}

// One converter unit: 8 rows (row group g of the stage) x W consecutive columns of an operand
// tile -> W K-major chunk pairs (hi, lo).  Source: the raw fp32 stage in shared memory, or
// global memory for chunks that did not go through the bulk-copy ring.  `msk`: relu mask of
// the same shape (value kept where mask > 0); `bsum`: running column sums of what was kept.
template <int W>
__device__ __forceinline__ void wt_convert_unit(const float* __restrict__ src, const float* __restrict__ msk,
                                                size_t pitch, int col, int row_first, int rows_left,
                                                unsigned char* dst, uint32_t half_bytes,
                                                float* bsum) {
  float v[8][W];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int c = 0; c < W; ++c) v[e][c] = 0.f;
    if (e < rows_left) {
      const float* p = src + (size_t)(row_first + e) * pitch + col;
      float m[W];
      if constexpr (W == 4) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[e][0] = a.x; v[e][1] = a.y; v[e][2] = a.z; v[e][3] = a.w;
        if (msk != nullptr) {
          const float4 b = *reinterpret_cast<const float4*>(msk + (size_t)(row_first + e) * pitch + col);
          m[0] = b.x; m[1] = b.y; m[2] = b.z; m[3] = b.w;
        }
      } else if constexpr (W == 2) {
        const float2 a = *reinterpret_cast<const float2*>(p);
        v[e][0] = a.x; v[e][1] = a.y;
        if (msk != nullptr) {
          const float2 b = *reinterpret_cast<const float2*>(msk + (size_t)(row_first + e) * pitch + col);
          m[0] = b.x; m[1] = b.y;
        }
      } else {
        v[e][0] = *p;
        if (msk != nullptr) m[0] = msk[(size_t)(row_first + e) * pitch + col];
      }
      if (msk != nullptr) {
#pragma unroll
        for (int c = 0; c < W; ++c) v[e][c] = m[c] > 0.f ? v[e][c] : 0.f;
      }
      if (bsum != nullptr) {
#pragma unroll
        for (int c = 0; c < W; ++c) bsum[c] += v[e][c];
      }
    }
  }
#pragma unroll
  for (int c = 0; c < W; ++c) {
    uint4 h, l;
    umma::split_bf16x2(v[0][c], v[1][c], h.x, l.x);
    umma::split_bf16x2(v[2][c], v[3][c], h.y, l.y);
    umma::split_bf16x2(v[4][c], v[5][c], h.z, l.z);
    umma::split_bf16x2(v[6][c], v[7][c], h.w, l.w);
    *reinterpret_cast<uint4*>(dst + c * 16) = h;
    *reinterpret_cast<uint4*>(dst + half_bytes + c * 16) = l;
  }
}

__global__ void __launch_bounds__(WT_THREADS, 1)
fc_wgrad_tc_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int ldy,
                   const float* __restrict__ mask, float* __restrict__ dw, float* __restrict__ db,
                   int rows_host, const int32_t* __restrict__ rows_dev, int k, int n,
                   uint32_t tmem_cols, int R, int raw_stages, uint32_t raw_stage_bytes,
                   uint32_t op_stage_bytes, int bulk) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bars[2 * WT_MAX_RAW + 2 * WT_OPS + 1];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  int rows = rows_host;
  if (rows_dev != nullptr) rows = min(rows, __ldg(rows_dev));
  const int nchunks = (rows + R - 1) / R;
  if ((int)blockIdx.x >= nchunks) return;
  const int my_chunks = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int mt_count = (k + 127) / 128, mrows = mt_count * 128;
  const int G = R / 8, gshift = (G == 4) ? 2 : 1;           // 8-row groups per stage
  // chunk pitch skewed by one 16-byte unit: lanes that differ in the row group then hit
  // different bank groups (see the converter's unit mapping)
  const uint32_t lbo_a = (uint32_t)mrows * 16 + 16, lbo_b = (uint32_t)n * 16 + 16;
  const uint32_t a_half = (uint32_t)G * lbo_a, b_half = (uint32_t)G * lbo_b;
  unsigned char* raw_base = smem;
  unsigned char* op_base = smem + (bulk ? (uint32_t)raw_stages * raw_stage_bytes : 0u);
  uint64_t* raw_full = bars;                         // [raw_stages] bulk copies landed
  uint64_t* raw_empty = bars + WT_MAX_RAW;           // [raw_stages] every converter thread
  uint64_t* op_full = bars + 2 * WT_MAX_RAW;         // [2] the converter warps
  uint64_t* op_empty = op_full + WT_OPS;             // [2] tcgen05.commit
  uint64_t* done = op_empty + WT_OPS;
  const uint32_t x_bytes = (uint32_t)R * (uint32_t)k * 4u, d_bytes = (uint32_t)R * (uint32_t)n * 4u;
  // a chunk goes through the bulk-copy ring when its byte counts are 16-byte multiples (always
  // for full chunks; a ragged last chunk of 9-float rows is read from global memory directly)
  auto chunk_is_bulk = [&](int r0) {
    const int live = min(R, rows - r0);
    return bulk && ((live * k) % 4 == 0);
  };

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, tmem_cols);
  if (t == 0) {
    for (int s = 0; s < raw_stages; ++s) {
      umma::mbar_init(&raw_full[s], 1);
      umma::mbar_init(&raw_empty[s], WT_CONV);          // every converter thread releases its reads
    }
    for (int s = 0; s < WT_OPS; ++s) {
      umma::mbar_init(&op_full[s], WT_CONV_WARPS);
      umma::mbar_init(&op_empty[s], 1);
    }
    umma::mbar_init(done, 1);
    umma::fence_barrier_init();
  }
  // the zero padding of the M dimension (rows k .. mrows of A') is written once
  for (int s = 0; s < WT_OPS; ++s)
    for (int u = t; u < 2 * G * (mrows - k); u += WT_THREADS) {
      const int part = u / (G * (mrows - k)), v = u % (G * (mrows - k));
      const int g = v / (mrows - k), m = k + v % (mrows - k);
      *reinterpret_cast<uint4*>(op_base + s * op_stage_bytes + part * a_half + g * lbo_a + m * 16) =
          make_uint4(0, 0, 0, 0);
    }
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == WT_CONV_WARPS + 1) {
    // ================================ producer ========================================
    if (lane == 0 && bulk) {
      int s = 0;
      uint32_t nuse = 0;
#pragma unroll 1
      for (int it = 0; it < my_chunks; ++it) {
        const int r0 = (blockIdx.x + it * gridDim.x) * R;
        if (nuse >= 1) umma::mbar_wait_relaxed(&raw_empty[s], (nuse - 1) & 1u);
        if (chunk_is_bulk(r0)) {
          const int live = min(R, rows - r0);
          const uint32_t xb = (uint32_t)live * k * 4u, db_ = (uint32_t)live * n * 4u;
          const uint32_t dst = umma::smem_u32(raw_base + s * raw_stage_bytes);
          umma::mbar_expect_tx(&raw_full[s], xb + db_ * (mask != nullptr ? 2u : 1u));
          umma::bulk_copy_g2s(dst, x + (size_t)r0 * k, xb, &raw_full[s]);
          umma::bulk_copy_g2s(dst + x_bytes, dy + (size_t)r0 * n, db_, &raw_full[s]);
          if (mask != nullptr)
            umma::bulk_copy_g2s(dst + x_bytes + d_bytes, mask + (size_t)r0 * n, db_, &raw_full[s]);
        } else {
          umma::mbar_arrive(&raw_full[s]);          // the converters read this chunk themselves
        }
        if (++s == raw_stages) { s = 0; ++nuse; }
      }
    }
  } else if (warp == WT_CONV_WARPS) {
    // ================================ MMA issuer =======================================
    {   // whole warp, warp-uniform operands; elect.sync picks the issuing lane (gn_umma.cuh)
      const uint32_t idesc = umma::idesc_bf16_f32(128, n);
      const uint32_t sbase = umma::smem_u32(op_base);
#pragma unroll 1
      for (int it = 0; it < my_chunks; ++it) {
        const int o = it & 1;
        // relaxed: a hot spin here would share its scheduler with two converter warps
        umma::mbar_wait_relaxed(&op_full[o], ((uint32_t)it >> 1) & 1u);
        umma::tc_fence_after();
        const uint32_t sa = sbase + o * op_stage_bytes, sb = sa + 2 * a_half;
        const uint64_t d_bh = umma::smem_desc(sb, lbo_b, 128), d_bl = umma::smem_desc(sb + b_half, lbo_b, 128);
        for (int mt = 0; mt < mt_count; ++mt) {
          const uint64_t d_ah = umma::smem_desc(sa + mt * 128 * 16, lbo_a, 128);
          const uint64_t d_al = umma::smem_desc(sa + a_half + mt * 128 * 16, lbo_a, 128);
          for (int ks = 0; ks < R / 16; ++ks)
            umma::mma_bf16x3_elect(tmem + (uint32_t)(mt * n), d_ah, d_al, d_bh, d_bl, ks * (2 * lbo_a >> 4),
                                   ks * (2 * lbo_b >> 4), idesc, (it | ks) != 0);
        }
        umma::mma_commit_elect(&op_empty[o]);
      }
      umma::mma_commit_elect(done);
    }
  } else {
    // ================================ converters ======================================
    // unit = (8-row group g, W consecutive columns), W = 2 for operands up to 128 columns and 4
    // above: every thread converts at most ONE unit of x and ONE of dy per stage, the x units
    // on the first threads, the dy units on the threads after them (small shapes keep eight
    // warps busy with one short pass each).  Lanes run over the row groups first, so with the
    // skewed chunk pitch a quarter-warp's 16-byte stores spread over the bank groups.
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};   // this thread's columns of dy (loop invariant)
    const bool vec_k = (k % 4 == 0);
    const int wa = k <= 64 ? 1 : k <= 128 ? 2 : 4, wb = n <= 64 ? 1 : n <= 128 ? 2 : 4;
    const int ua = vec_k ? G * (k / wa) : 0, ub = G * (n / wb);        // units per stage, <= 256
    const int toff = ua + ub <= WT_CONV ? ua : WT_CONV - ub;
    const int ubi = (t - toff) & (WT_CONV - 1);                         // this thread's dy unit
    int s = 0;
    uint32_t nuse = 0;
#pragma unroll 1
    for (int it = 0; it < my_chunks; ++it) {
      const int o = it & 1;
      const uint32_t ouse = (uint32_t)it >> 1;
      const int r0 = (blockIdx.x + it * gridDim.x) * R;
      const bool from_smem = chunk_is_bulk(r0);
      if (bulk) umma::mbar_wait_relaxed(&raw_full[s], nuse & 1u);
      if (ouse >= 1) umma::mbar_wait_relaxed(&op_empty[o], (ouse - 1) & 1u);
      const float* xs = reinterpret_cast<const float*>(raw_base + s * raw_stage_bytes);
      const float* ds = xs + R * k;
      const float* ms = ds + R * n;
      unsigned char* st = op_base + o * op_stage_bytes;
      unsigned char* sb_ = st + 2 * a_half;
      // ---- A' = x^T -----------------------------------------------------------------------
      if (vec_k) {
        if (t < ua) {
          const int g = t & (G - 1), kk = (t >> gshift) * wa;
          const float* src = from_smem ? xs : x + (size_t)r0 * ldx;
          const size_t pitch = from_smem ? (size_t)k : (size_t)ldx;
          unsigned char* dst = st + g * lbo_a + kk * 16;
          if (wa == 1) wt_convert_unit<1>(src, nullptr, pitch, kk, g * 8, rows - r0 - g * 8, dst, a_half, nullptr);
          else if (wa == 2) wt_convert_unit<2>(src, nullptr, pitch, kk, g * 8, rows - r0 - g * 8, dst, a_half, nullptr);
          else wt_convert_unit<4>(src, nullptr, pitch, kk, g * 8, rows - r0 - g * 8, dst, a_half, nullptr);
        }
      } else {
        for (int u = t; u < G * k; u += WT_CONV) {      // k = 9 raw pair features
          const int g = u / k, kk = u - g * k;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int r = g * 8 + e;
            v[e] = 0.f;
            if (r0 + r < rows) v[e] = from_smem ? xs[r * k + kk] : __ldg(x + (size_t)(r0 + r) * ldx + kk);
          }
          uint4 h, l;
          umma::split_bf16x2(v[0], v[1], h.x, l.x);
          umma::split_bf16x2(v[2], v[3], h.y, l.y);
          umma::split_bf16x2(v[4], v[5], h.z, l.z);
          umma::split_bf16x2(v[6], v[7], h.w, l.w);
          *reinterpret_cast<uint4*>(st + g * lbo_a + kk * 16) = h;
          *reinterpret_cast<uint4*>(st + a_half + g * lbo_a + kk * 16) = l;
        }
      }
      // ---- B' = (dy . relu')^T ----------------------------------------------------------------
      if (ubi < ub) {
        const int g = ubi & (G - 1), nn = (ubi >> gshift) * wb;
        const float* src = from_smem ? ds : dy + (size_t)r0 * ldy;
        const float* msk = mask == nullptr ? nullptr : (from_smem ? ms : mask + (size_t)r0 * ldy);
        const size_t pitch = from_smem ? (size_t)n : (size_t)ldy;
        unsigned char* dst = sb_ + g * lbo_b + nn * 16;
        if (wb == 1) wt_convert_unit<1>(src, msk, pitch, nn, g * 8, rows - r0 - g * 8, dst, b_half, bsum);
        else if (wb == 2) wt_convert_unit<2>(src, msk, pitch, nn, g * 8, rows - r0 - g * 8, dst, b_half, bsum);
        else wt_convert_unit<4>(src, msk, pitch, nn, g * 8, rows - r0 - g * 8, dst, b_half, bsum);
      }
      umma::fence_smem_to_async();
      if (bulk) umma::mbar_arrive(&raw_empty[s]);     // this thread's reads of the raw stage are done
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&op_full[o]);
      if (++s == raw_stages) { s = 0; ++nuse; }
    }
    if (db != nullptr && ubi < ub) {
      const int nn = (ubi >> gshift) * wb;
      for (int c = 0; c < wb; ++c)
        if (bsum[c] != 0.f) atomicAdd(db + nn + c, bsum[c]);
    }
    // ---- epilogue: accumulator rows (= weight rows) into dW ------------------------------
    umma::mbar_wait_relaxed(done, 0);
    umma::tc_fence_after();
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    for (int mt = 0; mt < mt_count && warp < 8; ++mt) {      // warps 0-7: two per lane quadrant
      const int kk = mt * 128 + quad * 32 + lane;
      for (int c0 = half * 32; c0 < n; c0 += 64) {       // the two warps of a quadrant alternate
        float v[32];
        umma::tmem_ld32(tmem + tlane + (uint32_t)(mt * n + c0), v);
        umma::tmem_ld_wait();
        if (kk < k) {
          float* dst = dw + (size_t)kk * n + c0;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"l"(dst + g * 4), "f"(v[g * 4]), "f"(v[g * 4 + 1]), "f"(v[g * 4 + 2]),
                           "f"(v[g * 4 + 3]) : "memory");
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, tmem_cols);
}

// ---------------------------------------------------------------------------------
// Operand images for gn_fc_fwd_tc.  table: 6 int32 per entry
//   (src offset in floats, k, n, dst offset in bytes, kpad, transposed)
// transposed = 0: B[nn][kk] = W[kk][nn]   (forward: y = x @ W,  W is [k, n], image N = n)
// transposed = 1: B[kk][nn] = W[kk][nn]   (input gradient: dx = dy @ W^T, image N = k, K = n;
//                                          kpad then pads n)
// Image layout: hi part then lo part, each (Kpad / 8) chunks x N rows x 16 bytes.
// ---------------------------------------------------------------------------------
__global__ void prepare_fc_images_kernel(const float* __restrict__ flat,
                                         const int32_t* __restrict__ table,
                                         unsigned char* __restrict__ image) {
  const int32_t* e = table + blockIdx.y * 6;
  const float* w = flat + e[0];
  const int k = e[1], n = e[2], kpad = e[4], tr = e[5];
  unsigned char* hi = image + e[3];
  const int N = tr ? k : n, K = tr ? n : k;             // image rows, reduction length
  unsigned char* lo = hi + (size_t)(kpad / 8) * N * 16;
  const int units = (kpad / 8) * N;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int row = u % N, j = u / N;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kk = j * 8 + i;
      x[i] = kk < K ? (tr ? __ldg(w + (size_t)row * n + kk) : __ldg(w + (size_t)kk * n + row)) : 0.f;
    }
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + (size_t)j * N * 16 + row * 16) = h;
    *reinterpret_cast<uint4*>(lo + (size_t)j * N * 16 + row * 16) = l;
  }
}

static uint32_t pow2_cols(int c) {
  uint32_t v = 32;
  while ((int)v < c) v <<= 1;
  return v;
}

}  // namespace gn

extern "C" int gn_prepare_fc_images(const float* flat_params, const int32_t* table, int entries,
                                    void* image, gn_stream_t stream) {
  GN_REQUIRE(entries >= 0, "gn_prepare_fc_images: negative entry count");
  if (entries == 0) return GN_OK;
  GN_REQUIRE(flat_params && table && image, "gn_prepare_fc_images: null pointer");
  GN_REQUIRE(((uintptr_t)image & 15) == 0, "gn_prepare_fc_images: image must be 16-byte aligned");
  gn::prepare_fc_images_kernel<<<dim3(8, entries), 256, 0, (cudaStream_t)stream>>>(
      flat_params, table, static_cast<unsigned char*>(image));
  GN_CHECK_LAUNCH("gn_prepare_fc_images");
  return GN_OK;
}

extern "C" int gn_fc_fwd_tc(const float* x, int ldx, const float* mask, const void* wimg,
                            const float* bias, const float* residual, int ld_res, int relu,
                            float* y, int ldy, int rows, const int32_t* rows_dev, int k, int kpad,
                            int n, gn_stream_t stream) {
  GN_REQUIRE(rows >= 0 && k > 0 && n > 0, "gn_fc_fwd_tc: bad shape rows=%d k=%d n=%d", rows, k, n);
  if (kpad % 16 != 0 || kpad < k || kpad > 256 || n % 32 != 0 || n > 256) {
    gn::set_error("gn_fc_fwd_tc: needs kpad %% 16 == 0, k <= kpad <= 256, n %% 32 == 0, n <= 256 "
                  "(got k=%d kpad=%d n=%d)", k, kpad, n);
    return GN_ERR_UNSUPPORTED;
  }
  GN_REQUIRE(ldx >= k && ldy >= n && ldy % 4 == 0, "gn_fc_fwd_tc: leading dimensions");
  GN_REQUIRE(residual == nullptr || (ld_res >= n && ld_res % 4 == 0), "gn_fc_fwd_tc: residual ld");
  if (rows == 0) return GN_OK;
  GN_REQUIRE(x && wimg && y, "gn_fc_fwd_tc: null pointer");
  GN_REQUIRE((((uintptr_t)wimg | (uintptr_t)y | (uintptr_t)residual) & 15) == 0,
             "gn_fc_fwd_tc: image, y and residual must be 16-byte aligned");
  const bool vec = ldx % 4 == 0 && k % 8 == 0 && (((uintptr_t)x | (uintptr_t)mask) & 15) == 0;
  const uint32_t smem = gn::ft_smem(n);
  const uint32_t cols = gn::pow2_cols(n);
  const void* kern = vec ? (const void*)gn::fc_tc_kernel<true> : (const void*)gn::fc_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_fc_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(rows, gn::FT_TILE);
  // two CTAs per SM share the 512 TMEM columns: that holds for every n <= 256
  const int cap = 2 * gn::sm_count();
  if (grid > cap) grid = cap;
  if (vec)
    gn::fc_tc_kernel<true><<<grid, gn::FT_THREADS, smem, (cudaStream_t)stream>>>(
        x, ldx, mask, static_cast<const unsigned char*>(wimg), bias, residual, ld_res, relu, y, ldy,
        rows, rows_dev, k, kpad, n, cols);
  else
    gn::fc_tc_kernel<false><<<grid, gn::FT_THREADS, smem, (cudaStream_t)stream>>>(
        x, ldx, mask, static_cast<const unsigned char*>(wimg), bias, residual, ld_res, relu, y, ldy,
        rows, rows_dev, k, kpad, n, cols);
  GN_CHECK_LAUNCH("gn_fc_fwd_tc");
  return GN_OK;
}

extern "C" int gn_fc_bwd_weight_tc(const float* x, int ldx, const float* dy, int ldy,
                                   const float* mask, float* dw, float* db, int rows,
                                   const int32_t* rows_dev, int k, int n, gn_stream_t stream) {
  GN_REQUIRE(rows >= 0 && k > 0 && n > 0, "gn_fc_bwd_weight_tc: bad shape");
  if (k > 256 || n > 256 || n % 32 != 0 || 256 % n != 0) {
    gn::set_error("gn_fc_bwd_weight_tc: needs k <= 256 and n in {32, 64, 128, 256} (got k=%d n=%d)",
                  k, n);
    return GN_ERR_UNSUPPORTED;
  }
  GN_REQUIRE(ldx >= k && ldy >= n && ldy % 4 == 0, "gn_fc_bwd_weight_tc: leading dimensions");
  GN_REQUIRE(k % 4 != 0 || ldx % 4 == 0, "gn_fc_bwd_weight_tc: ldx must be a multiple of 4 when k is");
  if (rows == 0) return GN_OK;
  GN_REQUIRE(x && dy && dw, "gn_fc_bwd_weight_tc: null pointer");
  GN_REQUIRE((((uintptr_t)dw | (uintptr_t)x | (uintptr_t)dy | (uintptr_t)mask) & 15) == 0,
             "gn_fc_bwd_weight_tc: pointers must be 16-byte aligned");
  // contiguous row blocks (ld == width) go through the bulk-copy ring
  const bool contiguous = ldx == k && ldy == n;
  const gn::WtPlan p = gn::wt_plan(k, n, mask != nullptr, contiguous);
  GN_REQUIRE(p.rows_per_stage > 0, "gn_fc_bwd_weight_tc: no shared-memory plan for k=%d n=%d", k, n);
  const int mrows = gn::ceil_div(k, 128) * 128;
  const uint32_t cols = gn::pow2_cols((mrows / 128) * n);
  cudaError_t e = cudaFuncSetAttribute(gn::fc_wgrad_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_fc_bwd_weight_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(rows, p.rows_per_stage);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  // at least 120 KB of dynamic shared memory: ONE CTA per SM, so the TMEM allocation
  // (up to all 512 columns) can never wait for a co-resident CTA
  gn::fc_wgrad_tc_kernel<<<grid, gn::WT_THREADS, p.smem, (cudaStream_t)stream>>>(
      x, ldx, dy, ldy, mask, dw, db, rows, rows_dev, k, n, cols, p.rows_per_stage, p.raw_stages,
      p.raw_stage_bytes, p.op_stage_bytes, p.bulk);
  GN_CHECK_LAUNCH("gn_fc_bwd_weight_tc");
  return GN_OK;
}
