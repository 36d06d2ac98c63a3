// One Gnet block's pair stage on the tensor cores (A7a + pw_fc1 + pw_fc2 + A7b),
// shipped shape w = 32, r = 32, f = 64:
//
//   x[p]   = [pw_feats[p] | feats[pair_c[p]] | nfeats[pair_n[p]] (0 if c == n)]   (96)
//   h1     = relu(x  @ W1 + b1)                                                    (64)
//   h2     = relu(h1 @ W2 + b2)                                                    (64)
//   pooled[c] = max over the pairs of c of h2                     (network.py:367-388)
//
// Tile = 128 consecutive pairs = the M dimension of one tcgen05.mma.  Per tile:
//   1. fill   : 256 threads gather the 96 fp32 inputs of each pair (coalesced
//               8-row x 128-byte warp requests), split every value into bf16
//               hi + lo and store both into K-major operand tiles in shared memory;
//   2. FC1    : one thread issues 6 k-steps x 3 (lo*hi, hi*lo, hi*hi) UMMAs,
//               M=128 N=64 K=16, fp32 accumulation in TMEM columns [0,64);
//   3. epi 1  : 8 warps read the accumulator (tcgen05.ld), add b1, ReLU, split to
//               bf16 hi/lo and store h1 as the A operand of FC2;
//   4. FC2    : 4 k-steps x 3 UMMAs into TMEM columns [64,128);
//   5. epi 2  : + b2, ReLU, fp32 tile to shared memory;
//   6. pool   : per column, runs of equal c are max-reduced and merged into
//               pooled[] with one integer atomicMax per (segment, tile, column):
//               activations are >= 0 and every detection owns its self pair, so
//               the result is exact and order independent.
// W1^T / W2^T (hi and lo) stay resident in shared memory for the CTA's lifetime.
// The phases of one CTA are sequential; two CTAs are resident per SM (92 KB smem,
// 128 TMEM columns each) so one CTA's UMMAs overlap the other's CUDA-core phases.
#include <cstdlib>
#include "gn_common.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int TC_TILE = 128;
constexpr int TC_THREADS = 256;
constexpr int TC_W = 32, TC_R = 32, TC_F = 64;
constexpr int TC_K1 = TC_W + 2 * TC_R;          // 96
constexpr int TC_CH1 = TC_K1 / 8;               // 12 chunks of 8 bf16
constexpr int TC_CH2 = TC_F / 8;                // 8
constexpr uint32_t TC_SBO = 128;                // 8 rows x 16 B
// A-side chunk pitch: 128 rows x 16 B + 32 B skew, so that the fill's
// (2 rows x 4 chunks) quarter-warp store hits 8 distinct 16-byte bank groups
constexpr uint32_t TC_LBO_A = TC_TILE * 16 + 32;
constexpr uint32_t TC_LBO_B = TC_F * 16;        // 64 rows x 16 B
constexpr int TC_LDH2 = TC_F + 4;               // fp32 h2 tile row pitch (floats)

constexpr uint32_t TC_OFF_B1H = 0;
constexpr uint32_t TC_OFF_B1L = TC_OFF_B1H + TC_CH1 * TC_LBO_B;
constexpr uint32_t TC_OFF_B2H = TC_OFF_B1L + TC_CH1 * TC_LBO_B;
constexpr uint32_t TC_OFF_B2L = TC_OFF_B2H + TC_CH2 * TC_LBO_B;
constexpr uint32_t TC_OFF_A = TC_OFF_B2L + TC_CH2 * TC_LBO_B;           // 40960
constexpr uint32_t TC_A_BYTES = 2 * TC_CH1 * TC_LBO_A;                   // hi + lo, 49920
constexpr uint32_t TC_OFF_IDX = TC_OFF_A + TC_A_BYTES;                   // c[128], n[128]
constexpr uint32_t TC_OFF_BIAS = TC_OFF_IDX + 2 * TC_TILE * 4;          // b1[64], b2[64]
constexpr uint32_t TC_OFF_BAR = TC_OFF_BIAS + 2 * TC_F * 4;
constexpr uint32_t TC_SMEM = TC_OFF_BAR + 16;   // two mbarriers: UMMA completion, weight image
static_assert(2 * TC_CH2 * TC_LBO_A <= TC_A_BYTES, "h1 operand tile must fit the A region");
static_assert(TC_TILE * TC_LDH2 * 4 <= TC_A_BYTES, "h2 tile must fit the A region");
static_assert(2 * TC_SMEM <= 227 * 1024, "two CTAs per SM");

// transpose + split a [k_total, 64] fp32 weight into K-major hi / lo operand tiles
__device__ __forceinline__ void stage_weight(const float* __restrict__ w, int chunks,
                                             unsigned char* hi, unsigned char* lo, int t) {
  for (int u = t; u < chunks * TC_F; u += TC_THREADS) {
    const int n = u & (TC_F - 1), j = u >> 6;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __ldg(w + (size_t)(j * 8 + e) * TC_F + n);
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + j * TC_LBO_B + n * 16) = h;
    *reinterpret_cast<uint4*>(lo + j * TC_LBO_B + n * 16) = l;
  }
}

// HL = false: feats / nfeats are fp32 [T,32] rows (split here, per pair);
// HL = true : they are bf16 rows [T,64] = [32 hi | 32 lo] written by
//             gn_block_det_fwd, copied verbatim (16-byte chunks) into the A tile.
template <bool HL>
__global__ void __launch_bounds__(TC_THREADS, 2)
block_pair_tc_kernel(const float* __restrict__ pw, const float* __restrict__ feats,
                     const float* __restrict__ nfeats, const int32_t* __restrict__ pair_c,
                     const int32_t* __restrict__ pair_n, const int32_t* __restrict__ num_pairs,
                     int capacity, const float* __restrict__ w1, const float* __restrict__ b1,
                     const float* __restrict__ w2, const float* __restrict__ b2,
                     const unsigned char* __restrict__ wimg, float* __restrict__ pooled) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + TC_TILE - 1) / TC_TILE;
  if ((int)blockIdx.x >= num_tiles) return;  // uniform per CTA: before any allocation

  unsigned char* a_hi = smem + TC_OFF_A;
  unsigned char* a_lo = a_hi + TC_CH1 * TC_LBO_A;
  unsigned char* h_hi = smem + TC_OFF_A;                       // aliases A (dead after FC1)
  unsigned char* h_lo = h_hi + TC_CH2 * TC_LBO_A;
  float* h2 = reinterpret_cast<float*>(smem + TC_OFF_A);        // aliases h1 (dead after FC2)
  int* c_idx = reinterpret_cast<int*>(smem + TC_OFF_IDX);
  unsigned* seg_end = reinterpret_cast<unsigned*>(c_idx + TC_TILE);   // 4 words: 8 x 16-row slices
  float* bias1 = reinterpret_cast<float*>(smem + TC_OFF_BIAS);
  float* bias2 = bias1 + TC_F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (t == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  uint64_t* wbar = bar + 1;
  if (wimg != nullptr) {
    // operand image prepared by gn_prepare_operands: [W1^T hi | lo | W2^T hi | lo], 40 KB,
    // laid out exactly like the shared-memory region -> one bulk copy
    if (t == 0) {
      umma::mbar_init(wbar, 1);
      umma::fence_barrier_init();
      umma::mbar_expect_tx(wbar, TC_OFF_A);
      umma::bulk_copy_g2s(umma::smem_u32(smem), wimg, TC_OFF_A, wbar);
    }
  } else {
    stage_weight(w1, TC_CH1, smem + TC_OFF_B1H, smem + TC_OFF_B1L, t);
    stage_weight(w2, TC_CH2, smem + TC_OFF_B2H, smem + TC_OFF_B2L, t);
  }
  if (t < TC_F) {
    bias1[t] = __ldg(b1 + t);
    bias2[t] = __ldg(b2 + t);
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + TC_F;
  // h1 (A operand of FC2) lives in tensor memory: bf16 pairs, 32 columns hi + 32 columns lo
  const uint32_t tmem_hh = tmem + 2 * TC_F, tmem_hl = tmem + 2 * TC_F + TC_F / 2;
  const uint32_t idesc = umma::idesc_bf16_f32(TC_TILE, TC_F);
  const uint32_t sa_hi = umma::smem_u32(a_hi), sa_lo = umma::smem_u32(a_lo);
  const uint32_t sh_hi = umma::smem_u32(h_hi), sh_lo = umma::smem_u32(h_lo);
  const uint32_t sb1h = umma::smem_u32(smem + TC_OFF_B1H), sb1l = umma::smem_u32(smem + TC_OFF_B1L);
  const uint32_t sb2h = umma::smem_u32(smem + TC_OFF_B2H), sb2l = umma::smem_u32(smem + TC_OFF_B2L);
  // kernel-lifetime operand descriptors (k-step 0); the issue loops only add offsets
  const uint64_t d_ah = umma::smem_desc(sa_hi, TC_LBO_A, TC_SBO), d_al = umma::smem_desc(sa_lo, TC_LBO_A, TC_SBO);
  const uint64_t d_hh = umma::smem_desc(sh_hi, TC_LBO_A, TC_SBO), d_hl = umma::smem_desc(sh_lo, TC_LBO_A, TC_SBO);
  const uint64_t d_b1h = umma::smem_desc(sb1h, TC_LBO_B, TC_SBO), d_b1l = umma::smem_desc(sb1l, TC_LBO_B, TC_SBO);
  const uint64_t d_b2h = umma::smem_desc(sb2h, TC_LBO_B, TC_SBO), d_b2l = umma::smem_desc(sb2l, TC_LBO_B, TC_SBO);

  // epilogue mapping: TMEM lane quadrant = warp % 4, column half = warp / 4
  const int erow = (warp & 3) * 32 + lane;
  const int ecol0 = (warp >> 2) * 32;
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;

  // The 96 inputs of a pair row are fetched one tile ahead into registers (6 units
  // of 8 floats per thread), so the L2 gather latency of tile i+1 hides behind the
  // UMMA / epilogue phases of tile i.  Warp task = 8 rows x one 128-byte part
  // (pw | c | n); lane = (row % 8) * 4 + piece: every warp request covers 8 rows x
  // 128 contiguous bytes.
  // Indices run one tile ahead of the data (two ahead of the UMMAs), so neither
  // load sits on the critical path of a tile.
  float4 pre[6][2];
  int idx_c[4], idx_n[2];          // units 2..5 need c; units 4,5 also need n
  auto prefetch_idx = [&](int tile_) {
    const int q0 = tile_ * TC_TILE;
#pragma unroll
    for (int it = 2; it < 6; ++it) {
      const int task = it * 8 + warp;
      const int p = q0 + (task & 15) * 8 + (lane >> 2);
      const bool ok = tile_ < num_tiles && p < P;
      idx_c[it - 2] = ok ? __ldg(pair_c + p) : -1;
      if (it >= 4) idx_n[it - 4] = ok ? __ldg(pair_n + p) : -1;
    }
  };
  auto prefetch = [&](int tile_) {
    const int q0 = tile_ * TC_TILE;
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const int task = it * 8 + warp;        // 0..47
      const int part = task >> 4;            // 0 pw, 1 c, 2 n   (warp uniform; it>>1 == part)
      const int row = (task & 15) * 8 + (lane >> 2);
      const int q = lane & 3;
      const int p = q0 + row;
      const float* src = nullptr;
      if (it < 2) {
        if (tile_ < num_tiles && p < P) src = pw + (size_t)p * TC_W;
      } else {
        const int c = idx_c[it - 2];
        if (c >= 0) {
          if (it < 4) src = feats + (size_t)c * TC_R;
          else {
            const int n = idx_n[it - 4];
            if (n != c) src = nfeats + (size_t)n * TC_R;   // self pair: zeros (network.py:372-374)
          }
        }
      }
      pre[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      pre[it][1] = pre[it][0];
      if (src != nullptr) {
        if (HL && part > 0) {       // 128-byte row: hi chunk q at +16q bytes, lo chunk q at +64+16q
          pre[it][0] = ldg4(src + q * 4);
          pre[it][1] = ldg4(src + 16 + q * 4);
        } else {
          pre[it][0] = ldg4(src + q * 8);
          pre[it][1] = ldg4(src + q * 8 + 4);
        }
      }
    }
  };
  prefetch_idx(blockIdx.x);
  prefetch(blockIdx.x);
  prefetch_idx(blockIdx.x + gridDim.x);
  if (wimg != nullptr) umma::mbar_wait(wbar, 0);   // weights have landed (async proxy writes)

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int p0 = tile * TC_TILE;

    // ---- 0. segment ids of the tile + "last row of its run" masks (pooling phase) ----
    if (t < TC_TILE) {
      const int p = p0 + t;
      const int c = p < P ? __ldg(pair_c + p) : -1;
      const int cnext = (p + 1 < P && (t & 15) != 15) ? __ldg(pair_c + p + 1) : -2;
      c_idx[t] = c;
      // a run ends where the next row has another c, at the end of its 16-row slice,
      // or at the last valid pair; rows past P never flush
      const unsigned m = __ballot_sync(0xffffffffu, c >= 0 && c != cnext);
      if (lane == 0) seg_end[warp] = m;     // bits 0-15: slice 2*warp, bits 16-31: slice 2*warp+1
    }

    // ---- 1. fill A (hi / lo) from the prefetched registers ---------------------------
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const int task = it * 8 + warp;
      const int part = task >> 4;
      const int row = (task & 15) * 8 + (lane >> 2);
      const int q = lane & 3;
      const float4 v0 = pre[it][0], v1 = pre[it][1];
      uint4 h, l;
      if (HL && part > 0) {
        h = make_uint4(__float_as_uint(v0.x), __float_as_uint(v0.y), __float_as_uint(v0.z), __float_as_uint(v0.w));
        l = make_uint4(__float_as_uint(v1.x), __float_as_uint(v1.y), __float_as_uint(v1.z), __float_as_uint(v1.w));
      } else {
        umma::split_bf16x2(v0.x, v0.y, h.x, l.x);
        umma::split_bf16x2(v0.z, v0.w, h.y, l.y);
        umma::split_bf16x2(v1.x, v1.y, h.z, l.z);
        umma::split_bf16x2(v1.z, v1.w, h.w, l.w);
      }
      const uint32_t off = (uint32_t)(part * 4 + q) * TC_LBO_A + (uint32_t)row * 16;
      *reinterpret_cast<uint4*>(a_hi + off) = h;
      *reinterpret_cast<uint4*>(a_lo + off) = l;
    }
    umma::fence_smem_to_async();
    umma::tc_fence_before();
    __syncthreads();

    // ---- 2. FC1 on the tensor core -------------------------------------------------
    if (t == 0) {
      umma::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < TC_K1 / 16; ++ks)
        umma::mma_bf16x3(tmem_d1, d_ah, d_al, d_b1h, d_b1l, ks * (2 * TC_LBO_A >> 4),
                         ks * (2 * TC_LBO_B >> 4), idesc, ks > 0);
      umma::mma_commit(bar);
    }
    prefetch(tile + gridDim.x);          // data of the next tile (its indices are here already)
    prefetch_idx(tile + 2 * gridDim.x);  // indices of the tile after that
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();

    // ---- 3. epilogue 1: h1 = relu(acc + b1) -> bf16 hi / lo A operand in TENSOR MEMORY --
    // (TS-form UMMA for FC2: no shared-memory round trip for h1, and FC2 reads only its
    // 2 KB weight slab per UMMA instead of 6 KB)
    {
      float v[32];
      umma::tmem_ld32(tmem_d1 + tlane + ecol0, v);
      umma::tmem_ld_wait();
      uint32_t hh[16], hl[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int col = ecol0 + 2 * e;
        umma::split_bf16x2(fmaxf(v[2 * e] + bias1[col], 0.f), fmaxf(v[2 * e + 1] + bias1[col + 1], 0.f),
                           hh[e], hl[e]);
      }
      const uint32_t c0 = (uint32_t)(ecol0 >> 1);     // 2 bf16 per 32-bit column
      umma::tmem_st8(tmem_hh + tlane + c0, reinterpret_cast<const uint32_t(&)[8]>(hh[0]));
      umma::tmem_st8(tmem_hh + tlane + c0 + 8, reinterpret_cast<const uint32_t(&)[8]>(hh[8]));
      umma::tmem_st8(tmem_hl + tlane + c0, reinterpret_cast<const uint32_t(&)[8]>(hl[0]));
      umma::tmem_st8(tmem_hl + tlane + c0 + 8, reinterpret_cast<const uint32_t(&)[8]>(hl[8]));
      umma::tmem_st_wait();
    }
    umma::tc_fence_before();
    __syncthreads();

    // ---- 4. FC2 (A from tensor memory) ----------------------------------------------------
    if (t == 0) {
      umma::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < TC_F / 16; ++ks) {
        const uint32_t boff = ks * (2 * TC_LBO_B >> 4);
        umma::mma_bf16_ts(tmem_d2, tmem_hl + ks * 8, d_b2h + boff, idesc, ks > 0);
        umma::mma_bf16_ts(tmem_d2, tmem_hh + ks * 8, d_b2l + boff, idesc, 1);
        umma::mma_bf16_ts(tmem_d2, tmem_hh + ks * 8, d_b2h + boff, idesc, 1);
      }
      umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, 1);
    umma::tc_fence_after();

    // ---- 5. epilogue 2: h2 = relu(acc + b2) -> fp32 tile ---------------------------
    {
      float v[32];
      umma::tmem_ld32(tmem_d2 + tlane + ecol0, v);
      umma::tmem_ld_wait();
      float* dst = h2 + erow * TC_LDH2 + ecol0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int col = ecol0 + g * 4;
        *reinterpret_cast<float4*>(dst + g * 4) =
            make_float4(fmaxf(v[g * 4 + 0] + bias2[col + 0], 0.f), fmaxf(v[g * 4 + 1] + bias2[col + 1], 0.f),
                        fmaxf(v[g * 4 + 2] + bias2[col + 2], 0.f), fmaxf(v[g * 4 + 3] + bias2[col + 3], 0.f));
      }
    }
    umma::tc_fence_before();
    __syncthreads();

    // ---- 6. segmented max over the tile's rows ---------------------------------------
    // thread = (16-row slice, column pair): one 8-byte shared load per row
    {
      const int j = (t & 31) * 2;      // columns j, j+1 (a warp covers all 64 columns of a slice)
      const int slice = t >> 5;        // 8 slices of 16 rows
      const int r0 = slice * 16;
      const unsigned ends = (seg_end[slice >> 1] >> ((slice & 1) * 16)) & 0xffffu;
      const float* col = h2 + r0 * TC_LDH2 + j;
      float cur0 = 0.f, cur1 = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 v = *reinterpret_cast<const float2*>(col + r * TC_LDH2);
        cur0 = fmaxf(cur0, v.x);
        cur1 = fmaxf(cur1, v.y);
        if ((ends >> r) & 1u) {        // warp uniform
          int* dst = reinterpret_cast<int*>(pooled + (size_t)c_idx[r0 + r] * TC_F + j);
          atomicMax(dst, __float_as_int(cur0));
          atomicMax(dst + 1, __float_as_int(cur1));
          cur0 = cur1 = 0.f;
        }
      }
    }
    __syncthreads();  // tile buffers free for the next fill
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace gn

static int launch_block_pair(const char* name, bool hl, const float* pw, int w, const void* feats,
                             const void* nfeats, int r, const int32_t* pair_c,
                             const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                             const float* w1, const float* b1, const float* w2, const float* b2,
                             const void* wimg, int f, float* pooled, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0, "%s: negative capacity", name);
  if (w != gn::TC_W || r != gn::TC_R || f != gn::TC_F) {
    gn::set_error("%s: fused kernel is built for w=%d r=%d f=%d (got %d, %d, %d)", name,
                  gn::TC_W, gn::TC_R, gn::TC_F, w, r, f);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(pw && feats && nfeats && pair_c && pair_n && num_pairs && b1 && b2 && pooled &&
                 ((w1 && w2) || wimg), "%s: null pointer", name);
  GN_REQUIRE(((uintptr_t)wimg & 15) == 0, "%s: weight image must be 16-byte aligned", name);
  GN_REQUIRE((((uintptr_t)pw | (uintptr_t)feats | (uintptr_t)nfeats) & 15) == 0,
             "%s: pointers must be 16-byte aligned", name);
  const void* kern = hl ? (const void*)gn::block_pair_tc_kernel<true>
                        : (const void*)gn::block_pair_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)gn::TC_SMEM);
  if (e != cudaSuccess) {
    gn::set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(capacity, gn::TC_TILE);
  int per_sm = 2;
  if (const char* e = getenv("GN_PAIR_CTAS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : 2;   // experiments only
  const int cap = per_sm * gn::sm_count();
  if (grid > cap) grid = cap;
  const float* fp = static_cast<const float*>(feats);
  const float* np = static_cast<const float*>(nfeats);
  if (hl)
    gn::block_pair_tc_kernel<true><<<grid, gn::TC_THREADS, gn::TC_SMEM, (cudaStream_t)stream>>>(
        pw, fp, np, pair_c, pair_n, num_pairs, capacity, w1, b1, w2, b2,
        static_cast<const unsigned char*>(wimg), pooled);
  else
    gn::block_pair_tc_kernel<false><<<grid, gn::TC_THREADS, gn::TC_SMEM, (cudaStream_t)stream>>>(
        pw, fp, np, pair_c, pair_n, num_pairs, capacity, w1, b1, w2, b2,
        static_cast<const unsigned char*>(wimg), pooled);
  GN_CHECK_LAUNCH(name);
  return GN_OK;
}

extern "C" int gn_block_pair_fwd(const float* pw, int w, const float* feats,
                                 const float* nfeats, int r, const int32_t* pair_c,
                                 const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                                 const float* w1, const float* b1, const float* w2,
                                 const float* b2, int f, float* pooled, gn_stream_t stream) {
  return launch_block_pair("gn_block_pair_fwd", false, pw, w, feats, nfeats, r, pair_c, pair_n,
                           num_pairs, capacity, w1, b1, w2, b2, nullptr, f, pooled, stream);
}

extern "C" int gn_block_pair_fwd_hl(const float* pw, int w, const void* feats_hl,
                                    const void* nfeats_hl, int r, const int32_t* pair_c,
                                    const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                                    const float* w1, const float* b1, const float* w2,
                                    const float* b2, const void* wimg, int f, float* pooled,
                                    gn_stream_t stream) {
  return launch_block_pair("gn_block_pair_fwd_hl", true, pw, w, feats_hl, nfeats_hl, r, pair_c,
                           pair_n, num_pairs, capacity, w1, b1, w2, b2, wimg, f, pooled, stream);
}

extern "C" int64_t gn_block_pair_image_bytes(void) { return (int64_t)gn::TC_OFF_A; }
