// Pair stage of a Gnet block with the first pair FC split by input block
// (SURVEY.md §7 restructuring (i)):
//
//   x @ W1 = pw @ W1[0:32] + feats[c] @ W1[32:64] + nfeats[n] @ W1[64:96]
//
// The last two terms depend on ONE detection each, so the detection-level kernel
// (gn_block_det_fwd_img) evaluates them once per detection:
//   AB[d, 0:64]   = feats[d]  @ W1[32:64] + b1        ("A": centre term, bias folded in)
//   AB[d, 64:128] = nfeats[d] @ W1[64:96]             ("B": neighbour term)
// and the pair kernel only runs the K = 32 product on the tensor cores:
//   h1[p] = relu(pw[p] @ W1[0:32] + A[c] + (c != n ? B[n] : 0))
//   h2[p] = relu(h1 @ W2 + b2);  pooled[c] = max over the pairs of c of h2
// Same math as network.py:367-388 up to fp32 summation order (the 1e-4 logit budget
// covers it; algorithmic flops in bench.py stay those of the reference formulation).
// Compared with gn_block_tc.cu this cuts the per-tile shared-memory traffic from
// ~245 KB to ~116 KB (A tile 16 KB instead of 48 KB, 6 instead of 18 FC1 UMMAs), which
// is what bounded that kernel; the gathers become 2 x 128 B fp32 row reads per pair
// served by L1/L2, issued before the FC1 wait.
//
// Tile = 128 pairs; two CTAs per SM (77 KB smem, 256 TMEM columns each):
//   TMEM: D1 [0,64) | D2 [64,128) | h1 hi [128,160) | h1 lo [160,192)
#include "gn_common.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int AB_TILE = 128, AB_THREADS = 256;
constexpr int AB_W = 32, AB_F = 64;
constexpr uint32_t AB_SBO = 128;
constexpr uint32_t AB_LBO_A = AB_TILE * 16 + 32;     // skewed A chunk pitch (see gn_block_tc.cu)
constexpr uint32_t AB_LBO_B = AB_F * 16;
constexpr int AB_LDH2 = AB_F + 4;

constexpr uint32_t AB_OFF_B1H = 0;                                   // W1[0:32]^T : 4 chunks x 64 rows
constexpr uint32_t AB_OFF_B1L = AB_OFF_B1H + 4 * AB_LBO_B;
constexpr uint32_t AB_OFF_B2H = AB_OFF_B1L + 4 * AB_LBO_B;           // W2^T : 8 chunks x 64 rows
constexpr uint32_t AB_OFF_B2L = AB_OFF_B2H + 8 * AB_LBO_B;
constexpr uint32_t AB_IMG_BYTES = AB_OFF_B2L + 8 * AB_LBO_B;         // 24576
constexpr uint32_t AB_OFF_A = AB_IMG_BYTES;                          // pw operand tile, hi + lo
constexpr uint32_t AB_A_BYTES = 2 * 4 * AB_LBO_A;
constexpr uint32_t AB_OFF_H2 = AB_OFF_A + AB_A_BYTES;                // fp32 h2 tile
constexpr uint32_t AB_OFF_IDX = AB_OFF_H2 + AB_TILE * AB_LDH2 * 4;   // c[128], n[128], seg_end[4]
constexpr uint32_t AB_OFF_BIAS = AB_OFF_IDX + 2 * AB_TILE * 4 + 16;  // b2[64]
constexpr uint32_t AB_OFF_BAR = AB_OFF_BIAS + AB_F * 4;
constexpr uint32_t AB_SMEM = AB_OFF_BAR + 16;
static_assert(2 * (AB_SMEM + 1024) <= 227 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(AB_THREADS, 2)
block_pair_ab_kernel(const float* __restrict__ pw, const float* __restrict__ ab,
                     const int32_t* __restrict__ pair_c, const int32_t* __restrict__ pair_n,
                     const int32_t* __restrict__ num_pairs, int capacity,
                     const float* __restrict__ b2, const unsigned char* __restrict__ wimg,
                     float* __restrict__ pooled) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + AB_TILE - 1) / AB_TILE;
  if ((int)blockIdx.x >= num_tiles) return;

  unsigned char* a_hi = smem + AB_OFF_A;
  unsigned char* a_lo = a_hi + 4 * AB_LBO_A;
  float* h2 = reinterpret_cast<float*>(smem + AB_OFF_H2);
  int* c_idx = reinterpret_cast<int*>(smem + AB_OFF_IDX);
  int* n_idx = c_idx + AB_TILE;
  unsigned* seg_end = reinterpret_cast<unsigned*>(n_idx + AB_TILE);
  float* bias2 = reinterpret_cast<float*>(smem + AB_OFF_BIAS);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + AB_OFF_BAR);
  uint64_t* wbar = bar + 1;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (t == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_init(wbar, 1);
    umma::fence_barrier_init();
    umma::mbar_expect_tx(wbar, AB_IMG_BYTES);
    umma::bulk_copy_g2s(umma::smem_u32(smem), wimg, AB_IMG_BYTES, wbar);
  }
  if (t < AB_F) bias2[t] = __ldg(b2 + t);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();

  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + AB_F;
  const uint32_t tmem_hh = tmem + 2 * AB_F, tmem_hl = tmem + 2 * AB_F + AB_F / 2;
  const uint32_t idesc = umma::idesc_bf16_f32(AB_TILE, AB_F);
  const uint32_t s0 = umma::smem_u32(smem);
  const uint64_t d_ah = umma::smem_desc(umma::smem_u32(a_hi), AB_LBO_A, AB_SBO);
  const uint64_t d_al = umma::smem_desc(umma::smem_u32(a_lo), AB_LBO_A, AB_SBO);
  const uint64_t d_b1h = umma::smem_desc(s0 + AB_OFF_B1H, AB_LBO_B, AB_SBO);
  const uint64_t d_b1l = umma::smem_desc(s0 + AB_OFF_B1L, AB_LBO_B, AB_SBO);
  const uint64_t d_b2h = umma::smem_desc(s0 + AB_OFF_B2H, AB_LBO_B, AB_SBO);
  const uint64_t d_b2l = umma::smem_desc(s0 + AB_OFF_B2L, AB_LBO_B, AB_SBO);

  // epilogue mapping: TMEM lane quadrant = warp % 4, column half = warp / 4
  const int erow = (warp & 3) * 32 + lane;
  const int ecol0 = (warp >> 2) * 32;
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;

  // pw rows of the next tile, in registers: warp task = 8 rows x 128 B, lane = (row%8)*4 + piece
  float4 pre[2][2];
  auto prefetch = [&](int tile_) {
    const int q0 = tile_ * AB_TILE;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int p = q0 + (it * 8 + warp) * 8 + (lane >> 2);
      pre[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      pre[it][1] = pre[it][0];
      if (tile_ < num_tiles && p < P) {
        const float* src = pw + (size_t)p * AB_W + (lane & 3) * 8;
        pre[it][0] = ldg4(src);
        pre[it][1] = ldg4(src + 4);
      }
    }
  };
  // pair ids of the next tile: thread t < 128 holds (c, c of the next row), t >= 128 holds n
  int nx_a = -1, nx_b = -2;
  auto prefetch_idx = [&](int tile_) {
    const int row = t & (AB_TILE - 1);
    const int p = tile_ * AB_TILE + row;
    const bool ok = tile_ < num_tiles && p < P;
    if (t < AB_TILE) {
      nx_a = ok ? __ldg(pair_c + p) : -1;
      nx_b = (ok && p + 1 < P && (row & 15) != 15) ? __ldg(pair_c + p + 1) : -2;
    } else {
      nx_a = ok ? __ldg(pair_n + p) : -1;
    }
  };
  prefetch_idx(blockIdx.x);
  prefetch(blockIdx.x);
  umma::mbar_wait(wbar, 0);   // weight image has landed

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    // ---- 0. pair ids + run-end masks of this tile -----------------------------------
    if (t < AB_TILE) {
      c_idx[t] = nx_a;
      const unsigned m = __ballot_sync(0xffffffffu, nx_a >= 0 && nx_a != nx_b);
      if (lane == 0) seg_end[warp] = m;      // bits 0-15: slice 2*warp, 16-31: slice 2*warp+1
    } else {
      n_idx[t - AB_TILE] = nx_a;
    }
    // ---- 1. fill A (pw, hi / lo) --------------------------------------------------------
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int row = (it * 8 + warp) * 8 + (lane >> 2), q = lane & 3;
      const float4 v0 = pre[it][0], v1 = pre[it][1];
      uint4 h, l;
      umma::split_bf16x2(v0.x, v0.y, h.x, l.x);
      umma::split_bf16x2(v0.z, v0.w, h.y, l.y);
      umma::split_bf16x2(v1.x, v1.y, h.z, l.z);
      umma::split_bf16x2(v1.z, v1.w, h.w, l.w);
      const uint32_t off = (uint32_t)q * AB_LBO_A + (uint32_t)row * 16;
      *reinterpret_cast<uint4*>(a_hi + off) = h;
      *reinterpret_cast<uint4*>(a_lo + off) = l;
    }
    umma::fence_smem_to_async();
    umma::tc_fence_before();
    __syncthreads();

    // ---- 2. FC1: pw @ W1[0:32] ------------------------------------------------------------
    if (t == 0) {
      umma::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < AB_W / 16; ++ks)
        umma::mma_bf16x3(tmem_d1, d_ah, d_al, d_b1h, d_b1l, ks * (2 * AB_LBO_A >> 4),
                         ks * (2 * AB_LBO_B >> 4), idesc, ks > 0);
      umma::mma_commit(bar);
    }
    // gathers of this thread's 32 columns, first 16 now (in flight across the UMMA wait)
    const int c = c_idx[erow], n = n_idx[erow];
    const bool live = c >= 0, has_n = live && n != c;     // self pair: no neighbour term
    const float* arow = ab + (size_t)(live ? c : 0) * (2 * AB_F) + ecol0;
    const float* brow = ab + (size_t)(has_n ? n : 0) * (2 * AB_F) + AB_F + ecol0;
    float4 ga[4], gb[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      ga[g] = live ? ldg4(arow + g * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      gb[g] = has_n ? ldg4(brow + g * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    prefetch(tile + gridDim.x);
    prefetch_idx(tile + gridDim.x);
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();

    // ---- 3. epilogue 1: h1 = relu(acc + A[c] + B[n]) -> bf16 hi / lo in tensor memory ------
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[16];
      umma::tmem_ld16(tmem_d1 + tlane + ecol0 + half * 16, v);
      float4 na[4], nb[4];
      if (half == 0) {      // second half's gathers overlap the first half's arithmetic
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          na[g] = live ? ldg4(arow + 16 + g * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          nb[g] = has_n ? ldg4(brow + 16 + g * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      umma::tmem_ld_wait();
      uint32_t hh[8], hl[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float x0 = fmaxf(v[g * 4 + 0] + (ga[g].x + gb[g].x), 0.f);
        const float x1 = fmaxf(v[g * 4 + 1] + (ga[g].y + gb[g].y), 0.f);
        const float x2 = fmaxf(v[g * 4 + 2] + (ga[g].z + gb[g].z), 0.f);
        const float x3 = fmaxf(v[g * 4 + 3] + (ga[g].w + gb[g].w), 0.f);
        umma::split_bf16x2(x0, x1, hh[g * 2], hl[g * 2]);
        umma::split_bf16x2(x2, x3, hh[g * 2 + 1], hl[g * 2 + 1]);
      }
      const uint32_t c0 = (uint32_t)((ecol0 + half * 16) >> 1);
      umma::tmem_st8(tmem_hh + tlane + c0, hh);
      umma::tmem_st8(tmem_hl + tlane + c0, hl);
      if (half == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          ga[g] = na[g];
          gb[g] = nb[g];
        }
      }
    }
    umma::tmem_st_wait();
    umma::tc_fence_before();
    __syncthreads();

    // ---- 4. FC2 (A from tensor memory) ----------------------------------------------------
    if (t == 0) {
      umma::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < AB_F / 16; ++ks) {
        const uint32_t boff = ks * (2 * AB_LBO_B >> 4);
        umma::mma_bf16_ts(tmem_d2, tmem_hl + ks * 8, d_b2h + boff, idesc, ks > 0);
        umma::mma_bf16_ts(tmem_d2, tmem_hh + ks * 8, d_b2l + boff, idesc, 1);
        umma::mma_bf16_ts(tmem_d2, tmem_hh + ks * 8, d_b2h + boff, idesc, 1);
      }
      umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, 1);
    umma::tc_fence_after();

    // ---- 5. epilogue 2: h2 = relu(acc + b2) -> fp32 tile ---------------------------------
    {
      float v[32];
      umma::tmem_ld32(tmem_d2 + tlane + ecol0, v);
      umma::tmem_ld_wait();
      float* dst = h2 + erow * AB_LDH2 + ecol0;
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

    // ---- 6. segmented max: thread = (16-row slice, column pair) ---------------------------
    {
      const int j = (t & 31) * 2;
      const int slice = t >> 5;
      const int r0 = slice * 16;
      const unsigned ends = (seg_end[slice >> 1] >> ((slice & 1) * 16)) & 0xffffu;
      const float* col = h2 + r0 * AB_LDH2 + j;
      float cur0 = 0.f, cur1 = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 v = *reinterpret_cast<const float2*>(col + r * AB_LDH2);
        cur0 = fmaxf(cur0, v.x);
        cur1 = fmaxf(cur1, v.y);
        if ((ends >> r) & 1u) {        // warp uniform
          int* dst = reinterpret_cast<int*>(pooled + (size_t)c_idx[r0 + r] * AB_F + j);
          atomicMax(dst, __float_as_int(cur0));
          atomicMax(dst + 1, __float_as_int(cur1));
          cur0 = cur1 = 0.f;
        }
      }
    }
    __syncthreads();  // c_idx / n_idx / h2 free for the next tile
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace gn

extern "C" int64_t gn_block_pair_ab_image_bytes(void) { return (int64_t)gn::AB_IMG_BYTES; }

extern "C" int gn_block_pair_fwd_ab(const float* pw, int w, const float* ab, int f,
                                    const int32_t* pair_c, const int32_t* pair_n,
                                    const int32_t* num_pairs, int capacity, const float* b2,
                                    const void* wimg, float* pooled, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0, "gn_block_pair_fwd_ab: negative capacity");
  if (w != gn::AB_W || f != gn::AB_F) {
    gn::set_error("gn_block_pair_fwd_ab: kernel is built for w=%d f=%d (got %d, %d)", gn::AB_W,
                  gn::AB_F, w, f);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(pw && ab && pair_c && pair_n && num_pairs && b2 && wimg && pooled,
             "gn_block_pair_fwd_ab: null pointer");
  GN_REQUIRE((((uintptr_t)pw | (uintptr_t)ab | (uintptr_t)wimg) & 15) == 0,
             "gn_block_pair_fwd_ab: pointers must be 16-byte aligned");
  cudaError_t e = cudaFuncSetAttribute(gn::block_pair_ab_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gn::AB_SMEM);
  if (e != cudaSuccess) {
    gn::set_error("gn_block_pair_fwd_ab: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(capacity, gn::AB_TILE);
  const int cap = 2 * gn::sm_count();
  if (grid > cap) grid = cap;
  gn::block_pair_ab_kernel<<<grid, gn::AB_THREADS, gn::AB_SMEM, (cudaStream_t)stream>>>(
      pw, ab, pair_c, pair_n, num_pairs, capacity, b2, static_cast<const unsigned char*>(wimg),
      pooled);
  GN_CHECK_LAUNCH("gn_block_pair_fwd_ab");
  return GN_OK;
}
