// Detection-level layers of a Gnet block on the tensor cores, fused across the
// block boundary (network.py:344-409):
//
//   stage A (post-pool of block b, skipped when `pooled` is null)
//     d1        = relu(pooled @ W_fc1 + b_fc1)                      64 -> 64    (:390-397)
//     feats_out = relu(feats_in + d1 @ W_fc2 + b_fc2)               64 -> 128   (:399-408)
//     pooled   <- 0   (the next block max-accumulates into it with atomicMax)
//   stage B (reduce_dim of block b+1, skipped when `w_rd` is null)
//     red       = relu(feats_out @ W_rd + b_rd)                     128 -> 32   (:348-354)
//     written as fp32 [T,32] and/or as bf16 (hi | lo) rows [T,64]: the operand
//     format the pair kernel copies straight into its A tile.
//
// Tile = 128 detections = M of one tcgen05.mma; persistent CTAs, weights (hi/lo,
// K-major) staged once per CTA.  Every GEMM is bf16x3 with fp32 accumulation in
// TMEM (gn_umma.cuh); epilogues run on 8 warps (TMEM lane quadrant = warp % 4,
// column half = warp / 4) and write the next GEMM's A operand back to shared
// memory, so one tile makes a single pass over HBM: read pooled + feats_in, write
// feats_out + red.
#include "gn_common.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int DT_TILE = 128, DT_THREADS = 256;
constexpr int DT_D = 128, DT_F = 64, DT_R = 32;       // shortcut, pairfeat, reduced dims
constexpr uint32_t DT_SBO = 128;
constexpr uint32_t DT_LBO_A = DT_TILE * 16 + 32;      // skewed like the pair kernel's A tile
constexpr uint32_t DT_OFF_W1H = 0;                                   // fc1^T: 8 chunks x 64 rows
constexpr uint32_t DT_OFF_W1L = DT_OFF_W1H + 8 * DT_F * 16;
constexpr uint32_t DT_OFF_W2H = DT_OFF_W1L + 8 * DT_F * 16;          // fc2^T: 8 chunks x 128 rows
constexpr uint32_t DT_OFF_W2L = DT_OFF_W2H + 8 * DT_D * 16;
constexpr uint32_t DT_OFF_WRH = DT_OFF_W2L + 8 * DT_D * 16;          // rd^T: 16 chunks x 32 rows
constexpr uint32_t DT_OFF_WRL = DT_OFF_WRH + 16 * DT_R * 16;
constexpr uint32_t DT_OFF_WABH = DT_OFF_WRL + 16 * DT_R * 16;        // [W1[32:64] | W1[64:96]]^T of
constexpr uint32_t DT_OFF_WABL = DT_OFF_WABH + 4 * DT_D * 16;        //   the next block's pw_fc1: 4 chunks x 128 rows
constexpr uint32_t DT_OFF_A = DT_OFF_WABL + 4 * DT_D * 16;           // 81920 = weight image bytes
constexpr uint32_t DT_A_BYTES = 2 * 16 * DT_LBO_A;                   // K up to 128, hi + lo
constexpr uint32_t DT_OFF_BIAS = DT_OFF_A + DT_A_BYTES;              // b_fc1[64] b_fc2[128] b_rd[32] b_ab[64]
constexpr uint32_t DT_OFF_BAR = DT_OFF_BIAS + (DT_F + DT_D + DT_R + DT_F) * 4;
// feats tile staging [128 rows][128 fp32], row pitch 512 + 16 bytes: the shortcut operand
// comes in and the block output leaves as whole 512-byte rows (one warp instruction = one
// row = 4 L1 wavefronts) and is transposed to / from the row-per-thread epilogue layout
// through shared memory (16-byte accesses at a 528-byte pitch are conflict free).  A
// row-per-thread LDG / STG costs one wavefront per lane: 8 192 wavefront cycles per tile.
constexpr uint32_t DT_STAGE_PITCH = DT_D * 4 + 16;
constexpr uint32_t DT_OFF_STAGE = (DT_OFF_BAR + 16 + 127) / 128 * 128;
constexpr uint32_t DT_SMEM = DT_OFF_STAGE + DT_TILE * DT_STAGE_PITCH;   // two mbarriers before it
static_assert(DT_SMEM <= 227 * 1024, "det tile exceeds shared memory");

// transpose + split w[k_total, n_total] (fp32, [in,out]) into K-major hi/lo tiles
__device__ __forceinline__ void dt_stage_weight(const float* __restrict__ w, int k_total,
                                                int n_total, unsigned char* hi, unsigned char* lo,
                                                int t) {
  const int chunks = k_total / 8;
  for (int u = t; u < chunks * n_total; u += DT_THREADS) {
    const int n = u % n_total, j = u / n_total;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __ldg(w + (size_t)(j * 8 + e) * n_total + n);
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + (size_t)j * n_total * 16 + n * 16) = h;
    *reinterpret_cast<uint4*>(lo + (size_t)j * n_total * 16 + n * 16) = l;
  }
}

// load a [128 x width] fp32 tile (row pitch = width floats) into the A operand region
template <int WIDTH>
__device__ __forceinline__ void dt_load_tile(const float* __restrict__ src, int row0, int rows,
                                             unsigned char* a_hi, unsigned char* a_lo, int warp,
                                             int lane, float* __restrict__ zero_after) {
  constexpr int PIECES = WIDTH / 8;           // 8-float pieces per row
  constexpr int ROWS_PER_REQ = 32 / PIECES;   // lane -> (row within request, piece)
  constexpr int ITERS = DT_TILE / ((DT_THREADS / 32) * ROWS_PER_REQ);
  // all loads first (ITERS x 32 bytes in flight per thread), then the stores / splits:
  // this phase is pure memory latency, so it is paid once per tile, not once per row group
  float4 v[ITERS][2];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int r = (it * (DT_THREADS / 32) + warp) * ROWS_PER_REQ + lane / PIECES, q = lane % PIECES;
    v[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
    v[it][1] = v[it][0];
    if (row0 + r < rows) {
      const float* p = src + (size_t)(row0 + r) * WIDTH + q * 8;
      v[it][0] = ldg4(p);
      v[it][1] = ldg4(p + 4);
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int r = (it * (DT_THREADS / 32) + warp) * ROWS_PER_REQ + lane / PIECES, q = lane % PIECES;
    if (zero_after != nullptr && row0 + r < rows) {
      float* z = zero_after + (size_t)(row0 + r) * WIDTH + q * 8;
      *reinterpret_cast<float4*>(z) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(z + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint4 h, l;
    umma::split_bf16x2(v[it][0].x, v[it][0].y, h.x, l.x);
    umma::split_bf16x2(v[it][0].z, v[it][0].w, h.y, l.y);
    umma::split_bf16x2(v[it][1].x, v[it][1].y, h.z, l.z);
    umma::split_bf16x2(v[it][1].z, v[it][1].w, h.w, l.w);
    const uint32_t off = (uint32_t)q * DT_LBO_A + (uint32_t)r * 16;
    *reinterpret_cast<uint4*>(a_hi + off) = h;
    *reinterpret_cast<uint4*>(a_lo + off) = l;
  }
}

// bf16x3 GEMM: D[tmem] = A[128 x K] (smem) * B[N x K]^T (smem); issued by one thread
// X3 = false: plain bf16 operands (hi parts only), one UMMA per k-step
template <int KSTEPS, bool X3>
__device__ __forceinline__ void dt_gemm(uint32_t tmem_d, uint64_t d_ah, uint64_t d_al,
                                        uint64_t d_bh, uint64_t d_bl, uint32_t lbo_b,
                                        uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    if (X3)
      umma::mma_bf16x3(tmem_d, d_ah, d_al, d_bh, d_bl, ks * (2 * DT_LBO_A >> 4),
                       ks * (2 * lbo_b >> 4), idesc, ks > 0);
    else
      umma::mma_bf16_ss(tmem_d, d_ah + ks * (2 * DT_LBO_A >> 4), d_bh + ks * (2 * lbo_b >> 4), idesc,
                        ks > 0);
  }
}

#ifdef DT_TRACE
__device__ long long dt_trace[32];
#define DT_TR(i) do { if (blockIdx.x == 0 && t == 0 && tile == (int)(2 * gridDim.x)) dt_trace[i] = clock64(); } while (0)
#else
#define DT_TR(i) do { } while (0)
#endif

template <bool X3>
__global__ void __launch_bounds__(DT_THREADS, 1)
block_det_tc_kernel(float* __restrict__ pooled, const float* __restrict__ feats_in,
                    const float* __restrict__ w_fc1, const float* __restrict__ b_fc1,
                    const float* __restrict__ w_fc2, const float* __restrict__ b_fc2,
                    const float* __restrict__ w_rd, const float* __restrict__ b_rd,
                    const unsigned char* __restrict__ wimg, float* __restrict__ feats_out,
                    float* __restrict__ red_f32, __nv_bfloat16* __restrict__ red_hl,
                    const float* __restrict__ b_ab, float* __restrict__ ab_out, int ab_cols,
                    int num_dets, int has_a, int has_b) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int num_tiles = (num_dets + DT_TILE - 1) / DT_TILE;
  if ((int)blockIdx.x >= num_tiles) return;
  const bool stage_a = has_a != 0, stage_b = has_b != 0;

  unsigned char* a_hi = smem + DT_OFF_A;
  unsigned char* a_lo = a_hi + 16 * DT_LBO_A;
  float* bias1 = reinterpret_cast<float*>(smem + DT_OFF_BIAS);
  float* bias2 = bias1 + DT_F;
  float* biasr = bias2 + DT_D;
  float* biasab = biasr + DT_R;
  const bool stage_ab = ab_out != nullptr;   // needs the prepared image (W_ab part)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + DT_OFF_BAR);

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (t == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  uint64_t* wbar = bar + 1;
  unsigned char* stage = smem + DT_OFF_STAGE;
  if (wimg != nullptr) {
    // operand image from gn_prepare_operands, laid out like the shared-memory weight
    // region [fc1^T hi|lo, fc2^T hi|lo, rd^T hi|lo] (64 KB): two bulk copies
    if (t == 0) {
      umma::mbar_init(wbar, 1);
      umma::fence_barrier_init();
      umma::mbar_expect_tx(wbar, DT_OFF_A);
      umma::bulk_copy_g2s(umma::smem_u32(smem), wimg, DT_OFF_A / 2, wbar);
      umma::bulk_copy_g2s(umma::smem_u32(smem) + DT_OFF_A / 2, wimg + DT_OFF_A / 2, DT_OFF_A / 2, wbar);
    }
  } else {
    if (stage_a) {
      dt_stage_weight(w_fc1, DT_F, DT_F, smem + DT_OFF_W1H, smem + DT_OFF_W1L, t);
      dt_stage_weight(w_fc2, DT_F, DT_D, smem + DT_OFF_W2H, smem + DT_OFF_W2L, t);
    }
    if (stage_b) dt_stage_weight(w_rd, DT_D, DT_R, smem + DT_OFF_WRH, smem + DT_OFF_WRL, t);
  }
  if (stage_a) {
    if (t < DT_F) bias1[t] = __ldg(b_fc1 + t);
    if (t < DT_D) bias2[t] = __ldg(b_fc2 + t);
  }
  if (stage_b && t < DT_R) biasr[t] = __ldg(b_rd + t);
  if (stage_ab && t < DT_F) biasab[t] = __ldg(b_ab + t);
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();

  const uint32_t tmem = tmem_base_s;
  const uint32_t tm1 = tmem, tm2 = tmem + 64, tmr = tmem + 192;
  const uint32_t sa_hi = umma::smem_u32(a_hi), sa_lo = umma::smem_u32(a_lo);
  const uint32_t s_w1h = umma::smem_u32(smem + DT_OFF_W1H), s_w1l = umma::smem_u32(smem + DT_OFF_W1L);
  const uint32_t s_w2h = umma::smem_u32(smem + DT_OFF_W2H), s_w2l = umma::smem_u32(smem + DT_OFF_W2L);
  const uint32_t s_wrh = umma::smem_u32(smem + DT_OFF_WRH), s_wrl = umma::smem_u32(smem + DT_OFF_WRL);
  const uint64_t d_ah = umma::smem_desc(sa_hi, DT_LBO_A, DT_SBO), d_al = umma::smem_desc(sa_lo, DT_LBO_A, DT_SBO);
  const uint64_t d_w1h = umma::smem_desc(s_w1h, DT_F * 16, DT_SBO), d_w1l = umma::smem_desc(s_w1l, DT_F * 16, DT_SBO);
  const uint64_t d_w2h = umma::smem_desc(s_w2h, DT_D * 16, DT_SBO), d_w2l = umma::smem_desc(s_w2l, DT_D * 16, DT_SBO);
  const uint64_t d_wrh = umma::smem_desc(s_wrh, DT_R * 16, DT_SBO), d_wrl = umma::smem_desc(s_wrl, DT_R * 16, DT_SBO);
  const uint64_t d_wabh = umma::smem_desc(umma::smem_u32(smem + DT_OFF_WABH), DT_D * 16, DT_SBO);
  const uint64_t d_wabl = umma::smem_desc(umma::smem_u32(smem + DT_OFF_WABL), DT_D * 16, DT_SBO);
  const int erow = (warp & 3) * 32 + lane;
  const int ehalf = warp >> 2;
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t par = 0;
  bool weights_pending = wimg != nullptr;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row0 = tile * DT_TILE;
    const int grow = row0 + erow;
    const bool live = grow < num_dets;
    DT_TR(0);

    if (stage_a) {
      // ---- pooled tile -> A (K = 64); pooled <- 0 for the next block ----------------
      dt_load_tile<DT_F>(pooled, row0, num_dets, a_hi, a_lo, warp, lane, pooled);
      DT_TR(1);
      if (weights_pending) { umma::mbar_wait(wbar, 0); weights_pending = false; }
      umma::fence_smem_to_async();
      umma::tc_fence_before();
      __syncthreads();
      DT_TR(2);
      if (t == 0) {
        umma::tc_fence_after();
        dt_gemm<DT_F / 16, X3>(tm1, d_ah, d_al, d_w1h, d_w1l, DT_F * 16, umma::idesc_bf16_f32(DT_TILE, DT_F));
        umma::mma_commit(bar);
      }
      umma::mbar_wait(bar, par);
      par ^= 1;
      umma::tc_fence_after();
      DT_TR(3);
      // ---- d1 = relu(acc + b_fc1) -> A (K = 64) ---------------------------------------
#pragma unroll
      for (int cc = 0; cc < 32; cc += 16) {
        const int col0 = ehalf * 32 + cc;
        float v[16];
        umma::tmem_ld16(tm1 + tlane + col0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = fmaxf(v[g * 8 + e] + bias1[col0 + g * 8 + e], 0.f);
          uint4 h, l;
          umma::split_bf16x2(x[0], x[1], h.x, l.x);
          umma::split_bf16x2(x[2], x[3], h.y, l.y);
          umma::split_bf16x2(x[4], x[5], h.z, l.z);
          umma::split_bf16x2(x[6], x[7], h.w, l.w);
          const uint32_t off = (uint32_t)((col0 >> 3) + g) * DT_LBO_A + (uint32_t)erow * 16;
          *reinterpret_cast<uint4*>(a_hi + off) = h;
          *reinterpret_cast<uint4*>(a_lo + off) = l;
        }
      }
      DT_TR(4);
      umma::fence_smem_to_async();
      umma::tc_fence_before();
      __syncthreads();
      DT_TR(5);
      if (t == 0) {
        umma::tc_fence_after();
        dt_gemm<DT_F / 16, X3>(tm2, d_ah, d_al, d_w2h, d_w2l, DT_D * 16, umma::idesc_bf16_f32(DT_TILE, DT_D));
        umma::mma_commit(bar);
      }
      // the shortcut tile (feats_in rows) is fetched while the fc2 UMMAs run: every warp
      // instruction reads one whole 512-byte row (16 rows per warp, all in flight), then the
      // rows go to the staging tile for the row-per-thread epilogue
      {
        float4 rowv[16];
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const int r = g * 8 + warp;
          rowv[g] = (row0 + r < num_dets) ? ldg4(feats_in + (size_t)(row0 + r) * DT_D + lane * 4)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int g = 0; g < 16; ++g)
          *reinterpret_cast<float4*>(stage + (g * 8 + warp) * DT_STAGE_PITCH + lane * 16) = rowv[g];
      }
      DT_TR(6);
      __syncthreads();
      DT_TR(7);
      float* srow = reinterpret_cast<float*>(stage + erow * DT_STAGE_PITCH) + ehalf * 64;
      umma::mbar_wait(bar, par);
      par ^= 1;
      umma::tc_fence_after();
      DT_TR(8);
      // ---- feats_out = relu(feats_in + acc + b_fc2) -> staging (in place), and -> A (K = 128)
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        const int col0 = ehalf * 64 + cc;
        float v[16];
        umma::tmem_ld16(tm2 + tlane + col0, v);
        float4 res[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) res[g] = *reinterpret_cast<const float4*>(srow + cc + g * 4);
        umma::tmem_ld_wait();
        float x[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          x[g * 4 + 0] = fmaxf(res[g].x + (v[g * 4 + 0] + bias2[col0 + g * 4 + 0]), 0.f);
          x[g * 4 + 1] = fmaxf(res[g].y + (v[g * 4 + 1] + bias2[col0 + g * 4 + 1]), 0.f);
          x[g * 4 + 2] = fmaxf(res[g].z + (v[g * 4 + 2] + bias2[col0 + g * 4 + 2]), 0.f);
          x[g * 4 + 3] = fmaxf(res[g].w + (v[g * 4 + 3] + bias2[col0 + g * 4 + 3]), 0.f);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<float4*>(srow + cc + g * 4) =
              make_float4(x[g * 4 + 0], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
        if (stage_b) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 h, l;
            umma::split_bf16x2(x[g * 8 + 0], x[g * 8 + 1], h.x, l.x);
            umma::split_bf16x2(x[g * 8 + 2], x[g * 8 + 3], h.y, l.y);
            umma::split_bf16x2(x[g * 8 + 4], x[g * 8 + 5], h.z, l.z);
            umma::split_bf16x2(x[g * 8 + 6], x[g * 8 + 7], h.w, l.w);
            const uint32_t off = (uint32_t)((col0 >> 3) + g) * DT_LBO_A + (uint32_t)erow * 16;
            *reinterpret_cast<uint4*>(a_hi + off) = h;
            *reinterpret_cast<uint4*>(a_lo + off) = l;
          }
        }
      }
    } else {
      // block 1 / stand-alone reduce: feats_in tile -> A (K = 128)
      dt_load_tile<DT_D>(feats_in, row0, num_dets, a_hi, a_lo, warp, lane, nullptr);
      if (weights_pending) { umma::mbar_wait(wbar, 0); weights_pending = false; }
    }

    // block output: staging rows -> feats_out, one whole 512-byte row per warp instruction
    auto store_feats_tile = [&]() {
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const int r = g * 8 + warp;
        const float4 x = *reinterpret_cast<const float4*>(stage + r * DT_STAGE_PITCH + lane * 16);
        if (row0 + r < num_dets)
          *reinterpret_cast<float4*>(feats_out + (size_t)(row0 + r) * DT_D + lane * 4) = x;
      }
    };
    if (stage_a && !stage_b) {
      __syncthreads();
      store_feats_tile();
    }
    if (stage_b) {
      DT_TR(9);
      umma::fence_smem_to_async();
      umma::tc_fence_before();
      __syncthreads();
      DT_TR(10);
      if (t == 0) {
        umma::tc_fence_after();
        dt_gemm<DT_D / 16, X3>(tmr, d_ah, d_al, d_wrh, d_wrl, DT_R * 16, umma::idesc_bf16_f32(DT_TILE, DT_R));
        umma::mma_commit(bar);
      }
      if (stage_a) store_feats_tile();      // while the reduce_dim UMMAs run
      DT_TR(11);
      umma::mbar_wait(bar, par);
      par ^= 1;
      umma::tc_fence_after();
      DT_TR(12);
      // ---- red = relu(acc + b_rd): 16 columns per thread --------------------------------
      {
        const int col0 = ehalf * 16;
        float v[16];
        umma::tmem_ld16(tmr + tlane + col0, v);
        umma::tmem_ld_wait();
        float x[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) x[e] = fmaxf(v[e] + biasr[col0 + e], 0.f);
        if (live) {
          if (red_f32 != nullptr) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<float4*>(red_f32 + (size_t)grow * DT_R + col0 + g * 4) =
                  make_float4(x[g * 4 + 0], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
          }
          if (red_hl != nullptr) {
            // row layout: [32 hi | 32 lo] bf16 = 128 B; this thread owns 16 of the 32
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 h, l;
              umma::split_bf16x2(x[g * 8 + 0], x[g * 8 + 1], h.x, l.x);
              umma::split_bf16x2(x[g * 8 + 2], x[g * 8 + 3], h.y, l.y);
              umma::split_bf16x2(x[g * 8 + 4], x[g * 8 + 5], h.z, l.z);
              umma::split_bf16x2(x[g * 8 + 6], x[g * 8 + 7], h.w, l.w);
              __nv_bfloat16* row = red_hl + (size_t)grow * 2 * DT_R;
              *reinterpret_cast<uint4*>(row + col0 + g * 8) = h;
              *reinterpret_cast<uint4*>(row + DT_R + col0 + g * 8) = l;
            }
          }
        }
        if (stage_ab) {
          // red -> A operand (K = 32) of the per-detection halves of the next pair FC
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 h, l;
            umma::split_bf16x2(x[g * 8 + 0], x[g * 8 + 1], h.x, l.x);
            umma::split_bf16x2(x[g * 8 + 2], x[g * 8 + 3], h.y, l.y);
            umma::split_bf16x2(x[g * 8 + 4], x[g * 8 + 5], h.z, l.z);
            umma::split_bf16x2(x[g * 8 + 6], x[g * 8 + 7], h.w, l.w);
            const uint32_t off = (uint32_t)((col0 >> 3) + g) * DT_LBO_A + (uint32_t)erow * 16;
            *reinterpret_cast<uint4*>(a_hi + off) = h;
            *reinterpret_cast<uint4*>(a_lo + off) = l;
          }
        }
      }
      if (stage_ab) {
        // AB[d, 0:64] = red @ W1[32:64] + b1 ; AB[d, 64:128] = red @ W1[64:96]   (gn_block_ab.cu)
        // ab_cols = 64: only the first half (gn_block_tma.cu gathers the neighbor rows itself)
        DT_TR(13);
        umma::fence_smem_to_async();
        umma::tc_fence_before();
        __syncthreads();
        DT_TR(14);
        if (t == 0) {
          umma::tc_fence_after();
          dt_gemm<DT_R / 16, X3>(tm2, d_ah, d_al, d_wabh, d_wabl, DT_D * 16, umma::idesc_bf16_f32(DT_TILE, ab_cols));
          umma::mma_commit(bar);
        }
        umma::mbar_wait(bar, par);
        par ^= 1;
        umma::tc_fence_after();
        DT_TR(15);
        const int per = ab_cols >> 1;          // columns per warp half: 64 or 32
        for (int cc = 0; cc < per; cc += 32) {
          const int col0 = ehalf * per + cc;
          float v[32];
          umma::tmem_ld32(tm2 + tlane + col0, v);
          umma::tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int col = col0 + g * 4;
              float4 o = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
              if (col < DT_F) {          // uniform per warp
                o.x += biasab[col]; o.y += biasab[col + 1]; o.z += biasab[col + 2]; o.w += biasab[col + 3];
              }
              *reinterpret_cast<float4*>(ab_out + (size_t)grow * ab_cols + col) = o;
            }
          }
        }
      }
    }
    DT_TR(16);
    umma::tc_fence_before();
    __syncthreads();   // A region and TMEM columns are reused by the next tile
    DT_TR(17);
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace gn

static int launch_block_det(const char* name, bool x3, float* pooled, const float* feats_in,
                            const float* w_fc1, const float* b_fc1, const float* w_fc2,
                            const float* b_fc2, const float* w_rd, const float* b_rd,
                            const void* wimg, int has_a, int has_b, float* feats_out,
                            float* red_f32, void* red_hl, const float* b_ab, float* ab_out,
                            int ab_cols, int num_dets, int shortcut_dim, int pairfeat_dim,
                            int reduced_dim, gn_stream_t stream) {
  GN_REQUIRE(num_dets >= 0, "%s: negative size", name);
  if (shortcut_dim != gn::DT_D || pairfeat_dim != gn::DT_F || reduced_dim != gn::DT_R) {
    gn::set_error("%s: fused kernel is built for d=%d f=%d r=%d (got %d, %d, %d)", name,
                  gn::DT_D, gn::DT_F, gn::DT_R, shortcut_dim, pairfeat_dim, reduced_dim);
    return GN_ERR_UNSUPPORTED;
  }
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(feats_in != nullptr, "%s: null feats_in", name);
  GN_REQUIRE(has_a || has_b, "%s: nothing to do", name);
  GN_REQUIRE(!has_a || (pooled && b_fc1 && b_fc2 && feats_out && ((w_fc1 && w_fc2) || wimg)),
             "%s: stage A needs pooled, fc1 / fc2 parameters and feats_out", name);
  GN_REQUIRE(!has_b || (b_rd && (red_f32 || red_hl || ab_out) && (w_rd || wimg)),
             "%s: stage B needs reduce_dim parameters and an output", name);
  GN_REQUIRE(ab_out == nullptr || (has_b && wimg && b_ab),
             "%s: the AB output needs stage B, the prepared image and the pw_fc1 bias", name);
  GN_REQUIRE(((uintptr_t)ab_out & 15) == 0, "%s: ab_out must be 16-byte aligned", name);
  GN_REQUIRE(ab_out == nullptr || ab_cols == 64 || ab_cols == 128, "%s: ab_cols must be 64 or 128", name);
  GN_REQUIRE((((uintptr_t)pooled | (uintptr_t)feats_in | (uintptr_t)feats_out |
               (uintptr_t)red_f32 | (uintptr_t)red_hl | (uintptr_t)wimg) & 15) == 0,
             "%s: pointers must be 16-byte aligned", name);
  const void* kern = x3 ? (const void*)gn::block_det_tc_kernel<true>
                        : (const void*)gn::block_det_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)gn::DT_SMEM);
  if (e != cudaSuccess) {
    gn::set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(num_dets, gn::DT_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  if (x3)
    gn::block_det_tc_kernel<true><<<grid, gn::DT_THREADS, gn::DT_SMEM, (cudaStream_t)stream>>>(
        pooled, feats_in, w_fc1, b_fc1, w_fc2, b_fc2, w_rd, b_rd,
        static_cast<const unsigned char*>(wimg), feats_out, red_f32,
        static_cast<__nv_bfloat16*>(red_hl), b_ab, ab_out, ab_cols, num_dets, has_a, has_b);
  else
    gn::block_det_tc_kernel<false><<<grid, gn::DT_THREADS, gn::DT_SMEM, (cudaStream_t)stream>>>(
        pooled, feats_in, w_fc1, b_fc1, w_fc2, b_fc2, w_rd, b_rd,
        static_cast<const unsigned char*>(wimg), feats_out, red_f32,
        static_cast<__nv_bfloat16*>(red_hl), b_ab, ab_out, ab_cols, num_dets, has_a, has_b);
  GN_CHECK_LAUNCH(name);
  return GN_OK;
}

extern "C" int gn_block_det_fwd(float* pooled, const float* feats_in, const float* w_fc1,
                                const float* b_fc1, const float* w_fc2, const float* b_fc2,
                                const float* w_rd, const float* b_rd, float* feats_out,
                                float* red_f32, void* red_hl, int num_dets, int shortcut_dim,
                                int pairfeat_dim, int reduced_dim, gn_stream_t stream) {
  return launch_block_det("gn_block_det_fwd", true, pooled, feats_in, w_fc1, b_fc1, w_fc2, b_fc2, w_rd,
                          b_rd, nullptr, pooled != nullptr, w_rd != nullptr, feats_out, red_f32,
                          red_hl, nullptr, nullptr, 128, num_dets, shortcut_dim, pairfeat_dim,
                          reduced_dim, stream);
}

extern "C" int gn_block_det_fwd_img(float* pooled, const float* feats_in, const void* wimg,
                                    const float* b_fc1, const float* b_fc2, const float* b_rd,
                                    int has_stage_a, int has_stage_b, float* feats_out,
                                    float* red_f32, void* red_hl, const float* b_ab,
                                    float* ab_out, int num_dets, int shortcut_dim,
                                    int pairfeat_dim, int reduced_dim, gn_stream_t stream) {
  GN_REQUIRE(wimg != nullptr, "gn_block_det_fwd_img: null weight image");
  return launch_block_det("gn_block_det_fwd_img", true, pooled, feats_in, nullptr, b_fc1, nullptr,
                          b_fc2, nullptr, b_rd, wimg, has_stage_a, has_stage_b, feats_out, red_f32,
                          red_hl, b_ab, ab_out, 128, num_dets, shortcut_dim, pairfeat_dim,
                          reduced_dim, stream);
}

extern "C" int gn_block_det_fwd_img_bf16(float* pooled, const float* feats_in, const void* wimg,
                                         const float* b_fc1, const float* b_fc2, const float* b_rd,
                                         int has_stage_a, int has_stage_b, float* feats_out,
                                         float* red_f32, void* red_hl, const float* b_ab,
                                         float* ab_out, int num_dets, int shortcut_dim,
                                         int pairfeat_dim, int reduced_dim, gn_stream_t stream) {
  GN_REQUIRE(wimg != nullptr, "gn_block_det_fwd_img_bf16: null weight image");
  return launch_block_det("gn_block_det_fwd_img_bf16", false, pooled, feats_in, nullptr, b_fc1,
                          nullptr, b_fc2, nullptr, b_rd, wimg, has_stage_a, has_stage_b, feats_out,
                          red_f32, red_hl, b_ab, ab_out, 128, num_dets, shortcut_dim, pairfeat_dim,
                          reduced_dim, stream);
}

// The same kernel writing only U = red @ pw_fc1[32:64] + b_pw_fc1 as u_out[T, 64]: the
// detection-level term of the TMA-fed pair stage (gn_block_tma.cu), which gathers the
// neighbor rows itself.  plain_bf16 != 0: the bf16 arithmetic.
extern "C" int gn_block_det_fwd_img_u(float* pooled, const float* feats_in, const void* wimg,
                                      const float* b_fc1, const float* b_fc2, const float* b_rd,
                                      int has_stage_a, int has_stage_b, float* feats_out,
                                      void* red_hl, const float* b_u, float* u_out, int plain_bf16,
                                      int num_dets, int shortcut_dim, int pairfeat_dim,
                                      int reduced_dim, gn_stream_t stream) {
  GN_REQUIRE(wimg != nullptr, "gn_block_det_fwd_img_u: null weight image");
  return launch_block_det("gn_block_det_fwd_img_u", plain_bf16 == 0, pooled, feats_in, nullptr, b_fc1,
                          nullptr, b_fc2, nullptr, b_rd, wimg, has_stage_a, has_stage_b, feats_out,
                          nullptr, red_hl, b_u, u_out, 64, num_dets, shortcut_dim, pairfeat_dim,
                          reduced_dim, stream);
}

#ifdef DT_TRACE
extern "C" int gn_block_det_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::dt_trace, sizeof(long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int64_t gn_block_det_image_bytes(void) { return (int64_t)gn::DT_OFF_A; }

// ---------------------------------------------------------------------------------
// gn_prepare_operands: fp32 [k, n] weights ([in, out]) of the flat parameter buffer ->
// bf16 hi / lo K-major operand tiles (chunk j of row n at j * n * 16 + n_row * 16).
// table: 6 int32 per entry: src offset (floats), k, n, dst_hi offset, dst_lo offset (bytes),
// chunk pitch in bytes (0 = n * 16; larger when several matrices share one N-wide tile).
// One launch converts every block's weights of a forward pass.
// ---------------------------------------------------------------------------------
namespace gn {
__global__ void prepare_operands_kernel(const float* __restrict__ flat,
                                        const int32_t* __restrict__ table,
                                        unsigned char* __restrict__ image) {
  const int32_t* e = table + blockIdx.y * 6;
  const float* w = flat + e[0];
  const int k = e[1], n = e[2];
  unsigned char* hi = image + e[3];
  unsigned char* lo = image + e[4];
  const size_t pitch = e[5] > 0 ? (size_t)e[5] : (size_t)n * 16;
  const int units = (k / 8) * n;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int col = u % n, j = u / n;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ldg(w + (size_t)(j * 8 + i) * n + col);
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + j * pitch + col * 16) = h;
    *reinterpret_cast<uint4*>(lo + j * pitch + col * 16) = l;
  }
}
}  // namespace gn

extern "C" int gn_prepare_operands(const float* flat_params, const int32_t* table, int entries,
                                   void* image, gn_stream_t stream) {
  GN_REQUIRE(entries >= 0, "gn_prepare_operands: negative entry count");
  if (entries == 0) return GN_OK;
  GN_REQUIRE(flat_params && table && image, "gn_prepare_operands: null pointer");
  GN_REQUIRE(((uintptr_t)image & 15) == 0, "gn_prepare_operands: image must be 16-byte aligned");
  gn::prepare_operands_kernel<<<dim3(8, entries), 256, 0, (cudaStream_t)stream>>>(
      flat_params, table, static_cast<unsigned char*>(image));
  GN_CHECK_LAUNCH("gn_prepare_operands");
  return GN_OK;
}
