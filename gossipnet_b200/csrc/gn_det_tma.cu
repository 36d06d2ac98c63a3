// Detection-level layers of a Gnet block with every tile transfer on the copy engine
// (same arithmetic as gn_det_tc.cu, network.py:390-408 of block b and :348-354 of block b+1):
//
//   d1        = relu(pooled @ W_fc1 + b_fc1)                       64 -> 64
//   feats_out = relu(feats_in + d1 @ W_fc2 + b_fc2)                64 -> 128
//   pooled   <- 0
//   red       = relu(feats_out @ W_rd + b_rd)                      128 -> 32   (bf16 hi | lo rows)
//   u         = red @ W_pw_fc1[32:64] + b_pw_fc1                   32 -> 64    (gn_block_tma.cu)
//
// gn_det_tc.cu spends ~18 000 of its ~23 000 cycles per 128-detection tile waiting for memory:
// the pooled tile, the shortcut rows, the block output and the U rows are moved by the same
// eight warps that run the epilogues, one phase after the other (clock64 trace,
// profiles/r2_det_pwfeat.md).  Here the four dependent GEMMs and their epilogues are all those
// warps do:
//   * a ninth warp drives tensor-map TMA: the shortcut tile of the NEXT tile is loaded while
//     this one computes (4 boxes of 128 rows x 32 floats, SWIZZLE_128B - the row-per-thread
//     epilogue reads and writes them without bank conflicts), the block output, U and the
//     bf16 red rows leave as TMA stores from shared memory (cp.async.bulk.tensor ... bulk_group);
//     rows beyond num_dets are zero filled on the way in and clipped on the way out;
//   * the pooled rows of the next tile are prefetched into registers (32 per thread) behind
//     the second epilogue and zeroed in global memory when they are consumed.
// Shared memory: weights 72 KB | operand region 65 KB | biases | feats tile 64 KB | red box 16 KB.
// The operand region holds, in turn: pooled / d1 (K = 64, upper half R1), the feats_out operand
// (K = 128, all of it), red (K = 32, R1) next to the two U boxes (lower half R0).
#include "gn_common.cuh"
#include "gn_tma.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int NT_TILE = 128, NT_D = 128, NT_F = 64, NT_R = 32;
constexpr int NT_EPI_THREADS = 256, NT_THREADS = NT_EPI_THREADS + 32;
constexpr uint32_t NT_SBO = 128;
constexpr uint32_t NT_LBO_A = NT_TILE * 16 + 32;          // skewed chunk pitch of the A operands
// weight operand image of gn_prepare_operands (gn_det_tc.cu): fc1^T, fc2^T, rd^T as hi | lo,
// then [W_pw_fc1[32:64] | W_pw_fc1[64:96]]^T in 128-row chunks, of which rows 0-63 are used
constexpr uint32_t NT_W1H = 0, NT_W1L = 8192, NT_W2H = 16384, NT_W2L = 32768;
constexpr uint32_t NT_WRH = 49152, NT_WRL = 57344, NT_W_MAIN = 65536;
constexpr uint32_t NT_IMG_WUH = 65536, NT_IMG_WUL = 65536 + 8192, NT_IMG_WU_PITCH = 2048;
constexpr uint32_t NT_WUH = 65536, NT_WUL = 69632, NT_WU_PITCH = 1024;     // compact copy: 64-row chunks
constexpr uint32_t NT_OFF_A = 73728;
constexpr uint32_t NT_A_BYTES = 2 * 16 * NT_LBO_A;        // K = 128, hi + lo
constexpr uint32_t NT_R1 = 32768;                          // upper half of the operand region
constexpr uint32_t NT_OFF_BIAS = NT_OFF_A + NT_A_BYTES;   // b_fc1[64] b_fc2[128] b_rd[32] b_u[64]
constexpr uint32_t NT_OFF_BAR = NT_OFF_BIAS + (NT_F + NT_D + NT_R + NT_F) * 4;
constexpr uint32_t NT_OFF_FEATS = (NT_OFF_BAR + 64 + 1023) / 1024 * 1024;   // 4 boxes x 16 KB
constexpr uint32_t NT_BOX = 16384;
constexpr uint32_t NT_OFF_RED = NT_OFF_FEATS + 4 * NT_BOX;                  // 1 box
constexpr uint32_t NT_SMEM = NT_OFF_RED + NT_BOX + 1024;                    // + alignment slack
static_assert(NT_OFF_A % 1024 == 0 && NT_OFF_FEATS % 1024 == 0, "TMA boxes need 1024-byte alignment");
static_assert(NT_R1 + 2 * 8 * NT_LBO_A <= NT_A_BYTES, "K = 64 operands do not fit the upper half");
static_assert(NT_SMEM <= 227 * 1024, "det TMA tile exceeds shared memory");

#ifdef NT_TRACE
__device__ long long nt_trace[32];
#define NT_TR(i) do { if (blockIdx.x == 0 && t == 0 && it == 2) nt_trace[i] = clock64(); } while (0)
#else
#define NT_TR(i) do { } while (0)
#endif

// issued by a whole converged warp: elect.sync picks the lane (gn_umma.cuh)
template <int KSTEPS, bool X3>
__device__ __forceinline__ void nt_gemm(uint32_t tmem_d, uint64_t d_ah, uint64_t d_al, uint64_t d_bh,
                                        uint64_t d_bl, uint32_t lbo_b, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    if (X3)
      umma::mma_bf16x3_elect(tmem_d, d_ah, d_al, d_bh, d_bl, ks * (2 * NT_LBO_A >> 4),
                             ks * (2 * lbo_b >> 4), idesc, ks > 0);
    else
      umma::mma_bf16_ss_elect(tmem_d, d_ah + ks * (2 * NT_LBO_A >> 4), d_bh + ks * (2 * lbo_b >> 4),
                              idesc, ks > 0);
  }
}

__device__ __forceinline__ void nt_epi_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT_EPI_THREADS) : "memory");
}

__device__ __forceinline__ void nt_split8(const float* x, uint4& h, uint4& l) {
  umma::split_bf16x2(x[0], x[1], h.x, l.x);
  umma::split_bf16x2(x[2], x[3], h.y, l.y);
  umma::split_bf16x2(x[4], x[5], h.z, l.z);
  umma::split_bf16x2(x[6], x[7], h.w, l.w);
}

template <bool X3>
__global__ void __launch_bounds__(NT_THREADS, 1)
block_det_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                     const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_red,
                     float* pooled, const unsigned char* __restrict__ wimg,
                     const float* __restrict__ b_fc1, const float* __restrict__ b_fc2,
                     const float* __restrict__ b_rd, const float* __restrict__ b_u, int num_dets,
                     int has_a, int has_b, int tile_rows) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // A tile covers tile_rows <= 128 detections (the UMMAs always run M = 128; rows beyond
  // tile_rows are never loaded or stored): the launcher picks it so that every SM gets the same
  // number of tiles - 500 tiles of 128 on 148 SMs are four rounds with the last one a quarter full.
  const int num_tiles = (num_dets + tile_rows - 1) / tile_rows;
  if ((int)blockIdx.x >= num_tiles) return;
  const int my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  // stage_a = false (block 1): feats_in goes straight into reduce_dim, nothing is stored back
  const bool stage_a = has_a != 0, stage_b = has_b != 0;

  const uint32_t sbase = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (sbase - umma::smem_u32(smem_raw));
  unsigned char* a_reg = smem + NT_OFF_A;
  float* bias1 = reinterpret_cast<float*>(smem + NT_OFF_BIAS);
  float* bias2 = bias1 + NT_F;
  float* biasr = bias2 + NT_D;
  float* biasu = biasr + NT_R;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NT_OFF_BAR);
  uint64_t* bar_mma = bars;          // tcgen05.commit of the GEMM in flight
  uint64_t* wbar = bars + 1;         // weight image landed
  uint64_t* sc_full = bars + 2;      // shortcut tile landed in the feats boxes
  uint64_t* feats_ready = bars + 3;  // block output complete in the feats boxes
  uint64_t* u_ready = bars + 4;      // U boxes + red box complete
  uint64_t* u_free = bars + 5;       // their stores have read shared memory
  unsigned char* feats = smem + NT_OFF_FEATS;
  unsigned char* redbox = smem + NT_OFF_RED;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (t == 0) {
    for (int i = 0; i < 6; ++i) umma::mbar_init(&bars[i], 1);
    umma::fence_barrier_init();
    umma::mbar_expect_tx(wbar, NT_W_MAIN + 2 * 4 * NT_WU_PITCH);
    umma::bulk_copy_g2s(sbase, wimg, NT_W_MAIN / 2, wbar);
    umma::bulk_copy_g2s(sbase + NT_W_MAIN / 2, wimg + NT_W_MAIN / 2, NT_W_MAIN / 2, wbar);
    for (int j = 0; j < 4; ++j) {
      umma::bulk_copy_g2s(sbase + NT_WUH + j * NT_WU_PITCH, wimg + NT_IMG_WUH + j * NT_IMG_WU_PITCH,
                          NT_WU_PITCH, wbar);
      umma::bulk_copy_g2s(sbase + NT_WUL + j * NT_WU_PITCH, wimg + NT_IMG_WUL + j * NT_IMG_WU_PITCH,
                          NT_WU_PITCH, wbar);
    }
  }
  if (stage_a && t < NT_F) bias1[t] = __ldg(b_fc1 + t);
  if (stage_a && t < NT_D) bias2[t] = __ldg(b_fc2 + t);
  if (stage_b && t < NT_R) biasr[t] = __ldg(b_rd + t);
  if (stage_b && t < NT_F) biasu[t] = __ldg(b_u + t);
  umma::griddep_launch_dependents();   // the next kernel's CTAs may be placed as ours retire
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // up to here only parameters and the weight image were read: pooled (pair kernel in front of
  // us) and feats_in are touched below
  umma::griddep_wait();

  if (warp == NT_EPI_THREADS / 32) {
    // ================================ copy-engine warp ====================================
    if (lane == 0) {
      umma::tma_prefetch_desc(&tm_in);
      umma::tma_prefetch_desc(&tm_out);
      if (stage_b) {
        umma::tma_prefetch_desc(&tm_u);
        umma::tma_prefetch_desc(&tm_red);
      }
      const uint32_t s_feats = sbase + NT_OFF_FEATS;
      auto load_shortcut = [&](int row0) {
        umma::mbar_expect_tx(sc_full, 4u * (uint32_t)tile_rows * 128u);   // 4 boxes of tile_rows x 128 B
        for (int b = 0; b < 4; ++b)
          umma::tma_load_2d(s_feats + b * NT_BOX, &tm_in, b * 32, row0, sc_full, umma::TMA_EVICT_FIRST);
      };
      load_shortcut(blockIdx.x * tile_rows);
      for (int it = 0; it < my_tiles; ++it) {
        const int row0 = (blockIdx.x + it * gridDim.x) * tile_rows;
        umma::mbar_wait_relaxed(feats_ready, (uint32_t)it & 1u);
        if (stage_a) {
          for (int b = 0; b < 4; ++b) umma::tma_store_2d(&tm_out, b * 32, row0, s_feats + b * NT_BOX);
          umma::bulk_commit_group();
          umma::bulk_wait_group_read0();        // the boxes may be overwritten
        }
        if (it + 1 < my_tiles) load_shortcut(row0 + (int)gridDim.x * tile_rows);
        if (stage_b) {
          umma::mbar_wait_relaxed(u_ready, (uint32_t)it & 1u);
          umma::tma_store_2d(&tm_u, 0, row0, sbase + NT_OFF_A);
          umma::tma_store_2d(&tm_u, 32, row0, sbase + NT_OFF_A + NT_BOX);
          umma::tma_store_2d(&tm_red, 0, row0, sbase + NT_OFF_RED);
          umma::bulk_commit_group();
          umma::bulk_wait_group_read0();
          umma::mbar_arrive(u_free);
        }
      }
      umma::bulk_wait_group0();                 // every store complete before the CTA retires
    }
  } else {
    // ================================ epilogue warps ======================================
    const uint32_t tm1 = tmem, tm2 = tmem + 64, tmr = tmem + 192;
    const uint32_t s_a = sbase + NT_OFF_A;
    // K = 64 operands (pooled, d1) and the K = 32 red operand live in the upper half R1
    const uint64_t d_a64h = umma::smem_desc(s_a + NT_R1, NT_LBO_A, NT_SBO);
    const uint64_t d_a64l = umma::smem_desc(s_a + NT_R1 + 8 * NT_LBO_A, NT_LBO_A, NT_SBO);
    const uint64_t d_a32h = d_a64h;
    const uint64_t d_a32l = umma::smem_desc(s_a + NT_R1 + 4 * NT_LBO_A, NT_LBO_A, NT_SBO);
    const uint64_t d_a128h = umma::smem_desc(s_a, NT_LBO_A, NT_SBO);
    const uint64_t d_a128l = umma::smem_desc(s_a + 16 * NT_LBO_A, NT_LBO_A, NT_SBO);
    const uint64_t d_w1h = umma::smem_desc(sbase + NT_W1H, NT_F * 16, NT_SBO), d_w1l = umma::smem_desc(sbase + NT_W1L, NT_F * 16, NT_SBO);
    const uint64_t d_w2h = umma::smem_desc(sbase + NT_W2H, NT_D * 16, NT_SBO), d_w2l = umma::smem_desc(sbase + NT_W2L, NT_D * 16, NT_SBO);
    const uint64_t d_wrh = umma::smem_desc(sbase + NT_WRH, NT_R * 16, NT_SBO), d_wrl = umma::smem_desc(sbase + NT_WRL, NT_R * 16, NT_SBO);
    const uint64_t d_wuh = umma::smem_desc(sbase + NT_WUH, NT_WU_PITCH, NT_SBO), d_wul = umma::smem_desc(sbase + NT_WUL, NT_WU_PITCH, NT_SBO);
    unsigned char* a64h = a_reg + NT_R1;
    unsigned char* a64l = a64h + 8 * NT_LBO_A;
    unsigned char* a32l = a64h + 4 * NT_LBO_A;
    unsigned char* a128h = a_reg;
    unsigned char* a128l = a_reg + 16 * NT_LBO_A;
    const int erow = (warp & 3) * 32 + lane;
    const int ehalf = warp >> 2;
    const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t sw = (uint32_t)(erow & 7);
    const uint32_t rowoff = (uint32_t)erow * 128;     // this thread's row inside a swizzled box
    uint32_t par = 0;
    bool weights_pending = true;

    // pooled rows of a tile: thread -> (row r, 8-float piece q) x 4, whole 256-byte rows per
    // quarter warp.  Plain loads: this kernel also writes pooled.
    float4 pv[4][2];
    auto load_pooled = [&](int row0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (i * 8 + warp) * 4 + (lane >> 3), q = lane & 7;
        pv[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        pv[i][1] = pv[i][0];
        if (r < tile_rows && row0 + r < num_dets) {
          const float4* p = reinterpret_cast<const float4*>(pooled + (size_t)(row0 + r) * NT_F + q * 8);
          pv[i][0] = p[0];
          pv[i][1] = p[1];
        }
      }
    };
    if (stage_a) load_pooled(blockIdx.x * tile_rows);

    for (int it = 0; it < my_tiles; ++it) {
      const int row0 = (blockIdx.x + it * gridDim.x) * tile_rows;
      NT_TR(0);
      if (stage_a) {
      // ---- pooled (prefetched) -> A (K = 64) ----------------------------------------------
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (i * 8 + warp) * 4 + (lane >> 3), q = lane & 7;
        const float x[8] = {pv[i][0].x, pv[i][0].y, pv[i][0].z, pv[i][0].w,
                            pv[i][1].x, pv[i][1].y, pv[i][1].z, pv[i][1].w};
        uint4 h, l;
        nt_split8(x, h, l);
        const uint32_t off = (uint32_t)q * NT_LBO_A + (uint32_t)r * 16;
        *reinterpret_cast<uint4*>(a64h + off) = h;
        *reinterpret_cast<uint4*>(a64l + off) = l;
      }
      NT_TR(11);
      umma::fence_smem_to_async();
      NT_TR(12);
      umma::tc_fence_before();
      nt_epi_sync();
      NT_TR(1);
      if (warp == 0) {
        if (weights_pending) umma::mbar_wait(wbar, 0);
        umma::tc_fence_after();
        nt_gemm<NT_F / 16, X3>(tm1, d_a64h, d_a64l, d_w1h, d_w1l, NT_F * 16, umma::idesc_bf16_f32(NT_TILE, NT_F));
        umma::mma_commit_elect(bar_mma);
      }
      weights_pending = false;
      // pooled <- 0 for the next block.  After the hand-off: fence.proxy.async waits for the
      // thread's outstanding global stores, which put their L2 round trip in front of fc1.
      // Every half warp writes one whole 256-byte row per instruction.
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (i * 8 + warp) * 4 + (lane >> 4) * 2, q = lane & 15;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (r + j < tile_rows && row0 + r + j < num_dets)
            *reinterpret_cast<float4*>(pooled + (size_t)(row0 + r + j) * NT_F + q * 4) =
                make_float4(0.f, 0.f, 0.f, 0.f);
      }
      umma::mbar_wait(bar_mma, par);
      par ^= 1;
      umma::tc_fence_after();
      NT_TR(2);
      // ---- d1 = relu(acc + b_fc1) -> A (K = 64, in place of pooled) -------------------------
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
          nt_split8(x, h, l);
          const uint32_t off = (uint32_t)((col0 >> 3) + g) * NT_LBO_A + (uint32_t)erow * 16;
          *reinterpret_cast<uint4*>(a64h + off) = h;
          *reinterpret_cast<uint4*>(a64l + off) = l;
        }
      }
      umma::fence_smem_to_async();
      umma::tc_fence_before();
      nt_epi_sync();
      NT_TR(3);
      if (warp == 0) {
        umma::tc_fence_after();
        nt_gemm<NT_F / 16, X3>(tm2, d_a64h, d_a64l, d_w2h, d_w2l, NT_D * 16, umma::idesc_bf16_f32(NT_TILE, NT_D));
        umma::mma_commit_elect(bar_mma);
      }
      }   // stage_a
      umma::mbar_wait(sc_full, (uint32_t)it & 1u);                     // shortcut rows are in the boxes
      if (stage_b && it > 0) umma::mbar_wait(u_free, (uint32_t)(it - 1) & 1u);   // U boxes of the last tile are out
      NT_TR(4);
      if (stage_a) {
        umma::mbar_wait(bar_mma, par);
        par ^= 1;
        umma::tc_fence_after();
      }
      NT_TR(5);
      // ---- feats_out = relu(feats_in + acc + b_fc2): in place in the boxes, and -> A (K = 128)
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        const int col0 = ehalf * 64 + cc;
        unsigned char* brow = feats + (uint32_t)(col0 >> 5) * NT_BOX + rowoff;
        const uint32_t ch0 = (uint32_t)(col0 & 31) >> 2;               // first 16-byte chunk: 0 or 4
        float v[16];
        if (stage_a) umma::tmem_ld16(tm2 + tlane + col0, v);
        float4 res[4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          res[g] = *reinterpret_cast<const float4*>(brow + (((ch0 + g) ^ sw) << 4));
        if (stage_a) umma::tmem_ld_wait();
        float x[16];
        if (stage_a) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            x[g * 4 + 0] = fmaxf(res[g].x + (v[g * 4 + 0] + bias2[col0 + g * 4 + 0]), 0.f);
            x[g * 4 + 1] = fmaxf(res[g].y + (v[g * 4 + 1] + bias2[col0 + g * 4 + 1]), 0.f);
            x[g * 4 + 2] = fmaxf(res[g].z + (v[g * 4 + 2] + bias2[col0 + g * 4 + 2]), 0.f);
            x[g * 4 + 3] = fmaxf(res[g].w + (v[g * 4 + 3] + bias2[col0 + g * 4 + 3]), 0.f);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(brow + (((ch0 + g) ^ sw) << 4)) =
                make_float4(x[g * 4 + 0], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            x[g * 4 + 0] = res[g].x; x[g * 4 + 1] = res[g].y; x[g * 4 + 2] = res[g].z; x[g * 4 + 3] = res[g].w;
          }
        }
        if (stage_b) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 h, l;
            nt_split8(x + g * 8, h, l);
            const uint32_t off = (uint32_t)((col0 >> 3) + g) * NT_LBO_A + (uint32_t)erow * 16;
            *reinterpret_cast<uint4*>(a128h + off) = h;
            *reinterpret_cast<uint4*>(a128l + off) = l;
          }
        }
      }
      umma::fence_smem_to_async();
      umma::tc_fence_before();
      nt_epi_sync();
      NT_TR(6);
      if (t == 0) umma::mbar_arrive(feats_ready);          // the copy-engine warp stores the block output
      if (warp == 0 && stage_b) {
        __syncwarp();
        if (weights_pending) umma::mbar_wait(wbar, 0);     // block 1: the first GEMM of the kernel
        umma::tc_fence_after();
        nt_gemm<NT_D / 16, X3>(tmr, d_a128h, d_a128l, d_wrh, d_wrl, NT_R * 16, umma::idesc_bf16_f32(NT_TILE, NT_R));
        umma::mma_commit_elect(bar_mma);
      }
      weights_pending = false;
      if (stage_a && it + 1 < my_tiles) load_pooled(row0 + (int)gridDim.x * tile_rows);
      if (stage_b) {
        umma::mbar_wait(bar_mma, par);
        par ^= 1;
        umma::tc_fence_after();
        NT_TR(7);
        // ---- red = relu(acc + b_rd): 16 columns per thread -> red box (bf16 hi | lo row) and
        //      -> A (K = 32) for the U GEMM ----------------------------------------------------
        {
          const int col0 = ehalf * 16;
          float v[16];
          umma::tmem_ld16(tmr + tlane + col0, v);
          umma::tmem_ld_wait();
          float x[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) x[e] = fmaxf(v[e] + biasr[col0 + e], 0.f);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 h, l;
            nt_split8(x + g * 8, h, l);
            const uint32_t c = (uint32_t)(col0 >> 3) + g;            // hi chunk 0..3, lo chunk 4 + c
            *reinterpret_cast<uint4*>(redbox + rowoff + ((c ^ sw) << 4)) = h;
            *reinterpret_cast<uint4*>(redbox + rowoff + (((4 + c) ^ sw) << 4)) = l;
            const uint32_t off = c * NT_LBO_A + (uint32_t)erow * 16;
            *reinterpret_cast<uint4*>(a64h + off) = h;
            *reinterpret_cast<uint4*>(a32l + off) = l;
          }
        }
        umma::fence_smem_to_async();
        umma::tc_fence_before();
        nt_epi_sync();
        NT_TR(8);
        if (warp == 0) {
          umma::tc_fence_after();
          nt_gemm<NT_R / 16, X3>(tm2, d_a32h, d_a32l, d_wuh, d_wul, NT_WU_PITCH, umma::idesc_bf16_f32(NT_TILE, NT_F));
          umma::mma_commit_elect(bar_mma);
        }
        umma::mbar_wait(bar_mma, par);
        par ^= 1;
        umma::tc_fence_after();
        NT_TR(9);
        // ---- U = acc + b_pw_fc1: 32 columns per thread -> U box `ehalf` (lower half of A) ---------
        {
          const int col0 = ehalf * 32;
          float v[32];
          umma::tmem_ld32(tm2 + tlane + col0, v);
          umma::tmem_ld_wait();
          unsigned char* brow = a_reg + (uint32_t)ehalf * NT_BOX + rowoff;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(brow + (((uint32_t)g ^ sw) << 4)) =
                make_float4(v[g * 4 + 0] + biasu[col0 + g * 4 + 0], v[g * 4 + 1] + biasu[col0 + g * 4 + 1],
                            v[g * 4 + 2] + biasu[col0 + g * 4 + 2], v[g * 4 + 3] + biasu[col0 + g * 4 + 3]);
        }
        umma::fence_smem_to_async();
        umma::tc_fence_before();
        nt_epi_sync();
        NT_TR(10);
        if (t == 0) umma::mbar_arrive(u_ready);
      }
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace gn

static int launch_block_det_tma(const char* name, bool x3, float* pooled, const float* feats_in,
                                const void* wimg, const float* b_fc1, const float* b_fc2,
                                const float* b_rd, int has_a, int has_b, float* feats_out, void* red_hl,
                                const float* b_u, float* u_out, int num_dets, int shortcut_dim,
                                int pairfeat_dim, int reduced_dim, gn_stream_t stream) {
  GN_REQUIRE(num_dets >= 0, "%s: negative size", name);
  if (shortcut_dim != gn::NT_D || pairfeat_dim != gn::NT_F || reduced_dim != gn::NT_R) {
    gn::set_error("%s: fused kernel is built for d=%d f=%d r=%d (got %d, %d, %d)", name, gn::NT_D,
                  gn::NT_F, gn::NT_R, shortcut_dim, pairfeat_dim, reduced_dim);
    return GN_ERR_UNSUPPORTED;
  }
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(feats_in && wimg, "%s: null pointer", name);
  GN_REQUIRE(has_a || has_b, "%s: nothing to do", name);
  GN_REQUIRE(!has_a || (pooled && b_fc1 && b_fc2 && feats_out),
             "%s: stage A needs pooled, fc1 / fc2 biases and feats_out", name);
  GN_REQUIRE(!has_b || (b_rd && red_hl && b_u && u_out),
             "%s: stage B needs reduce_dim / pw_fc1 biases, red_hl and u_out", name);
  GN_REQUIRE((((uintptr_t)pooled | (uintptr_t)feats_in | (uintptr_t)feats_out | (uintptr_t)wimg |
               (uintptr_t)red_hl | (uintptr_t)u_out) & 15) == 0,
             "%s: pointers must be 16-byte aligned", name);
  // rows per tile: the smallest that still gives every SM the same number of tiles
  const int sms = gn::sm_count();
  int tile_rows = gn::NT_TILE;
  {
    const int tiles128 = gn::ceil_div(num_dets, gn::NT_TILE);
    if (tiles128 > sms) {
      const int rounds = gn::ceil_div(tiles128, sms);
      tile_rows = gn::ceil_div(num_dets, rounds * sms);
      if (tile_rows > gn::NT_TILE) tile_rows = gn::NT_TILE;
    }
  }
  CUtensorMap tm_in, tm_out, tm_u, tm_red;
  const uint64_t rows = (uint64_t)num_dets;
  int r = gn::encode_tmap_2d_f32(&tm_in, feats_in, rows, gn::NT_D, gn::NT_D * 4, tile_rows, 32);
  if (has_a) {
    if (r == 0) r = gn::encode_tmap_2d_f32(&tm_out, feats_out, rows, gn::NT_D, gn::NT_D * 4, tile_rows, 32);
  } else {
    tm_out = tm_in;     // not used
  }
  if (has_b) {
    if (r == 0) r = gn::encode_tmap_2d_f32(&tm_u, u_out, rows, gn::NT_F, gn::NT_F * 4, tile_rows, 32);
    if (r == 0) r = gn::encode_tmap_2d_bf16(&tm_red, red_hl, rows, 2 * gn::NT_R, 4 * gn::NT_R, tile_rows, 2 * gn::NT_R);
  } else {
    tm_u = tm_out;
    tm_red = tm_out;
  }
  if (r != 0) {
    gn::set_error("%s: cuTensorMapEncodeTiled failed (%d)", name, r);
    return GN_ERR_CUDA;
  }
  const void* kern = x3 ? (const void*)gn::block_det_tma_kernel<true>
                        : (const void*)gn::block_det_tma_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)gn::NT_SMEM);
  if (e != cudaSuccess) {
    gn::set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(num_dets, tile_rows);
  if (grid > sms) grid = sms;
  if (x3)
    e = gn::launch_kernel(gn::block_det_tma_kernel<true>, grid, gn::NT_THREADS, gn::NT_SMEM,
                          (cudaStream_t)stream, gn::pdl_enabled(), tm_in, tm_out, tm_u, tm_red, pooled,
                          static_cast<const unsigned char*>(wimg), b_fc1, b_fc2, b_rd, b_u, num_dets,
                          has_a, has_b, tile_rows);
  else
    e = gn::launch_kernel(gn::block_det_tma_kernel<false>, grid, gn::NT_THREADS, gn::NT_SMEM,
                          (cudaStream_t)stream, gn::pdl_enabled(), tm_in, tm_out, tm_u, tm_red, pooled,
                          static_cast<const unsigned char*>(wimg), b_fc1, b_fc2, b_rd, b_u, num_dets,
                          has_a, has_b, tile_rows);
  if (e != cudaSuccess) {
    gn::set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  GN_CHECK_LAUNCH(name);
  return GN_OK;
}

// gn_block_det_fwd_img_u on the copy-engine kernel (stage A when pooled is given, stage B when
// has_stage_b; pooled == NULL: block 1, feats_in straight into reduce_dim).
// feats_in and feats_out may not alias partially (equal or disjoint); u_out is [T, 64] fp32,
// red_hl [>= T, 64] bf16.  plain_bf16 != 0: the bf16 arithmetic.
extern "C" int gn_block_det_fwd_tma(float* pooled, const float* feats_in, const void* wimg,
                                    const float* b_fc1, const float* b_fc2, const float* b_rd,
                                    int has_stage_b, float* feats_out, void* red_hl,
                                    const float* b_u, float* u_out, int plain_bf16, int num_dets,
                                    int shortcut_dim, int pairfeat_dim, int reduced_dim,
                                    gn_stream_t stream) {
  return launch_block_det_tma("gn_block_det_fwd_tma", plain_bf16 == 0, pooled, feats_in, wimg, b_fc1,
                              b_fc2, b_rd, pooled != nullptr, has_stage_b, feats_out, red_hl, b_u, u_out, num_dets,
                              shortcut_dim, pairfeat_dim, reduced_dim, stream);
}

#ifdef NT_TRACE
extern "C" int gn_block_det_tma_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::nt_trace, sizeof(long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif
