// One Gnet block's pair stage (A7a + pw_fc1 + pw_fc2 + A7b, network.py:367-388) with the
// operand tile fed without any register staging and every pipeline stage on its own warps.
//
// Formulation.  pw_fc1 acts on the concatenation [pw_feats | c_feats | n_feats]
// (network.py:376), so its pre-activation splits into
//     pw[p] . W1[0:32]  +  red[c] . W1[32:64] + b1  +  (c != n) red[n] . W1[64:96].
// The middle term depends on the detection only: the detection-level kernel
// (gn_det_tc.cu, "AB" output) evaluates U[c] = red[c] . W1[32:64] + b1 once per detection
// and the first epilogue adds it (one broadcast row per run of equal c).  What is left for
// the tensor core per pair is K = 64: the pair's own pw row and the neighbor's reduced
// feature row, both stored as bf16 (hi | lo) rows of 128 bytes = one SWIZZLE_128B row:
//   * pw rows are contiguous in pair order  -> ONE tensor-map TMA tile load per 128 pairs
//     (cp.async.bulk.tensor, UTMALDG), written by the pair-feature kernel in this format;
//   * neighbor rows are a gather by pair_n  -> cp.async (LDGSTS) 16-byte copies straight
//     into the swizzled tile, one warp, two tiles in flight.  (TMA gather4 does the same
//     job functionally - gn_selftest_tma pins it - but costs ~120 cycles per instruction:
//     32 of them per tile made the producer the bottleneck, profiles/r2_pair_tma.md.)
//     The self pair (n == c, zeroed in the reference, network.py:372-374) gathers an
//     all-zero row kept at index `zero_row` of the table.
// bf16x3 with the two hi-operand products stacked along N: one N = 128 UMMA computes
// x_hi . [W_hi | W_lo] into 128 accumulator columns (tensor floor 64 cycles = its 8 KB of
// operand reads), one N = 64 UMMA adds x_lo . W_hi; the epilogue sums the column halves.
//
// Roles (one persistent CTA per SM; every hand-off is an mbarrier):
//   warps  0-7  epilogue 1 : D1[:, j] + D1[:, 64+j] + U[c] -> relu -> bf16 hi/lo -> h1 (TMEM)
//   warps  8-11 epilogue 2 : D2 -> fp32 staging tile (raw accumulators)
//   warps 12-19 pool       : segmented max over runs of equal c, + b2, relu, atomicMax
//   warps 20-22 producers  : one per A-ring stage: tile load of the pw rows + neighbor-row gather
//   warp  23    FC1 issuer : one thread: FC1(it) -> D1[it%2] (K = 64, SS form)
//   warp  24    FC2 issuer : one thread: FC2(it) -> D2[it%2] (A operand h1 in TMEM)
// max / relu / + b2 commute as  max_rows relu(d + b2) = relu(max_rows d + b2), so the staging
// tile holds raw accumulators and the bias is applied once per run.
#include "gn_common.cuh"
#include "gn_tma.cuh"
#include "gn_umma.cuh"

#include <float.h>

namespace gn {

constexpr int BT_TILE = 128, BT_F = 64;
constexpr int BT_STAGES = 3;
constexpr int BT_EPI1_WARPS = 8, BT_EPI2_WARPS = 4, BT_POOL_WARPS = 8;
constexpr int BT_WARP_EPI2 = BT_EPI1_WARPS;                    // 8  (warp % 4 = TMEM lane quadrant)
constexpr int BT_WARP_POOL = BT_WARP_EPI2 + BT_EPI2_WARPS;     // 12
#define BT_LOAD_WARPS BT_STAGES   /* one producer warp per ring stage */
constexpr int BT_WARP_LOAD = BT_WARP_POOL + BT_POOL_WARPS;     // 20: producer warps
constexpr int BT_WARP_MMA = BT_WARP_LOAD + BT_LOAD_WARPS;      // FC1 issuer, then the FC2 issuer
constexpr int BT_THREADS = (BT_WARP_MMA + 2) * 32;             // 736
constexpr uint32_t BT_ATOM = BT_TILE * 128;            // 128 rows x 128 bytes (SWIZZLE_128B): 16 KB
constexpr uint32_t BT_A_BYTES = 2 * BT_ATOM;           // pw atom | neighbor atom
constexpr uint32_t BT_WATOM = BT_F * 128;              // 64-row weight atom: 8 KB
constexpr uint32_t BT_W_BYTES = 4 * BT_WATOM;          // W1 (128 rows: hi | lo) | W2 hi | W2 lo
constexpr int BT_LDS = BT_F + 4;                       // staging row pitch (floats)
constexpr uint32_t BT_STG_BYTES = BT_TILE * BT_LDS * 4;
constexpr uint32_t BT_OFF_W = 0;
constexpr uint32_t BT_OFF_A = BT_OFF_W + BT_W_BYTES;
constexpr uint32_t BT_OFF_STG = BT_OFF_A + BT_STAGES * BT_A_BYTES;
constexpr uint32_t BT_OFF_BAR = BT_OFF_STG + 2 * BT_STG_BYTES;
constexpr int BT_NBAR = 2 * BT_STAGES + 13;
constexpr uint32_t BT_SMEM = BT_OFF_BAR + BT_NBAR * 8 + 1024;   // + slack for the 1024-byte alignment
static_assert(BT_SMEM <= 227 * 1024, "pair TMA pipeline exceeds shared memory");
// BT_STACK = 1: the two hi-operand products of FC1 stacked along N (one N = 128 UMMA into
// 128 accumulator columns + one N = 64 UMMA; 25 % fewer tensor cycles and operand reads, but
// twice the D1 drain through the TMEM read port).  Measured slower (profiles/r2_pair_tma.md):
// the TMEM read port (64 B/clk: accumulator drains + the TS-form A operand of FC2) is the
// resource this kernel saturates first.  Default 0: three N = 64 UMMAs per k-step.
#ifndef BT_STACK
#define BT_STACK 0
#endif
// TMEM columns: D1[2] (128 apart; 64 used unstacked) | h1[2] (hi 32 + lo 32) | D2[2] (64 each)
constexpr uint32_t BT_TM_D1 = 0, BT_TM_H1 = 256, BT_TM_D2 = 384;

#ifdef BT_TRACE
__device__ long long bt_trace[64];
#define BT_TR(i) do { if (blockIdx.x == 0 && (it == 6 || it == 7)) bt_trace[(i) + 32 * (it & 1)] = clock64(); } while (0)
__device__ long long bt_acc[32];
#define BT_ACC_BEGIN() const long long _acc0 = clock64()
#define BT_ACC(i, stmt) do { const long long _t0 = clock64(); stmt; if (blockIdx.x == 0) _acc[(i)] += clock64() - _t0; } while (0)
#define BT_ACC_DECL() long long _acc[4] = {0, 0, 0, 0}
#define BT_ACC_FLUSH(base) do { if (blockIdx.x == 0) { for (int _i = 0; _i < 4; ++_i) bt_acc[(base) + _i] = _acc[_i]; } } while (0)
#else
#define BT_TR(i) do { } while (0)
#define BT_ACC(i, stmt) do { stmt; } while (0)
#define BT_ACC_DECL() do { } while (0)
#define BT_ACC_FLUSH(base) do { } while (0)
#endif

template <bool X3>
__global__ void __launch_bounds__(BT_THREADS, 1)
block_pair_tma_kernel(const __grid_constant__ CUtensorMap tm_pw,
                      const unsigned char* __restrict__ red_hl, const float* __restrict__ u,
                      int u_pitch, const int32_t* __restrict__ pair_c,
                      const int32_t* __restrict__ pair_n, const int32_t* __restrict__ num_pairs,
                      int capacity, int zero_row, const float* __restrict__ b2,
                      const unsigned char* __restrict__ wimg, float* __restrict__ pooled) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + BT_TILE - 1) / BT_TILE;
  if ((int)blockIdx.x >= num_tiles) return;   // uniform per CTA: before any allocation
  const int my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  const uint32_t sbase = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (sbase - umma::smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BT_OFF_BAR);
  uint64_t* a_full = bars;                       // [STAGES] tile load (expect_tx) + gather warp
  uint64_t* a_empty = a_full + BT_STAGES;        // [STAGES] tcgen05.commit
  uint64_t* fc1_done = a_empty + BT_STAGES;      // [2] tcgen05.commit
  uint64_t* h1_full = fc1_done + 2;              // [2] 8 epilogue-1 warps
  uint64_t* fc2_done = h1_full + 2;              // [2] tcgen05.commit
  uint64_t* d2_free = fc2_done + 2;              // [2] 4 epilogue-2 warps
  uint64_t* stg_full = d2_free + 2;              // [2] 128 epilogue-2 threads
  uint64_t* stg_free = stg_full + 2;             // [2] 256 pool threads
  uint64_t* wbar = stg_free + 2;                 // weight image landed

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (t == 32) {
    for (int s = 0; s < BT_STAGES; ++s) {
      umma::mbar_init(&a_full[s], 2);   // the tile load's expect_tx arrival + the stage's producer warp
      umma::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&fc1_done[s], 1);
      umma::mbar_init(&h1_full[s], BT_EPI1_WARPS);
      umma::mbar_init(&fc2_done[s], 1);
      umma::mbar_init(&d2_free[s], BT_EPI2_WARPS);
      umma::mbar_init(&stg_full[s], BT_EPI2_WARPS * 32);   // staging tile: every thread releases /
      umma::mbar_init(&stg_free[s], BT_POOL_WARPS * 32);   // acquires its own rows (racecheck clean)
    }
    umma::mbar_init(wbar, 1);
    umma::fence_barrier_init();
    // operand image prepared by gn_prepare_pair_tma_image (already swizzled)
    umma::mbar_expect_tx(wbar, BT_W_BYTES);
    umma::bulk_copy_g2s(sbase + BT_OFF_W, wimg, BT_W_BYTES, wbar);
  }
  if (t == BT_WARP_LOAD * 32) umma::tma_prefetch_desc(&tm_pw);
  umma::griddep_launch_dependents();   // the next kernel's CTAs may be placed as ours retire
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // everything above touched only this launch's own constants (pair count, weight image);
  // red_hl, u and pooled come from the det kernel in front of us
  umma::griddep_wait();

  if (warp >= BT_WARP_LOAD && warp < BT_WARP_MMA) {
    // ================================== producer ======================================
    // pw atom: one tensor-map tile load.  Neighbor atom: 32 cp.async requests, each 4 rows x
    // 128 bytes (lane = row-in-group * 8 + 16-byte chunk) into the swizzled position.
    // Producer warp w owns ring stage w (tiles w, w + STAGES, ...): wait until FC1 has
    // released the stage, issue the tile's loads, wait for its own cp.async copies, make them
    // visible to the tensor core's proxy and signal.  The other stages' warps keep their
    // tiles in flight meanwhile, so the gather latency (an L2 round trip per tile) is
    // overlapped STAGES-fold; a single warp waiting per tile was the bottleneck.
    const int rl = lane >> 3, ch = lane & 7;
    const int s = warp - BT_WARP_LOAD;
    // lane l holds the gather row index of rows l, l+32, l+64, l+96 of a tile; fetched one
    // tile ahead so the pair-list latency stays off the issue loop
    int idx[4], idx_next[4];
    auto load_idx = [&](int it_, int (&out)[4]) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int p = (blockIdx.x + it_ * gridDim.x) * BT_TILE + k * 32 + lane;
        out[k] = zero_row;                       // rows past P and self pairs read zeros
        if (it_ < my_tiles && p < P) {
          const int c = __ldg(pair_c + p), nn = __ldg(pair_n + p);
          if (nn != c) out[k] = nn;
        }
      }
    };
    load_idx(s, idx_next);
    BT_ACC_DECL();
    const uint32_t sa = sbase + BT_OFF_A + (uint32_t)s * BT_A_BYTES;
#pragma unroll 1
    for (int it = s; it < my_tiles; it += BT_STAGES) {
      const uint32_t n = (uint32_t)(it / BT_STAGES);
      const int tile = blockIdx.x + it * gridDim.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) idx[k] = idx_next[k];
      load_idx(it + BT_STAGES, idx_next);
      if (n >= 1) BT_ACC(0, umma::mbar_wait_relaxed(&a_empty[s], (n - 1) & 1u));
#ifndef BT_NOTILE
      if (lane == 0) {
        umma::mbar_expect_tx(&a_full[s], BT_ATOM);
        umma::tma_load_2d(sa, &tm_pw, 0, tile * BT_TILE, &a_full[s], umma::TMA_EVICT_FIRST);
      }
#else
      if (lane == 0) umma::mbar_arrive(&a_full[s]);
#endif
#ifndef BT_NOGATHER
#pragma unroll
      for (int g = 0; g < 32; ++g) {
        const int row = 4 * g + rl;
        const int src_row = __shfl_sync(0xffffffffu, idx[g >> 3], row & 31);
        const uint32_t dst = sa + BT_ATOM + (uint32_t)row * 128 + (uint32_t)((ch ^ (row & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                     ::"r"(dst), "l"(red_hl + (size_t)src_row * 128 + ch * 16) : "memory");
      }
      cp_async_commit();
      BT_ACC(1, cp_async_wait<0>());
#endif
      umma::fence_smem_to_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&a_full[s]);
    }
    if (lane == 0 && s == 0) BT_ACC_FLUSH(0);
  } else if (warp == BT_WARP_MMA) {
    // ============================== FC1 issuer ==========================================
    // Two issuing threads (this one and the FC2 issuer below) keep the in-order tensor pipe
    // fed: FC1 is bound by shared-memory operand reads, FC2 by the TMEM read of its A
    // operand, and a single issuer would idle the pipe across its own mbarrier waits.
    {   // whole warp, warp-uniform operands; elect.sync picks the issuing lane (gn_umma.cuh)
      const uint32_t idesc64 = umma::idesc_bf16_f32(BT_TILE, 64);
      const uint32_t idesc128 = umma::idesc_bf16_f32(BT_TILE, 128);
      const uint64_t d_w1 = umma::smem_desc_sw128(sbase + BT_OFF_W);
      // A row (both atoms): [hi k 0-15 | hi k 16-31 | lo k 0-15 | lo k 16-31], 32 bytes each.
      // W1 image row n < 64: [W1pw_hi ks0 | ks1 | W1n_hi ks0 | ks1], row 64 + n the lo parts:
      // slab q of the image pairs with the hi slab (q & 1) of A atom (q >> 1).
      umma::mbar_wait(wbar, 0);
      BT_ACC_DECL();
#pragma unroll 1
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % BT_STAGES, b = it & 1;
        const uint32_t n2 = (uint32_t)it >> 1;
        BT_ACC(0, umma::mbar_wait(&a_full[s], (uint32_t)(it / BT_STAGES) & 1u));
        // D1[b] is free once epilogue 1 of tile it-2 has read it (it signals h1_full then)
        if (n2 >= 1) BT_ACC(1, umma::mbar_wait(&h1_full[b], (n2 - 1) & 1u));
        umma::tc_fence_after();
        BT_TR(2);
#ifdef BT_TRACE
        if (blockIdx.x == 0 && (it == 6 || it == 106)) {   // SM clock over 100 tiles of this CTA
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          bt_trace[it == 6 ? 20 : 22] = clock64();
          bt_trace[it == 6 ? 21 : 23] = (long long)gt;
        }
#endif
        const uint64_t d_a0 = umma::smem_desc_sw128(sbase + BT_OFF_A + (uint32_t)s * BT_A_BYTES);
        const uint32_t d1 = tmem + BT_TM_D1 + (uint32_t)b * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint64_t da = d_a0 + (q >> 1) * (BT_ATOM >> 4);
          const int j = q & 1;
          if (X3 && BT_STACK) {
            // x_hi . [W_hi | W_lo] -> columns [0,64) | [64,128);  x_lo . W_hi -> [0,64)
            umma::mma_bf16_ss_elect(d1, da + j * 2, d_w1 + q * 2, idesc128, q > 0);
            umma::mma_bf16_ss_elect(d1, da + (2 + j) * 2, d_w1 + q * 2, idesc64, 1);
          } else if (X3) {
            umma::mma_bf16_ss_elect(d1, da + (2 + j) * 2, d_w1 + q * 2, idesc64, q > 0);            // lo . hi
            umma::mma_bf16_ss_elect(d1, da + j * 2, d_w1 + (BT_WATOM >> 4) + q * 2, idesc64, 1);     // hi . lo
            umma::mma_bf16_ss_elect(d1, da + j * 2, d_w1 + q * 2, idesc64, 1);                       // hi . hi
          } else {
            umma::mma_bf16_ss_elect(d1, da + j * 2, d_w1 + q * 2, idesc64, q > 0);
          }
        }
        umma::mma_commit_elect(&fc1_done[b]);
        umma::mma_commit_elect(&a_empty[s]);
        BT_TR(3);
      }
      BT_ACC_FLUSH(4);
    }
  } else if (warp == BT_WARP_MMA + 1) {
    // ============================== FC2 issuer ==========================================
    {   // whole warp, warp-uniform operands; elect.sync picks the issuing lane (gn_umma.cuh)
      const uint32_t idesc64 = umma::idesc_bf16_f32(BT_TILE, 64);
      const uint64_t d_w2h = umma::smem_desc_sw128(sbase + BT_OFF_W + 2 * BT_WATOM);
      const uint64_t d_w2l = umma::smem_desc_sw128(sbase + BT_OFF_W + 3 * BT_WATOM);
      umma::mbar_wait(wbar, 0);
      BT_ACC_DECL();
#pragma unroll 1
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        const uint32_t n2 = (uint32_t)it >> 1;
        BT_ACC(0, umma::mbar_wait(&h1_full[b], n2 & 1u));
        if (n2 >= 1) BT_ACC(1, umma::mbar_wait(&d2_free[b], (n2 - 1) & 1u));
        umma::tc_fence_after();
        BT_TR(0);
        const uint32_t d2 = tmem + BT_TM_D2 + (uint32_t)b * 64;
        const uint32_t hh = tmem + BT_TM_H1 + (uint32_t)b * 64, hl = hh + 32;
#pragma unroll
        for (int ks = 0; ks < BT_F / 16; ++ks) {
          if (X3) {
            umma::mma_bf16_ts_elect(d2, hl + ks * 8, d_w2h + ks * 2, idesc64, ks > 0);
            umma::mma_bf16_ts_elect(d2, hh + ks * 8, d_w2l + ks * 2, idesc64, 1);
            umma::mma_bf16_ts_elect(d2, hh + ks * 8, d_w2h + ks * 2, idesc64, 1);
          } else {
            umma::mma_bf16_ts_elect(d2, hh + ks * 8, d_w2h + ks * 2, idesc64, ks > 0);
          }
        }
        umma::mma_commit_elect(&fc2_done[b]);
        BT_TR(1);
      }
      BT_ACC_FLUSH(8);
    }
  } else if (warp < BT_WARP_EPI2) {
    // ====================== epilogue 1: D1 + U[c] -> relu -> h1 (TMEM) ======================
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane, col0 = half * 32;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    auto load_c = [&](int it_) {
      int c = 0;                                  // rows past P: any valid row, result unused
      if (it_ < my_tiles) {
        const int p = (blockIdx.x + it_ * gridDim.x) * BT_TILE + row;
        if (p < P) c = __ldg(pair_c + p);
      }
      return c;
    };
    int c = load_c(0);
    BT_ACC_DECL();
#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const uint32_t n2 = (uint32_t)it >> 1;
      // the detection-level half of the pre-activation (rows of one run share c: broadcast)
      const float4* up = reinterpret_cast<const float4*>(u + (size_t)c * u_pitch + col0);
      float4 uv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) uv[k] = __ldg(up + k);
      c = load_c(it + 1);
      BT_ACC(0, umma::mbar_wait_relaxed(&fc1_done[b], n2 & 1u));
      // h1[b] may be overwritten once FC2 of tile it-2 has consumed it (the two issuers are
      // not ordered against each other, so this is an explicit wait)
      if (n2 >= 1) BT_ACC(1, umma::mbar_wait_relaxed(&fc2_done[b], (n2 - 1) & 1u));
      umma::tc_fence_after();
      if (t == 0) BT_TR(8);
      const uint32_t td = tmem + BT_TM_D1 + (uint32_t)b * 128 + tlane + col0;
      const uint32_t th = tmem + BT_TM_H1 + (uint32_t)b * 64 + tlane + (uint32_t)(col0 >> 1);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {          // 16 columns at a time
        float v[16], w[16];
        umma::tmem_ld16(td + hf * 16, v);
        if (X3 && BT_STACK) umma::tmem_ld16(td + 64 + hf * 16, w);
        umma::tmem_ld_wait();
        uint32_t hh[8], hl[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 uu = uv[hf * 4 + k];
          float x0 = v[4 * k] + uu.x, x1 = v[4 * k + 1] + uu.y, x2 = v[4 * k + 2] + uu.z,
                x3 = v[4 * k + 3] + uu.w;
          if (X3 && BT_STACK) { x0 += w[4 * k]; x1 += w[4 * k + 1]; x2 += w[4 * k + 2]; x3 += w[4 * k + 3]; }
          umma::split_bf16x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f), hh[2 * k], hl[2 * k]);
          umma::split_bf16x2(fmaxf(x2, 0.f), fmaxf(x3, 0.f), hh[2 * k + 1], hl[2 * k + 1]);
        }
        umma::tmem_st8(th + hf * 8, hh);
        if (X3) umma::tmem_st8(th + 32 + hf * 8, hl);
      }
      umma::tmem_st_wait();
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&h1_full[b]);
      if (t == 0) BT_TR(9);
    }
    if (t == 0) BT_ACC_FLUSH(12);
  } else if (warp < BT_WARP_POOL) {
    // ========================= epilogue 2: D2 -> fp32 staging tile ==========================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    BT_ACC_DECL();
#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const uint32_t n2 = (uint32_t)it >> 1;
      BT_ACC(0, umma::mbar_wait_relaxed(&fc2_done[b], n2 & 1u));
      umma::tc_fence_after();
      if (t == BT_WARP_EPI2 * 32) BT_TR(10);
      float v0[32], v1[32];
      umma::tmem_ld32(tmem + BT_TM_D2 + (uint32_t)b * 64 + tlane, v0);
      umma::tmem_ld32(tmem + BT_TM_D2 + (uint32_t)b * 64 + tlane + 32, v1);
      umma::tmem_ld_wait();
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&d2_free[b]);
      if (n2 >= 1) BT_ACC(1, umma::mbar_wait_relaxed(&stg_free[b], (n2 - 1) & 1u));
      float4* dst = reinterpret_cast<float4*>(smem + BT_OFF_STG + b * BT_STG_BYTES) + row * (BT_LDS / 4);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dst[k] = make_float4(v0[4 * k], v0[4 * k + 1], v0[4 * k + 2], v0[4 * k + 3]);
        dst[8 + k] = make_float4(v1[4 * k], v1[4 * k + 1], v1[4 * k + 2], v1[4 * k + 3]);
      }
      umma::mbar_arrive(&stg_full[b]);
      if (t == BT_WARP_EPI2 * 32) BT_TR(11);
    }
    if (t == BT_WARP_EPI2 * 32) BT_ACC_FLUSH(16);
  } else {
    // ==================== pool: segmented max over runs of equal c ==========================
    // warp = 16-row slice, lane = column pair; a run ends where the next row has another c,
    // at the end of the slice, or at the last valid pair (rows past P never flush)
    const int slice = warp - BT_WARP_POOL;
    const int j = lane * 2;
    const float bb0 = __ldg(b2 + j), bb1 = __ldg(b2 + j + 1);
    int c_cur = -1, c_nxt = -2;
    auto load_seg = [&](int it_) {
      c_cur = -1;
      c_nxt = -2;
      if (it_ < my_tiles && lane < 16) {
        const int p = (blockIdx.x + it_ * gridDim.x) * BT_TILE + slice * 16 + lane;
        if (p < P) c_cur = __ldg(pair_c + p);
        if (p + 1 < P && lane != 15) c_nxt = __ldg(pair_c + p + 1);
      }
    };
    load_seg(0);
    BT_ACC_DECL();
#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const uint32_t n2 = (uint32_t)it >> 1;
      const unsigned ends = __ballot_sync(0xffffffffu, c_cur >= 0 && c_cur != c_nxt);
      const int cc = c_cur;
      load_seg(it + 1);
      BT_ACC(0, umma::mbar_wait_relaxed(&stg_full[b], n2 & 1u));
      if (lane == 0 && slice == 0) BT_TR(12);
      const float* col = reinterpret_cast<const float*>(smem + BT_OFF_STG + b * BT_STG_BYTES) +
                         slice * 16 * BT_LDS + j;
      float2 x[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) x[r] = *reinterpret_cast<const float2*>(col + r * BT_LDS);
      umma::mbar_arrive(&stg_free[b]);   // this thread's rows live in registers now
      float cur0 = -FLT_MAX, cur1 = -FLT_MAX;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        cur0 = fmaxf(cur0, x[r].x);
        cur1 = fmaxf(cur1, x[r].y);
        if ((ends >> r) & 1u) {        // warp uniform
          const int cr = __shfl_sync(0xffffffffu, cc, r);
          int* dst = reinterpret_cast<int*>(pooled + (size_t)cr * BT_F + j);
          atomicMax(dst, __float_as_int(fmaxf(cur0 + bb0, 0.f)));
          atomicMax(dst + 1, __float_as_int(fmaxf(cur1 + bb1, 0.f)));
          cur0 = cur1 = -FLT_MAX;
        }
      }
      if (lane == 0 && slice == 0) BT_TR(13);
    }
    if (lane == 0 && slice == 0) BT_ACC_FLUSH(20);
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------
// Operand image of one block for the kernel above (32 KB, byte-for-byte the shared-memory
// weight region), SWIZZLE_128B rows of 128 bytes:
//   rows   0-63  (W1 hi)  row n = [W1[0:32]^T hi (k 0..31) | W1[64:96]^T hi (k 0..31)]
//   rows  64-127 (W1 lo)  the lo parts, same positions   (one 128-row B tile: N-stacked hi|lo)
//   rows 128-191 W2^T hi  row n = hi k 0..63        rows 192-255 W2^T lo
// 16-byte chunk j of row r sits at r * 128 + ((j ^ (r % 8)) * 16).
// table: 2 int32 per block = offsets (floats) of pw_fc1/weights [96,64] and pw_fc2/weights
// [64,64] in the flat parameter buffer.
// ---------------------------------------------------------------------------------
__global__ void prepare_pair_tma_image_kernel(const float* __restrict__ flat,
                                              const int32_t* __restrict__ table,
                                              unsigned char* __restrict__ image) {
  const float* w1 = flat + table[blockIdx.x * 2];
  const float* w2 = flat + table[blockIdx.x * 2 + 1];
  unsigned char* img = image + (size_t)blockIdx.x * BT_W_BYTES;
  // unit = (matrix, row n, chunk): 8 k values -> one hi chunk and one lo chunk
  for (int uidx = threadIdx.x; uidx < 2 * BT_F * 8; uidx += blockDim.x) {
    const int mat = uidx / (BT_F * 8), n = (uidx / 8) % BT_F, jj = uidx & 7;
    float x[8];
    uint32_t off_hi, off_lo;
    if (mat == 0) {
      // chunks 0-3: pw rows k = jj*8.. ; chunks 4-7: neighbor rows k = 64 + (jj-4)*8..
      const int k0 = jj < 4 ? jj * 8 : 64 + (jj - 4) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __ldg(w1 + (size_t)(k0 + e) * BT_F + n);
      off_hi = n * 128 + ((jj ^ (n & 7)) * 16);
      off_lo = off_hi + BT_WATOM;                 // row 64 + n: same n % 8
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __ldg(w2 + (size_t)(jj * 8 + e) * BT_F + n);
      off_hi = 2 * BT_WATOM + n * 128 + ((jj ^ (n & 7)) * 16);
      off_lo = off_hi + BT_WATOM;
    }
    uint4 h, l;
    umma::split_bf16x2(x[0], x[1], h.x, l.x);
    umma::split_bf16x2(x[2], x[3], h.y, l.y);
    umma::split_bf16x2(x[4], x[5], h.z, l.z);
    umma::split_bf16x2(x[6], x[7], h.w, l.w);
    *reinterpret_cast<uint4*>(img + off_hi) = h;
    *reinterpret_cast<uint4*>(img + off_lo) = l;
  }
}

}  // namespace gn

#ifdef BT_TRACE
extern "C" int gn_block_pair_tma_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::bt_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
extern "C" int gn_block_pair_tma_acc(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::bt_acc, sizeof(long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int64_t gn_block_pair_tma_image_bytes(void) { return (int64_t)gn::BT_W_BYTES; }

extern "C" int gn_prepare_pair_tma_image(const float* flat_params, const int32_t* table,
                                         int num_blocks, void* image, gn_stream_t stream) {
  GN_REQUIRE(num_blocks >= 0, "gn_prepare_pair_tma_image: negative block count");
  if (num_blocks == 0) return GN_OK;
  GN_REQUIRE(flat_params && table && image, "gn_prepare_pair_tma_image: null pointer");
  GN_REQUIRE(((uintptr_t)image & 15) == 0, "gn_prepare_pair_tma_image: image must be 16-byte aligned");
  gn::prepare_pair_tma_image_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>(
      flat_params, table, static_cast<unsigned char*>(image));
  GN_CHECK_LAUNCH("gn_prepare_pair_tma_image");
  return GN_OK;
}

static int launch_pair_tma(const char* name, bool x3, const void* pw_hl, const void* red_hl,
                           int num_dets, const float* u, int u_pitch, const int32_t* pair_c,
                           const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                           const float* b2, const void* wimg, float* pooled, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0 && num_dets >= 0, "%s: negative size", name);
  if (capacity == 0 || num_dets == 0) return GN_OK;
  GN_REQUIRE(pw_hl && red_hl && u && pair_c && pair_n && num_pairs && b2 && wimg && pooled,
             "%s: null pointer", name);
  GN_REQUIRE(u_pitch >= gn::BT_F && u_pitch % 4 == 0, "%s: u_pitch=%d must be a multiple of 4, >= %d",
             name, u_pitch, gn::BT_F);
  GN_REQUIRE((((uintptr_t)pw_hl | (uintptr_t)red_hl | (uintptr_t)u | (uintptr_t)wimg) & 15) == 0,
             "%s: pointers must be 16-byte aligned", name);
  CUtensorMap tm_pw;
  const int r = gn::encode_tmap_2d_bf16(&tm_pw, pw_hl, (uint64_t)capacity, 64, 128, gn::BT_TILE, 64);
  if (r != 0) {
    gn::set_error("%s: cuTensorMapEncodeTiled failed (%d)", name, r);
    return GN_ERR_CUDA;
  }
  const void* kern = x3 ? (const void*)gn::block_pair_tma_kernel<true>
                        : (const void*)gn::block_pair_tma_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)gn::BT_SMEM);
  if (e != cudaSuccess) {
    gn::set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(capacity, gn::BT_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  if (x3)
    e = gn::launch_kernel(gn::block_pair_tma_kernel<true>, grid, gn::BT_THREADS, gn::BT_SMEM,
                          (cudaStream_t)stream, gn::pdl_enabled(), tm_pw,
                          static_cast<const unsigned char*>(red_hl), u, u_pitch, pair_c, pair_n,
                          num_pairs, capacity, num_dets, b2, static_cast<const unsigned char*>(wimg),
                          pooled);
  else
    e = gn::launch_kernel(gn::block_pair_tma_kernel<false>, grid, gn::BT_THREADS, gn::BT_SMEM,
                          (cudaStream_t)stream, gn::pdl_enabled(), tm_pw,
                          static_cast<const unsigned char*>(red_hl), u, u_pitch, pair_c, pair_n,
                          num_pairs, capacity, num_dets, b2, static_cast<const unsigned char*>(wimg),
                          pooled);
  if (e != cudaSuccess) {
    gn::set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  GN_CHECK_LAUNCH(name);
  return GN_OK;
}

extern "C" int gn_block_pair_fwd_tma(const void* pw_hl, const void* red_hl, int num_dets,
                                     const float* u, int u_pitch, const int32_t* pair_c,
                                     const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                                     const float* b2, const void* wimg, float* pooled,
                                     gn_stream_t stream) {
  return launch_pair_tma("gn_block_pair_fwd_tma", true, pw_hl, red_hl, num_dets, u, u_pitch, pair_c,
                         pair_n, num_pairs, capacity, b2, wimg, pooled, stream);
}

extern "C" int gn_block_pair_fwd_tma_bf16(const void* pw_hl, const void* red_hl, int num_dets,
                                          const float* u, int u_pitch, const int32_t* pair_c,
                                          const int32_t* pair_n, const int32_t* num_pairs,
                                          int capacity, const float* b2, const void* wimg,
                                          float* pooled, gn_stream_t stream) {
  return launch_pair_tma("gn_block_pair_fwd_tma_bf16", false, pw_hl, red_hl, num_dets, u, u_pitch,
                         pair_c, pair_n, num_pairs, capacity, b2, wimg, pooled, stream);
}
