// One Gnet block's pair stage (A7a + pw_fc1 + pw_fc2 + A7b, network.py:367-388) as a
// warp-specialised pipeline: the same arithmetic as gn_block_tc.cu (bf16x3 UMMAs,
// K = 96 -> 64 -> 64, exact integer atomicMax pooling), but instead of two co-resident
// CTAs stepping through their phases, ONE persistent CTA per SM keeps three tiles in
// flight through dedicated warps:
//
//   warps  0-7   fill      : gather the 96 inputs of each pair (pw_feats fp32 + the bf16
//                            hi|lo reduced-feature rows of c and n), split / copy them into
//                            the K-major A operand tile A[it % 2]; loads run one tile ahead
//   warp   8     MMA       : one thread issues FC1(it+1) -> D1[(it+1) % 2], then FC2(it) ->
//                            D2[it % 2] (A operand h1 in tensor memory)
//   warps  9-16  epilogue 0: tiles it = 0, 2, 4, ...   |  epi 1: D1 -> +b1, relu, bf16 hi/lo
//   warps 17-24  epilogue 1: tiles it = 1, 3, 5, ...   |         -> h1 in TMEM
//                                                        epi 2: D2 -> +b2, relu -> fp32 tile
//                                                        pool : segmented max + atomicMax
//
// Hand-offs are mbarriers: a_full / a_empty (fill <-> MMA, tcgen05.commit frees the
// slot), fc1_done / fc2_done (MMA -> epilogue), h1_full (epilogue -> MMA).  Each epilogue
// group owns its TMEM columns (D1 | D2 | h1 = 192), its fp32 staging tile and its named
// barrier, so the two groups never synchronise with each other.
#include "gn_common.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int BP_TILE = 128;
constexpr int BP_W = 32, BP_R = 32, BP_F = 64;
constexpr int BP_K1 = BP_W + 2 * BP_R;          // 96
constexpr int BP_CH1 = BP_K1 / 8, BP_CH2 = BP_F / 8;
constexpr int BP_FILL_WARPS = 8, BP_EPI_WARPS = 8;
constexpr int BP_WARP_MMA = BP_FILL_WARPS;
constexpr int BP_WARP_EPI = BP_WARP_MMA + 1;
constexpr int BP_THREADS = (BP_WARP_EPI + 2 * BP_EPI_WARPS) * 32;   // 800
constexpr uint32_t BP_SBO = 128;
constexpr uint32_t BP_LBO_A = BP_TILE * 16 + 32;    // skewed chunk pitch (see gn_block_tc.cu)
constexpr uint32_t BP_LBO_B = BP_F * 16;
constexpr int BP_LDH2 = BP_F + 4;                   // fp32 staging row pitch (floats)

constexpr uint32_t BP_OFF_B1H = 0;
constexpr uint32_t BP_OFF_B1L = BP_OFF_B1H + BP_CH1 * BP_LBO_B;
constexpr uint32_t BP_OFF_B2H = BP_OFF_B1L + BP_CH1 * BP_LBO_B;
constexpr uint32_t BP_OFF_B2L = BP_OFF_B2H + BP_CH2 * BP_LBO_B;
constexpr uint32_t BP_W_BYTES = BP_OFF_B2L + BP_CH2 * BP_LBO_B;          // 40960 = weight image
constexpr uint32_t BP_A_BYTES = 2 * BP_CH1 * BP_LBO_A;                   // hi + lo: 49920
constexpr uint32_t BP_OFF_A = BP_W_BYTES;                                // 2 slots
constexpr uint32_t BP_OFF_H2 = BP_OFF_A + 2 * BP_A_BYTES;                // 2 groups
constexpr uint32_t BP_H2_BYTES = BP_TILE * BP_LDH2 * 4;                  // 34816
constexpr uint32_t BP_OFF_IDX = BP_OFF_H2 + 2 * BP_H2_BYTES;             // 2 x (c[128], ends[4])
constexpr uint32_t BP_IDX_BYTES = (BP_TILE + 4) * 4;
constexpr uint32_t BP_OFF_BIAS = BP_OFF_IDX + 2 * BP_IDX_BYTES;          // b1[64], b2[64]
constexpr uint32_t BP_OFF_BAR = BP_OFF_BIAS + 2 * BP_F * 4;
constexpr int BP_NBAR = 11;
constexpr uint32_t BP_SMEM = BP_OFF_BAR + BP_NBAR * 8;
static_assert(BP_SMEM <= 227 * 1024, "pair pipeline exceeds shared memory");
static_assert(BP_OFF_BAR % 8 == 0 && BP_OFF_H2 % 16 == 0 && BP_OFF_A % 128 == 0, "alignment");

#ifdef BP_TRACE
__device__ long long bp_trace[64];
#define BP_TR(i) do { if (blockIdx.x == 0 && (it == 4 || it == 5)) bp_trace[(i) + 32 * (it & 1)] = clock64(); } while (0)
#else
#define BP_TR(i) do { } while (0)
#endif

// X3 = true : fp32 semantics, every product as three bf16 UMMAs (hi*hi + hi*lo + lo*hi);
// X3 = false: plain bf16 operands (the hi parts only) with fp32 accumulation - the "bf16"
//             arithmetic BASELINE configs[2] names; a third of the UMMAs, half the fill.
template <bool X3>
__global__ void __launch_bounds__(BP_THREADS, 1)
block_pair_pipe_kernel(const float* __restrict__ pw, const float* __restrict__ feats_hl,
                       const float* __restrict__ nfeats_hl, const int32_t* __restrict__ pair_c,
                       const int32_t* __restrict__ pair_n, const int32_t* __restrict__ num_pairs,
                       int capacity, const float* __restrict__ b1, const float* __restrict__ b2,
                       const unsigned char* __restrict__ wimg, float* __restrict__ pooled) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + BP_TILE - 1) / BP_TILE;
  if ((int)blockIdx.x >= num_tiles) return;   // uniform per CTA: before any allocation
  const int my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  float* bias1 = reinterpret_cast<float*>(smem + BP_OFF_BIAS);
  float* bias2 = bias1 + BP_F;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BP_OFF_BAR);
  uint64_t* a_full = bars;            // [2] count 8 (fill warps)
  uint64_t* a_empty = bars + 2;       // [2] tcgen05.commit
  uint64_t* fc1_done = bars + 4;      // [2]
  uint64_t* h1_full = bars + 6;       // [2] count 8 (epilogue warps of the group)
  uint64_t* fc2_done = bars + 8;      // [2]
  uint64_t* wbar = bars + 10;         // weight image landed

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (t == 0) {
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&a_full[s], BP_FILL_WARPS);
      umma::mbar_init(&a_empty[s], 1);
      umma::mbar_init(&fc1_done[s], 1);
      umma::mbar_init(&h1_full[s], BP_EPI_WARPS);
      umma::mbar_init(&fc2_done[s], 1);
    }
    umma::mbar_init(wbar, 1);
    umma::fence_barrier_init();
    // operand image prepared by gn_prepare_operands: [W1^T hi | lo | W2^T hi | lo]
    umma::mbar_expect_tx(wbar, BP_W_BYTES);
    umma::bulk_copy_g2s(umma::smem_u32(smem), wimg, BP_W_BYTES, wbar);
  }
  if (t < BP_F) {
    bias1[t] = __ldg(b1 + t);
    bias2[t] = __ldg(b2 + t);
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp < BP_FILL_WARPS) {
    // ================================ fill warps ======================================
    // Warp task = 8 rows x one 128-byte part (pw | c | n); lane = (row % 8) * 4 + piece, so
    // every warp request covers 8 rows x 128 contiguous bytes (and a quarter-warp only two
    // rows: the piece-major mapping that would make the stores conflict free without the
    // chunk-pitch skew costs 4x the L1 tag lookups on the loads; measured 339 vs 245 us).  6 units of 8 floats per
    // thread, fetched one tile ahead; indices two tiles ahead.
    float4 pre[6][2];
    int idx_c[4], idx_n[2];
    auto prefetch_idx = [&](int it_) {
      const int q0 = (blockIdx.x + it_ * gridDim.x) * BP_TILE;
#pragma unroll
      for (int u = 2; u < 6; ++u) {
        const int task = u * 8 + warp;
        const int p = q0 + (task & 15) * 8 + (lane >> 2);
        const bool ok = it_ < my_tiles && p < P;
        idx_c[u - 2] = ok ? __ldg(pair_c + p) : -1;
        if (u >= 4) idx_n[u - 4] = ok ? __ldg(pair_n + p) : -1;
      }
    };
    auto prefetch = [&](int it_) {
      const int q0 = (blockIdx.x + it_ * gridDim.x) * BP_TILE;
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int task = u * 8 + warp;
        const int row = (task & 15) * 8 + (lane >> 2);
        const int q = lane & 3;
        const int p = q0 + row;
        const float* src = nullptr;
        if (u < 2) {
          if (it_ < my_tiles && p < P) src = pw + (size_t)p * BP_W + q * 8;
        } else {
          const int c = idx_c[u - 2];
          if (c >= 0) {
            // 128-byte row: hi chunk q at +16q bytes, lo chunk q at +64+16q bytes
            if (u < 4) src = feats_hl + (size_t)c * BP_R + q * 4;
            else {
              const int n = idx_n[u - 4];
              if (n != c) src = nfeats_hl + (size_t)n * BP_R + q * 4;   // self pair: zeros (network.py:372-374)
            }
          }
        }
        pre[u][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        pre[u][1] = pre[u][0];
        if (src != nullptr) {
          pre[u][0] = ldg4(src);
          if (X3 || u < 2) pre[u][1] = ldg4(src + (u < 2 ? 4 : 16));   // (the lo chunk of an hl row)
        }
      }
    };
    prefetch_idx(0);
    prefetch(0);
    prefetch_idx(1);
#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int slot = it & 1;
      const uint32_t n = (uint32_t)it >> 1;
      if (n >= 1) umma::mbar_wait_relaxed(&a_empty[slot], (n - 1) & 1u);
      if (t == 0) BP_TR(20);
      unsigned char* a_hi = smem + BP_OFF_A + slot * BP_A_BYTES;
      unsigned char* a_lo = a_hi + BP_CH1 * BP_LBO_A;
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int task = u * 8 + warp;
        const int part = task >> 4;
        const int row = (task & 15) * 8 + (lane >> 2);
        const int q = lane & 3;
        const float4 v0 = pre[u][0], v1 = pre[u][1];
        uint4 h, l;
        if (u >= 2) {
          h = make_uint4(__float_as_uint(v0.x), __float_as_uint(v0.y), __float_as_uint(v0.z), __float_as_uint(v0.w));
          l = make_uint4(__float_as_uint(v1.x), __float_as_uint(v1.y), __float_as_uint(v1.z), __float_as_uint(v1.w));
        } else {
          umma::split_bf16x2(v0.x, v0.y, h.x, l.x);
          umma::split_bf16x2(v0.z, v0.w, h.y, l.y);
          umma::split_bf16x2(v1.x, v1.y, h.z, l.z);
          umma::split_bf16x2(v1.z, v1.w, h.w, l.w);
        }
        const uint32_t off = (uint32_t)(part * 4 + q) * BP_LBO_A + (uint32_t)row * 16;
        *reinterpret_cast<uint4*>(a_hi + off) = h;
        if (X3) *reinterpret_cast<uint4*>(a_lo + off) = l;
      }
      umma::fence_smem_to_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&a_full[slot]);
      if (t == 0) BP_TR(21);
      prefetch(it + 1);
      prefetch_idx(it + 2);
    }
  } else if (warp == BP_WARP_MMA) {
    // ================================ MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = umma::idesc_bf16_f32(BP_TILE, BP_F);
      const uint32_t sb = umma::smem_u32(smem);
      const uint64_t d_b1h = umma::smem_desc(sb + BP_OFF_B1H, BP_LBO_B, BP_SBO), d_b1l = umma::smem_desc(sb + BP_OFF_B1L, BP_LBO_B, BP_SBO);
      const uint64_t d_b2h = umma::smem_desc(sb + BP_OFF_B2H, BP_LBO_B, BP_SBO), d_b2l = umma::smem_desc(sb + BP_OFF_B2L, BP_LBO_B, BP_SBO);
      auto issue_fc1 = [&](int it_) {
        const int slot = it_ & 1;
        umma::mbar_wait(&a_full[slot], ((uint32_t)it_ >> 1) & 1u);
        umma::tc_fence_after();
        { [[maybe_unused]] const int it = it_; BP_TR(2); }
        const uint32_t sa = sb + BP_OFF_A + slot * BP_A_BYTES;
        const uint64_t d_ah = umma::smem_desc(sa, BP_LBO_A, BP_SBO);
        const uint64_t d_al = umma::smem_desc(sa + BP_CH1 * BP_LBO_A, BP_LBO_A, BP_SBO);
        const uint32_t d1 = tmem + (uint32_t)slot * 192;
#pragma unroll
        for (int ks = 0; ks < BP_K1 / 16; ++ks) {
          if (X3)
            umma::mma_bf16x3(d1, d_ah, d_al, d_b1h, d_b1l, ks * (2 * BP_LBO_A >> 4),
                             ks * (2 * BP_LBO_B >> 4), idesc, ks > 0);
          else
            umma::mma_bf16_ss(d1, d_ah + ks * (2 * BP_LBO_A >> 4), d_b1h + ks * (2 * BP_LBO_B >> 4),
                              idesc, ks > 0);
        }
        umma::mma_commit(&fc1_done[slot]);
        umma::mma_commit(&a_empty[slot]);
        { [[maybe_unused]] const int it = it_; BP_TR(3); }
      };
      umma::mbar_wait(wbar, 0);
      issue_fc1(0);
#pragma unroll 1
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) issue_fc1(it + 1);
        const int slot = it & 1;
        umma::mbar_wait(&h1_full[slot], ((uint32_t)it >> 1) & 1u);
        umma::tc_fence_after();
        BP_TR(0);
        const uint32_t d2 = tmem + (uint32_t)slot * 192 + 64;
        const uint32_t hh = tmem + (uint32_t)slot * 192 + 128, hl = hh + 32;
#pragma unroll
        for (int ks = 0; ks < BP_F / 16; ++ks) {
          const uint32_t boff = ks * (2 * BP_LBO_B >> 4);
          if (X3) {
            umma::mma_bf16_ts(d2, hl + ks * 8, d_b2h + boff, idesc, ks > 0);
            umma::mma_bf16_ts(d2, hh + ks * 8, d_b2l + boff, idesc, 1);
            umma::mma_bf16_ts(d2, hh + ks * 8, d_b2h + boff, idesc, 1);
          } else {
            umma::mma_bf16_ts(d2, hh + ks * 8, d_b2h + boff, idesc, ks > 0);
          }
        }
        umma::mma_commit(&fc2_done[slot]);
        BP_TR(1);
      }
    }
  } else {
    // ============================== epilogue groups =====================================
    const int g = (warp - BP_WARP_EPI) / BP_EPI_WARPS;          // group = tile parity
    const int ew = (warp - BP_WARP_EPI) % BP_EPI_WARPS;         // warp within the group
    const int lt = ew * 32 + lane;                              // thread within the group
    const int quad = warp & 3;                 // TMEM lane quadrant = CTA-level warp index % 4
    const int erow = quad * 32 + lane;
    const int ecol0 = (ew >> 2) * 32;          // the two warps sharing a quadrant split the 64 columns
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    const uint32_t tm_d1 = tmem + (uint32_t)g * 192, tm_d2 = tm_d1 + 64, tm_hh = tm_d1 + 128, tm_hl = tm_hh + 32;
    float* h2 = reinterpret_cast<float*>(smem + BP_OFF_H2 + g * BP_H2_BYTES);
    int* c_idx = reinterpret_cast<int*>(smem + BP_OFF_IDX + g * BP_IDX_BYTES);
    unsigned* seg_end = reinterpret_cast<unsigned*>(c_idx + BP_TILE);
    const int bar_id = 1 + g;

    // segment ids of this group's next tile, fetched a whole tile period ahead
    int c_pre = -1, cn_pre = -2;
    auto prefetch_seg = [&](int it_) {
      c_pre = -1;
      cn_pre = -2;
      if (lt < BP_TILE && it_ < my_tiles) {
        const int p = (blockIdx.x + it_ * gridDim.x) * BP_TILE + lt;
        if (p < P) c_pre = __ldg(pair_c + p);
        if (p + 1 < P && (lt & 15) != 15) cn_pre = __ldg(pair_c + p + 1);
      }
    };
    prefetch_seg(g);
#pragma unroll 1
    for (int it = g; it < my_tiles; it += 2) {
      const uint32_t n = (uint32_t)it >> 1;
      // ---- segment table of the tile: c per row + "last row of its run" masks ----------
      if (lt < BP_TILE) {
        c_idx[lt] = c_pre;
        // a run ends where the next row has another c, at the end of its 16-row slice, or
        // at the last valid pair; rows past P never flush
        const unsigned m = __ballot_sync(0xffffffffu, c_pre >= 0 && c_pre != cn_pre);
        if (lane == 0) seg_end[ew] = m;       // bits 0-15: slice 2*ew, bits 16-31: slice 2*ew+1
      }
      prefetch_seg(it + 2);

      // ---- epilogue 1: h1 = relu(acc + b1) -> bf16 hi / lo A operand in tensor memory -----
      umma::mbar_wait_relaxed(&fc1_done[g], n & 1u);
      umma::tc_fence_after();
      if (lt == 0) BP_TR(8);
      {
        float v[32];
        umma::tmem_ld32(tm_d1 + tlane + ecol0, v);
        umma::tmem_ld_wait();
        uint32_t hh[16], hl[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = ecol0 + 2 * e;
          umma::split_bf16x2(fmaxf(v[2 * e] + bias1[col], 0.f), fmaxf(v[2 * e + 1] + bias1[col + 1], 0.f),
                             hh[e], hl[e]);
        }
        const uint32_t c0 = (uint32_t)(ecol0 >> 1);
        umma::tmem_st8(tm_hh + tlane + c0, reinterpret_cast<const uint32_t(&)[8]>(hh[0]));
        umma::tmem_st8(tm_hh + tlane + c0 + 8, reinterpret_cast<const uint32_t(&)[8]>(hh[8]));
        if (X3) {
          umma::tmem_st8(tm_hl + tlane + c0, reinterpret_cast<const uint32_t(&)[8]>(hl[0]));
          umma::tmem_st8(tm_hl + tlane + c0 + 8, reinterpret_cast<const uint32_t(&)[8]>(hl[8]));
        }
        umma::tmem_st_wait();
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&h1_full[g]);
      if (lt == 0) BP_TR(9);

      // ---- epilogue 2: h2 = relu(acc + b2) -> fp32 staging tile ----------------------------
      umma::mbar_wait_relaxed(&fc2_done[g], n & 1u);
      umma::tc_fence_after();
      if (lt == 0) BP_TR(10);
      {
        float v[32];
        umma::tmem_ld32(tm_d2 + tlane + ecol0, v);
        umma::tmem_ld_wait();
        float* dst = h2 + erow * BP_LDH2 + ecol0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int col = ecol0 + k * 4;
          *reinterpret_cast<float4*>(dst + k * 4) =
              make_float4(fmaxf(v[k * 4 + 0] + bias2[col + 0], 0.f), fmaxf(v[k * 4 + 1] + bias2[col + 1], 0.f),
                          fmaxf(v[k * 4 + 2] + bias2[col + 2], 0.f), fmaxf(v[k * 4 + 3] + bias2[col + 3], 0.f));
        }
      }
      umma::tc_fence_before();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(BP_EPI_WARPS * 32) : "memory");
      if (lt == 0) BP_TR(11);

      // ---- segmented max over the tile's rows -----------------------------------------------
      // thread = (16-row slice, column pair): one 8-byte shared load per row
      {
        const int j = (lt & 31) * 2;
        const int slice = lt >> 5;
        const int r0 = slice * 16;
        const unsigned ends = (seg_end[slice >> 1] >> ((slice & 1) * 16)) & 0xffffu;
        const float* col = h2 + r0 * BP_LDH2 + j;
        float cur0 = 0.f, cur1 = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 x = *reinterpret_cast<const float2*>(col + r * BP_LDH2);
          cur0 = fmaxf(cur0, x.x);
          cur1 = fmaxf(cur1, x.y);
          if ((ends >> r) & 1u) {        // warp uniform
            int* dst = reinterpret_cast<int*>(pooled + (size_t)c_idx[r0 + r] * BP_F + j);
            atomicMax(dst, __float_as_int(cur0));
            atomicMax(dst + 1, __float_as_int(cur1));
            cur0 = cur1 = 0.f;
          }
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(BP_EPI_WARPS * 32) : "memory");
      if (lt == 0) BP_TR(12);
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace gn

#ifdef BP_TRACE
extern "C" int gn_block_pair_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::bp_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

static int launch_pair_pipe(const char* name, bool x3, const float* pw, int w, const void* feats_hl,
                            const void* nfeats_hl, int r, const int32_t* pair_c,
                            const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                            const float* b1, const float* b2, const void* wimg, int f,
                            float* pooled, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0, "%s: negative capacity", name);
  if (w != gn::BP_W || r != gn::BP_R || f != gn::BP_F) {
    gn::set_error("%s: fused kernel is built for w=%d r=%d f=%d (got %d, %d, %d)", name,
                  gn::BP_W, gn::BP_R, gn::BP_F, w, r, f);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(pw && feats_hl && nfeats_hl && pair_c && pair_n && num_pairs && b1 && b2 && wimg &&
                 pooled, "%s: null pointer", name);
  GN_REQUIRE((((uintptr_t)pw | (uintptr_t)feats_hl | (uintptr_t)nfeats_hl | (uintptr_t)wimg) & 15) == 0,
             "%s: pointers must be 16-byte aligned", name);
  const void* kern = x3 ? (const void*)gn::block_pair_pipe_kernel<true>
                        : (const void*)gn::block_pair_pipe_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)gn::BP_SMEM);
  if (e != cudaSuccess) {
    gn::set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(capacity, gn::BP_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  if (x3)
    gn::block_pair_pipe_kernel<true><<<grid, gn::BP_THREADS, gn::BP_SMEM, (cudaStream_t)stream>>>(
        pw, static_cast<const float*>(feats_hl), static_cast<const float*>(nfeats_hl), pair_c, pair_n,
        num_pairs, capacity, b1, b2, static_cast<const unsigned char*>(wimg), pooled);
  else
    gn::block_pair_pipe_kernel<false><<<grid, gn::BP_THREADS, gn::BP_SMEM, (cudaStream_t)stream>>>(
        pw, static_cast<const float*>(feats_hl), static_cast<const float*>(nfeats_hl), pair_c, pair_n,
        num_pairs, capacity, b1, b2, static_cast<const unsigned char*>(wimg), pooled);
  GN_CHECK_LAUNCH(name);
  return GN_OK;
}

extern "C" int gn_block_pair_fwd_pipe(const float* pw, int w, const void* feats_hl,
                                      const void* nfeats_hl, int r, const int32_t* pair_c,
                                      const int32_t* pair_n, const int32_t* num_pairs,
                                      int capacity, const float* b1, const float* b2,
                                      const void* wimg, int f, float* pooled, gn_stream_t stream) {
  return launch_pair_pipe("gn_block_pair_fwd_pipe", true, pw, w, feats_hl, nfeats_hl, r, pair_c,
                          pair_n, num_pairs, capacity, b1, b2, wimg, f, pooled, stream);
}

extern "C" int gn_block_pair_fwd_pipe_bf16(const float* pw, int w, const void* feats_hl,
                                           const void* nfeats_hl, int r, const int32_t* pair_c,
                                           const int32_t* pair_n, const int32_t* num_pairs,
                                           int capacity, const float* b1, const float* b2,
                                           const void* wimg, int f, float* pooled,
                                           gn_stream_t stream) {
  return launch_pair_pipe("gn_block_pair_fwd_pipe_bf16", false, pw, w, feats_hl, nfeats_hl, r,
                          pair_c, pair_n, num_pairs, capacity, b1, b2, wimg, f, pooled, stream);
}
