// Pair features on the tensor cores: geometry (A4) + the 3-layer pair-feature
// MLP (A5) width -> 256 -> 256 -> 32 with ReLU after every layer, fused.
//
// Tile = 128 pairs (the M of one tcgen05.mma), one persistent warp-specialised CTA per SM
// (pwfeat_pipe_kernel below: epilogue / MMA / W2-producer / geometry warps, mbarrier
// hand-offs, the hidden layers processed in 64-column quarters).
//   geometry : one thread per pair computes the 9 hand-crafted features in fp32
//              (pair_geometry, same rounding as the reference) and writes a K=16
//              bf16 hi/lo operand row, one tile ahead of the tensor core;
//   layer 1  : [128 x 16] x [16 x 64] per quarter (3 UMMAs) -> a 64-column accumulator;
//              single class: K rows = [c_score, n_score, 7 geometry]; multi class:
//              the one-hot score block degenerates to two weight-row gathers that
//              the epilogue adds in fp32 (K rows = 7 geometry only);
//   layer 2  : [128 x 256] x [256 x 256] (16 k-steps x 3 UMMAs, A operand in tensor
//              memory) -> TMEM [256,512).  W2 (256 KB as bf16 hi+lo) does not fit next
//              to the activations, so it is streamed per k-step (16 KB) from a pre-split
//              operand image in global memory (L2 resident) through a cp.async.bulk ring
//              (mbarrier complete_tx / tcgen05.commit);
//   layer 3  : [128 x 256] x [256 x (32 hi | 32 lo)] -> the first 64 columns of the layer-2
//              accumulator (hi and lo weight parts stacked along N, added by the epilogue),
//              again quarter by quarter as the epilogue drains it.
// Only pw_out[P,32] goes back to HBM (whole 128-byte rows through a padded staging tile);
// the [P,256] activations never leave the SM.
// bf16x3: every fp32 operand is split into bf16 hi + lo and each product is
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with fp32 accumulation (gn_umma.cuh).
// Measured per-tile timeline and what bounds each phase: profiles/r1_tc_kernels.md.
#include "gn_pairfeat.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int PT_TILE = 128;
constexpr int PT_H = 256, PT_O = 32;
constexpr uint32_t PT_SBO = 128;
constexpr uint32_t PT_LBO_A = PT_TILE * 16;      // activation / A1 chunk pitch (2048)
constexpr uint32_t PT_LBO_W = PT_H * 16;         // 256-row weight chunk pitch (4096)
constexpr uint32_t PT_LBO_W3 = 2 * PT_O * 16;    // W3 chunk pitch (1024): 32 hi rows, then 32 lo rows
constexpr uint32_t PT_STAGE = 2 * 2 * PT_LBO_W;  // one W2 k-step: (hi, lo) x 2 chunks = 16 KB

// prepared weight image (global workspace), bytes
// The W2 image can be replicated (CTA i streams copy i % COPIES).  Measured: 8 copies
// do not change the kernel time (profiles/r1_tc_kernels.md), so L2 line contention is
// not what limits the W2 stream; one copy is kept.
constexpr int PT_W2_COPIES = 1;
constexpr uint32_t PT_IMG_W2 = 0;                            // COPIES x 16 k-steps x 16 KB
constexpr uint32_t PT_IMG_B1 = PT_IMG_W2 + PT_W2_COPIES * 16 * PT_STAGE;   // hi 8 KB, lo 8 KB
constexpr uint32_t PT_IMG_B3 = PT_IMG_B1 + 2 * 2 * PT_LBO_W; // hi 16 KB, lo 16 KB
constexpr uint32_t PT_IMG_BYTES = PT_IMG_B3 + 32 * PT_LBO_W3;


// ---------------------------------------------------------------------------------
// weight preparation: fp32 [in,out] -> bf16 hi/lo K-major operand images
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void put_split(unsigned char* hi_base, unsigned char* lo_base,
                                          uint32_t off, int e, float x) {
  __nv_bfloat16 h, l;
  umma::split_bf16(x, h, l);
  reinterpret_cast<__nv_bfloat16*>(hi_base + off)[e] = h;
  reinterpret_cast<__nv_bfloat16*>(lo_base + off)[e] = l;
}

__global__ void pwfeat_prepare_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                      const float* __restrict__ w3, int num_classes,
                                      unsigned char* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // W2: k-step ks, chunk c (8 k), row n
  for (int i = tid; i < PT_H * PT_H; i += nth) {
    const int k = i / PT_H, n = i - k * PT_H;
    const int ks = k >> 4, c = (k >> 3) & 1, e = k & 7;
    const float x = __ldg(w2 + i);
    for (int rep = 0; rep < PT_W2_COPIES; ++rep) {
      unsigned char* st = img + PT_IMG_W2 + (rep * 16 + ks) * PT_STAGE;
      put_split(st, st + 2 * PT_LBO_W, c * PT_LBO_W + n * 16, e, x);
    }
  }
  // B1: 16 K rows selected from W1 (zero padded)
  const bool multi = num_classes > 1;
  const int gbase = multi ? 2 * num_classes : 2;
  for (int i = tid; i < 16 * PT_H; i += nth) {
    const int k = i / PT_H, n = i - k * PT_H;
    int src_row;
    if (multi) src_row = k < 7 ? gbase + k : -1;      // geometry only
    else src_row = k < 9 ? k : -1;                     // c_score, n_score, geometry
    const float x = src_row >= 0 ? __ldg(w1 + (size_t)src_row * PT_H + n) : 0.f;
    unsigned char* st = img + PT_IMG_B1;
    put_split(st, st + 2 * PT_LBO_W, (k >> 3) * PT_LBO_W + n * 16, k & 7, x);
  }
  // B3: W3^T, 32 chunks x (32 hi rows | 32 lo rows): the hi and lo parts stacked along N, so
  // that one N = 64 UMMA forms a_hi * [b_hi | b_lo] (layer 3 below)
  for (int i = tid; i < PT_H * PT_O; i += nth) {
    const int k = i / PT_O, n = i - k * PT_O;
    unsigned char* st = img + PT_IMG_B3;
    put_split(st, st + PT_O * 16, (k >> 3) * PT_LBO_W3 + n * 16, k & 7, __ldg(w3 + i));
  }
}

// ---------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------
// Pipelined version: warp-specialised roles, the hidden layer processed in four
// 64-column quarters with double-buffered layer-1 accumulators and activation operands,
// so the epilogue of quarter q+1 runs while the layer-2 UMMAs of quarter q do.
//
//   warps 0-15  epilogue : TMEM lane quadrant = warp % 4, 16 of a quarter's 64 columns
//   warp  16    MMA      : one thread issues every tcgen05.mma / commit
//   warp  17    producer : one thread streams W2 k-steps through the cp.async.bulk ring
//   warps 18-21 geometry : one thread per pair, one tile ahead (A1 double-buffered)
//
// TMEM (512 columns): L1 accumulators 2 x 64 | activation operands H[2] (K = 64 each:
// 32 columns bf16 hi + 32 lo) | layer-2 accumulator 256 (its first 64 columns are reused
// as the layer-3 accumulator once the first quarter has been drained).
// Hand-offs are mbarriers: a1_full/a1_empty (geometry <-> MMA), l1_done (MMA -> epilogue),
// h_full/h_empty (epilogue <-> MMA), acc2_done, acc3_done, tab_free (epilogue -> geometry).
// Every operand-buffer production p uses H[p % 2]; it may start once consumption p - 2 has
// completed (tcgen05.commit -> h_empty).
// ---------------------------------------------------------------------------------
constexpr int PP_EPI_WARPS = 16;
constexpr int PP_WARP_MMA = 16, PP_WARP_LOAD = 17, PP_WARP_GEO = 18;
constexpr int PP_THREADS = (PP_WARP_GEO + 4) * 32;   // 704
constexpr int PP_Q = 64;                              // hidden columns per quarter
#ifndef PP_RING_N
#define PP_RING_N 8
#endif
constexpr int PP_RING = PP_RING_N;                    // W2 k-step stages in flight
constexpr uint32_t PP_OUT_PITCH = 144;                // output staging row pitch (bytes): 128 + 16,
                                                      // so 8 consecutive rows hit 8 distinct 16-byte bank groups

constexpr uint32_t PQ_RING = 0;
constexpr uint32_t PQ_B1 = PQ_RING + PP_RING * PT_STAGE;
constexpr uint32_t PQ_B3 = PQ_B1 + 2 * 2 * PT_LBO_W;
constexpr uint32_t PQ_A1 = PQ_B3 + 32 * PT_LBO_W3;                // 2 buffers x (hi 4 KB + lo 4 KB)
constexpr uint32_t PQ_BIAS = PQ_A1 + 2 * 2 * 2 * PT_LBO_A;
constexpr uint32_t PQ_ROW = PQ_BIAS + (2 * PT_H + PT_O) * 4;      // 2 x (sc, sn, rc, rn)[128]
constexpr uint32_t PQ_OUT = PQ_ROW + 2 * 4 * PT_TILE * 4;          // output staging tile
constexpr uint32_t PQ_BAR = PQ_OUT + PT_TILE * PP_OUT_PITCH;
constexpr int PQ_NBAR = 2 * PP_RING + 14;
constexpr uint32_t PQ_BYTES = PQ_BAR + PQ_NBAR * 8;
static_assert(PQ_BYTES <= 227 * 1024, "pipelined pair MLP exceeds shared memory");

#ifdef PP_TRACE
__device__ long long pp_trace[64];
#define PP_TR(i) do { if (blockIdx.x == 0 && it == 2) pp_trace[i] = clock64(); } while (0)
#else
#define PP_TR(i) do { } while (0)
#endif

// X3: bf16x3 products (fp32 semantics); !X3: plain bf16 operands, a third of the UMMAs and
// half of the W2 stream (BASELINE configs[2] arithmetic).
template <bool MULTI, bool X3>
__global__ void __launch_bounds__(PP_THREADS, 1)
pwfeat_pipe_kernel(const float* __restrict__ dets, const float* __restrict__ scores,
                   const int32_t* __restrict__ classes, const int32_t* __restrict__ pair_c,
                   const int32_t* __restrict__ pair_n, const float* __restrict__ pair_iou,
                   const int32_t* __restrict__ num_pairs, int capacity, int num_classes, float mult,
                   const float* __restrict__ w1, const float* __restrict__ b1,
                   const float* __restrict__ b2, const float* __restrict__ b3,
                   const unsigned char* __restrict__ img, float* __restrict__ pw_out,
                   uint4* __restrict__ pw_hl) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + PT_TILE - 1) / PT_TILE;
  if ((int)blockIdx.x >= num_tiles) return;
  const int my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  float* bias1 = reinterpret_cast<float*>(smem + PQ_BIAS);
  float* bias2 = bias1 + PT_H;
  float* bias3 = bias2 + PT_H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PQ_BAR);
  uint64_t* full = bars;                       // [RING] W2 stage landed
  uint64_t* empty = full + PP_RING;            // [RING] W2 stage consumed
  uint64_t* a1_full = empty + PP_RING;         // [2]
  uint64_t* a1_empty = a1_full + 2;            // [2]
  uint64_t* l1_done = a1_empty + 2;            // [2]
  uint64_t* h_full = l1_done + 2;              // [2]
  uint64_t* h_empty = h_full + 2;              // [2]
  uint64_t* tab_free = h_empty + 2;            // [2]
  uint64_t* acc2_done = tab_free + 2;          // [1]
  uint64_t* acc3_done = acc2_done + 1;         // [1]

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (t == 0) {
    for (int s = 0; s < PP_RING; ++s) {
      umma::mbar_init(&full[s], 1);
      umma::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(&a1_full[b], 4 * 32);        // every geometry thread arrives itself
      umma::mbar_init(&a1_empty[b], 1);
      umma::mbar_init(&l1_done[b], 1);
      umma::mbar_init(&h_full[b], PP_EPI_WARPS);
      umma::mbar_init(&h_empty[b], 1);
      umma::mbar_init(&tab_free[b], PP_EPI_WARPS * 32);
    }
    umma::mbar_init(acc2_done, 1);
    umma::mbar_init(acc3_done, 1);
    umma::fence_barrier_init();
  }
  for (int i = t; i < (int)(2 * 2 * PT_LBO_W) / 16; i += PP_THREADS)
    reinterpret_cast<uint4*>(smem + PQ_B1)[i] = __ldg(reinterpret_cast<const uint4*>(img + PT_IMG_B1) + i);
  for (int i = t; i < (int)(32 * PT_LBO_W3) / 16; i += PP_THREADS)
    reinterpret_cast<uint4*>(smem + PQ_B3)[i] = __ldg(reinterpret_cast<const uint4*>(img + PT_IMG_B3) + i);
  for (int i = t; i < PT_H; i += PP_THREADS) {
    bias1[i] = __ldg(b1 + i);
    bias2[i] = __ldg(b2 + i);
  }
  if (t < PT_O) bias3[t] = __ldg(b3 + t);
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();

  const uint32_t tmem = tmem_base_s;
  const uint32_t tm_l1 = tmem;              // L1 accumulator buffers: +0, +64
  const uint32_t tm_h = tmem + 128;         // H[b] at +64 b: hi 32 columns, lo 32 columns
  const uint32_t tm_l2 = tmem + 256;        // layer-2 accumulator (256), layer-3 in its first 64

  if (warp < PP_EPI_WARPS) {
    // =========================== epilogue warps =====================================
    const int quad = warp & 3, cg = warp >> 2;
    const int erow = quad * 32 + lane;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    uint32_t prod = 0;
    // relu(v + bias) of this thread's 16 columns -> H[prod % 2] (bf16 hi | lo), then hand over
    auto produce = [&](float (&v)[16], const float* bias16) {
      uint32_t hh[8], hl[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        umma::split_bf16x2(fmaxf(v[2 * e] + bias16[2 * e], 0.f),
                           fmaxf(v[2 * e + 1] + bias16[2 * e + 1], 0.f), hh[e], hl[e]);
      const uint32_t b = prod & 1u, u = prod >> 1;
      if (u >= 1) umma::mbar_wait_relaxed(&h_empty[b], (u - 1) & 1u);
      umma::tc_fence_after();
      const uint32_t dst = tm_h + b * 64 + tlane + (uint32_t)cg * 8;
      umma::tmem_st8(dst, hh);
      if (X3) umma::tmem_st8(dst + 32, hl);
      umma::tmem_st_wait();
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&h_full[b]);
      ++prod;
    };
    // The output tile of the previous iteration: its values wait in registers (o) while the
    // first two layer-1 quarters of the next tile are handed to the tensor core, and go out
    // through the padded staging tile as whole 128-byte rows (a row-per-thread store would
    // cost one L1 wavefront per lane) while the layer-2 UMMAs of that tile run - the epilogue
    // warps are idle there, and layer 2 starts without waiting for these stores.
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int o_tile = -1;
    auto store_output = [&]() {
      unsigned char* stage = smem + PQ_OUT;
      if (pw_out != nullptr) {
        float4* dst = reinterpret_cast<float4*>(stage + erow * PP_OUT_PITCH + cg * 32);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        asm volatile("bar.sync 1, %0;" ::"n"(PP_EPI_WARPS * 32) : "memory");
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int idx = i * (PP_EPI_WARPS * 32) + t;       // 1024 chunks of 16 bytes
          const int r = idx >> 3, ch = idx & 7;
          const int p = o_tile * PT_TILE + r;
          const float4 x = *reinterpret_cast<const float4*>(stage + r * PP_OUT_PITCH + ch * 16);
          if (p < P) *reinterpret_cast<float4*>(pw_out + (size_t)p * PT_O + ch * 4) = x;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PP_EPI_WARPS * 32) : "memory");   // staging reusable
      }
      if (pw_hl != nullptr) {
        // the same row as the block kernels' tensor-core operand: [32 bf16 hi | 32 bf16 lo]
        // (128 bytes per pair, gn_block_tma.cu loads it with one tensor-map TMA per tile)
        uint4 h, l;
        umma::split_bf16x2(o[0], o[1], h.x, l.x);
        umma::split_bf16x2(o[2], o[3], h.y, l.y);
        umma::split_bf16x2(o[4], o[5], h.z, l.z);
        umma::split_bf16x2(o[6], o[7], h.w, l.w);
        *reinterpret_cast<uint4*>(stage + erow * PP_OUT_PITCH + cg * 16) = h;
        *reinterpret_cast<uint4*>(stage + erow * PP_OUT_PITCH + 64 + cg * 16) = l;
        asm volatile("bar.sync 1, %0;" ::"n"(PP_EPI_WARPS * 32) : "memory");
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int idx = i * (PP_EPI_WARPS * 32) + t;
          const int r = idx >> 3, ch = idx & 7;
          const int p = o_tile * PT_TILE + r;
          const uint4 x = *reinterpret_cast<const uint4*>(stage + r * PP_OUT_PITCH + ch * 16);
          if (p < P) pw_hl[(size_t)p * 8 + ch] = x;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PP_EPI_WARPS * 32) : "memory");
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int tb = it & 1;
      const float* row_sc = reinterpret_cast<const float*>(smem + PQ_ROW) + tb * 4 * PT_TILE;
      const float* row_sn = row_sc + PT_TILE;
      const int* row_rc = reinterpret_cast<const int*>(row_sn + PT_TILE);
      const int* row_rn = row_rc + PT_TILE;
      // the score tables were written by the geometry warps before they arrived on a1_full;
      // the epilogue otherwise only sees that through the MMA thread (a1_full -> UMMA ->
      // commit -> l1_done), so it acquires the barrier itself (always already complete;
      // the phase cannot advance again before this tile's tab_free arrival below)
      if (MULTI) umma::mbar_wait(&a1_full[tb], ((uint32_t)it >> 1) & 1u);
      // ---- layer-1 quarters -> H ------------------------------------------------------
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const uint32_t n = 2u * (uint32_t)it + (uint32_t)(q >> 1);
        umma::mbar_wait_relaxed(&l1_done[q & 1], n & 1u);
        umma::tc_fence_after();
        if (t == 0) PP_TR(32 + 2 * q);
        float v[16];
        umma::tmem_ld16(tm_l1 + (uint32_t)(q & 1) * 64 + tlane + (uint32_t)cg * 16, v);
        umma::tmem_ld_wait();
        const int col = q * PP_Q + cg * 16;
        if (MULTI) {
          const float sc = row_sc[erow], sn = row_sn[erow];
          const float* wc = w1 + (size_t)row_rc[erow] * PT_H + col;
          const float* wn = w1 + (size_t)row_rn[erow] * PT_H + col;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 a = ldg4(wc + g * 4), b = ldg4(wn + g * 4);
            v[g * 4 + 0] += sc * a.x + sn * b.x;
            v[g * 4 + 1] += sc * a.y + sn * b.y;
            v[g * 4 + 2] += sc * a.z + sn * b.z;
            v[g * 4 + 3] += sc * a.w + sn * b.w;
          }
        }
        produce(v, bias1 + col);
        if (t == 0) PP_TR(33 + 2 * q);
        if (q == 1 && it > 0) {   // both operand buffers are queued: idle until l1_done q2
          store_output();
          if (t == 0) PP_TR(46);
        }
      }
      umma::mbar_arrive(&tab_free[tb]);   // this thread's score-table reads of the tile are done
      // ---- layer-2 accumulator quarters -> H --------------------------------------------
      umma::mbar_wait_relaxed(acc2_done, (uint32_t)it & 1u);
      umma::tc_fence_after();
      if (t == 0) PP_TR(40);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        float v[16];
        const int col = q * PP_Q + cg * 16;
        umma::tmem_ld16(tm_l2 + tlane + (uint32_t)col, v);
        umma::tmem_ld_wait();
        produce(v, bias2 + col);
        if (t == 0) PP_TR(41 + q);
      }
      // ---- output: relu(acc3 + b3), this thread's 8 columns of its row, into registers; the
      //      stores are deferred into the next tile (store_output above) ----------------------
      umma::mbar_wait_relaxed(acc3_done, (uint32_t)it & 1u);
      umma::tc_fence_after();
      if (t == 0) PP_TR(45);
      {
        float v[8];
        umma::tmem_ld8(tm_l2 + tlane + (uint32_t)cg * 8, v);
        if (X3) {
          float v2[8];
          umma::tmem_ld8(tm_l2 + tlane + (uint32_t)(PT_O + cg * 8), v2);
          umma::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += v2[e];
        } else {
          umma::tmem_ld_wait();
        }
        if (t == 0) PP_TR(47);
        const float* bb = bias3 + cg * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[e] + bb[e], 0.f);
        o_tile = tile;
      }
      umma::tc_fence_before();   // the next production's arrive orders these loads before the
                                 // UMMAs that overwrite the accumulator
    }
    store_output();              // the last tile's
  } else if (warp == PP_WARP_MMA) {
    // =============================== MMA issuer =======================================
    // the whole warp runs the issue sequence with warp-uniform operands; elect.sync inside the
    // *_elect forms picks the issuing lane (gn_umma.cuh)
    {
      const uint32_t idesc64 = umma::idesc_bf16_f32(PT_TILE, PP_Q);
      const uint32_t idesc256 = umma::idesc_bf16_f32(PT_TILE, PT_H);
      const uint32_t idesc32 = umma::idesc_bf16_f32(PT_TILE, PT_O);
      const uint32_t s_ring = umma::smem_u32(smem + PQ_RING);
      const uint32_t s_a1 = umma::smem_u32(smem + PQ_A1);
      const uint32_t s_b1h = umma::smem_u32(smem + PQ_B1), s_b1l = s_b1h + 2 * PT_LBO_W;
      const uint32_t s_b3 = umma::smem_u32(smem + PQ_B3);
      const uint64_t d_b1h = umma::smem_desc(s_b1h, PT_LBO_W, PT_SBO), d_b1l = umma::smem_desc(s_b1l, PT_LBO_W, PT_SBO);
      const uint64_t d_ringh = umma::smem_desc(s_ring, PT_LBO_W, PT_SBO);
      const uint64_t d_ringl = umma::smem_desc(s_ring + 2 * PT_LBO_W, PT_LBO_W, PT_SBO);
      const uint64_t d_b3 = umma::smem_desc(s_b3, PT_LBO_W3, PT_SBO);
      uint32_t cons = 0;
      // layer 1 of quarter q of the tile whose A1 sits in buffer ab -> L1 accumulator q % 2
      auto issue_l1 = [&](int ab, int q) {
        const uint64_t d_ah = umma::smem_desc(s_a1 + (uint32_t)ab * (4 * PT_LBO_A), PT_LBO_A, PT_SBO);
        const uint64_t d_al = umma::smem_desc(s_a1 + (uint32_t)ab * (4 * PT_LBO_A) + 2 * PT_LBO_A, PT_LBO_A, PT_SBO);
        if (X3)
          umma::mma_bf16x3_elect(tm_l1 + (uint32_t)(q & 1) * 64, d_ah, d_al, d_b1h, d_b1l, 0,
                           (uint32_t)q * (PP_Q * 16 >> 4), idesc64, 0);
        else
          umma::mma_bf16_ss_elect(tm_l1 + (uint32_t)(q & 1) * 64, d_ah,
                            d_b1h + (uint32_t)q * (PP_Q * 16 >> 4), idesc64, 0);
        umma::mma_commit_elect(&l1_done[q & 1]);
      };
      umma::mbar_wait(&a1_full[0], 0);
      umma::tc_fence_after();
      issue_l1(0, 0);
      issue_l1(0, 1);
#pragma unroll 1
      for (int it = 0; it < my_tiles; ++it) {
        const int ab = it & 1;
        PP_TR(0);
        // ---- layer 2: quarter q of K as soon as its operand is there -------------------
        // The issuing thread is the pace maker of this phase (48 N = 256 UMMAs of 128 cycles each
        // against ~90 cycles of issue path per UMMA plus ~125 per mbarrier poll), so the whole
        // sequence is unrolled: 16 k-steps per tile over a ring of 8 stages make the stage and
        // the parity of every wait compile-time constants (and so every descriptor), and the
        // four ring stages of a quarter are polled together instead of one round trip each.
        static_assert(PP_RING == 8 || PP_RING == 4, "stage / parity constants below assume 16 k-steps over 8 or 4 stages");
        // operand q + 1 and its ring stages are polled right AFTER quarter q has been issued:
        // the ~12 x 128 cycles of queued UMMAs cover the polls' round trips
        auto acquire = [&](int q) {
          const uint32_t hb = (uint32_t)q & 1u, hpar = ((uint32_t)q >> 1) & 1u;   // cons = 8 it + q
#ifndef PP_NORING
          // the ring stages first (they are there long before the operand is)
          // g = 16 it + 4 q + ksl: stage g % RING, parity (g / RING) & 1 - both independent of it
          const uint32_t s0 = (4u * q) & (PP_RING - 1), par = ((4u * q) / PP_RING) & 1u;
          bool ok = umma::mbar_try_wait(&full[s0], par);
          ok &= umma::mbar_try_wait(&full[s0 + 1], par);
          ok &= umma::mbar_try_wait(&full[s0 + 2], par);
          ok &= umma::mbar_try_wait(&full[s0 + 3], par);
          if (!ok) {
#pragma unroll
            for (int ksl = 0; ksl < 4; ++ksl) umma::mbar_wait(&full[s0 + ksl], par);
          }
#endif
          umma::mbar_wait(&h_full[hb], hpar);
          umma::tc_fence_after();
        };
        acquire(0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t hb = (uint32_t)q & 1u;
          PP_TR(1 + 2 * q);
          const uint32_t a_hi = tm_h + hb * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int ksl = 0; ksl < 4; ++ksl) {
            const uint32_t s = (4u * q + ksl) & (PP_RING - 1);
            const uint32_t boff = s * (PT_STAGE >> 4);
            const uint32_t acc = (q | ksl) != 0;
            if (X3) {
              umma::mma_bf16_ts_elect(tm_l2, a_lo + ksl * 8, d_ringh + boff, idesc256, acc);
              umma::mma_bf16_ts_elect(tm_l2, a_hi + ksl * 8, d_ringl + boff, idesc256, 1);
              umma::mma_bf16_ts_elect(tm_l2, a_hi + ksl * 8, d_ringh + boff, idesc256, 1);
            } else {
              umma::mma_bf16_ts_elect(tm_l2, a_hi + ksl * 8, d_ringh + boff, idesc256, acc);
            }
            umma::mma_commit_elect(&empty[s]);
          }
          umma::mma_commit_elect(&h_empty[hb]);
          PP_TR(2 + 2 * q);
          if (q < 2) {
            issue_l1(ab, q + 2);
            if (q == 1) umma::mma_commit_elect(&a1_empty[ab]);   // A1[ab] has been read for the last time
          }
          if (q < 3) acquire(q + 1);
        }
        cons += 4;
        umma::mma_commit_elect(acc2_done);
        // ---- layer 3 ------------------------------------------------------------------------
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const uint32_t hb = cons & 1u;
          umma::mbar_wait(&h_full[hb], (cons >> 1) & 1u);
          umma::tc_fence_after();
          PP_TR(10 + 2 * q);
          const uint32_t a_hi = tm_h + hb * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int ksl = 0; ksl < 4; ++ksl) {
            const uint32_t boff = (uint32_t)((q * 4 + ksl) * 2) * (PT_LBO_W3 >> 4);
            const uint32_t acc = (q | ksl) != 0;
            if (X3) {
              // a_hi * [b_hi | b_lo] -> columns [0,32) | [32,64), a_lo * b_hi -> [0,32): two reads
              // of the activation operand through the TMEM port instead of three (the port
              // bounds this phase); the epilogue adds the two column blocks
              umma::mma_bf16_ts_elect(tm_l2, a_hi + ksl * 8, d_b3 + boff, idesc64, acc);
              umma::mma_bf16_ts_elect(tm_l2, a_lo + ksl * 8, d_b3 + boff, idesc32, 1);
            } else {
              umma::mma_bf16_ts_elect(tm_l2, a_hi + ksl * 8, d_b3 + boff, idesc32, acc);
            }
          }
          umma::mma_commit_elect(&h_empty[hb]);
          ++cons;
        }
        umma::mma_commit_elect(acc3_done);
        PP_TR(20);
        // ---- the next tile's first two layer-1 quarters queue up behind layer 3 ---------------
        if (it + 1 < my_tiles) {
          const int nb = (it + 1) & 1;
          umma::mbar_wait(&a1_full[nb], (uint32_t)((it + 1) >> 1) & 1u);
          umma::tc_fence_after();
          issue_l1(nb, 0);
          issue_l1(nb, 1);
        }
      }
    }
  } else if (warp == PP_WARP_LOAD) {
    // ============================== W2 producer =========================================
    if (lane == 0) {
      const uint32_t s_ring = umma::smem_u32(smem + PQ_RING);
      uint32_t total = (uint32_t)my_tiles * 16u;
#ifdef PP_NORING
      total = 0;
#endif
      for (uint32_t i = 0; i < total; ++i) {
        const uint32_t s = i % PP_RING;
        if (i >= PP_RING) umma::mbar_wait_relaxed(&empty[s], ((i / PP_RING) - 1) & 1u);
        // a stage image is [hi: 2 chunks | lo: 2 chunks]; plain bf16 streams the hi half only
        bulk_g2s(s_ring + s * PT_STAGE, img + PT_IMG_W2 + (size_t)((blockIdx.x % PT_W2_COPIES) * 16 + (i & 15u)) * PT_STAGE,
                 X3 ? PT_STAGE : PT_STAGE / 2, &full[s]);
      }
    }
  } else {
    // ================================ geometry ============================================
    const int row = (warp - PP_WARP_GEO) * 32 + lane;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int ab = it & 1;
      const uint32_t u = (uint32_t)it >> 1;
      const int p = tile * PT_TILE + row;
      const bool live = p < P;
      int c = 0, n = 0;
      float4 cbx = make_float4(0.f, 0.f, 1.f, 1.f), nbx = cbx;
      float iou = 0.f, sc = 0.f, sn = 0.f;
      int rc = 0, rn = 0;
      if (live) {
        c = __ldg(pair_c + p);
        n = __ldg(pair_n + p);
        iou = __ldg(pair_iou + p);
        cbx = ldg4(dets + (size_t)c * 4);
        nbx = ldg4(dets + (size_t)n * 4);
        sc = __fmul_rn(__ldg(scores + c), mult);
        sn = __fmul_rn(__ldg(scores + n), mult);
        if (MULTI) {
          rc = __ldg(classes + c) - 1;               // one-based classes (network.py:413-419)
          rn = num_classes + __ldg(classes + n) - 1;
        }
      }
      float g[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (live) {
        pair_geometry_dist(cbx, nbx, iou, mult, g);
        pair_geometry_logs(cbx, nbx, mult, g + 4);
      }
      uint4 h0, l0, h1 = make_uint4(0u, 0u, 0u, 0u), l1 = h1;
      if (MULTI) {      // [iou xd yd l2 | wd hd ad 0] [0 ...]
        umma::split_bf16x2(g[0], g[1], h0.x, l0.x);
        umma::split_bf16x2(g[2], g[3], h0.y, l0.y);
        umma::split_bf16x2(g[4], g[5], h0.z, l0.z);
        umma::split_bf16x2(g[6], 0.f, h0.w, l0.w);
      } else {          // [sc sn | iou xd yd l2 | wd hd] [ad 0 ...]
        umma::split_bf16x2(sc, sn, h0.x, l0.x);
        umma::split_bf16x2(g[0], g[1], h0.y, l0.y);
        umma::split_bf16x2(g[2], g[3], h0.z, l0.z);
        umma::split_bf16x2(g[4], g[5], h0.w, l0.w);
        umma::split_bf16x2(g[6], 0.f, h1.x, l1.x);
      }
      if (u >= 1) {
        umma::mbar_wait_relaxed(&a1_empty[ab], (u - 1) & 1u);
        umma::mbar_wait_relaxed(&tab_free[ab], (u - 1) & 1u);
      }
      unsigned char* a_hi = smem + PQ_A1 + ab * (4 * PT_LBO_A);
      unsigned char* a_lo = a_hi + 2 * PT_LBO_A;
      *reinterpret_cast<uint4*>(a_hi + row * 16) = h0;
      *reinterpret_cast<uint4*>(a_lo + row * 16) = l0;
      *reinterpret_cast<uint4*>(a_hi + PT_LBO_A + row * 16) = h1;
      *reinterpret_cast<uint4*>(a_lo + PT_LBO_A + row * 16) = l1;
      if (MULTI) {
        float* t_sc = reinterpret_cast<float*>(smem + PQ_ROW) + ab * 4 * PT_TILE;
        t_sc[row] = sc;
        t_sc[PT_TILE + row] = sn;
        reinterpret_cast<int*>(t_sc)[2 * PT_TILE + row] = rc;
        reinterpret_cast<int*>(t_sc)[3 * PT_TILE + row] = rn;
      }
      // every thread releases its own stores (operand rows and score-table entries) with its
      // own arrival: the hand-off is then a plain per-thread release / acquire pair, which is
      // also what compute-sanitizer's racecheck can follow (an elected lane arriving for its
      // warp was reported as 4 hazards on the tables in round 1)
      umma::fence_smem_to_async();
      umma::mbar_arrive(&a1_full[ab]);
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace gn

#ifdef PP_TRACE
extern "C" int gn_pwfeat_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, gn::pp_trace, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int64_t gn_pwfeat_prep_bytes(void) { return (int64_t)gn::PT_IMG_BYTES; }

static int launch_pwfeat(bool x3, const float* dets, const float* scores, const int32_t* classes,
                                 const int32_t* pair_c, const int32_t* pair_n,
                                 const float* pair_iou, const int32_t* num_pairs, int capacity,
                                 int num_classes, float multiplier, const float* w1,
                                 const float* b1, const float* w2, const float* b2,
                                 const float* w3, const float* b3, int hidden, int out_dim,
                                 void* wprep, float* pw_out, void* pw_hl, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0 && num_classes >= 1, "gn_pwfeat_mlp_fwd: bad sizes");
  if (hidden != gn::PT_H || out_dim != gn::PT_O) {
    gn::set_error("gn_pwfeat_mlp_fwd: fused kernel is built for hidden=%d out=%d (got %d, %d)",
                  gn::PT_H, gn::PT_O, hidden, out_dim);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(dets && scores && pair_c && pair_n && pair_iou && num_pairs && (pw_out || pw_hl) &&
                 w1 && b1 && w2 && b2 && w3 && b3 && wprep,
             "gn_pwfeat_mlp_fwd: null pointer");
  GN_REQUIRE(num_classes == 1 || classes != nullptr, "gn_pwfeat_mlp_fwd: classes required");
  GN_REQUIRE((((uintptr_t)w1 | (uintptr_t)pw_out | (uintptr_t)pw_hl | (uintptr_t)dets |
               (uintptr_t)wprep) & 15) == 0,
             "gn_pwfeat_mlp_fwd: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* img = static_cast<unsigned char*>(wprep);
  gn::pwfeat_prepare_kernel<<<64, 256, 0, s>>>(w1, w2, w3, num_classes, img);
  GN_CHECK_LAUNCH("gn_pwfeat_mlp_fwd(prepare)");
  int grid = gn::ceil_div(capacity, gn::PT_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  cudaError_t e;
#define GN_PWFEAT_LAUNCH(MULTI_, X3_)                                                          \
  do {                                                                                         \
    e = cudaFuncSetAttribute(gn::pwfeat_pipe_kernel<MULTI_, X3_>,                              \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gn::PQ_BYTES);  \
    if (e == cudaSuccess)                                                                      \
      gn::pwfeat_pipe_kernel<MULTI_, X3_><<<grid, gn::PP_THREADS, gn::PQ_BYTES, s>>>(          \
          dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity, num_classes,   \
          multiplier, w1, b1, b2, b3, img, pw_out, static_cast<uint4*>(pw_hl));                \
  } while (0)
  if (num_classes > 1) {
    if (x3) GN_PWFEAT_LAUNCH(true, true); else GN_PWFEAT_LAUNCH(true, false);
  } else {
    if (x3) GN_PWFEAT_LAUNCH(false, true); else GN_PWFEAT_LAUNCH(false, false);
  }
#undef GN_PWFEAT_LAUNCH
  if (e != cudaSuccess) {
    gn::set_error("gn_pwfeat_mlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  GN_CHECK_LAUNCH("gn_pwfeat_mlp_fwd");
  return GN_OK;
}

extern "C" int gn_pwfeat_mlp_fwd(const float* dets, const float* scores, const int32_t* classes,
                                 const int32_t* pair_c, const int32_t* pair_n,
                                 const float* pair_iou, const int32_t* num_pairs, int capacity,
                                 int num_classes, float multiplier, const float* w1,
                                 const float* b1, const float* w2, const float* b2,
                                 const float* w3, const float* b3, int hidden, int out_dim,
                                 void* wprep, float* pw_out, gn_stream_t stream) {
  return launch_pwfeat(true, dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity,
                       num_classes, multiplier, w1, b1, w2, b2, w3, b3, hidden, out_dim, wprep,
                       pw_out, nullptr, stream);
}

extern "C" int gn_pwfeat_mlp_fwd_bf16(const float* dets, const float* scores,
                                      const int32_t* classes, const int32_t* pair_c,
                                      const int32_t* pair_n, const float* pair_iou,
                                      const int32_t* num_pairs, int capacity, int num_classes,
                                      float multiplier, const float* w1, const float* b1,
                                      const float* w2, const float* b2, const float* w3,
                                      const float* b3, int hidden, int out_dim, void* wprep,
                                      float* pw_out, gn_stream_t stream) {
  return launch_pwfeat(false, dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity,
                       num_classes, multiplier, w1, b1, w2, b2, w3, b3, hidden, out_dim, wprep,
                       pw_out, nullptr, stream);
}

// The same MLP with the output (also) written as bf16 (hi | lo) operand rows pw_hl[P, 64]:
// the format gn_block_pair_fwd_tma streams into its A tile.  pw_out may be null.
extern "C" int gn_pwfeat_mlp_fwd_hl(const float* dets, const float* scores, const int32_t* classes,
                                    const int32_t* pair_c, const int32_t* pair_n,
                                    const float* pair_iou, const int32_t* num_pairs, int capacity,
                                    int num_classes, float multiplier, const float* w1,
                                    const float* b1, const float* w2, const float* b2,
                                    const float* w3, const float* b3, int hidden, int out_dim,
                                    void* wprep, float* pw_out, void* pw_hl, int plain_bf16,
                                    gn_stream_t stream) {
  GN_REQUIRE(pw_hl != nullptr, "gn_pwfeat_mlp_fwd_hl: null pw_hl");
  return launch_pwfeat(plain_bf16 == 0, dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs,
                       capacity, num_classes, multiplier, w1, b1, w2, b2, w3, b3, hidden, out_dim,
                       wprep, pw_out, pw_hl, stream);
}
