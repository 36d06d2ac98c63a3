// Pair features on the tensor cores: geometry (A4) + the 3-layer pair-feature
// MLP (A5) width -> 256 -> 256 -> 32 with ReLU after every layer, fused.
//
// Tile = 128 pairs (the M of one tcgen05.mma), one persistent CTA per SM.
//   geometry : one thread per pair computes the 9 hand-crafted features in fp32
//              (pair_geometry, same rounding as the reference) and writes a K=16
//              bf16 hi/lo operand row;
//   layer 1  : [128 x 16] x [16 x 256]  (1 k-step x 3 UMMAs, N = 256) -> TMEM [0,256)
//              single class: K rows = [c_score, n_score, 7 geometry]; multi class:
//              the one-hot score block degenerates to two weight-row gathers that
//              the epilogue adds in fp32 (K rows = 7 geometry only);
//   layer 2  : [128 x 256] x [256 x 256] (16 k-steps x 3 UMMAs) -> TMEM [256,512).
//              W2 (256 KB as bf16 hi+lo) does not fit next to the activations, so it
//              is streamed per k-step (16 KB) from a pre-split operand image in
//              global memory (L2 resident) through a 4-stage cp.async.bulk ring
//              (mbarrier complete_tx / tcgen05.commit) that runs ahead of the UMMAs
//              across the epilogue phases;
//   layer 3  : [128 x 256] x [256 x 32] -> TMEM [0,32) (layer-1 columns are dead).
// The activation operand tile holds 128 of the 256 K columns at a time, so each of
// layers 2 and 3 runs as two (epilogue-half, 8 k-step) rounds.  Only pw_out[P,32]
// goes back to HBM; the [P,256] activations never leave the SM.
// bf16x3: every fp32 operand is split into bf16 hi + lo and each product is
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with fp32 accumulation (gn_umma.cuh).
#include "gn_pairfeat.cuh"
#include "gn_umma.cuh"

namespace gn {

constexpr int PT_TILE = 128;
constexpr int PT_THREADS = 512;   // 16 warps: 4 TMEM lane quadrants x 4 column groups
constexpr int PT_H = 256, PT_O = 32;
constexpr int PT_RING = 6;
constexpr uint32_t PT_SBO = 128;
constexpr uint32_t PT_LBO_A = PT_TILE * 16;      // activation / A1 chunk pitch (2048)
constexpr uint32_t PT_LBO_W = PT_H * 16;         // 256-row weight chunk pitch (4096)
constexpr uint32_t PT_LBO_W3 = PT_O * 16;        // 32-row weight chunk pitch (512)
constexpr uint32_t PT_STAGE = 2 * 2 * PT_LBO_W;  // one W2 k-step: (hi, lo) x 2 chunks = 16 KB

// prepared weight image (global workspace), bytes
// The W2 image can be replicated (CTA i streams copy i % COPIES).  Measured: 8 copies
// do not change the kernel time (profiles/r1_tc_kernels.md), so L2 line contention is
// not what limits the W2 stream; one copy is kept.
constexpr int PT_W2_COPIES = 1;
constexpr uint32_t PT_IMG_W2 = 0;                            // COPIES x 16 k-steps x 16 KB
constexpr uint32_t PT_IMG_B1 = PT_IMG_W2 + PT_W2_COPIES * 16 * PT_STAGE;   // hi 8 KB, lo 8 KB
constexpr uint32_t PT_IMG_B3 = PT_IMG_B1 + 2 * 2 * PT_LBO_W; // hi 16 KB, lo 16 KB
constexpr uint32_t PT_IMG_BYTES = PT_IMG_B3 + 2 * 32 * PT_LBO_W3;

// shared memory map
constexpr uint32_t PS_RING = 0;                                   // PT_RING x 16 KB
constexpr uint32_t PS_B1 = PS_RING + PT_RING * PT_STAGE;          // 16 KB
constexpr uint32_t PS_B3 = PS_B1 + 2 * 2 * PT_LBO_W;              // 32 KB
constexpr uint32_t PS_A1 = PS_B3 + 2 * 32 * PT_LBO_W3;            // hi 4 KB + lo 4 KB
constexpr uint32_t PS_BIAS = PS_A1 + 2 * 2 * PT_LBO_A;            // b1[256] b2[256] b3[32]
constexpr uint32_t PS_ROW = PS_BIAS + (2 * PT_H + PT_O) * 4;      // sc[128] sn[128] rc[128] rn[128]
constexpr uint32_t PS_BAR = PS_ROW + 4 * PT_TILE * 4;             // full[R] empty[R] done
constexpr uint32_t PS_BYTES = PS_BAR + 16 * 8;
static_assert(PS_BYTES <= 227 * 1024, "pair MLP tile exceeds shared memory");

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
  // B3: W3^T, 32 chunks x 32 rows
  for (int i = tid; i < PT_H * PT_O; i += nth) {
    const int k = i / PT_O, n = i - k * PT_O;
    unsigned char* st = img + PT_IMG_B3;
    put_split(st, st + 32 * PT_LBO_W3, (k >> 3) * PT_LBO_W3 + n * 16, k & 7, __ldg(w3 + i));
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

struct PtRing {
  uint64_t* full;
  uint64_t* empty;
  uint32_t ring_smem;
  const unsigned char* w2img;
  uint32_t loads_issued, total_loads;
  uint32_t k0;   // per-CTA rotation of the k-step order inside each K half (spreads the
                 // CTAs over the L2 lines of the W2 image instead of all hitting the same one)

  // issue every W2 k-step load up to (exclusive) index `upto`
  __device__ __forceinline__ void fill(uint32_t upto) {
    while (loads_issued < upto && loads_issued < total_loads) {
      const uint32_t i = loads_issued, s = i % PT_RING;
      if (i >= PT_RING) umma::mbar_wait(&empty[s], ((i / PT_RING) - 1) & 1);
      bulk_g2s(ring_smem + s * PT_STAGE, w2img + (size_t)((i & 8u) | (((i & 7u) + k0) & 7u)) * PT_STAGE, PT_STAGE, &full[s]);
      ++loads_issued;
    }
  }
};

template <bool MULTI>
__global__ void __launch_bounds__(PT_THREADS, 1)
pwfeat_tc_kernel(const float* __restrict__ dets, const float* __restrict__ scores,
                 const int32_t* __restrict__ classes, const int32_t* __restrict__ pair_c,
                 const int32_t* __restrict__ pair_n, const float* __restrict__ pair_iou,
                 const int32_t* __restrict__ num_pairs, int capacity, int num_classes, float mult,
                 const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ b2, const float* __restrict__ b3,
                 const unsigned char* __restrict__ img, float* __restrict__ pw_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + PT_TILE - 1) / PT_TILE;
  if ((int)blockIdx.x >= num_tiles) return;
  const int my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  unsigned char* a1_hi = smem + PS_A1;
  unsigned char* a1_lo = a1_hi + 2 * PT_LBO_A;
  float* bias1 = reinterpret_cast<float*>(smem + PS_BIAS);
  float* bias2 = bias1 + PT_H;
  float* bias3 = bias2 + PT_H;
  float* row_sc = reinterpret_cast<float*>(smem + PS_ROW);
  float* row_sn = row_sc + PT_TILE;
  int* row_rc = reinterpret_cast<int*>(row_sn + PT_TILE);
  int* row_rn = row_rc + PT_TILE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PS_BAR);
  uint64_t* empty = full + PT_RING;
  uint64_t* done = empty + PT_RING;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (t == 0) {
    for (int s = 0; s < PT_RING; ++s) {
      umma::mbar_init(&full[s], 1);
      umma::mbar_init(&empty[s], 1);
    }
    umma::mbar_init(done, 1);
    umma::fence_barrier_init();
  }
  // resident operands: B1, B3 (already split, copied verbatim), biases
  for (int i = t; i < (int)(2 * 2 * PT_LBO_W) / 16; i += PT_THREADS)
    reinterpret_cast<uint4*>(smem + PS_B1)[i] = __ldg(reinterpret_cast<const uint4*>(img + PT_IMG_B1) + i);
  for (int i = t; i < (int)(2 * 32 * PT_LBO_W3) / 16; i += PT_THREADS)
    reinterpret_cast<uint4*>(smem + PS_B3)[i] = __ldg(reinterpret_cast<const uint4*>(img + PT_IMG_B3) + i);
  for (int i = t; i < PT_H; i += PT_THREADS) {
    bias1[i] = __ldg(b1 + i);
    bias2[i] = __ldg(b2 + i);
  }
  if (t < PT_O) bias3[t] = __ldg(b3 + t);
  umma::fence_smem_to_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();

  const uint32_t tmem = tmem_base_s;
  // TMEM map: layer-1 accumulator of one 128-column half [0,128) (layer-3 accumulator
  // [0,32) later), activation operand hi [128,192) + lo [192,256) (128 K elements, two
  // bf16 per column), layer-2 accumulator [256,512)
  const uint32_t tm_l1 = tmem, tm_hh = tmem + 128, tm_hl = tmem + 192, tm_l2 = tmem + PT_H, tm_l3 = tmem;
  const uint32_t idesc128 = umma::idesc_bf16_f32(PT_TILE, 128);
  const uint32_t idesc256 = umma::idesc_bf16_f32(PT_TILE, PT_H);
  const uint32_t idesc32 = umma::idesc_bf16_f32(PT_TILE, PT_O);
  const uint32_t s_ring = umma::smem_u32(smem + PS_RING);
  const uint32_t s_a1h = umma::smem_u32(a1_hi), s_a1l = umma::smem_u32(a1_lo);
  const uint32_t s_b1h = umma::smem_u32(smem + PS_B1), s_b1l = s_b1h + 2 * PT_LBO_W;
  const uint32_t s_b3h = umma::smem_u32(smem + PS_B3), s_b3l = s_b3h + 32 * PT_LBO_W3;

  // kernel-lifetime operand descriptors; the issue loops only add start-address offsets
  const uint64_t d_a1h = umma::smem_desc(s_a1h, PT_LBO_A, PT_SBO), d_a1l = umma::smem_desc(s_a1l, PT_LBO_A, PT_SBO);
  const uint64_t d_b1h = umma::smem_desc(s_b1h, PT_LBO_W, PT_SBO), d_b1l = umma::smem_desc(s_b1l, PT_LBO_W, PT_SBO);
  const uint64_t d_ringh = umma::smem_desc(s_ring, PT_LBO_W, PT_SBO);
  const uint64_t d_ringl = umma::smem_desc(s_ring + 2 * PT_LBO_W, PT_LBO_W, PT_SBO);
  const uint64_t d_b3h = umma::smem_desc(s_b3h, PT_LBO_W3, PT_SBO), d_b3l = umma::smem_desc(s_b3l, PT_LBO_W3, PT_SBO);

  PtRing ring{full, empty, s_ring,
              img + PT_IMG_W2 + (size_t)(blockIdx.x % PT_W2_COPIES) * 16 * PT_STAGE, 0u, (uint32_t)my_tiles * 16u,
              0u};   // rotation off: results stay independent of the tile -> CTA mapping
  uint32_t mma_k = 0;     // W2 k-steps consumed so far (thread 0 only)
  uint32_t done_par = 0;  // parity of the next `done` completion (all threads)
  uint32_t load_target = PT_RING;   // W2 k-steps requested so far (producer thread only)
  if (t == 32) ring.fill(load_target);

  // epilogue mapping: TMEM lane quadrant = warp % 4; within a 128-column half the
  // warp owns columns [32 * (warp / 4), +32)
  const int erow = (warp & 3) * 32 + lane;
  const int ecol = (warp >> 2) * 32;
  const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
  // one half-epilogue: relu(acc[:, ecol..+32] + bias[col0 + ecol ..] (+ score rows)) -> the
  // activation operand in TENSOR MEMORY (bf16 hi / lo pairs; TS-form UMMAs read it)
  auto epilogue_to_h = [&](uint32_t tm_src, const float* bias, int col0, bool add_scores) {
    float va[32];
    umma::tmem_ld32(tm_src + tlane + ecol, va);
    umma::tmem_ld_wait();
    uint32_t hh[16], hl[16];
#pragma unroll
    for (int cc = 0; cc < 32; cc += 16) {
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = va[cc + e];
      const int col = col0 + ecol + cc;
      if (MULTI && add_scores) {
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
#pragma unroll
      for (int e = 0; e < 8; ++e)
        umma::split_bf16x2(fmaxf(v[2 * e] + bias[col + 2 * e], 0.f),
                           fmaxf(v[2 * e + 1] + bias[col + 2 * e + 1], 0.f),
                           hh[(cc >> 1) + e], hl[(cc >> 1) + e]);
    }
    const uint32_t c0 = (uint32_t)(ecol >> 1);
    umma::tmem_st8(tm_hh + tlane + c0, reinterpret_cast<const uint32_t(&)[8]>(hh[0]));
    umma::tmem_st8(tm_hh + tlane + c0 + 8, reinterpret_cast<const uint32_t(&)[8]>(hh[8]));
    umma::tmem_st8(tm_hl + tlane + c0, reinterpret_cast<const uint32_t(&)[8]>(hl[0]));
    umma::tmem_st8(tm_hl + tlane + c0 + 8, reinterpret_cast<const uint32_t(&)[8]>(hl[8]));
    umma::tmem_st_wait();
    umma::tc_fence_before();
    __syncthreads();
  };
  auto wait_done = [&]() {
    umma::mbar_wait(done, done_par);
    done_par ^= 1;
    umma::tc_fence_after();
  };

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int p0 = tile * PT_TILE;

    // ---- geometry -> A1 (K = 16): two threads per pair -------------------------------
    // role 0 (t < 128): iou and the three distances; role 1 (128 <= t < 256): the three
    // log ratios, the scores and the zero padding.  Each stores its own bf16 hi / lo
    // elements of the operand row (single class: [sc sn | iou xd yd l2 | wd hd] [ad 0..];
    // multi class: [iou xd yd l2 | wd hd ad 0] [0..]).
    if (t < 2 * PT_TILE) {
      const int row = t & (PT_TILE - 1), role = t >> 7;
      const int p = p0 + row;
      const bool live = p < P;
      int c = 0, n = 0;
      float4 cbx = make_float4(0.f, 0.f, 1.f, 1.f), nbx = cbx;
      if (live) {
        c = __ldg(pair_c + p);
        n = __ldg(pair_n + p);
        cbx = ldg4(dets + (size_t)c * 4);
        nbx = ldg4(dets + (size_t)n * 4);
      }
      __nv_bfloat16* hi0 = reinterpret_cast<__nv_bfloat16*>(a1_hi + row * 16);
      __nv_bfloat16* lo0 = reinterpret_cast<__nv_bfloat16*>(a1_lo + row * 16);
      auto put2 = [&](int e, float x0, float x1) {      // elements e, e+1 of chunk 0
        uint32_t h, l;
        umma::split_bf16x2(x0, x1, h, l);
        *reinterpret_cast<uint32_t*>(hi0 + e) = h;
        *reinterpret_cast<uint32_t*>(lo0 + e) = l;
      };
      if (role == 0) {
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) pair_geometry_dist(cbx, nbx, __ldg(pair_iou + p), mult, g);
        const int e0 = MULTI ? 0 : 2;
        put2(e0, g[0], g[1]);
        put2(e0 + 2, g[2], g[3]);
      } else {
        float g[3] = {0.f, 0.f, 0.f};
        float sc = 0.f, sn = 0.f;
        int rc = 0, rn = 0;
        if (live) {
          pair_geometry_logs(cbx, nbx, mult, g);
          sc = __fmul_rn(__ldg(scores + c), mult);
          sn = __fmul_rn(__ldg(scores + n), mult);
          if (MULTI) {
            rc = __ldg(classes + c) - 1;               // one-based classes (network.py:413-419)
            rn = num_classes + __ldg(classes + n) - 1;
          }
        }
        uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (MULTI) {
          row_sc[row] = sc; row_sn[row] = sn; row_rc[row] = rc; row_rn[row] = rn;
          put2(4, g[0], g[1]);
          put2(6, g[2], 0.f);
          *reinterpret_cast<uint4*>(a1_hi + PT_LBO_A + row * 16) = z;
          *reinterpret_cast<uint4*>(a1_lo + PT_LBO_A + row * 16) = z;
        } else {
          put2(0, sc, sn);
          put2(6, g[0], g[1]);
          uint32_t h, l;
          umma::split_bf16x2(g[2], 0.f, h, l);
          uint4 zh = z, zl = z;
          zh.x = h;
          zl.x = l;
          *reinterpret_cast<uint4*>(a1_hi + PT_LBO_A + row * 16) = zh;
          *reinterpret_cast<uint4*>(a1_lo + PT_LBO_A + row * 16) = zl;
        }
      }
    }
    umma::fence_smem_to_async();
    umma::tc_fence_before();
    __syncthreads();

    // ---- layers 1 + 2, per 128-column half of the hidden layer ---------------------------
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if (t == 0) {
        umma::tc_fence_after();
        // layer 1, output columns [128*half, +128): rows 128*half.. of the B1 tile
        umma::mma_bf16x3(tm_l1, d_a1h, d_a1l, d_b1h, d_b1l, 0, half * (128 * 16 >> 4), idesc128, 0);
        umma::mma_commit(done);
      }
      // UMMAs complete in order: this also covers the layer-2 UMMAs of the previous half,
      // so the activation operand may be overwritten
      wait_done();
      epilogue_to_h(tm_l1, bias1, half * 128, true);
      if (t == 0) {
        umma::tc_fence_after();
#pragma unroll 1
        for (int ksl = 0; ksl < 8; ++ksl) {
          const uint32_t g = mma_k++, s = g % PT_RING;
          umma::mbar_wait(&full[s], (g / PT_RING) & 1);
          umma::tc_fence_after();
          const uint32_t boff = s * (PT_STAGE >> 4);
          const uint32_t acc = (half | ksl) != 0;
          umma::mma_bf16_ts(tm_l2, tm_hl + ksl * 8, d_ringh + boff, idesc256, acc);
          umma::mma_bf16_ts(tm_l2, tm_hh + ksl * 8, d_ringl + boff, idesc256, 1);
          umma::mma_bf16_ts(tm_l2, tm_hh + ksl * 8, d_ringh + boff, idesc256, 1);
          umma::mma_commit(&empty[s]);   // slot s is free once these UMMAs have run
        }
        if (half == 1) umma::mma_commit(done);
      } else if (t == 32) {
        // W2 producer (its own thread, so the UMMA issuer never waits on a completion:
        // the commit -> mbarrier latency would otherwise sit on the issue path of every
        // k-step).  Refill every slot this round frees; the last wait resolves with the
        // round's last UMMA.
        load_target += 8;
        ring.fill(load_target);
      }
    }
    wait_done();   // layer 2 complete

    // ---- layer 3, two K halves ------------------------------------------------------
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      epilogue_to_h(tm_l2 + half * 128, bias2, half * 128, false);
      if (t == 0) {
        umma::tc_fence_after();
#pragma unroll
        for (int ksl = 0; ksl < 8; ++ksl) {
          const uint32_t boff = (uint32_t)((half * 8 + ksl) * 2) * (PT_LBO_W3 >> 4);
          const uint32_t acc = (half | ksl) != 0;
          umma::mma_bf16_ts(tm_l3, tm_hl + ksl * 8, d_b3h + boff, idesc32, acc);
          umma::mma_bf16_ts(tm_l3, tm_hh + ksl * 8, d_b3l + boff, idesc32, 1);
          umma::mma_bf16_ts(tm_l3, tm_hh + ksl * 8, d_b3h + boff, idesc32, 1);
        }
        umma::mma_commit(done);
      }
      wait_done();   // also frees the activation operand for the next half / tile
    }

    // ---- output: relu(acc + b3) -> pw_out[p, 32] ---------------------------------------
    if (warp < 4) {
      const int p = p0 + erow;
#pragma unroll
      for (int cc = 0; cc < PT_O; cc += 16) {
        float v[16];
        umma::tmem_ld16(tm_l3 + tlane + cc, v);
        umma::tmem_ld_wait();
        if (p < P) {
          float* dst = pw_out + (size_t)p * PT_O + cc;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(dst + g * 4) = make_float4(
                fmaxf(v[g * 4 + 0] + bias3[cc + g * 4 + 0], 0.f), fmaxf(v[g * 4 + 1] + bias3[cc + g * 4 + 1], 0.f),
                fmaxf(v[g * 4 + 2] + bias3[cc + g * 4 + 2], 0.f), fmaxf(v[g * 4 + 3] + bias3[cc + g * 4 + 3], 0.f));
        }
      }
    }
    umma::tc_fence_before();
    __syncthreads();   // TMEM [0,32) and the row tables are free for the next tile
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace gn

extern "C" int64_t gn_pwfeat_prep_bytes(void) { return (int64_t)gn::PT_IMG_BYTES; }

extern "C" int gn_pwfeat_mlp_fwd(const float* dets, const float* scores, const int32_t* classes,
                                 const int32_t* pair_c, const int32_t* pair_n,
                                 const float* pair_iou, const int32_t* num_pairs, int capacity,
                                 int num_classes, float multiplier, const float* w1,
                                 const float* b1, const float* w2, const float* b2,
                                 const float* w3, const float* b3, int hidden, int out_dim,
                                 void* wprep, float* pw_out, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0 && num_classes >= 1, "gn_pwfeat_mlp_fwd: bad sizes");
  if (hidden != gn::PT_H || out_dim != gn::PT_O) {
    gn::set_error("gn_pwfeat_mlp_fwd: fused kernel is built for hidden=%d out=%d (got %d, %d)",
                  gn::PT_H, gn::PT_O, hidden, out_dim);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(dets && scores && pair_c && pair_n && pair_iou && num_pairs && pw_out && w1 && b1 &&
                 w2 && b2 && w3 && b3 && wprep,
             "gn_pwfeat_mlp_fwd: null pointer");
  GN_REQUIRE(num_classes == 1 || classes != nullptr, "gn_pwfeat_mlp_fwd: classes required");
  GN_REQUIRE((((uintptr_t)w1 | (uintptr_t)pw_out | (uintptr_t)dets | (uintptr_t)wprep) & 15) == 0,
             "gn_pwfeat_mlp_fwd: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* img = static_cast<unsigned char*>(wprep);
  gn::pwfeat_prepare_kernel<<<64, 256, 0, s>>>(w1, w2, w3, num_classes, img);
  GN_CHECK_LAUNCH("gn_pwfeat_mlp_fwd(prepare)");
  int grid = gn::ceil_div(capacity, gn::PT_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  cudaError_t e;
  if (num_classes > 1) {
    e = cudaFuncSetAttribute(gn::pwfeat_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gn::PS_BYTES);
    if (e == cudaSuccess)
      gn::pwfeat_tc_kernel<true><<<grid, gn::PT_THREADS, gn::PS_BYTES, s>>>(
          dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity, num_classes,
          multiplier, w1, b1, b2, b3, img, pw_out);
  } else {
    e = cudaFuncSetAttribute(gn::pwfeat_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gn::PS_BYTES);
    if (e == cudaSuccess)
      gn::pwfeat_tc_kernel<false><<<grid, gn::PT_THREADS, gn::PS_BYTES, s>>>(
          dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity, num_classes,
          multiplier, w1, b1, b2, b3, img, pw_out);
  }
  if (e != cudaSuccess) {
    gn::set_error("gn_pwfeat_mlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  GN_CHECK_LAUNCH("gn_pwfeat_mlp_fwd");
  return GN_OK;
}
