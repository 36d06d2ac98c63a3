// Pair features: raw geometry (A4) and the fused geometry + 3-layer MLP (A5).
//
// Fused CUDA-core (FFMA) variant, kept as gn_pwfeat_mlp_fwd_ffma: the fp32
// cross-check of the tensor-core kernel in gn_pairfeat_tc.cu.  A persistent CTA
// takes tiles of 64 pairs;
// the geometry of the tile is computed in registers, layer 1 (K = 9 effective
// inputs: with one-hot class scores the 2C+7 wide first layer degenerates to two
// weight-row gathers scaled by the scores plus 7 geometry rows) is evaluated
// straight into a pair-major shared tile, layer 2 (256 x 256, 86 % of the
// flops) runs as an 8x8 register-tiled GEMM with W2 streamed through a
// cp.async double buffer, and layer 3 (256 x 32) finishes from shared memory.
// Only pw_out[P,32] goes back to HBM; the [P,256] activations never leave the SM.
#include "gn_pairfeat.cuh"

namespace gn {

// ---------------------------------------------------------------------------
// raw features, one thread per pair (attribute / generic-shape path)
// ---------------------------------------------------------------------------
__global__ void pair_geometry_kernel(const float* __restrict__ dets,
                                     const float* __restrict__ scores,
                                     const int32_t* __restrict__ classes,
                                     const int32_t* __restrict__ pair_c,
                                     const int32_t* __restrict__ pair_n,
                                     const float* __restrict__ pair_iou,
                                     const int32_t* __restrict__ num_pairs, int capacity,
                                     int num_classes, float mult, float* __restrict__ out) {
  const int P = min(__ldg(num_pairs), capacity);
  const bool multi = num_classes > 1;
  const int width = multi ? 2 * num_classes + 7 : 9;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const int c = __ldg(pair_c + p), n = __ldg(pair_n + p);
    float g[7];
    pair_geometry(ldg4(dets + (size_t)c * 4), ldg4(dets + (size_t)n * 4), __ldg(pair_iou + p),
                  mult, g);
    float* o = out + (size_t)p * width;
    const float sc = __fmul_rn(__ldg(scores + c), mult), sn = __fmul_rn(__ldg(scores + n), mult);
    int gbase;
    if (multi) {
      for (int j = 0; j < 2 * num_classes; ++j) o[j] = 0.f;
      o[__ldg(classes + c) - 1] = sc;                 // scatter_nd one-hot rows,
      o[num_classes + __ldg(classes + n) - 1] = sn;   // classes are one-based
      gbase = 2 * num_classes;
    } else {
      o[0] = sc;
      o[1] = sn;
      gbase = 2;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) o[gbase + j] = g[j];
  }
}

// ---------------------------------------------------------------------------
// fused geometry + MLP (width -> 256 -> 256 -> 32)
// ---------------------------------------------------------------------------
constexpr int PW_TILE = 64;         // pairs per tile
constexpr int PW_H = 256;           // hidden width (cfg.gnet.pwfeat_dim)
constexpr int PW_O = 32;            // output width (cfg.gnet.pwfeat_narrow_dim)
constexpr int PW_THREADS = 256;
constexpr int PW_LDA = PW_H + 4;    // pair-major activation tile stride (floats)
constexpr int PW_KC = 16;           // W2 rows per cp.async stage

struct PwSmem {
  float a1[PW_TILE * PW_LDA];       // layer-1 output  [pair][256]
  float a2[PW_TILE * PW_LDA];       // layer-2 output  [pair][256]
  float w2s[2][PW_KC * PW_H];       // W2 stage double buffer [k][256]
  float w3s[PW_H * PW_O];           // W3 [256][32]
  float feat[PW_TILE][12];          // c_score, n_score, 7 geometry values
  int rows[PW_TILE][2];             // W1 row of the c / n score
};

__global__ void __launch_bounds__(PW_THREADS, 1)
pwfeat_mlp_kernel(const float* __restrict__ dets, const float* __restrict__ scores,
                  const int32_t* __restrict__ classes, const int32_t* __restrict__ pair_c,
                  const int32_t* __restrict__ pair_n, const float* __restrict__ pair_iou,
                  const int32_t* __restrict__ num_pairs, int capacity, int num_classes,
                  float mult, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w2, const float* __restrict__ b2,
                  const float* __restrict__ w3, const float* __restrict__ b3,
                  float* __restrict__ pw_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PwSmem& s = *reinterpret_cast<PwSmem*>(smem_raw);
  const int t = threadIdx.x;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + PW_TILE - 1) / PW_TILE;
  const bool multi = num_classes > 1;
  const int gbase = multi ? 2 * num_classes : 2;

  // W3 is tile-invariant: stage once per CTA
  for (int i = t; i < PW_H * PW_O / 4; i += PW_THREADS)
    reinterpret_cast<float4*>(s.w3s)[i] = __ldg(reinterpret_cast<const float4*>(w3) + i);

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int p0 = tile * PW_TILE;
    __syncthreads();  // previous tile fully consumed (a1/a2/feat reuse)

    // kick off the first W2 stage while the geometry is computed
    {
      const float4* src = reinterpret_cast<const float4*>(w2);
      float4* dst = reinterpret_cast<float4*>(s.w2s[0]);
#pragma unroll
      for (int i = 0; i < PW_KC * PW_H / 4 / PW_THREADS; ++i)
        cp_async16(dst + t + i * PW_THREADS, src + t + i * PW_THREADS);
      cp_async_commit();
    }

    // ---- geometry: one thread per pair -------------------------------------
    if (t < PW_TILE) {
      const int p = p0 + t;
      float g[7], sc = 0.f, sn = 0.f;
      int rc = 0, rn = 1;
      if (p < P) {
        const int c = __ldg(pair_c + p), n = __ldg(pair_n + p);
        pair_geometry(ldg4(dets + (size_t)c * 4), ldg4(dets + (size_t)n * 4),
                      __ldg(pair_iou + p), mult, g);
        sc = __fmul_rn(__ldg(scores + c), mult);
        sn = __fmul_rn(__ldg(scores + n), mult);
        if (multi) {
          rc = __ldg(classes + c) - 1;
          rn = num_classes + __ldg(classes + n) - 1;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 7; ++j) g[j] = 0.f;
      }
      s.feat[t][0] = sc;
      s.feat[t][1] = sn;
#pragma unroll
      for (int j = 0; j < 7; ++j) s.feat[t][2 + j] = g[j];
      s.rows[t][0] = rc;
      s.rows[t][1] = rn;
    }
    __syncthreads();

    // ---- layer 1: a1[m][:] = relu(b1 + sum_i feat[m][i] * W1[row_i(m)][:]) ---
    {
      const int j4 = t & 63;   // columns 4*j4 .. 4*j4+3
      const int mg = t >> 6;   // rows mg*16 .. mg*16+15 (warp-uniform)
      float4 wg[7];
#pragma unroll
      for (int i = 0; i < 7; ++i)
        wg[i] = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)(gbase + i) * PW_H) + j4);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b1) + j4);
#pragma unroll 4
      for (int mi = 0; mi < 16; ++mi) {
        const int m = mg * 16 + mi;
        const float4 wc = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)s.rows[m][0] * PW_H) + j4);
        const float4 wn = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)s.rows[m][1] * PW_H) + j4);
        float4 acc = bb;
        const float fc_ = s.feat[m][0], fn_ = s.feat[m][1];
        acc.x = fmaf(fc_, wc.x, acc.x); acc.y = fmaf(fc_, wc.y, acc.y);
        acc.z = fmaf(fc_, wc.z, acc.z); acc.w = fmaf(fc_, wc.w, acc.w);
        acc.x = fmaf(fn_, wn.x, acc.x); acc.y = fmaf(fn_, wn.y, acc.y);
        acc.z = fmaf(fn_, wn.z, acc.z); acc.w = fmaf(fn_, wn.w, acc.w);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const float f = s.feat[m][2 + i];
          acc.x = fmaf(f, wg[i].x, acc.x); acc.y = fmaf(f, wg[i].y, acc.y);
          acc.z = fmaf(f, wg[i].z, acc.z); acc.w = fmaf(f, wg[i].w, acc.w);
        }
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
        acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        *reinterpret_cast<float4*>(&s.a1[m * PW_LDA + j4 * 4]) = acc;
      }
    }
    // (the first __syncthreads inside the layer-2 loop orders a1 writes vs reads)

    // ---- layer 2: a2 = relu(a1 @ W2 + b2), 8 x 8 register tile per thread ----
    {
      const int tm = t & 7;    // rows tm + 8*i
      const int tn = t >> 3;   // columns tn*8 .. tn*8+7
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

      constexpr int NCHUNK = PW_H / PW_KC;
      for (int ch = 0; ch < NCHUNK; ++ch) {
        if (ch + 1 < NCHUNK) {
          const float4* src = reinterpret_cast<const float4*>(w2 + (size_t)(ch + 1) * PW_KC * PW_H);
          float4* dst = reinterpret_cast<float4*>(s.w2s[(ch + 1) & 1]);
#pragma unroll
          for (int i = 0; i < PW_KC * PW_H / 4 / PW_THREADS; ++i)
            cp_async16(dst + t + i * PW_THREADS, src + t + i * PW_THREADS);
          cp_async_commit();
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
        const float* wb = s.w2s[ch & 1];
#pragma unroll
        for (int k4 = 0; k4 < PW_KC; k4 += 4) {
          float4 av[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            av[i] = *reinterpret_cast<const float4*>(&s.a1[(tm + 8 * i) * PW_LDA + ch * PW_KC + k4]);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 b0 = *reinterpret_cast<const float4*>(&wb[(k4 + kk) * PW_H + tn * 8]);
            const float4 b1v = *reinterpret_cast<const float4*>(&wb[(k4 + kk) * PW_H + tn * 8 + 4]);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a, bv[j], acc[i][j]);
            }
          }
        }
        __syncthreads();  // stage (ch&1) free for the prefetch of chunk ch+2
      }
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(b2 + tn * 8));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(b2 + tn * 8 + 4));
      const float bias[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(acc[i][j] + bias[j], 0.f);
        float* dst = &s.a2[(tm + 8 * i) * PW_LDA + tn * 8];
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();

    // ---- layer 3: out = relu(a2 @ W3 + b3), 2 rows x 4 columns per thread ----
    {
      const int tn = t & 7;    // columns tn*4 .. +3
      const int tm = t >> 3;   // rows tm and tm + 32
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
      for (int k4 = 0; k4 < PW_H; k4 += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(&s.a2[tm * PW_LDA + k4]);
        const float4 a1v = *reinterpret_cast<const float4*>(&s.a2[(tm + 32) * PW_LDA + k4]);
        const float a0s[4] = {a0.x, a0.y, a0.z, a0.w};
        const float a1s[4] = {a1v.x, a1v.y, a1v.z, a1v.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 b = *reinterpret_cast<const float4*>(&s.w3s[(k4 + kk) * PW_O + tn * 4]);
          acc[0][0] = fmaf(a0s[kk], b.x, acc[0][0]); acc[0][1] = fmaf(a0s[kk], b.y, acc[0][1]);
          acc[0][2] = fmaf(a0s[kk], b.z, acc[0][2]); acc[0][3] = fmaf(a0s[kk], b.w, acc[0][3]);
          acc[1][0] = fmaf(a1s[kk], b.x, acc[1][0]); acc[1][1] = fmaf(a1s[kk], b.y, acc[1][1]);
          acc[1][2] = fmaf(a1s[kk], b.z, acc[1][2]); acc[1][3] = fmaf(a1s[kk], b.w, acc[1][3]);
        }
      }
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b3) + tn);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int p = p0 + tm + 32 * i;
        if (p < P) {
          float4 v;
          v.x = fmaxf(acc[i][0] + bb.x, 0.f); v.y = fmaxf(acc[i][1] + bb.y, 0.f);
          v.z = fmaxf(acc[i][2] + bb.z, 0.f); v.w = fmaxf(acc[i][3] + bb.w, 0.f);
          *reinterpret_cast<float4*>(pw_out + (size_t)p * PW_O + tn * 4) = v;
        }
      }
    }
  }
}

}  // namespace gn

extern "C" int gn_pair_geometry(const float* dets, const float* scores, const int32_t* classes,
                                const int32_t* pair_c, const int32_t* pair_n,
                                const float* pair_iou, const int32_t* num_pairs, int capacity,
                                int num_classes, float multiplier, float* out,
                                gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0 && num_classes >= 1, "gn_pair_geometry: bad sizes");
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(dets && scores && pair_c && pair_n && pair_iou && num_pairs && out,
             "gn_pair_geometry: null pointer");
  GN_REQUIRE(num_classes == 1 || classes != nullptr, "gn_pair_geometry: classes required");
  const int threads = 128;
  int grid = gn::ceil_div(capacity, threads);
  const int cap = 32 * gn::sm_count();
  if (grid > cap) grid = cap;
  gn::pair_geometry_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
      dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity, num_classes,
      multiplier, out);
  GN_CHECK_LAUNCH("gn_pair_geometry");
  return GN_OK;
}

extern "C" int gn_pwfeat_mlp_fwd_ffma(const float* dets, const float* scores, const int32_t* classes,
                                 const int32_t* pair_c, const int32_t* pair_n,
                                 const float* pair_iou, const int32_t* num_pairs, int capacity,
                                 int num_classes, float multiplier, const float* w1,
                                 const float* b1, const float* w2, const float* b2,
                                 const float* w3, const float* b3, int hidden, int out_dim,
                                 float* pw_out, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0 && num_classes >= 1, "gn_pwfeat_mlp_fwd_ffma: bad sizes");
  if (hidden != gn::PW_H || out_dim != gn::PW_O) {
    gn::set_error("gn_pwfeat_mlp_fwd_ffma: fused kernel is built for hidden=%d out=%d (got %d, %d)",
                  gn::PW_H, gn::PW_O, hidden, out_dim);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(dets && scores && pair_c && pair_n && pair_iou && num_pairs && pw_out && w1 && b1 &&
                 w2 && b2 && w3 && b3,
             "gn_pwfeat_mlp_fwd_ffma: null pointer");
  GN_REQUIRE(num_classes == 1 || classes != nullptr, "gn_pwfeat_mlp_fwd_ffma: classes required");
  GN_REQUIRE((((uintptr_t)w1 | (uintptr_t)b1 | (uintptr_t)w2 | (uintptr_t)b2 | (uintptr_t)w3 |
               (uintptr_t)b3 | (uintptr_t)pw_out | (uintptr_t)dets) & 15) == 0,
             "gn_pwfeat_mlp_fwd_ffma: pointers must be 16-byte aligned");
  static_assert(sizeof(gn::PwSmem) <= 227 * 1024, "pair MLP tile exceeds shared memory");
  const int smem = (int)sizeof(gn::PwSmem);
  cudaError_t e = cudaFuncSetAttribute(gn::pwfeat_mlp_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_pwfeat_mlp_fwd_ffma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(capacity, gn::PW_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  gn::pwfeat_mlp_kernel<<<grid, gn::PW_THREADS, smem, (cudaStream_t)stream>>>(
      dets, scores, classes, pair_c, pair_n, pair_iou, num_pairs, capacity, num_classes,
      multiplier, w1, b1, w2, b2, w3, b3, pw_out);
  GN_CHECK_LAUNCH("gn_pwfeat_mlp_fwd_ffma");
  return GN_OK;
}
