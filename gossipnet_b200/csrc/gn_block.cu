// One Gnet block's pair stage (A7a, A7b and the two pair FCs of A7).
//
// Unfused pieces in the reference's formulation (parity cross-check and the
// path for non-shipped shapes): gather+concat, segment max.
// Fused CUDA-core (FFMA) variant, kept as gn_block_pair_fwd_ffma: the fp32
// cross-check of the tensor-core kernel in gn_block_tc.cu.  A persistent CTA takes tiles of 128
// consecutive pairs; pw_feats rows and the gathered detection features are
// cp.async'd straight into a pair-major [128][96] shared tile (self-pair
// neighbour halves and rows past P are zero-filled by the copy itself), both
// pair FCs run as register-tiled GEMMs out of shared memory, and the segmented
// max is finished with one integer atomicMax per (segment, tile, column):
// activations are >= 0 after the ReLU and every detection owns its self pair,
// so the max is exact and order independent.  Nothing of size P leaves the SM.
#include "gn_common.cuh"

namespace gn {

// ---------------------------------------------------------------------------
// unfused pieces
// ---------------------------------------------------------------------------
__global__ void gather_concat_kernel(const float* __restrict__ pw, int w,
                                     const float* __restrict__ feats,
                                     const float* __restrict__ nfeats, int r,
                                     const int32_t* __restrict__ pair_c,
                                     const int32_t* __restrict__ pair_n,
                                     const int32_t* __restrict__ num_pairs, int capacity,
                                     float* __restrict__ x) {
  const int P = min(__ldg(num_pairs), capacity);
  const int width = w + 2 * r;
  const int64_t total = (int64_t)P * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / width);
    const int j = (int)(i - (int64_t)p * width);
    float v;
    if (j < w) {
      v = __ldg(pw + (size_t)p * w + j);
    } else if (j < w + r) {
      v = __ldg(feats + (size_t)__ldg(pair_c + p) * r + (j - w));
    } else {
      const int c = __ldg(pair_c + p), n = __ldg(pair_n + p);
      v = (c == n) ? 0.f : __ldg(nfeats + (size_t)n * r + (j - w - r));
    }
    x[i] = v;
  }
}

// the same with one 16-byte chunk per thread (w, r multiples of 4, 16-byte aligned rows):
// four times fewer index computations and memory instructions - this op moves 160 MB per block
// of the training step
__global__ void gather_concat_vec4_kernel(const float4* __restrict__ pw, int w4,
                                          const float4* __restrict__ feats,
                                          const float4* __restrict__ nfeats, int r4,
                                          const int32_t* __restrict__ pair_c,
                                          const int32_t* __restrict__ pair_n,
                                          const int32_t* __restrict__ num_pairs, int capacity,
                                          float4* __restrict__ x) {
  const int P = min(__ldg(num_pairs), capacity);
  const int width4 = w4 + 2 * r4;
  const int64_t total = (int64_t)P * width4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / width4);
    const int j = (int)(i - (int64_t)p * width4);
    float4 v;
    if (j < w4) {
      v = __ldg(pw + (size_t)p * w4 + j);
    } else if (j < w4 + r4) {
      v = __ldg(feats + (size_t)__ldg(pair_c + p) * r4 + (j - w4));
    } else {
      const int c = __ldg(pair_c + p), n = __ldg(pair_n + p);
      v = (c == n) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(nfeats + (size_t)n * r4 + (j - w4 - r4));
    }
    x[i] = v;
  }
}

// (segment_max itself stays one column per thread: it is a serial walk over each detection's
// pairs, bound by latency, and a 16-byte version has four times fewer threads - measured 48 us
// against 27.)
__global__ void segment_max_kernel(const float* __restrict__ x, int f,
                                   const int32_t* __restrict__ row_ptr, int num_dets,
                                   float* __restrict__ out) {
  const int64_t total = (int64_t)num_dets * f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / f);
    const int j = (int)(i - (int64_t)row * f);
    const int b = __ldg(row_ptr + row), e = __ldg(row_ptr + row + 1);
    float m = 0.f;  // tf.segment_max yields 0 for an empty segment
    if (e > b) {
      m = __ldg(x + (size_t)b * f + j);
      for (int p = b + 1; p < e; ++p) m = fmaxf(m, __ldg(x + (size_t)p * f + j));
    }
    out[i] = m;
  }
}

// ---------------------------------------------------------------------------
// fused pair stage, shipped shape: w = 32, r = 32, f = 64
// ---------------------------------------------------------------------------
constexpr int BP_TILE = 128;
constexpr int BP_THREADS = 256;
constexpr int BP_W = 32, BP_R = 32, BP_F = 64;
constexpr int BP_K1 = BP_W + 2 * BP_R;     // 96
constexpr int BP_LDX = BP_K1 + 4;          // 100 floats: rows 16-byte aligned, bank-skewed
constexpr int BP_LDH = BP_F + 4;           // 68

struct BpSmem {
  float x[BP_TILE * BP_LDX];    // [pair][pw | c | n]; reused for the FC2 output
  float h1[BP_TILE * BP_LDH];   // FC1 output
  float w1[BP_K1 * BP_F];       // [k][64]
  float w2[BP_F * BP_F];
  int c[BP_TILE];
};

// acc[i][j] += A[tm + 16 i][k] * B[k][tn*4 + j]
template <int K, int LDA>
__device__ __forceinline__ void tile_gemm_128x64(const float* __restrict__ A,
                                                 const float* __restrict__ B, int tm, int tn,
                                                 float (&acc)[8][4]) {
#pragma unroll 2
  for (int k4 = 0; k4 < K; k4 += 4) {
    float4 av[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      av[i] = *reinterpret_cast<const float4*>(&A[(tm + 16 * i) * LDA + k4]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(&B[(k4 + kk) * BP_F + tn * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
        acc[i][0] = fmaf(a, b.x, acc[i][0]);
        acc[i][1] = fmaf(a, b.y, acc[i][1]);
        acc[i][2] = fmaf(a, b.z, acc[i][2]);
        acc[i][3] = fmaf(a, b.w, acc[i][3]);
      }
    }
  }
}

__global__ void __launch_bounds__(BP_THREADS, 1)
block_pair_fwd_kernel(const float* __restrict__ pw, const float* __restrict__ feats,
                      const float* __restrict__ nfeats, const int32_t* __restrict__ pair_c,
                      const int32_t* __restrict__ pair_n, const int32_t* __restrict__ num_pairs,
                      int capacity, const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2,
                      float* __restrict__ pooled) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BpSmem& s = *reinterpret_cast<BpSmem*>(smem_raw);
  const int t = threadIdx.x;
  const int P = min(__ldg(num_pairs), capacity);
  const int num_tiles = (P + BP_TILE - 1) / BP_TILE;
  const int tm = t & 15, tn = t >> 4;

  for (int i = t; i < BP_K1 * BP_F / 4; i += BP_THREADS)
    reinterpret_cast<float4*>(s.w1)[i] = __ldg(reinterpret_cast<const float4*>(w1) + i);
  for (int i = t; i < BP_F * BP_F / 4; i += BP_THREADS)
    reinterpret_cast<float4*>(s.w2)[i] = __ldg(reinterpret_cast<const float4*>(w2) + i);
  const float4 bias1 = __ldg(reinterpret_cast<const float4*>(b1) + tn);
  const float4 bias2 = __ldg(reinterpret_cast<const float4*>(b2) + tn);

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int p0 = tile * BP_TILE;
    __syncthreads();  // previous tile's x/h1/c fully consumed

    // ---- gather: 24 x 16-byte chunks per pair row ----------------------------
    constexpr int CHUNKS = BP_K1 / 4;  // 24
#pragma unroll
    for (int it = 0; it < BP_TILE * CHUNKS / BP_THREADS; ++it) {
      const int q = t + it * BP_THREADS;
      const int m = q / CHUNKS, ch = q - m * CHUNKS;
      const int p = p0 + m;
      const bool valid = p < P;
      const int pc = valid ? __ldg(pair_c + p) : 0;
      const float* src;
      bool keep = valid;
      if (ch < BP_W / 4) {
        src = pw + (size_t)(valid ? p : 0) * BP_W + ch * 4;
      } else if (ch < (BP_W + BP_R) / 4) {
        src = feats + (size_t)pc * BP_R + (ch - BP_W / 4) * 4;
      } else {
        const int pn = valid ? __ldg(pair_n + p) : 0;
        src = nfeats + (size_t)pn * BP_R + (ch - (BP_W + BP_R) / 4) * 4;
        keep = valid && (pn != pc);  // self pair: neighbour half is zero (network.py:372-374)
      }
      cp_async16_zfill(&s.x[m * BP_LDX + ch * 4], src, keep);
      if (ch == 0) s.c[m] = valid ? pc : -1;
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- FC1: h1 = relu(x @ W1 + b1) -----------------------------------------
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = bias1.x; acc[i][1] = bias1.y; acc[i][2] = bias1.z; acc[i][3] = bias1.w;
    }
    tile_gemm_128x64<BP_K1, BP_LDX>(s.x, s.w1, tm, tn, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(&s.h1[(tm + 16 * i) * BP_LDH + tn * 4]) =
          make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f),
                      fmaxf(acc[i][3], 0.f));
    __syncthreads();

    // ---- FC2: h2 = relu(h1 @ W2 + b2) -> reuse x as [pair][68] ----------------
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = bias2.x; acc[i][1] = bias2.y; acc[i][2] = bias2.z; acc[i][3] = bias2.w;
    }
    tile_gemm_128x64<BP_F, BP_LDH>(s.h1, s.w2, tm, tn, acc);
    float* h2 = s.x;  // all reads of x finished before the barrier above
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(&h2[(tm + 16 * i) * BP_LDH + tn * 4]) =
          make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f),
                      fmaxf(acc[i][3], 0.f));
    __syncthreads();

    // ---- segmented max over the tile's rows ----------------------------------
    {
      const int j = t & 63;            // column
      const int r0 = (t >> 6) * 32;    // 32-row slice
      int cur_c = -1;
      float cur = 0.f;
      for (int r = r0; r < r0 + 32; ++r) {
        const int c = s.c[r];
        if (c < 0) break;  // rows past P
        if (c != cur_c) {
          if (cur_c >= 0) atomicMax(reinterpret_cast<int*>(pooled + (size_t)cur_c * BP_F + j), __float_as_int(cur));
          cur_c = c;
          cur = 0.f;
        }
        cur = fmaxf(cur, h2[r * BP_LDH + j]);
      }
      if (cur_c >= 0) atomicMax(reinterpret_cast<int*>(pooled + (size_t)cur_c * BP_F + j), __float_as_int(cur));
    }
  }
}

}  // namespace gn

extern "C" int gn_block_gather_concat(const float* pw, int w, const float* feats,
                                      const float* nfeats, int r, const int32_t* pair_c,
                                      const int32_t* pair_n, const int32_t* num_pairs,
                                      int capacity, float* x, gn_stream_t stream) {
  GN_REQUIRE(w > 0 && r > 0 && capacity >= 0, "gn_block_gather_concat: bad sizes");
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(pw && feats && nfeats && pair_c && pair_n && num_pairs && x,
             "gn_block_gather_concat: null pointer");
  const int threads = 256;
  const int cap = 32 * gn::sm_count();
  if (w % 4 == 0 && r % 4 == 0 &&
      (((uintptr_t)pw | (uintptr_t)feats | (uintptr_t)nfeats | (uintptr_t)x) & 15) == 0) {
    const int64_t blocks = gn::ceil_div64((int64_t)capacity * ((w + 2 * r) / 4), threads);
    gn::gather_concat_vec4_kernel<<<(int)(blocks < cap ? blocks : cap), threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(pw), w / 4, reinterpret_cast<const float4*>(feats),
        reinterpret_cast<const float4*>(nfeats), r / 4, pair_c, pair_n, num_pairs, capacity,
        reinterpret_cast<float4*>(x));
    GN_CHECK_LAUNCH("gn_block_gather_concat");
    return GN_OK;
  }
  int64_t blocks = gn::ceil_div64((int64_t)capacity * (w + 2 * r), threads);
  const int grid = (int)(blocks < cap ? blocks : cap);
  gn::gather_concat_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
      pw, w, feats, nfeats, r, pair_c, pair_n, num_pairs, capacity, x);
  GN_CHECK_LAUNCH("gn_block_gather_concat");
  return GN_OK;
}

extern "C" int gn_segment_max(const float* x, int f, const int32_t* row_ptr, int num_dets,
                              float* out, gn_stream_t stream) {
  GN_REQUIRE(f > 0 && num_dets >= 0, "gn_segment_max: bad sizes");
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(x && row_ptr && out, "gn_segment_max: null pointer");
  const int threads = 256;
  const int cap = 32 * gn::sm_count();
  int64_t blocks = gn::ceil_div64((int64_t)num_dets * f, threads);
  const int grid = (int)(blocks < cap ? blocks : cap);
  gn::segment_max_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(x, f, row_ptr, num_dets, out);
  GN_CHECK_LAUNCH("gn_segment_max");
  return GN_OK;
}

extern "C" int gn_block_pair_fwd_ffma(const float* pw, int w, const float* feats,
                                 const float* nfeats, int r, const int32_t* pair_c,
                                 const int32_t* pair_n, const int32_t* num_pairs, int capacity,
                                 const float* w1, const float* b1, const float* w2,
                                 const float* b2, int f, float* pooled, gn_stream_t stream) {
  GN_REQUIRE(capacity >= 0, "gn_block_pair_fwd_ffma: negative capacity");
  if (w != gn::BP_W || r != gn::BP_R || f != gn::BP_F) {
    gn::set_error("gn_block_pair_fwd_ffma: fused kernel is built for w=%d r=%d f=%d (got %d, %d, %d)",
                  gn::BP_W, gn::BP_R, gn::BP_F, w, r, f);
    return GN_ERR_UNSUPPORTED;
  }
  if (capacity == 0) return GN_OK;
  GN_REQUIRE(pw && feats && nfeats && pair_c && pair_n && num_pairs && w1 && b1 && w2 && b2 &&
                 pooled,
             "gn_block_pair_fwd_ffma: null pointer");
  GN_REQUIRE((((uintptr_t)pw | (uintptr_t)feats | (uintptr_t)nfeats | (uintptr_t)w1 |
               (uintptr_t)b1 | (uintptr_t)w2 | (uintptr_t)b2) & 15) == 0,
             "gn_block_pair_fwd_ffma: pointers must be 16-byte aligned");
  static_assert(sizeof(gn::BpSmem) <= 227 * 1024, "block tile exceeds shared memory");
  const int smem = (int)sizeof(gn::BpSmem);
  cudaError_t e = cudaFuncSetAttribute(gn::block_pair_fwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    gn::set_error("gn_block_pair_fwd_ffma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return GN_ERR_CUDA;
  }
  int grid = gn::ceil_div(capacity, gn::BP_TILE);
  const int sms = gn::sm_count();
  if (grid > sms) grid = sms;
  gn::block_pair_fwd_kernel<<<grid, gn::BP_THREADS, smem, (cudaStream_t)stream>>>(
      pw, feats, nfeats, pair_c, pair_n, num_pairs, capacity, w1, b1, w2, b2, pooled);
  GN_CHECK_LAUNCH("gn_block_pair_fwd_ffma");
  return GN_OK;
}
