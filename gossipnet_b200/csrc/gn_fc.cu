// Generic fully connected layer on CUDA cores: y = act(res + x @ W + b).
// Any k, n (the reference's shapes range from 9 x 256 to 256 x 256 and 128 x 1).
// This is the shape-agnostic path (detection-level layers, predict head,
// non-shipped pair-MLP shapes); the per-pair hot loops have their own fused
// kernels (gn_pairfeat.cu, gn_block.cu).
//
// 64 x 64 output tile per CTA, 256 threads, 4 x 4 register tile per thread,
// K in chunks of 16 through shared memory (A stored k-major so both operands
// are read with 128-bit shared loads).
#include "gn_common.cuh"

namespace gn {

constexpr int FC_BM = 64, FC_BN = 64, FC_BK = 16, FC_THREADS = 256, FC_PAD = 4;

__global__ void __launch_bounds__(FC_THREADS)
fc_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
              const float* __restrict__ bias, const float* __restrict__ res, int ld_res,
              int relu, float* __restrict__ y, int ldy, int rows_host,
              const int32_t* __restrict__ rows_dev, int k, int n) {
  __shared__ __align__(16) float As[FC_BK][FC_BM + FC_PAD];
  __shared__ __align__(16) float Bs[FC_BK][FC_BN + FC_PAD];
  int rows = rows_host;
  if (rows_dev != nullptr) rows = min(rows, __ldg(rows_dev));
  const int col0 = blockIdx.y * FC_BN;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;

  for (int row0 = blockIdx.x * FC_BM; row0 < rows; row0 += gridDim.x * FC_BM) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < k; k0 += FC_BK) {
      {
        const int kk = t & 15;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (t >> 4) + 16 * i;
          const int gr = row0 + r, gk = k0 + kk;
          As[kk][r] = (gr < rows && gk < k) ? __ldg(x + (size_t)gr * ldx + gk) : 0.f;
        }
        const int c = t & 63;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kb = (t >> 6) + 4 * i;
          const int gk = k0 + kb, gc = col0 + c;
          Bs[kb][c] = (gk < k && gc < n) ? __ldg(w + (size_t)gk * n + gc) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < FC_BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      }
      __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = row0 + ty * 4 + i;
      if (gr >= rows) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gc = col0 + tx * 4 + j;
        if (gc >= n) continue;
        float v = acc[i][j] + __ldg(bias + gc);
        if (res != nullptr) v += __ldg(res + (size_t)gr * ld_res + gc);
        if (relu) v = fmaxf(v, 0.f);
        y[(size_t)gr * ldy + gc] = v;
      }
    }
  }
}

}  // namespace gn

extern "C" int gn_fc_fwd(const float* x, int ldx, const float* w, const float* b,
                         const float* residual, int ld_res, int relu, float* y, int ldy,
                         int rows, const int32_t* rows_dev, int k, int n,
                         gn_stream_t stream) {
  GN_REQUIRE(rows >= 0 && k > 0 && n > 0, "gn_fc_fwd: bad shape rows=%d k=%d n=%d", rows, k, n);
  GN_REQUIRE(ldx >= k && ldy >= n, "gn_fc_fwd: leading dimension smaller than width");
  GN_REQUIRE(residual == nullptr || ld_res >= n, "gn_fc_fwd: residual leading dimension");
  if (rows == 0) return GN_OK;
  GN_REQUIRE(x && w && b && y, "gn_fc_fwd: null pointer");
  const int row_tiles = gn::ceil_div(rows, gn::FC_BM);
  const int col_tiles = gn::ceil_div(n, gn::FC_BN);
  const int max_x = 16 * gn::sm_count();
  dim3 grid(row_tiles < max_x ? row_tiles : max_x, col_tiles);
  gn::fc_fwd_kernel<<<grid, gn::FC_THREADS, 0, (cudaStream_t)stream>>>(
      x, ldx, w, b, residual, ld_res, relu, y, ldy, rows, rows_dev, k, n);
  GN_CHECK_LAUNCH("gn_fc_fwd");
  return GN_OK;
}

// ---------------------------------------------------------------------------------
// Predict head (A8, network.py:257-273).  Its hidden layers are LINEAR
// (activation_fn=None), so the chain x -> (x W1 + b1) -> (. W2 + b2) -> (. w3 + b3) is one
// affine map  x . w_eff + b_eff  with  w_eff = W1 (W2 w3),  b_eff = b3 + b2 . w3 + b1 . (W2 w3).
// gn_predict_collapse folds the chain (right to left, fp32, one small CTA; weights change
// every training step, so it runs once per forward), gn_rowdot_fwd applies it: one warp per
// detection row, the 512-byte row read once.  Same result as the staged FCs up to fp32
// re-association (~1e-6 relative, far inside the 1e-4 logit tolerance).
// table: n_layers x 4 int32 = (weight offset, bias offset, in, out) into flat_params; the
// last layer must have out == 1.  scratch: 2 * max(in) floats.
// ---------------------------------------------------------------------------------
namespace gn {

__global__ void predict_collapse_kernel(const float* __restrict__ flat,
                                        const int32_t* __restrict__ table, int n_layers,
                                        float* __restrict__ scratch, int max_dim,
                                        float* __restrict__ w_eff, float* __restrict__ b_eff) {
  __shared__ float bsum;
  const int t = threadIdx.x;
  const int32_t* last = table + (n_layers - 1) * 4;
  float* cur = scratch;
  float* nxt = scratch + max_dim;
  if (n_layers == 1) cur = w_eff;
  for (int i = t; i < last[2]; i += blockDim.x) cur[i] = flat[last[0] + i];   // [in, 1]
  if (t == 0) bsum = flat[last[1]];
  __syncthreads();
  for (int l = n_layers - 2; l >= 0; --l) {
    const int32_t* e = table + l * 4;
    const float* w = flat + e[0];
    const float* b = flat + e[1];
    const int in = e[2], out = e[3];
    if (l == 0) nxt = w_eff;
    for (int r = t; r < in; r += blockDim.x) {       // nxt = W[in,out] . cur[out]
      float acc = 0.f;
      for (int c = 0; c < out; ++c) acc = fmaf(__ldg(w + (size_t)r * out + c), cur[c], acc);
      nxt[r] = acc;
    }
    if (t == 0) {                                     // bias of layer l seen through the rest
      float acc = bsum;
      for (int c = 0; c < out; ++c) acc = fmaf(__ldg(b + c), cur[c], acc);
      bsum = acc;
    }
    __syncthreads();
    float* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  if (t == 0) *b_eff = bsum;
}

__global__ void __launch_bounds__(256)
rowdot_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                  const float* __restrict__ b, float* __restrict__ y, int rows, int k) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float bias = __ldg(b);
  for (int r = warp; r < rows; r += nwarps) {
    const float* xr = x + (size_t)r * ldx;
    float acc = 0.f;
    if ((k & 3) == 0 && (ldx & 3) == 0) {
      for (int c = lane * 4; c < k; c += 128) {
        const float4 a = ldg4(xr + c), ww = ldg4(w + c);
        acc = fmaf(a.x, ww.x, acc);
        acc = fmaf(a.y, ww.y, acc);
        acc = fmaf(a.z, ww.z, acc);
        acc = fmaf(a.w, ww.w, acc);
      }
    } else {
      for (int c = lane; c < k; c += 32) acc = fmaf(__ldg(xr + c), __ldg(w + c), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[r] = acc + bias;
  }
}

}  // namespace gn

extern "C" int gn_predict_collapse(const float* flat_params, const int32_t* table, int n_layers,
                                   int max_dim, float* scratch, float* w_eff, float* b_eff,
                                   gn_stream_t stream) {
  GN_REQUIRE(n_layers >= 1 && max_dim > 0, "gn_predict_collapse: bad sizes");
  GN_REQUIRE(flat_params && table && scratch && w_eff && b_eff, "gn_predict_collapse: null pointer");
  gn::predict_collapse_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(flat_params, table, n_layers,
                                                                  scratch, max_dim, w_eff, b_eff);
  GN_CHECK_LAUNCH("gn_predict_collapse");
  return GN_OK;
}

extern "C" int gn_rowdot_fwd(const float* x, int ldx, const float* w, const float* b, float* y,
                             int rows, int k, gn_stream_t stream) {
  GN_REQUIRE(rows >= 0 && k > 0 && ldx >= k, "gn_rowdot_fwd: bad shape rows=%d k=%d", rows, k);
  if (rows == 0) return GN_OK;
  GN_REQUIRE(x && w && b && y, "gn_rowdot_fwd: null pointer");
  GN_REQUIRE((((uintptr_t)x | (uintptr_t)w) & 15) == 0, "gn_rowdot_fwd: pointers must be 16-byte aligned");
  int grid = gn::ceil_div(rows, 8);
  const int cap = 8 * gn::sm_count();
  if (grid > cap) grid = cap;
  gn::rowdot_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, b, y, rows, k);
  GN_CHECK_LAUNCH("gn_rowdot_fwd");
  return GN_OK;
}
