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
