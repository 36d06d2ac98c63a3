// Dense box IoU / IoA (A1 + A2).  HBM-store bound: 4 bytes out per pair, the
// 16-byte boxes are reused from registers / L1.
//
// Layout: a[batch,n,4], b[batch,m,4] -> out[batch,n,m].  One CTA owns a strip
// of ROWS_PER_CTA rows and all m columns of one image; each thread keeps the
// boxes of 4 consecutive columns in registers, walks the strip's rows (row box
// is a warp-uniform broadcast load) and emits one 16-byte streaming store per
// row, so a warp writes 512 contiguous bytes per instruction.
#include "gn_common.cuh"

namespace gn {

constexpr int IOU_THREADS = 256;
constexpr int IOU_ROWS_PER_CTA = 16;

template <bool VEC4>
__global__ void __launch_bounds__(IOU_THREADS)
iou_dense_kernel(const float* __restrict__ a, const float* __restrict__ b,
                 const uint8_t* __restrict__ crowd, const int32_t* __restrict__ a_cls,
                 const int32_t* __restrict__ b_cls, int n, int m, int strips_per_image,
                 float* __restrict__ out) {
  const int img = blockIdx.x / strips_per_image;
  const int strip = blockIdx.x - img * strips_per_image;
  const int row0 = strip * IOU_ROWS_PER_CTA;
  const int row1 = min(row0 + IOU_ROWS_PER_CTA, n);
  const float* ai = a + (size_t)img * n * 4;
  const float* bi = b + (size_t)img * m * 4;
  float* oi = out + (size_t)img * n * m;

  for (int c0 = threadIdx.x * 4; c0 < m; c0 += IOU_THREADS * 4) {
    Box cb[4];
    bool cr[4];
    int cc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = min(c0 + j, m - 1);
      cb[j] = make_box(ldg4(bi + (size_t)c * 4));
      cr[j] = crowd ? (crowd[(size_t)img * m + c] != 0) : false;
      cc[j] = b_cls ? b_cls[(size_t)img * m + c] : 0;
    }
    for (int r = row0; r < row1; ++r) {
      const Box rb = make_box(ldg4(ai + (size_t)r * 4));
      const int rc = a_cls ? a_cls[(size_t)img * n + r] : 0;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float inter = box_intersection(rb, cb[j]);
        const float uni = __fsub_rn(__fadd_rn(rb.area, cb[j].area), inter);
        // crowd column: intersection over the DETECTION's area (network.py:485-488)
        const float den = cr[j] ? rb.area : uni;
        float q = __fdiv_rn(inter, den);
        if (a_cls && rc != cc[j]) q = 0.0f;  // network.py:177-187
        v[j] = q;
      }
      float* dst = oi + (size_t)r * m + c0;
      if (VEC4) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < m) __stcs(dst + j, v[j]);
      }
    }
  }
}

}  // namespace gn

extern "C" int gn_iou_dense(const float* a, const float* b, const uint8_t* crowd,
                            const int32_t* a_cls, const int32_t* b_cls, int batch,
                            int n, int m, float* out, gn_stream_t stream) {
  GN_REQUIRE(batch >= 0 && n >= 0 && m >= 0, "gn_iou_dense: negative size");
  GN_REQUIRE((a_cls == nullptr) == (b_cls == nullptr),
             "gn_iou_dense: a_cls and b_cls must both be given or both be null");
  if (batch == 0 || n == 0 || m == 0) return GN_OK;
  GN_REQUIRE(a && b && out, "gn_iou_dense: null pointer");
  GN_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0,
             "gn_iou_dense: box arrays must be 16-byte aligned");
  const int strips = gn::ceil_div(n, gn::IOU_ROWS_PER_CTA);
  const int64_t grid = (int64_t)strips * batch;
  GN_REQUIRE(grid < (1ll << 31), "gn_iou_dense: problem too large for one launch");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (m % 4 == 0) && (((uintptr_t)out & 15) == 0);
  if (vec)
    gn::iou_dense_kernel<true><<<(unsigned)grid, gn::IOU_THREADS, 0, s>>>(
        a, b, crowd, a_cls, b_cls, n, m, strips, out);
  else
    gn::iou_dense_kernel<false><<<(unsigned)grid, gn::IOU_THREADS, 0, s>>>(
        a, b, crowd, a_cls, b_cls, n, m, strips, out);
  GN_CHECK_LAUNCH("gn_iou_dense");
  return GN_OK;
}
