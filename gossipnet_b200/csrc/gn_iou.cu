// Dense box IoU / IoA (A1 + A2).  HBM-store bound: 4 bytes out per pair, the
// 16-byte boxes are reused from registers / shared memory.
//
// Layout: a[batch,n,4], b[batch,m,4] -> out[batch,n,m].  A CTA owns a tile of
// IOU_ROWS rows x (IOU_THREADS*4) columns of one image: the row boxes (+areas)
// are staged in shared memory once, each thread keeps the boxes of 4
// consecutive columns in registers, walks the rows and emits one 16-byte
// streaming store per row, so a warp writes 512 contiguous bytes per
// instruction.  The grid is (column tiles, row tiles, images): thousands of
// small CTAs, so the 148 SMs stay evenly loaded at any N.
//
// Division: most pairs are disjoint (inter == 0).  IEEE division with a zero
// numerator leaves the fast path of div.rn.f32, so for inter == 0 and a
// positive union the quotient (+0) is produced by a select instead; every other
// case (including zero / NaN unions of degenerate boxes) goes through the exact
// __fdiv_rn, so results stay bit-identical to the reference arithmetic.
#include <stdlib.h>

#include "gn_common.cuh"

namespace gn {

constexpr int IOU_THREADS = 128;
constexpr int IOU_ROWS = 32;
constexpr int IOU_COLS = IOU_THREADS * 4;

template <bool CROWD, bool CLS, bool VEC4>
__global__ void __launch_bounds__(IOU_THREADS)
iou_dense_kernel(const float* __restrict__ a, const float* __restrict__ b,
                 const uint8_t* __restrict__ crowd, const int32_t* __restrict__ a_cls,
                 const int32_t* __restrict__ b_cls, int batch, int n, int m,
                 float* __restrict__ out) {
  __shared__ float4 row_box[IOU_ROWS];
  __shared__ float row_area[IOU_ROWS];
  __shared__ int row_cls[IOU_ROWS];
  // persistent CTAs walk the (image, row strip, column chunk) items in memory order, so
  // the CTAs resident at any moment write one contiguous band of the output
  const int col_chunks = (m + IOU_COLS - 1) / IOU_COLS;
  const int strips = (n + IOU_ROWS - 1) / IOU_ROWS;
  const int64_t items = (int64_t)batch * strips * col_chunks;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
  const int chunk = (int)(item % col_chunks);
  const int strip = (int)((item / col_chunks) % strips);
  const int img = (int)(item / ((int64_t)col_chunks * strips));
  const int row0 = strip * IOU_ROWS;
  const int nrows = min(IOU_ROWS, n - row0);
  const float* ai = a + (size_t)img * n * 4;
  const float* bi = b + (size_t)img * m * 4;
  float* oi = out + (size_t)img * n * m;

  __syncthreads();   // previous item's row table fully consumed
  if (threadIdx.x < nrows) {
    const float4 v = ldg4(ai + (size_t)(row0 + threadIdx.x) * 4);
    row_box[threadIdx.x] = v;
    row_area[threadIdx.x] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
    if (CLS) row_cls[threadIdx.x] = a_cls[(size_t)img * n + row0 + threadIdx.x];
  }
  __syncthreads();

  const int c0 = chunk * IOU_COLS + threadIdx.x * 4;
  if (c0 >= m) continue;
  Box cb[4];
  bool cr[4];
  int cc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = min(c0 + j, m - 1);
    cb[j] = make_box(ldg4(bi + (size_t)c * 4));
    cr[j] = CROWD ? (crowd[(size_t)img * m + c] != 0) : false;
    cc[j] = CLS ? b_cls[(size_t)img * m + c] : 0;
  }
  float* dst = oi + (size_t)row0 * m + c0;
#pragma unroll 2
  for (int r = 0; r < nrows; ++r, dst += m) {
    const float4 rv = row_box[r];
    const float ra = row_area[r];
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float w = fmaxf(0.0f, __fsub_rn(fminf(rv.z, cb[j].x2), fmaxf(rv.x, cb[j].x1)));
      const float h = fmaxf(0.0f, __fsub_rn(fminf(rv.w, cb[j].y2), fmaxf(rv.y, cb[j].y1)));
      const float inter = __fmul_rn(w, h);
      const float uni = __fsub_rn(__fadd_rn(ra, cb[j].area), inter);
      // crowd column: intersection over the DETECTION's area (network.py:485-488)
      const float den = (CROWD && cr[j]) ? ra : uni;
      const bool zero = (inter == 0.0f) && (den > 0.0f);   // 0 / positive = +0, exactly
      float q = __fdiv_rn(zero ? den : inter, den);
      q = zero ? 0.0f : q;
      if (CLS && row_cls[r] != cc[j]) q = 0.0f;  // network.py:177-187
      v[j] = q;
    }
    if (VEC4) {
      *reinterpret_cast<float4*>(dst) = (make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < m) dst[j] = v[j];
    }
  }
  }
}

// ---------------------------------------------------------------------------------
// Symmetric variant for the det x det matrix (a == b, no crowd / class columns): the
// kernel is bound by instruction issue (~22 per element, mostly the exact division),
// not by the 4-byte store, and iou(i,j) == iou(j,i) bit for bit (min / max / + commute).
// A CTA computes one 64 x 64 tile of the upper triangle (tile row I <= tile column J),
// stores it, and stores its transpose into tile (J, I) through a shared-memory
// transposition so both stores are coalesced 256-byte row segments.
// ---------------------------------------------------------------------------------
constexpr int SYM_T = 64;
constexpr int SYM_THREADS = 256;

__global__ void __launch_bounds__(SYM_THREADS)
iou_symmetric_kernel(const float* __restrict__ a, int n, int tiles, float* __restrict__ out) {
  __shared__ float tr[SYM_T][SYM_T + 1];
  // linear upper-triangle tile index -> (I, J), I <= J
  const int img = blockIdx.y;
  int rem = blockIdx.x, I = 0;
  // row I of the triangle holds (tiles - I) tiles
  {
    // solve I from rem with a closed form, then fix up (tiles <= 4096 here)
    const float tf = (float)tiles + 0.5f;
    I = (int)(tf - sqrtf(tf * tf - 2.0f * (float)rem));
    if (I < 0) I = 0;
    while (I > 0 && (int64_t)I * tiles - (int64_t)I * (I - 1) / 2 > rem) --I;
    while ((int64_t)(I + 1) * tiles - (int64_t)(I + 1) * I / 2 <= rem) ++I;
    rem -= (int)((int64_t)I * tiles - (int64_t)I * (I - 1) / 2);
  }
  const int J = I + rem;
  const float* ai = a + (size_t)img * n * 4;
  float* oi = out + (size_t)img * n * n;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int r0 = I * SYM_T + ty * 4, c0 = J * SYM_T + tx * 4;

  Box rb[4], cb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rb[i] = make_box(ldg4(ai + (size_t)min(r0 + i, n - 1) * 4));
    cb[i] = make_box(ldg4(ai + (size_t)min(c0 + i, n - 1) * 4));
  }
  float v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = box_iou(rb[i], cb[j]);

  const bool vec = (n % 4 == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + i;
    if (r < n) {
      float* dst = oi + (size_t)r * n + c0;
      if (vec && c0 + 3 < n) {
        *reinterpret_cast<float4*>(dst) = (make_float4(v[i][0], v[i][1], v[i][2], v[i][3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < n) dst[j] = v[i][j];
      }
    }
  }
  if (I == J) return;   // diagonal tile: already complete (uniform per CTA)
  // transpose: tr[c][r] = v[r][c]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) tr[tx * 4 + j][ty * 4 + i] = v[i][j];
  __syncthreads();
  const int tr0 = J * SYM_T + ty * 4, tc0 = I * SYM_T + tx * 4;   // rows of tile (J, I)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr0 + i;
    if (r < n) {
      const float* src = &tr[ty * 4 + i][tx * 4];
      float* dst = oi + (size_t)r * n + tc0;
      if (vec && tc0 + 3 < n) {
        *reinterpret_cast<float4*>(dst) = (make_float4(src[0], src[1], src[2], src[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (tc0 + j < n) dst[j] = src[j];
      }
    }
  }
}

template <bool CROWD, bool CLS>
static void launch_iou(const float* a, const float* b, const uint8_t* crowd, const int32_t* a_cls,
                       const int32_t* b_cls, int batch, int n, int m, float* out, bool vec,
                       cudaStream_t s) {
  const int64_t items = (int64_t)batch * ceil_div(n, IOU_ROWS) * ceil_div(m, IOU_COLS);
  const int64_t cap = (int64_t)sm_count() * 16;     // 16 x 128 threads resident per SM
  const unsigned grid = (unsigned)(items < cap ? items : cap);
  if (vec)
    iou_dense_kernel<CROWD, CLS, true><<<grid, IOU_THREADS, 0, s>>>(a, b, crowd, a_cls, b_cls, batch, n, m, out);
  else
    iou_dense_kernel<CROWD, CLS, false><<<grid, IOU_THREADS, 0, s>>>(a, b, crowd, a_cls, b_cls, batch, n, m, out);
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
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (m % 4 == 0) && (((uintptr_t)out & 15) == 0);
  const bool has_crowd = crowd != nullptr, has_cls = a_cls != nullptr;
  static const int variant = getenv("GN_IOU_VARIANT") ? atoi(getenv("GN_IOU_VARIANT")) : 0;
  if (variant != 1 && a == b && n == m && !has_crowd && !has_cls && n >= 2 * gn::SYM_T &&
      (((uintptr_t)out & 15) == 0)) {
    const int tiles = gn::ceil_div(n, gn::SYM_T);
    const int64_t tri = (int64_t)tiles * (tiles + 1) / 2;
    if (tri < (1ll << 31) && batch <= 65535) {
      gn::iou_symmetric_kernel<<<dim3((unsigned)tri, batch), gn::SYM_THREADS, 0, s>>>(a, n, tiles, out);
      GN_CHECK_LAUNCH("gn_iou_dense(symmetric)");
      return GN_OK;
    }
  }
  if (has_crowd && has_cls) gn::launch_iou<true, true>(a, b, crowd, a_cls, b_cls, batch, n, m, out, vec, s);
  else if (has_crowd) gn::launch_iou<true, false>(a, b, crowd, a_cls, b_cls, batch, n, m, out, vec, s);
  else if (has_cls) gn::launch_iou<false, true>(a, b, crowd, a_cls, b_cls, batch, n, m, out, vec, s);
  else gn::launch_iou<false, false>(a, b, crowd, a_cls, b_cls, batch, n, m, out, vec, s);
  GN_CHECK_LAUNCH("gn_iou_dense");
  return GN_OK;
}
