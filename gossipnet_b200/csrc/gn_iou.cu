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
// The transposition goes through shared memory as 16-byte units in an XOR-swizzled 16 x 16
// block grid (conflict free both ways); the first version's scalar transposed stores were
// 4-way bank conflicted and made the kernel L1-bound (profiles/r2_iou.md).
// ---------------------------------------------------------------------------------
constexpr int SYM_T = 64;
constexpr int SYM_THREADS = 256;

__global__ void __launch_bounds__(SYM_THREADS)
iou_symmetric_kernel(const float* __restrict__ a, int n, float* __restrict__ out) {
  // 16 x 16 blocks of 4 x 4 elements; plane i holds row i of every block as one 16-byte unit
  __shared__ float4 blk[4][256];
  // The upper triangle folded into a rectangle: tile row p (T - p tiles) and tile row
  // T - 1 - p (p + 1 tiles) share grid row p of width T + 1, so no CTA is empty (but the
  // second half of the middle row when T is odd) and the decode is a compare and a subtract.
  const int img = blockIdx.z, T = gridDim.x - 1, p = blockIdx.y, k = blockIdx.x;
  int I, J;
  if (k < T - p) {
    I = p;
    J = p + k;
  } else {
    I = T - 1 - p;
    if (I == p) return;                 // middle row of an odd T: already covered above
    J = I + (k - (T - p));
  }
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
  // Transposed copy for tile (J, I): block (ty, tx) goes to unit ty * 16 + (tx ^ ty) of each
  // plane (16-byte stores, conflict free: a quarter-warp covers 8 distinct bank groups), the
  // thread then reads block (tx, ty) the same way and transposes the 4 x 4 in registers.
#pragma unroll
  for (int i = 0; i < 4; ++i)
    blk[i][ty * 16 + (tx ^ ty)] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  __syncthreads();
  float4 rr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) rr[i] = blk[i][tx * 16 + (ty ^ tx)];   // rows tx*4+i, cols ty*4.. of tile (I, J)
  const int tr0 = J * SYM_T + ty * 4, tc0 = I * SYM_T + tx * 4;       // this thread's block of tile (J, I)
  const float o[4][4] = {{rr[0].x, rr[1].x, rr[2].x, rr[3].x}, {rr[0].y, rr[1].y, rr[2].y, rr[3].y},
                         {rr[0].z, rr[1].z, rr[2].z, rr[3].z}, {rr[0].w, rr[1].w, rr[2].w, rr[3].w}};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = tr0 + j;
    if (r < n) {
      float* dst = oi + (size_t)r * n + tc0;
      if (vec && tc0 + 3 < n) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[j][0], o[j][1], o[j][2], o[j][3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (tc0 + i < n) dst[i] = o[j][i];
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
  // the same buffer on both sides = the det x det matrix: symmetric kernel (pass a copy of the
  // boxes to get the general kernel; tests/test_gpu_iou_neighbors.py compares the two)
  if (a == b && n == m && !has_crowd && !has_cls && n >= 2 * gn::SYM_T &&
      (((uintptr_t)out & 15) == 0)) {
    const int tiles = gn::ceil_div(n, gn::SYM_T);
    if (tiles <= 65535 && batch <= 65535) {
      gn::iou_symmetric_kernel<<<dim3(tiles + 1, (tiles + 1) / 2, batch), gn::SYM_THREADS, 0, s>>>(a, n, out);
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
