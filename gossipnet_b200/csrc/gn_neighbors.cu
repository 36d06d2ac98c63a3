// Neighbor build (A3): ordered compaction of {(c, n) : iou(c, n) >= thresh} into
// CSR form without materialising the dense N x N matrix.  IoU is recomputed from
// the boxes with the same rounded arithmetic as gn_iou_dense, so the lists are
// bit-identical to thresholding the dense matrix (tf.where order: row-major).
//
// One warp owns NB_ROWS_PER_WARP consecutive rows; lanes stride over the columns
// of that image, so each column box is loaded once per warp (coalesced 512 B)
// and reused from registers for all of the warp's rows.  Order inside a row is
// kept with ballot + popc prefix counts.
#include <type_traits>
#include "gn_common.cuh"

namespace gn {

constexpr int NB_THREADS = 256;
constexpr int NB_ROWS_PER_WARP = 4;
constexpr int NB_ROWS_PER_CTA = (NB_THREADS / 32) * NB_ROWS_PER_WARP;

template <bool FILL>
__global__ void __launch_bounds__(NB_THREADS)
neighbor_kernel(const float* __restrict__ dets, const int32_t* __restrict__ img_off,
                int num_images, int num_dets, float thresh,
                const int32_t* __restrict__ row_ptr, int capacity,
                int32_t* __restrict__ degree, int32_t* __restrict__ pair_c,
                int32_t* __restrict__ pair_n, float* __restrict__ pair_iou,
                int32_t* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * NB_ROWS_PER_CTA + warp * NB_ROWS_PER_WARP;
  if (row0 >= num_dets) return;

  // The warp's rows may straddle an image boundary: handle each row's own range.
  Box rb[NB_ROWS_PER_WARP];
  int lo[NB_ROWS_PER_WARP], hi[NB_ROWS_PER_WARP], cnt[NB_ROWS_PER_WARP], base[NB_ROWS_PER_WARP];
  int col_lo = 0x7fffffff, col_hi = 0;
#pragma unroll
  for (int i = 0; i < NB_ROWS_PER_WARP; ++i) {
    const int r = row0 + i;
    cnt[i] = 0;
    if (r < num_dets) {
      const int img = find_image(img_off, num_images, r);
      lo[i] = __ldg(img_off + img);
      hi[i] = __ldg(img_off + img + 1);
      rb[i] = make_box(ldg4(dets + (size_t)r * 4));
      base[i] = FILL ? __ldg(row_ptr + r) : 0;
      col_lo = min(col_lo, lo[i]);
      col_hi = max(col_hi, hi[i]);
    } else {
      lo[i] = hi[i] = 0;
      base[i] = 0;
      rb[i] = Box{0.f, 0.f, 0.f, 0.f, 0.f};
    }
  }

  for (int c0 = col_lo; c0 < col_hi; c0 += 32) {
    const int c = c0 + lane;
    const bool in_range = c < col_hi;
    const Box cb = make_box(in_range ? ldg4(dets + (size_t)c * 4) : make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
    for (int i = 0; i < NB_ROWS_PER_WARP; ++i) {
      const float v = box_iou(rb[i], cb);
      const bool hit = in_range && c >= lo[i] && c < hi[i] && (v >= thresh);
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (FILL) {
        if (hit) {
          const int pos = base[i] + cnt[i] + __popc(mask & ((1u << lane) - 1u));
          if (pos < capacity) {
            pair_c[pos] = row0 + i;
            pair_n[pos] = c;
            pair_iou[pos] = v;
          }
        }
      }
      cnt[i] += __popc(mask);
    }
  }

  if (!FILL) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NB_ROWS_PER_WARP; ++i)
        if (row0 + i < num_dets) degree[row0 + i] = cnt[i];
    }
  } else if (lane == 0 && overflow != nullptr) {
#pragma unroll
    for (int i = 0; i < NB_ROWS_PER_WARP; ++i)
      if (row0 + i < num_dets && base[i] + cnt[i] > capacity) *overflow = 1;
  }
}


// ---------------------------------------------------------------------------------
// Mask variant (the shipped path): the count pass decides  iou >= thresh  WITHOUT the IEEE
// division for all but borderline pairs and leaves one 32-bit hit mask per (row, 32
// columns of its image); the fill pass touches only the hits (P << N^2) and computes the
// exact IoU value there.  The decision is bit-identical to thresholding fl(inter / union):
// with t = thresh * union (one rounding, 2^-24) and d = 2^-20,
//     inter >= t (1 + d)  =>  inter/union > thresh (1 + 2^-21)  =>  fl(inter/union) >= thresh
//     inter <= t (1 - d)  =>  inter/union < thresh (1 - 2^-21)  =>  fl(inter/union) <  thresh
// (rounding is monotonic and both bounds are several ulps away from thresh); the band in
// between takes the exact division.  Mask word w of row r covers columns lo + 32 w .. + 31
// of r's own image [lo, hi); `stride` words per row (>= ceil(largest image / 32)).
// ---------------------------------------------------------------------------------
// image that owns detection row r, found by the whole warp at once: the number of images
// that end at or before r (one round of independent loads per 32 images, where the binary
// search of find_image is a chain of dependent ones - the prologue of a warp that only owns
// four rows must not cost as much as its main loop)
__device__ __forceinline__ int find_image_warp(const int32_t* __restrict__ off, int num_images,
                                               int r, int lane) {
  int count = 0;
  for (int base = 0; base < num_images; base += 32) {
    const int idx = base + lane;
    const bool ended = idx < num_images && __ldg(off + idx + 1) <= r;
    count += __popc(__ballot_sync(0xffffffffu, ended));
  }
  return min(count, num_images - 1);
}

// EXACT = false: the division-free decision; `border` collects (per lane, no vote in the hot
// loop) whether any pair fell into the band where only the exact test decides.
// EXACT = true : fl(inter / union) >= thresh, the reference's own test.
template <bool EXACT>
__device__ __forceinline__ bool iou_at_least(const Box& a, const Box& b, float thresh, bool& border) {
  if (EXACT) return box_iou(a, b) >= thresh;
  const float inter = box_intersection(a, b);
  const float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
  const float t = __fmul_rn(thresh, uni);
  const float hi = __fmul_rn(t, 1.0f + 9.5367431640625e-07f);   // 1 + 2^-20
  const float lo = __fmul_rn(t, 1.0f - 9.5367431640625e-07f);
  border |= !(uni > 0.f) || (inter > lo && inter < hi);
  return inter >= hi;
}

__global__ void __launch_bounds__(NB_THREADS)
neighbor_mask_kernel(const float* __restrict__ dets, const int32_t* __restrict__ img_off,
                     int num_images, int num_dets, float thresh, int stride,
                     int32_t* __restrict__ degree, uint32_t* __restrict__ masks) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * NB_ROWS_PER_CTA + warp * NB_ROWS_PER_WARP;
  if (row0 >= num_dets) return;
  Box rb[NB_ROWS_PER_WARP];
  int lo[NB_ROWS_PER_WARP], hi[NB_ROWS_PER_WARP], cnt[NB_ROWS_PER_WARP];
#pragma unroll
  for (int i = 0; i < NB_ROWS_PER_WARP; ++i) {
    const int r = row0 + i;
    cnt[i] = 0;
    lo[i] = hi[i] = -1;
    rb[i] = Box{0.f, 0.f, 0.f, 0.f, 0.f};
    const int img = find_image_warp(img_off, num_images, min(r, num_dets - 1), lane);
    if (r < num_dets) {
      lo[i] = __ldg(img_off + img);
      hi[i] = __ldg(img_off + img + 1);
      rb[i] = make_box(ldg4(dets + (size_t)r * 4));
    }
  }
  // One pass over the warp's rows with the division-free test; in the (rare) case that some
  // pair of the warp landed in the borderline band - or thresh <= 0 - the warp repeats its rows
  // with the exact test and overwrites its masks and counts.
  bool border = !(thresh > 0.f);
  auto walk = [&](auto exact_tag) {
    constexpr bool EXACT = decltype(exact_tag)::value;
#pragma unroll
    for (int i = 0; i < NB_ROWS_PER_WARP; ++i) cnt[i] = 0;
    // rows of one image are contiguous: walk the (at most 4) images the warp's rows touch, so
    // every row's mask words are aligned to its own image start.  All array indices are
    // compile-time constants (dynamic ones would send lo / hi / rb to local memory).
  #pragma unroll
    for (int s0 = 0; s0 < NB_ROWS_PER_WARP; ++s0) {
      const bool starts = lo[s0] >= 0 && (s0 == 0 || lo[s0] != lo[s0 > 0 ? s0 - 1 : 0]);
      if (!starts) continue;                          // warp uniform
      const int seg_lo = lo[s0], seg_hi = hi[s0];
      // lane (w % 32) keeps the mask of iteration w; every 32 iterations (and at the end) the
      // warp writes 32 words of each row with one coalesced store
      unsigned acc[NB_ROWS_PER_WARP] = {0u, 0u, 0u, 0u};
      int w = 0;
      for (int c0 = seg_lo; c0 < seg_hi; c0 += 32, ++w) {
        const int c = c0 + lane;
        const bool in_range = c < seg_hi;
        const Box cb = make_box(in_range ? ldg4(dets + (size_t)c * 4) : make_float4(0.f, 0.f, 1.f, 1.f));
  #pragma unroll
        for (int i = s0; i < NB_ROWS_PER_WARP; ++i) {
          if (lo[i] != seg_lo) continue;              // warp uniform
          const bool over = iou_at_least<EXACT>(rb[i], cb, thresh, border);
          const unsigned mask = __ballot_sync(0xffffffffu, in_range && over);
          if ((w & 31) == lane) acc[i] = mask;
          cnt[i] += __popc(mask);
        }
        if ((w & 31) == 31) {
  #pragma unroll
          for (int i = s0; i < NB_ROWS_PER_WARP; ++i)
            if (lo[i] == seg_lo && (w - 31 + lane) < stride)
              masks[(size_t)(row0 + i) * stride + (w - 31) + lane] = acc[i];
        }
      }
      if ((w & 31) != 0) {
        const int wbase = w & ~31;
  #pragma unroll
        for (int i = s0; i < NB_ROWS_PER_WARP; ++i)
          if (lo[i] == seg_lo && lane < (w & 31) && wbase + lane < stride)
            masks[(size_t)(row0 + i) * stride + wbase + lane] = acc[i];
      }
    }
  };
  walk(std::false_type{});
  if (__any_sync(0xffffffffu, border)) walk(std::true_type{});
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NB_ROWS_PER_WARP; ++i)
      if (row0 + i < num_dets) degree[row0 + i] = cnt[i];
  }
}

// one warp per row: lane l takes mask words l, l + 32, ...; ordered positions from a warp
// prefix sum of the popcounts; the exact IoU only for the set bits
__global__ void __launch_bounds__(NB_THREADS)
neighbor_fill_mask_kernel(const float* __restrict__ dets, const int32_t* __restrict__ img_off,
                          int num_images, int num_dets, const int32_t* __restrict__ row_ptr,
                          int capacity, const uint32_t* __restrict__ masks, int stride,
                          int32_t* __restrict__ pair_c, int32_t* __restrict__ pair_n,
                          float* __restrict__ pair_iou) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (NB_THREADS / 32) + (threadIdx.x >> 5);
  if (r >= num_dets) return;
  const int img = find_image_warp(img_off, num_images, r, lane);
  const int lo = __ldg(img_off + img), hi = __ldg(img_off + img + 1);
  const Box rb = make_box(ldg4(dets + (size_t)r * 4));
  int base = __ldg(row_ptr + r);
  const int words = min((hi - lo + 31) >> 5, stride);
  for (int w0 = 0; w0 < words; w0 += 32) {
    const int w = w0 + lane;
    unsigned m = w < words ? __ldg(masks + (size_t)r * stride + w) : 0u;
    const int n = __popc(m);
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    int pos = base + incl - n;
    base += __shfl_sync(0xffffffffu, incl, 31);
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int c = lo + (w << 5) + b;
      if (pos < capacity) {
        pair_c[pos] = r;
        pair_n[pos] = c;
        pair_iou[pos] = box_iou(rb, make_box(ldg4(dets + (size_t)c * 4)));
      }
      ++pos;
    }
  }
}

// Single-CTA exclusive scan with a running carry: out[0..n] (out[n] = total).
// n is at most a few hundred thousand detections and the data is L2 resident, so the scan is
// bound by latency and by how the one CTA touches memory.  A round covers 32 768 elements as
// eight segments of 4 096; thread t owns elements [4t, 4t + 4) of every segment, so each load /
// store of a round is ONE coalesced 16-byte access per thread and the eight loads are in flight
// together (a thread that owns 16 consecutive elements stores with a 64-byte stride between
// lanes: 32 sectors per instruction, ~9 us per round).  The segment scans share one pair of
// barriers; the 64 k degrees of the bench batch are two rounds (42 -> ~10 us).
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_SEGS = 8;
constexpr int SCAN_ROUND = SCAN_THREADS * 4 * SCAN_SEGS;

__global__ void __launch_bounds__(SCAN_THREADS)
exclusive_scan_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out) {
  __shared__ int warp_sums[SCAN_SEGS][SCAN_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  int carry = 0;                       // every thread keeps the same running total
  for (int base = 0; base < n; base += SCAN_ROUND) {
    int4 q[SCAN_SEGS];
    int incl[SCAN_SEGS], own[SCAN_SEGS];
#pragma unroll
    for (int j = 0; j < SCAN_SEGS; ++j) {
      const int i0 = base + (j * SCAN_THREADS + t) * 4;
      if (vec && i0 + 4 <= n) {
        q[j] = *reinterpret_cast<const int4*>(in + i0);
      } else {
        q[j].x = i0 < n ? in[i0] : 0;
        q[j].y = i0 + 1 < n ? in[i0 + 1] : 0;
        q[j].z = i0 + 2 < n ? in[i0 + 2] : 0;
        q[j].w = i0 + 3 < n ? in[i0 + 3] : 0;
      }
      own[j] = q[j].x + q[j].y + q[j].z + q[j].w;
      incl[j] = own[j];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
      for (int j = 0; j < SCAN_SEGS; ++j) {
        const int u = __shfl_up_sync(0xffffffffu, incl[j], d);
        if (lane >= d) incl[j] += u;
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int j = 0; j < SCAN_SEGS; ++j) warp_sums[j][warp] = incl[j];
    }
    __syncthreads();
    if (warp < SCAN_SEGS) {            // warp j scans the 32 warp totals of segment j
      int w = warp_sums[warp][lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += u;
      }
      warp_sums[warp][lane] = w;
    }
    __syncthreads();
    int seg_off = carry;
#pragma unroll
    for (int j = 0; j < SCAN_SEGS; ++j) {
      const int i0 = base + (j * SCAN_THREADS + t) * 4;
      int4 o;
      o.x = seg_off + (warp > 0 ? warp_sums[j][warp - 1] : 0) + incl[j] - own[j];
      o.y = o.x + q[j].x;
      o.z = o.y + q[j].y;
      o.w = o.z + q[j].z;
      if (vec && i0 + 4 <= n) {
        *reinterpret_cast<int4*>(out + i0) = o;
      } else {
        if (i0 < n) out[i0] = o.x;
        if (i0 + 1 < n) out[i0 + 1] = o.y;
        if (i0 + 2 < n) out[i0 + 2] = o.z;
        if (i0 + 3 < n) out[i0 + 3] = o.w;
      }
      seg_off += warp_sums[j][SCAN_THREADS / 32 - 1];
    }
    carry = seg_off;
    __syncthreads();                   // warp_sums are rewritten by the next round
  }
  if (t == 0) out[n] = carry;
}

}  // namespace gn

extern "C" int gn_neighbor_count(const float* dets, const int32_t* img_off, int num_images,
                                 int num_dets, float thresh, int32_t* degree,
                                 gn_stream_t stream) {
  GN_REQUIRE(num_images >= 0 && num_dets >= 0, "gn_neighbor_count: negative size");
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(dets && img_off && degree, "gn_neighbor_count: null pointer");
  GN_REQUIRE(num_images > 0, "gn_neighbor_count: detections without images");
  GN_REQUIRE(((uintptr_t)dets & 15) == 0, "gn_neighbor_count: dets must be 16-byte aligned");
  const int grid = gn::ceil_div(num_dets, gn::NB_ROWS_PER_CTA);
  gn::neighbor_kernel<false><<<grid, gn::NB_THREADS, 0, (cudaStream_t)stream>>>(
      dets, img_off, num_images, num_dets, thresh, nullptr, 0, degree, nullptr, nullptr,
      nullptr, nullptr);
  GN_CHECK_LAUNCH("gn_neighbor_count");
  return GN_OK;
}

extern "C" int gn_exclusive_scan(const int32_t* in, int n, int32_t* out, gn_stream_t stream) {
  GN_REQUIRE(n >= 0, "gn_exclusive_scan: negative size");
  GN_REQUIRE(out != nullptr && (in != nullptr || n == 0), "gn_exclusive_scan: null pointer");
  gn::exclusive_scan_kernel<<<1, gn::SCAN_THREADS, 0, (cudaStream_t)stream>>>(in, n, out);
  GN_CHECK_LAUNCH("gn_exclusive_scan");
  return GN_OK;
}

extern "C" int gn_neighbor_fill(const float* dets, const int32_t* img_off, int num_images,
                                int num_dets, float thresh, const int32_t* row_ptr,
                                int capacity, int32_t* pair_c, int32_t* pair_n,
                                float* pair_iou, int32_t* overflow, gn_stream_t stream) {
  GN_REQUIRE(num_images >= 0 && num_dets >= 0 && capacity >= 0, "gn_neighbor_fill: negative size");
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(dets && img_off && row_ptr && pair_c && pair_n && pair_iou,
             "gn_neighbor_fill: null pointer");
  GN_REQUIRE(num_images > 0, "gn_neighbor_fill: detections without images");
  const int grid = gn::ceil_div(num_dets, gn::NB_ROWS_PER_CTA);
  gn::neighbor_kernel<true><<<grid, gn::NB_THREADS, 0, (cudaStream_t)stream>>>(
      dets, img_off, num_images, num_dets, thresh, row_ptr, capacity, nullptr, pair_c, pair_n,
      pair_iou, overflow);
  GN_CHECK_LAUNCH("gn_neighbor_fill");
  return GN_OK;
}

extern "C" int gn_neighbor_count_masks(const float* dets, const int32_t* img_off, int num_images,
                                       int num_dets, float thresh, int stride_words,
                                       int32_t* degree, uint32_t* masks, gn_stream_t stream) {
  GN_REQUIRE(num_images >= 0 && num_dets >= 0 && stride_words > 0, "gn_neighbor_count_masks: bad size");
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(dets && img_off && degree && masks, "gn_neighbor_count_masks: null pointer");
  GN_REQUIRE(num_images > 0, "gn_neighbor_count_masks: detections without images");
  GN_REQUIRE(((uintptr_t)dets & 15) == 0, "gn_neighbor_count_masks: dets must be 16-byte aligned");
  const int grid = gn::ceil_div(num_dets, gn::NB_ROWS_PER_CTA);
  gn::neighbor_mask_kernel<<<grid, gn::NB_THREADS, 0, (cudaStream_t)stream>>>(
      dets, img_off, num_images, num_dets, thresh, stride_words, degree, masks);
  GN_CHECK_LAUNCH("gn_neighbor_count_masks");
  return GN_OK;
}

extern "C" int gn_neighbor_fill_masks(const float* dets, const int32_t* img_off, int num_images,
                                      int num_dets, const int32_t* row_ptr, int capacity,
                                      const uint32_t* masks, int stride_words, int32_t* pair_c,
                                      int32_t* pair_n, float* pair_iou, gn_stream_t stream) {
  GN_REQUIRE(num_images >= 0 && num_dets >= 0 && capacity >= 0 && stride_words > 0,
             "gn_neighbor_fill_masks: bad size");
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(dets && img_off && row_ptr && masks && pair_c && pair_n && pair_iou,
             "gn_neighbor_fill_masks: null pointer");
  GN_REQUIRE(num_images > 0, "gn_neighbor_fill_masks: detections without images");
  const int grid = gn::ceil_div(num_dets, gn::NB_THREADS / 32);
  gn::neighbor_fill_mask_kernel<<<grid, gn::NB_THREADS, 0, (cudaStream_t)stream>>>(
      dets, img_off, num_images, num_dets, row_ptr, capacity, masks, stride_words, pair_c, pair_n,
      pair_iou);
  GN_CHECK_LAUNCH("gn_neighbor_fill_masks");
  return GN_OK;
}
