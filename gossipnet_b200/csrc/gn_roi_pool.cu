// ROI max pooling forward / backward: the RoiPool /
// RoiPoolGrad ops of nms_net/roi_pooling_layer/roi_pooling_op.cc (op definitions
// :35-54, CPU forward :128-187, CPU backward :374-449 - the CPU kernels are the
// semantic reference; the reference's own CUDA forward stores the wrong element
// for batch index > 0, roi_pooling_op_gpu.cu:75-76).
//
// NHWC feature map, channels innermost: a thread owns 4 consecutive channels of
// one output (forward) or input (backward) position, so every global access is a
// coalesced 16-byte vector.  Bin geometry is evaluated with the reference's float
// arithmetic (explicitly rounded mul/div, roundf / floorf / ceilf) so the integer
// bin edges - and with them argmax - are bit-identical.
//
// Backward is a GATHER in the reference's summation order (rois ascending, then
// ph, pw ascending), which keeps the float sums bit-identical and deterministic
// (no atomics).  The reference tests every roi against every input element; here
// a CTA owns one (image, h, w) position, compacts the rois that contain it once
// (ordered ballot compaction, 256 rois per round) and its threads then visit only
// those rois for their channels.
#include "gn_common.cuh"

namespace gn {

struct RoiGeom {
  int batch, start_w, start_h, end_w, end_h;
  float bin_h, bin_w;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale,
                                                int pooled_h, int pooled_w) {
  RoiGeom g;
  g.batch = (int)__ldg(roi + 0);
  g.start_w = (int)roundf(__fmul_rn(__ldg(roi + 1), scale));
  g.start_h = (int)roundf(__fmul_rn(__ldg(roi + 2), scale));
  g.end_w = (int)roundf(__fmul_rn(__ldg(roi + 3), scale));
  g.end_h = (int)roundf(__fmul_rn(__ldg(roi + 4), scale));
  const int rw = max(g.end_w - g.start_w + 1, 1);   // malformed rois become 1x1
  const int rh = max(g.end_h - g.start_h + 1, 1);
  g.bin_h = __fdiv_rn((float)rh, (float)pooled_h);
  g.bin_w = __fdiv_rn((float)rw, (float)pooled_w);
  return g;
}

template <int VEC>
__global__ void __launch_bounds__(256)
roi_pool_fwd_kernel(const float* __restrict__ data, const float* __restrict__ rois, int num_rois,
                    int height, int width, int channels, int pooled_h, int pooled_w, float scale,
                    float* __restrict__ top, int32_t* __restrict__ argmax) {
  const int cv = channels / VEC;
  const int64_t total = (int64_t)num_rois * pooled_h * pooled_w * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i;
    const int c = (int)(n % cv) * VEC; n /= cv;
    const int pw = (int)(n % pooled_w); n /= pooled_w;
    const int ph = (int)(n % pooled_h); n /= pooled_h;
    const RoiGeom g = roi_geometry(rois + n * 5, scale, pooled_h, pooled_w);
    int hstart = (int)floorf(__fmul_rn((float)ph, g.bin_h));
    int wstart = (int)floorf(__fmul_rn((float)pw, g.bin_w));
    int hend = (int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h));
    int wend = (int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w));
    hstart = min(max(hstart + g.start_h, 0), height);
    hend = min(max(hend + g.start_h, 0), height);
    wstart = min(max(wstart + g.start_w, 0), width);
    wend = min(max(wend + g.start_w, 0), width);
    const bool is_empty = (hend <= hstart) || (wend <= wstart);
    float best[VEC];
    int idx[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      best[e] = is_empty ? 0.f : -3.402823466e+38f;   // -FLT_MAX
      idx[e] = -1;
    }
    const float* img = data + (size_t)g.batch * height * width * channels;
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) {
        const int base = (h * width + w) * channels + c;
        float v[VEC];
        if (VEC == 4) {
          const float4 q = ldg4(img + base);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
          v[0] = __ldg(img + base);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (v[e] > best[e]) {   // strict: the first maximum wins
            best[e] = v[e];
            idx[e] = base + e;
          }
      }
    const size_t o = (size_t)i * VEC;
    if (VEC == 4) {
      *reinterpret_cast<float4*>(top + o) = make_float4(best[0], best[1], best[2], best[3]);
      *reinterpret_cast<int4*>(argmax + o) = make_int4(idx[0], idx[1], idx[2], idx[3]);
    } else {
      top[o] = best[0];
      argmax[o] = idx[0];
    }
  }
}

constexpr int RPB_THREADS = 256;

template <int VEC>
__global__ void __launch_bounds__(RPB_THREADS)
roi_pool_bwd_kernel(const float* __restrict__ rois, int num_rois, const int32_t* __restrict__ argmax,
                    const float* __restrict__ top_diff, int height, int width, int channels,
                    int pooled_h, int pooled_w, float scale, float* __restrict__ bottom_diff) {
  __shared__ int hits[RPB_THREADS];
  __shared__ int warp_cnt[RPB_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int pos = blockIdx.x;                 // (image, h, w)
  const int w = pos % width;
  const int h = (pos / width) % height;
  const int n = pos / (width * height);
  const int cv = channels / VEC;
  const int ncv = (cv + RPB_THREADS - 1) / RPB_THREADS;   // channel vectors per thread
  // per-thread accumulators for up to 4 channel-vector slots (C <= 4096 with VEC = 4)
  float acc[4][VEC];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[s][e] = 0.f;

  for (int r0 = 0; r0 < num_rois; r0 += RPB_THREADS) {
    // ---- which of these 256 rois contain (n, h, w)?  ordered compaction ----------
    const int r = r0 + t;
    bool inside = false;
    if (r < num_rois) {
      const RoiGeom g = roi_geometry(rois + (size_t)r * 5, scale, pooled_h, pooled_w);
      inside = g.batch == n && w >= g.start_w && w <= g.end_w && h >= g.start_h && h <= g.end_h;
    }
    const unsigned m = __ballot_sync(0xffffffffu, inside);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < RPB_THREADS / 32; ++i) {
      if (i < warp) base += warp_cnt[i];
      total += warp_cnt[i];
    }
    if (inside) hits[base + __popc(m & ((1u << lane) - 1u))] = r;
    __syncthreads();

    // ---- accumulate, rois ascending, then ph, pw ascending ---------------------------
    for (int k = 0; k < total; ++k) {
      const int roi_n = hits[k];
      const RoiGeom g = roi_geometry(rois + (size_t)roi_n * 5, scale, pooled_h, pooled_w);
      int phstart = (int)floorf(__fdiv_rn((float)(h - g.start_h), g.bin_h));
      int phend = (int)ceilf(__fdiv_rn((float)(h - g.start_h + 1), g.bin_h));
      int pwstart = (int)floorf(__fdiv_rn((float)(w - g.start_w), g.bin_w));
      int pwend = (int)ceilf(__fdiv_rn((float)(w - g.start_w + 1), g.bin_w));
      phstart = min(max(phstart, 0), pooled_h);
      phend = min(max(phend, 0), pooled_h);
      pwstart = min(max(pwstart, 0), pooled_w);
      pwend = min(max(pwend, 0), pooled_w);
      const size_t off = (size_t)roi_n * pooled_h * pooled_w * channels;
      for (int ph = phstart; ph < phend; ++ph)
        for (int pw = pwstart; pw < pwend; ++pw) {
          const size_t o = off + (size_t)(ph * pooled_w + pw) * channels;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const int c = (s * RPB_THREADS + t) * VEC;
            if (s < ncv && c < channels) {
              const int want = (h * width + w) * channels + c;
              if (VEC == 4) {
                const int4 a = __ldg(reinterpret_cast<const int4*>(argmax + o + c));
                const float4 d = ldg4(top_diff + o + c);
                if (a.x == want + 0) acc[s][0] += d.x;
                if (a.y == want + 1) acc[s][1] += d.y;
                if (a.z == want + 2) acc[s][2] += d.z;
                if (a.w == want + 3) acc[s][3] += d.w;
              } else {
                if (__ldg(argmax + o + c) == want) acc[s][0] += __ldg(top_diff + o + c);
              }
            }
          }
        }
    }
    __syncthreads();   // hits[] is rewritten by the next round
  }
  float* dst = bottom_diff + (size_t)pos * channels;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int c = (s * RPB_THREADS + t) * VEC;
    if (s < ncv && c < channels) {
      if (VEC == 4) *reinterpret_cast<float4*>(dst + c) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
      else dst[c] = acc[s][0];
    }
  }
}

}  // namespace gn

extern "C" int gn_roi_pool_fwd(const float* bottom_data, int batch, int height, int width,
                               int channels, const float* bottom_rois, int num_rois,
                               int pooled_height, int pooled_width, float spatial_scale,
                               float* top_data, int32_t* argmax, gn_stream_t stream) {
  GN_REQUIRE(pooled_height >= 0, "Need pooled_height >= 0, got %d", pooled_height);   // :64-66
  GN_REQUIRE(pooled_width >= 0, "Need pooled_width >= 0, got %d", pooled_width);      // :71-73
  GN_REQUIRE(batch >= 0 && height >= 0 && width >= 0 && channels >= 0 && num_rois >= 0,
             "gn_roi_pool_fwd: negative size");
  const int64_t total = (int64_t)num_rois * pooled_height * pooled_width * channels;
  if (total == 0) return GN_OK;
  GN_REQUIRE(bottom_data && bottom_rois && top_data && argmax, "gn_roi_pool_fwd: null pointer");
  GN_REQUIRE((int64_t)height * width * channels < (1ll << 31),
             "gn_roi_pool_fwd: one image must have fewer than 2^31 elements (int32 argmax)");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = channels % 4 == 0 && (((uintptr_t)bottom_data | (uintptr_t)top_data |
                                          (uintptr_t)argmax) & 15) == 0;
  const int64_t work = vec ? total / 4 : total;
  int64_t blocks = gn::ceil_div64(work, 256);
  const int cap = 32 * gn::sm_count();
  const int grid = (int)(blocks < cap ? blocks : cap);
  if (vec)
    gn::roi_pool_fwd_kernel<4><<<grid, 256, 0, s>>>(bottom_data, bottom_rois, num_rois, height, width,
                                                    channels, pooled_height, pooled_width,
                                                    spatial_scale, top_data, argmax);
  else
    gn::roi_pool_fwd_kernel<1><<<grid, 256, 0, s>>>(bottom_data, bottom_rois, num_rois, height, width,
                                                    channels, pooled_height, pooled_width,
                                                    spatial_scale, top_data, argmax);
  GN_CHECK_LAUNCH("gn_roi_pool_fwd");
  return GN_OK;
}

extern "C" int gn_roi_pool_bwd(int batch, int height, int width, int channels,
                               const float* bottom_rois, int num_rois, const int32_t* argmax,
                               const float* top_diff, int pooled_height, int pooled_width,
                               float spatial_scale, float* bottom_diff, gn_stream_t stream) {
  GN_REQUIRE(pooled_height >= 0, "Need pooled_height >= 0, got %d", pooled_height);
  GN_REQUIRE(pooled_width >= 0, "Need pooled_width >= 0, got %d", pooled_width);
  GN_REQUIRE(batch >= 0 && height >= 0 && width >= 0 && channels >= 0 && num_rois >= 0,
             "gn_roi_pool_bwd: negative size");
  const int64_t positions = (int64_t)batch * height * width;
  if (positions == 0 || channels == 0) return GN_OK;
  GN_REQUIRE(bottom_diff && (num_rois == 0 || (bottom_rois && argmax && top_diff)),
             "gn_roi_pool_bwd: null pointer");
  GN_REQUIRE(positions < (1ll << 31), "gn_roi_pool_bwd: too many positions for one launch");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = channels % 4 == 0 && (((uintptr_t)top_diff | (uintptr_t)bottom_diff |
                                          (uintptr_t)argmax) & 15) == 0;
  const int per_thread = vec ? 4 : 1;
  GN_REQUIRE(channels <= 4 * gn::RPB_THREADS * per_thread,
             "gn_roi_pool_bwd: at most %d channels supported", 4 * gn::RPB_THREADS * per_thread);
  if (vec)
    gn::roi_pool_bwd_kernel<4><<<(unsigned)positions, gn::RPB_THREADS, 0, s>>>(
        bottom_rois, num_rois, argmax, top_diff, height, width, channels, pooled_height,
        pooled_width, spatial_scale, bottom_diff);
  else
    gn::roi_pool_bwd_kernel<1><<<(unsigned)positions, gn::RPB_THREADS, 0, s>>>(
        bottom_rois, num_rois, argmax, top_diff, height, width, channels, pooled_height,
        pooled_width, spatial_scale, bottom_diff);
  GN_CHECK_LAUNCH("gn_roi_pool_bwd");
  return GN_OK;
}

// ---------------------------------------------------------------------------------
// Detection boxes -> Fast R-CNN rois for the image-feature head (network.py:78-100,
// enlarge_windows + to_frcn_coords): every box grows by `padding` times its size on each
// side around its centre, batch index 0:
//   cx = (x1 + x2) / 2, nw2 = (x2 - x1) * (0.5 + padding)  ->  (0, cx - nw2, cy - nh2, cx + nw2, cy + nh2)
// float32 ops in the reference's order.
// ---------------------------------------------------------------------------------
namespace gn {
__global__ void frcn_boxes_kernel(const float* __restrict__ dets, int n, float half_plus_pad,
                                  float batch_index, float* __restrict__ rois) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 d = ldg4(dets + (size_t)i * 4);
  const float w = __fsub_rn(d.z, d.x), h = __fsub_rn(d.w, d.y);
  const float cx = __fdiv_rn(__fadd_rn(d.x, d.z), 2.0f), cy = __fdiv_rn(__fadd_rn(d.y, d.w), 2.0f);
  const float nw2 = __fmul_rn(w, half_plus_pad), nh2 = __fmul_rn(h, half_plus_pad);
  float* r = rois + (size_t)i * 5;
  r[0] = batch_index;
  r[1] = __fsub_rn(cx, nw2);
  r[2] = __fsub_rn(cy, nh2);
  r[3] = __fadd_rn(cx, nw2);
  r[4] = __fadd_rn(cy, nh2);
}
}  // namespace gn

extern "C" int gn_frcn_boxes(const float* dets, int num_dets, float padding, int batch_index,
                             float* rois, gn_stream_t stream) {
  GN_REQUIRE(num_dets >= 0, "gn_frcn_boxes: negative size");
  if (num_dets == 0) return GN_OK;
  GN_REQUIRE(dets && rois, "gn_frcn_boxes: null pointer");
  GN_REQUIRE(((uintptr_t)dets & 15) == 0, "gn_frcn_boxes: dets must be 16-byte aligned");
  gn::frcn_boxes_kernel<<<gn::ceil_div(num_dets, 256), 256, 0, (cudaStream_t)stream>>>(
      dets, num_dets, (float)(0.5 + (double)padding), (float)batch_index, rois);
  GN_CHECK_LAUNCH("gn_frcn_boxes");
  return GN_OK;
}
