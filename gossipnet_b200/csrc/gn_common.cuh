// Shared device/host helpers for the gossipnet_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gossipnet_b200.h"

namespace gn {

void set_error(const char* fmt, ...);

#define GN_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      gn::set_error(__VA_ARGS__);            \
      return GN_ERR_INVALID_ARGUMENT;        \
    }                                        \
  } while (0)

#define GN_CHECK_LAUNCH(name)                                              \
  do {                                                                     \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) {                                               \
      gn::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(_e)); \
      return GN_ERR_CUDA;                                                  \
    }                                                                      \
  } while (0)

int sm_count();

// gn_set_pdl (gossipnet_b200.h): launch the persistent block kernels with programmatic stream
// serialization, so that each one's prologue overlaps the tail of its predecessor
bool pdl_enabled();

// <<<grid, block, smem, stream>>> with the programmatic-serialization attribute when `pdl`
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem,
                                        cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// Box overlap.  Every operation is an explicitly rounded float32 op (the _rn
// intrinsics are never contracted into FMAs), in the reference's operation
// order (network.py:462-511): area = (x2-x1)*(y2-y1);
// inter = max(0, min(ax2,bx2)-max(ax1,bx1)) * max(0, min(ay2,by2)-max(ay1,by1));
// iou = inter / ((a_area + b_area) - inter).
// ---------------------------------------------------------------------------
struct Box {
  float x1, y1, x2, y2, area;
};

__device__ __forceinline__ Box make_box(float4 v) {
  Box b;
  b.x1 = v.x; b.y1 = v.y; b.x2 = v.z; b.y2 = v.w;
  b.area = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
  return b;
}

__device__ __forceinline__ float box_intersection(const Box& a, const Box& b) {
  const float x1 = fmaxf(a.x1, b.x1);
  const float y1 = fmaxf(a.y1, b.y1);
  const float x2 = fminf(a.x2, b.x2);
  const float y2 = fminf(a.y2, b.y2);
  const float w = fmaxf(0.0f, __fsub_rn(x2, x1));
  const float h = fmaxf(0.0f, __fsub_rn(y2, y1));
  return __fmul_rn(w, h);
}

__device__ __forceinline__ float box_iou(const Box& a, const Box& b) {
  const float inter = box_intersection(a, b);
  const float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
  // 0 / positive is +0 exactly; a zero numerator would leave div.rn's fast path,
  // and most pairs are disjoint.  Everything else takes the exact IEEE division.
  const bool zero = (inter == 0.0f) && (uni > 0.0f);
  const float q = __fdiv_rn(zero ? uni : inter, uni);
  return zero ? 0.0f : q;
}

// index of the image that owns detection row `row`: largest i with off[i] <= row
__device__ __forceinline__ int find_image(const int32_t* __restrict__ off, int num_images, int row) {
  int lo = 0, hi = num_images;  // invariant: off[lo] <= row < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= row) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}  // namespace gn
