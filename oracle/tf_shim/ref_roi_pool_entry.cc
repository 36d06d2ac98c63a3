// ORACLE SUPPORT: C entry points that drive the reference's own
// RoiPoolOp<CPUDevice,float> / RoiPoolGradOp<CPUDevice,float>::Compute (compiled
// from /root/reference/nms_net/roi_pooling_layer/roi_pooling_op.cc against
// tf_shim).  Shard() runs the work on std::threads; the CUDA launchers the file
// declares are stubbed (the GPU variants are never constructed here).
#include <cmath>
#include <thread>

#include "third_party/eigen3/unsupported/Eigen/CXX11/Tensor"
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
void Shard(int max_parallelism, thread::ThreadPool*, int64 total, int64,
           std::function<void(int64, int64)> work) {
  int nt = std::max(1, std::min<int>(max_parallelism, (int)std::thread::hardware_concurrency()));
  if (total < 4096 || nt == 1) { work(0, total); return; }
  std::vector<std::thread> th;
  const int64 per = (total + nt - 1) / nt;
  for (int i = 0; i < nt; ++i) {
    const int64 lo = i * per, hi = std::min<int64>(total, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
}
}  // namespace tensorflow

bool ROIPoolForwardLaucher(const float*, const float, const int, const int, const int, const int,
                           const int, const int, const float*, float*, int*,
                           const Eigen::GpuDevice&) { return false; }
bool ROIPoolBackwardLaucher(const float*, const float, const int, const int, const int, const int,
                            const int, const int, const int, const float*, float*, const int*,
                            const Eigen::GpuDevice&) { return false; }

using namespace tensorflow;

static int run(const char* key, OpKernelConstruction* cons, OpKernelContext* ctx, char* err,
               int err_len) {
  auto it = KernelRegistry().find(key);
  if (it == KernelRegistry().end()) return 2;
  std::unique_ptr<OpKernel> k(it->second(cons));
  const Status* bad = nullptr;
  if (!cons->status().ok()) bad = &cons->status();
  else {
    k->Compute(ctx);
    if (!ctx->status().ok()) bad = &ctx->status();
  }
  if (bad) {
    if (err && err_len > 0) {
      std::strncpy(err, bad->error_message().c_str(), err_len - 1);
      err[err_len - 1] = 0;
    }
    return 1;
  }
  return 0;
}

extern "C" int ref_roi_pool_fwd(const float* data, int b, int h, int w, int c, const float* rois,
                                int r, int rois_rank_is_2, int ph, int pw, float scale,
                                float* top, int32_t* argmax, int threads, char* err, int err_len) {
  OpKernelConstruction cons;
  cons.attrs = {{"pooled_height", ph}, {"pooled_width", pw}, {"spatial_scale", scale}};
  Tensor t_data((void*)data, TensorShape({b, h, w, c}));
  Tensor t_rois((void*)rois, rois_rank_is_2 ? TensorShape({r, 5}) : TensorShape({r * 5}));
  Tensor t_top(top, TensorShape({r, ph, pw, c}));
  Tensor t_arg(argmax, TensorShape({r, ph, pw, c}));
  OpKernelContext ctx;
  ctx.dev.cpu_threads.num_threads = threads;
  ctx.inputs = {&t_data, &t_rois};
  ctx.outputs = {&t_top, &t_arg};
  return run("RoiPool:CPU", &cons, &ctx, err, err_len);
}

extern "C" int ref_roi_pool_bwd(const float* data, int b, int h, int w, int c, const float* rois,
                                int r, const int32_t* argmax, const float* grad, int ph, int pw,
                                float scale, float* out, int threads, char* err, int err_len) {
  OpKernelConstruction cons;
  cons.attrs = {{"pooled_height", ph}, {"pooled_width", pw}, {"spatial_scale", scale}};
  Tensor t_data((void*)data, TensorShape({b, h, w, c}));
  Tensor t_rois((void*)rois, TensorShape({r, 5}));
  Tensor t_arg((void*)argmax, TensorShape({r, ph, pw, c}));
  Tensor t_grad((void*)grad, TensorShape({r, ph, pw, c}));
  Tensor t_out(out, TensorShape({b, h, w, c}));
  OpKernelContext ctx;
  ctx.dev.cpu_threads.num_threads = threads;
  ctx.inputs = {&t_data, &t_rois, &t_arg, &t_grad};
  ctx.outputs = {&t_out};
  return run("RoiPoolGrad:CPU", &cons, &ctx, err, err_len);
}
