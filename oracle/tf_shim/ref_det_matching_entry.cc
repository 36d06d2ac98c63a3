// ORACLE SUPPORT: C entry point that drives the reference's own
// DetectionMatchingOp<float>::Compute (compiled from
// /root/reference/nms_net/matching_module/det_matching.cc against tf_shim).
#include "tensorflow/core/framework/op_kernel.h"

using namespace tensorflow;

extern "C" int ref_detection_matching(const float* iou, const float* score,
                                      const uint8_t* ignore, int n_dets, int n_gt,
                                      float* labels, float* weights,
                                      int32_t* assignment, char* err, int err_len) {
  auto it = KernelRegistry().find("DetectionMatching:CPU");
  if (it == KernelRegistry().end()) return 2;
  OpKernelConstruction cons;
  std::unique_ptr<OpKernel> k(it->second(&cons));
  static_assert(sizeof(bool) == 1, "bool is one byte");
  Tensor t_iou((void*)iou, TensorShape({n_dets, n_gt}));
  Tensor t_score((void*)score, TensorShape({n_dets}));
  Tensor t_ign((void*)ignore, TensorShape({n_gt}));
  Tensor t_lab(labels, TensorShape({n_dets}));
  Tensor t_w(weights, TensorShape({n_dets}));
  Tensor t_as(assignment, TensorShape({n_dets}));
  OpKernelContext ctx;
  ctx.inputs = {&t_iou, &t_score, &t_ign};
  ctx.outputs = {&t_lab, &t_w, &t_as};
  k->Compute(&ctx);
  if (!ctx.status().ok()) {
    if (err && err_len > 0) {
      std::strncpy(err, ctx.status().error_message().c_str(), err_len - 1);
      err[err_len - 1] = 0;
    }
    return 1;
  }
  return 0;
}
