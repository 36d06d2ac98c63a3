// ORACLE SUPPORT: stand-in for TF's shape-inference types (see op_kernel.h).
#ifndef ORACLE_TF_SHIM_SHAPE_INFERENCE_H_
#define ORACLE_TF_SHIM_SHAPE_INFERENCE_H_
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
namespace shape_inference {
struct DimensionHandle {};
struct ShapeHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int) { return ShapeHandle(); }
  Status WithRank(ShapeHandle, int, ShapeHandle*) { return Status::OK(); }
  void set_output(int, ShapeHandle) {}
};
}  // namespace shape_inference

#define TF_RETURN_IF_ERROR(expr)            \
  do {                                      \
    ::tensorflow::Status _s = (expr);       \
    if (!_s.ok()) return _s;                \
  } while (0)
}  // namespace tensorflow
#endif
