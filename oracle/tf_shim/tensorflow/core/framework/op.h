// ORACLE SUPPORT: stand-in for TF's REGISTER_OP builder (see op_kernel.h).
#ifndef ORACLE_TF_SHIM_OP_H_
#define ORACLE_TF_SHIM_OP_H_
#include "tensorflow/core/framework/shape_inference.h"

namespace tensorflow {
class OpDefBuilderShim {
 public:
  explicit OpDefBuilderShim(const char*) {}
  OpDefBuilderShim& Attr(const char*) { return *this; }
  OpDefBuilderShim& Input(const char*) { return *this; }
  OpDefBuilderShim& Output(const char*) { return *this; }
  OpDefBuilderShim& Doc(const char*) { return *this; }
  template <typename F> OpDefBuilderShim& SetShapeFn(F) { return *this; }
};
#define REGISTER_OP(NAME) \
  static ::tensorflow::OpDefBuilderShim TF_SHIM_CAT(_shim_op_, __COUNTER__) = \
      ::tensorflow::OpDefBuilderShim(NAME)
}  // namespace tensorflow
#endif
