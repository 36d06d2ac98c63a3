// ORACLE SUPPORT: int32 / int64 typedefs live in the shim's op_kernel.h.
#include "tensorflow/core/framework/op_kernel.h"
