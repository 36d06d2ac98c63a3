// ORACLE SUPPORT: thread::ThreadPool lives in the shim's op_kernel.h.
#include "tensorflow/core/framework/op_kernel.h"
