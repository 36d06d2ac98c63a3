// ORACLE SUPPORT (test infrastructure): a minimal stand-in for the handful of
// TensorFlow C++ framework types that the reference's custom CPU op sources
// touch, so that /root/reference/nms_net/matching_module/det_matching.cc can be
// (and nms_net/roi_pooling_layer/roi_pooling_op.cc) can be
// compiled UNMODIFIED, from where they lie, into oracle/_ref/ and run as the
// ground truth for the DetectionMatching / RoiPool parity tests.  Nothing here is
// TensorFlow code; it only mimics names and call shapes (TF ~0.12 API).
#ifndef ORACLE_TF_SHIM_OP_KERNEL_H_
#define ORACLE_TF_SHIM_OP_KERNEL_H_

#include <math.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

namespace tensorflow {

typedef int32_t int32;
typedef int64_t int64;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline void Append(std::ostringstream&) {}
template <typename A, typename... R>
void Append(std::ostringstream& os, const A& a, const R&... r) { os << a; Append(os, r...); }
template <typename... Args>
Status InvalidArgument(const Args&... args) {
  std::ostringstream os;
  Append(os, args...);
  return Status(os.str());
}
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : dims_(d) {}
  int dims() const { return (int)dims_.size(); }
  int64 dim_size(int i) const { return dims_[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (int64 d : dims_) n *= d;
    return n;
  }
 private:
  std::vector<int64> dims_;
};

struct TensorShapeUtils {
  static bool IsVector(const TensorShape& s) { return s.dims() == 1; }
  static bool IsMatrix(const TensorShape& s) { return s.dims() == 2; }
  template <typename I>
  static Status MakeShape(const I* dims, int n, TensorShape* out) {
    std::vector<int64> d(dims, dims + n);
    *out = TensorShape(d);
    return Status::OK();
  }
};

template <typename T, int R>
class TensorView {
 public:
  TensorView(T* data, const int64* dims) : data_(data) {
    for (int i = 0; i < R; ++i) dims_[i] = dims[i];
  }
  int64 dimension(int i) const { return dims_[i]; }
  T* data() const { return data_; }
  int64 size() const {
    int64 n = 1;
    for (int i = 0; i < R; ++i) n *= dims_[i];
    return n;
  }
  T& operator()(int64 i) const { return data_[i]; }
  T& operator()(int64 i, int64 j) const { return data_[i * dims_[1] + j]; }
  void setZero() { setConstant(T(0)); }
  void setConstant(T v) {
    int64 n = 1;
    for (int i = 0; i < R; ++i) n *= dims_[i];
    for (int64 k = 0; k < n; ++k) data_[k] = v;
  }
 private:
  T* data_;
  int64 dims_[R];
};

template <typename T>
struct TTypes {
  typedef TensorView<T, 1> Flat;
  typedef TensorView<const T, 1> ConstFlat;
  typedef TensorView<T, 2> Matrix;
  typedef TensorView<const T, 2> ConstMatrix;
};

class Tensor {
 public:
  Tensor() : data_(nullptr) {}
  // view over caller memory
  Tensor(void* data, const TensorShape& shape) : shape_(shape), data_(data) {}
  // owning
  Tensor(size_t elem_size, const TensorShape& shape)
      : shape_(shape), store_(new char[elem_size * (size_t)std::max<int64>(shape.num_elements(), 1)]),
        data_(store_.get()) {}
  const TensorShape& shape() const { return shape_; }
  int dims() const { return shape_.dims(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  void* raw() const { return data_; }

  template <typename T> TensorView<T, 1> flat() {
    int64 n = shape_.num_elements();
    return TensorView<T, 1>(static_cast<T*>(data_), &n);
  }
  template <typename T> TensorView<const T, 1> flat() const {
    int64 n = shape_.num_elements();
    return TensorView<const T, 1>(static_cast<const T*>(data_), &n);
  }
  template <typename T, int R> TensorView<T, R> tensor() {
    int64 d[R];
    for (int i = 0; i < R; ++i) d[i] = shape_.dim_size(i);
    return TensorView<T, R>(static_cast<T*>(data_), d);
  }
  template <typename T, int R> TensorView<const T, R> tensor() const {
    int64 d[R];
    for (int i = 0; i < R; ++i) d[i] = shape_.dim_size(i);
    return TensorView<const T, R>(static_cast<const T*>(data_), d);
  }
 private:
  TensorShape shape_;
  std::shared_ptr<char> store_;
  void* data_;
};

namespace thread { class ThreadPool {}; }
struct DeviceBase {
  struct CpuWorkerThreads {
    int num_threads = 1;
    thread::ThreadPool* workers = nullptr;
  };
  CpuWorkerThreads cpu_threads;
  const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &cpu_threads; }
};

// op attributes are supplied by the driver before the kernel is constructed
class OpKernelConstruction {
 public:
  std::map<std::string, double> attrs;
  Status status_;
  const Status& status() const { return status_; }
  template <typename V>
  Status GetAttr(const char* name, V* out) const {
    auto it = attrs.find(name);
    if (it == attrs.end()) return Status(std::string("missing attr ") + name);
    *out = static_cast<V>(it->second);
    return Status::OK();
  }
  void CtxFailure(const Status& s) { status_ = s; }
};

class OpKernelContext {
 public:
  std::vector<const Tensor*> inputs;
  std::vector<Tensor*> outputs;  // caller-provided views, indexed by output slot
  Status status_;
  const Status& status() const { return status_; }
  const Tensor& input(int i) { return *inputs[i]; }
  Status allocate_output(int i, const TensorShape&, Tensor** out) {
    *out = outputs[i];
    return Status::OK();
  }
  void CtxFailure(const Status& s) { status_ = s; }
  DeviceBase dev;
  DeviceBase* device() { return &dev; }
  template <typename D> const D& eigen_device() const { static D d; return d; }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

// TF ~0.12 spelled these as plain if-blocks (no trailing ';' required).
#define OP_REQUIRES(CTX, EXP, STATUS) \
  if (!(EXP)) {                       \
    (CTX)->CtxFailure((STATUS));      \
    return;                           \
  }
#define OP_REQUIRES_OK(CTX, STATUS)   \
  do {                                \
    ::tensorflow::Status _s(STATUS);  \
    if (!_s.ok()) {                   \
      (CTX)->CtxFailure(_s);          \
      return;                         \
    }                                 \
  } while (0)

// ---- kernel registry --------------------------------------------------------
static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

struct KernelDef {
  std::string op, device;
};
class Name {
 public:
  explicit Name(const char* op) { def_.op = op; }
  Name& Device(const char* d) { def_.device = d; return *this; }
  template <typename T> Name& TypeConstraint(const char*) { return *this; }
  const KernelDef& def() const { return def_; }
 private:
  KernelDef def_;
};

typedef std::function<OpKernel*(OpKernelConstruction*)> KernelFactory;
inline std::map<std::string, KernelFactory>& KernelRegistry() {
  static std::map<std::string, KernelFactory> r;
  return r;
}
struct KernelRegistrar {
  KernelRegistrar(const Name& n, KernelFactory f) {
    KernelRegistry()[n.def().op + ":" + n.def().device] = f;
  }
};
#define TF_SHIM_CAT2(a, b) a##b
#define TF_SHIM_CAT(a, b) TF_SHIM_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(BUILDER, ...)                                 \
  static ::tensorflow::KernelRegistrar TF_SHIM_CAT(_shim_kernel_, __COUNTER__)( \
      BUILDER, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow

#endif  // ORACLE_TF_SHIM_OP_KERNEL_H_
