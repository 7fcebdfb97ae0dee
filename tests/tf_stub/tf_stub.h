// tf_stub.h -- a small functional stand-in for the part of TensorFlow's C++ custom-op API that
// maskrcnn_tf2_b200/tf_shim/mrcnn_roi_ops.cc uses (REGISTER_OP / shape functions / OpKernel / OpKernelContext /
// Tensor / Status / OP_REQUIRES).  TEST INFRASTRUCTURE ONLY: TensorFlow cannot be installed in this image, so the shim
// is compiled against these declarations instead, which (1) type-checks every launcher call of the shim against
// include/mrcnn_roi_b200.h, (2) lets the CPU tests read back the registered op signatures and run the shape
// functions, and (3) lets the GPU tests execute each OpKernel::Compute through a fake OpKernelContext (device
// buffers from cudaMalloc) and compare the outputs with the ctypes path.  Written from the documented behaviour of
// the TF API (tensorflow/core/framework/{op,op_kernel,shape_inference}.h); nothing here ships.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {
struct GpuDevice {
  void* stream_ = nullptr;
  void* stream() const { return stream_; }  // cudaStream_t in the real API
};
}  // namespace Eigen

namespace tensorflow {
using int64 = long long;
using int32 = int;
using uint8 = unsigned char;

enum DataType { DT_INVALID = 0, DT_FLOAT = 1, DT_DOUBLE = 2, DT_INT32 = 3, DT_UINT8 = 4, DT_BOOL = 10 };
template <class T> struct DataTypeOf;
template <> struct DataTypeOf<float> { static constexpr DataType v = DT_FLOAT; };
template <> struct DataTypeOf<double> { static constexpr DataType v = DT_DOUBLE; };
template <> struct DataTypeOf<int32> { static constexpr DataType v = DT_INT32; };
template <> struct DataTypeOf<uint8> { static constexpr DataType v = DT_UINT8; };
template <> struct DataTypeOf<bool> { static constexpr DataType v = DT_BOOL; };
inline size_t DataTypeSize(DataType t) { return t == DT_DOUBLE ? 8 : (t == DT_FLOAT || t == DT_INT32) ? 4 : 1; }

class Status {
 public:
  Status() = default;
  Status(int code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  bool ok() const { return code_ == 0; }
  int code() const { return code_; }
  const std::string& error_message() const { return msg_; }

 private:
  int code_ = 0;
  std::string msg_;
};

namespace errors {
template <class... A> std::string StrCat(const A&... a) {
  std::ostringstream os;
  (void)std::initializer_list<int>{((os << a), 0)...};
  return os.str();
}
template <class... A> Status InvalidArgument(const A&... a) { return Status(3, StrCat(a...)); }
template <class... A> Status Internal(const A&... a) { return Status(13, StrCat(a...)); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() = default;
  TensorShape(std::initializer_list<int64> d) : d_(d) {}
  explicit TensorShape(std::vector<int64> d) : d_(std::move(d)) {}
  int dims() const { return (int)d_.size(); }
  int64 dim_size(int i) const { return d_.at(i); }
  int64 num_elements() const {
    int64 n = 1;
    for (int64 v : d_) n *= v;
    return n;
  }
  bool operator==(const TensorShape& o) const { return d_ == o.d_; }
  bool operator!=(const TensorShape& o) const { return d_ != o.d_; }
  const std::vector<int64>& vec() const { return d_; }

 private:
  std::vector<int64> d_;
};

void* StubDeviceAlloc(size_t bytes);  // tf_stub_harness.cc: cudaMalloc
void StubDeviceFree(void* p);

class Tensor {
 public:
  template <class T> struct Flat {
    T* p;
    T* data() const { return p; }
  };
  Tensor() = default;
  Tensor(DataType dt, TensorShape s, void* borrowed) : dt_(dt), shape_(std::move(s)), data_(borrowed) {}
  Tensor(DataType dt, TensorShape s) : dt_(dt), shape_(std::move(s)) {
    const size_t bytes = (size_t)shape_.num_elements() * DataTypeSize(dt);
    data_ = StubDeviceAlloc(bytes ? bytes : 1);
    own_ = std::shared_ptr<void>(data_, StubDeviceFree);
  }
  DataType dtype() const { return dt_; }
  const TensorShape& shape() const { return shape_; }
  int dims() const { return shape_.dims(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  int64 NumElements() const { return shape_.num_elements(); }
  void* raw() const { return data_; }
  template <class T> Flat<T> flat() {
    CheckType(DataTypeOf<T>::v);
    return Flat<T>{static_cast<T*>(data_)};
  }
  template <class T> Flat<const T> flat() const {
    CheckType(DataTypeOf<T>::v);
    return Flat<const T>{static_cast<const T*>(data_)};
  }

 private:
  void CheckType(DataType want) const;  // aborts the test process on a dtype mismatch, like TF's CHECK
  DataType dt_ = DT_INVALID;
  TensorShape shape_;
  void* data_ = nullptr;
  std::shared_ptr<void> own_;
};

// ---- op registry ------------------------------------------------------------------------------------------
struct AttrDef {
  std::string name, type, def;  // type: "int" | "float" | "bool" | "list(float)"; def: text after '=' ("" = required)
  bool has_default = false;
};
struct ArgDef {
  std::string name, type, number_attr;  // "N * float" -> type "float", number_attr "N"
};
namespace shape_inference {
class InferenceContext;
}
struct OpDef {
  std::string name;
  std::vector<ArgDef> inputs, outputs;
  std::vector<AttrDef> attrs;
  std::function<Status(shape_inference::InferenceContext*)> shape_fn;
};
using AttrMap = std::map<std::string, std::string>;  // attr name -> value text ("1000", "0.7", "true", "0.1,0.1,0.2,0.2")

Status ParseAttr(const std::string& text, int* v);
Status ParseAttr(const std::string& text, float* v);
Status ParseAttr(const std::string& text, bool* v);
Status ParseAttr(const std::string& text, std::vector<float>* v);
Status LookupAttr(const OpDef& op, const AttrMap& given, const std::string& name, std::string* text);

class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char* name) { def_.name = name; }
  OpDefBuilderWrapper& Input(const std::string& spec);
  OpDefBuilderWrapper& Output(const std::string& spec);
  OpDefBuilderWrapper& Attr(const std::string& spec);
  template <class F> OpDefBuilderWrapper& SetShapeFn(F f) {
    def_.shape_fn = f;
    return *this;
  }
  const OpDef& def() const { return def_; }

 private:
  OpDef def_;
};
struct OpRegistrar {
  OpRegistrar(const OpDefBuilderWrapper& b);  // NOLINT: implicit, as in REGISTER_OP(...) = builder chain
};

namespace shape_inference {
struct DimensionHandle {
  int64 v = -1;
};
struct ShapeHandle {
  std::vector<int64> d;
  bool known = false;
};
struct DimensionOrConstant {
  int64 v;
  DimensionOrConstant(DimensionHandle h) : v(h.v) {}  // NOLINT
  DimensionOrConstant(int64 c) : v(c) {}              // NOLINT
};
class InferenceContext {
 public:
  InferenceContext(const OpDef* op, AttrMap attrs, std::vector<ShapeHandle> in)
      : op_(op), attrs_(std::move(attrs)), in_(std::move(in)), out_(op->outputs.size()) {}
  template <class T> Status GetAttr(const std::string& name, T* v) const {
    std::string text;
    Status s = LookupAttr(*op_, attrs_, name, &text);
    return s.ok() ? ParseAttr(text, v) : s;
  }
  ShapeHandle input(int i) const { return in_.at(i); }
  DimensionHandle Dim(const ShapeHandle& s, int i) const { return DimensionHandle{s.known ? s.d.at(i) : -1}; }
  DimensionHandle UnknownDim() const { return DimensionHandle{-1}; }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant> dims) const {
    ShapeHandle h;
    h.known = true;
    for (const auto& d : dims) h.d.push_back(d.v);
    return h;
  }
  void set_output(int i, const ShapeHandle& s) { out_.at(i) = s; }
  const std::vector<ShapeHandle>& outputs() const { return out_; }

 private:
  const OpDef* op_;
  AttrMap attrs_;
  std::vector<ShapeHandle> in_, out_;
};
}  // namespace shape_inference

// ---- kernels ------------------------------------------------------------------------------------------------
class OpKernelConstruction {
 public:
  OpKernelConstruction(const OpDef* op, AttrMap attrs) : op_(op), attrs_(std::move(attrs)) {}
  template <class T> Status GetAttr(const std::string& name, T* v) const {
    std::string text;
    Status s = LookupAttr(*op_, attrs_, name, &text);
    return s.ok() ? ParseAttr(text, v) : s;
  }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  void CtxFailureWithWarning(const Status& s) { CtxFailure(s); }
  const Status& status() const { return status_; }
  const OpDef* op_def() const { return op_; }

 private:
  const OpDef* op_;
  AttrMap attrs_;
  Status status_;
};

class OpKernelContext {
 public:
  OpKernelContext(const OpDef* op, std::vector<Tensor> inputs, void* stream)
      : op_(op), in_(std::move(inputs)), out_(op->outputs.size()) { dev_.stream_ = stream; }
  const Tensor& input(int i) const { return in_.at(i); }
  int num_inputs() const { return (int)in_.size(); }
  Status allocate_output(int i, const TensorShape& s, Tensor** t);
  Status allocate_temp(DataType dt, const TensorShape& s, Tensor* t) {
    *t = Tensor(dt, s);
    temps_.push_back(*t);  // keeps scratch alive until the context goes away (TF: until the stream is done with it)
    return Status();
  }
  template <class D> const D& eigen_device() const { return dev_; }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  void CtxFailureWithWarning(const Status& s) { CtxFailure(s); }
  const Status& status() const { return status_; }
  std::vector<std::unique_ptr<Tensor>>& outputs() { return out_; }

 private:
  const OpDef* op_;
  std::vector<Tensor> in_;
  std::vector<std::unique_ptr<Tensor>> out_;
  std::vector<Tensor> temps_;
  Eigen::GpuDevice dev_;
  Status status_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() = default;
  virtual void Compute(OpKernelContext* ctx) = 0;
};

constexpr const char* DEVICE_GPU = "GPU";
constexpr const char* DEVICE_CPU = "CPU";
struct KernelDefBuilder {
  std::string op, device;
  explicit KernelDefBuilder(const char* name) : op(name) {}
  KernelDefBuilder& Device(const char* d) {
    device = d;
    return *this;
  }
};
inline KernelDefBuilder Name(const char* n) { return KernelDefBuilder(n); }
struct KernelRegistrar {
  KernelRegistrar(const KernelDefBuilder& b, std::function<OpKernel*(OpKernelConstruction*)> factory);
};
}  // namespace tensorflow

#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpRegistrar TF_STUB_CAT(tf_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderWrapper(name)
#define REGISTER_KERNEL_BUILDER(builder, ...)                                                  \
  static ::tensorflow::KernelRegistrar TF_STUB_CAT(tf_stub_kernel_, __COUNTER__)(              \
      ::tensorflow::builder,                                                                   \
      [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })
#define TF_RETURN_IF_ERROR(expr)                  \
  do {                                            \
    const ::tensorflow::Status _s = (expr);       \
    if (!_s.ok()) return _s;                      \
  } while (0)
#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)                    \
  do {                                              \
    const ::tensorflow::Status _s(__VA_ARGS__);     \
    if (!_s.ok()) {                                 \
      (CTX)->CtxFailureWithWarning(_s);             \
      return;                                       \
    }                                               \
  } while (0)
