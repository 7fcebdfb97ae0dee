// tf_stub_harness.cc -- registries behind tests/tf_stub/tf_stub.h plus the extern "C" surface the Python tests drive
// (tests/test_tf_shim_stub.py): list the registered ops, run a shape function, construct an OpKernel from attribute
// text and run its Compute on device buffers.  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tf_stub.h"

namespace tensorflow {
namespace {
std::vector<OpDef>& Ops() {
  static std::vector<OpDef> v;
  return v;
}
struct KernelEntry {
  std::string op, device;
  std::function<OpKernel*(OpKernelConstruction*)> factory;
};
std::vector<KernelEntry>& Kernels() {
  static std::vector<KernelEntry> v;
  return v;
}
std::string Trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
ArgDef ParseArg(const std::string& spec) {
  ArgDef a;
  const size_t colon = spec.find(':');
  a.name = Trim(spec.substr(0, colon));
  std::string t = Trim(spec.substr(colon + 1));
  const size_t star = t.find('*');
  if (star != std::string::npos) {
    a.number_attr = Trim(t.substr(0, star));
    t = Trim(t.substr(star + 1));
  }
  a.type = t;
  return a;
}
const OpDef* FindOp(const char* name) {
  for (const auto& o : Ops())
    if (o.name == name) return &o;
  return nullptr;
}
DataType TypeOf(const std::string& t) {
  if (t == "float") return DT_FLOAT;
  if (t == "double") return DT_DOUBLE;
  if (t == "int32") return DT_INT32;
  if (t == "uint8") return DT_UINT8;
  if (t == "bool") return DT_BOOL;
  return DT_INVALID;
}
AttrMap ParseAttrText(const char* text) {  // "a=1;b=0.5;c=0.1,0.1,0.2,0.2"
  AttrMap m;
  std::string s = text ? text : "";
  size_t pos = 0;
  while (pos < s.size()) {
    size_t semi = s.find(';', pos);
    if (semi == std::string::npos) semi = s.size();
    const std::string item = s.substr(pos, semi - pos);
    const size_t eq = item.find('=');
    if (eq != std::string::npos) m[Trim(item.substr(0, eq))] = Trim(item.substr(eq + 1));
    pos = semi + 1;
  }
  return m;
}
void SetErr(char* err, int len, const std::string& msg) {
  if (err && len > 0) {
    std::snprintf(err, (size_t)len, "%s", msg.c_str());
  }
}
}  // namespace

void* StubDeviceAlloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    std::fprintf(stderr, "tf_stub: cudaMalloc(%zu) failed\n", bytes);
    std::abort();
  }
  return p;
}
void StubDeviceFree(void* p) { cudaFree(p); }

void Tensor::CheckType(DataType want) const {
  if (want != dt_) {
    std::fprintf(stderr, "tf_stub: flat<T>() dtype mismatch: tensor is %d, asked for %d\n", (int)dt_, (int)want);
    std::abort();
  }
}

OpDefBuilderWrapper& OpDefBuilderWrapper::Input(const std::string& spec) {
  def_.inputs.push_back(ParseArg(spec));
  return *this;
}
OpDefBuilderWrapper& OpDefBuilderWrapper::Output(const std::string& spec) {
  def_.outputs.push_back(ParseArg(spec));
  return *this;
}
OpDefBuilderWrapper& OpDefBuilderWrapper::Attr(const std::string& spec) {
  AttrDef a;
  const size_t colon = spec.find(':');
  a.name = Trim(spec.substr(0, colon));
  std::string rest = Trim(spec.substr(colon + 1));
  const size_t ge = rest.find(">=");
  size_t eq = rest.find('=');
  if (ge != std::string::npos && eq == ge + 1) eq = rest.find('=', ge + 2);  // "int >= 1": a constraint, not a default
  if (eq != std::string::npos) {
    a.has_default = true;
    a.def = Trim(rest.substr(eq + 1));
    rest = Trim(rest.substr(0, eq));
  }
  if (ge != std::string::npos) rest = Trim(rest.substr(0, ge));
  a.type = rest;
  if (a.has_default && a.type == "list(float)") {  // "[0.1, 0.2]" -> "0.1,0.2"
    std::string d;
    for (char ch : a.def)
      if (ch != '[' && ch != ']' && ch != ' ') d.push_back(ch);
    a.def = d;
  }
  def_.attrs.push_back(a);
  return *this;
}
OpRegistrar::OpRegistrar(const OpDefBuilderWrapper& b) { Ops().push_back(b.def()); }
KernelRegistrar::KernelRegistrar(const KernelDefBuilder& b, std::function<OpKernel*(OpKernelConstruction*)> factory) {
  Kernels().push_back(KernelEntry{b.op, b.device, std::move(factory)});
}

Status ParseAttr(const std::string& text, int* v) {
  char* end = nullptr;
  const long x = std::strtol(text.c_str(), &end, 10);
  if (end == text.c_str() || *end) return errors::InvalidArgument("not an int: '", text, "'");
  *v = (int)x;
  return Status();
}
Status ParseAttr(const std::string& text, float* v) {
  char* end = nullptr;
  const float x = std::strtof(text.c_str(), &end);
  if (end == text.c_str() || *end) return errors::InvalidArgument("not a float: '", text, "'");
  *v = x;
  return Status();
}
Status ParseAttr(const std::string& text, bool* v) {
  if (text == "true" || text == "True" || text == "1") *v = true;
  else if (text == "false" || text == "False" || text == "0") *v = false;
  else return errors::InvalidArgument("not a bool: '", text, "'");
  return Status();
}
Status ParseAttr(const std::string& text, std::vector<float>* v) {
  v->clear();
  size_t pos = 0;
  while (pos <= text.size() && !text.empty()) {
    size_t comma = text.find(',', pos);
    if (comma == std::string::npos) comma = text.size();
    float x;
    TF_RETURN_IF_ERROR(ParseAttr(Trim(text.substr(pos, comma - pos)), &x));
    v->push_back(x);
    pos = comma + 1;
  }
  return Status();
}
Status LookupAttr(const OpDef& op, const AttrMap& given, const std::string& name, std::string* text) {
  for (const auto& a : op.attrs) {
    if (a.name != name) continue;
    auto it = given.find(name);
    if (it != given.end()) *text = it->second;
    else if (a.has_default) *text = a.def;
    else return errors::InvalidArgument("op ", op.name, ": required attr '", name, "' not given");
    return Status();
  }
  return errors::InvalidArgument("op ", op.name, " has no attr '", name, "'");
}

Status OpKernelContext::allocate_output(int i, const TensorShape& s, Tensor** t) {
  if (i < 0 || i >= (int)out_.size()) return errors::InvalidArgument("output index ", i, " out of range");
  const DataType dt = TypeOf(op_->outputs[(size_t)i].type);
  if (dt == DT_INVALID) return errors::Internal("output ", i, " has an unknown type");
  out_[(size_t)i] = std::make_unique<Tensor>(dt, s);
  *t = out_[(size_t)i].get();
  return Status();
}

struct StubKernel {
  const OpDef* op;
  std::unique_ptr<OpKernel> kernel;
  std::unique_ptr<OpKernelContext> last;
};
}  // namespace tensorflow

namespace tf = tensorflow;

extern "C" {

int tfstub_num_ops() { return (int)tf::Ops().size(); }

// "Name|in0:type[*N],in1:type|out0:type|attr:type[=default];..." ; returns the length written (0 = no such op)
int tfstub_op_signature(int index, char* buf, int len) {
  if (index < 0 || index >= (int)tf::Ops().size()) return 0;
  const tf::OpDef& o = tf::Ops()[(size_t)index];
  std::string s = o.name + "|";
  for (size_t i = 0; i < o.inputs.size(); ++i)
    s += (i ? "," : "") + o.inputs[i].name + ":" + o.inputs[i].type +
         (o.inputs[i].number_attr.empty() ? "" : "*" + o.inputs[i].number_attr);
  s += "|";
  for (size_t i = 0; i < o.outputs.size(); ++i) s += (i ? "," : "") + o.outputs[i].name + ":" + o.outputs[i].type;
  s += "|";
  for (size_t i = 0; i < o.attrs.size(); ++i)
    s += (i ? ";" : "") + o.attrs[i].name + ":" + o.attrs[i].type + (o.attrs[i].has_default ? "=" + o.attrs[i].def : "");
  std::snprintf(buf, (size_t)len, "%s", s.c_str());
  return (int)s.size();
}

// devices a kernel is registered for, comma separated ("GPU"); the north star allows no CPU kernel
int tfstub_kernel_devices(const char* op, char* buf, int len) {
  std::string s;
  for (const auto& k : tf::Kernels())
    if (k.op == op) s += (s.empty() ? "" : ",") + k.device;
  std::snprintf(buf, (size_t)len, "%s", s.c_str());
  return (int)s.size();
}

// shape function: in_dims holds the input shapes back to back (in_ndims[i] entries each); out_dims receives the
// output shapes back to back (-1 = unknown dimension), out_ndims[i] their ranks
int tfstub_infer_shapes(const char* op_name, const char* attrs, int n_in, const int* in_ndims, const long long* in_dims,
                        int max_out, int* out_ndims, long long* out_dims, int max_out_dims, char* err, int errlen) {
  const tf::OpDef* op = tf::FindOp(op_name);
  if (!op) { tf::SetErr(err, errlen, "no such op"); return -1; }
  std::vector<tf::shape_inference::ShapeHandle> in((size_t)n_in);
  size_t off = 0;
  for (int i = 0; i < n_in; ++i) {
    in[(size_t)i].known = true;
    for (int d = 0; d < in_ndims[i]; ++d) in[(size_t)i].d.push_back(in_dims[off++]);
  }
  tf::shape_inference::InferenceContext c(op, tf::ParseAttrText(attrs), in);
  if (!op->shape_fn) { tf::SetErr(err, errlen, "op has no shape function"); return -1; }
  const tf::Status s = op->shape_fn(&c);
  if (!s.ok()) { tf::SetErr(err, errlen, s.error_message()); return -1; }
  const auto& outs = c.outputs();
  if ((int)outs.size() > max_out) { tf::SetErr(err, errlen, "too many outputs"); return -1; }
  size_t o = 0;
  for (size_t i = 0; i < outs.size(); ++i) {
    out_ndims[i] = outs[i].known ? (int)outs[i].d.size() : -1;
    for (long long d : outs[i].d) {
      if ((int)o >= max_out_dims) { tf::SetErr(err, errlen, "too many output dims"); return -1; }
      out_dims[o++] = d;
    }
  }
  return (int)outs.size();
}

void* tfstub_kernel_create(const char* op_name, const char* attrs, char* err, int errlen) {
  const tf::OpDef* op = tf::FindOp(op_name);
  if (!op) { tf::SetErr(err, errlen, "no such op"); return nullptr; }
  for (const auto& k : tf::Kernels()) {
    if (k.op != op_name || k.device != "GPU") continue;
    tf::OpKernelConstruction c(op, tf::ParseAttrText(attrs));
    std::unique_ptr<tf::OpKernel> kernel(k.factory(&c));
    if (!c.status().ok()) { tf::SetErr(err, errlen, c.status().error_message()); return nullptr; }
    auto* sk = new tf::StubKernel{op, std::move(kernel), nullptr};
    return sk;
  }
  tf::SetErr(err, errlen, "no GPU kernel registered");
  return nullptr;
}

// inputs: borrowed device pointers.  Returns 0, or the TF error code with the message in `err`.
int tfstub_kernel_compute(void* handle, int n_in, void* const* in_ptrs, const int* in_dtypes, const int* in_ndims,
                          const long long* in_dims, void* stream, char* err, int errlen) {
  auto* sk = static_cast<tf::StubKernel*>(handle);
  std::vector<tf::Tensor> in;
  size_t off = 0;
  for (int i = 0; i < n_in; ++i) {
    std::vector<tf::int64> d;
    for (int k = 0; k < in_ndims[i]; ++k) d.push_back(in_dims[off++]);
    in.emplace_back((tf::DataType)in_dtypes[i], tf::TensorShape(d), in_ptrs[i]);
  }
  sk->last = std::make_unique<tf::OpKernelContext>(sk->op, std::move(in), stream);
  sk->kernel->Compute(sk->last.get());
  if (!sk->last->status().ok()) {
    tf::SetErr(err, errlen, sk->last->status().error_message());
    return sk->last->status().code();
  }
  return 0;
}

int tfstub_output_info(void* handle, int i, int* dtype, int* ndims, long long* dims, int max_dims, void** data) {
  auto* sk = static_cast<tf::StubKernel*>(handle);
  if (!sk->last || i < 0 || i >= (int)sk->last->outputs().size() || !sk->last->outputs()[(size_t)i]) return -1;
  const tf::Tensor& t = *sk->last->outputs()[(size_t)i];
  if (t.dims() > max_dims) return -1;
  *dtype = (int)t.dtype();
  *ndims = t.dims();
  for (int d = 0; d < t.dims(); ++d) dims[d] = t.dim_size(d);
  *data = t.raw();
  return 0;
}

// device-to-device copy of output i into a caller-owned buffer, ordered on `stream`
int tfstub_output_copy(void* handle, int i, void* dst, size_t bytes, void* stream) {
  auto* sk = static_cast<tf::StubKernel*>(handle);
  if (!sk->last || i < 0 || i >= (int)sk->last->outputs().size() || !sk->last->outputs()[(size_t)i]) return -1;
  const tf::Tensor& t = *sk->last->outputs()[(size_t)i];
  if (bytes != (size_t)t.NumElements() * tf::DataTypeSize(t.dtype())) return -2;
  if (bytes == 0) return 0;
  return (int)cudaMemcpyAsync(dst, t.raw(), bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}

void tfstub_kernel_destroy(void* handle) { delete static_cast<tf::StubKernel*>(handle); }

}  // extern "C"
