// tf_stub.h — the minimum of the TensorFlow-1.x C++ API that /root/reference/utils.cu touches, so that the reference's UNMODIFIED
// source file compiles and runs without TensorFlow.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py and oracle/Makefile).
//
// Nothing here is copied from TensorFlow; each class is the smallest stand-in with the same names and call signatures:
//   tensorflow::{Tensor, TensorShape, PersistentTensor, Status, OpKernel, OpKernelConstruction, OpKernelContext, DeviceContext,
//                shape_inference::InferenceContext, REGISTER_OP, REGISTER_KERNEL_BUILDER, OP_REQUIRES_OK, CHECK_*}
//   perftools::gputools::{DeviceMemory, DeviceMemoryBase, ScratchAllocator, Stream, port::StatusOr, blas::Transpose}
// Third-party behaviour restated (TensorFlow stream_executor, not part of /root/reference):
//   Stream::ThenBlasGemmBatchedWithScratch forwards its arguments, in order and unchanged, to cublasSgemmBatched (column-major), with the
//   per-matrix pointer arrays copied to device scratch obtained from the ScratchAllocator — which is what TF's CUDA BLAS plugin does.
#pragma once
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tensorflow {
typedef long long int64;
typedef unsigned char uint8;
typedef int int32;

class Status {
public:
    Status() : ok_(true) {}
    explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
    static Status OK() { return Status(); }
    bool ok() const { return ok_; }
    const std::string& error_message() const { return msg_; }
private:
    bool ok_; std::string msg_;
};

enum DataType { DT_INVALID = 0, DT_FLOAT = 1, DT_UINT8 = 4 };
static const char* const DEVICE_GPU = "GPU";

class TensorShape {
public:
    TensorShape() {}
    TensorShape(std::initializer_list<int64> d) : d_(d) {}
    int64 dim_size(int i) const { return d_[i]; }
    void set_dim(int i, int64 v) { d_[i] = v; }
    void AddDim(int64 v) { d_.push_back(v); }
    int dims() const { return (int)d_.size(); }
    int64 num_elements() const { int64 n = 1; for (int64 v : d_) n *= v; return n; }
private:
    std::vector<int64> d_;
};

namespace stub {
struct Buffer {
    void* p = nullptr; bool own = false;
    Buffer(void* q, bool o) : p(q), own(o) {}
    ~Buffer() { if (own && p) cudaFree(p); }
};
template <typename T> struct Flat {
    T* p; size_t n;
    T* data() const { return p; }
    size_t size() const { return n; }
};
inline size_t dtype_size(DataType t) { return t == DT_UINT8 ? 1 : 4; }
}  // namespace stub

class Tensor {
public:
    Tensor() : dt_(DT_INVALID) {}
    Tensor(DataType dt, const TensorShape& sh, void* external) : dt_(dt), sh_(sh), buf_(std::make_shared<stub::Buffer>(external, false)) {}
    static bool Allocate(DataType dt, const TensorShape& sh, Tensor* out) {
        void* p = nullptr;
        const size_t bytes = std::max<size_t>((size_t)sh.num_elements() * stub::dtype_size(dt), 16);
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
        out->dt_ = dt; out->sh_ = sh; out->buf_ = std::make_shared<stub::Buffer>(p, true);
        return true;
    }
    const TensorShape& shape() const { return sh_; }
    DataType dtype() const { return dt_; }
    template <typename T> stub::Flat<T> flat() const { return stub::Flat<T>{reinterpret_cast<T*>(buf_->p), (size_t)sh_.num_elements()}; }
private:
    DataType dt_; TensorShape sh_; std::shared_ptr<stub::Buffer> buf_;
};

class PersistentTensor {
public:
    Tensor* AccessTensor() { return &t_; }
    Tensor t_;
};
}  // namespace tensorflow

namespace perftools { namespace gputools {
typedef long long int64;
typedef unsigned char uint8;

class DeviceMemoryBase {
public:
    explicit DeviceMemoryBase(void* p = nullptr, uint64_t size = 0) : p_(p), size_(size) {}
    void* opaque() const { return p_; }
    uint64_t size() const { return size_; }
private:
    void* p_; uint64_t size_;
};
template <typename T> class DeviceMemory : public DeviceMemoryBase {
public:
    DeviceMemory() {}
    explicit DeviceMemory(const DeviceMemoryBase& o) : DeviceMemoryBase(o.opaque(), o.size()) {}
    static DeviceMemory<T> MakeFromByteSize(void* p, uint64_t bytes) { return DeviceMemory<T>(DeviceMemoryBase(p, bytes)); }
};
namespace port {
template <typename T> class StatusOr {
public:
    StatusOr(const T& v) : v_(v) {}
    bool ok() const { return v_.opaque() != nullptr; }
    T ValueOrDie() const { return v_; }
private:
    T v_;
};
}  // namespace port
namespace blas { enum class Transpose { kNoTranspose, kTranspose, kConjugateTranspose }; }

class Stream;
class ScratchAllocator {
public:
    virtual ~ScratchAllocator() {}
    virtual int64 GetMemoryLimitInBytes(Stream* stream) = 0;
    virtual port::StatusOr<DeviceMemory<uint8>> AllocateBytes(Stream* stream, int64 byte_size) = 0;
};

class StreamImplementation {
public:
    explicit StreamImplementation(cudaStream_t s) : s_(s) {}
    void* GpuStreamMemberHack() { return &s_; }        // TF: pointer to the CUstream member (utils.cu:82-89 casts it to cudaStream_t*)
    cudaStream_t s_;
};

class Stream {
public:
    explicit Stream(cudaStream_t s) : impl_(s), ok_(true) {
        if (cublasCreate(&h_) != CUBLAS_STATUS_SUCCESS) { ok_ = false; h_ = nullptr; }
        else cublasSetStream(h_, s);
    }
    ~Stream() { if (h_) cublasDestroy(h_); }
    StreamImplementation* implementation() { return &impl_; }
    bool ok() const { return ok_; }
    // TF stream_executor: DoBlasGemmBatched -> cublasSgemmBatched(handle, transa, transb, m, n, k, &alpha, a[], lda, b[], ldb, &beta, c[], ldc, batch)
    Stream& ThenBlasGemmBatchedWithScratch(blas::Transpose ta, blas::Transpose tb, uint64_t m, uint64_t n, uint64_t k, float alpha,
                                           const std::vector<DeviceMemory<float>*>& a, int lda,
                                           const std::vector<DeviceMemory<float>*>& b, int ldb, float beta,
                                           const std::vector<DeviceMemory<float>*>& c, int ldc, int batch_count,
                                           ScratchAllocator* scratch) {
        if (!ok_) return *this;
        std::vector<const float*> ha(batch_count), hb(batch_count);
        std::vector<float*> hc(batch_count);
        for (int i = 0; i < batch_count; ++i) {
            ha[i] = static_cast<const float*>(a[i]->opaque()); hb[i] = static_cast<const float*>(b[i]->opaque());
            hc[i] = static_cast<float*>(c[i]->opaque());
        }
        const size_t bytes = sizeof(void*) * (size_t)batch_count;
        auto sa = scratch->AllocateBytes(this, (int64)bytes), sb = scratch->AllocateBytes(this, (int64)bytes), sc = scratch->AllocateBytes(this, (int64)bytes);
        if (!sa.ok() || !sb.ok() || !sc.ok()) { ok_ = false; return *this; }
        void *da = sa.ValueOrDie().opaque(), *db = sb.ValueOrDie().opaque(), *dc = sc.ValueOrDie().opaque();
        cudaMemcpyAsync(da, ha.data(), bytes, cudaMemcpyHostToDevice, impl_.s_);
        cudaMemcpyAsync(db, hb.data(), bytes, cudaMemcpyHostToDevice, impl_.s_);
        cudaMemcpyAsync(dc, hc.data(), bytes, cudaMemcpyHostToDevice, impl_.s_);
        cudaStreamSynchronize(impl_.s_);                 // the host arrays die at return
        auto op = [](blas::Transpose t) { return t == blas::Transpose::kNoTranspose ? CUBLAS_OP_N : (t == blas::Transpose::kTranspose ? CUBLAS_OP_T : CUBLAS_OP_C); };
        cublasStatus_t st = cublasSgemmBatched(h_, op(ta), op(tb), (int)m, (int)n, (int)k, &alpha,
                                               reinterpret_cast<const float* const*>(da), lda, reinterpret_cast<const float* const*>(db), ldb, &beta,
                                               reinterpret_cast<float* const*>(dc), ldc, batch_count);
        if (st != CUBLAS_STATUS_SUCCESS) { ok_ = false; std::fprintf(stderr, "tf_stub: cublasSgemmBatched failed (%d)\n", (int)st); }
        return *this;
    }
private:
    StreamImplementation impl_; cublasHandle_t h_ = nullptr; bool ok_;
};
}}  // namespace perftools::gputools

namespace tensorflow {
class DeviceContext {
public:
    explicit DeviceContext(perftools::gputools::Stream* s) : s_(s) {}
    perftools::gputools::Stream* stream() { return s_; }
private:
    perftools::gputools::Stream* s_;
};
struct GpuDeviceInfo { int gpu_id = 0; };
class DeviceBase {
public:
    const GpuDeviceInfo* tensorflow_gpu_device_info() const { return &info_; }
    GpuDeviceInfo info_;
};
class OpKernelConstruction {
public:
    DeviceBase* device() { return &dev_; }
    DeviceBase dev_;
};

class OpKernelContext {
public:
    explicit OpKernelContext(DeviceContext* dc) : dc_(dc) {}
    void add_input(const Tensor& t) { in_.push_back(t); }
    void set_output_buffer(int idx, float* p) { if ((int)outbuf_.size() <= idx) outbuf_.resize(idx + 1, nullptr); outbuf_[idx] = p; }
    const Tensor& input(int i) const { return in_[i]; }
    DeviceContext* op_device_context() { return dc_; }
    Status allocate_output(int idx, const TensorShape& sh, Tensor** out) {
        if ((int)out_.size() <= idx) out_.resize(idx + 1);
        if (idx < (int)outbuf_.size() && outbuf_[idx]) out_[idx] = std::make_shared<Tensor>(DT_FLOAT, sh, outbuf_[idx]);
        else { out_[idx] = std::make_shared<Tensor>(); if (!Tensor::Allocate(DT_FLOAT, sh, out_[idx].get())) return Status("allocate_output failed"); }
        *out = out_[idx].get();
        return Status::OK();
    }
    Status allocate_temp(DataType dt, const TensorShape& sh, Tensor* out) { return Tensor::Allocate(dt, sh, out) ? Status::OK() : Status("allocate_temp failed"); }
    Status allocate_persistent(DataType dt, const TensorShape& sh, PersistentTensor* pt, Tensor** out) {
        if (!Tensor::Allocate(dt, sh, &pt->t_)) return Status("allocate_persistent failed");
        if (out) *out = &pt->t_;
        return Status::OK();
    }
    void SetStatus(const Status& s) { status_ = s; }
    const Status& status() const { return status_; }
    const TensorShape& output_shape(int idx) const { return out_[idx]->shape(); }
private:
    DeviceContext* dc_; std::vector<Tensor> in_; std::vector<float*> outbuf_; std::vector<std::shared_ptr<Tensor>> out_; Status status_;
};

class OpKernel {
public:
    explicit OpKernel(OpKernelConstruction*) {}
    virtual ~OpKernel() {}
    virtual void Compute(OpKernelContext* context) = 0;
};

namespace shape_inference {
struct DimensionHandle { int64 v = -1; };
struct ShapeHandle { std::vector<int64> d; };
class InferenceContext {
public:
    std::vector<ShapeHandle> inputs, outputs;
    ShapeHandle input(int i) const { return inputs[i]; }
    DimensionHandle Dim(const ShapeHandle& s, int i) const { DimensionHandle h; h.v = s.d[i]; return h; }
    ShapeHandle Vector(DimensionHandle a) const { ShapeHandle s; s.d = {a.v}; return s; }
    ShapeHandle Matrix(DimensionHandle a, DimensionHandle b) const { ShapeHandle s; s.d = {a.v, b.v}; return s; }
    Status Concatenate(const ShapeHandle& a, const ShapeHandle& b, ShapeHandle* out) const { out->d = a.d; out->d.insert(out->d.end(), b.d.begin(), b.d.end()); return Status::OK(); }
    void set_output(int i, const ShapeHandle& s) { if ((int)outputs.size() <= i) outputs.resize(i + 1); outputs[i] = s; }
};
}  // namespace shape_inference

namespace stub {
typedef std::function<Status(shape_inference::InferenceContext*)> ShapeFn;
typedef std::function<OpKernel*(OpKernelConstruction*)> KernelFactory;
inline std::map<std::string, ShapeFn>& shape_fns() { static std::map<std::string, ShapeFn> m; return m; }
inline std::map<std::string, KernelFactory>& kernel_factories() { static std::map<std::string, KernelFactory> m; return m; }
class OpBuilder {
public:
    explicit OpBuilder(const char* name) : name_(name) {}
    OpBuilder& Input(const char*) { return *this; }
    OpBuilder& Output(const char*) { return *this; }
    OpBuilder& Attr(const char*) { return *this; }
    OpBuilder& SetShapeFn(ShapeFn fn) { fn_ = fn; return *this; }
    std::string name_; ShapeFn fn_;
};
struct OpReg { OpReg(const OpBuilder& b) { if (b.fn_) shape_fns()[b.name_] = b.fn_; } };
class KernelName {
public:
    explicit KernelName(const char* n) : name(n) {}
    KernelName& Device(const char*) { return *this; }
    std::string name;
};
struct KernelReg { KernelReg(const KernelName& n, KernelFactory f) { kernel_factories()[n.name] = f; } };
}  // namespace stub
inline stub::KernelName Name(const char* n) { return stub::KernelName(n); }
}  // namespace tensorflow

#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::stub::OpReg TF_STUB_CAT(tf_stub_op_reg_, __COUNTER__) = ::tensorflow::stub::OpBuilder(name)
#define REGISTER_KERNEL_BUILDER(kname, cls) \
    static ::tensorflow::stub::KernelReg TF_STUB_CAT(tf_stub_kernel_reg_, __COUNTER__)(kname, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new cls(c); })
#define OP_REQUIRES_OK(ctx, expr) do { ::tensorflow::Status s__(expr); if (!s__.ok()) { (ctx)->SetStatus(s__); return; } } while (0)
#define CHECK_GT(a, b) do { if (!((a) > (b))) { std::fprintf(stderr, "CHECK_GT failed: %s:%d\n", __FILE__, __LINE__); std::abort(); } } while (0)
#define CHECK_NOTNULL(p) (p)
