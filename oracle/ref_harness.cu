// ref_harness.cu — C entry points around the reference's OWN op kernels (EquationConstruction / EquationConstructionGrad, compiled
// unmodified from /root/reference/utils.cu by oracle/Makefile against oracle/tf_stub).  TEST INFRASTRUCTURE ONLY.
// Device pointers in, device pointers out, everything on the default stream, synchronous.
// The reference op keeps process-static scratch sized by its FIRST call (utils.cu:210-216, 259-296, 519-524): one library instance
// serves ONE shape; the Python loader (oracle/ref_lib.py) loads a fresh copy of the .so per shape.  The gradient op reuses the forward
// op's scratch (utils.cu:515-516), so banet_ref_eqc_fwd must have run first, with the same shape.
#include "tf_stub.h"

using namespace tensorflow;

namespace {
perftools::gputools::Stream* g_stream = nullptr;
DeviceContext* g_dc = nullptr;
OpKernel* g_fwd = nullptr;
OpKernel* g_bwd = nullptr;
int g_shape[4] = {0, 0, 0, 0};

int ensure(int nb, int N, int C, int P)
{
    if (!g_stream) {
        g_stream = new perftools::gputools::Stream((cudaStream_t)0);
        if (!g_stream->ok()) return -3;
        g_dc = new DeviceContext(g_stream);
        OpKernelConstruction ctor;
        auto& f = stub::kernel_factories();
        if (!f.count("EquationConstruction") || !f.count("EquationConstructionGrad")) return -4;
        g_fwd = f["EquationConstruction"](&ctor);
        g_bwd = f["EquationConstructionGrad"](&ctor);
        g_shape[0] = nb; g_shape[1] = N; g_shape[2] = C; g_shape[3] = P;
    }
    if (g_shape[0] != nb || g_shape[1] != N || g_shape[2] != C || g_shape[3] != P) return -5;     // static scratch: one shape per instance
    return 0;
}
}  // namespace

extern "C" int banet_ref_eqc_fwd(const float* J, const float* G, const float* d, int nb, int N, int C, int P, float* AtA, float* Atb)
{
    if (N < 2) return -1;                                  // ColumnReduceSimpleKernel reads rows 0 and 1 unconditionally (utils.cu:194)
    int rc = ensure(nb, N, C, P);
    if (rc) return rc;
    OpKernelContext ctx(g_dc);
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, N, 2, P}), const_cast<float*>(J)));
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, N, C, 2}), const_cast<float*>(G)));
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, N, C, 1}), const_cast<float*>(d)));
    ctx.set_output_buffer(0, AtA); ctx.set_output_buffer(1, Atb);
    g_fwd->Compute(&ctx);
    if (cudaDeviceSynchronize() != cudaSuccess) return -3;
    if (!ctx.status().ok() || !g_stream->ok()) return -3;
    // the op's own shape function (utils.cu:156-171) must describe what was produced
    shape_inference::InferenceContext ic;
    ic.inputs = {{{nb, N, 2, P}}, {{nb, N, C, 2}}, {{nb, N, C, 1}}};
    auto& sf = stub::shape_fns();
    if (sf.count("EquationConstruction")) {
        if (!sf["EquationConstruction"](&ic).ok()) return -6;
        if (ic.outputs.size() != 2 || ic.outputs[0].d != std::vector<int64>({nb, P, P}) || ic.outputs[1].d != std::vector<int64>({nb, P, 1})) return -6;
    }
    return 0;
}

extern "C" int banet_ref_eqc_bwd(const float* J, const float* G, const float* d, const float* gAtA, const float* gAtb,
                                 int nb, int N, int C, int P, float* dJ, float* dG, float* dd)
{
    if (!g_fwd) return -2;                                 // the gradient op tiles into the forward op's buffer (utils.cu:515-516)
    int rc = ensure(nb, N, C, P);
    if (rc) return rc;
    OpKernelContext ctx(g_dc);
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, N, 2, P}), const_cast<float*>(J)));
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, N, C, 2}), const_cast<float*>(G)));
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, N, C, 1}), const_cast<float*>(d)));
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, P, P}), const_cast<float*>(gAtA)));
    ctx.add_input(Tensor(DT_FLOAT, TensorShape({nb, P, 1}), const_cast<float*>(gAtb)));
    ctx.set_output_buffer(0, dJ); ctx.set_output_buffer(1, dG); ctx.set_output_buffer(2, dd);
    g_bwd->Compute(&ctx);
    if (cudaDeviceSynchronize() != cudaSuccess) return -3;
    return (ctx.status().ok() && g_stream->ok()) ? 0 : -3;
}
