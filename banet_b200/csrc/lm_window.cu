// lm_window.cu — joint LM step of a keyframe window: nf frame pairs (keyframe -> frame f) that share the keyframe's depth D + B.W.
//
// SURVEY.md section 8f-4.  NOT in the reference: its BA layer is 2-view (one pose + one W per pair, bundlenet.py:193-278); BA-Net's
// 5-frame use case runs it as 4 independent pairs (legacy/seq_example.py).  Here the pairs of a window share ONE W, so the unknowns are
// 6 nf + K and the normal matrix is block-arrow:
//
//        | Hcc_0              Hcd_0 |        per-pair blocks exactly as the 2-view build produces them (banet_lm_build, nb = nf):
//   Hj = |        ...          ...  |        Hcc_f 6x6, Hcd_f 6xK, Hdd_f KxK, g_f;
//        |             Hcc_nf  Hcd_nf|        the depth block and the depth right-hand side are the sums over the frames
//        | Hcd_0' ...  Hcd_nf' S Hdd |        (the residuals of all frames depend on the same W).
//
// The rest follows the 2-view iteration: lambda from the mean |residual| over ALL points of ALL frames through the same MLP
// (bundlenet.py:241-253), damping of every diagonal entry but the last depth coefficient (:264-266), one solve, every frame's pose
// updated with its own 6 entries (:269-275), W with the shared K.  The solve is the fused lm_step kernel on the one (6 nf + K) system.
#include "common.cuh"
#include "lm_build.h"

namespace banet {
namespace {

__global__ void window_assemble_kernel(const float* __restrict__ H, const float* __restrict__ g, const float* __restrict__ rbar_sum,
                                       int nf, int K, int C, float* __restrict__ Hj, float* __restrict__ gj, float* __restrict__ rbar_j,
                                       float* __restrict__ zero_w)
{
    const int P = 6 + K, np = 6 * nf, Pj = np + K;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int idx = t0; idx < Pj * Pj; idx += stride) {
        const int i = idx / Pj, j = idx - i * Pj;
        float v = 0.f;
        if (i < np && j < np) {
            const int fi = i / 6, fj = j / 6;
            if (fi == fj) v = H[((size_t)fi * P + (i - 6 * fi)) * P + (j - 6 * fj)];
        } else if (i < np) {
            const int f = i / 6;
            v = H[((size_t)f * P + (i - 6 * f)) * P + 6 + (j - np)];
        } else if (j < np) {
            const int f = j / 6;
            v = H[((size_t)f * P + 6 + (i - np)) * P + (j - 6 * f)];
        } else {
            double acc = 0.0;                                       // fixed order over the frames: bit-reproducible
            for (int f = 0; f < nf; ++f) acc += (double)H[((size_t)f * P + 6 + (i - np)) * P + 6 + (j - np)];
            v = (float)acc;
        }
        Hj[idx] = v;
    }
    for (int i = t0; i < Pj; i += stride) {
        if (i < np) { const int f = i / 6; gj[i] = g[(size_t)f * P + (i - 6 * f)]; }
        else { double acc = 0.0; for (int f = 0; f < nf; ++f) acc += (double)g[(size_t)f * P + 6 + (i - np)]; gj[i] = (float)acc; }
    }
    for (int c = t0; c < C; c += stride) {
        double acc = 0.0;
        for (int f = 0; f < nf; ++f) acc += (double)rbar_sum[(size_t)f * C + c];
        rbar_j[c] = (float)acc;
    }
    for (int i = t0; i < Pj; i += stride) zero_w[i] = 0.f;      // the "W" the solve kernel updates on the side (unused)
}

__global__ void window_scatter_kernel(const float* __restrict__ delta_j, const int32_t* __restrict__ status_j, int nf, int K,
                                      float* __restrict__ delta_f, float* __restrict__ W, int32_t* __restrict__ status)
{
    const int P = 6 + K, np = 6 * nf;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int idx = t0; idx < nf * P; idx += stride) {
        const int f = idx / P, c = idx - f * P;
        const float v = c < 6 ? delta_j[6 * f + c] : delta_j[np + c - 6];
        delta_f[idx] = v;
        if (c >= 6) W[(size_t)f * K + c - 6] += v;               // every frame's copy of the shared W gets the same update
    }
    for (int f = t0; f < nf; f += stride) status[f] |= status_j[0];
}

__global__ void window_broadcast_w_kernel(float* __restrict__ W, int nf, int K)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (nf - 1) * K) W[K + i] = W[i % K];
}

}  // namespace

bool lm_window_supported(int nf, int K, int C) { return nf >= 1 && K >= 1 && lm_step_supported(6 * nf + K, C); }

size_t lm_window_step_workspace_floats(int nf, int K, int C)
{
    const size_t Pj = 6 * (size_t)nf + K, P = 6 + (size_t)K;
    return Pj * Pj + Pj + (size_t)C + Pj + 1 + (size_t)nf * P + 12 + 2 * Pj + 8;
}

int lm_window_broadcast_w(float* W, int nf, int K, cudaStream_t st)
{
    if (nf > 1) {
        window_broadcast_w_kernel<<<((nf - 1) * K + 255) / 256, 256, 0, st>>>(W, nf, K);
        BANET_CUDA_LAUNCH_CHECK("window_broadcast_w_kernel launch");
    }
    return BANET_OK;
}

int lm_window_step(const float* H, const float* g, const float* rbar_sum, int nf, int N, int C, int K, const float* mlp, float base,
                   const float* lambda_in, const banet_solve_opts_t& opts, float* R, float* T, float* W, float* ws, float* lambda_out,
                   int32_t* status, cudaStream_t st)
{
    const int P = 6 + K, Pj = 6 * nf + K;
    BANET_REQUIRE(lm_window_supported(nf, K, C), BANET_ERR_UNSUPPORTED, "lm_window_step: 6*%d+%d unknowns with C=%d do not fit the solve kernel", nf, K, C);
    float* Hj = ws;                 float* gj = Hj + (size_t)Pj * Pj;   float* rbar_j = gj + Pj;        float* delta_j = rbar_j + C;
    float* lam = delta_j + Pj;      float* delta_f = lam + 1;           float* dumR = delta_f + (size_t)nf * P;
    float* dumT = dumR + 9;         float* zero_w = dumT + 3;           float* dumW = zero_w + Pj;
    int32_t* status_j = reinterpret_cast<int32_t*>(dumW + Pj);
    window_assemble_kernel<<<64, 256, 0, st>>>(H, g, rbar_sum, nf, K, C, Hj, gj, rbar_j, zero_w);
    BANET_CUDA_LAUNCH_CHECK("window_assemble_kernel launch");
    // one system of 6 nf + K unknowns: the solve kernel sees "pose" = frame 0's six and "W" = everything else; its own pose / W outputs go
    // to scratch, the real update is the scatter below.  The mean |residual| divides by all nf * N points.
    int rc = lm_step(Hj, gj, rbar_j, 1, N * nf, C, Pj - 6, mlp, base, lambda_in, kStepBundleNet, nullptr, opts, R, T, zero_w, dumR, dumT, dumW,
                     delta_j, lam, status_j, 0, st);
    if (rc) return rc;
    window_scatter_kernel<<<(nf * P + 255) / 256, 256, 0, st>>>(delta_j, status_j, nf, K, delta_f, W, status);
    BANET_CUDA_LAUNCH_CHECK("window_scatter_kernel launch");
    if (lambda_out) {
        cudaError_t e = cudaMemcpyAsync(lambda_out, lam, sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("lm_window_step: %s", cudaGetErrorString(e)); return BANET_ERR_CUDA; }
    }
    return launch_pose_update(delta_f, nf, P, R, T, R, T, st);
}

}  // namespace banet
