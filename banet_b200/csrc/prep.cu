// Level pre-/post-steps: rays, grad_fixed(+concat,+half swap), bilinear resampler, depth composition.
// All are one-pass streaming kernels (HBM-bound, coalesced along the channel axis).
#include "common.cuh"
#include "lm_build.h"

namespace banet {

// BundleNet.computeCoordinates (reference bundlenet.py:112-120; legacy/ba.py:27-34 when !normalize)
__global__ void compute_coordinates_kernel(const float* __restrict__ points, const float* __restrict__ intr,
                                           int nb, int N, int normalize, float* __restrict__ p)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nb * N) return;
    const int b = (int)(i / N), n = (int)(i - (long long)b * N);
    const float fx = intr[b * 4], fy = intr[b * 4 + 1], ox = intr[b * 4 + 2], oy = intr[b * 4 + 3];
    const float2 uv = reinterpret_cast<const float2*>(points)[i];
    float x = (uv.x - ox) / fx, y = (uv.y - oy) / fy, z = 1.f;
    if (normalize) {                                   // tf.nn.l2_normalize: x * rsqrt(max(sum sq, 1e-12))
        const float inv = 1.0f / sqrtf(fmaxf(x * x + y * y + 1.f, 1e-12f));
        x *= inv; y *= inv; z *= inv;
    }
    float* pb = p + (size_t)b * 3 * N;
    pb[n] = x; pb[(size_t)N + n] = y; pb[2 * (size_t)N + n] = z;
}

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// grad_fixed + concat (+ half swap): bundlenet.py:92-100, 386-389.  One thread per (texel, 4 channels).
template <int VEC>
__global__ void grad_fixed_concat_kernel(const float* __restrict__ F, int nb, int h, int w, int C, int swap_halves,
                                         float* __restrict__ out)
{
    const int cv = C / VEC;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nb * h * w * cv;
    if (i >= total) return;
    const int c = (int)(i % cv) * VEC;
    long long t = i / cv;
    const int x = (int)(t % w); t /= w;
    const int y = (int)(t % h); const int b = (int)(t / h);
    const int bs = swap_halves ? (b + nb / 2) % nb : b;            // layers[nb/2:nb] ++ layers[0:nb/2]
    const float* img = F + (size_t)bs * h * w * C;
    const int xe = reflect1(x + 1, w), xw = reflect1(x - 1, w), ys = reflect1(y + 1, h), yn = reflect1(y - 1, h);
    float* o = out + (((size_t)b * h + y) * w + x) * 3 * C;
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
        const float f = img[((size_t)y * w + x) * C + c + u];
        const float gx = 0.5f * (img[((size_t)y * w + xe) * C + c + u] - img[((size_t)y * w + xw) * C + c + u]);
        const float gy = 0.5f * (img[((size_t)ys * w + x) * C + c + u] - img[((size_t)yn * w + x) * C + c + u]);
        o[c + u] = f; o[C + c + u] = gx; o[2 * C + c + u] = gy;
    }
}

// tf.contrib.resampler.resampler: bilinear, zero outside.  Warp per point, lanes over channels.
__global__ void resample_kernel(const float* __restrict__ data, const float* __restrict__ xy, float cs,
                                int nb, int h, int w, int C, int N, float* __restrict__ out)
{
    const long long pt = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= (long long)nb * N) return;
    const int b = (int)(pt / N);
    const float x = xy[pt * 2] * cs, y = xy[pt * 2 + 1] * cs;
    const float fx = floorf(x), fy = floorf(y);
    const float dx = x - fx, dy = y - fy;
    const bool fin = isfinite(x) && isfinite(y) && fabsf(x) < 1e9f && fabsf(y) < 1e9f;
    const int x0 = fin ? (int)fx : -10, y0 = fin ? (int)fy : -10;
    const float wt[4] = {(1.f - dx) * (1.f - dy), dx * (1.f - dy), (1.f - dx) * dy, dx * dy};
    const float* img = data + (size_t)b * h * w * C;
    float* o = out + (size_t)pt * C;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
            const int xx = x0 + (tp & 1), yy = y0 + (tp >> 1);
            if (xx >= 0 && xx < w && yy >= 0 && yy < h) acc = fmaf(wt[tp], __ldg(img + ((size_t)yy * w + xx) * C + c), acc);
        }
        o[c] = acc;
    }
}

// bundlenet.py:397: depth = init_depth + basis . W.   Warp per output texel.
__global__ void depth_compose_kernel(const float* __restrict__ init_depth, const float* __restrict__ basis,
                                     const float* __restrict__ W, int nb, int M, int K, float* __restrict__ out)
{
    const long long pt = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= (long long)nb * M) return;
    const int b = (int)(pt / M);
    const float* br = basis + (size_t)pt * K;
    const float* wb = W + (size_t)b * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(ld_stream_f1(br + k), __ldg(wb + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[pt] = init_depth[pt] + acc;
}

}  // namespace banet

using namespace banet;

extern "C" int banet_compute_coordinates(const float* points, const float* intr, int nb, int N, int normalize,
                                         float* p, banet_stream_t stream)
{
    BANET_REQUIRE(points && intr && p && nb > 0 && N > 0, BANET_ERR_BAD_ARG, "compute_coordinates: bad argument");
    const long long tot = (long long)nb * N;
    compute_coordinates_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, intr, nb, N, normalize, p);
    BANET_CUDA_LAUNCH_CHECK("compute_coordinates");
    return BANET_OK;
}

extern "C" int banet_grad_fixed_concat(const float* F, int nb, int h, int w, int C, int swap_halves,
                                       float* conv2, banet_stream_t stream)
{
    BANET_REQUIRE(F && conv2 && nb > 0 && h >= 2 && w >= 2 && C > 0, BANET_ERR_BAD_ARG, "grad_fixed_concat: bad argument (need h,w >= 2)");
    BANET_REQUIRE(!swap_halves || nb % 2 == 0, BANET_ERR_BAD_ARG, "grad_fixed_concat: swap_halves needs an even batch");
    if (C % 4 == 0) {
        const long long tot = (long long)nb * h * w * (C / 4);
        grad_fixed_concat_kernel<4><<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(F, nb, h, w, C, swap_halves, conv2);
    } else {
        const long long tot = (long long)nb * h * w * C;
        grad_fixed_concat_kernel<1><<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(F, nb, h, w, C, swap_halves, conv2);
    }
    BANET_CUDA_LAUNCH_CHECK("grad_fixed_concat");
    return BANET_OK;
}

extern "C" int banet_resample(const float* data, const float* xy, float coord_scale, int nb, int h, int w, int C, int N,
                              float* out, banet_stream_t stream)
{
    BANET_REQUIRE(data && xy && out && nb > 0 && h > 0 && w > 0 && C > 0 && N > 0, BANET_ERR_BAD_ARG, "resample: bad argument");
    const long long thr = (long long)nb * N * 32;
    resample_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, (cudaStream_t)stream>>>(data, xy, coord_scale, nb, h, w, C, N, out);
    BANET_CUDA_LAUNCH_CHECK("resample");
    return BANET_OK;
}

extern "C" int banet_depth_compose(const float* init_depth, const float* basis, const float* W, int nb, int M, int K,
                                   float* out, banet_stream_t stream)
{
    BANET_REQUIRE(init_depth && basis && W && out && nb > 0 && M > 0 && K > 0, BANET_ERR_BAD_ARG, "depth_compose: bad argument");
    const long long thr = (long long)nb * M * 32;
    depth_compose_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, (cudaStream_t)stream>>>(init_depth, basis, W, nb, M, K, out);
    BANET_CUDA_LAUNCH_CHECK("depth_compose");
    return BANET_OK;
}
