// Op-level drop-in for the reference's native TF op pair (the reference's own plugin boundary):
//   EquationConstruction      utils.cu:150-171 (op), :219-417 (kernel: 5 batched SGEMMs + 2 reductions)
//   EquationConstructionGrad  utils.cu:420-428 (op), :465-694 (kernel: tile + 6 batched SGEMMs)
//
// Same tensors in and out, but streaming: per pixel only M = G^T G (2x2) and q = G^T d (2) are formed,
//   AtA = sum_n J^T (M J)   as a [2N x P]^T [2N x P] register-tiled product,   Atb = sum_n J^T q,
// so the reference's [nb,N,2+P,P] persistent scratch (utils.cu:259-264) does not exist here.
// Backward uses the factored form of utils.cu:625-690 (A = G J is never materialised):
//   Y = J (2 Ghat)   Z = J ghat   S = Y J^T
//   dJ = M Y + q ghat^T     dd = G Z     dG = G S + d Z^T
#include "common.cuh"

namespace banet {

constexpr int EQC_THREADS = 256;
constexpr int EQC_TPX = 32;                  // pixels per sub-tile (64 J rows)
constexpr int EQC_ROWS = 2 * EQC_TPX;

struct EqcPlan { int T, BT, nblk, nblk_tri, nchunks, chunk_px; size_t ws_bytes; };

static EqcPlan eqc_plan(int nb, int N, int P, int num_sms)
{
    EqcPlan pl;
    pl.T = (P <= 32) ? 2 : 9;
    pl.BT = 16 * pl.T;
    pl.nblk = (P + pl.BT - 1) / pl.BT;
    pl.nblk_tri = pl.nblk * (pl.nblk + 1) / 2;
    int want = (2 * num_sms + nb * pl.nblk_tri - 1) / (nb * pl.nblk_tri);
    if (want < 1) want = 1;
    const int max_chunks = (N + EQC_TPX - 1) / EQC_TPX;
    pl.nchunks = want < max_chunks ? want : max_chunks;
    const int tiles = (max_chunks + pl.nchunks - 1) / pl.nchunks;
    pl.chunk_px = tiles * EQC_TPX;
    pl.nchunks = (N + pl.chunk_px - 1) / pl.chunk_px;
    pl.ws_bytes = align_up((size_t)nb * pl.nchunks * ((size_t)P * P + P) * sizeof(float), 256);
    return pl;
}

// per-pixel M (3) and q (2): warp per pixel, lanes over channels
__device__ __forceinline__ void pixel_Mq(const float* __restrict__ Gp, const float* __restrict__ dp, int C, int lane,
                                         float& m11, float& m12, float& m22, float& q1, float& q2)
{
    m11 = m12 = m22 = q1 = q2 = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float2 gv = *reinterpret_cast<const float2*>(Gp + 2 * (size_t)c);
        const float dv = dp[c];
        m11 = fmaf(gv.x, gv.x, m11); m12 = fmaf(gv.x, gv.y, m12); m22 = fmaf(gv.y, gv.y, m22);
        q1 = fmaf(gv.x, dv, q1); q2 = fmaf(gv.y, dv, q2);
    }
    m11 = warp_sum(m11); m12 = warp_sum(m12); m22 = warp_sum(m22); q1 = warp_sum(q1); q2 = warp_sum(q2);
}

template <int T>
__global__ void __launch_bounds__(EQC_THREADS)
eqc_fwd_kernel(const float* __restrict__ J, const float* __restrict__ G, const float* __restrict__ d,
               int N, int C, int P, int chunk_px, int nchunks, int nblk, float* __restrict__ partial)
{
    constexpr int BT = 16 * T, LD = BT + 1;
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm;                          // [EQC_ROWS][LD]  J rows, column block bi
    float* Us = Xs + EQC_ROWS * LD;          // [EQC_ROWS][LD]  (M J) rows, column block bj
    float* sMq = Us + EQC_ROWS * LD;         // [EQC_TPX][5]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, b = blockIdx.z;
    int bi = 0, rem = blockIdx.y;            // lower-triangular block enumeration
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    const int i0 = bi * BT, j0 = bj * BT;
    const int ti = tid >> 4, tj = tid & 15;
    const int px0 = chunk * chunk_px, px1 = min(N, px0 + chunk_px);

    float acc[T][T];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int c = 0; c < T; ++c) acc[a][c] = 0.f;
    float gacc[2] = {0.f, 0.f};              // Atb columns j0+tid, j0+tid+256 (only the bi==bj... see below)

    for (int n0 = px0; n0 < px1; n0 += EQC_TPX) {
        const int cnt = min(EQC_TPX, px1 - n0);
        // phase 1: M, q
        for (int i = warp; i < EQC_TPX; i += EQC_THREADS / 32) {
            float m11 = 0.f, m12 = 0.f, m22 = 0.f, q1 = 0.f, q2 = 0.f;
            if (i < cnt) {
                const size_t pix = (size_t)b * N + n0 + i;
                pixel_Mq(G + pix * C * 2, d + pix * C, C, lane, m11, m12, m22, q1, q2);
            }
            if (lane == 0) { sMq[i * 5] = m11; sMq[i * 5 + 1] = m12; sMq[i * 5 + 2] = m22; sMq[i * 5 + 3] = q1; sMq[i * 5 + 4] = q2; }
        }
        __syncthreads();
        // phase 2: stage J column blocks and U = M J
        for (int e = tid; e < EQC_TPX * BT; e += EQC_THREADS) {
            const int n = e / BT, cidx = e - n * BT;
            float xa0 = 0.f, xa1 = 0.f, xb0 = 0.f, xb1 = 0.f;
            if (n < cnt) {
                const float* Jp = J + ((size_t)b * N + n0 + n) * 2 * P;
                if (i0 + cidx < P) { xa0 = Jp[i0 + cidx]; xa1 = Jp[P + i0 + cidx]; }
                if (j0 + cidx < P) { xb0 = Jp[j0 + cidx]; xb1 = Jp[P + j0 + cidx]; }
            }
            const float m11 = sMq[n * 5], m12 = sMq[n * 5 + 1], m22 = sMq[n * 5 + 2];
            Xs[(2 * n) * LD + cidx] = xa0; Xs[(2 * n + 1) * LD + cidx] = xa1;
            Us[(2 * n) * LD + cidx] = m11 * xb0 + m12 * xb1; Us[(2 * n + 1) * LD + cidx] = m12 * xb0 + m22 * xb1;
        }
        __syncthreads();
        // phase 3: acc += Xs^T Us ; diagonal blocks also accumulate Atb = sum J^T q
#pragma unroll 2
        for (int k = 0; k < EQC_ROWS; ++k) {
            float xa[T], ub[T];
#pragma unroll
            for (int a = 0; a < T; ++a) { xa[a] = Xs[k * LD + ti + 16 * a]; ub[a] = Us[k * LD + tj + 16 * a]; }
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int c = 0; c < T; ++c) acc[a][c] = fmaf(xa[a], ub[c], acc[a][c]);
        }
        if (bi == bj && tid < BT) {
            for (int n = 0; n < EQC_TPX; ++n)
                gacc[0] += Xs[(2 * n) * LD + tid] * sMq[n * 5 + 3] + Xs[(2 * n + 1) * LD + tid] * sMq[n * 5 + 4];
        }
        __syncthreads();
    }
    float* slot = partial + ((size_t)b * nchunks + chunk) * ((size_t)P * P + P);
#pragma unroll
    for (int a = 0; a < T; ++a) {
        const int i = i0 + ti + 16 * a;
#pragma unroll
        for (int c = 0; c < T; ++c) {
            const int j = j0 + tj + 16 * c;
            if (i < P && j < P) slot[(size_t)i * P + j] = acc[a][c];
        }
    }
    if (bi == bj && tid < BT && i0 + tid < P) slot[(size_t)P * P + i0 + tid] = gacc[0];
    (void)nblk;
}

__global__ void eqc_reduce_kernel(const float* __restrict__ partial, int nchunks, int P, int BT,
                                  float* __restrict__ AtA, float* __restrict__ Atb)
{
    const int b = blockIdx.y;
    const size_t stride = (size_t)P * P + P;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int)stride) return;
    int i = 0, j = 0;
    bool is_rhs = e >= P * P;
    if (!is_rhs) { i = e / P; j = e - i * P; if (j / BT > i / BT) return; }   // upper blocks were not computed
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += (double)partial[((size_t)b * nchunks + c) * stride + e];
    if (is_rhs) { Atb[(size_t)b * P + e - P * P]  = (float)s; return; }
    if (j <= i) { AtA[((size_t)b * P + i) * P + j] = (float)s; AtA[((size_t)b * P + j) * P + i] = (float)s; }
    // (entries with j > i inside a diagonal block are the mirror of (j,i); skip them)
}

// ---------------------------------------------------------------------------------------------------
// backward: CTA per 32 pixels.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EQC_THREADS)
eqc_bwd_kernel(const float* __restrict__ J, const float* __restrict__ G, const float* __restrict__ d,
               const float* __restrict__ gAtA, const float* __restrict__ gAtb, int N, int C, int P, int PP, int exact_sym,
               float* __restrict__ dJ, float* __restrict__ dG, float* __restrict__ dd)
{
    extern __shared__ __align__(16) float sm[];
    float* Js = sm;                           // [EQC_ROWS][PP]  (PP = P rounded up to 4)
    float* Ys = Js + EQC_ROWS * PP;           // [EQC_ROWS][PP]
    float* sMq = Ys + EQC_ROWS * PP;          // [EQC_TPX][5]
    float* sZ = sMq + EQC_TPX * 5;            // [EQC_ROWS]
    float* sS = sZ + EQC_ROWS;                // [EQC_TPX][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, n0 = blockIdx.x * EQC_TPX, cnt = min(EQC_TPX, N - n0);
    const float* Gh = gAtA + (size_t)b * P * P;
    const float* gh = gAtb + (size_t)b * P;

    for (int e = tid; e < EQC_ROWS * PP; e += EQC_THREADS) {
        const int r = e / PP, j = e - r * PP;
        float v = 0.f;
        if ((r >> 1) < cnt && j < P) v = J[(((size_t)b * N + n0 + (r >> 1)) * 2 + (r & 1)) * P + j];
        Js[e] = v;
    }
    for (int i = warp; i < EQC_TPX; i += EQC_THREADS / 32) {
        float m11 = 0.f, m12 = 0.f, m22 = 0.f, q1 = 0.f, q2 = 0.f;
        if (i < cnt) {
            const size_t pix = (size_t)b * N + n0 + i;
            pixel_Mq(G + pix * C * 2, d + pix * C, C, lane, m11, m12, m22, q1, q2);
        }
        if (lane == 0) { sMq[i * 5] = m11; sMq[i * 5 + 1] = m12; sMq[i * 5 + 2] = m22; sMq[i * 5 + 3] = q1; sMq[i * 5 + 4] = q2; }
    }
    __syncthreads();

    // Y = J * Ghat_s,  Ghat_s = 2 Ghat (reference, utils.cu:648-657) or Ghat + Ghat^T (exact)
    for (int j = tid; j < P; j += EQC_THREADS) {
        float y[EQC_ROWS];
#pragma unroll
        for (int r = 0; r < EQC_ROWS; ++r) y[r] = 0.f;
        for (int i = 0; i < PP; i += 4) {
            float gs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ii = i + u;
                gs[u] = (ii < P) ? (exact_sym ? __ldg(Gh + (size_t)ii * P + j) + __ldg(Gh + (size_t)j * P + ii)
                                              : 2.0f * __ldg(Gh + (size_t)ii * P + j)) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < EQC_ROWS; ++r) {
                const float4 jv = *reinterpret_cast<const float4*>(Js + r * PP + i);
                y[r] = fmaf(jv.x, gs[0], fmaf(jv.y, gs[1], fmaf(jv.z, gs[2], fmaf(jv.w, gs[3], y[r]))));
            }
        }
#pragma unroll
        for (int r = 0; r < EQC_ROWS; ++r) Ys[r * PP + j] = y[r];
    }
    __syncthreads();

    // Z = J ghat (per row), S = Y J^T (2x2 per pixel): warp per pixel
    for (int i = warp; i < EQC_TPX; i += EQC_THREADS / 32) {
        float z0 = 0.f, z1 = 0.f, s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
        for (int j = lane; j < P; j += 32) {
            const float j0v = Js[(2 * i) * PP + j], j1v = Js[(2 * i + 1) * PP + j];
            const float y0v = Ys[(2 * i) * PP + j], y1v = Ys[(2 * i + 1) * PP + j], gv = __ldg(gh + j);
            z0 = fmaf(j0v, gv, z0); z1 = fmaf(j1v, gv, z1);
            s00 = fmaf(y0v, j0v, s00); s01 = fmaf(y0v, j1v, s01); s10 = fmaf(y1v, j0v, s10); s11 = fmaf(y1v, j1v, s11);
        }
        z0 = warp_sum(z0); z1 = warp_sum(z1); s00 = warp_sum(s00); s01 = warp_sum(s01); s10 = warp_sum(s10); s11 = warp_sum(s11);
        if (lane == 0) { sZ[2 * i] = z0; sZ[2 * i + 1] = z1; sS[4 * i] = s00; sS[4 * i + 1] = s01; sS[4 * i + 2] = s10; sS[4 * i + 3] = s11; }
    }
    __syncthreads();

    // dJ = M Y + q ghat^T
    for (int e = tid; e < cnt * 2 * P; e += EQC_THREADS) {
        const int r = e / P, j = e - r * P, n = r >> 1, rr = r & 1;
        const float m_r0 = rr ? sMq[n * 5 + 1] : sMq[n * 5], m_r1 = rr ? sMq[n * 5 + 2] : sMq[n * 5 + 1];
        const float qv = sMq[n * 5 + 3 + rr];
        dJ[((size_t)b * N + n0) * 2 * P + e] = m_r0 * Ys[(2 * n) * PP + j] + m_r1 * Ys[(2 * n + 1) * PP + j] + qv * __ldg(gh + j);
    }
    // dd = G Z ; dG = G S + d Z^T
    for (int e = tid; e < cnt * C; e += EQC_THREADS) {
        const int n = e / C;
        const size_t gi = ((size_t)b * N + n0) * C + e;
        const float2 gv = *reinterpret_cast<const float2*>(G + 2 * gi);
        const float dv = d[gi];
        const float z0 = sZ[2 * n], z1 = sZ[2 * n + 1];
        dd[gi] = gv.x * z0 + gv.y * z1;
        float2 o;
        o.x = gv.x * sS[4 * n] + gv.y * sS[4 * n + 2] + dv * z0;
        o.y = gv.x * sS[4 * n + 1] + gv.y * sS[4 * n + 3] + dv * z1;
        *reinterpret_cast<float2*>(dG + 2 * gi) = o;
    }
}

int num_sms();

}  // namespace banet

using namespace banet;

extern "C" size_t banet_eqc_workspace_bytes(int nb, int N, int C, int P)
{
    (void)C;
    if (nb <= 0 || N <= 0 || P <= 0) return 0;
    return eqc_plan(nb, N, P, num_sms()).ws_bytes;
}

extern "C" int banet_eqc_fwd(const float* J, const float* G, const float* d, int nb, int N, int C, int P,
                             float* AtA, float* Atb, void* ws, size_t ws_bytes, banet_stream_t stream)
{
    BANET_REQUIRE(J && G && d && AtA && Atb, BANET_ERR_BAD_ARG, "eqc_fwd: null pointer");
    BANET_REQUIRE(nb > 0 && N > 0 && C > 0 && P > 0, BANET_ERR_BAD_ARG, "eqc_fwd: bad shape nb=%d N=%d C=%d P=%d", nb, N, C, P);
    BANET_REQUIRE(nb <= 65535, BANET_ERR_UNSUPPORTED, "eqc_fwd: nb=%d > 65535", nb);
    const EqcPlan pl = eqc_plan(nb, N, P, num_sms());
    BANET_REQUIRE(ws && ws_bytes >= pl.ws_bytes, BANET_ERR_WORKSPACE, "eqc_fwd: workspace %zu < %zu bytes", ws_bytes, pl.ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(ws);
    const dim3 grid(pl.nchunks, pl.nblk_tri, nb);
    const size_t smem = ((size_t)2 * EQC_ROWS * (pl.BT + 1) + EQC_TPX * 5) * sizeof(float);
    cudaError_t e;
    if (pl.T == 2) {
        eqc_fwd_kernel<2><<<grid, EQC_THREADS, smem, st>>>(J, G, d, N, C, P, pl.chunk_px, pl.nchunks, pl.nblk, partial);
    } else {
        e = cudaFuncSetAttribute(eqc_fwd_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("eqc_fwd smem attr: %s", cudaGetErrorString(e)); return BANET_ERR_CUDA; }
        eqc_fwd_kernel<9><<<grid, EQC_THREADS, smem, st>>>(J, G, d, N, C, P, pl.chunk_px, pl.nchunks, pl.nblk, partial);
    }
    BANET_CUDA_LAUNCH_CHECK("eqc_fwd_kernel launch");
    const int nel = P * P + P;
    eqc_reduce_kernel<<<dim3((nel + 255) / 256, nb), 256, 0, st>>>(partial, pl.nchunks, P, pl.BT, AtA, Atb);
    BANET_CUDA_LAUNCH_CHECK("eqc_reduce_kernel launch");
    return BANET_OK;
}

extern "C" int banet_eqc_bwd(const float* J, const float* G, const float* d, const float* gAtA, const float* gAtb,
                             int nb, int N, int C, int P, int exact_sym, float* dJ, float* dG, float* dd, banet_stream_t stream)
{
    BANET_REQUIRE(J && G && d && gAtA && gAtb && dJ && dG && dd, BANET_ERR_BAD_ARG, "eqc_bwd: null pointer");
    BANET_REQUIRE(nb > 0 && N > 0 && C > 0 && P > 0, BANET_ERR_BAD_ARG, "eqc_bwd: bad shape");
    BANET_REQUIRE(nb <= 65535, BANET_ERR_UNSUPPORTED, "eqc_bwd: nb=%d > 65535", nb);
    const int PP = (P + 3) / 4 * 4;
    const size_t smem = ((size_t)2 * EQC_ROWS * PP + EQC_TPX * 5 + EQC_ROWS + EQC_TPX * 4) * sizeof(float);
    BANET_REQUIRE(smem <= 220 * 1024, BANET_ERR_UNSUPPORTED, "eqc_bwd: P=%d too large for shared memory", P);
    cudaError_t e = cudaFuncSetAttribute(eqc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("eqc_bwd smem attr: %s", cudaGetErrorString(e)); return BANET_ERR_CUDA; }
    const dim3 grid((N + EQC_TPX - 1) / EQC_TPX, nb);
    eqc_bwd_kernel<<<grid, EQC_THREADS, smem, (cudaStream_t)stream>>>(J, G, d, gAtA, gAtb, N, C, P, PP, exact_sym, dJ, dG, dd);
    BANET_CUDA_LAUNCH_CHECK("eqc_bwd_kernel launch");
    return BANET_OK;
}
