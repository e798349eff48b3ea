// One kernel for everything of an LM iteration that is not the normal-equation build: lambda-MLP, damping, blocked Cholesky with the
// right-hand side carried as an extra row, blocked back substitution, SE(3) / depth-coefficient update.
//
// Replaces, per iteration, lm_lambda_kernel + lm_solve_kernel + pose_update_kernel (lm_solve.cu; reference bundlenet.py:241-253, 264-276):
// those are three launches of nb CTAs whose run time is pure dependent latency (measured round 1: 64 + 215..258 + 5 us, about 11 % of a cfg2
// solve).  Here: one launch, one CTA (1024 threads) per pair:
//   1. rbar = rbar_sum / N, ||rbar||; 5 dense layers C->2C->4C->2C->C->1 (selu x4, tanh): warps take (32 outputs x an input slice) tasks,
//      lanes over consecutive outputs (coalesced 128-B weight rows, 8 rows in flight per lane), slices combined through shared memory;
//      lambda = base * ||rbar||^(2 + tanh(.))                                                            (bundlenet.py:243-253)
//   2. packed lower triangle of H (+ damping on the diagonal) and g as row P of the same packed array, in S = double (P <= 200) or float;
//   3. right-looking Cholesky in panels of 4 columns: one thread per row of the panel (the block's own rows, the rows below, and row P = the
//      right-hand side, so that forward substitution comes for free); every row owner factors the 4x4 diagonal block redundantly IN REGISTERS and
//      solves its own row against it (a single thread factoring through shared memory was 37 % of the first version's run time, measured);
//      warp-per-row trailing update; 3 block barriers per panel instead of one per column;
//   4. back substitution L^T x = y panel by panel (4 warps form the 4 dot products of a panel, thread 0 solves the 4x4 triangle in registers);
//   5. delta, W' = W + delta_d, status; R' = exp(w) R, T' = V(w) t + exp(w) T in double by thread 0 (per-pair VMatrix).
// The reference's batch-interleaved VMatrix (vmatrix_batch_scramble, bundlenet.py:45) needs every pair's delta first: the host falls back to
// the three-kernel path for that option.
#include "common.cuh"
#include "lm_build.h"

namespace banet {

constexpr int STEP_THREADS = 1024;
constexpr int STEP_WARPS = STEP_THREADS / 32;
constexpr int STEP_NB = 4;                       // Cholesky panel width (the panel owners keep the NB x NB block in registers: 1024 threads leave 64 registers each)

__device__ __forceinline__ float selu_s(float x) {
    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
    return scale * (x > 0.f ? x : alpha * expm1f(x));
}
__host__ __device__ __forceinline__ int tri3(int i, int k) { return i * (i + 1) / 2 + k; }

// one dense layer: out[j] = act(bias[j] + sum_i in[i] W[i][j]);  W row-major [cin][cout].  Tasks = (32-output block) x (input slice); every task
// stores its partial sums into its own row of `part` ([slices][cout], at most max(1024, cout) floats) and the activation pass adds the slices in
// a fixed order: bit-reproducible (shared-memory float atomics were not: lambda, and with it the whole step, moved in the last ulp run to run).
__device__ __forceinline__ void dense_layer(const float* __restrict__ in, const float* __restrict__ Wm, const float* __restrict__ bias,
                                            int cin, int cout, bool last, float* __restrict__ part, float* __restrict__ out, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int jblocks = (cout + 31) / 32;
    int slices = STEP_WARPS / jblocks; if (slices < 1) slices = 1; if (slices > cin / 8) slices = max(1, cin / 8);
    const int ntask = jblocks * slices;
    const int rows = (cin + slices - 1) / slices;
    for (int t = warp; t < ntask; t += STEP_WARPS) {
        const int jb = t % jblocks, sl = t / jblocks;
        const int j = jb * 32 + lane, i0 = sl * rows, i1 = min(cin, i0 + rows);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (j < cout) {
            int i = i0;
            for (; i + 15 < i1; i += 16) {                       // 16 weight rows in flight per lane (the loop is pure L2 latency)
                float wv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) wv[u] = __ldg(Wm + (size_t)(i + u) * cout + j);
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    a0 = fmaf(in[i + u], wv[u], a0); a1 = fmaf(in[i + u + 1], wv[u + 1], a1); a2 = fmaf(in[i + u + 2], wv[u + 2], a2); a3 = fmaf(in[i + u + 3], wv[u + 3], a3);
                }
            }
            for (; i + 7 < i1; i += 8) {
                const float w0 = __ldg(Wm + (size_t)i * cout + j), w1 = __ldg(Wm + (size_t)(i + 1) * cout + j), w2 = __ldg(Wm + (size_t)(i + 2) * cout + j),
                            w3 = __ldg(Wm + (size_t)(i + 3) * cout + j), w4 = __ldg(Wm + (size_t)(i + 4) * cout + j), w5 = __ldg(Wm + (size_t)(i + 5) * cout + j),
                            w6 = __ldg(Wm + (size_t)(i + 6) * cout + j), w7 = __ldg(Wm + (size_t)(i + 7) * cout + j);
                a0 = fmaf(in[i], w0, a0); a1 = fmaf(in[i + 1], w1, a1); a2 = fmaf(in[i + 2], w2, a2); a3 = fmaf(in[i + 3], w3, a3);
                a0 = fmaf(in[i + 4], w4, a0); a1 = fmaf(in[i + 5], w5, a1); a2 = fmaf(in[i + 6], w6, a2); a3 = fmaf(in[i + 7], w7, a3);
            }
            for (; i < i1; ++i) a0 = fmaf(in[i], __ldg(Wm + (size_t)i * cout + j), a0);
            part[sl * cout + j] = (a0 + a1) + (a2 + a3);
        }
    }
    __syncthreads();
    for (int j = tid; j < cout; j += STEP_THREADS) {
        float z = part[j];
        for (int sl = 1; sl < slices; ++sl) z += part[sl * cout + j];
        z += __ldg(bias + j);
        out[j] = last ? tanhf(z) : selu_s(z);
    }
    __syncthreads();
}

template <typename S, bool FULL>
__global__ void __launch_bounds__(STEP_THREADS)
lm_step_kernel(const float* __restrict__ H, const float* __restrict__ g, const float* __restrict__ rbar_sum, int N, int C,
               const float* __restrict__ mlp, float base, const float* __restrict__ lambda_in, const StepMode mode, const float* __restrict__ nvalid,
               int P, float eps, int ndamped,
               const float* __restrict__ R, const float* __restrict__ T, const float* __restrict__ W,
               float* __restrict__ R_out, float* __restrict__ T_out, float* __restrict__ W_out,
               float* __restrict__ delta, float* __restrict__ lambda_out, int32_t* __restrict__ status, int status_accumulate)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    // lower triangle of the damped matrix, rows 0..P (row P = right-hand side).  FULL: square storage with an odd row pitch (column walks over
    // consecutive rows are bank-conflict free for 64-bit words; the packed triangle's varying row offsets were 2..4-way conflicted); else packed.
    S* A = reinterpret_cast<S*>(smraw);
    const int LD = (P + 1) | 1;
    const size_t nA = FULL ? (size_t)(P + 1) * LD : (size_t)(P + 1) * (P + 2) / 2;
    auto IX = [&](int i, int k) -> int { return FULL ? i * LD + k : i * (i + 1) / 2 + k; };
    S* xs = A + nA;                                                  // [P] solution
    S* dinv = xs + P;                                                // [P] reciprocals of the Cholesky diagonal
    S* dots = dinv + P;                                              // [STEP_NB]
    float* mbuf = reinterpret_cast<float*>(dots + STEP_NB);          // MLP buffers: 2 x 4C floats + max(4C, 1024) floats of slice partials
    __shared__ int s_flag;
    __shared__ float s_wpart[STEP_WARPS], s_lam;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K = P - 6;
    if (tid == 0) s_flag = 0;
    __syncthreads();

    // ---- 1. lambda -------------------------------------------------------------------------------------------------------------------
    float lam;
    if (!lambda_in) {
        float* bufA = mbuf; float* bufB = mbuf + 4 * C; float* acc = mbuf + 8 * C;
        // bundlenet.py:243 divides by N; the legacy tracker rescales by N / valid, i.e. divides by the in-bounds count (legacy/ba.py:256,274)
        const float invN = 1.0f / (mode.rbar_per_valid ? nvalid[b] : (float)N);
        float part = 0.f;
        for (int c = tid; c < C; c += STEP_THREADS) { const float r = rbar_sum[(size_t)b * C + c] * invN; bufA[c] = r; part += r * r; }
        part = warp_sum(part);
        if (lane == 0) s_wpart[warp] = part;
        __syncthreads();
        const int dims[6] = {C, 2 * C, 4 * C, 2 * C, C, 1};
        const float* wp = mlp;
        float* in = bufA; float* out = bufB;
        for (int l = 0; l < (mlp ? 5 : 0); ++l) {
            const int cin = dims[l], cout = dims[l + 1];
            dense_layer(in, wp, wp + (size_t)cin * cout, cin, cout, l == 4, acc, out, tid);
            wp += (size_t)cin * cout + cout;
            float* tmp = in; in = out; out = tmp;
            __syncthreads();
        }
        // bundlenet.py:249,253: base * ||rbar||^(2 + h); legacy/ba.py:280: ||rbar||^(1 + h); no MLP (legacy/ba.py:190): h = 0
        if (tid == 0) {
            float norm2 = 0.f;
            for (int wq = 0; wq < STEP_WARPS; ++wq) norm2 += s_wpart[wq];               // fixed order
            s_lam = base * powf(sqrtf(norm2), mode.lambda_exp0 + (mlp ? in[0] : 0.f)); lambda_out[b] = s_lam;
        }
        __syncthreads();
        lam = s_lam;
    } else {
        lam = lambda_in[b];
        if (tid == 0) lambda_out[b] = lam;
    }

    // ---- 2. load (+ damping, bundlenet.py:264-266 / :181-182) -------------------------------------------------------------------------------
    const float* Hb = H + (size_t)b * P * P;
    int bad = 0;
    for (int i = warp; i < P; i += STEP_WARPS)
        for (int k = lane; k <= i; k += 32) {
            const float v = Hb[(size_t)i * P + k];
            if (!isfinite(v)) bad = 1;
            S sv = (S)v;
            if (k == i && i < ndamped) sv += ((S)v + (S)eps) * (S)lam;
            A[IX(i, k)] = sv;
        }
    for (int k = tid; k < P; k += STEP_THREADS) { const float v = g[(size_t)b * P + k]; if (!isfinite(v)) bad = 1; A[IX(P, k)] = (S)v; }
    if (!isfinite(lam)) bad = 1;
    if (bad) atomicOr(&s_flag, 2);
    __syncthreads();

    // ---- 3. blocked Cholesky; row P rides along: afterwards A[P][:] = y = L^-1 g ----------------------------------------------------------------
    for (int j0 = 0; j0 < P; j0 += STEP_NB) {
        const int jb = min(STEP_NB, P - j0);
        // Panel: thread t owns row j0 + t (the block's own rows first, then the rows below, then the rhs row P).  EVERY owner factors the 8x8 diagonal
        // block redundantly in registers (36 independent loads, then a register-only chain) instead of waiting for one thread to do it through
        // shared memory (measured: that serial section and its barrier were 37 % of the kernel), then solves its own row against it.
        {
            const int i = j0 + tid;
            const bool owner = i <= P;
            S L[STEP_NB][STEP_NB], row[STEP_NB], Linv[STEP_NB];
            if (owner) {
#pragma unroll
                for (int r = 0; r < STEP_NB; ++r)
#pragma unroll
                    for (int c = 0; c < STEP_NB; ++c) L[r][c] = (c <= r && r < jb) ? A[IX(j0 + r, j0 + c)] : (S)(r == c ? 1 : 0);
#pragma unroll
                for (int c = 0; c < STEP_NB; ++c) row[c] = (c < jb && (tid >= jb || c <= tid)) ? A[IX(i, j0 + c)] : (S)0;
            }
            __syncthreads();                                         // every owner has read the block before its rows are overwritten
            if (owner) {
                bool notpd = false;
#pragma unroll
                for (int c = 0; c < STEP_NB; ++c) {                  // unblocked Cholesky of the block, registers only
                    S d = L[c][c];
#pragma unroll
                    for (int m = 0; m < STEP_NB; ++m) if (m < c) d -= L[c][m] * L[c][m];
                    if (c < jb && !(d > (S)0)) { notpd = true; d = (S)1; }
                    const S inv = rsqrt(d), ld = d * inv;                // one reciprocal square root per column; no divisions anywhere on the path
                    L[c][c] = ld; Linv[c] = inv;
#pragma unroll
                    for (int r = 0; r < STEP_NB; ++r) {
                        if (r > c) {
                            S v = L[r][c];
#pragma unroll
                            for (int m = 0; m < STEP_NB; ++m) if (m < c) v -= L[r][m] * L[c][m];
                            L[r][c] = v * inv;
                        }
                    }
                }
                if (tid == 0 && notpd) s_flag |= 1;
                if (tid < jb) {                                      // a block row: its part of the factor
#pragma unroll
                    for (int c = 0; c < STEP_NB; ++c) if (c <= tid) { S v = (S)0;
#pragma unroll
                        for (int r = 0; r < STEP_NB; ++r) if (r == tid) v = L[r][c];
                        A[IX(i, j0 + c)] = v; }
#pragma unroll
                    for (int c = 0; c < STEP_NB; ++c) if (c == tid) dinv[j0 + c] = Linv[c];
                } else {                                             // a row below (or the rhs row): L21[i][:] = A21[i][:] L11^-T
#pragma unroll
                    for (int c = 0; c < STEP_NB; ++c) {
                        if (c < jb) {
                            S v = row[c];
#pragma unroll
                            for (int m = 0; m < STEP_NB; ++m) if (m < c) v -= row[m] * L[c][m];
                            row[c] = v * Linv[c];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < STEP_NB; ++c) if (c < jb) A[IX(i, j0 + c)] = row[c];
                }
            }
        }
        __syncthreads();
        for (int i = j0 + jb + warp; i <= P; i += STEP_WARPS) {      // trailing update, warp per row, lanes over columns
            S li[STEP_NB];
#pragma unroll
            for (int c = 0; c < STEP_NB; ++c) li[c] = (c < jb) ? A[IX(i, j0 + c)] : (S)0;
            const int kend = (i == P) ? P - 1 : i;
            for (int k0 = j0 + jb; k0 <= kend; k0 += 128) {         // 4 independent column chunks in flight (no store between their loads)
                S sacc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = k0 + 32 * q + lane;
                    S sv = (S)0;
                    if (k <= kend) {
#pragma unroll
                        for (int c = 0; c < STEP_NB; ++c) if (c < jb) sv += li[c] * A[IX(k, j0 + c)];
                    }
                    sacc[q] = sv;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { const int k = k0 + 32 * q + lane; if (k <= kend) A[IX(i, k)] -= sacc[q]; }
            }
        }
        __syncthreads();
    }

    // ---- 4. back substitution L^T x = y, panels from the bottom ----------------------------------------------------------------------------
    const int npan = (P + STEP_NB - 1) / STEP_NB;
    for (int pnl = npan - 1; pnl >= 0; --pnl) {
        const int j0 = pnl * STEP_NB, jb = min(STEP_NB, P - j0);
        if (warp < jb) {                                             // dots[c] = sum_{i >= j0+jb} L[i][j0+c] x[i]
            S s = (S)0;
            for (int i = j0 + jb + lane; i < P; i += 32) s += A[IX(i, j0 + warp)] * xs[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) dots[warp] = s;
        }
        __syncthreads();
        if (tid == 0) {                                              // 8x8 triangle in registers (independent loads first)
            S L[STEP_NB][STEP_NB], y[STEP_NB], x[STEP_NB];
#pragma unroll
            for (int r = 0; r < STEP_NB; ++r)
#pragma unroll
                for (int c = 0; c < STEP_NB; ++c) L[r][c] = (c <= r && r < jb) ? A[IX(j0 + r, j0 + c)] : (S)(r == c ? 1 : 0);
#pragma unroll
            for (int c = 0; c < STEP_NB; ++c) y[c] = (c < jb) ? A[IX(P, j0 + c)] - dots[c] : (S)0;
#pragma unroll
            for (int c = STEP_NB - 1; c >= 0; --c) {
                S v = y[c];
#pragma unroll
                for (int m = 0; m < STEP_NB; ++m) if (m > c) v -= L[m][c] * x[m];
                x[c] = v * (c < jb ? dinv[j0 + c] : (S)1);
            }
#pragma unroll
            for (int c = 0; c < STEP_NB; ++c) if (c < jb) xs[j0 + c] = x[c];
        }
        __syncthreads();
    }

    // ---- 5. outputs ----------------------------------------------------------------------------------------------------------------------------
    const int flag = s_flag;
    for (int i = tid; i < P; i += STEP_THREADS) {
        float dv = flag ? 0.f : (float)xs[i];
        if (!isfinite(dv)) dv = 0.f;
        delta[(size_t)b * P + i] = dv;
        if (i >= 6) W_out[(size_t)b * K + i - 6] = W[(size_t)b * K + i - 6] + dv;      // bundlenet.py:276
    }
    if (tid == 0) {
        status[b] = status_accumulate ? (status[b] | flag) : flag;
        double dl[6];
        for (int i = 0; i < 6; ++i) { float dv = flag ? 0.f : (float)xs[i]; if (!isfinite(dv)) dv = 0.f; dl[i] = (double)dv; }
        const double wx = dl[0], wy = dl[1], wz = dl[2], tx = dl[3], ty = dl[4], tz = dl[5];
        const double th_raw = sqrt(wx * wx + wy * wy + wz * wz);
        const double th = mode.clamp_theta ? fmax(th_raw, 1e-6) : fmax(th_raw, 1e-300);      // AngleaAxisRotation (bundlenet.py:17-37; legacy/ba.py:60-80 has no clamp)
        const double kx = wx / th, ky = wy / th, kz = wz / th, c = cos(th), s = sin(th), oc = 1.0 - c;
        const double dr[9] = {c + kx * kx * oc,      kx * ky * oc - kz * s, ky * s + kx * kz * oc,
                              kz * s + kx * ky * oc, c + ky * ky * oc,      -kx * s + ky * kz * oc,
                              -ky * s + kx * kz * oc, kx * s + ky * kz * oc, c + kz * kz * oc};
        double ca, cb;                                               // VMatrix (bundlenet.py:39-46), series below 1e-4
        if (th_raw < 1e-4) { ca = 0.5 - th_raw * th_raw / 24.0; cb = 1.0 / 6.0 - th_raw * th_raw / 120.0; }
        else { ca = (1.0 - cos(th_raw)) / (th_raw * th_raw); cb = (th_raw - sin(th_raw)) / (th_raw * th_raw * th_raw); }
        const double sk[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
        double V[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            const double sk2 = sk[i * 3] * sk[j] + sk[i * 3 + 1] * sk[3 + j] + sk[i * 3 + 2] * sk[6 + j];
            V[i * 3 + j] = ((i == j) ? 1.0 : 0.0) + ca * sk[i * 3 + j] + cb * sk2;
        }
        double Rin[9], Tin[3];
        for (int q = 0; q < 9; ++q) Rin[q] = R[(size_t)b * 9 + q];
        for (int q = 0; q < 3; ++q) Tin[q] = T[(size_t)b * 3 + q];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j)
                R_out[(size_t)b * 9 + i * 3 + j] = (float)(dr[i * 3] * Rin[j] + dr[i * 3 + 1] * Rin[3 + j] + dr[i * 3 + 2] * Rin[6 + j]);
            const double vt = mode.use_vmatrix ? (V[i * 3] * tx + V[i * 3 + 1] * ty + V[i * 3 + 2] * tz) : dl[3 + i];     // legacy/ba.py:213: T' = t + dr T
            T_out[(size_t)b * 3 + i] = (float)(vt + dr[i * 3] * Tin[0] + dr[i * 3 + 1] * Tin[1] + dr[i * 3 + 2] * Tin[2]);
        }
    }
}

size_t lm_step_smem(int P, int C, bool use_double, bool full)
{
    const size_t nA = (full ? (size_t)(P + 1) * ((P + 1) | 1) : (size_t)(P + 1) * (P + 2) / 2) + 2 * (size_t)P + STEP_NB;
    return nA * (use_double ? sizeof(double) : sizeof(float)) + ((size_t)8 * C + (4 * C > 1024 ? 4 * C : 1024)) * sizeof(float);
}

bool lm_step_supported(int P, int C) { return lm_step_smem(P, C, false, false) <= 220 * 1024; }

// lambda_in != nullptr: used as is; else lambda = base * ||rbar||^(exp0 + MLP(rbar)) (MLP term 0 when mlp == nullptr).
// In-place R/T/W (R_out == R ...) is fine: a pair's CTA reads before it writes.
int lm_step(const float* H, const float* g, const float* rbar_sum, int nb, int N, int C, int K, const float* mlp, float base, const float* lambda_in,
            const StepMode& mode, const float* nvalid, const banet_solve_opts_t& opts, const float* R, const float* T, const float* W, float* R_out, float* T_out, float* W_out,
            float* delta, float* lambda_out, int32_t* status, int status_accumulate, cudaStream_t st)
{
    const int P = 6 + K;
    const int ndamped = opts.undamped_last ? P - 1 : P;
    // storage: square double (P <= ~150), packed double (P <= ~200), packed float beyond
    const bool full = lm_step_smem(P, C, true, true) <= 200 * 1024;
    const bool use_double = full || lm_step_smem(P, C, true, false) <= 200 * 1024;
    const size_t smem = lm_step_smem(P, C, use_double, full);
    BANET_REQUIRE(smem <= 220 * 1024, BANET_ERR_UNSUPPORTED, "lm_step: P=%d, C=%d do not fit shared memory", P, C);
    auto launch = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("lm_step smem attr: %s", cudaGetErrorString(e)); return BANET_ERR_CUDA; }
        kern<<<nb, STEP_THREADS, smem, st>>>(H, g, rbar_sum, N, C, mlp, base, lambda_in, mode, nvalid, P, opts.damping_eps, ndamped, R, T, W,
                                             R_out, T_out, W_out, delta, lambda_out, status, status_accumulate);
        return BANET_OK;
    };
    int rc;
    if (full) rc = launch(lm_step_kernel<double, true>);
    else if (use_double) rc = launch(lm_step_kernel<double, false>);
    else rc = launch(lm_step_kernel<float, false>);
    if (rc) return rc;
    BANET_CUDA_LAUNCH_CHECK("lm_step_kernel launch");
    return BANET_OK;
}

}  // namespace banet
