// Fused normal-equation construction, tensor-core path (sm_100a: TMA + tcgen05 + TMEM), K = 128.
//
// Same maths and partial-slot contract as lm_build.cu (reference bundlenet.py:206-263 + utils.cu:219-417),
// but the basis contraction runs on the 5th-gen tensor cores:
//
//     D[128 x 160] += Bt^T R          per 64-pixel tile, kind::tf32, fp32 accumulate in TMEM
//        Bt [64 px x 128]  the basis tile exactly as it lies in HBM (TMA, 128B/32B-atom swizzle) = MN-major "A"
//        R  [64 px x 160]  row n = [ s_n * b_n (128) | v_n (6) | t_n | 0 ... ]  built by the gather warps = MN-major "B"
//     => D[i][j<128] = H_dd[i][j],  D[i][128+r] = H_cd[r][i] (r<6),  D[i][134] = g_d[i]
//
// Precision: MODE 1 = one tf32 pass (A truncated by the tensor core, R rounded to nearest);
//            MODE 2 = split-A: a second pass with A_lo = b - trunc(b), so only R's (unbiased) rounding remains.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over a contiguous tile range):
//   warp 0    TMA producer: basis tile -> smem stage (mbarrier complete_tx)
//   warp 1    MMA issuer (one thread): 8 (16) tcgen05.mma per tile, tcgen05.commit frees the stage
//   warps 2-9 gather warps: half-warp per pixel, lanes over channels: D~ = D + b.W (from the staged tile), warp,
//             4-tap (or 12-texel, F2-only) gather, M/q reductions, per-pixel 2x7 algebra, R rows; at a pair
//             boundary they drain TMEM (tcgen05.ld) into the partial slot.
#include "common.cuh"
#include "lm_build.h"
#include "tc_utils.cuh"
#include "tmap.h"

namespace banet {
using namespace tc;

constexpr int TC_TILE = 64;
constexpr int TC_GW = 8;
constexpr int TC_THREADS = (2 + TC_GW) * 32;
constexpr int TC_K = 128;
constexpr int TC_N = 160;
constexpr int TC_STAGE_A = 4 * TC_TILE * 128;      // 32 KB: 4 blocks of [64 rows][128 B]
constexpr int TC_STAGE_R = 5 * TC_TILE * 128;      // 40 KB
constexpr int TC_REC = 16;                         // floats per pixel record

template <int MODE> struct TcSmem {
    static constexpr int off_A = 0;
    static constexpr int off_R = 2 * TC_STAGE_A;
    static constexpr int off_Alo = off_R + 2 * TC_STAGE_R;
    static constexpr int off_misc = off_Alo + (MODE == 2 ? 2 * TC_STAGE_A : 0);
    static constexpr int off_bar = off_misc;                       // 8 mbarriers
    static constexpr int off_tmem = off_bar + 64;
    static constexpr int off_pose = off_misc + 128;                // [2][16] floats
    static constexpr int off_W = off_pose + 128;                   // [2][128] floats
    static constexpr int off_rec = off_W + 1024;                   // [GW][8][TC_REC] floats
    static constexpr int off_cc = off_rec + TC_GW * 8 * TC_REC * 4;   // [GW][8][28] floats
    static constexpr int total = off_cc + TC_GW * 8 * 28 * 4;
    static constexpr int bytes = total + 1024;                     // slack for manual 1024-B alignment
};

__device__ __forceinline__ void gather_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ int reflect_i(int i, int n) { i = i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); return i < 0 ? 0 : i; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float hsum16(float v) {            // sum over the 16 lanes of a half-warp
    v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

template <int NCH, bool FLY, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
lm_build_tc_kernel(const __grid_constant__ CUtensorMap tmapB, const BuildParams prm)
{
    using SM = TcSmem<MODE>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + SM::off_bar);
    uint64_t* fullB = bars;          // [2]
    uint64_t* ready = bars + 2;      // [2]
    uint64_t* empty = bars + 4;      // [2]
    uint64_t* flushb = bars + 6;
    uint64_t* tmemfree = bars + 7;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(base + SM::off_tmem);
    float* sPose = reinterpret_cast<float*>(base + SM::off_pose);
    float* sW = reinterpret_cast<float*>(base + SM::off_W);
    float* sRec = reinterpret_cast<float*>(base + SM::off_rec);
    float* sCC = reinterpret_cast<float*>(base + SM::off_cc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = prm.N, h = prm.h, w = prm.w, c2 = prm.c2;
    constexpr int C = 64 * NCH;
    const long long t_begin = part_begin(prm.total_tiles, gridDim.x, blockIdx.x);
    const long long t_end   = part_begin(prm.total_tiles, gridDim.x, blockIdx.x + 1);

    if (tid == 0) {
        mbar_init(&fullB[0], 1); mbar_init(&fullB[1], 1);
        mbar_init(&ready[0], TC_GW); mbar_init(&ready[1], TC_GW);
        mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
        mbar_init(flushb, 1); mbar_init(tmemfree, TC_GW);
        fence_barrier_init();
        prefetch_tmap(&tmapB);
    }
    if (warp == 0) tmem_alloc<256>(s_tmem);
    // the pad chunks of R's 5th block (columns 136..159) stay zero for the whole kernel
    for (int i = tid; i < 2 * TC_TILE * 8; i += TC_THREADS) {
        const int s = i / (TC_TILE * 8), r = (i / 8) % TC_TILE, c = i & 7;
        *reinterpret_cast<float4*>(base + SM::off_R + s * TC_STAGE_R + 4 * 8192 + sw128_32b_off(r, c)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int it = 0;
            for (long long t = t_begin; t < t_end; ++t, ++it) {
                const int s = it & 1, ph = (it >> 1) & 1;
                const int b = (int)(t / prm.tiles_per_pair);
                const int n0 = (int)(t - (long long)b * prm.tiles_per_pair) * TC_TILE;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&fullB[s], TC_STAGE_A);
                const int row = b * N + n0;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk)
                    tma_load_2d(base + SM::off_A + s * TC_STAGE_A + blk * 8192, &tmapB, blk * 32, row, &fullB[s]);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32_mn_mn(128, TC_N);
            int it = 0, span = 0, cur_b = -1;
            uint32_t acc = 0;
            for (long long t = t_begin; t < t_end; ++t, ++it) {
                const int s = it & 1, ph = (it >> 1) & 1;
                const int b = (int)(t / prm.tiles_per_pair);
                if (b != cur_b) {
                    if (cur_b >= 0) { mma_commit(flushb); ++span; }
                    mbar_wait(tmemfree, (span & 1) ^ 1);          // TMEM drained by the previous span's flush
                    tc_fence_after_sync();
                    acc = 0; cur_b = b;
                }
                mbar_wait(&ready[s], ph);
                tc_fence_after_sync();
                const uint32_t r0 = smem_u32(base + SM::off_R + s * TC_STAGE_R);
#pragma unroll
                for (int pass = 0; pass < (MODE == 2 ? 2 : 1); ++pass) {
                    const uint32_t a0 = smem_u32(base + (pass ? SM::off_Alo : SM::off_A) + s * TC_STAGE_A);
#pragma unroll
                    for (int kk = 0; kk < TC_TILE / 8; ++kk) {
                        mma_tf32_ss(tmem, make_desc_mn_sw128_32b(a0 + kk * 1024, 8192, 512),
                                    make_desc_mn_sw128_32b(r0 + kk * 1024, 8192, 512), idesc, acc);
                        acc = 1;
                    }
                }
                mma_commit(&empty[s]);
            }
            if (cur_b >= 0) mma_commit(flushb);
        }
    } else {
        // ===================================================================== gather warps
        const int g = warp - 2, hw = lane >> 4, hl = lane & 15;
        const int gtid = tid - 64;
        const int blkA = hl >> 3, ccA = hl & 7;                 // this lane's two 16-B chunks of a 128-float row: blocks blkA and 2+blkA
        float wreg[8];
        float rb[NCH * 4];
        float* myRec = sRec + g * 8 * TC_REC;
        float* myCC = sCC + (g * 8 + (lane & 7)) * 28;
        const SlotLayout L{TC_K, C};
        int it = 0, span = 0, cur_b = -1;

        auto flush = [&](int sp) {
            float* slot = prm.partials + ((size_t)blockIdx.x * prm.max_span + sp) * prm.slot_floats;
            mbar_wait(flushb, sp & 1);
            tc_fence_after_sync();
            const int q = warp & 3, row = q * 32 + lane;
            const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
            float v[32];
#pragma unroll 1
            for (int cbi = 0; cbi < 2; ++cbi) {
                const int cb = (g >= 4 ? 2 : 0) + cbi;
                tmem_ld_32x32(tq + cb * 32, v);
                float4* dst = reinterpret_cast<float4*>(slot + (size_t)row * TC_K + cb * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (g < 4) {
                tmem_ld_32x32(tq + 128, v);
#pragma unroll
                for (int r = 0; r < 7; ++r) slot[L.off_ext() + r * TC_K + row] = v[r];
            }
            tc_fence_before_sync();
            // rbar / cc through a scratch aliased on R stage 0 (every MMA of this span has completed)
            float* scratch = reinterpret_cast<float*>(base + SM::off_R);
#pragma unroll
            for (int u = 0; u < NCH * 4; ++u) rb[u] += __shfl_xor_sync(0xffffffffu, rb[u], 16);
            if (hw == 0) {
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    *reinterpret_cast<float4*>(scratch + g * C + 64 * j + 4 * hl) = make_float4(rb[4 * j], rb[4 * j + 1], rb[4 * j + 2], rb[4 * j + 3]);
            }
            gather_bar();
            if (gtid < C) {
                float s = 0.f;
#pragma unroll
                for (int wq = 0; wq < TC_GW; ++wq) s += scratch[wq * C + gtid];
                slot[L.off_rbar() + gtid] = s;
            }
            if (gtid < 28) {
                float s = 0.f;
                for (int e = 0; e < TC_GW * 8; ++e) s += sCC[e * 28 + gtid];
                slot[L.off_cc() + gtid] = s;
            }
            gather_bar();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmemfree);
        };

        for (long long t = t_begin; t < t_end; ++t, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            const int b = (int)(t / prm.tiles_per_pair);
            const int n0 = (int)(t - (long long)b * prm.tiles_per_pair) * TC_TILE;
            const int cnt = min(TC_TILE, N - n0);
            if (b != cur_b) {
                if (cur_b >= 0) { flush(span); ++span; }
                float* pose = sPose + (span & 1) * 16;
                float* Wsm = sW + (span & 1) * TC_K;
                if (g == 0) {
                    if (lane < 9) pose[lane] = prm.R[b * 9 + lane];
                    else if (lane < 12) pose[lane] = prm.T[b * 3 + lane - 9];
                    else if (lane < 16) pose[lane] = prm.intr[b * 4 + lane - 12];
                    for (int k = lane; k < TC_K; k += 32) Wsm[k] = prm.W[b * TC_K + k];
                }
                gather_bar();
                {
                    const float4 w0 = *reinterpret_cast<const float4*>(Wsm + blkA * 32 + ccA * 4);
                    const float4 w1 = *reinterpret_cast<const float4*>(Wsm + (2 + blkA) * 32 + ccA * 4);
                    wreg[0] = w0.x; wreg[1] = w0.y; wreg[2] = w0.z; wreg[3] = w0.w;
                    wreg[4] = w1.x; wreg[5] = w1.y; wreg[6] = w1.z; wreg[7] = w1.w;
                }
                if (lane < 8) {
#pragma unroll
                    for (int q = 0; q < 28; ++q) myCC[q] = 0.f;
                }
#pragma unroll
                for (int u = 0; u < NCH * 4; ++u) rb[u] = 0.f;
                cur_b = b;
            }
            const float* pose = sPose + (span & 1) * 16;
            const unsigned char* As = base + SM::off_A + s * TC_STAGE_A;
            unsigned char* Rs = base + SM::off_R + s * TC_STAGE_R;

            mbar_wait(&fullB[s], ph);

            // ---------------------------------------------------------------- gather: 4 x (2 pixels per warp)
#pragma unroll 1
            for (int i4 = 0; i4 < 4; ++i4) {
                const int pl = i4 * 2 + hw;                      // pixel slot within this warp (0..7)
                const int nl = g * 8 + pl;                       // pixel within the tile
                const bool in_tile = nl < cnt;
                const uint32_t offA = blkA * 8192 + sw128_32b_off(nl, ccA);
                const float4 b0 = *reinterpret_cast<const float4*>(As + offA);
                const float4 b1 = *reinterpret_cast<const float4*>(As + offA + 2 * 8192);
                float dot = b0.x * wreg[0] + b0.y * wreg[1] + b0.z * wreg[2] + b0.w * wreg[3]
                          + b1.x * wreg[4] + b1.y * wreg[5] + b1.z * wreg[6] + b1.w * wreg[7];
                dot = hsum16(dot);
                float mask = 0.f, x = 0.f, y = 0.f, iZ = 0.f, rx = 0.f, ry = 0.f, rz = 0.f, dx = 0.f, dy = 0.f;
                int x0 = 0, y0 = 0;
                if (in_tile) {
                    const size_t gi = (size_t)b * N + n0 + nl;
                    const float* pp = prm.p + (size_t)b * 3 * N + n0 + nl;
                    const float p0 = __ldg(pp), p1 = __ldg(pp + N), p2 = __ldg(pp + 2 * (size_t)N);
                    const float Dt = __ldg(prm.D + gi) + dot;                                   // bundlenet.py:208
                    rx = pose[0] * p0 + pose[1] * p1 + pose[2] * p2;
                    ry = pose[3] * p0 + pose[4] * p1 + pose[5] * p2;
                    rz = pose[6] * p0 + pose[7] * p1 + pose[8] * p2;
                    const float X = rx * Dt + pose[9], Y = ry * Dt + pose[10], Z = rz * Dt + pose[11];
                    x = X / Z; y = Y / Z; iZ = 1.0f / Z;
                    const float u = pose[12] * x + pose[14], v = pose[13] * y + pose[15];
                    if ((u >= 0.f) && (u <= (float)(w - 1)) && (v >= 0.f) && (v <= (float)(h - 1)) && isfinite(iZ)) {
                        mask = 1.f;
                        const float fu = floorf(u), fv = floorf(v);
                        x0 = (int)fu; y0 = (int)fv; dx = u - fu; dy = v - fv;
                    }
                }
                float m11 = 0.f, m12 = 0.f, m22 = 0.f, q1 = 0.f, q2 = 0.f;
                if (mask != 0.f) {
                    const float w00 = (1.f - dx) * (1.f - dy), w01 = dx * (1.f - dy), w10 = (1.f - dx) * dy, w11 = dx * dy;
                    const float* img = prm.conv2 + (size_t)b * h * w * c2;
                    const float* c1 = prm.conv1 + ((size_t)b * N + n0 + nl) * C + 4 * hl;
                    if constexpr (!FLY) {
                        const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
                        const float* t00 = img + ((size_t)y0 * w + x0) * c2 + 4 * hl;
                        const float* t01 = img + ((size_t)y0 * w + x1) * c2 + 4 * hl;
                        const float* t10 = img + ((size_t)y1 * w + x0) * c2 + 4 * hl;
                        const float* t11 = img + ((size_t)y1 * w + x1) * c2 + 4 * hl;
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
                            const int co = 64 * j;
                            const float4 f1 = ld_stream_f4(c1 + co);
                            const float4 a00 = ldg4(t00 + co), a01 = ldg4(t01 + co), a10 = ldg4(t10 + co), a11 = ldg4(t11 + co);
                            const float4 g00 = ldg4(t00 + C + co), g01 = ldg4(t01 + C + co), g10 = ldg4(t10 + C + co), g11 = ldg4(t11 + C + co);
                            const float4 e00 = ldg4(t00 + 2 * C + co), e01 = ldg4(t01 + 2 * C + co), e10 = ldg4(t10 + 2 * C + co), e11 = ldg4(t11 + 2 * C + co);
#define BANET_CH(F)                                                                                              \
                            {                                                                                    \
                                const float f2 = w00 * a00.F + w01 * a01.F + w10 * a10.F + w11 * a11.F;          \
                                const float gx = w00 * g00.F + w01 * g01.F + w10 * g10.F + w11 * g11.F;          \
                                const float gy = w00 * e00.F + w01 * e01.F + w10 * e10.F + w11 * e11.F;          \
                                const float d = f1.F - f2;                                                       \
                                m11 = fmaf(gx, gx, m11); m12 = fmaf(gx, gy, m12); m22 = fmaf(gy, gy, m22);       \
                                q1 = fmaf(gx, d, q1); q2 = fmaf(gy, d, q2);                                      \
                                rb[4 * j + ci] += fabsf(d); ++ci;                                                \
                            }
                            int ci = 0;
                            BANET_CH(x) BANET_CH(y) BANET_CH(z) BANET_CH(w)
#undef BANET_CH
                        }
                    } else {
                        // F2-only map: central differences with REFLECT-by-one borders (bundlenet.py:92-100) folded into the
                        // bilinear blend: 12 texels = rows y0,Y1 x columns XM,x0,X1,XP  +  rows YM,YP x columns x0,X1
                        const int X1 = reflect_i(x0 + 1, w), XM = reflect_i(x0 - 1, w), XP = reflect_i(x0 + 2, w);
                        const int Y1 = reflect_i(y0 + 1, h), YM = reflect_i(y0 - 1, h), YP = reflect_i(y0 + 2, h);
                        const float* r0 = img + (size_t)y0 * w * c2 + 4 * hl;
                        const float* r1 = img + (size_t)Y1 * w * c2 + 4 * hl;
                        const float* rm = img + (size_t)YM * w * c2 + 4 * hl;
                        const float* rp = img + (size_t)YP * w * c2 + 4 * hl;
                        const size_t oM = (size_t)XM * c2, o0 = (size_t)x0 * c2, o1 = (size_t)X1 * c2, oP = (size_t)XP * c2;
                        const float h00 = 0.5f * w00, h01 = 0.5f * w01, h10 = 0.5f * w10, h11 = 0.5f * w11;
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
                            const int co = 64 * j;
                            const float4 f1 = ld_stream_f4(c1 + co);
                            const float4 aM0 = ldg4(r0 + oM + co), a00 = ldg4(r0 + o0 + co), a10 = ldg4(r0 + o1 + co), aP0 = ldg4(r0 + oP + co);
                            const float4 aM1 = ldg4(r1 + oM + co), a01 = ldg4(r1 + o0 + co), a11 = ldg4(r1 + o1 + co), aP1 = ldg4(r1 + oP + co);
                            const float4 a0m = ldg4(rm + o0 + co), a1m = ldg4(rm + o1 + co), a0p = ldg4(rp + o0 + co), a1p = ldg4(rp + o1 + co);
                            // naming: aXY = F[column X in {M,0,1,P}][row Y in {m,0,1,p}]
#define BANET_CH(F)                                                                                              \
                            {                                                                                    \
                                const float f2 = w00 * a00.F + w01 * a10.F + w10 * a01.F + w11 * a11.F;          \
                                const float gx = h00 * (a10.F - aM0.F) + h01 * (aP0.F - a00.F)                   \
                                               + h10 * (a11.F - aM1.F) + h11 * (aP1.F - a01.F);                  \
                                const float gy = h00 * (a01.F - a0m.F) + h10 * (a0p.F - a00.F)                   \
                                               + h01 * (a11.F - a1m.F) + h11 * (a1p.F - a10.F);                  \
                                const float d = f1.F - f2;                                                       \
                                m11 = fmaf(gx, gx, m11); m12 = fmaf(gx, gy, m12); m22 = fmaf(gy, gy, m22);       \
                                q1 = fmaf(gx, d, q1); q2 = fmaf(gy, d, q2);                                      \
                                rb[4 * j + ci] += fabsf(d); ++ci;                                                \
                            }
                            int ci = 0;
                            BANET_CH(x) BANET_CH(y) BANET_CH(z) BANET_CH(w)
#undef BANET_CH
                        }
                    }
                }
                m11 = hsum16(m11); m12 = hsum16(m12); m22 = hsum16(m22); q1 = hsum16(q1); q2 = hsum16(q2);
                if (hl == 0) {
                    float4* rp4 = reinterpret_cast<float4*>(myRec + pl * TC_REC);
                    rp4[0] = make_float4(m11, m12, m22, q1);
                    rp4[1] = make_float4(q2, x, y, iZ);
                    rp4[2] = make_float4(rx, ry, rz, mask);
                }
            }
            __syncwarp();

            // ---------------------------------------------------------------- per-pixel 2x7 algebra (lanes 0..7), bundlenet.py:49-74
            if (lane < 8) {
                float* rec = myRec + lane * TC_REC;
                const float4 ra = *reinterpret_cast<const float4*>(rec), rbq = *reinterpret_cast<const float4*>(rec + 4),
                             rc = *reinterpret_cast<const float4*>(rec + 8);
                float ext[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (rc.w != 0.f) {
                    const float m11 = ra.x, m12 = ra.y, m22 = ra.z, q1 = ra.w, q2 = rbq.x, x = rbq.y, y = rbq.z, iZ = rbq.w;
                    const float rx = rc.x, ry = rc.y, rz = rc.z;
                    const float fx = pose[12], fy = pose[13];
                    const float a0[6] = {-fx * (x * y), -fx * (-1.f - x * x), -fx * y, -fx * (-iZ), 0.f, -fx * (x * iZ)};
                    const float a1[6] = {-fy * (1.f + y * y), -fy * (-(x * y)), -fy * (-x), 0.f, -fy * (-iZ), -fy * (y * iZ)};
                    float ux[6], uy[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { ux[i] = m11 * a0[i] + m12 * a1[i]; uy[i] = m12 * a0[i] + m22 * a1[i]; }
                    int q = 0;
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int jj = i; jj < 6; ++jj) { myCC[q] += a0[i] * ux[jj] + a1[i] * uy[jj]; ++q; }
#pragma unroll
                    for (int i = 0; i < 6; ++i) myCC[21 + i] += a0[i] * q1 + a1[i] * q2;
                    myCC[27] += 1.f;
                    const float jd0 = fx * ((rx - rz * x) * iZ), jd1 = fy * ((ry - rz * y) * iZ);
                    const float u0 = m11 * jd0 + m12 * jd1, u1 = m12 * jd0 + m22 * jd1;
#pragma unroll
                    for (int i = 0; i < 6; ++i) ext[i] = a0[i] * u0 + a1[i] * u1;
                    ext[6] = jd0 * q1 + jd1 * q2;
                    ext[7] = jd0 * u0 + jd1 * u1;
                }
                *reinterpret_cast<float4*>(rec) = make_float4(ext[0], ext[1], ext[2], ext[3]);
                *reinterpret_cast<float4*>(rec + 4) = make_float4(ext[4], ext[5], ext[6], ext[7]);
            }
            __syncwarp();

            // ---------------------------------------------------------------- R rows (and A_lo) for this warp's 8 pixels
#pragma unroll 2
            for (int i4 = 0; i4 < 4; ++i4) {
                const int pl = i4 * 2 + hw, nl = g * 8 + pl;
                const float4 e0 = *reinterpret_cast<const float4*>(myRec + pl * TC_REC);
                const float4 e1 = *reinterpret_cast<const float4*>(myRec + pl * TC_REC + 4);
                const float sn = e1.w;
                const uint32_t offA = blkA * 8192 + sw128_32b_off(nl, ccA);
                const float4 b0 = *reinterpret_cast<const float4*>(As + offA);
                const float4 b1 = *reinterpret_cast<const float4*>(As + offA + 2 * 8192);
                *reinterpret_cast<float4*>(Rs + offA) = make_float4(tf32_rna(sn * b0.x), tf32_rna(sn * b0.y), tf32_rna(sn * b0.z), tf32_rna(sn * b0.w));
                *reinterpret_cast<float4*>(Rs + offA + 2 * 8192) = make_float4(tf32_rna(sn * b1.x), tf32_rna(sn * b1.y), tf32_rna(sn * b1.z), tf32_rna(sn * b1.w));
                if constexpr (MODE == 2) {
                    unsigned char* Al = base + SM::off_Alo + s * TC_STAGE_A;
                    *reinterpret_cast<float4*>(Al + offA) = make_float4(b0.x - tf32_trunc(b0.x), b0.y - tf32_trunc(b0.y), b0.z - tf32_trunc(b0.z), b0.w - tf32_trunc(b0.w));
                    *reinterpret_cast<float4*>(Al + offA + 2 * 8192) = make_float4(b1.x - tf32_trunc(b1.x), b1.y - tf32_trunc(b1.y), b1.z - tf32_trunc(b1.z), b1.w - tf32_trunc(b1.w));
                }
                if (hl == 0) *reinterpret_cast<float4*>(Rs + 4 * 8192 + sw128_32b_off(nl, 0)) = make_float4(tf32_rna(e0.x), tf32_rna(e0.y), tf32_rna(e0.z), tf32_rna(e0.w));
                if (hl == 1) *reinterpret_cast<float4*>(Rs + 4 * 8192 + sw128_32b_off(nl, 1)) = make_float4(tf32_rna(e1.x), tf32_rna(e1.y), tf32_rna(e1.z), 0.f);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
        }
        if (cur_b >= 0) flush(span);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// ---- host side ------------------------------------------------------------------------------------
bool tc_supported(const banet_level_t* lv)
{
    return lv->K == TC_K && (lv->C == 64 || lv->C == 128) && (lv->conv2_channels == lv->C || lv->conv2_channels == 3 * lv->C) &&
           ((reinterpret_cast<uintptr_t>(lv->conv1) | reinterpret_cast<uintptr_t>(lv->conv2) | reinterpret_cast<uintptr_t>(lv->B)) % 16 == 0) &&
           (long long)lv->nb * lv->N < (1LL << 31);
}

int build_plan_tc(const banet_level_t* lv, int num_sms, BuildPlan* plan)
{
    plan->KP = TC_K;
    plan->tiles_per_pair = (lv->N + TC_TILE - 1) / TC_TILE;
    plan->total_tiles = (long long)lv->nb * plan->tiles_per_pair;
    long long grid = num_sms;
    if (grid > plan->total_tiles) grid = plan->total_tiles;
    if (grid < 1) grid = 1;
    plan->grid = (int)grid;
    const long long tiles_per_cta = (plan->total_tiles + grid - 1) / grid;
    plan->max_span = (int)((tiles_per_cta + plan->tiles_per_pair - 2) / plan->tiles_per_pair) + 1;
    SlotLayout L{lv->K, lv->C};
    plan->slot_floats = L.floats();
    plan->ws_bytes = align_up((size_t)plan->grid * plan->max_span * plan->slot_floats * sizeof(float), 256);
    return BANET_OK;
}

template <int NCH, bool FLY, int MODE>
static int launch_tc(const CUtensorMap& tm, const BuildParams& prm, int grid, cudaStream_t st)
{
    auto kern = lm_build_tc_kernel<NCH, FLY, MODE>;
    const int smem = TcSmem<MODE>::bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("lm_build_tc: smem attr (%d B): %s", smem, cudaGetErrorString(e)); return BANET_ERR_CUDA; }
    kern<<<grid, TC_THREADS, smem, st>>>(tm, prm);
    BANET_CUDA_LAUNCH_CHECK("lm_build_tc_kernel launch");
    return BANET_OK;
}

int lm_build_tc(const banet_level_t* lv, const BuildPlan& plan, int mode, const float* R, const float* T, const float* W,
                float* H, float* g, float* rbar_sum, float* nvalid, void* ws, cudaStream_t st)
{
    BANET_REQUIRE(tc_supported(lv), BANET_ERR_UNSUPPORTED,
                  "lm_build (tensor-core path) needs K=128, C in {64,128}, 16-B aligned tensors; got K=%d C=%d", lv->K, lv->C);
    CUtensorMap tm;
    int rc = make_tmap_f32_2d_sw128_32b(&tm, lv->B, (uint64_t)lv->nb * lv->N, TC_K, TC_TILE, 32);
    if (rc) return rc;
    BuildParams prm;
    prm.nb = lv->nb; prm.N = lv->N; prm.C = lv->C; prm.K = lv->K; prm.h = lv->h; prm.w = lv->w; prm.c2 = lv->conv2_channels;
    prm.conv1 = lv->conv1; prm.conv2 = lv->conv2; prm.intr = lv->intr; prm.p = lv->p; prm.D = lv->D; prm.B = lv->B;
    prm.R = R; prm.T = T; prm.W = W;
    prm.partials = reinterpret_cast<float*>(ws);
    prm.slot_floats = plan.slot_floats; prm.max_span = plan.max_span;
    prm.tiles_per_pair = plan.tiles_per_pair; prm.total_tiles = plan.total_tiles;
    const bool fly = lv->conv2_channels == lv->C;
    const int nch = lv->C / 64;
#define BANET_TC(NCHV, FLYV)                                                                        \
    rc = (mode == 2) ? launch_tc<NCHV, FLYV, 2>(tm, prm, plan.grid, st) : launch_tc<NCHV, FLYV, 1>(tm, prm, plan.grid, st)
    if (nch == 2) { if (fly) BANET_TC(2, true); else BANET_TC(2, false); }
    else          { if (fly) BANET_TC(1, true); else BANET_TC(1, false); }
#undef BANET_TC
    if (rc) return rc;
    return launch_lm_reduce(prm, plan.grid, H, g, rbar_sum, nvalid, st);
}

}  // namespace banet
