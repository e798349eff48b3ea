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
// Precision modes (tf32 keeps 10 mantissa bits; products are exact, accumulation is fp32):
//   MODE 1  one pass:   A truncated by the tensor core, R rounded to nearest
//   MODE 2  two passes: + A_lo = b - trunc(b)                       (only R's unbiased rounding remains)
//   MODE 3  three passes: + R_lo = s*b - rna(s*b)                   (fp32-grade: the dropped term is ~2^-22)
//
// Tiles: 64 points.  With the dense-grid hint (banet_level_t::grid_w/h) a tile is an 8x8 pixel patch fetched by ONE
// 3-D TMA box and the gather warps' conv2 taps of a tile overlap (~1.3 texel fetches per pixel instead of ~2);
// without the hint a tile is 64 consecutive points (2-D TMA box).
//
// TMEM accumulators: the tensor core TRUNCATES when it adds into the fp32 accumulator (measured: about -2.5e-8 relative
// per accumulation step, tests/test_gpu_tensorcore.py::test_tcgen05_accumulator_rounding), so accumulation chains are kept
// short: the hi-pass accumulates at most TC_CHAIN tiles into one of two ping-pong TMEM regions, which the gather warps
// then drain (tcgen05.ld) and add, round-to-nearest, into the partial slot while the other region fills; the lo passes
// (2^-11 of the magnitude) use a third region that is drained once per pair.
//
// Warp roles (384 threads = 3 warpgroups, 1 CTA / SM, persistent over a contiguous tile range):
//   warp 0     TMA producer: basis tile -> smem stage (2 stages, mbarrier complete_tx)          } warpgroup 0 gives its
//   warp 1     MMA issuer (one thread): 8 tcgen05.mma per pass and tile; tcgen05.commit         } registers away
//   warp 2     L2 prefetcher, two tiles ahead: conv1 / p / D rows of the tile and (dense grid) the conv2 footprint the   } (setmaxnreg.dec 32)
//              tile will sample, predicted from the flow of the tile just finished (cp.async.bulk.prefetch[.tensor])
//   warp 3     idle
//   warps 4-11 gather warps (setmaxnreg.inc 232), 8 pixels each per tile: D~ = D + b.W from the staged tile, thread-per-pixel warp geometry,
//             then half-warp per pixel / lanes over channels for the 4-tap (or 12-texel, F2-only) gather with the loads of
//             the next two (pixel pair, channel chunk) units in flight while the current one is reduced; per-pixel 2x7
//             algebra; R rows; TMEM drains.
#include "common.cuh"
#include "lm_build.h"
#include "tc_utils.cuh"
#include "tmap.h"
#include <stdlib.h>
#include <limits.h>

namespace banet {
using namespace tc;

constexpr int TC_TILE = 64;
constexpr int TC_GW0 = 4;                          // first gather warp (warpgroup 1); warps 0-3 = producer, MMA issuer, 2 idle
constexpr int TC_GWMAX = 16;                       // gather warps: 8 (8 pixels each, 3 load units in flight) or 16 (4 pixels, 1 unit)
constexpr int TC_CHAIN = 8;                        // tiles per hi-accumulator chain (64 accumulation steps: ~ -1.6e-6 relative bias)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_ACCL = 320;                       // TMEM column of the lo accumulator (hi: 0 and 160)
constexpr int TC_K = 128;
constexpr int TC_N = 160;
constexpr int TC_STAGE_A = 4 * TC_TILE * 128;      // 32 KB: 4 blocks of [64 rows][128 B]
constexpr int TC_STAGE_R = 5 * TC_TILE * 128;      // 40 KB
constexpr int TC_REC = 12;                         // floats per pixel record

template <int MODE> struct TcSmem {
    static constexpr int off_A = 0;                                   // 3 stages (TMA landing zone, also MMA operand A_hi)
    static constexpr int off_R = 3 * TC_STAGE_A;                      // 1 stage
    static constexpr int off_Alo = off_R + TC_STAGE_R;                // 1 stage (MODE >= 2)
    static constexpr int off_Rlo = off_Alo + (MODE >= 2 ? TC_STAGE_A : 0);   // 1 stage (MODE 3)
    static constexpr int off_misc = off_Rlo + (MODE == 3 ? TC_STAGE_R : 0);
    static constexpr int off_bar = off_misc;                          // 14 mbarriers
    static constexpr int off_tmem = off_bar + 112;
    static constexpr int off_pose = off_misc + 128;                   // [TC_GWMAX warps][2 pair parities][16] floats
    static constexpr int off_rec = off_pose + TC_GWMAX * 32 * 4;      // [2 tile parities][64 pixels][TC_REC] floats
    static constexpr int off_cc = off_rec + 2 * TC_TILE * TC_REC * 4; // [64 pixel slots][28] floats
    static constexpr int off_bb = off_cc + TC_TILE * 28 * 4;          // [2][TC_GWMAX][4] ints: tap-origin bounding boxes
    static constexpr int total = off_bb + 2 * TC_GWMAX * 4 * 4;
    static constexpr int bytes = total + 1024;                        // slack for manual 1024-B alignment
};

template <int NT> __device__ __forceinline__ void gather_bar() { asm volatile("bar.sync 1, %0;" :: "n"(NT) : "memory"); }
__device__ __forceinline__ int reflect_i(int i, int n) { i = i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); return i < 0 ? 0 : i; }
__device__ __forceinline__ void pf_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_TRACE(slot) do { if (prm.trace && blockIdx.x == 1 && lane == 0 && (g == 0 || g == GW - 1) && it >= 16 && it < 48) \
        prm.trace[(((g == 0 ? 0 : 1) * 32 + (it - 16)) * 8) + (slot)] = gtime(); } while (0)
__device__ __forceinline__ float hsum16(float v) {            // sum over the 16 lanes of a half-warp
    v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

struct TileCoord { int b, n0, cnt, tx0, ty0; };

__device__ __forceinline__ TileCoord tile_coord(const BuildParams& prm, long long tl) {
    TileCoord tc;
    const unsigned t = (unsigned)tl, tpp = (unsigned)prm.tiles_per_pair;      // total_tiles < 2^31 (checked on the host): 32-bit maths
    tc.b = (int)(t / tpp);
    const int r = (int)(t - (unsigned)tc.b * tpp);
    if (prm.grid_w > 0) { const int tyi = r / prm.tiles_x; tc.ty0 = tyi * 8; tc.tx0 = (r - tyi * prm.tiles_x) * 8; tc.n0 = 0; tc.cnt = TC_TILE; }
    else { tc.n0 = r * TC_TILE; tc.cnt = min(TC_TILE, prm.N - tc.n0); tc.tx0 = tc.ty0 = 0; }
    return tc;
}

template <int NCH, bool FLY, int MODE, int GW>
__global__ void __launch_bounds__((TC_GW0 + GW) * 32, 1)
lm_build_tc_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapC2, const BuildParams prm)
{
    using SM = TcSmem<MODE>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays in the shared state space (LDS/STS)
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + SM::off_bar);
    uint64_t* fullB = bars;          // [3]  TMA landed
    uint64_t* emptyB = bars + 3;     // [3]  MMAs that read A stage s have completed
    uint64_t* ready = bars + 6;      //      R (Alo, Rlo) of the current tile written by all gather warps
    uint64_t* rfree = bars + 7;      //      MMAs of the current tile completed -> R may be overwritten
    uint64_t* flushb = bars + 8;     //      every MMA of the span completed
    uint64_t* tmemfree = bars + 9;   //      lo accumulator drained by the gather warps
    uint64_t* chain_done = bars + 10; // [2] every hi-pass MMA of the chain that used accumulator `set` completed
    uint64_t* drained = bars + 12;   // [2]  hi accumulator `set` drained by the gather warps
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(base + SM::off_tmem);
    float* sPose = reinterpret_cast<float*>(base + SM::off_pose);
    float* sRec = reinterpret_cast<float*>(base + SM::off_rec);
    float* sCC = reinterpret_cast<float*>(base + SM::off_cc);
    int* sBB = reinterpret_cast<int*>(base + SM::off_bb);

    constexpr int TC_THREADS = (TC_GW0 + GW) * 32;
    constexpr int PXW = TC_TILE / GW;                 // pixels per gather warp and tile (8 or 4)
    constexpr int DEPTH = (GW == 8) ? 3 : 1;          // (pixel pair, 64-channel chunk) load units in flight per warp
    constexpr int GREG = (GW == 8) ? 232 : 112;       // registers per gather thread after setmaxnreg
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = prm.N, h = prm.h, w = prm.w, c2 = prm.c2;
    const bool grid2d = prm.grid_w > 0;
    constexpr int C = 64 * NCH;
    const long long t_begin = part_begin(prm.total_tiles, gridDim.x, blockIdx.x);
    const long long t_end   = part_begin(prm.total_tiles, gridDim.x, blockIdx.x + 1);

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        mbar_init(ready, GW); mbar_init(rfree, 1);
        mbar_init(flushb, 1); mbar_init(tmemfree, GW);
        mbar_init(&chain_done[0], 1); mbar_init(&chain_done[1], 1);
        mbar_init(&drained[0], GW); mbar_init(&drained[1], GW);
        fence_barrier_init();
        prefetch_tmap(&tmapB);
    }
    if (warp == 0) tmem_alloc<TC_TMEM_COLS>(s_tmem);
    // the pad chunks of the 5th block of R / R_lo (columns 136..159) stay zero for the whole kernel
    for (int i = tid; i < TC_TILE * 8; i += TC_THREADS) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<float4*>(base + SM::off_R + 4 * 8192 + sw128_32b_off(r, c)) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 3) *reinterpret_cast<float4*>(base + SM::off_Rlo + 4 * 8192 + sw128_32b_off(r, c)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp < TC_GW0) {
      setmaxnreg_dec<32>();
      if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int it = 0;
            for (long long t = t_begin; t < t_end; ++t, ++it) {
                const int s = it % 3, ph = (it / 3) & 1;
                const TileCoord tc = tile_coord(prm, t);
                mbar_wait_sleep(&emptyB[s], ph ^ 1);
                mbar_arrive_expect_tx(&fullB[s], TC_STAGE_A);
                unsigned char* dst = base + SM::off_A + s * TC_STAGE_A;
                if (grid2d) {
#pragma unroll
                    for (int blk = 0; blk < 4; ++blk) tma_load_3d(dst + blk * 8192, &tmapB, blk * 32, tc.tx0, tc.b * prm.grid_h + tc.ty0, &fullB[s]);
                } else {
                    const int row = tc.b * N + tc.n0;
#pragma unroll
                    for (int blk = 0; blk < 4; ++blk) tma_load_2d(dst + blk * 8192, &tmapB, blk * 32, row, &fullB[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32_mn_mn(128, TC_N);
            int it = 0, span = 0, cur_b = -1, chain = -1, tic = 0, set = 0;
            uint32_t accH = 0, accL = 0;
            const uint32_t rhi = smem_u32(base + SM::off_R), rlo = smem_u32(base + SM::off_Rlo), alo = smem_u32(base + SM::off_Alo);
            for (long long t = t_begin; t < t_end; ++t, ++it) {
                const int s = it % 3;
                const int b = (int)((unsigned)t / (unsigned)prm.tiles_per_pair);
                if (b != cur_b) {
                    if (cur_b >= 0) { if (tic > 0) mma_commit(&chain_done[set]); mma_commit(flushb); ++span; }
                    mbar_wait_sleep(tmemfree, (span & 1) ^ 1);    // lo accumulator drained by the previous span's flush
                    accL = 0; cur_b = b; tic = 0;
                }
                if (tic == 0) {                                   // new hi chain on the other accumulator
                    ++chain; set = chain & 1;
                    mbar_wait_sleep(&drained[set], ((chain >> 1) & 1) ^ 1);
                    accH = 0;
                }
                mbar_wait_sleep(ready, it & 1);
                tc_fence_after_sync();
                const uint32_t ahi = smem_u32(base + SM::off_A + s * TC_STAGE_A);
#pragma unroll
                for (int pass = 0; pass < MODE; ++pass) {
                    const uint32_t a0 = (pass == 1) ? alo : ahi;
                    const uint32_t r0 = (pass == 2) ? rlo : rhi;
                    const uint32_t dcol = tmem + (pass == 0 ? set * TC_N : TC_ACCL);
#pragma unroll
                    for (int kk = 0; kk < TC_TILE / 8; ++kk) {
                        mma_tf32_ss(dcol, make_desc_mn_sw128_32b(a0 + kk * 1024, 8192, 512),
                                    make_desc_mn_sw128_32b(r0 + kk * 1024, 8192, 512), idesc, pass == 0 ? accH : accL);
                        if (pass == 0) accH = 1; else accL = 1;
                    }
                }
                mma_commit(&emptyB[s]);
                mma_commit(rfree);
                if (++tic == TC_CHAIN) { mma_commit(&chain_done[set]); tic = 0; }
            }
            if (cur_b >= 0) { if (tic > 0) mma_commit(&chain_done[set]); mma_commit(flushb); }
        }
      } else if (warp == 2) {
        // ===================================================================== L2 prefetcher (whole warp), two tiles ahead
        constexpr int PF_AHEAD = 2;
        constexpr int PF_M = FLY ? 2 : 1;                 // margin left/top of the predicted tap origin
        int it = 0;
        for (long long t = t_begin; t < t_end; ++t, ++it) {
            const long long tp = t + PF_AHEAD;
            if (tp >= t_end) break;
            const TileCoord tc = tile_coord(prm, t), tn = tile_coord(prm, tp);
            // streaming inputs of tile tp (addresses known exactly)
            if (grid2d) {
                if (lane < 8) {
                    const int gy = tn.ty0 + lane;
                    if (gy < prm.grid_h && tn.tx0 < prm.grid_w) {
                        const size_t n = (size_t)gy * prm.grid_w + tn.tx0;
                        const int wpx = min(8, prm.grid_w - tn.tx0);
                        prefetch_l2_bulk(prm.conv1 + ((size_t)tn.b * N + n) * C, (uint32_t)(wpx * C * 4));
                        if ((n & 3) == 0 && (N & 3) == 0) {
                            const uint32_t by = (uint32_t)(((wpx * 4) + 15) & ~15);
                            prefetch_l2_bulk(prm.D + (size_t)tn.b * N + n, by);
#pragma unroll
                            for (int k = 0; k < 3; ++k) prefetch_l2_bulk(prm.p + ((size_t)tn.b * 3 + k) * N + n, by);
                        }
                    }
                }
            } else if (lane == 0) {
                prefetch_l2_bulk(prm.conv1 + ((size_t)tn.b * N + tn.n0) * C, (uint32_t)(tn.cnt * C * 4));
                if ((N & 3) == 0) {
                    const uint32_t by = (uint32_t)(((tn.cnt * 4) + 15) & ~15);
                    prefetch_l2_bulk(prm.D + (size_t)tn.b * N + tn.n0, by);
#pragma unroll
                    for (int k = 0; k < 3; ++k) prefetch_l2_bulk(prm.p + ((size_t)tn.b * 3 + k) * N + tn.n0, by);
                }
            }
            // conv2 footprint of tile tp, predicted from the flow of tile t (dense grid only).  Off by default: measured on
            // B200 it raised DRAM traffic by ~35% (boxes over-cover, lines evicted before use) without shortening the kernel,
            // whose gather phase is bound by the per-warp dependency chain, not by tap latency (profiles/r01_*).
            if (grid2d && prm.pf_conv2) {
                const bool have_bb = mbar_wait_bounded(ready, it & 1, 2000);   // tile t's geometry (and bounding boxes) complete
                int xmn = INT_MAX, ymn = INT_MAX;
                if (have_bb && lane < GW) { xmn = sBB[((it & 1) * TC_GWMAX + lane) * 4 + 0]; ymn = sBB[((it & 1) * TC_GWMAX + lane) * 4 + 2]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { xmn = min(xmn, __shfl_xor_sync(0xffffffffu, xmn, o)); ymn = min(ymn, __shfl_xor_sync(0xffffffffu, ymn, o)); }
                const int fx = (xmn == INT_MAX || tn.b != tc.b) ? 0 : xmn - tc.tx0;
                const int fy = (ymn == INT_MAX || tn.b != tc.b) ? 0 : ymn - tc.ty0;
                if (lane < c2 / C) prefetch_l2_tensor_4d(&tmapC2, lane * C, tn.tx0 + fx - PF_M, tn.ty0 + fy - PF_M, tn.b);
            }
        }
      }     // warp 3 of warpgroup 0 idle
    } else {
        // ===================================================================== gather warps
        // Software-pipelined over tiles so that the ALU-only phases hide the tap-load latency of the gather:
        //   iteration j:  issue loads of unit 0 of tile j
        //                 2x7 algebra + R rows of tile j-1 (-> MMA j-1)          [hides unit 0's latency]
        //                 reduce unit 0, issue unit 1
        //                 b.W + warp geometry of tile j+1 (its basis tile has landed: 3 TMA stages)   [hides unit 1's]
        //                 remaining units of tile j
        setmaxnreg_inc<GREG>();
        const int g = warp - TC_GW0, hw = lane >> 4, hl = lane & 15;
        const int gtid = tid - TC_GW0 * 32;
        const int blkA = hl >> 3, ccA = hl & 7;                 // this lane's two 16-B chunks of a 128-float row: blocks blkA and 2+blkA
        float wreg[8];
        float rb[NCH * 4];
#pragma unroll
        for (int u = 0; u < NCH * 4; ++u) rb[u] = 0.f;
        float* myCC = sCC + (g * PXW + (lane & (PXW - 1))) * 28;
        float* myPose = sPose + g * 32;                          // [2 pair parities][16]
        const SlotLayout L{TC_K, C};
        unsigned char* Rs = base + SM::off_R;
        const int ntiles = (int)(t_end - t_begin);
        int chain = -1, tic = 0, next_drain = 0;                 // TMEM chain bookkeeping (mirrors the MMA issuer), advanced per scaled tile
        bool first_drain = true;
        int gpar = 1, geom_b = -1;                               // pair parity / pair of the geometry stage
        int spar = 1, scale_b = -1, sspan = -1;                  // ... of the algebra+scale stage; sspan = slot index within this CTA

        // add (or store) this warp's share of a 128 x 160 TMEM region into the slot: H_dd transposed (coalesced), ext rows
        auto drain_region = [&](float* slot, uint32_t col0, bool overwrite) {
            const int q = warp & 3, row = q * 32 + lane;
            const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + col0;
            float v[32];
            constexpr int NSUB = GW / 4;                 // warps per TMEM lane quadrant; they split the 4 H_dd column blocks
            const int sub = g >> 2;
#pragma unroll 1
            for (int cbi = 0; cbi < 4 / NSUB; ++cbi) {
                const int cb = sub * (4 / NSUB) + cbi;
                tmem_ld_32x32(tq + cb * 32, v);
                float* dst = slot + (size_t)(cb * 32) * TC_K + row;
                if (overwrite) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) dst[(size_t)j * TC_K] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) dst[(size_t)j * TC_K] += v[j];
                }
            }
            if (sub == 0) {
                tmem_ld_32x32(tq + 128, v);
                float* dst = slot + L.off_ext() + row;
#pragma unroll
                for (int r = 0; r < 7; ++r) { if (overwrite) dst[r * TC_K] = v[r]; else dst[r * TC_K] += v[r]; }
            }
        };
        auto drain_hi = [&](int c, float* slot) {
            const int set = c & 1;
            mbar_wait(&chain_done[set], (c >> 1) & 1);
            tc_fence_after_sync();
            drain_region(slot, set * TC_N, first_drain);
            first_drain = false;
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&drained[set]);
        };
        auto flush = [&](int sp) {
            float* slot = prm.partials + ((size_t)blockIdx.x * prm.max_span + sp) * prm.slot_floats;
            for (; next_drain <= chain; ++next_drain) drain_hi(next_drain, slot);
            mbar_wait(flushb, sp & 1);
            tc_fence_after_sync();
            if constexpr (MODE >= 2) drain_region(slot, TC_ACCL, false);
            tc_fence_before_sync();
            // rbar / cc through a scratch aliased on R (every MMA of this span has completed)
            float* scratch = reinterpret_cast<float*>(base + SM::off_R);
#pragma unroll
            for (int u = 0; u < NCH * 4; ++u) rb[u] += __shfl_xor_sync(0xffffffffu, rb[u], 16);
            if (hw == 0) {
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    *reinterpret_cast<float4*>(scratch + g * C + 64 * j + 4 * hl) = make_float4(rb[4 * j], rb[4 * j + 1], rb[4 * j + 2], rb[4 * j + 3]);
            }
            gather_bar<GW * 32>();
            if (gtid < C) {
                float s = 0.f;
#pragma unroll
                for (int wq = 0; wq < GW; ++wq) s += scratch[wq * C + gtid];
                slot[L.off_rbar() + gtid] = s;
            }
            if (gtid < 28) {
                float s = 0.f;
                for (int e = 0; e < TC_TILE; ++e) s += sCC[e * 28 + gtid];
                slot[L.off_cc() + gtid] = s;
            }
            gather_bar<GW * 32>();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmemfree);
#pragma unroll
            for (int u = 0; u < NCH * 4; ++u) rb[u] = 0.f;
        };

        // ---- stage G: D~ - D = b.W (geom_a) and the warp geometry of tile j (geom_b2, thread per pixel), records -> sRec[j&1]
        float mydot = 0.f;
        auto geom_a = [&](int j) {
            const TileCoord tc = tile_coord(prm, t_begin + j);
            const int b = tc.b;
            if (b != geom_b) {                                   // new pair: this warp's private copy of pose/intrinsics, W chunks
                gpar ^= 1; geom_b = b;
                float* pose = myPose + gpar * 16;
                __syncwarp();
                if (lane < 9) pose[lane] = prm.R[b * 9 + lane];
                else if (lane < 12) pose[lane] = prm.T[b * 3 + lane - 9];
                else if (lane < 16) pose[lane] = prm.intr[b * 4 + lane - 12];
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(prm.W + (size_t)b * TC_K + blkA * 32 + ccA * 4));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(prm.W + (size_t)b * TC_K + (2 + blkA) * 32 + ccA * 4));
                wreg[0] = w0.x; wreg[1] = w0.y; wreg[2] = w0.z; wreg[3] = w0.w;
                wreg[4] = w1.x; wreg[5] = w1.y; wreg[6] = w1.z; wreg[7] = w1.w;
                __syncwarp();
            }
            const float* pose = myPose + gpar * 16;
            const int s = j % 3, ph = (j / 3) & 1;
            const unsigned char* As = base + SM::off_A + s * TC_STAGE_A;
            float* rec = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            mbar_wait(&fullB[s], ph);
            mydot = 0.f;
#pragma unroll
            for (int i4 = 0; i4 < PXW / 2; ++i4) {
                const int nl = g * PXW + i4 * 2 + hw;
                const uint32_t offA = blkA * 8192 + sw128_32b_off(nl, ccA);
                const float4 b0 = *reinterpret_cast<const float4*>(As + offA);
                const float4 b1 = *reinterpret_cast<const float4*>(As + offA + 2 * 8192);
                float dot = b0.x * wreg[0] + b0.y * wreg[1] + b0.z * wreg[2] + b0.w * wreg[3]
                          + b1.x * wreg[4] + b1.y * wreg[5] + b1.z * wreg[6] + b1.w * wreg[7];
                dot = hsum16(dot);
                const float other = __shfl_sync(0xffffffffu, dot, 16);
                if (lane == 2 * i4) mydot = dot;
                if (lane == 2 * i4 + 1) mydot = other;
            }
        };
        auto geom_b2 = [&](int j) {
            const TileCoord tc = tile_coord(prm, t_begin + j);
            const int b = tc.b;
            const float* pose = myPose + gpar * 16;
            float* rec = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            if (lane < PXW) {                                    // bundlenet.py:208-224, mask :231
                int n; bool valid;
                const int nl = g * PXW + lane;
                if (grid2d) { const int gx = tc.tx0 + (nl & 7), gy = tc.ty0 + (nl >> 3); valid = gx < prm.grid_w && gy < prm.grid_h; n = gy * prm.grid_w + gx; }
                else { valid = nl < tc.cnt; n = tc.n0 + nl; }
                float mask = 0.f, x = 0.f, y = 0.f, iZ = 0.f, rx = 0.f, ry = 0.f, rz = 0.f, dx = 0.f, dy = 0.f;
                int x0 = 0, y0 = 0;
                if (valid) {
                    const float* pp = prm.p + (size_t)b * 3 * N + n;
                    const float p0 = __ldg(pp), p1 = __ldg(pp + N), p2 = __ldg(pp + 2 * (size_t)N);
                    const float Dt = __ldg(prm.D + (size_t)b * N + n) + mydot;
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
                float4* r4 = reinterpret_cast<float4*>(rec + lane * TC_REC);
                r4[0] = make_float4(__int_as_float(x0), __int_as_float(y0), dx, dy);
                r4[1] = make_float4(mask, x, y, iZ);
                r4[2] = make_float4(rx, ry, rz, __int_as_float(valid ? n : 0));
            }
            __syncwarp();
        };

        // ---- stage S: per-pixel 2x7 algebra (bundlenet.py:49-74), R rows (A_lo, R_lo), TMEM chain bookkeeping of tile j
        auto s3 = [&](int j) {
            const TileCoord tc = tile_coord(prm, t_begin + j);
            if (tc.b != scale_b) {                               // first tile of a pair in this CTA
                spar ^= 1; scale_b = tc.b; ++sspan; tic = 0; first_drain = true;
                if (lane < PXW) {
#pragma unroll
                    for (int q = 0; q < 28; ++q) myCC[q] = 0.f;
                }
            }
            const float* pose = myPose + spar * 16;
            float* recw = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            const unsigned char* As = base + SM::off_A + (j % 3) * TC_STAGE_A;
            if (tic == 0) ++chain;
            if (tic == TC_CHAIN / 2 && next_drain < chain) {     // lazy drain of the previous hi chain: its MMAs completed long ago
                drain_hi(next_drain, prm.partials + ((size_t)blockIdx.x * prm.max_span + sspan) * prm.slot_floats);
                ++next_drain;
            }
            if (++tic == TC_CHAIN) tic = 0;
            if (lane < PXW) {
                float* rec = recw + lane * TC_REC;
                const float4 ra = *reinterpret_cast<const float4*>(rec), rbq = *reinterpret_cast<const float4*>(rec + 4),
                             rc = *reinterpret_cast<const float4*>(rec + 8);
                float ext[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (rbq.x != 0.f) {
                    const float m11 = ra.x, m12 = ra.y, m22 = ra.z, q1 = ra.w, q2 = rc.w, x = rbq.y, y = rbq.z, iZ = rbq.w;
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
        };
        auto scale = [&](int j) {
            const TileCoord tc = tile_coord(prm, t_begin + j);
            float* recw = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            const unsigned char* As = base + SM::off_A + (j % 3) * TC_STAGE_A;
            if (j > 0) mbar_wait(rfree, (j - 1) & 1);            // the previous tile's MMAs no longer read R / A_lo / R_lo
#pragma unroll 2
            for (int i4 = 0; i4 < PXW / 2; ++i4) {
                const int pl = i4 * 2 + hw, nl = g * PXW + pl;
                const float4 e0 = *reinterpret_cast<const float4*>(recw + pl * TC_REC);
                const float4 e1 = *reinterpret_cast<const float4*>(recw + pl * TC_REC + 4);
                const float sn = e1.w;
                const uint32_t offA = blkA * 8192 + sw128_32b_off(nl, ccA);
                const float4 b0 = *reinterpret_cast<const float4*>(As + offA);
                const float4 b1 = *reinterpret_cast<const float4*>(As + offA + 2 * 8192);
                const float4 p0 = make_float4(sn * b0.x, sn * b0.y, sn * b0.z, sn * b0.w);
                const float4 p1 = make_float4(sn * b1.x, sn * b1.y, sn * b1.z, sn * b1.w);
                const float4 h0 = make_float4(tf32_rna(p0.x), tf32_rna(p0.y), tf32_rna(p0.z), tf32_rna(p0.w));
                const float4 h1 = make_float4(tf32_rna(p1.x), tf32_rna(p1.y), tf32_rna(p1.z), tf32_rna(p1.w));
                *reinterpret_cast<float4*>(Rs + offA) = h0;
                *reinterpret_cast<float4*>(Rs + offA + 2 * 8192) = h1;
                if constexpr (MODE >= 2) {
                    unsigned char* Al = base + SM::off_Alo;
                    *reinterpret_cast<float4*>(Al + offA) = make_float4(b0.x - tf32_trunc(b0.x), b0.y - tf32_trunc(b0.y), b0.z - tf32_trunc(b0.z), b0.w - tf32_trunc(b0.w));
                    *reinterpret_cast<float4*>(Al + offA + 2 * 8192) = make_float4(b1.x - tf32_trunc(b1.x), b1.y - tf32_trunc(b1.y), b1.z - tf32_trunc(b1.z), b1.w - tf32_trunc(b1.w));
                }
                if constexpr (MODE == 3) {
                    unsigned char* Rl = base + SM::off_Rlo;
                    *reinterpret_cast<float4*>(Rl + offA) = make_float4(p0.x - h0.x, p0.y - h0.y, p0.z - h0.z, p0.w - h0.w);
                    *reinterpret_cast<float4*>(Rl + offA + 2 * 8192) = make_float4(p1.x - h1.x, p1.y - h1.y, p1.z - h1.z, p1.w - h1.w);
                }
                if (hl < 2) {
                    const float4 ev = hl == 0 ? e0 : make_float4(e1.x, e1.y, e1.z, 0.f);
                    const float4 eh = make_float4(tf32_rna(ev.x), tf32_rna(ev.y), tf32_rna(ev.z), tf32_rna(ev.w));
                    const uint32_t offE = 4 * 8192 + sw128_32b_off(nl, hl);
                    *reinterpret_cast<float4*>(Rs + offE) = eh;
                    if constexpr (MODE == 3)
                        *reinterpret_cast<float4*>(base + SM::off_Rlo + offE) = make_float4(ev.x - eh.x, ev.y - eh.y, ev.z - eh.z, ev.w - eh.w);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready);
            // last tile of this pair inside this CTA?  then drain / publish its slot now
            const bool last_of_pair = (j + 1 >= ntiles) || ((int)((unsigned)(t_begin + j + 1) / (unsigned)prm.tiles_per_pair) != tc.b);
            if (last_of_pair) flush(sspan);
        };

        // ---- stage L: the tap loads / blends of tile j, (pixel pair, 64-channel chunk) units
        constexpr int NUNIT = (PXW / 2) * NCH;
        float4 tb[13];
        float m11 = 0.f, m12 = 0.f, m22 = 0.f, q1 = 0.f, q2 = 0.f;
        auto issue = [&](int j, int b, int u) {
            const float* rec = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            const int pl = 2 * (u / NCH) + hw, co = 64 * (u % NCH);
            const float4 ra = *reinterpret_cast<const float4*>(rec + pl * TC_REC);
            const float mask = rec[pl * TC_REC + 4];
            if (mask != 0.f) {
                const int x0 = __float_as_int(ra.x), y0 = __float_as_int(ra.y);
                const int n = __float_as_int(rec[pl * TC_REC + 11]);
                const float* img = prm.conv2 + (size_t)b * h * w * c2 + 4 * hl + co;
                tb[0] = ld_stream_f4(prm.conv1 + ((size_t)b * N + n) * C + 4 * hl + co);
                if constexpr (!FLY) {
                    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
                    const float* t00 = img + ((size_t)y0 * w + x0) * c2;
                    const float* t01 = img + ((size_t)y0 * w + x1) * c2;
                    const float* t10 = img + ((size_t)y1 * w + x0) * c2;
                    const float* t11 = img + ((size_t)y1 * w + x1) * c2;
                    tb[1] = ldg4(t00); tb[2] = ldg4(t01); tb[3] = ldg4(t10); tb[4] = ldg4(t11);
                    tb[5] = ldg4(t00 + C); tb[6] = ldg4(t01 + C); tb[7] = ldg4(t10 + C); tb[8] = ldg4(t11 + C);
                    tb[9] = ldg4(t00 + 2 * C); tb[10] = ldg4(t01 + 2 * C); tb[11] = ldg4(t10 + 2 * C); tb[12] = ldg4(t11 + 2 * C);
                } else {
                    // F2-only map: 12 texels = rows y0,Y1 x columns XM,x0,X1,XP  +  rows YM,YP x columns x0,X1 (REFLECT-by-one, bundlenet.py:97)
                    const int X1 = reflect_i(x0 + 1, w), XM = reflect_i(x0 - 1, w), XP = reflect_i(x0 + 2, w);
                    const int Y1 = reflect_i(y0 + 1, h), YM = reflect_i(y0 - 1, h), YP = reflect_i(y0 + 2, h);
                    const float* r0 = img + (size_t)y0 * w * c2;
                    const float* r1 = img + (size_t)Y1 * w * c2;
                    const float* rm = img + (size_t)YM * w * c2;
                    const float* rp = img + (size_t)YP * w * c2;
                    const size_t oM = (size_t)XM * c2, o0 = (size_t)x0 * c2, o1 = (size_t)X1 * c2, oP = (size_t)XP * c2;
                    tb[1] = ldg4(r0 + oM); tb[2] = ldg4(r0 + o0); tb[3] = ldg4(r0 + o1); tb[4] = ldg4(r0 + oP);      // aM0 a00 a10 aP0
                    tb[5] = ldg4(r1 + oM); tb[6] = ldg4(r1 + o0); tb[7] = ldg4(r1 + o1); tb[8] = ldg4(r1 + oP);      // aM1 a01 a11 aP1
                    tb[9] = ldg4(rm + o0); tb[10] = ldg4(rm + o1); tb[11] = ldg4(rp + o0); tb[12] = ldg4(rp + o1);   // a0m a1m a0p a1p
                }
            }
        };
        auto pf_unit = [&](int j, int b, int u) {        // L2 prefetch of exactly the lines issue(j,b,u) will load
            const float* rec = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            const int pl = 2 * (u / NCH) + hw, co = 64 * (u % NCH);
            const float4 ra = *reinterpret_cast<const float4*>(rec + pl * TC_REC);
            const float mask = rec[pl * TC_REC + 4];
            if (mask != 0.f) {
                const int x0 = __float_as_int(ra.x), y0 = __float_as_int(ra.y);
                const int n = __float_as_int(rec[pl * TC_REC + 11]);
                const float* img = prm.conv2 + (size_t)b * h * w * c2 + 4 * hl + co;
                
                if constexpr (!FLY) {
                    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
                    const float* t00 = img + ((size_t)y0 * w + x0) * c2;
                    const float* t01 = img + ((size_t)y0 * w + x1) * c2;
                    const float* t10 = img + ((size_t)y1 * w + x0) * c2;
                    const float* t11 = img + ((size_t)y1 * w + x1) * c2;
                    pf_l2(t00); pf_l2(t01); pf_l2(t10); pf_l2(t11);
                    pf_l2(t00 + C); pf_l2(t01 + C); pf_l2(t10 + C); pf_l2(t11 + C);
                    pf_l2(t00 + 2 * C); pf_l2(t01 + 2 * C); pf_l2(t10 + 2 * C); pf_l2(t11 + 2 * C);
                } else {
                    // F2-only map: 12 texels = rows y0,Y1 x columns XM,x0,X1,XP  +  rows YM,YP x columns x0,X1 (REFLECT-by-one, bundlenet.py:97)
                    const int X1 = reflect_i(x0 + 1, w), XM = reflect_i(x0 - 1, w), XP = reflect_i(x0 + 2, w);
                    const int Y1 = reflect_i(y0 + 1, h), YM = reflect_i(y0 - 1, h), YP = reflect_i(y0 + 2, h);
                    const float* r0 = img + (size_t)y0 * w * c2;
                    const float* r1 = img + (size_t)Y1 * w * c2;
                    const float* rm = img + (size_t)YM * w * c2;
                    const float* rp = img + (size_t)YP * w * c2;
                    const size_t oM = (size_t)XM * c2, o0 = (size_t)x0 * c2, o1 = (size_t)X1 * c2, oP = (size_t)XP * c2;
                    pf_l2(r0 + oM); pf_l2(r0 + o0); pf_l2(r0 + o1); pf_l2(r0 + oP);      // aM0 a00 a10 aP0
                    pf_l2(r1 + oM); pf_l2(r1 + o0); pf_l2(r1 + o1); pf_l2(r1 + oP);      // aM1 a01 a11 aP1
                    pf_l2(rm + o0); pf_l2(rm + o1); pf_l2(rp + o0); pf_l2(rp + o1);   // a0m a1m a0p a1p
                }
            }
        };
        auto compute = [&](int j, int u) {
            float* rec = sRec + ((j & 1) * TC_TILE + g * PXW) * TC_REC;
            const float4* t = tb;
            const int pl = 2 * (u / NCH) + hw, jc = u % NCH;
            const float4 ra = *reinterpret_cast<const float4*>(rec + pl * TC_REC);
            const float mask = rec[pl * TC_REC + 4];
            if (jc == 0) { m11 = m12 = m22 = q1 = q2 = 0.f; }
            if (mask != 0.f) {
                const float dx = ra.z, dy = ra.w;
                const float w00 = (1.f - dx) * (1.f - dy), w01 = dx * (1.f - dy), w10 = (1.f - dx) * dy, w11 = dx * dy;
                if constexpr (!FLY) {
#define BANET_CH(F, CI)                                                                                              \
                    {                                                                                                \
                        const float f2 = w00 * t[1].F + w01 * t[2].F + w10 * t[3].F + w11 * t[4].F;                  \
                        const float gx = w00 * t[5].F + w01 * t[6].F + w10 * t[7].F + w11 * t[8].F;                  \
                        const float gy = w00 * t[9].F + w01 * t[10].F + w10 * t[11].F + w11 * t[12].F;               \
                        const float d = t[0].F - f2;                                                                 \
                        m11 = fmaf(gx, gx, m11); m12 = fmaf(gx, gy, m12); m22 = fmaf(gy, gy, m22);                   \
                        q1 = fmaf(gx, d, q1); q2 = fmaf(gy, d, q2);                                                  \
                        rb[4 * jc + CI] += fabsf(d);                                                                 \
                    }
                    BANET_CH(x, 0) BANET_CH(y, 1) BANET_CH(z, 2) BANET_CH(w, 3)
#undef BANET_CH
                } else {
                    const float h00 = 0.5f * w00, h01 = 0.5f * w01, h10 = 0.5f * w10, h11 = 0.5f * w11;
                    // t: 1 aM0, 2 a00, 3 a10, 4 aP0, 5 aM1, 6 a01, 7 a11, 8 aP1, 9 a0m, 10 a1m, 11 a0p, 12 a1p  (aXY: column X, row Y)
#define BANET_CH(F, CI)                                                                                              \
                    {                                                                                                \
                        const float f2 = w00 * t[2].F + w01 * t[3].F + w10 * t[6].F + w11 * t[7].F;                  \
                        const float gx = h00 * (t[3].F - t[1].F) + h01 * (t[4].F - t[2].F)                           \
                                       + h10 * (t[7].F - t[5].F) + h11 * (t[8].F - t[6].F);                          \
                        const float gy = h00 * (t[6].F - t[9].F) + h10 * (t[11].F - t[2].F)                          \
                                       + h01 * (t[7].F - t[10].F) + h11 * (t[12].F - t[3].F);                        \
                        const float d = t[0].F - f2;                                                                 \
                        m11 = fmaf(gx, gx, m11); m12 = fmaf(gx, gy, m12); m22 = fmaf(gy, gy, m22);                   \
                        q1 = fmaf(gx, d, q1); q2 = fmaf(gy, d, q2);                                                  \
                        rb[4 * jc + CI] += fabsf(d);                                                                 \
                    }
                    BANET_CH(x, 0) BANET_CH(y, 1) BANET_CH(z, 2) BANET_CH(w, 3)
#undef BANET_CH
                }
            }
            if (jc == NCH - 1) {
                // totals overwrite (x0,y0,dx,dy) / n of this pixel's record, which are no longer needed
                // (five independent butterflies pipeline better than a packed 8-shuffle reduction: measured 2.48 vs 2.63 ms)
                m11 = hsum16(m11); m12 = hsum16(m12); m22 = hsum16(m22); q1 = hsum16(q1); q2 = hsum16(q2);
                if (hl == 0) {
                    *reinterpret_cast<float4*>(rec + pl * TC_REC) = make_float4(m11, m12, m22, q1);
                    rec[pl * TC_REC + 11] = q2;
                }
            }
        };

        if (ntiles > 0) { geom_a(0); geom_b2(0); }
        for (int j = 0; j < ntiles; ++j) {
            const int b = (int)((unsigned)(t_begin + j) / (unsigned)prm.tiles_per_pair);
            // the ALU-only phases of the neighbouring tiles sit behind the first two batches of tap loads of this tile
            // (measured: spreading them over all four batches is slower — more live state while loads are in flight)
            issue(j, b, 0);
            if (j > 0) { s3(j - 1); scale(j - 1); }
            compute(j, 0);
            if (NUNIT > 1) issue(j, b, 1);
            if (j + 1 < ntiles) {
                geom_a(j + 1); geom_b2(j + 1);
                if (prm.pf_taps) {       // one tile of lead: by the time tile j+1's taps are loaded they sit in L2
                    const int bn = (int)((unsigned)(t_begin + j + 1) / (unsigned)prm.tiles_per_pair);
#pragma unroll
                    for (int u = 0; u < NUNIT; ++u) pf_unit(j + 1, bn, u);
                }
            }
            if (NUNIT > 1) compute(j, 1);
#pragma unroll
            for (int u = 2; u < NUNIT; ++u) { issue(j, b, u); compute(j, u); }
            __syncwarp();
        }
        if (ntiles > 0) { s3(ntiles - 1); scale(ntiles - 1); }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TC_TMEM_COLS>(tmem);
}

// ---- host side ------------------------------------------------------------------------------------
// tuning knob (not part of the ABI yet): BANET_TC_SMALLK=1 lets K = 64 / 32 take the generation-6 tensor-core kernel too
// (two-pass and fp32-grade modes; TF32X1 falls back to TF32X2).  Off by default: not yet through the GPU parity suite.
static bool tc_small_k() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("BANET_TC_SMALLK"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}
static int tc_generation();

bool tc_supported(const banet_level_t* lv)
{
    const bool k_ok = lv->K == TC_K || (tc_small_k() && tc_generation() == 6 && (lv->K == 64 || lv->K == 32));
    return k_ok && (lv->C == 64 || lv->C == 128) && (lv->conv2_channels == lv->C || lv->conv2_channels == 3 * lv->C) &&
           ((reinterpret_cast<uintptr_t>(lv->conv1) | reinterpret_cast<uintptr_t>(lv->conv2) | reinterpret_cast<uintptr_t>(lv->B)) % 16 == 0) &&
           (long long)lv->nb * lv->N < (1LL << 31) && (long long)lv->nb * ((lv->N + 63) / 64 + 80) < (1LL << 31);
}

int build_plan_tc(const banet_level_t* lv, int num_sms, BuildPlan* plan)
{
    plan->KP = TC_K;
    if (lv->grid_w > 0) plan->tiles_per_pair = ((lv->grid_w + 7) / 8) * ((lv->grid_h + 7) / 8);
    else plan->tiles_per_pair = (lv->N + TC_TILE - 1) / TC_TILE;
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

int lm_build_tc6_launch(int mode, bool fly, int nch, int kblk, const CUtensorMap& tm, const BuildParams& prm, int grid, cudaStream_t st);
// tuning knob (not part of the ABI): BANET_TC_GEN=5 selects the previous kernel generation (default 6: helper warpgroup)
static int tc_generation() {
    static int gen = 0;
    if (!gen) { const char* e = getenv("BANET_TC_GEN"); gen = (e && atoi(e) == 5) ? 5 : 6; }
    return gen;
}
// tuning knob (not part of the ABI): BANET_TC_GW=8 selects the 8-gather-warp variant (default 16)
static int tc_gather_warps() {
    static int gw = 0;
    if (!gw) { const char* e = getenv("BANET_TC_GW"); gw = (e && atoi(e) == 8) ? 8 : 16; }
    return gw;
}

template <int NCH, bool FLY, int MODE, int GW>
static int launch_tc_gw(const CUtensorMap& tm, const CUtensorMap& tm2, const BuildParams& prm, int grid, cudaStream_t st)
{
    auto kern = lm_build_tc_kernel<NCH, FLY, MODE, GW>;
    const int smem = TcSmem<MODE>::bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("lm_build_tc: smem attr (%d B): %s", smem, cudaGetErrorString(e)); return BANET_ERR_CUDA; }
    kern<<<grid, (TC_GW0 + GW) * 32, smem, st>>>(tm, tm2, prm);
    BANET_CUDA_LAUNCH_CHECK("lm_build_tc_kernel launch");
    return BANET_OK;
}

template <int NCH, bool FLY, int MODE>
static int launch_tc(const CUtensorMap& tm, const CUtensorMap& tm2, const BuildParams& prm, int grid, cudaStream_t st)
{
    return tc_gather_warps() == 8 ? launch_tc_gw<NCH, FLY, MODE, 8>(tm, tm2, prm, grid, st) : launch_tc_gw<NCH, FLY, MODE, 16>(tm, tm2, prm, grid, st);
}

template <int NCH, bool FLY>
static int launch_tc_mode(int mode, const CUtensorMap& tm, const CUtensorMap& tm2, const BuildParams& prm, int grid, cudaStream_t st)
{
    if (mode == 1) return launch_tc<NCH, FLY, 1>(tm, tm2, prm, grid, st);
    if (mode == 2) return launch_tc<NCH, FLY, 2>(tm, tm2, prm, grid, st);
    return launch_tc<NCH, FLY, 3>(tm, tm2, prm, grid, st);
}

int lm_build_tc(const banet_level_t* lv, const BuildPlan& plan, int mode, const float* R, const float* T, const float* W,
                float* H, float* g, float* rbar_sum, float* nvalid, void* ws, cudaStream_t st)
{
    BANET_REQUIRE(tc_supported(lv), BANET_ERR_UNSUPPORTED,
                  "lm_build (tensor-core path) needs K=128, C in {64,128}, 16-B aligned tensors; got K=%d C=%d", lv->K, lv->C);
    CUtensorMap tm;
    int rc;
    if (lv->grid_w > 0) rc = make_tmap_f32_3d_sw128_32b(&tm, lv->B, (uint64_t)lv->nb * lv->grid_h, lv->grid_w, lv->K, 8, 8, 32);
    else rc = make_tmap_f32_2d_sw128_32b(&tm, lv->B, (uint64_t)lv->nb * lv->N, lv->K, TC_TILE, 32);
    if (rc) return rc;
    const bool fly = lv->conv2_channels == lv->C;
    CUtensorMap tm2;        // conv2 footprint prefetch boxes: C channels x (12|14)^2 texels
    rc = make_tmap_f32_nhwc_prefetch(&tm2, lv->conv2, lv->nb, lv->h, lv->w, lv->conv2_channels, lv->C, fly ? 14 : 12, fly ? 14 : 12);
    if (rc) return rc;
    BuildParams prm;
    prm.nb = lv->nb; prm.N = lv->N; prm.C = lv->C; prm.K = lv->K; prm.h = lv->h; prm.w = lv->w; prm.c2 = lv->conv2_channels;
    prm.conv1 = lv->conv1; prm.conv2 = lv->conv2; prm.intr = lv->intr; prm.p = lv->p; prm.D = lv->D; prm.B = lv->B;
    prm.R = R; prm.T = T; prm.W = W;
    prm.partials = reinterpret_cast<float*>(ws);
    prm.slot_floats = plan.slot_floats; prm.max_span = plan.max_span;
    prm.tiles_per_pair = plan.tiles_per_pair; prm.total_tiles = plan.total_tiles;
    prm.grid_w = lv->grid_w; prm.grid_h = lv->grid_h; prm.tiles_x = lv->grid_w > 0 ? (lv->grid_w + 7) / 8 : 0;
    prm.hdd_transposed = 1;
    { const char* e = getenv("BANET_TC_PF_CONV2"); prm.pf_conv2 = (e && atoi(e) == 1) ? 1 : 0; }
    { const char* e = getenv("BANET_TC_PF_TAPS"); prm.pf_taps = (e && atoi(e) == 1) ? 1 : 0; }
    { const char* e = getenv("BANET_TC_TRACE_PTR"); prm.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
    const int nch = lv->C / 64;
    if (tc_generation() == 6) rc = lm_build_tc6_launch(lv->K == TC_K ? mode : (mode == 3 ? 3 : 2), fly, nch, lv->K / 32, tm, prm, plan.grid, st);
    else if (nch == 2) rc = fly ? launch_tc_mode<2, true>(mode, tm, tm2, prm, plan.grid, st) : launch_tc_mode<2, false>(mode, tm, tm2, prm, plan.grid, st);
    else          rc = fly ? launch_tc_mode<1, true>(mode, tm, tm2, prm, plan.grid, st) : launch_tc_mode<1, false>(mode, tm, tm2, prm, plan.grid, st);
    if (rc) return rc;
    return launch_lm_reduce(prm, plan.grid, H, g, rbar_sum, nvalid, st);
}

}  // namespace banet
