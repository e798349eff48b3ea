// Shared helpers for the sm_100a kernels of libbanet_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/banet_abi.h"

namespace banet {

void set_error(const char* fmt, ...);

#define BANET_REQUIRE(cond, code, ...)                 \
    do { if (!(cond)) { ::banet::set_error(__VA_ARGS__); return (code); } } while (0)

#define BANET_CUDA_LAUNCH_CHECK(what)                                              \
    do { cudaError_t e__ = cudaGetLastError();                                     \
         if (e__ != cudaSuccess) { ::banet::set_error("%s: %s", what, cudaGetErrorString(e__)); \
                                   return BANET_ERR_CUDA; } } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kMaxSMs = 148;          // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not pollute L1 (read-once data: conv1, B)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// Static contiguous partition of `total` work units over `parts` workers (deterministic).
__host__ __device__ __forceinline__ long long part_begin(long long total, int parts, int i) {
    return total * (long long)i / (long long)parts;
}

// ---- layout of one partial-sum slot written by the build kernel (floats) --------------------
//   [0, K*K)            Hdd   row-major full K x K
//   [K*K, K*K+7K)       ext   rows 0..5: H_cd[i][k] ; row 6: g_d[k]
//   then 32 floats      cc    21 upper-tri H_cc (row-major i<=j) + 6 g_c + nvalid + pad
//   then C floats       rbar partial sums
struct SlotLayout {
    int K, C;
    __host__ __device__ int off_ext()  const { return K * K; }
    __host__ __device__ int off_cc()   const { return K * K + 7 * K; }
    __host__ __device__ int off_rbar() const { return K * K + 7 * K + 32; }
    __host__ __device__ int floats()   const { return ((K * K + 7 * K + 32 + C) + 3) / 4 * 4; }
};

}  // namespace banet
