// Internal interfaces between the translation units of libbanet_sm100.so.
#pragma once
#include "common.cuh"

namespace banet {

struct BuildParams {
    int nb, N, C, K, h, w, c2;
    const float *conv1, *conv2, *intr, *p, *D, *B, *R, *T, *W;
    float* partials;
    int slot_floats, max_span, tiles_per_pair;
    long long total_tiles;
    int grid_w, grid_h, tiles_x;      // dense-grid 8x8 tiling (tensor-core path); grid_w == 0 -> linear 64-pixel tiles
    int hdd_transposed;
    int pf_taps;                      // tuning (generation 5 only, BANET_TC_PF_TAPS=1): per-lane L2 prefetch of the next tile's tap lines; off: not a win
    int pf_conv2;                     // tuning (generation 5 only, BANET_TC_PF_CONV2=1): predicted conv2-footprint L2 prefetch; off: not a win
    long long* trace;                 // optional debug timeline buffer (NULL in production)               // tensor-core path stores the H_dd block of a slot column-major (coalesced TMEM drains)
};

struct BuildPlan {
    int KP, grid, max_span, slot_floats, tiles_per_pair;
    long long total_tiles;
    size_t ws_bytes;
};

int num_sms();

// fp32 SIMT path (lm_build.cu)
int build_plan(const banet_level_t* lv, int num_sms, BuildPlan* plan);
int lm_build_simt(const banet_level_t* lv, const BuildPlan& plan, const float* R, const float* T, const float* W,
                  float* H, float* g, float* rbar_sum, float* nvalid, void* ws, cudaStream_t st);

int launch_lm_reduce(const BuildParams& prm, int grid_build, float* H, float* g, float* rbar_sum, float* nvalid, cudaStream_t st);

// tensor-core path (lm_build_tc.cu): K = 128, C in {64,128}
bool tc_supported(const banet_level_t* lv);
int build_plan_tc(const banet_level_t* lv, int num_sms, BuildPlan* plan);
int lm_build_tc(const banet_level_t* lv, const BuildPlan& plan, int mode, const float* R, const float* T, const float* W,
                float* H, float* g, float* rbar_sum, float* nvalid, void* ws, cudaStream_t st);

// precision resolution + dispatch (abi.cu)
int resolve_precision(const banet_level_t* lv, int precision);      // -> BANET_PREC_* actually used, or <0 (error set)
int plan_for(const banet_level_t* lv, int resolved, BuildPlan* plan);
int build_dispatch(const banet_level_t* lv, int resolved, const BuildPlan& plan, const float* R, const float* T, const float* W,
                   float* H, float* g, float* rbar_sum, float* nvalid, void* ws, cudaStream_t st);

// lambda MLP / solve / update (lm_solve.cu)
int lm_lambda(const float* rbar_sum, int nb, int N, int C, const float* mlp, float base, float* lambda_out, cudaStream_t st);
int lm_solve_update(const float* H, const float* g, const float* lambda, int nb, int K, const banet_solve_opts_t& opts,
                    const float* R, const float* T, const float* W, float* R_out, float* T_out, float* W_out,
                    float* delta, int32_t* status, int status_accumulate, cudaStream_t st);

}  // namespace banet
