// The legacy pose-only tracker loop on the device: reference legacy/ba.py:83-145 (`Tracker.trackTF`) with its two iteration kinds,
//   CameraIteration  (:147-214)  lambda = ||rbar||^2, T' = t + dr T, fixed iteration count per level, and
//   CameraIteration2 (:226-345)  lambda-MLP step, then the residual is RE-EVALUATED at the updated pose and the step is kept only if it
//                                decreased (:304-345); the level ends early once the update is small or a step was rejected (:132-141).
// The reference runs this as a tf.while_loop with a host-visible condition, one pair at a time.  Here every pair of the batch carries its own
// `active` flag on the device and the loop runs the level's maximum count without any host synchronisation:
//   build(R,T) -> [ step (lm_step_kernel, candidate pose) -> build(candidate) -> decide ] x level_iters
// The candidate's build yields exactly the residual statistics the acceptance test needs (sum |diff| per channel, in-bounds count), and when the
// step is accepted its normal equations ARE the next iteration's: an accepted iteration costs one build, like a plain one.
// Conventions of the legacy code that differ from bundlenet.py and cancel in H, g: diff = F2w - conv1 and an un-negated camera Jacobian
// (legacy/ba.py:36-48, :263).  rbar is rescaled by N / valid (:256,274); lambda = ||rbar||^(1 + tanh(MLP)) (:280); every diagonal entry is damped.
#include "common.cuh"
#include "lm_build.h"

namespace banet {

__global__ void legacy_init_kernel(int nb, int32_t* active, int32_t* iters_done) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) { active[b] = 1; if (iters_done) iters_done[b] = 0; }
}

// thread per pair: accept / reject (legacy/ba.py:304-345) + the continuation test of the while_loop (:132-133)
__global__ void legacy_decide_kernel(int nb, int C, int N, float residual_ratio, float angle_change, float translation_change,
                                     const float* __restrict__ rbar_cand, const float* __restrict__ nvalid_cand, const float* __restrict__ Hc,
                                     const float* __restrict__ gc, const float* __restrict__ Rc, const float* __restrict__ Tc, const float* __restrict__ delta,
                                     float* __restrict__ rbar_cur, float* __restrict__ nvalid_cur, float* __restrict__ H, float* __restrict__ g,
                                     float* __restrict__ R, float* __restrict__ T, int32_t* __restrict__ active, int32_t* __restrict__ iters_done,
                                     float* __restrict__ ratio_out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb || !active[b]) return;
    const float vcur = nvalid_cur[b], vcand = nvalid_cand[b];
    ratio_out[b] = (float)N / vcur;                                  // `num_valid` of this iteration (:255, returned :343-345)
    double scur = 0.0, scand = 0.0;
    for (int c = 0; c < C; ++c) { scur += rbar_cur[(size_t)b * C + c]; scand += rbar_cand[(size_t)b * C + c]; }
    const float avg_cur = (float)(scur / (double)C) / vcur, avg_cand = (float)(scand / (double)C) / vcand;       // mean_c (N/valid) mean_n |diff|
    const bool accept = avg_cand < residual_ratio * avg_cur;         // NaN / zero valid count -> false, like tf.less
    if (iters_done) iters_done[b] += 1;
    if (!accept) { active[b] = 0; return; }                          // (R,T,0,0): zero updates end the while_loop
    for (int q = 0; q < 9; ++q) R[(size_t)b * 9 + q] = Rc[(size_t)b * 9 + q];
    for (int q = 0; q < 3; ++q) T[(size_t)b * 3 + q] = Tc[(size_t)b * 3 + q];
    for (int q = 0; q < 36; ++q) H[(size_t)b * 36 + q] = Hc[(size_t)b * 36 + q];
    for (int q = 0; q < 6; ++q) g[(size_t)b * 6 + q] = gc[(size_t)b * 6 + q];
    for (int c = 0; c < C; ++c) rbar_cur[(size_t)b * C + c] = rbar_cand[(size_t)b * C + c];
    nvalid_cur[b] = vcand;
    const float* d = delta + (size_t)b * 6;
    const float uw = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), ut = sqrtf(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
    active[b] = (angle_change < uw) && (translation_change < ut);
}

__global__ void legacy_ratio_kernel(int nb, int N, const float* __restrict__ nvalid, float* __restrict__ ratio_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) ratio_out[b] = nvalid[b] / (float)N;                 // CameraIteration returns reduce_sum(mask) / npixels (:214)
}

namespace {
struct LegacyCarve { size_t build, cur, cand, lam, delta, Rc, Tc, active, st, total; int maxC; };
int legacy_carve(const banet_level_t* levels, int nlevels, LegacyCarve* c)
{
    size_t build = 0; int maxC = 0;
    const int nb = levels[0].nb;
    for (int l = 0; l < nlevels; ++l) {
        BuildPlan plan;
        int rc = build_plan(&levels[l], num_sms(), &plan);
        if (rc) return rc;
        if (plan.ws_bytes > build) build = plan.ws_bytes;
        if (levels[l].C > maxC) maxC = levels[l].C;
    }
    const size_t eq = align_up((size_t)nb * (36 + 6 + maxC + 1) * 4, 256);
    size_t off = 0;
    c->build = off; off += align_up(build, 256);
    c->cur = off; off += eq;
    c->cand = off; off += eq;
    c->lam = off; off += align_up((size_t)nb * 4, 256);
    c->delta = off; off += align_up((size_t)nb * 6 * 4, 256);
    c->Rc = off; off += align_up((size_t)nb * 9 * 4, 256);
    c->Tc = off; off += align_up((size_t)nb * 3 * 4, 256);
    c->active = off; off += align_up((size_t)nb * 4, 256);
    c->st = off; off += align_up((size_t)nb * 4, 256);
    c->total = off; c->maxC = maxC;
    return BANET_OK;
}
struct Eq { float *H, *g, *rbar, *nvalid; };
Eq eq_at(char* base, size_t off, int nb, int maxC) {
    float* p = reinterpret_cast<float*>(base + off);
    return Eq{p, p + (size_t)nb * 36, p + (size_t)nb * 42, p + (size_t)nb * (42 + maxC)};
}
}  // namespace

size_t lm_track_legacy_workspace_bytes(const banet_level_t* levels, int nlevels)
{
    LegacyCarve c;
    if (legacy_carve(levels, nlevels, &c) != BANET_OK) return 0;
    return c.total;
}

int lm_track_legacy(const banet_level_t* levels, int nlevels, const int* level_iters, const float* const* mlp_weights, const banet_legacy_opts_t& o,
                    float* R, float* T, int32_t* iters_done, float* valid_ratio, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t st)
{
    LegacyCarve c;
    int rc = legacy_carve(levels, nlevels, &c);
    if (rc) return rc;
    BANET_REQUIRE(ws && ws_bytes >= c.total, BANET_ERR_WORKSPACE, "lm_track_legacy: workspace %zu < %zu bytes", ws_bytes, c.total);
    const int nb = levels[0].nb;
    char* base = reinterpret_cast<char*>(ws);
    Eq cur = eq_at(base, c.cur, nb, c.maxC), cand = eq_at(base, c.cand, nb, c.maxC);
    float* lam = reinterpret_cast<float*>(base + c.lam);
    float* delta = reinterpret_cast<float*>(base + c.delta);
    float* Rc = reinterpret_cast<float*>(base + c.Rc);
    float* Tc = reinterpret_cast<float*>(base + c.Tc);
    int32_t* active = reinterpret_cast<int32_t*>(base + c.active);
    int32_t* st_tmp = reinterpret_cast<int32_t*>(base + c.st);
    const banet_solve_opts_t sopts = {1e-5f, 0, 0};                  // every diagonal entry damped (legacy/ba.py:200, :285)
    const int tb = (nb + 127) / 128;
    cudaMemsetAsync(status, 0, (size_t)nb * sizeof(int32_t), st);
    for (int l = 0; l < nlevels; ++l) {
        const banet_level_t* lv = &levels[l];
        BuildPlan plan;
        rc = build_plan(lv, num_sms(), &plan);
        if (rc) return rc;
        int32_t* itd = iters_done ? iters_done + (size_t)l * nb : nullptr;
        legacy_init_kernel<<<tb, 128, 0, st>>>(nb, active, itd);
        if (!o.early_termination) {
            const StepMode mode = {2.0f, 0, 0, 0};                   // lambda = ||rbar||^2 (:190), T' = t + dr T (:213), no theta clamp (:60-80)
            for (int it = 0; it < level_iters[l]; ++it) {
                rc = lm_build_simt(lv, plan, R, T, nullptr, cur.H, cur.g, cur.rbar, cur.nvalid, base + c.build, st);
                if (rc) return rc;
                rc = lm_step(cur.H, cur.g, cur.rbar, nb, lv->N, lv->C, 0, nullptr, 1.0f, nullptr, mode, cur.nvalid, sopts, R, T, nullptr, R, T, nullptr,
                             delta, lam, status, 1, st);
                if (rc) return rc;
            }
            if (level_iters[l] > 0) legacy_ratio_kernel<<<tb, 128, 0, st>>>(nb, lv->N, cur.nvalid, valid_ratio);
            continue;
        }
        BANET_REQUIRE(mlp_weights && mlp_weights[l], BANET_ERR_BAD_ARG, "lm_track_legacy: level %d has no lambda-MLP weights (CameraIteration2 needs them)", l);
        const StepMode mode = {1.0f, 1, 1, 0};                       // rbar per valid point (:274), exponent 1 + tanh (:280), V matrix (:302)
        if (level_iters[l] > 0) {
            rc = lm_build_simt(lv, plan, R, T, nullptr, cur.H, cur.g, cur.rbar, cur.nvalid, base + c.build, st);
            if (rc) return rc;
        }
        for (int it = 0; it < level_iters[l]; ++it) {
            rc = lm_step(cur.H, cur.g, cur.rbar, nb, lv->N, lv->C, 0, mlp_weights[l], 1.0f, nullptr, mode, cur.nvalid, sopts, R, T, nullptr, Rc, Tc, nullptr,
                         delta, lam, st_tmp, 0, st);
            if (rc) return rc;
            rc = lm_build_simt(lv, plan, Rc, Tc, nullptr, cand.H, cand.g, cand.rbar, cand.nvalid, base + c.build, st);
            if (rc) return rc;
            legacy_decide_kernel<<<tb, 128, 0, st>>>(nb, lv->C, lv->N, o.residual_ratio, o.angle_change, o.translation_change, cand.rbar, cand.nvalid,
                                                     cand.H, cand.g, Rc, Tc, delta, cur.rbar, cur.nvalid, cur.H, cur.g, R, T, active, itd, valid_ratio);
        }
    }
    BANET_CUDA_LAUNCH_CHECK("lm_track_legacy");
    return BANET_OK;
}

}  // namespace banet
