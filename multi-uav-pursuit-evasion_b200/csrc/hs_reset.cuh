// hs_reset.cuh -- reset scatter kernel and the AoS <-> arena field copy
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"

namespace {

// =========================================================================================
// Reset scatter: hideandseek.py:698-717, multirotor.py:635-650
// =========================================================================================
template <int A>
__global__ void __launch_bounds__(128)
hs_reset_scatter_kernel(const __grid_constant__ KParams P) {
    const hs_config& c = P.c;
    // one thread per (env, body): bodies 0..A-1 = pursuers, body A = the evader + cylinders + stats of the env
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = gt / (A + 1);
    const int slot = (int)(gt - e * (A + 1));
    const int E = c.num_envs;
    if (e >= E) return;
    const bool masked = (P.env_mask == nullptr) || (P.env_mask[e] != 0);
    const int C = c.num_cylinders;
    if (slot < A && masked) {
        const int64_t row = e * A + slot;
        for (int k = 0; k < 3; ++k) *DROW(D_POS + k) = P.init_drone_pos[row * 3 + k];
        for (int k = 0; k < 4; ++k) *DROW(D_ROT + k) = P.init_drone_rot[row * 4 + k];
        for (int k = 0; k < 3; ++k) { *DROW(D_LIN + k) = 0.f; *DROW(D_ANG + k) = 0.f; }
        const float h = c.hover_throttle;
        for (int k = 0; k < 4; ++k) *DROW(D_THR + k) = h;
        const float cmd_init = 2.0f * (h * h) - 1.0f;
        P.b.prev_action[row * 4 + 3] = 0.5f * (c.max_thrust_ratio + cmd_init);
    }
    if (slot == A) {
        if (masked) {
            for (int k = 0; k < 3; ++k) *EROW(E_TPOS + k) = P.init_target_pos[e * 3 + k];
            for (int k = 0; k < 3 * C; ++k) *EROW(E_CYL + k) = P.init_cyl_pos[e * 3 * C + k];
            for (int k = 0; k < HS_NUM_STATS; ++k) P.b.stats[(int64_t)k * E + e] = 0.f;
        }
        // every env, not just the masked ones (hideandseek.py:712)
        P.b.stats[(int64_t)HS_STAT_FIRST_CAPTURE_STEP * E + e] = (float)c.max_episode_length;
    }
}

// ---- AoS <-> arena field copies (views/* replacement, used by tests and tools) -----------
__global__ void hs_field_copy_kernel(float* arena, int R, int row0, int n_slots, int width,
                                     int row_stride_slot, int row_stride_comp, int E, float* aos, int to_aos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_env = (int64_t)n_slots * width;
    if (i >= per_env * E) return;
    const int64_t e = i / per_env;
    const int r = (int)(i - e * per_env);
    const int a = r / width, k = r - a * width;
    const int64_t arow = (int64_t)row0 + (int64_t)k * row_stride_comp + (int64_t)a * row_stride_slot;
    float* ap = arena + ((e >> 5) * R + arow) * 32 + (e & 31);
    if (to_aos) aos[i] = *ap; else *ap = aos[i];
}


}  // namespace
