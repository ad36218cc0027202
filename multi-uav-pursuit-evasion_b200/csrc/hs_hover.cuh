// hs_hover.cuh -- Hover (BASELINE config 1): the task's observation / reward / stats after a tick, one thread per env.
// omni_drones/envs/single/hover.py:334-523.  The vehicle (CTBR transform, rate PID, rotors, rigid body) is the same tick
// kernel as HideAndSeek with one pursuer, no cylinders and the evader slot parked far away.
#pragma once
#include "hs_common.cuh"

namespace {

// stats slots: declaration order of the reference's stats spec (hover.py:239-279)
enum { HV_RETURN = 0, HV_POS_BONUS, HV_HEAD_BONUS, HV_REWARD_POS, HV_REWARD_UP, HV_REWARD_VEL, HV_REWARD_ACC, HV_REWARD_JERK,
       HV_EPISODE_LEN, HV_POS_ERROR, HV_HEADING_ALIGNMENT, HV_UPRIGHTNESS, HV_ACTION_SMOOTHNESS, HV_LIN_V_MAX, HV_ANG_V_MAX,
       HV_LIN_A_MAX, HV_ANG_A_MAX, HV_LIN_J_MAX, HV_ANG_J_MAX, HV_LIN_V_MEAN, HV_ANG_V_MEAN, HV_LIN_A_MEAN, HV_ANG_A_MEAN,
       HV_LIN_J_MEAN, HV_ANG_J_MEAN, HV_MOTOR1, HV_MOTOR2, HV_MOTOR3, HV_MOTOR4, HV_CMD_R, HV_CMD_P, HV_CMD_Y, HV_CMD_THRUST,
       HV_TARGET_R, HV_TARGET_P, HV_TARGET_Y, HV_REAL_R, HV_REAL_P, HV_REAL_Y };
// persistent state rows
enum { HS_LAST_LV = 0, HS_LAST_AV, HS_LAST_LA, HS_LAST_AA, HS_LAST_LJ, HS_LAST_AJ, HS_SUM_LV, HS_SUM_AV, HS_SUM_LA, HS_SUM_AA,
       HS_SUM_LJ, HS_SUM_AJ };

__global__ void __launch_bounds__(128)
hs_hover_post_kernel(const __grid_constant__ KParams P, const hs_hover_params hp, const hs_hover_io io) {
    const hs_config& c = P.c;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int E = c.num_envs;
    if (e >= E) return;
    constexpr int A = 1;
    const int slot = 0;
    (void)slot;
    const float* ds = P.b.drone_state + e * 13;
    const V3 p = mk(ds[0], ds[1], ds[2]);
    Q4 q; q.w = ds[3]; q.x = ds[4]; q.y = ds[5]; q.z = ds[6];
    const V3 lv = mk(ds[7], ds[8], ds[9]), av = mk(ds[10], ds[11], ds[12]);
    const float progress = *EROW(E_PROGRESS);
    float* S = io.stats + e;
    float* X = io.state + e;
    const int64_t Es = E;
#define ST(k) S[(int64_t)(k) * Es]
#define XS(k) X[(int64_t)(k) * Es]
    if (hp.with_reward) {
        // _pre_sim_step's logging (hover.py:334-359): motor commands, CTBR command, target rates
        const float4 cmd = *(reinterpret_cast<const float4*>(P.b.rotor_cmds) + e);
        const float4 ct = *(reinterpret_cast<const float4*>(P.b.ctbr) + e);
        ST(HV_MOTOR1) = cmd.x; ST(HV_MOTOR2) = cmd.y; ST(HV_MOTOR3) = cmd.z; ST(HV_MOTOR4) = cmd.w;
        ST(HV_CMD_R) = ct.x; ST(HV_CMD_P) = ct.y; ST(HV_CMD_Y) = ct.z; ST(HV_CMD_THRUST) = ct.w;
        ST(HV_TARGET_R) = P.b.target_rate[e * 3]; ST(HV_TARGET_P) = P.b.target_rate[e * 3 + 1]; ST(HV_TARGET_Y) = P.b.target_rate[e * 3 + 2];
    }
    // ---- _compute_state_and_obs, hover.py:361-437
    const V3 br0 = qrot_inv_exact(q, av);
    const float pi_f = 3.14159265358979323846f;
    ST(HV_REAL_R) = ex::div(ex::mul(br0.x, 180.0f), pi_f);
    ST(HV_REAL_P) = ex::div(ex::mul(br0.y, 180.0f), pi_f);
    ST(HV_REAL_Y) = ex::div(ex::mul(br0.z, 180.0f), pi_f);
    V3 heading, up;
    heading_up(q, heading, up);
    const V3 rpos = mk(hp.target_pos[0] - p.x, hp.target_pos[1] - p.y, hp.target_pos[2] - p.z);
    const V3 th = mk(io.target_heading[e * 3], io.target_heading[e * 3 + 1], io.target_heading[e * 3 + 2]);
    const V3 rheading = th - heading;
    const int D = 16 + (hp.omega ? 3 : 0) + (hp.motor ? 4 : 0) + (hp.time_encoding ? 4 : 0);
    float* o = io.observation + e * D;
    o[0] = rpos.x; o[1] = rpos.y; o[2] = rpos.z;
    o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z; o[7] = lv.x; o[8] = lv.y; o[9] = lv.z;
    o[10] = heading.x; o[11] = heading.y; o[12] = heading.z; o[13] = up.x; o[14] = up.y; o[15] = up.z;
    int k = 16;
    if (hp.omega) { o[k] = av.x; o[k + 1] = av.y; o[k + 2] = av.z; k += 3; }
    if (hp.motor) { for (int r = 0; r < 4; ++r) o[k + r] = *DROW(D_THR + r) * 2.0f - 1.0f; k += 4; }
    if (hp.time_encoding) { const float t = fdiv(progress, (float)c.max_episode_length); o[k] = o[k + 1] = o[k + 2] = o[k + 3] = t; }
    // velocity / acceleration / jerk magnitudes with running max and episode mean (:388-417)
    const float inv_n = frcp(progress + 1.0f), inv_dt = frcp(c.dt);
    const float lin_v = norm3(lv), ang_v = norm3(av);
    const float lin_a = fabsf(lin_v - XS(HS_LAST_LV)) * inv_dt, ang_a = fabsf(ang_v - XS(HS_LAST_AV)) * inv_dt;
    const float lin_j = fabsf(lin_a - XS(HS_LAST_LA)) * inv_dt, ang_j = fabsf(ang_a - XS(HS_LAST_AA)) * inv_dt;
    const float vals[6] = {lin_v, ang_v, lin_a, ang_a, lin_j, ang_j};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float v = fabsf(vals[i]);
        ST(HV_LIN_V_MAX + i) = fmaxf(ST(HV_LIN_V_MAX + i), v);
        const float sum = XS(HS_SUM_LV + i) + v;
        XS(HS_SUM_LV + i) = sum;
        ST(HV_LIN_V_MEAN + i) = sum * inv_n;
        XS(HS_LAST_LV + i) = vals[i];
    }
    if (!hp.with_reward) return;
    // ---- _compute_reward_and_done, hover.py:439-523
    const float pos_error = norm3(rpos), head_error = norm3(rheading);
    const float heading_alignment = dot3(heading, th);
    const float reward_pos = -pos_error * hp.reward_distance_scale;
    const float bonus = (pos_error <= 0.02f) ? 10.0f : 0.0f;
    const float near = (bonus > 0.0f) ? 1.0f : 0.0f;
    const float reward_head = -head_error * near;
    const float head_bonus = ((head_error <= 0.02f) ? 10.0f : 0.0f) * near;
    const float upz = (up.z + 1.0f) / 2.0f;
    const float reward_up = upz * upz;
    const float reward_v = (hp.reward_v_scale * near) * ((lin_v < hp.linear_vel_max) ? 1.0f : 0.0f);
    const float reward_acc = (hp.reward_acc_scale * near) * ((lin_a < hp.linear_acc_max) ? 1.0f : 0.0f);
    const float reward_jerk = (hp.reward_jerk_scale * near) * (-lin_j);
    const float reward = ((((((reward_pos + bonus) + reward_head) + head_bonus) + reward_up) + reward_v) + reward_acc) + reward_jerk;
    io.reward[e] = reward;
    io.done[e] = (progress >= (float)c.max_episode_length) ? 1 : 0;
    const float w = 1.0f - hp.alpha;                      // Tensor.lerp_(end, w) = start + w * (end - start)
    ST(HV_POS_ERROR) = ST(HV_POS_ERROR) + w * (pos_error - ST(HV_POS_ERROR));
    ST(HV_HEADING_ALIGNMENT) = ST(HV_HEADING_ALIGNMENT) + w * (heading_alignment - ST(HV_HEADING_ALIGNMENT));
    ST(HV_UPRIGHTNESS) = ST(HV_UPRIGHTNESS) + w * (up.z - ST(HV_UPRIGHTNESS));
    const float tdiff = (P.b.throttle_diff != nullptr) ? P.b.throttle_diff[e] : 0.0f;
    ST(HV_ACTION_SMOOTHNESS) = ST(HV_ACTION_SMOOTHNESS) + w * (-tdiff - ST(HV_ACTION_SMOOTHNESS));
    ST(HV_RETURN) = ST(HV_RETURN) + reward;
    ST(HV_REWARD_POS) = reward_pos; ST(HV_POS_BONUS) = bonus; ST(HV_HEAD_BONUS) = head_bonus;
    ST(HV_REWARD_VEL) = reward_v; ST(HV_REWARD_ACC) = reward_acc; ST(HV_REWARD_JERK) = reward_jerk;
    ST(HV_EPISODE_LEN) = progress;
#undef ST
#undef XS
}

}  // namespace
