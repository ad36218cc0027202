// hs_stages.cuh -- the arithmetic of one control tick as per-pursuer / per-environment device functions.
// Both work decompositions call exactly these functions, in the same order, so their results are bit-identical:
//   * hs_tick.cuh       4 adjacent lanes own one environment (lanes 0..A-1 = pursuers, lane A = evader); the
//                       latency-bound small-batch mapping and the first phase of the one-launch tick + predictor kernel
//   * hs_tick_wide.cuh  one lane owns one environment and loops over its pursuers; the bandwidth-bound large-batch
//                       mapping (TMA tensor loads of the SoA state tile) and the only one for more than 3 pursuers
// Part of the single translation unit hs_kernels.cu (and of hs_tick_exact.cu, where HS_EXACT_MATH=1 swaps the SFU
// approximations for IEEE operations and the ill-conditioned stages follow the reference's operation order).
#pragma once
#include "hs_common.cuh"

namespace {

// ---- CTBR transform + body-rate PID ----------------------------------------------------------
// omni_drones/utils/torchrl/transforms.py:425-459 feeding omni_drones/controllers/lee_position_controller.py:476-550.
// Explicit round-to-nearest operations in the reference's order: the D term amplifies 1-ulp differences of the body
// rate by 180/pi * kd / dt ~ 1.4e4.
struct CtbrOut {
    float cmd[4];            // rotor commands in [-1, 1]          ("agents","action") after the transform
    float4 prev_new;         // [rate3, thrust]                    ("info","prev_action")
    float4 ctbr;             // [r, p, y, thrust * 2^16]           'ctbr'
    V3 trate;                // target body rate, deg/s            'target_rate'
    float action_err;        // |[rate, thrust] - prev_action|     ("stats","action_error_order1")
};
__device__ __forceinline__ void stage_ctbr_pid(const hs_config& c, const float4 act, const float4 prev, const bool pid_reset,
                                               const Q4 q, const V3 av, V3& integ, V3& last, CtbrOut& o) {
    using namespace ex;
    const float dt = c.dt;
    const float a0 = tanhf(act.x), a1 = tanhf(act.y), a3 = tanhf(act.w);
    float a2 = tanhf(act.z);
    const float thrust = clampf(mul(add(a3, 1.0f), 0.5f), 0.0f, c.max_thrust_ratio);
    if (c.fixed_yaw) a2 = 0.0f;
    const float d0 = sub(a0, prev.x), d1 = sub(a1, prev.y), d2 = sub(a2, prev.z), d3 = sub(thrust, prev.w);
    o.action_err = __fsqrt_rn(add(add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)), mul(d3, d3)));
    o.prev_new = make_float4(a0, a1, a2, thrust);
    o.trate = mk(mul(mul(a0, 180.0f), c.target_clip), mul(mul(a1, 180.0f), c.target_clip), mul(mul(a2, 180.0f), c.target_clip));
    const float tthrust = mul(thrust, 65536.0f);
    if (pid_reset) { integ = mk(0, 0, 0); last = mk(0, 0, 0); }
    const V3 br0 = qrot_inv_exact(q, av);
    const float pi_f = 3.14159265358979323846f;
    const V3 br = mk(div(mul(br0.x, 180.0f), pi_f), div(mul(br0.y, 180.0f), pi_f), div(mul(br0.z, 180.0f), pi_f));
    float out3[3];
    const float errv[3] = {sub(o.trate.x, br.x), sub(o.trate.y, br.y), sub(o.trate.z, br.z)};
    const float brv[3] = {br.x, br.y, br.z};
    const float lastv[3] = {last.x, last.y, last.z};
    float integv[3] = {integ.x, integ.y, integ.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float outP = mul(errv[k], c.pid_kp[k]);
        float deriv = div(-sub(brv[k], lastv[k]), dt);
        if (isnan(deriv)) deriv = 0.0f;
        const float outD = mul(deriv, c.pid_kd[k]);
        integv[k] = clampf(add(integv[k], mul(errv[k], dt)), -c.pid_ilimit[k], c.pid_ilimit[k]);
        const float outI = mul(integv[k], c.pid_ki[k]);
        float out = add(add(outP, outD), outI);
        if (isnan(out)) out = 0.0f;
        out3[k] = clampf(out, -c.pid_out_limit, c.pid_out_limit);
    }
    integ = mk(integv[0], integv[1], integv[2]);
    last = br;
    const float r = out3[0] * 0.5f, pp = out3[1] * 0.5f, y = out3[2];
    const float m[4] = {add(sub(add(tthrust, r), pp), y), sub(add(add(tthrust, r), pp), y),
                        add(add(sub(tthrust, r), pp), y), sub(sub(sub(tthrust, r), pp), y)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = sub(mul(mul(m[k], 1.0f / 65536.0f), 2.0f), c.max_thrust_ratio);
        if (isnan(v)) v = 0.0f;                       // torch.nan_to_num_(cmds, 0.)
        else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        o.cmd[k] = v;
    }
    o.ctbr = make_float4(r, pp, y, tthrust);
}

// ---- rotor model, omni_drones/actuators/rotor_group.py:55-71 ---------------------------------
__device__ __forceinline__ void stage_rotor(const hs_config& c, const float (&cmd)[4], float (&thr)[4], float (&T)[4],
                                            float& yaw_torque, float& throttle_diff) {
    float dsq = 0.f;
    yaw_torque = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float target = fsqrt(clampf((cmd[k] + 1.0f) / 2.0f, 0.0f, 1.0f));
        const float nt = thr[k] + c.rotor_alpha * (target - thr[k]);
        const float dth = nt - thr[k];
        dsq = (k == 0) ? dth * dth : dsq + dth * dth;
        thr[k] = nt;
        const float t = clampf(nt * nt + 0.0f, 0.0f, 1.0f);
        T[k] = t * c.kf;
        const float mom = (t * c.km) * (-c.rotor_dirs[k]);
        yaw_torque = (k == 0) ? mom : yaw_torque + mom;
    }
    throttle_diff = fsqrt(dsq);
}

// ---- downwash of source drone j (thrust Fj in world frame, at pj) on the drone at p -----------
// omni_drones/robots/drone/multirotor.py:488-494, 724-753
__device__ __forceinline__ V3 downwash_term(const hs_config& c, const V3 Fj, const V3 pj, const V3 p) {
    const V3 d = Fj / (norm3(Fj) + 1e-6f);
    const V3 rel = pj - p;
    const float zd = dot3(rel, d);
    const float rr = norm3(rel - d * zd);
    const float z = zd < 0.0f ? 0.0f : zd;
    const float qq = fdiv(c.downwash_kr * rr, z);
    const float den = 1.0f + c.downwash_kz * z;
    const float v = fdiv(fexp(-0.5f * (qq * qq)), den * den);
    return neg(Fj) * v;
}

// ---- evader (potential field), omni_drones/envs/hide_and_seek/hideandseek.py:1067-1141 + 737-744 ----
// repulsion of the evader at tp from the pursuer at p
__device__ __forceinline__ V3 evader_pursuer_term(const hs_config& c, const V3 p, const V3 tp, const bool blocked) {
    const V3 rel = p - tp;
    const float dist = norm3(rel);
    const float active = ((dist < c.target_detect_radius) && !blocked) ? 1.0f : 0.0f;
#if HS_EXACT_MATH
    // reference order: (-rel / (d + 1e-5)) * (1 / (d + 1e-5)) * active
    const float d1 = dist + 1e-5f;
    const float inv = __fdiv_rn(1.0f, d1);
    return mk((__fdiv_rn(-rel.x, d1) * inv) * active, (__fdiv_rn(-rel.y, d1) * inv) * active, (__fdiv_rn(-rel.z, d1) * inv) * active);
#else
    const float inv_d = frcp(dist + 1e-5f);
    return (neg(rel) * (inv_d * inv_d)) * active;
#endif
}
// force_p = sum of the pursuer terms (agent order); returns the evader's new velocity
template <int CT>
__device__ __forceinline__ V3 evader_velocity(const hs_config& c, const V3 force_p, const V3 tp, const float (&cx)[CT],
                                              const float (&cy)[CT], const float (&cz)[CT], const int C, const float v_prey,
                                              bool& out_of_arena) {
    V3 force = mk(0.f, 0.f, 0.f) + force_p;
    const float rho = fsqrt(tp.x * tp.x + tp.y * tp.y);
    out_of_arena = (tp.x * tp.x + tp.y * tp.y) > c.arena_size_sq;
    const float o = out_of_arena ? 1.0f : 0.0f, no = out_of_arena ? 0.0f : 1.0f;
    V3 fr;
#if HS_EXACT_MATH
    const float rho1 = rho + 1e-5f;
    const float inx = __fdiv_rn(-tp.x, rho1), iny = __fdiv_rn(-tp.y, rho1);
    const float wall = __fdiv_rn(1.0f, (c.arena_size - rho) + 1e-5f);
    fr.x = (o * inx) * 1e5f + (no * inx) * wall;
    fr.y = (o * iny) * 1e5f + (no * iny) * wall;
    const bool hi = tp.z > c.max_height;
    const float hz = c.max_height - tp.z;
    fr.z = (hi ? 1.0f : 0.0f) * -1e5f + __fdiv_rn((hi ? 0.0f : 1.0f) * -hz, hz * hz + 1e-5f);
    const bool lo = tp.z < 0.0f;
    const float lz = 0.0f - tp.z;
    fr.z = fr.z + ((lo ? 1.0f : 0.0f) * 1e5f + __fdiv_rn((lo ? 0.0f : 1.0f) * -lz, lz * lz + 1e-5f));
#else
    const float inv_rho = frcp(rho + 1e-5f);
    const float inx = -tp.x * inv_rho, iny = -tp.y * inv_rho;
    const float wall = frcp((c.arena_size - rho) + 1e-5f);
    fr.x = (o * inx) * 1e5f + (no * inx) * wall;
    fr.y = (o * iny) * 1e5f + (no * iny) * wall;
    const bool hi = tp.z > c.max_height;
    const float hz = c.max_height - tp.z;
    fr.z = hi ? -1e5f : fdiv(-hz, hz * hz + 1e-5f);
    const bool lo = tp.z < 0.0f;
    const float lz = 0.0f - tp.z;
    fr.z = fr.z + (lo ? 1e5f : fdiv(-lz, lz * lz + 1e-5f));
#endif
    force = force + fr;
    float fcx = 0.f, fcy = 0.f;
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        if (k < C && !(cz[k] < 0.0f)) {
            const float tx = tp.x - cx[k], ty = tp.y - cy[k];
            const float dxy = fsqrt(tx * tx + ty * ty);
            if (dxy < c.target_detect_radius) {
#if HS_EXACT_MATH
                const float d1 = dxy + 1e-5f;
                const float g = __fdiv_rn(1.0f, (dxy - c.cylinder_size) + 1e-5f);
                fcx = fcx + __fdiv_rn(tx, d1) * g;
                fcy = fcy + __fdiv_rn(ty, d1) * g;
#else
                const float sc = frcp(dxy + 1e-5f) * frcp((dxy - c.cylinder_size) + 1e-5f);
                fcx = fcx + tx * sc;
                fcy = fcy + ty * sc;
#endif
            }
        }
    }
    force = force + mk(fcx, fcy, 0.f);
    // per-component normalisation: torch.norm over the size-1 agent dim, hideandseek.py:741
    return mk(fdiv(v_prey * force.x, fabsf(force.x) + 1e-5f), fdiv(v_prey * force.y, fabsf(force.y) + 1e-5f),
              fdiv(v_prey * force.z, fabsf(force.z) + 1e-5f));
}

// ---- rigid-body integration (PhysX stand-in; oracle/hs_oracle.py rigid_body_step) ------------
// FORCED=false: the unforced tick inside reset (hideandseek.py:722-723)
template <bool FORCED>
__device__ __forceinline__ void stage_integrate(const hs_config& c, V3& p, Q4& q, V3& lv, V3& av, const float (&T)[4],
                                                const float yaw_torque, const V3 ext) {
    const float dt = c.dt;
    V3 force = mk(0.f, 0.f, 0.f), tau = mk(0.f, 0.f, 0.f);
    if (FORCED) {
        const float total_thrust = ((T[0] + T[1]) + T[2]) + T[3];
        force = qrot<false>(q, mk(0.f, 0.f, total_thrust));
        tau.x = ((c.rotor_y[0] * T[0] + c.rotor_y[1] * T[1]) + c.rotor_y[2] * T[2]) + c.rotor_y[3] * T[3];
        tau.y = (((-c.rotor_x[0]) * T[0] + (-c.rotor_x[1]) * T[1]) + (-c.rotor_x[2]) * T[2]) + (-c.rotor_x[3]) * T[3];
        tau.z = yaw_torque;
        force = force + ext;
    }
    V3 acc = force / c.total_mass;
    acc.z = acc.z - c.gravity;
    V3 v = lv + acc * dt;
    const V3 I = mk(c.inertia[0], c.inertia[1], c.inertia[2]);
    V3 wb = qrot<true>(q, av);
    const V3 gyro = cross3(wb, mk(I.x * wb.x, I.y * wb.y, I.z * wb.z));
    const V3 tg = tau - gyro;
#if HS_EXACT_MATH
    wb = wb + mk(__fdiv_rn(tg.x, I.x), __fdiv_rn(tg.y, I.y), __fdiv_rn(tg.z, I.z)) * dt;
#else
    wb = wb + mk(tg.x * c.inv_inertia[0], tg.y * c.inv_inertia[1], tg.z * c.inv_inertia[2]) * dt;
#endif
    V3 w = qrot<false>(q, wb);
    v = v * c.lin_damp_factor;
    w = w * c.ang_damp_factor;
    const float vn = norm3(v);
    if (vn > c.max_linear_velocity) v = v * fdiv(c.vmax_clamped, vn);
    float wn = norm3(w);
    if (wn > c.max_angular_velocity) w = w * fdiv(c.max_angular_velocity, wn);
    p = p + v * dt;
    wn = norm3(w);
    const float half = (0.5f * dt) * wn;
    const bool small = wn < 1e-6f;
    float sh, ch;
    sincosf(half, &sh, &ch);
    const float kk = small ? (0.5f * dt) : fdiv(sh, fmaxf(wn, 1e-6f));
    Q4 dq; dq.w = small ? 1.0f : ch; dq.x = w.x * kk; dq.y = w.y * kk; dq.z = w.z * kk;
    const Q4 qn = qmul(dq, q);
#if HS_EXACT_MATH
    const float qnorm = __fsqrt_rn(((qn.w * qn.w + qn.x * qn.x) + qn.y * qn.y) + qn.z * qn.z);
    q.w = __fdiv_rn(qn.w, qnorm); q.x = __fdiv_rn(qn.x, qnorm); q.y = __fdiv_rn(qn.y, qnorm); q.z = __fdiv_rn(qn.z, qnorm);
#else
    const float qinv = rsqrtf(((qn.w * qn.w + qn.x * qn.x) + qn.y * qn.y) + qn.z * qn.z);
    q.w = qn.w * qinv; q.x = qn.x * qinv; q.y = qn.y * qinv; q.z = qn.z * qinv;
#endif
    if (c.ground_clamp && p.z < c.ground_z) {
        p.z = c.ground_z;
        if (v.z < 0.0f) v.z = 0.0f;
    }
    lv = v; av = w;
}

// ---- analytic contacts (hs_config.contact_mode = 1; PhysX stand-in, PARITY UNPINNED) -------------------------
// After the integration: project the pursuer out of every standing cylinder it penetrates (2-D, below the cylinder top)
// and out of the evader's sphere (evader position at the start of the tick), and remove the velocity component that
// points into the obstacle (inelastic).  oracle/hs_oracle.py::apply_contacts is the same arithmetic.
template <int CT>
__device__ __forceinline__ void stage_contacts(const hs_config& c, V3& p, V3& v, const V3 tp, const float (&cx)[CT],
                                               const float (&cy)[CT], const float (&cz)[CT], const int C) {
    if (!c.contact_mode) return;
    const float Rc = c.cylinder_size + c.drone_radius;
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        if (k < C && cz[k] > 0.0f && p.z < 2.0f * cz[k]) {
            const float dx = p.x - cx[k], dy = p.y - cy[k];
            const float d = fsqrt(dx * dx + dy * dy);
            if (d < Rc) {
                const float inv = frcp(fmaxf(d, 1e-6f));
                const float nx = dx * inv, ny = dy * inv;
                p.x = cx[k] + nx * Rc;
                p.y = cy[k] + ny * Rc;
                const float vn = v.x * nx + v.y * ny;
                if (vn < 0.0f) { v.x = v.x - vn * nx; v.y = v.y - vn * ny; }
            }
        }
    }
    const float Re = c.evader_radius + c.drone_radius;
    const V3 rel = p - tp;
    const float d = norm3(rel);
    if (d < Re) {
        const V3 n = rel * frcp(fmaxf(d, 1e-6f));
        p = tp + n * Re;
        const float vn = dot3(v, n);
        if (vn < 0.0f) v = v - n * vn;
    }
}

// ---- k nearest cylinders [K,5] of the pursuer at p; lowest index wins ties (hideandseek.py:757-778) ----
// also counts the cylinder collisions among them (hideandseek.py:961-968)
template <int CT>
__device__ __forceinline__ void stage_knearest(const hs_config& c, const V3 p, const float (&cx)[CT], const float (&cy)[CT],
                                               const float (&cz)[CT], const int C, const int K, float* r, float& hit_cyl) {
    float key[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k)
        key[k] = (k < C) ? (norm3(mk(p.x - cx[k], p.y - cy[k], p.z - cz[k])) - c.cylinder_size) : INFINITY;
    unsigned taken = 0u;
    hit_cyl = 0.f;
#pragma unroll
    for (int n = 0; n < KMAX; ++n) {
        if (n < K) {
            int best = 0; float bk = INFINITY; bool found = false;
#pragma unroll
            for (int k = 0; k < CT; ++k) {
                const bool cand = (k < C) && !((taken >> k) & 1u);
                if (cand && (!found || key[k] < bk)) { best = k; bk = key[k]; found = true; }
            }
            taken |= 1u << best;
            float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
            for (int k = 0; k < CT; ++k) if (k == best) { bx = cx[k]; by = cy[k]; bz = cz[k]; }
            const bool inactive = bz < 0.0f;
            const float rx = p.x - bx, ry = p.y - by, rz = p.z - bz;
            const float mv = c.mask_value;
            r[n * 5 + 0] = inactive ? mv : rx;
            r[n * 5 + 1] = inactive ? mv : ry;
            r[n * 5 + 2] = inactive ? mv : rz;
            r[n * 5 + 3] = inactive ? mv : c.max_height;
            r[n * 5 + 4] = inactive ? mv : c.cylinder_size;
            const float dxy = fsqrt(rx * rx + ry * ry);
            const float hit = ((dxy - c.cylinder_size) < c.collision_radius) ? 1.0f : 0.0f;
            hit_cyl = hit_cyl + (inactive ? 0.0f : hit);
        }
    }
}

// ---- per-pursuer reward terms, hideandseek.py:919-1006 (the cooperative terms are added by the caller) ----
struct RewardTerms { float r_dist, r_speed, r_coll, r_smooth, hit_wall; bool seen_capture; };
__device__ __forceinline__ RewardTerms stage_reward_terms(const hs_config& c, const V3 p, const V3 lv, const V3 tp, const bool blocked,
                                                         const float hit_cyl, const float hit_drone, const float action_err,
                                                         const float smoothness_coef) {
    RewardTerms t;
    const float dist = norm3(tp - p);
    t.r_dist = (-c.dist_reward_coef * dist) * ((dist > c.catch_radius) ? 1.0f : 0.0f);
    t.seen_capture = (dist < c.catch_radius) && !blocked;
    t.r_speed = -c.speed_coef * ((norm3(lv) > c.v_drone) ? 1.0f : 0.0f);
    t.r_coll = -c.collision_coef * hit_cyl;
    t.r_coll = t.r_coll + (-c.collision_coef * hit_drone);
    t.hit_wall = ((p.z > c.max_height) ? 1.0f : 0.0f) + (((p.x * p.x + p.y * p.y) > c.arena_size_sq) ? 1.0f : 0.0f);
    t.r_coll = t.r_coll + (-c.collision_coef * t.hit_wall);
    t.r_smooth = c.smoothness_gated ? 0.0f : smoothness_coef * fexp(-action_err);
    return t;
}

// ---- per-environment statistics, hideandseek.py:400-425, 1017-1056 ---------------------------
// the per-env means over the pursuers (sum in agent order, then * 1/A like torch.mean) and maxima of one tick
struct EnvTick {
    float m_ae, m_dist, m_detect, m_catch, m_speed, m_hcyl, m_hdrone, m_hwall, m_coll, m_smooth, m_tdiff, m_reward, x_tdiff;
    float r_catch;
    bool bdetect, all_blocked, any_coll, out_of_arena;
};
// OLD(k) reads the previous value of stat k, ST(k, v) stores the new one
template <class Old, class Store>
__device__ __forceinline__ void stage_stats(const hs_config& c, const EnvTick& t, const float progress, const float smoothness_coef,
                                            Old OLD, Store ST) {
    const bool done = progress >= (float)c.max_episode_length;
    const float inv_len = done ? frcp(progress) : 1.0f;
#if HS_EXACT_MATH
#define HS_DIVLEN(x) (done ? __fdiv_rn((x), progress) : (x))
#else
#define HS_DIVLEN(x) ((x) * inv_len)
#endif
    (void)inv_len;
    // accumulators that are divided by the episode length on the done tick
    ST(HS_STAT_ACTION_ERROR_MEAN, HS_DIVLEN(OLD(HS_STAT_ACTION_ERROR_MEAN) + t.m_ae));
    ST(HS_STAT_ACTION_ERROR_MAX, fmaxf(OLD(HS_STAT_ACTION_ERROR_MAX), t.m_ae));
    ST(HS_STAT_OUT_OF_ARENA, ((OLD(HS_STAT_OUT_OF_ARENA) != 0.0f) || t.out_of_arena) ? 1.0f : 0.0f);
    ST(HS_STAT_DISTANCE_REWARD, HS_DIVLEN(OLD(HS_STAT_DISTANCE_REWARD) + t.m_dist));
    ST(HS_STAT_SUM_DETECT_STEP, OLD(HS_STAT_SUM_DETECT_STEP) + 1.0f * (t.bdetect ? 1.0f : 0.0f));
    ST(HS_STAT_DETECT_REWARD, HS_DIVLEN(OLD(HS_STAT_DETECT_REWARD) + t.m_detect));
    ST(HS_STAT_BLOCKED, OLD(HS_STAT_BLOCKED) + (t.all_blocked ? 1.0f : 0.0f));
    const bool capture_flag = t.r_catch != 0.0f;
    ST(HS_STAT_SUCCESS, (capture_flag || (OLD(HS_STAT_SUCCESS) != 0.0f)) ? 1.0f : 0.0f);
    const float step_now = (capture_flag ? 1.0f : 0.0f) * progress + (capture_flag ? 0.0f : 1.0f) * (float)c.max_episode_length;
    ST(HS_STAT_FIRST_CAPTURE_STEP, fminf(OLD(HS_STAT_FIRST_CAPTURE_STEP), step_now));
    ST(HS_STAT_CATCH_REWARD, HS_DIVLEN(OLD(HS_STAT_CATCH_REWARD) + t.m_catch));
    ST(HS_STAT_SPEED_REWARD, HS_DIVLEN(OLD(HS_STAT_SPEED_REWARD) + t.m_speed));
    ST(HS_STAT_COLLISION_CYLINDER, HS_DIVLEN(OLD(HS_STAT_COLLISION_CYLINDER) + t.m_hcyl));
    ST(HS_STAT_COLLISION_DRONE, HS_DIVLEN(OLD(HS_STAT_COLLISION_DRONE) + t.m_hdrone));
    ST(HS_STAT_COLLISION, HS_DIVLEN(OLD(HS_STAT_COLLISION) + (t.any_coll ? 1.0f : 0.0f)));
    ST(HS_STAT_COLLISION_WALL, HS_DIVLEN(OLD(HS_STAT_COLLISION_WALL) + t.m_hwall));
    ST(HS_STAT_COLLISION_REWARD, HS_DIVLEN(OLD(HS_STAT_COLLISION_REWARD) + t.m_coll));
    if (c.write_smoothness_coef_stat) ST(HS_STAT_SMOOTHNESS_COEF, smoothness_coef);
    ST(HS_STAT_SMOOTHNESS_REWARD, HS_DIVLEN(OLD(HS_STAT_SMOOTHNESS_REWARD) + t.m_smooth));
    ST(HS_STAT_SMOOTHNESS_MEAN, HS_DIVLEN(OLD(HS_STAT_SMOOTHNESS_MEAN) + t.m_tdiff));
    ST(HS_STAT_SMOOTHNESS_MAX, fmaxf(t.x_tdiff, OLD(HS_STAT_SMOOTHNESS_MAX)));
    ST(HS_STAT_RETURN, OLD(HS_STAT_RETURN) + t.m_reward);
    // target_predicted_error is only ever divided (stays 0); distance_predicted_reward and
    // distance_threshold_L are never written (hideandseek.py:1023-1025).
#undef HS_DIVLEN
}

}  // namespace
