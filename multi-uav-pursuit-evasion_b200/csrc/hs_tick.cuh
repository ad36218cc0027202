// hs_tick.cuh -- hs_tick_kernel (the fused control tick) and hs_fill_kernel (prediction-dependent rows from a caller-supplied prediction)
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"

namespace {

// =========================================================================================
// The tick.  RESET=false: full control tick.  RESET=true: the unforced physics tick + obs
// that closes a reset (hideandseek.py:722-723, isaac_env.py:220-224).
// =========================================================================================
// The body works on ONE warp's 8 envs (warp_g = index of the warp in the batch) with the per-warp shared-memory
// pieces passed in, so that it can run as hs_tick_kernel or as the first phase of the fused tick + predictor
// kernel (hs_tick_tp_fused_kernel, hs_predictor_tcgen05.cuh).  Only warp-level synchronisation inside.
constexpr int TICK_STAT_WORDS = ENVS_PER_WARP * HS_NUM_STATS;

// previous TP window -> shared tile for the uncommon shapes (ragged tile, H != 10, frame width not a multiple of 4)
__device__ __noinline__ void tp_prefetch_slow(float* tp_tile, const float* src, int nenv, int per_env, int keep, int FD, int lane) {
    if ((FD & 3) == 0) {
        const int pe4 = per_env >> 2, keep4 = keep >> 2, fd4 = FD >> 2;
        const int total = nenv * keep4;
        int env = 0, j = lane;                       // i = env*keep4 + j, kept incrementally
        for (int i = lane; i < total; i += 32, j += 32) {
            while (j >= keep4) { j -= keep4; ++env; }
            cp_async16(reinterpret_cast<float4*>(tp_tile) + env * pe4 + j,
                       reinterpret_cast<const float4*>(src) + env * pe4 + j + fd4);
        }
    } else {
        const int total = nenv * keep;
        int env = 0, j = lane;
        for (int i = lane; i < total; i += 32, j += 32) {
            while (j >= keep) { j -= keep; ++env; }
            cp_async4(tp_tile + env * per_env + j, src + env * per_env + j + FD);
        }
    }
}
template <int A, bool RESET, int CT>
__device__ __forceinline__ void hs_tick_body(const KParams& P, const int64_t warp_g, float* stage0, float* stage1,
                                             float* tp_tile, float* stat_tile) {
    const hs_config& c = P.c;
    const int lane = threadIdx.x & 31;
    const int slot = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const int64_t e0 = warp_g * ENVS_PER_WARP;           // first env of this warp
    const int E = c.num_envs;
    if (e0 >= E) return;                                 // whole warp out of range
    const int64_t e_raw = e0 + (lane >> 2);
    const bool valid = e_raw < E;
    const int64_t e = valid ? e_raw : (E - 1);           // clamp: idle lanes shadow the last env (no stores)
    const bool is_drone = slot < A;
    const bool is_ev = slot == A;
    const int nenv = (int)min((int64_t)ENVS_PER_WARP, E - e0);
    const bool full_tile = nenv == ENVS_PER_WARP;
    const int C = c.num_cylinders, K = c.obs_max_cylinder;
    const int FD = 7 + 3 * A;
    const int H = c.history_step;
    const float dt = c.dt;
    // arena offsets fit 32 bits (checked at hs_create): one IMAD + one wide add per access
    const uint32_t Ep32 = (uint32_t)P.Ep;
    const uint32_t o_drone = (uint32_t)slot * Ep32 + (uint32_t)e;     // + k * (A*Ep32)
    const uint32_t o_env = (uint32_t)(ND * A) * Ep32 + (uint32_t)e;   // + k * Ep32
    float* const arena = P.b.arena;
#undef DROW
#undef EROW
#define DROW(k) (arena + (o_drone + (uint32_t)(k) * ((uint32_t)A * Ep32)))
#define EROW(k) (arena + (o_env + (uint32_t)(k) * Ep32))

    Stager st;
    st.buf[0] = stage0;
    st.buf[1] = stage1;
    st.cur = 0;
    st.lane = lane;

    // ---- prefetch (no registers held): the previous TP_input rows 1..H-1 land in the warp's
    // shared tile already shifted to rows 0..H-2, and the env's stats row lands in stat_mem.
    const int per_env = H * FD, keep = (H - 1) * FD;
    if (c.use_tp_net && !P.tp_init) {
        const float* src = P.b.tp_input_prev + e0 * per_env;
        if ((FD & 3) == 0 && H == 10 && full_tile) {
            // common shape (A=3, H=10): 8 envs x 36 float4 = 9 per lane, all indices compile-time
            constexpr int fd4 = FD / 4, pe4 = 10 * fd4, keep4 = 9 * fd4;
#pragma unroll
            for (int it = 0; it < (ENVS_PER_WARP * keep4 + 31) / 32; ++it) {
                const int i = it * 32 + lane;
                const int env = i / keep4, j = i - env * keep4;      // division by a constant
                if (i < ENVS_PER_WARP * keep4)
                    cp_async16(reinterpret_cast<float4*>(tp_tile) + env * pe4 + j,
                               reinterpret_cast<const float4*>(src) + env * pe4 + j + fd4);
            }
        } else {
            tp_prefetch_slow(tp_tile, src, nenv, per_env, keep, FD, lane);
        }
    }
    if (!RESET && valid && is_ev) {
#pragma unroll
        for (int k = 0; k < HS_NUM_STATS; ++k) cp_async4(stat_tile + (lane >> 2) * HS_NUM_STATS + k, P.b.stats + (int64_t)k * E + e);
    }
    cp_async_commit();

    // ---- load state ------------------------------------------------------------------
    V3 p = mk(0, 0, 0), lv = mk(0, 0, 0), av = mk(0, 0, 0);
    Q4 q; q.w = 1.f; q.x = q.y = q.z = 0.f;
    float thr[4] = {0, 0, 0, 0};
    V3 integ = mk(0, 0, 0), last = mk(0, 0, 0);
    if (is_drone) {
        p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
        q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
        lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
        av = mk(*DROW(D_ANG), *DROW(D_ANG + 1), *DROW(D_ANG + 2));
        if (!RESET) {
#pragma unroll
            for (int k = 0; k < 4; ++k) thr[k] = *DROW(D_THR + k);
            integ = mk(*DROW(D_INT), *DROW(D_INT + 1), *DROW(D_INT + 2));
            last = mk(*DROW(D_LAST), *DROW(D_LAST + 1), *DROW(D_LAST + 2));
        }
    }
    // per-env scalars: every lane of the group reads the same address (one broadcast sector)
    V3 tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
    V3 tv = mk(*EROW(E_TVEL), *EROW(E_TVEL + 1), *EROW(E_TVEL + 2));
    float progress = *EROW(E_PROGRESS);
    float4 act4 = make_float4(0.f, 0.f, 0.f, 0.f), prev4 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool pid_reset = false;
    float v_prey = 0.f;
    if (!RESET) {
        if (is_drone) {
            const int64_t row = e * A + slot;
            act4 = __ldg(reinterpret_cast<const float4*>(P.action) + row);
            if (P.action_is_raw) {
                prev4 = *(reinterpret_cast<const float4*>(P.b.prev_action) + row);
                pid_reset = (P.reset_pid != nullptr) && (P.reset_pid[e] != 0);
            }
        }
        if (is_ev) v_prey = __ldg(P.b.v_prey);
    }
    float cx[CT], cy[CT], cz[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        if (k < C) {
            cx[k] = __ldg(EROW(E_CYL + 3 * k));
            cy[k] = __ldg(EROW(E_CYL + 3 * k + 1));
            cz[k] = __ldg(EROW(E_CYL + 3 * k + 2));
        } else { cx[k] = 0.f; cy[k] = 0.f; cz[k] = -20.f; }
    }

    float action_err = 0.f, throttle_diff = 0.f;
    float T[4] = {0, 0, 0, 0};
    float yaw_torque = 0.f;
    V3 ext = mk(0, 0, 0);
    bool out_of_arena = false;

    if (!RESET) {
        // ---- CTBR transform + body-rate PID (transforms.py:425-459, lee_position_controller.py:476-550)
        float cmd[4] = {0, 0, 0, 0};
        if (is_drone) {
            const int64_t row = e * A + slot;
            const float4 act = act4;
            if (P.action_is_raw) {
                using namespace ex;
                const float4 prev = prev4;
                const float a0 = tanhf(act.x), a1 = tanhf(act.y), a3 = tanhf(act.w);
                float a2 = tanhf(act.z);
                const float thrust = clampf(mul(add(a3, 1.0f), 0.5f), 0.0f, c.max_thrust_ratio);
                if (c.fixed_yaw) a2 = 0.0f;
                const float d0 = sub(a0, prev.x), d1 = sub(a1, prev.y), d2 = sub(a2, prev.z), d3 = sub(thrust, prev.w);
                action_err = __fsqrt_rn(add(add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)), mul(d3, d3)));
                if (valid) *(reinterpret_cast<float4*>(P.b.prev_action) + row) = make_float4(a0, a1, a2, thrust);
                const V3 trate = mk(mul(mul(a0, 180.0f), c.target_clip), mul(mul(a1, 180.0f), c.target_clip),
                                    mul(mul(a2, 180.0f), c.target_clip));
                const float tthrust = mul(thrust, 65536.0f);
                if (pid_reset) { integ = mk(0, 0, 0); last = mk(0, 0, 0); }
                const V3 br0 = qrot_inv_exact(q, av);
                const float pi_f = 3.14159265358979323846f;
                const V3 br = mk(div(mul(br0.x, 180.0f), pi_f), div(mul(br0.y, 180.0f), pi_f), div(mul(br0.z, 180.0f), pi_f));
                float o[3];
                const float errv[3] = {sub(trate.x, br.x), sub(trate.y, br.y), sub(trate.z, br.z)};
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
                    o[k] = clampf(out, -c.pid_out_limit, c.pid_out_limit);
                }
                integ = mk(integv[0], integv[1], integv[2]);
                last = br;
                const float r = o[0] * 0.5f, pp = o[1] * 0.5f, y = o[2];
                const float m[4] = {add(sub(add(tthrust, r), pp), y), sub(add(add(tthrust, r), pp), y),
                                    add(add(sub(tthrust, r), pp), y), sub(sub(sub(tthrust, r), pp), y)};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v = sub(mul(mul(m[k], 1.0f / 65536.0f), 2.0f), c.max_thrust_ratio);
                    if (isnan(v)) v = 0.0f;                       // torch.nan_to_num_(cmds, 0.)
                    else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
                    cmd[k] = v;
                }
                if (valid) {
                    *(reinterpret_cast<float4*>(P.b.rotor_cmds) + row) = make_float4(cmd[0], cmd[1], cmd[2], cmd[3]);
                    *(reinterpret_cast<float4*>(P.b.ctbr) + row) = make_float4(r, pp, y, tthrust);
                    P.b.target_rate[row * 3 + 0] = trate.x;
                    P.b.target_rate[row * 3 + 1] = trate.y;
                    P.b.target_rate[row * 3 + 2] = trate.z;
                    P.b.action_error[row] = action_err;
                }
            } else {
                cmd[0] = act.x; cmd[1] = act.y; cmd[2] = act.z; cmd[3] = act.w;
                action_err = P.b.action_error[row];
            }
            // ---- rotor model, rotor_group.py:55-71
            float dsq = 0.f;
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
        // ---- downwash all-pairs, multirotor.py:488-494, 724-753
        const float total_thrust = ((T[0] + T[1]) + T[2]) + T[3];
        const V3 Fw = qrot<false>(q, mk(0.f, 0.f, total_thrust));
        V3 dw = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const V3 Fj = gshfl3(Fw, gbase + j);
            const V3 pj = gshfl3(p, gbase + j);
            if (is_drone && j != slot) {
                const V3 d = Fj / (norm3(Fj) + 1e-6f);
                const V3 rel = pj - p;
                const float zd = dot3(rel, d);
                const float rr = norm3(rel - d * zd);
                const float z = zd < 0.0f ? 0.0f : zd;
                const float qq = fdiv(c.downwash_kr * rr, z);
                const float den = 1.0f + c.downwash_kz * z;
                const float v = fdiv(__expf(-0.5f * (qq * qq)), den * den);
                dw = dw + neg(Fj) * v;
            }
        }
        ext = dw + lv * c.drag_coef_times_mass;

        // ---- evader, hideandseek.py:1067-1141 + 737-744
        V3 fp = mk(0.f, 0.f, 0.f);
        if (is_drone) {
            const V3 rel = p - tp;
            const float dist = norm3(rel);
            const bool blocked = los_blocked(p, tp, cx, cy, cz, C, c.cylinder_size);
            const float active = ((dist < c.target_detect_radius) && !blocked) ? 1.0f : 0.0f;
            const float inv_d = frcp(dist + 1e-5f);
            fp = (neg(rel) * (inv_d * inv_d)) * active;
        }
        V3 force = gshfl3(fp, gbase);
#pragma unroll
        for (int j = 1; j < A; ++j) force = force + gshfl3(fp, gbase + j);
        if (is_ev) {
            force = mk(0.f, 0.f, 0.f) + force;
            const float rho = fsqrt(tp.x * tp.x + tp.y * tp.y);
            const float inv_rho = frcp(rho + 1e-5f);
            const float inx = -tp.x * inv_rho, iny = -tp.y * inv_rho;
            out_of_arena = (tp.x * tp.x + tp.y * tp.y) > c.arena_size_sq;
            const float o = out_of_arena ? 1.0f : 0.0f, no = out_of_arena ? 0.0f : 1.0f;
            const float wall = frcp((c.arena_size - rho) + 1e-5f);
            V3 fr;
            fr.x = (o * inx) * 1e5f + (no * inx) * wall;
            fr.y = (o * iny) * 1e5f + (no * iny) * wall;
            const bool hi = tp.z > c.max_height;
            const float hz = c.max_height - tp.z;
            fr.z = hi ? -1e5f : fdiv(-hz, hz * hz + 1e-5f);
            const bool lo = tp.z < 0.0f;
            const float lz = 0.0f - tp.z;
            fr.z = fr.z + (lo ? 1e5f : fdiv(-lz, lz * lz + 1e-5f));
            force = force + fr;
            float fcx = 0.f, fcy = 0.f;
#pragma unroll
            for (int k = 0; k < CT; ++k) {
                if (k < C && !(cz[k] < 0.0f)) {
                    const float tx = tp.x - cx[k], ty = tp.y - cy[k];
                    const float dxy = fsqrt(tx * tx + ty * ty);
                    if (dxy < c.target_detect_radius) {
                        const float sc = frcp(dxy + 1e-5f) * frcp((dxy - c.cylinder_size) + 1e-5f);
                        fcx = fcx + tx * sc;
                        fcy = fcy + ty * sc;
                    }
                }
            }
            force = force + mk(fcx, fcy, 0.f);
            const float vp = v_prey;
            tv = mk(fdiv(vp * force.x, fabsf(force.x) + 1e-5f), fdiv(vp * force.y, fabsf(force.y) + 1e-5f),
                    fdiv(vp * force.z, fabsf(force.z) + 1e-5f));
        }
    }

    // ---- rigid-body integration (PhysX stand-in; oracle/hs_oracle.py rigid_body_step) ----
    if (is_drone) {
        V3 force = mk(0.f, 0.f, 0.f), tau = mk(0.f, 0.f, 0.f);
        if (!RESET) {
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
        wb = wb + mk(tg.x * c.inv_inertia[0], tg.y * c.inv_inertia[1], tg.z * c.inv_inertia[2]) * dt;
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
        Q4 qn = qmul(dq, q);
        const float qinv = rsqrtf(((qn.w * qn.w + qn.x * qn.x) + qn.y * qn.y) + qn.z * qn.z);
        q.w = qn.w * qinv; q.x = qn.x * qinv; q.y = qn.y * qinv; q.z = qn.z * qinv;
        if (c.ground_clamp && p.z < c.ground_z) {
            p.z = c.ground_z;
            if (v.z < 0.0f) v.z = 0.0f;
        }
        lv = v; av = w;
    }
    if (is_ev) tp = tp + tv * dt;
    // everyone needs the evader's new position/velocity
    tp = gshfl3(tp, gbase + A);
    tv = gshfl3(tv, gbase + A);
    if (!RESET) progress = progress + 1.0f;
    else if (P.env_mask == nullptr || P.env_mask[e]) progress = 0.0f;

    // ---- write back state ------------------------------------------------------------
    if (valid && is_drone) {
        *DROW(D_POS) = p.x; *DROW(D_POS + 1) = p.y; *DROW(D_POS + 2) = p.z;
        *DROW(D_ROT) = q.w; *DROW(D_ROT + 1) = q.x; *DROW(D_ROT + 2) = q.y; *DROW(D_ROT + 3) = q.z;
        *DROW(D_LIN) = lv.x; *DROW(D_LIN + 1) = lv.y; *DROW(D_LIN + 2) = lv.z;
        *DROW(D_ANG) = av.x; *DROW(D_ANG + 1) = av.y; *DROW(D_ANG + 2) = av.z;
        if (!RESET) {
#pragma unroll
            for (int k = 0; k < 4; ++k) *DROW(D_THR + k) = thr[k];
            if (P.action_is_raw) {
                *DROW(D_INT) = integ.x; *DROW(D_INT + 1) = integ.y; *DROW(D_INT + 2) = integ.z;
                *DROW(D_LAST) = last.x; *DROW(D_LAST + 1) = last.y; *DROW(D_LAST + 2) = last.z;
            }
        }
    }
    if (valid && is_ev) {
        *EROW(E_TPOS) = tp.x; *EROW(E_TPOS + 1) = tp.y; *EROW(E_TPOS + 2) = tp.z;
        if (!RESET) { *EROW(E_TVEL) = tv.x; *EROW(E_TVEL + 1) = tv.y; *EROW(E_TVEL + 2) = tv.z; }
        *EROW(E_PROGRESS) = progress;
    }

    // ---- observation, hideandseek.py:746-917 ------------------------------------------
    const int row_l = (lane >> 2) * A + slot;            // row of this lane inside the warp tile
    const int64_t tile_row0 = e0 * A;                    // first [E*A] row of the warp
    V3 heading, up;
    heading_up(q, heading, up);

    // info.drone_state [E,A,13]
    {
        float* s = st.begin();
        if (is_drone) {
            float* r = s + row_l * 13;
            r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = q.w; r[4] = q.x; r[5] = q.y; r[6] = q.z;
            r[7] = lv.x; r[8] = lv.y; r[9] = lv.z; r[10] = av.x; r[11] = av.y; r[12] = av.z;
        }
        st.flush(P.b.drone_state + tile_row0 * 13, nenv * A * 13, full_tile);
    }
    // state_others [E,A,A-1,3] = p_a - p_j, j != a ascending; also drone-drone collisions
    float hit_drone = 0.f;
    if (A > 1) {
        float* s = st.begin();
        int o = 0;
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const V3 pj = gshfl3(p, gbase + j);
            if (is_drone && j != slot) {
                const V3 d = p - pj;
                float* r = s + row_l * ((A - 1) * 3) + o * 3;
                r[0] = d.x; r[1] = d.y; r[2] = d.z;
                hit_drone = hit_drone + ((norm3(d) < c.coll_radius_x2) ? 1.0f : 0.0f);
                ++o;
            }
        }
        st.flush(P.b.state_others + tile_row0 * ((A - 1) * 3), nenv * A * (A - 1) * 3, full_tile);
    }
    // k nearest cylinders [E,A,K,5]; lowest index wins ties
    float hit_cyl = 0.f;
    if (K > 0) {
        float* s = st.begin();
        if (is_drone) {
            float key[CT];
#pragma unroll
            for (int k = 0; k < CT; ++k)
                key[k] = (k < C) ? (norm3(mk(p.x - cx[k], p.y - cy[k], p.z - cz[k])) - c.cylinder_size) : INFINITY;
            unsigned taken = 0u;
            float* r = s + row_l * (K * 5);
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
        st.flush(P.b.obs_cylinders + tile_row0 * (K * 5), nenv * A * K * 5, full_tile);
    }
    // target visibility
    const V3 t_rpos = p - tp;
    bool blocked = false, detect = false;
    if (is_drone) {
        blocked = los_blocked(p, tp, cx, cy, cz, C, c.cylinder_size);
        detect = (norm3(t_rpos) < c.drone_detect_radius) && !blocked;
    }
    const unsigned gmask = ((1u << A) - 1u) << gbase;
    const unsigned det_ballot = __ballot_sync(FULL, detect);
    const bool bdetect = (det_ballot & gmask) != 0u;
    const float mv = c.mask_value;
    const float tfrac = fdiv(progress, (float)c.max_episode_length);

    if (c.use_tp_net) {
        // new TP frame [progress, tpos_masked3, tvel_masked3, p_0..p_{A-1}] = last row of the tile
        cp_async_wait_all();
        float* fr = tp_tile + (lane >> 2) * per_env + keep;
        if (is_drone) { fr[7 + 3 * slot] = p.x; fr[8 + 3 * slot] = p.y; fr[9 + 3 * slot] = p.z; }
        if (is_ev) {
            fr[0] = progress;
            fr[1] = bdetect ? tp.x : mv; fr[2] = bdetect ? tp.y : mv; fr[3] = bdetect ? tp.z : mv;
            fr[4] = bdetect ? tv.x : mv; fr[5] = bdetect ? tv.y : mv; fr[6] = bdetect ? tv.z : mv;
        }
        if (P.tp_init) {                                 // very first frame: every history row = this frame
            __syncwarp();
            float* row0 = tp_tile + (lane >> 2) * per_env;
            int k = slot;
            for (int i = slot; i < keep; i += G, k += G) {
                while (k >= FD) k -= FD;
                row0[i] = fr[k];
            }
        }
        {
            float* gdst = P.b.tp_input + e0 * per_env;
            const int nwords = nenv * per_env;
            const bool bulk = HS_USE_BULK_STORE && full_tile && ((nwords & 3) == 0) &&
                              ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0);
            if (bulk) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { bulk_store(gdst, tp_tile, (uint32_t)nwords * 4u); bulk_commit(); }
            } else {
                __syncwarp();
                warp_copy_slow(gdst, tp_tile, nwords, lane);
            }
        }
        if (valid && is_ev) {
            const float inv_ha = frcp(c.half_arena);
            P.b.tp_groundtruth[e * 3 + 0] = tp.x * inv_ha;
            P.b.tp_groundtruth[e * 3 + 1] = tp.y * inv_ha;
            P.b.tp_groundtruth[e * 3 + 2] = fdiv(tp.z, c.max_height) * 2.0f - 1.0f;
            P.b.tp_done[e] = (progress <= (float)(c.max_episode_length - c.future_step)) ? 1 : 0;
            *EROW(E_BDETECT) = bdetect ? 1.0f : 0.0f;
        }
    } else {
        // no predictor: the rows are complete now (width 20)
        const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
        float* s = st.begin();
        if (is_drone) write_self_row(s + row_l * 20, head_m, 0, nullptr, q, lv, heading, up, tfrac);
        st.flush(P.b.state_self + tile_row0 * 20, nenv * A * 20, full_tile);
        s = st.begin();
        if (is_drone) write_self_row(s + row_l * 20, t_rpos, 0, nullptr, q, lv, heading, up, tfrac);
        st.flush(P.b.state_drones + tile_row0 * 20, nenv * A * 20, full_tile);
    }

    if (RESET) {
        if (valid && is_ev && P.b.truncated != nullptr)
            P.b.truncated[e] = (progress > (float)c.max_episode_length) ? 1 : 0;
        st.finish();
        return;
    }

    // ---- reward / done / stats, hideandseek.py:919-1065 --------------------------------
    float r_dist = 0.f, r_speed = 0.f, r_coll = 0.f, r_smooth = 0.f, hit_wall = 0.f;
    bool seen_capture = false;
    if (is_drone) {
        const float dist = norm3(tp - p);
        r_dist = (-c.dist_reward_coef * dist) * ((dist > c.catch_radius) ? 1.0f : 0.0f);
        seen_capture = (dist < c.catch_radius) && !blocked;
        r_speed = -c.speed_coef * ((norm3(lv) > c.v_drone) ? 1.0f : 0.0f);
        r_coll = -c.collision_coef * hit_cyl;
        r_coll = r_coll + (-c.collision_coef * hit_drone);
        hit_wall = ((p.z > c.max_height) ? 1.0f : 0.0f) +
                   (((p.x * p.x + p.y * p.y) > c.arena_size_sq) ? 1.0f : 0.0f);
        r_coll = r_coll + (-c.collision_coef * hit_wall);
        r_smooth = c.smoothness_gated ? 0.0f : c.smoothness_coef * __expf(-action_err);
    }
    const bool any_capture = (__ballot_sync(FULL, seen_capture) & gmask) != 0u;
    const bool all_blocked = (__ballot_sync(FULL, blocked) & gmask) == gmask;
    const bool any_coll = (__ballot_sync(FULL, is_drone && (r_coll < 0.0f)) & gmask) != 0u;
    const float r_detect = c.detect_reward_coef * (bdetect ? 1.0f : 0.0f);
    const float r_catch = c.catch_reward_coef * (any_capture ? 1.0f : 0.0f);
    const float reward = ((((r_dist + r_detect) + r_catch) + r_coll) + r_speed) + r_smooth;
    if (valid && is_drone) P.b.reward[e * A + slot] = reward;

    // per-env means over the A pursuers (sum in agent order, then / A like torch.mean)
    // xor-butterfly over the 4 lanes of the group; non-pursuer lanes contribute the neutral
    // element, so for A=3 the sum is ((x0+x1)+(x2+0)) = the reference's left-to-right order
    const float inv_A = 1.0f / (float)A;
    auto gmean = [&](float x) {
        float s = is_drone ? x : 0.0f;
        s = s + __shfl_xor_sync(FULL, s, 1);
        s = s + __shfl_xor_sync(FULL, s, 2);
        return s * inv_A;
    };
    auto gmax = [&](float x) {
        float s = is_drone ? x : -INFINITY;
        s = fmaxf(s, __shfl_xor_sync(FULL, s, 1));
        s = fmaxf(s, __shfl_xor_sync(FULL, s, 2));
        return s;
    };
    const float m_ae = gmean(action_err), m_dist = gmean(r_dist), m_detect = gmean(r_detect),
                m_catch = gmean(r_catch), m_speed = gmean(r_speed), m_hcyl = gmean(hit_cyl),
                m_hdrone = gmean(hit_drone), m_hwall = gmean(hit_wall), m_coll = gmean(r_coll),
                m_smooth = gmean(r_smooth), m_tdiff = gmean(throttle_diff), m_reward = gmean(reward),
                x_tdiff = gmax(throttle_diff);

    if (valid && is_ev) {
        const bool done = progress >= (float)c.max_episode_length;
        P.b.done[e] = done ? 1 : 0;
        const float inv_len = done ? frcp(progress) : 1.0f;
        float* S = P.b.stats + e;
        const int64_t Es = E;
        cp_async_wait_all();
        const float* SO = stat_tile + (lane >> 2) * HS_NUM_STATS;     // values prefetched at kernel entry
#define ST(k) S[(int64_t)(k) * Es]
#define OLD(k) SO[k]
        // accumulators that are divided by the episode length on the done tick
        ST(HS_STAT_ACTION_ERROR_MEAN) = (OLD(HS_STAT_ACTION_ERROR_MEAN) + m_ae) * inv_len;
        ST(HS_STAT_ACTION_ERROR_MAX) = fmaxf(OLD(HS_STAT_ACTION_ERROR_MAX), m_ae);
        ST(HS_STAT_OUT_OF_ARENA) = ((OLD(HS_STAT_OUT_OF_ARENA) != 0.0f) || out_of_arena) ? 1.0f : 0.0f;
        ST(HS_STAT_DISTANCE_REWARD) = (OLD(HS_STAT_DISTANCE_REWARD) + m_dist) * inv_len;
        ST(HS_STAT_SUM_DETECT_STEP) = OLD(HS_STAT_SUM_DETECT_STEP) + 1.0f * (bdetect ? 1.0f : 0.0f);
        ST(HS_STAT_DETECT_REWARD) = (OLD(HS_STAT_DETECT_REWARD) + m_detect) * inv_len;
        ST(HS_STAT_BLOCKED) = OLD(HS_STAT_BLOCKED) + (all_blocked ? 1.0f : 0.0f);
        const bool capture_flag = r_catch != 0.0f;
        ST(HS_STAT_SUCCESS) = (capture_flag || (OLD(HS_STAT_SUCCESS) != 0.0f)) ? 1.0f : 0.0f;
        const float step_now = (capture_flag ? 1.0f : 0.0f) * progress +
                               (capture_flag ? 0.0f : 1.0f) * (float)c.max_episode_length;
        ST(HS_STAT_FIRST_CAPTURE_STEP) = fminf(OLD(HS_STAT_FIRST_CAPTURE_STEP), step_now);
        ST(HS_STAT_CATCH_REWARD) = (OLD(HS_STAT_CATCH_REWARD) + m_catch) * inv_len;
        ST(HS_STAT_SPEED_REWARD) = (OLD(HS_STAT_SPEED_REWARD) + m_speed) * inv_len;
        ST(HS_STAT_COLLISION_CYLINDER) = (OLD(HS_STAT_COLLISION_CYLINDER) + m_hcyl) * inv_len;
        ST(HS_STAT_COLLISION_DRONE) = (OLD(HS_STAT_COLLISION_DRONE) + m_hdrone) * inv_len;
        ST(HS_STAT_COLLISION) = (OLD(HS_STAT_COLLISION) + (any_coll ? 1.0f : 0.0f)) * inv_len;
        ST(HS_STAT_COLLISION_WALL) = (OLD(HS_STAT_COLLISION_WALL) + m_hwall) * inv_len;
        ST(HS_STAT_COLLISION_REWARD) = (OLD(HS_STAT_COLLISION_REWARD) + m_coll) * inv_len;
        if (c.write_smoothness_coef_stat) ST(HS_STAT_SMOOTHNESS_COEF) = c.smoothness_coef;
        ST(HS_STAT_SMOOTHNESS_REWARD) = (OLD(HS_STAT_SMOOTHNESS_REWARD) + m_smooth) * inv_len;
        ST(HS_STAT_SMOOTHNESS_MEAN) = (OLD(HS_STAT_SMOOTHNESS_MEAN) + m_tdiff) * inv_len;
        ST(HS_STAT_SMOOTHNESS_MAX) = fmaxf(x_tdiff, OLD(HS_STAT_SMOOTHNESS_MAX));
        ST(HS_STAT_RETURN) = OLD(HS_STAT_RETURN) + m_reward;
        // target_predicted_error is only ever divided (stays 0); distance_predicted_reward and
        // distance_threshold_L are never written (hideandseek.py:1023-1025).
#undef ST
#undef OLD
    }
    st.finish();
}

template <int A, bool RESET, int CT>
__global__ void __launch_bounds__(128, 5)
hs_tick_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(128) float stage_mem[4][2][TICK_STAGE_WORDS];
    __shared__ __align__(128) float tp_mem[4][ENVS_PER_WARP * TP_ENV_WORDS_MAX];   // TP_input tile of the warp
    __shared__ __align__(16) float stat_mem[4][TICK_STAT_WORDS];
    const int wib = threadIdx.x >> 5;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    hs_tick_body<A, RESET, CT>(P, warp_g, stage_mem[wib][0], stage_mem[wib][1], tp_mem[wib], stat_mem[wib]);
}

#undef DROW
#undef EROW
#define DROW(k) AROW((k) * A + slot)
#define EROW(k) AROW(ND * A + (k))

// =========================================================================================
// Second half with the trajectory predictor: state_self / state_drones rows (width 20+3F).
// hideandseek.py:834-887
// =========================================================================================
template <int A>
__global__ void __launch_bounds__(128)
hs_fill_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(128) float stage_mem[4][2][FILL_STAGE_WORDS];
    const hs_config& c = P.c;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slot = lane & (G - 1);
    const int64_t e0 = warp_g * ENVS_PER_WARP;
    const int E = c.num_envs;
    if (e0 >= E) return;
    const int64_t e_raw = e0 + (lane >> 2);
    const bool valid = e_raw < E;
    const int64_t e = valid ? e_raw : (E - 1);
    const bool is_drone = slot < A;
    const int nenv = (int)min((int64_t)ENVS_PER_WARP, E - e0);
    const bool full_tile = nenv == ENVS_PER_WARP;
    const int F = c.future_step, F3 = 3 * F, D = 20 + F3;

    Stager st;
    st.buf[0] = stage_mem[wib][0];
    st.buf[1] = stage_mem[wib][1];
    st.cur = 0;
    st.lane = lane;

    V3 p = mk(0, 0, 0), lv = mk(0, 0, 0);
    Q4 q; q.w = 1.f; q.x = q.y = q.z = 0.f;
    if (is_drone) {
        p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
        q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
        lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
    }
    const V3 tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
    const float progress = *EROW(E_PROGRESS);
    const bool bdetect = *EROW(E_BDETECT) != 0.0f;
    float rp[3 * FMAX];
    const float* pr = P.tp_pred + e * F3;
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
        if (f < F) {
            const float px = (__ldg(pr + 3 * f) * 0.5f) * c.arena_size;
            const float py = (__ldg(pr + 3 * f + 1) * 0.5f) * c.arena_size;
            const float pz = ((__ldg(pr + 3 * f + 2) + 1.0f) / 2.0f) * c.max_height;
            rp[3 * f] = p.x - px; rp[3 * f + 1] = p.y - py; rp[3 * f + 2] = p.z - pz;
        } else { rp[3 * f] = rp[3 * f + 1] = rp[3 * f + 2] = 0.f; }
    }
    V3 heading, up;
    heading_up(q, heading, up);
    const float tfrac = fdiv(progress, (float)c.max_episode_length);
    const V3 t_rpos = p - tp;
    const float mv = c.mask_value;
    const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
    const int row_l = (lane >> 2) * A + slot;
    const int64_t tile_row0 = e0 * A;

    float* s = st.begin();
    if (is_drone) write_self_row(s + row_l * D, head_m, F3, rp, q, lv, heading, up, tfrac);
    st.flush(P.b.state_self + tile_row0 * D, nenv * A * D, full_tile);
    s = st.begin();
    if (is_drone) write_self_row(s + row_l * D, t_rpos, F3, rp, q, lv, heading, up, tfrac);
    st.flush(P.b.state_drones + tile_row0 * D, nenv * A * D, full_tile);
    st.finish();
}


}  // namespace
