// hs_tick.cuh -- hs_tick_kernel (the fused control tick) and hs_fill_kernel (prediction-dependent rows from a caller-supplied prediction)
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"
#include "hs_stages.cuh"

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
// B / action: the buffer table and the action of THIS tick (P.b / P.action for a one-tick launch; the rollout kernel
// passes a different table every tick).  TP_IN_SMEM: tp_tile still holds the previous tick's window (rollout kernel,
// common shape only): it is shifted in place instead of being fetched from B.tp_input_prev.
// Hook: called by the whole warp once the new TP frame of its 8 envs is complete in the shared tile (the rollout kernel
// writes it straight into the predictor's MMA operand ring).
struct NoFrameHook {
    static constexpr bool ACTIVE = false;
    template <int FD> __device__ __forceinline__ void frame(const float*, int, int, int) const {}
    __device__ __forceinline__ void state(bool, bool, int, int, const V3&, const Q4&, const V3&, const V3&, float, bool) const {}
};
template <int A, bool RESET, int CT, bool TP_IN_SMEM = false, class Hook = NoFrameHook>
__device__ __forceinline__ void hs_tick_body(const KParams& P, const hs_buffers& B, const float* __restrict__ action,
                                             const int64_t warp_g, float* stage0, float* stage1, float* tp_tile, float* stat_tile,
                                             const Hook hook = Hook()) {
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
    const uint32_t Ep32 = 32u;                                         // row pitch inside a tile (hs_common.cuh AROW)
    const uint32_t tile_off = (uint32_t)(e >> 5) * ((uint32_t)P.R * 32u) + ((uint32_t)e & 31u);
    const uint32_t o_drone = (uint32_t)slot * Ep32 + tile_off;         // + k * (A*Ep32)
    const uint32_t o_env = (uint32_t)(ND * A) * Ep32 + tile_off;       // + k * Ep32
    float* const arena = B.arena;
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
    if (TP_IN_SMEM && c.use_tp_net && (FD & 3) == 0) {
        // rows 1..H-1 of the 8 windows move to rows 0..H-2 inside the tile: all loads, then all stores
        constexpr int fd4 = FD / 4, pe4 = 10 * fd4, keep4 = 9 * fd4, NIT = (ENVS_PER_WARP * keep4 + 31) / 32;
        float4 v[NIT];
        float4* t4 = reinterpret_cast<float4*>(tp_tile);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = it * 32 + lane;
            const int env = i / keep4, j = i - env * keep4;
            if (i < ENVS_PER_WARP * keep4) v[it] = t4[env * pe4 + j + fd4];
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = it * 32 + lane;
            const int env = i / keep4, j = i - env * keep4;
            if (i < ENVS_PER_WARP * keep4) t4[env * pe4 + j] = v[it];
        }
        __syncwarp();
    } else if (c.use_tp_net && !P.tp_init) {
        const float* src = B.tp_input_prev + e0 * per_env;
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
        for (int k = 0; k < HS_NUM_STATS; ++k) cp_async4(stat_tile + (lane >> 2) * HS_NUM_STATS + k, B.stats + (int64_t)k * E + e);
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
            act4 = __ldg(reinterpret_cast<const float4*>(action) + row);
            if (P.action_is_raw) {
                prev4 = *(reinterpret_cast<const float4*>(B.prev_action) + row);
                pid_reset = (P.reset_pid != nullptr) && (P.reset_pid[e] != 0);
            }
        }
        if (is_ev) v_prey = __ldg(B.v_prey);
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
            if (P.action_is_raw) {
                CtbrOut o;
                stage_ctbr_pid(c, act4, prev4, pid_reset, q, av, integ, last, o);
                action_err = o.action_err;
#pragma unroll
                for (int k = 0; k < 4; ++k) cmd[k] = o.cmd[k];
                if (valid) {
                    *(reinterpret_cast<float4*>(B.prev_action) + row) = o.prev_new;
                    *(reinterpret_cast<float4*>(B.rotor_cmds) + row) = make_float4(cmd[0], cmd[1], cmd[2], cmd[3]);
                    *(reinterpret_cast<float4*>(B.ctbr) + row) = o.ctbr;
                    B.target_rate[row * 3 + 0] = o.trate.x;
                    B.target_rate[row * 3 + 1] = o.trate.y;
                    B.target_rate[row * 3 + 2] = o.trate.z;
                    B.action_error[row] = action_err;
                }
            } else {
                cmd[0] = act4.x; cmd[1] = act4.y; cmd[2] = act4.z; cmd[3] = act4.w;
                action_err = B.action_error[row];
            }
            // ---- rotor model, rotor_group.py:55-71
            stage_rotor(c, cmd, thr, T, yaw_torque, throttle_diff);
            if (B.throttle_diff != nullptr && valid) B.throttle_diff[row] = throttle_diff;
        }
        // ---- downwash all-pairs, multirotor.py:488-494, 724-753
        const float total_thrust = ((T[0] + T[1]) + T[2]) + T[3];
        const V3 Fw = qrot<false>(q, mk(0.f, 0.f, total_thrust));
        V3 dw = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const V3 Fj = gshfl3(Fw, gbase + j);
            const V3 pj = gshfl3(p, gbase + j);
            if (is_drone && j != slot) dw = dw + downwash_term(c, Fj, pj, p);
        }
        ext = dw + lv * c.drag_coef_times_mass;

        // ---- evader, hideandseek.py:1067-1141 + 737-744
        V3 fp = mk(0.f, 0.f, 0.f);
        if (is_drone) fp = evader_pursuer_term(c, p, tp, los_blocked(p, tp, cx, cy, cz, C, c.cylinder_size));
        V3 force = gshfl3(fp, gbase);
#pragma unroll
        for (int j = 1; j < A; ++j) force = force + gshfl3(fp, gbase + j);
        if (is_ev) tv = evader_velocity(c, force, tp, cx, cy, cz, C, v_prey, out_of_arena);
    }

    // ---- rigid-body integration (PhysX stand-in; oracle/hs_oracle.py rigid_body_step) ----
    if (is_drone) {
        stage_integrate<!RESET>(c, p, q, lv, av, T, yaw_torque, ext);
        stage_contacts(c, p, lv, tp, cx, cy, cz, C);
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
        st.flush(B.drone_state + tile_row0 * 13, nenv * A * 13, full_tile);
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
        st.flush(B.state_others + tile_row0 * ((A - 1) * 3), nenv * A * (A - 1) * 3, full_tile);
    }
    // k nearest cylinders [E,A,K,5]; lowest index wins ties
    float hit_cyl = 0.f;
    if (K > 0) {
        float* s = st.begin();
        if (is_drone) stage_knearest(c, p, cx, cy, cz, C, K, s + row_l * (K * 5), hit_cyl);
        st.flush(B.obs_cylinders + tile_row0 * (K * 5), nenv * A * K * 5, full_tile);
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
        if (Hook::ACTIVE) {
            __syncwarp();
            hook.template frame<FD>(tp_tile, per_env, keep, lane);
            hook.state(is_drone, is_ev, slot, lane >> 2, p, q, lv, tp, progress, bdetect);     // what the row kernels read back
        }
        {
            float* gdst = B.tp_input + e0 * per_env;
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
            B.tp_groundtruth[e * 3 + 0] = fdiv(tp.x, c.half_arena);
            B.tp_groundtruth[e * 3 + 1] = fdiv(tp.y, c.half_arena);
            B.tp_groundtruth[e * 3 + 2] = fdiv(tp.z, c.max_height) * 2.0f - 1.0f;
            B.tp_done[e] = (progress <= (float)(c.max_episode_length - c.future_step)) ? 1 : 0;
            *EROW(E_BDETECT) = bdetect ? 1.0f : 0.0f;
        }
    } else {
        // no predictor: the rows are complete now (width 20)
        const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
        float* s = st.begin();
        if (is_drone) write_self_row(s + row_l * 20, head_m, 0, nullptr, q, lv, heading, up, tfrac);
        st.flush(B.state_self + tile_row0 * 20, nenv * A * 20, full_tile);
        s = st.begin();
        if (is_drone) write_self_row(s + row_l * 20, t_rpos, 0, nullptr, q, lv, heading, up, tfrac);
        st.flush(B.state_drones + tile_row0 * 20, nenv * A * 20, full_tile);
    }

    if (RESET) {
        if (valid && is_ev && B.truncated != nullptr)
            B.truncated[e] = (progress > (float)c.max_episode_length) ? 1 : 0;
        st.finish();
        return;
    }

    // ---- reward / done / stats, hideandseek.py:919-1065 --------------------------------
    float r_dist = 0.f, r_speed = 0.f, r_coll = 0.f, r_smooth = 0.f, hit_wall = 0.f;
    bool seen_capture = false;
    // device scalar when bound: follows update_epoch without re-capturing graphs (hideandseek.py:988-991)
    const float sm_coef = (B.smoothness_coef != nullptr) ? __ldg(B.smoothness_coef) : c.smoothness_coef;
    if (is_drone) {
        const RewardTerms rt = stage_reward_terms(c, p, lv, tp, blocked, hit_cyl, hit_drone, action_err, sm_coef);
        r_dist = rt.r_dist; r_speed = rt.r_speed; r_coll = rt.r_coll; r_smooth = rt.r_smooth; hit_wall = rt.hit_wall;
        seen_capture = rt.seen_capture;
    }
    const bool any_capture = (__ballot_sync(FULL, seen_capture) & gmask) != 0u;
    const bool all_blocked = (__ballot_sync(FULL, blocked) & gmask) == gmask;
    const bool any_coll = (__ballot_sync(FULL, is_drone && (r_coll < 0.0f)) & gmask) != 0u;
    const float r_detect = c.detect_reward_coef * (bdetect ? 1.0f : 0.0f);
    const float r_catch = c.catch_reward_coef * (any_capture ? 1.0f : 0.0f);
    const float reward = ((((r_dist + r_detect) + r_catch) + r_coll) + r_speed) + r_smooth;
    if (valid && is_drone) B.reward[e * A + slot] = reward;

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
    EnvTick et;
    et.m_ae = gmean(action_err); et.m_dist = gmean(r_dist); et.m_detect = gmean(r_detect); et.m_catch = gmean(r_catch);
    et.m_speed = gmean(r_speed); et.m_hcyl = gmean(hit_cyl); et.m_hdrone = gmean(hit_drone); et.m_hwall = gmean(hit_wall);
    et.m_coll = gmean(r_coll); et.m_smooth = gmean(r_smooth); et.m_tdiff = gmean(throttle_diff); et.m_reward = gmean(reward);
    et.x_tdiff = gmax(throttle_diff);
    et.r_catch = r_catch; et.bdetect = bdetect; et.all_blocked = all_blocked; et.any_coll = any_coll; et.out_of_arena = out_of_arena;

    if (valid && is_ev) {
        B.done[e] = (progress >= (float)c.max_episode_length) ? 1 : 0;
        float* S = B.stats + e;
        const int64_t Es = E;
        cp_async_wait_all();
        const float* SO = stat_tile + (lane >> 2) * HS_NUM_STATS;     // values prefetched at kernel entry
        stage_stats(c, et, progress, sm_coef, [&](int k) { return SO[k]; }, [&](int k, float v) { S[(int64_t)k * Es] = v; });
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
    hs_tick_body<A, RESET, CT>(P, P.b, P.action, warp_g, stage_mem[wib][0], stage_mem[wib][1], tp_mem[wib], stat_mem[wib]);
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
