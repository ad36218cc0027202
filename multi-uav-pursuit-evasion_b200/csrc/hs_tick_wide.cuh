// hs_tick_wide.cuh -- the control tick with ONE LANE PER ENVIRONMENT (hs_tick_wide_kernel) and the matching row-fill kernel.
//
// The bandwidth-bound mapping for batches that fill the machine, and the only one for more than 3 pursuers (the 4-lane
// mapping of hs_tick.cuh fixes A <= 3).  A warp owns a tile of 32 consecutive envs:
//   * the tile's SoA state - rows [23 A + 8 + 3 C] x 32 envs, 128 B per row - arrives with ONE TMA tensor copy
//     (cp.async.bulk.tensor.2d, SASS UTMALDG) into shared memory, the 24 stats rows with a second one, the AoS action and
//     prev_action spans with two bulk copies; one mbarrier (complete_tx) per warp, no block barrier anywhere;
//   * every lane then advances its env with the pursuers as an unrolled loop (A independent dependency chains per lane
//     instead of 3 of 4 lanes busy), reading state as tile[row][lane] - conflict free - and updating it in place;
//   * the updated state / stats tiles leave with TMA tensor stores (UTMASTG), every AoS output ([32 envs][W words], one
//     contiguous span per tensor) is staged in shared memory and leaves with one bulk store (UBLKCP);
//   * the previous TP window is shifted by one frame global -> global with warp-coalesced 16 B copies (it never enters
//     shared memory), the new frame is written directly (64 B per env, sector exact).
// Arithmetic = the device functions of hs_stages.cuh in the same order as hs_tick_body: results are bit-identical to
// the 4-lane kernel (tests/test_gpu_wide.py).
#pragma once
#include <cuda.h>                    // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "hs_common.cuh"
#include "hs_stages.cuh"

namespace {

constexpr int WIDE_WARPS = 2;        // warps (= 32-env tiles) per CTA; warps never synchronise with each other
constexpr int WIDE_MAX_A = HS_MAX_AGENTS;

// shared-memory plan of one warp (offsets in floats, every region 128 B aligned)
struct WidePlan {
    int rows_all, rows_rw;           // state tile rows: all (loaded), read-write prefix (stored back)
    int o_state, o_stats, o_action, o_prev;
    int o_dstate, o_others, o_cyl, o_cmds, o_ctbr, o_trate, o_aerr, o_reward, o_gt, o_self, o_drones;
    int total;                       // floats per warp (the mbarrier lives in the first 32 floats)
};
__host__ __device__ inline int wide_up(int words) { return (words + 31) & ~31; }
__host__ __device__ inline WidePlan wide_plan(int A, int C, int K, bool tp) {
    WidePlan w;
    w.rows_rw = ND * A + E_CYL;
    w.rows_all = w.rows_rw + 3 * C;
    int o = 32;                                                   // [0, 32): mbarrier
    w.o_state = o;  o += w.rows_all * 32;
    w.o_stats = o;  o += HS_NUM_STATS * 32;
    w.o_action = o; o += wide_up(32 * A * 4);
    w.o_prev = o;   o += wide_up(32 * A * 4);
    w.o_dstate = o; o += wide_up(32 * A * 13);
    w.o_others = o; o += wide_up(32 * A * (A - 1) * 3);
    w.o_cyl = o;    o += wide_up(32 * A * K * 5);
    w.o_cmds = o;   o += wide_up(32 * A * 4);
    w.o_ctbr = o;   o += wide_up(32 * A * 4);
    w.o_trate = o;  o += wide_up(32 * A * 3);
    w.o_aerr = o;   o += wide_up(32 * A);
    w.o_reward = o; o += wide_up(32 * A);
    w.o_gt = o;     o += tp ? wide_up(32 * 3) : 0;
    w.o_self = o;   o += tp ? 0 : wide_up(32 * A * 20);
    w.o_drones = o; o += tp ? 0 : wide_up(32 * A * 20);
    w.total = (o + 255) & ~255;                                   // 1 KB granularity keeps every warp's base aligned
    return w;
}

__device__ __forceinline__ uint32_t wd_smem(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void wd_mbar_init(uint32_t bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void wd_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wd_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void wd_tma_load_2d(uint32_t sdst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(sdst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void wd_tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t ssrc) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(ssrc), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void wd_bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

// One AoS output of the tile: [nenv][W] words staged at `s`, destination `g` (start of the tile's span).
// Full tiles leave with one bulk store issued by lane 0 (caller fences and commits); ragged tiles with a plain copy.
__device__ __forceinline__ void wd_store_span(float* g, const float* s, int W, int nenv, bool full, int lane) {
    if (W == 0 || g == nullptr) return;
    if (full && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
        if (lane == 0) bulk_store(g, s, (uint32_t)(32 * W) * 4u);
    } else {
        warp_copy_slow(g, s, nenv * W, lane);
    }
}

// Previous TP window -> this tick's window shifted by one frame, global -> global, warp-coalesced:
// dst[env][i] = src[env][FD + i], i < keep = (H-1) FD, for the nenv envs of the tile.
template <int VEC>
__device__ __forceinline__ void wd_window_shift(float* dst, const float* src, int nenv, int per_env, int keep, int FD, int lane) {
    constexpr int U = 12;
    const int keepv = keep / VEC;
    const int total = nenv * keepv;
    for (int base = 0; base < total; base += 32 * U) {
        float4 r4[U];
        float r1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * 32 + lane;
            if (i < total) {
                const int env = i / keepv, j = i - env * keepv;
                if (VEC == 4) r4[u] = __ldg(reinterpret_cast<const float4*>(src + env * per_env + FD) + j);
                else r1[u] = __ldg(src + env * per_env + FD + j);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * 32 + lane;
            if (i < total) {
                const int env = i / keepv, j = i - env * keepv;
                if (VEC == 4) *(reinterpret_cast<float4*>(dst + env * per_env) + j) = r4[u];
                else dst[env * per_env + j] = r1[u];
            }
        }
    }
}

template <int A, int CT, bool RESET>
__global__ void __launch_bounds__(WIDE_WARPS * 32)
hs_tick_wide_kernel(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tm_state_ld,
                    const __grid_constant__ CUtensorMap tm_state_st, const __grid_constant__ CUtensorMap tm_stats) {
    extern __shared__ __align__(1024) float wide_mem[];
    const hs_config& c = P.c;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int E = c.num_envs;
    const int64_t e0 = ((int64_t)blockIdx.x * WIDE_WARPS + wib) * 32;
    if (e0 >= E) return;                                          // whole warp out of range (no block barrier below)
    const int C = c.num_cylinders, K = c.obs_max_cylinder, H = c.history_step;
    const int FD = 7 + 3 * A;
    const bool tp_on = c.use_tp_net != 0;
    const WidePlan w = wide_plan(A, C, K, tp_on);
    float* const mem = wide_mem + (size_t)wib * w.total;
    const int nenv = (int)min((int64_t)32, E - e0);
    const bool full = nenv == 32;
    const int64_t e = e0 + lane;
    const bool valid = lane < nenv;
    const int64_t ec = valid ? e : (int64_t)E - 1;               // clamped index for the few direct global reads
    const float dt = c.dt;
    const bool raw = P.action_is_raw != 0;

    float* const S = mem + w.o_state;                            // S[row * 32 + lane]
    float* const ST = mem + w.o_stats;
    float4* const s_act = reinterpret_cast<float4*>(mem + w.o_action);
    float4* const s_prev = reinterpret_cast<float4*>(mem + w.o_prev);
    const uint32_t bar = wd_smem(mem);

    // ---- loads: one mbarrier per warp, lane 0 issues ------------------------------------------
    if (lane == 0) {
        wd_mbar_init(bar);
        uint32_t bytes = (uint32_t)w.rows_all * 128u;
        if (!RESET) {
            bytes += HS_NUM_STATS * 128u + (uint32_t)nenv * A * 16u;
            if (raw) bytes += (uint32_t)nenv * A * 16u;
        }
        wd_mbar_expect(bar, bytes);
        wd_tma_load_2d(wd_smem(S), &tm_state_ld, (int)e0, 0, bar);
        if (!RESET) {
            wd_tma_load_2d(wd_smem(ST), &tm_stats, (int)e0, 0, bar);
            wd_bulk_load(wd_smem(s_act), P.action + e0 * A * 4, (uint32_t)nenv * A * 16u, bar);
            if (raw) wd_bulk_load(wd_smem(s_prev), P.b.prev_action + e0 * A * 4, (uint32_t)nenv * A * 16u, bar);
        }
    }
    __syncwarp();
    // ---- the TP window moves global -> global while the tile is in flight -----------------------
    const int per_env = H * FD, keep = (H - 1) * FD;
    if (tp_on && !P.tp_init && keep > 0) {
        const float* src = P.b.tp_input_prev + e0 * per_env;
        float* dst = P.b.tp_input + e0 * per_env;
        if ((FD & 3) == 0) wd_window_shift<4>(dst, src, nenv, per_env, keep, FD, lane);
        else wd_window_shift<1>(dst, src, nenv, per_env, keep, FD, lane);
    }
    bool pid_reset = false;
    float v_prey = 0.f;
    if (!RESET) {
        pid_reset = raw && (P.reset_pid != nullptr) && (P.reset_pid[ec] != 0);
        v_prey = __ldg(P.b.v_prey);
    }
    wd_mbar_wait(bar, 0);

#define SD(k, a) S[((k) * A + (a)) * 32 + lane]
#define SE(k) S[(ND * A + (k)) * 32 + lane]
    V3 tp = mk(SE(E_TPOS), SE(E_TPOS + 1), SE(E_TPOS + 2));
    V3 tv = mk(SE(E_TVEL), SE(E_TVEL + 1), SE(E_TVEL + 2));
    float progress = SE(E_PROGRESS);
    float cx[CT], cy[CT], cz[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        if (k < C) { cx[k] = SE(E_CYL + 3 * k); cy[k] = SE(E_CYL + 3 * k + 1); cz[k] = SE(E_CYL + 3 * k + 2); }
        else { cx[k] = 0.f; cy[k] = 0.f; cz[k] = -20.f; }
    }

    V3 pos[A];                                                    // positions: old until the integration, new afterwards
    float Tt[A][4], yaw[A], aerr[A], tdiff[A];
    V3 ext[A];
    bool out_of_arena = false;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        pos[a] = mk(SD(D_POS, a), SD(D_POS + 1, a), SD(D_POS + 2, a));
        Tt[a][0] = Tt[a][1] = Tt[a][2] = Tt[a][3] = 0.f;
        yaw[a] = 0.f; aerr[a] = 0.f; tdiff[a] = 0.f;
        ext[a] = mk(0.f, 0.f, 0.f);
    }

    if (!RESET) {
        // ---- CTBR transform + body-rate PID + rotor model, per pursuer --------------------------
        V3 Fw[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            Q4 q; q.w = SD(D_ROT, a); q.x = SD(D_ROT + 1, a); q.y = SD(D_ROT + 2, a); q.z = SD(D_ROT + 3, a);
            float thr[4] = {SD(D_THR, a), SD(D_THR + 1, a), SD(D_THR + 2, a), SD(D_THR + 3, a)};
            const float4 act = s_act[lane * A + a];
            float cmd[4];
            if (raw) {
                const V3 av = mk(SD(D_ANG, a), SD(D_ANG + 1, a), SD(D_ANG + 2, a));
                V3 integ = mk(SD(D_INT, a), SD(D_INT + 1, a), SD(D_INT + 2, a));
                V3 last = mk(SD(D_LAST, a), SD(D_LAST + 1, a), SD(D_LAST + 2, a));
                CtbrOut o;
                stage_ctbr_pid(c, act, s_prev[lane * A + a], pid_reset, q, av, integ, last, o);
                aerr[a] = o.action_err;
#pragma unroll
                for (int k = 0; k < 4; ++k) cmd[k] = o.cmd[k];
                s_prev[lane * A + a] = o.prev_new;
                reinterpret_cast<float4*>(mem + w.o_cmds)[lane * A + a] = make_float4(cmd[0], cmd[1], cmd[2], cmd[3]);
                reinterpret_cast<float4*>(mem + w.o_ctbr)[lane * A + a] = o.ctbr;
                float* tr = mem + w.o_trate + (lane * A + a) * 3;
                tr[0] = o.trate.x; tr[1] = o.trate.y; tr[2] = o.trate.z;
                mem[w.o_aerr + lane * A + a] = o.action_err;
                if (valid) {
                    SD(D_INT, a) = integ.x; SD(D_INT + 1, a) = integ.y; SD(D_INT + 2, a) = integ.z;
                    SD(D_LAST, a) = last.x; SD(D_LAST + 1, a) = last.y; SD(D_LAST + 2, a) = last.z;
                }
            } else {
                cmd[0] = act.x; cmd[1] = act.y; cmd[2] = act.z; cmd[3] = act.w;
                aerr[a] = P.b.action_error[ec * A + a];
            }
            stage_rotor(c, cmd, thr, Tt[a], yaw[a], tdiff[a]);
            if (valid) { SD(D_THR, a) = thr[0]; SD(D_THR + 1, a) = thr[1]; SD(D_THR + 2, a) = thr[2]; SD(D_THR + 3, a) = thr[3]; }
            const float total_thrust = ((Tt[a][0] + Tt[a][1]) + Tt[a][2]) + Tt[a][3];
            Fw[a] = qrot<false>(q, mk(0.f, 0.f, total_thrust));
        }
        // ---- downwash all-pairs and the evader's repulsion sum (agent order) --------------------
        V3 force = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int a = 0; a < A; ++a) {
            V3 dw = mk(0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < A; ++j)
                if (j != a) dw = dw + downwash_term(c, Fw[j], pos[j], pos[a]);
            const V3 lv = mk(SD(D_LIN, a), SD(D_LIN + 1, a), SD(D_LIN + 2, a));
            ext[a] = dw + lv * c.drag_coef_times_mass;
            const V3 fp = evader_pursuer_term(c, pos[a], tp, los_blocked(pos[a], tp, cx, cy, cz, C, c.cylinder_size));
            force = (a == 0) ? fp : force + fp;
        }
        tv = evader_velocity(c, force, tp, cx, cy, cz, C, v_prey, out_of_arena);
    }

    // ---- rigid-body integration ------------------------------------------------------------------
#pragma unroll
    for (int a = 0; a < A; ++a) {
        Q4 q; q.w = SD(D_ROT, a); q.x = SD(D_ROT + 1, a); q.y = SD(D_ROT + 2, a); q.z = SD(D_ROT + 3, a);
        V3 lv = mk(SD(D_LIN, a), SD(D_LIN + 1, a), SD(D_LIN + 2, a));
        V3 av = mk(SD(D_ANG, a), SD(D_ANG + 1, a), SD(D_ANG + 2, a));
        stage_integrate<!RESET>(c, pos[a], q, lv, av, Tt[a], yaw[a], ext[a]);
        if (valid) {
            SD(D_POS, a) = pos[a].x; SD(D_POS + 1, a) = pos[a].y; SD(D_POS + 2, a) = pos[a].z;
            SD(D_ROT, a) = q.w; SD(D_ROT + 1, a) = q.x; SD(D_ROT + 2, a) = q.y; SD(D_ROT + 3, a) = q.z;
            SD(D_LIN, a) = lv.x; SD(D_LIN + 1, a) = lv.y; SD(D_LIN + 2, a) = lv.z;
            SD(D_ANG, a) = av.x; SD(D_ANG + 1, a) = av.y; SD(D_ANG + 2, a) = av.z;
        }
        // info.drone_state [E,A,13]
        float* r = mem + w.o_dstate + (lane * A + a) * 13;
        r[0] = pos[a].x; r[1] = pos[a].y; r[2] = pos[a].z; r[3] = q.w; r[4] = q.x; r[5] = q.y; r[6] = q.z;
        r[7] = lv.x; r[8] = lv.y; r[9] = lv.z; r[10] = av.x; r[11] = av.y; r[12] = av.z;
    }
    tp = tp + tv * dt;
    if (!RESET) progress = progress + 1.0f;
    else if (P.env_mask == nullptr || P.env_mask[ec]) progress = 0.0f;
    if (valid) {
        SE(E_TPOS) = tp.x; SE(E_TPOS + 1) = tp.y; SE(E_TPOS + 2) = tp.z;
        if (!RESET) { SE(E_TVEL) = tv.x; SE(E_TVEL + 1) = tv.y; SE(E_TVEL + 2) = tv.z; }
        SE(E_PROGRESS) = progress;
    }

    // ---- observation, hideandseek.py:746-917 -----------------------------------------------------
    float hit_drone[A], hit_cyl[A];
    bool blocked[A];
    bool bdetect = false;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        hit_drone[a] = 0.f; hit_cyl[a] = 0.f;
        if (A > 1) {
            float* r = mem + w.o_others + (lane * A + a) * ((A - 1) * 3);
            int o = 0;
#pragma unroll
            for (int j = 0; j < A; ++j) {
                if (j != a) {
                    const V3 d = pos[a] - pos[j];
                    r[o * 3] = d.x; r[o * 3 + 1] = d.y; r[o * 3 + 2] = d.z;
                    hit_drone[a] = hit_drone[a] + ((norm3(d) < c.coll_radius_x2) ? 1.0f : 0.0f);
                    ++o;
                }
            }
        }
        if (K > 0) stage_knearest(c, pos[a], cx, cy, cz, C, K, mem + w.o_cyl + (lane * A + a) * (K * 5), hit_cyl[a]);
        blocked[a] = los_blocked(pos[a], tp, cx, cy, cz, C, c.cylinder_size);
        const bool detect = (norm3(pos[a] - tp) < c.drone_detect_radius) && !blocked[a];
        bdetect = bdetect || detect;
    }
    const float mv = c.mask_value;
    const float tfrac = fdiv(progress, (float)c.max_episode_length);
    if (tp_on) {
        // new TP frame [progress, tpos_masked3, tvel_masked3, p_0..p_{A-1}], written straight to its slot (row H-1;
        // every row on the very first frame)
        float fr[7 + 3 * A];
        fr[0] = progress;
        fr[1] = bdetect ? tp.x : mv; fr[2] = bdetect ? tp.y : mv; fr[3] = bdetect ? tp.z : mv;
        fr[4] = bdetect ? tv.x : mv; fr[5] = bdetect ? tv.y : mv; fr[6] = bdetect ? tv.z : mv;
#pragma unroll
        for (int a = 0; a < A; ++a) { fr[7 + 3 * a] = pos[a].x; fr[8 + 3 * a] = pos[a].y; fr[9 + 3 * a] = pos[a].z; }
        if (valid) {
            float* win = P.b.tp_input + e * per_env;
            const int h0 = P.tp_init ? 0 : H - 1;
            for (int h = h0; h < H; ++h) {
                float* row = win + h * FD;
                if (((7 + 3 * A) & 3) == 0) {
#pragma unroll
                    for (int k = 0; k < (7 + 3 * A) / 4; ++k)
                        reinterpret_cast<float4*>(row)[k] = make_float4(fr[4 * k], fr[4 * k + 1], fr[4 * k + 2], fr[4 * k + 3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 7 + 3 * A; ++k) row[k] = fr[k];
                }
            }
            P.b.tp_done[e] = (progress <= (float)(c.max_episode_length - c.future_step)) ? 1 : 0;
            SE(E_BDETECT) = bdetect ? 1.0f : 0.0f;
        }
        float* gt = mem + w.o_gt + lane * 3;
        gt[0] = fdiv(tp.x, c.half_arena);
        gt[1] = fdiv(tp.y, c.half_arena);
        gt[2] = fdiv(tp.z, c.max_height) * 2.0f - 1.0f;
    } else {
        // no predictor: the rows are complete now (width 20)
#pragma unroll
        for (int a = 0; a < A; ++a) {
            Q4 q; q.w = SD(D_ROT, a); q.x = SD(D_ROT + 1, a); q.y = SD(D_ROT + 2, a); q.z = SD(D_ROT + 3, a);
            const V3 lv = mk(SD(D_LIN, a), SD(D_LIN + 1, a), SD(D_LIN + 2, a));
            V3 heading, up;
            heading_up(q, heading, up);
            const V3 t_rpos = pos[a] - tp;
            const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
            write_self_row(mem + w.o_self + (lane * A + a) * 20, head_m, 0, nullptr, q, lv, heading, up, tfrac);
            write_self_row(mem + w.o_drones + (lane * A + a) * 20, t_rpos, 0, nullptr, q, lv, heading, up, tfrac);
        }
    }

    if (RESET) {
        if (valid && P.b.truncated != nullptr) P.b.truncated[e] = (progress > (float)c.max_episode_length) ? 1 : 0;
    } else {
        // ---- reward / done / stats, hideandseek.py:919-1065 ---------------------------------------
        const float sm_coef = (P.b.smoothness_coef != nullptr) ? __ldg(P.b.smoothness_coef) : c.smoothness_coef;
        RewardTerms rt[A];
        bool any_capture = false, all_blocked = true, any_coll = false;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const V3 lv = mk(SD(D_LIN, a), SD(D_LIN + 1, a), SD(D_LIN + 2, a));
            rt[a] = stage_reward_terms(c, pos[a], lv, tp, blocked[a], hit_cyl[a], hit_drone[a], aerr[a], sm_coef);
            any_capture = any_capture || rt[a].seen_capture;
            all_blocked = all_blocked && blocked[a];
            any_coll = any_coll || (rt[a].r_coll < 0.0f);
        }
        const float r_detect = c.detect_reward_coef * (bdetect ? 1.0f : 0.0f);
        const float r_catch = c.catch_reward_coef * (any_capture ? 1.0f : 0.0f);
        const float inv_A = 1.0f / (float)A;
        EnvTick et;
        float s_ae = 0.f, s_dist = 0.f, s_speed = 0.f, s_hcyl = 0.f, s_hdrone = 0.f, s_hwall = 0.f, s_coll = 0.f, s_smooth = 0.f,
              s_tdiff = 0.f, s_reward = 0.f, s_detect = 0.f, s_catch = 0.f, x_tdiff = -INFINITY;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float reward = ((((rt[a].r_dist + r_detect) + r_catch) + rt[a].r_coll) + rt[a].r_speed) + rt[a].r_smooth;
            mem[w.o_reward + lane * A + a] = reward;
            // sums in agent order (for A <= 3 the same association as the 4-lane butterfly: (x0 + x1) + x2)
            if (a == 0) {
                s_ae = aerr[a]; s_dist = rt[a].r_dist; s_speed = rt[a].r_speed; s_hcyl = hit_cyl[a]; s_hdrone = hit_drone[a];
                s_hwall = rt[a].hit_wall; s_coll = rt[a].r_coll; s_smooth = rt[a].r_smooth; s_tdiff = tdiff[a]; s_reward = reward;
                s_detect = r_detect; s_catch = r_catch;
            } else {
                s_ae += aerr[a]; s_dist += rt[a].r_dist; s_speed += rt[a].r_speed; s_hcyl += hit_cyl[a]; s_hdrone += hit_drone[a];
                s_hwall += rt[a].hit_wall; s_coll += rt[a].r_coll; s_smooth += rt[a].r_smooth; s_tdiff += tdiff[a]; s_reward += reward;
                s_detect += r_detect; s_catch += r_catch;
            }
            x_tdiff = fmaxf(x_tdiff, tdiff[a]);
        }
        et.m_ae = s_ae * inv_A; et.m_dist = s_dist * inv_A; et.m_detect = s_detect * inv_A; et.m_catch = s_catch * inv_A;
        et.m_speed = s_speed * inv_A; et.m_hcyl = s_hcyl * inv_A; et.m_hdrone = s_hdrone * inv_A; et.m_hwall = s_hwall * inv_A;
        et.m_coll = s_coll * inv_A; et.m_smooth = s_smooth * inv_A; et.m_tdiff = s_tdiff * inv_A; et.m_reward = s_reward * inv_A;
        et.x_tdiff = x_tdiff;
        et.r_catch = r_catch; et.bdetect = bdetect; et.all_blocked = all_blocked; et.any_coll = any_coll; et.out_of_arena = out_of_arena;
        if (valid) {
            P.b.done[e] = (progress >= (float)c.max_episode_length) ? 1 : 0;
            stage_stats(c, et, progress, sm_coef, [&](int k) { return ST[k * 32 + lane]; },
                        [&](int k, float v) { ST[k * 32 + lane] = v; });
        }
    }
#undef SD
#undef SE

    // ---- stores: state / stats tiles by TMA tensor store, AoS outputs by bulk store -----------------
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
        wd_tma_store_2d(&tm_state_st, (int)e0, 0, wd_smem(S));
        if (!RESET) wd_tma_store_2d(&tm_stats, (int)e0, 0, wd_smem(ST));
    }
    const int64_t r0 = e0 * A;
    wd_store_span(P.b.drone_state + r0 * 13, mem + w.o_dstate, A * 13, nenv, full, lane);
    if (A > 1) wd_store_span(P.b.state_others + r0 * ((A - 1) * 3), mem + w.o_others, A * (A - 1) * 3, nenv, full, lane);
    if (K > 0) wd_store_span(P.b.obs_cylinders + r0 * (K * 5), mem + w.o_cyl, A * K * 5, nenv, full, lane);
    if (tp_on) {
        wd_store_span(P.b.tp_groundtruth + e0 * 3, mem + w.o_gt, 3, nenv, full, lane);
    } else {
        wd_store_span(P.b.state_self + r0 * 20, mem + w.o_self, A * 20, nenv, full, lane);
        wd_store_span(P.b.state_drones + r0 * 20, mem + w.o_drones, A * 20, nenv, full, lane);
    }
    if (!RESET) {
        wd_store_span(P.b.reward + r0, mem + w.o_reward, A, nenv, full, lane);
        if (raw) {
            wd_store_span(P.b.prev_action + r0 * 4, mem + w.o_prev, A * 4, nenv, full, lane);
            wd_store_span(P.b.rotor_cmds + r0 * 4, mem + w.o_cmds, A * 4, nenv, full, lane);
            wd_store_span(P.b.ctbr + r0 * 4, mem + w.o_ctbr, A * 4, nenv, full, lane);
            wd_store_span(P.b.target_rate + r0 * 3, mem + w.o_trate, A * 3, nenv, full, lane);
            wd_store_span(P.b.action_error + r0, mem + w.o_aerr, A, nenv, full, lane);
        }
    }
    if (lane == 0) {
        bulk_commit();
        bulk_wait_read<0>();                                      // shared memory must outlive the reads of the bulk engine
    }
    __syncwarp();
}

// =========================================================================================
// Second half with a caller-supplied prediction, one lane per env: state_self / state_drones rows (width 20 + 3F).
// hideandseek.py:834-887.  Used when A > 3 (hs_fill_kernel's 4-lane mapping stops at 3 pursuers).
// =========================================================================================
template <int A>
__global__ void __launch_bounds__(128)
hs_fill_wide_kernel(const __grid_constant__ KParams P) {
    const hs_config& c = P.c;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int E = c.num_envs;
    if (e >= E) return;
    const int F = c.future_step, F3 = 3 * F, D = 20 + F3;
    const float* const arena = P.b.arena;
    const int64_t Ep = P.Ep;
    const V3 tp = mk(arena[(ND * A + E_TPOS) * Ep + e], arena[(ND * A + E_TPOS + 1) * Ep + e], arena[(ND * A + E_TPOS + 2) * Ep + e]);
    const float progress = arena[(ND * A + E_PROGRESS) * Ep + e];
    const bool bdetect = arena[(ND * A + E_BDETECT) * Ep + e] != 0.0f;
    const float tfrac = fdiv(progress, (float)c.max_episode_length);
    const float mv = c.mask_value;
    float pw[3 * FMAX];
    const float* pr = P.tp_pred + e * F3;
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
        if (f < F) {
            pw[3 * f] = (__ldg(pr + 3 * f) * 0.5f) * c.arena_size;
            pw[3 * f + 1] = (__ldg(pr + 3 * f + 1) * 0.5f) * c.arena_size;
            pw[3 * f + 2] = ((__ldg(pr + 3 * f + 2) + 1.0f) / 2.0f) * c.max_height;
        } else { pw[3 * f] = pw[3 * f + 1] = pw[3 * f + 2] = 0.f; }
    }
    for (int a = 0; a < A; ++a) {
#define FR(k) arena[((int64_t)(k) * A + a) * Ep + e]
        const V3 p = mk(FR(D_POS), FR(D_POS + 1), FR(D_POS + 2));
        Q4 q; q.w = FR(D_ROT); q.x = FR(D_ROT + 1); q.y = FR(D_ROT + 2); q.z = FR(D_ROT + 3);
        const V3 lv = mk(FR(D_LIN), FR(D_LIN + 1), FR(D_LIN + 2));
#undef FR
        float rp[3 * FMAX];
#pragma unroll
        for (int f = 0; f < FMAX; ++f) { rp[3 * f] = p.x - pw[3 * f]; rp[3 * f + 1] = p.y - pw[3 * f + 1]; rp[3 * f + 2] = p.z - pw[3 * f + 2]; }
        V3 heading, up;
        heading_up(q, heading, up);
        const V3 t_rpos = p - tp;
        const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
        write_self_row(P.b.state_self + (e * A + a) * D, head_m, F3, rp, q, lv, heading, up, tfrac);
        write_self_row(P.b.state_drones + (e * A + a) * D, t_rpos, F3, rp, q, lv, heading, up, tfrac);
    }
}

}  // namespace
