// hs_tick_wide.cuh -- the control tick with ONE LANE PER ENVIRONMENT (hs_tick_wide_kernel) and the matching row-fill kernel.
//
// The bandwidth-bound mapping for batches that fill the machine, and the only one for more than 3 pursuers (the 4-lane
// mapping of hs_tick.cuh fixes A <= 3).  A CTA of A + 1 warps owns a tile of 32 consecutive envs: LANE = ENV, WARP = BODY
// (warps 0..A-1 the pursuers, warp A the evader and the env's bookkeeping), i.e. the 4-lane mapping transposed, so that
// every warp-wide access to the SoA state is one 128 B row and every lane of every warp is busy:
//   * the tile's SoA state - rows [23 A + 8 + 3 C] x 32 envs, 128 B per row - arrives with ONE TMA tensor copy
//     (cp.async.bulk.tensor.2d, SASS UTMALDG) into shared memory, the 24 stats rows with a second one, the AoS action and
//     prev_action spans with two bulk copies; one mbarrier (complete_tx) per tile;
//   * the warps read state as tile[row][lane] - conflict free - and update it in place; what one body needs from another
//     (thrust vectors and old positions for the downwash, the evader's repulsion terms, the per-pursuer reward terms)
//     goes through a small exchange area and four block barriers per tick;
//   * the updated state / stats tiles leave with TMA tensor stores (UTMASTG), every AoS output ([32 envs][W words], one
//     contiguous span per tensor) is staged in shared memory and leaves with one bulk store (UBLKCP);
//   * the previous TP window is shifted by one frame global -> global with coalesced 16 B copies while the tile is in
//     flight (it never enters shared memory), the new frame is written directly (64 B per env, sector exact).
// Arithmetic = the device functions of hs_stages.cuh in the same order as hs_tick_body: results are bit-identical to
// the 4-lane kernel (tests/test_gpu_wide.py).
#pragma once
#include <cuda.h>                    // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "hs_common.cuh"
#include "hs_stages.cuh"

namespace {

constexpr int WIDE_WARPS = 1;        // tiles per CTA (a CTA = one 32-env tile: A pursuer warps + one evader warp)
constexpr int WIDE_MAX_A = HS_MAX_AGENTS;

// shared-memory plan of one warp (offsets in floats, every region 128 B aligned)
struct WidePlan {
    int rows_all, rows_rw;           // state tile rows: all (loaded), read-write prefix (stored back)
    int o_state, o_action, o_prev;
    int o_dstate, o_others, o_cyl, o_cmds, o_ctbr, o_trate, o_aerr, o_reward, o_gt, o_self, o_drones, o_xchg;
    int total;                       // floats per warp (the mbarrier lives in the first 32 floats)
};
__host__ __device__ inline int wide_up(int words) { return (words + 31) & ~31; }
__host__ __device__ inline WidePlan wide_plan(int A, int C, int K, bool tp) {
    WidePlan w;
    w.rows_rw = ND * A + E_CYL;
    w.rows_all = w.rows_rw + 3 * C;
    int o = 32;                                                   // [0, 32): mbarrier
    w.o_state = o;  o += w.rows_all * 32;
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
    w.o_xchg = o;   o += 32 * (9 * A + 4);                        // see WX_* below (the reward terms alias the action tile)
    w.total = (o + 31) & ~31;
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
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");      // suspend-time hint: sleep, do not poll
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
template <int VEC, int NT>
__device__ __forceinline__ void wd_window_shift(float* dst, const float* src, int nenv, int per_env, int keep, int FD, int lane) {
    constexpr int U = 12;                                        // loads of a batch first, then its stores (12 x 16 B in flight per lane)
    const int keepv = keep / VEC;
    const int total = nenv * keepv;
    for (int base = 0; base < total; base += NT * U) {
        float4 r4[U];
        float r1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * NT + lane;
            if (i < total) {
                const int env = i / keepv, j = i - env * keepv;
                if (VEC == 4) r4[u] = __ldg(reinterpret_cast<const float4*>(src + env * per_env + FD) + j);
                else r1[u] = __ldg(src + env * per_env + FD + j);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * NT + lane;
            if (i < total) {
                const int env = i / keepv, j = i - env * keepv;
                if (VEC == 4) *(reinterpret_cast<float4*>(dst + env * per_env) + j) = r4[u];
                else dst[env * per_env + j] = r1[u];
            }
        }
    }
}

// The common shape (frame width a multiple of 4, H = 10): every index is a compile-time constant, the NT threads of the
// CTA move the tile's 32 x KEEP4 16-byte chunks in one batch - all loads, then all stores.
template <int KEEP4, int PE4, int FD4, int NT>
__device__ __forceinline__ void wd_window_shift_fixed(float* dst, const float* src, int nenv, int tid) {
    constexpr int U = (32 * KEEP4 + NT - 1) / NT;
    const int total = nenv * KEEP4;
    float4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = u * NT + tid;
        const int env = i / KEEP4, j = i - env * KEEP4;            // division by a constant
        if (i < total) r[u] = __ldg(reinterpret_cast<const float4*>(src) + env * PE4 + FD4 + j);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = u * NT + tid;
        const int env = i / KEEP4, j = i - env * KEEP4;
        if (i < total) *(reinterpret_cast<float4*>(dst) + env * PE4 + j) = r[u];
    }
}

// Exchange area of a tile (floats, [slot][32 lanes]): what the pursuer warps and the evader warp hand each other.
// Phase 1 -> 2: thrust in world frame, repulsion term of the evader, position before the integration (3 floats each per
// pursuer, slots 9 a + ...), then the evader's new position and the progress counter (slots 9 A ...).
// Phase 3 -> 4: four words per pursuer - r_dist, r_smooth, throttle difference and a word of small integers (collision
// counts, indicator bits) from which the other terms are rebuilt with the same operations; they live in the action tile,
// which nobody reads after phase 1.
constexpr int WX_FW = 0, WX_FP = 3, WX_POLD = 6;                 // + 9 * a
enum { WT_DIST = 0, WT_SMOOTH, WT_TDIFF, WT_FLAGS, WX_TERMS };    // + WX_TERMS * a in the action tile
// flags word: bit 0 seen_capture, 1 blocked, 2 detect, 3 speeding, 4..7 hit_wall, 8..11 hit_cyl, 12..15 hit_drone
__device__ __forceinline__ float wd_coll_reward(const hs_config& c, int flags) {
    float r = -c.collision_coef * (float)((flags >> 8) & 15);
    r = r + (-c.collision_coef * (float)((flags >> 12) & 15));
    return r + (-c.collision_coef * (float)((flags >> 4) & 15));      // same association as stage_reward_terms
}

template <int A, int CT, bool RESET>
__global__ void __launch_bounds__((A + 1) * 32, (A <= 3) ? 6 : 1)
hs_tick_wide_kernel(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tm_state_ld,
                    const __grid_constant__ CUtensorMap tm_state_st) {
    extern __shared__ __align__(1024) float wide_mem[];
    const hs_config& c = P.c;
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;   // role < A: pursuer `role`; role == A: evader + env bookkeeping
    const bool is_drone = role < A;
    const int a = is_drone ? role : 0;
    const int E = c.num_envs;
    const int64_t e0 = (int64_t)blockIdx.x * 32;
    const int C = c.num_cylinders, K = c.obs_max_cylinder, H = c.history_step;
    constexpr int FD0 = 7 + 3 * A;                               // frame without cylinders
    const bool tp_on = c.use_tp_net != 0;
    const int FD = FD0 + (c.use_obstacles ? 3 * C : 0);          // hideandseek.py:808-817
    const WidePlan w = wide_plan(A, C, K, tp_on);
    float* const mem = wide_mem;
    const int nenv = (int)min((int64_t)32, E - e0);
    const bool full = nenv == 32;
    const int64_t e = e0 + lane;
    const bool valid = lane < nenv;
    const int64_t ec = valid ? e : (int64_t)E - 1;               // clamped index for the few direct global reads
    const float dt = c.dt;
    const bool raw = P.action_is_raw != 0;

    float* const S = mem + w.o_state;                            // S[row * 32 + lane]
    float* const TERMS = mem + w.o_action;                       // reward terms of phase 3 reuse the action tile
    float* const X = mem + w.o_xchg;                             // X[slot * 32 + lane]
    float4* const s_act = reinterpret_cast<float4*>(mem + w.o_action);
    float4* const s_prev = reinterpret_cast<float4*>(mem + w.o_prev);
    const uint32_t bar = wd_smem(mem);

    // ---- loads: one mbarrier per tile, one thread issues ----------------------------------------
    if (threadIdx.x == 0) {
        wd_mbar_init(bar);
        uint32_t bytes = (uint32_t)w.rows_all * 128u;
        if (!RESET) {
            bytes += (uint32_t)nenv * A * 16u;
            if (raw) bytes += (uint32_t)nenv * A * 16u;
        }
        wd_mbar_expect(bar, bytes);
        wd_tma_load_2d(wd_smem(S), &tm_state_ld, 0, (int)(e0 >> 5) * P.R, bar);      // the tile = R consecutive 128 B rows
        if (!RESET) {
            wd_bulk_load(wd_smem(s_act), P.action + e0 * A * 4, (uint32_t)nenv * A * 16u, bar);
            if (raw) wd_bulk_load(wd_smem(s_prev), P.b.prev_action + e0 * A * 4, (uint32_t)nenv * A * 16u, bar);
        }
    }
    __syncthreads();                                              // the barrier is initialised before anyone polls it
    // ---- the TP window moves global -> global while the tile is in flight (all warps: one round trip, under the TMA wait)
    const int per_env = H * FD, keep = (H - 1) * FD;
    // ring form of the window (hs_buffers.tp_ring): nothing to shift - the new frame is written twice in phase 4
    const bool ring = tp_on && P.b.tp_ring != nullptr;
    const int ring_pos = ring ? P.b.tp_ring_pos[blockIdx.x] : 0;
    if (tp_on && !ring && !P.tp_init && keep > 0) {
        const float* src = P.b.tp_input_prev + e0 * per_env;
        float* dst = P.b.tp_input + e0 * per_env;
        constexpr int NT = (A + 1) * 32;
        if ((FD0 & 3) == 0 && H == 10 && FD == FD0) wd_window_shift_fixed<9 * (FD0 / 4), 10 * (FD0 / 4), FD0 / 4, NT>(dst, src, nenv, threadIdx.x);
        else if ((FD & 3) == 0) wd_window_shift<4, NT>(dst, src, nenv, per_env, keep, FD, threadIdx.x);
        else wd_window_shift<1, NT>(dst, src, nenv, per_env, keep, FD, threadIdx.x);
    }
    bool pid_reset = false;
    float v_prey = 0.f;
    if (!RESET) {
        if (is_drone) pid_reset = raw && (P.reset_pid != nullptr) && (P.reset_pid[ec] != 0);
        else v_prey = __ldg(P.b.v_prey);
    }
    wd_mbar_wait(bar, 0);

#define SD(k, aa) S[((k) * A + (aa)) * 32 + lane]
#define SE(k) S[(ND * A + (k)) * 32 + lane]
#define XS(slot) X[(slot) * 32 + lane]
    V3 tp = mk(SE(E_TPOS), SE(E_TPOS + 1), SE(E_TPOS + 2));
    V3 tv = mk(SE(E_TVEL), SE(E_TVEL + 1), SE(E_TVEL + 2));
    float progress = SE(E_PROGRESS);
    float cx[CT], cy[CT], cz[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        if (k < C) { cx[k] = SE(E_CYL + 3 * k); cy[k] = SE(E_CYL + 3 * k + 1); cz[k] = SE(E_CYL + 3 * k + 2); }
        else { cx[k] = 0.f; cy[k] = 0.f; cz[k] = -20.f; }
    }

    // ================= phase 1: control (pursuer warps) ============================================
    V3 p = mk(0, 0, 0), lv = mk(0, 0, 0), av = mk(0, 0, 0);
    Q4 q; q.w = 1.f; q.x = q.y = q.z = 0.f;
    float T[4] = {0, 0, 0, 0};
    float yaw_torque = 0.f, action_err = 0.f, throttle_diff = 0.f;
    bool out_of_arena = false;
    if (is_drone) {
        p = mk(SD(D_POS, a), SD(D_POS + 1, a), SD(D_POS + 2, a));
        q.w = SD(D_ROT, a); q.x = SD(D_ROT + 1, a); q.y = SD(D_ROT + 2, a); q.z = SD(D_ROT + 3, a);
        lv = mk(SD(D_LIN, a), SD(D_LIN + 1, a), SD(D_LIN + 2, a));
        av = mk(SD(D_ANG, a), SD(D_ANG + 1, a), SD(D_ANG + 2, a));
        if (!RESET) {
            float thr[4] = {SD(D_THR, a), SD(D_THR + 1, a), SD(D_THR + 2, a), SD(D_THR + 3, a)};
            const float4 act = s_act[lane * A + a];
            float cmd[4];
            if (raw) {
                V3 integ = mk(SD(D_INT, a), SD(D_INT + 1, a), SD(D_INT + 2, a));
                V3 last = mk(SD(D_LAST, a), SD(D_LAST + 1, a), SD(D_LAST + 2, a));
                CtbrOut o;
                stage_ctbr_pid(c, act, s_prev[lane * A + a], pid_reset, q, av, integ, last, o);
                action_err = o.action_err;
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
                action_err = P.b.action_error[ec * A + a];
            }
            stage_rotor(c, cmd, thr, T, yaw_torque, throttle_diff);
            if (P.b.throttle_diff != nullptr && valid) P.b.throttle_diff[e * A + a] = throttle_diff;
            if (valid) { SD(D_THR, a) = thr[0]; SD(D_THR + 1, a) = thr[1]; SD(D_THR + 2, a) = thr[2]; SD(D_THR + 3, a) = thr[3]; }
            const float total_thrust = ((T[0] + T[1]) + T[2]) + T[3];
            const V3 Fw = qrot<false>(q, mk(0.f, 0.f, total_thrust));
            const V3 fp = evader_pursuer_term(c, p, tp, los_blocked(p, tp, cx, cy, cz, C, c.cylinder_size));
            XS(9 * a + WX_FW) = Fw.x; XS(9 * a + WX_FW + 1) = Fw.y; XS(9 * a + WX_FW + 2) = Fw.z;
            XS(9 * a + WX_FP) = fp.x; XS(9 * a + WX_FP + 1) = fp.y; XS(9 * a + WX_FP + 2) = fp.z;
            XS(9 * a + WX_POLD) = p.x; XS(9 * a + WX_POLD + 1) = p.y; XS(9 * a + WX_POLD + 2) = p.z;
        }
    }
    // (1) thrusts, repulsion terms, old positions published - and every warp has read the evader rows of the tile, which
    // the evader warp overwrites next (also in the RESET variant: compute-sanitizer racecheck found that one)
    __syncthreads();

    // ================= phase 2: wrench + integration (pursuers) | evader policy (evader warp) =======
    if (is_drone) {
        V3 ext = mk(0.f, 0.f, 0.f);
        if (!RESET) {
            V3 dw = mk(0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < A; ++j) {
                if (j != a) {
                    const V3 Fj = mk(XS(9 * j + WX_FW), XS(9 * j + WX_FW + 1), XS(9 * j + WX_FW + 2));
                    const V3 pj = mk(XS(9 * j + WX_POLD), XS(9 * j + WX_POLD + 1), XS(9 * j + WX_POLD + 2));
                    dw = dw + downwash_term(c, Fj, pj, p);
                }
            }
            ext = dw + lv * c.drag_coef_times_mass;
        }
        stage_integrate<!RESET>(c, p, q, lv, av, T, yaw_torque, ext);
        stage_contacts(c, p, lv, tp, cx, cy, cz, C);
        if (valid) {
            SD(D_POS, a) = p.x; SD(D_POS + 1, a) = p.y; SD(D_POS + 2, a) = p.z;
            SD(D_ROT, a) = q.w; SD(D_ROT + 1, a) = q.x; SD(D_ROT + 2, a) = q.y; SD(D_ROT + 3, a) = q.z;
            SD(D_LIN, a) = lv.x; SD(D_LIN + 1, a) = lv.y; SD(D_LIN + 2, a) = lv.z;
            SD(D_ANG, a) = av.x; SD(D_ANG + 1, a) = av.y; SD(D_ANG + 2, a) = av.z;
        }
        // info.drone_state [E,A,13]
        float* r = mem + w.o_dstate + (lane * A + a) * 13;
        r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = q.w; r[4] = q.x; r[5] = q.y; r[6] = q.z;
        r[7] = lv.x; r[8] = lv.y; r[9] = lv.z; r[10] = av.x; r[11] = av.y; r[12] = av.z;
    } else {
        if (!RESET) {
            V3 force = mk(XS(WX_FP), XS(WX_FP + 1), XS(WX_FP + 2));
#pragma unroll
            for (int j = 1; j < A; ++j) force = force + mk(XS(9 * j + WX_FP), XS(9 * j + WX_FP + 1), XS(9 * j + WX_FP + 2));
            tv = evader_velocity(c, force, tp, cx, cy, cz, C, v_prey, out_of_arena);
        }
        tp = tp + tv * dt;
        if (!RESET) progress = progress + 1.0f;
        else if (P.env_mask == nullptr || P.env_mask[ec]) progress = 0.0f;
        // published through the exchange-free rows of the tile: pursuer warps read them after barrier (2)
        if (valid) {
            SE(E_TPOS) = tp.x; SE(E_TPOS + 1) = tp.y; SE(E_TPOS + 2) = tp.z;
            if (!RESET) { SE(E_TVEL) = tv.x; SE(E_TVEL + 1) = tv.y; SE(E_TVEL + 2) = tv.z; }
            SE(E_PROGRESS) = progress;
        }
        XS(9 * A + 0) = tp.x; XS(9 * A + 1) = tp.y; XS(9 * A + 2) = tp.z;      // (also for invalid lanes: no garbage downstream)
        XS(9 * A + 3) = progress;
    }
    __syncthreads();                                              // (2) new poses, new evader position published

    // ================= phase 3: observation + per-pursuer reward terms ============================
    const V3 tpn = mk(XS(9 * A + 0), XS(9 * A + 1), XS(9 * A + 2));
    progress = XS(9 * A + 3);
    const float mv = c.mask_value;
    const float tfrac = fdiv(progress, (float)c.max_episode_length);
    float hit_drone = 0.f, hit_cyl = 0.f;
    bool blocked = false, detect = false;
    RewardTerms rt;
    rt.r_dist = rt.r_speed = rt.r_coll = rt.r_smooth = rt.hit_wall = 0.f; rt.seen_capture = false;
    // ... and while the pursuer warps build the observation, it fetches the env's 24 stats (coalesced 128 B rows, in
    // registers until phase 4)
    float st_old[HS_NUM_STATS];
    if (!RESET && !is_drone) {
#pragma unroll
        for (int k = 0; k < HS_NUM_STATS; ++k) st_old[k] = __ldg(P.b.stats + (int64_t)k * E + ec);
    }
    const float sm_coef = (!RESET && P.b.smoothness_coef != nullptr) ? __ldg(P.b.smoothness_coef) : c.smoothness_coef;
    if (is_drone) {
        if (A > 1) {
            float* r = mem + w.o_others + (lane * A + a) * ((A - 1) * 3);
            int o = 0;
#pragma unroll
            for (int j = 0; j < A; ++j) {
                if (j != a) {
                    const V3 pj = mk(SD(D_POS, j), SD(D_POS + 1, j), SD(D_POS + 2, j));
                    const V3 d = p - pj;
                    r[o * 3] = d.x; r[o * 3 + 1] = d.y; r[o * 3 + 2] = d.z;
                    hit_drone = hit_drone + ((norm3(d) < c.coll_radius_x2) ? 1.0f : 0.0f);
                    ++o;
                }
            }
        }
        if (K > 0) stage_knearest(c, p, cx, cy, cz, C, K, mem + w.o_cyl + (lane * A + a) * (K * 5), hit_cyl);
        blocked = los_blocked(p, tpn, cx, cy, cz, C, c.cylinder_size);
        detect = (norm3(p - tpn) < c.drone_detect_radius) && !blocked;
        if (!RESET) rt = stage_reward_terms(c, p, lv, tpn, blocked, hit_cyl, hit_drone, action_err, sm_coef);
    }
    if (is_drone) {
        float* t = TERMS + (WX_TERMS * a) * 32 + lane;
        t[WT_DIST * 32] = rt.r_dist; t[WT_SMOOTH * 32] = rt.r_smooth; t[WT_TDIFF * 32] = throttle_diff;
        t[WT_FLAGS * 32] = __int_as_float((rt.seen_capture ? 1 : 0) | (blocked ? 2 : 0) | (detect ? 4 : 0) | ((rt.r_speed != 0.0f) ? 8 : 0) |
                                          ((int)rt.hit_wall << 4) | ((int)hit_cyl << 8) | ((int)hit_drone << 12));
    }
    __syncthreads();                                              // (3) terms published

    // ================= phase 4: cooperative terms, rows / TP frame, stats ==========================
    bool bdetect = false, any_capture = false, all_blocked = true, any_coll = false;
#pragma unroll
    for (int j = 0; j < A; ++j) {
        const int f = __float_as_int(TERMS[(WX_TERMS * j + WT_FLAGS) * 32 + lane]);
        any_capture = any_capture || (f & 1);
        all_blocked = all_blocked && (f & 2);
        bdetect = bdetect || (f & 4);
        any_coll = any_coll || (wd_coll_reward(c, f) < 0.0f);
    }
    const float r_detect = c.detect_reward_coef * (bdetect ? 1.0f : 0.0f);
    const float r_catch = c.catch_reward_coef * (any_capture ? 1.0f : 0.0f);
    if (is_drone) {
        if (!RESET)
            mem[w.o_reward + lane * A + a] = ((((rt.r_dist + r_detect) + r_catch) + rt.r_coll) + rt.r_speed) + rt.r_smooth;
        if (!tp_on) {
            // no predictor: the rows are complete now (width 20)
            V3 heading, up;
            heading_up(q, heading, up);
            const V3 t_rpos = p - tpn;
            const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
            write_self_row(mem + w.o_self + (lane * A + a) * 20, head_m, 0, nullptr, q, lv, heading, up, tfrac);
            write_self_row(mem + w.o_drones + (lane * A + a) * 20, t_rpos, 0, nullptr, q, lv, heading, up, tfrac);
        }
    } else {
        if (tp_on) {
            // new TP frame [progress, tpos_masked3, tvel_masked3, p_0..p_{A-1}], written straight to its slot (row H-1;
            // every row on the very first frame)
            float fr[FD0 + 3 * CT];
            fr[0] = progress;
            fr[1] = bdetect ? tp.x : mv; fr[2] = bdetect ? tp.y : mv; fr[3] = bdetect ? tp.z : mv;
            fr[4] = bdetect ? tv.x : mv; fr[5] = bdetect ? tv.y : mv; fr[6] = bdetect ? tv.z : mv;
#pragma unroll
            for (int j = 0; j < A; ++j) { fr[7 + 3 * j] = SD(D_POS, j); fr[8 + 3 * j] = SD(D_POS + 1, j); fr[9 + 3 * j] = SD(D_POS + 2, j); }
#pragma unroll
            for (int k = 0; k < CT; ++k) { fr[FD0 + 3 * k] = cx[k]; fr[FD0 + 3 * k + 1] = cy[k]; fr[FD0 + 3 * k + 2] = c.cylinder_size; }
            if (valid) {
                // plain: row H-1 of the shifted window (every row on the very first frame); ring: slots p and p + H of
                // the env's 2H slots (every slot on the very first frame)
                float* win = ring ? P.b.tp_ring + e * (2 * per_env) : P.b.tp_input + e * per_env;
                const int h0 = P.tp_init ? 0 : (ring ? ring_pos : H - 1);
                const int h1 = ring ? (P.tp_init ? 2 * H : ring_pos + H + 1) : H;
                const int hs = (ring && !P.tp_init) ? H : 1;
                for (int h = h0; h < h1; h += hs) {
                    float* row = win + h * FD;
                    if ((FD0 & 3) == 0 && FD == FD0) {
#pragma unroll
                        for (int k = 0; k < FD0 / 4; ++k)
                            reinterpret_cast<float4*>(row)[k] = make_float4(fr[4 * k], fr[4 * k + 1], fr[4 * k + 2], fr[4 * k + 3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < FD0 + 3 * CT; ++k) if (k < FD) row[k] = fr[k];
                    }
                }
                P.b.tp_done[e] = (progress <= (float)(c.max_episode_length - c.future_step)) ? 1 : 0;
                SE(E_BDETECT) = bdetect ? 1.0f : 0.0f;
            }
            float* gt = mem + w.o_gt + lane * 3;
            gt[0] = fdiv(tp.x, c.half_arena);
            gt[1] = fdiv(tp.y, c.half_arena);
            gt[2] = fdiv(tp.z, c.max_height) * 2.0f - 1.0f;
        }
        if (RESET) {
            if (valid && P.b.truncated != nullptr) P.b.truncated[e] = (progress > (float)c.max_episode_length) ? 1 : 0;
        } else {
            // per-env means over the pursuers: sums in agent order (for A <= 3 the same association as the 4-lane butterfly)
            const float inv_A = 1.0f / (float)A;
            float s_ae = 0.f, s_dist = 0.f, s_speed = 0.f, s_hcyl = 0.f, s_hdrone = 0.f, s_hwall = 0.f, s_coll = 0.f, s_smooth = 0.f,
                  s_tdiff = 0.f, s_reward = 0.f, x_tdiff = -INFINITY;
#pragma unroll
            for (int j = 0; j < A; ++j) {
                const float* t = TERMS + (WX_TERMS * j) * 32 + lane;
                const int f = __float_as_int(t[WT_FLAGS * 32]);
                const float r_dist = t[WT_DIST * 32], r_smooth = t[WT_SMOOTH * 32], tdiff = t[WT_TDIFF * 32];
                const float r_speed = -c.speed_coef * ((f & 8) ? 1.0f : 0.0f);
                const float r_coll = wd_coll_reward(c, f);
                const float aerr = raw ? mem[w.o_aerr + lane * A + j] : P.b.action_error[ec * A + j];
                const float reward = ((((r_dist + r_detect) + r_catch) + r_coll) + r_speed) + r_smooth;
                if (j == 0) {
                    s_ae = aerr; s_dist = r_dist; s_speed = r_speed; s_hcyl = (float)((f >> 8) & 15); s_hdrone = (float)((f >> 12) & 15);
                    s_hwall = (float)((f >> 4) & 15); s_coll = r_coll; s_smooth = r_smooth; s_tdiff = tdiff; s_reward = reward;
                } else {
                    s_ae += aerr; s_dist += r_dist; s_speed += r_speed; s_hcyl += (float)((f >> 8) & 15); s_hdrone += (float)((f >> 12) & 15);
                    s_hwall += (float)((f >> 4) & 15); s_coll += r_coll; s_smooth += r_smooth; s_tdiff += tdiff; s_reward += reward;
                }
                x_tdiff = fmaxf(x_tdiff, tdiff);
            }
            float s_detect = r_detect, s_catch = r_catch;
#pragma unroll
            for (int j = 1; j < A; ++j) { s_detect += r_detect; s_catch += r_catch; }
            EnvTick et;
            et.m_ae = s_ae * inv_A; et.m_dist = s_dist * inv_A; et.m_detect = s_detect * inv_A; et.m_catch = s_catch * inv_A;
            et.m_speed = s_speed * inv_A; et.m_hcyl = s_hcyl * inv_A; et.m_hdrone = s_hdrone * inv_A;
            et.m_hwall = s_hwall * inv_A; et.m_coll = s_coll * inv_A; et.m_smooth = s_smooth * inv_A;
            et.m_tdiff = s_tdiff * inv_A; et.m_reward = s_reward * inv_A; et.x_tdiff = x_tdiff;
            et.r_catch = r_catch; et.bdetect = bdetect; et.all_blocked = all_blocked; et.any_coll = any_coll; et.out_of_arena = out_of_arena;
            if (valid) {
                P.b.done[e] = (progress >= (float)c.max_episode_length) ? 1 : 0;
                float* SG = P.b.stats + e;
                const int64_t Es = E;
                stage_stats(c, et, progress, sm_coef, [&](int k) { return st_old[k]; }, [&](int k, float v) { SG[(int64_t)k * Es] = v; });
            }
        }
    }
#undef SD
#undef SE
#undef XS

    // ---- stores: state / stats tiles by TMA tensor store, AoS outputs by bulk store (warp 0 issues) ----
    fence_async_smem();
    __syncthreads();                                              // (4)
    if (role == 0) {
        if (lane == 0) {
            wd_tma_store_2d(&tm_state_st, 0, (int)(e0 >> 5) * P.R, wd_smem(S));
            if (ring) P.b.tp_ring_pos[blockIdx.x] = (ring_pos + 1 == H) ? 0 : ring_pos + 1;
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
            bulk_wait_read<0>();                                  // the CTA's shared memory lives until its last thread exits:
        }                                                         // this one stays until the bulk engine has read it
    }
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
    const float* const tile = P.b.arena + (e >> 5) * P.R * 32 + (e & 31);       // + row * 32
    const V3 tp = mk(tile[(ND * A + E_TPOS) * 32], tile[(ND * A + E_TPOS + 1) * 32], tile[(ND * A + E_TPOS + 2) * 32]);
    const float progress = tile[(ND * A + E_PROGRESS) * 32];
    const bool bdetect = tile[(ND * A + E_BDETECT) * 32] != 0.0f;
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
#define FR(k) tile[((k) * A + a) * 32]
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
