// hs_rollout_pair.cuh -- hs_rollout_pair_kernel: the rollout kernel of hs_rollout_fused.cuh with the predictor advancing
// a GROUP of NS = 2 or 3 consecutive ticks at a time.  Part of the single translation unit hs_kernels.cu.
//
// Why: the phase timeline of hs_rollout_fused_kernel (tools/rollout_phases.py) shows the recurrence bound by the tensor
// pipe's INSTRUCTION rate - 120 tcgen05.mma of shape 128 x 16 x 8 per LSTM step (two 16-env halves x two M-tiles x 30) at
// ~15 ns each = the 1.7 us step.  The halves exist only to give the pipe something to do while the epilogue warps turn the
// other half's accumulator into h.  Two consecutive TICKS are just as independent (the LSTM starts from zero state every
// tick), so this kernel ping-pongs the full 32-env tiles of ticks 2p and 2p+1 instead: 128 x 32 x 8 MMAs, 60 per tick and
// LSTM step - half the tensor-pipe instructions for the same arithmetic.
//
// NS = 3 (a third stream in the bubbles of the MMA -> commit -> epilogue -> h chain) is built too but buys nothing: with
// two streams the predictor already keeps up with the tick warps, whose ~8 us tick body takes ~15 us next to the busy
// epilogue warps and now bounds the kernel (1 / 2 / 3 ticks per pass: 17.96 / 15.35 / 15.35 us per tick at 4096 envs).
//
// What changes against hs_rollout_fused_kernel:
//   * the tick warps run one group ahead; the B-operand ring has H + 2 NS - 1 step slots (frame f lives in slot
//     f mod S): the NS + 9 frames of a group's windows and the NS frames the tick warps write meanwhile never share a slot;
//   * what FC + rows need from the state of a tick (pose, velocity, evader position, progress, detect flag) is stashed by
//     the tick warps in shared memory (ring of two groups: the arena already holds a later tick when the rows are built);
//   * buffer tables: ring of two groups; "tick done" uses one named barrier per position in the group (NS arrivals can be
//     pending); "group free" is raised by the predictor warps at the top of a group, once the previous one is complete.
// Results are bit-identical to hs_rollout_fused_kernel and to T calls of hs_step_fused.
#pragma once
#include "hs_rollout_fused.cuh"

namespace {

constexpr int RP_H = 10;                                       // history_step this kernel is built for (checked by the host)
constexpr int RP_STASH = 35;                                   // floats per env: 3 x (pos3, quat4, linvel3) + tpos3 + progress + detect
constexpr int RP_BAR_MAIN = 1, RP_BAR_GROUP_FREE = 3, RP_BAR_TICKW = 4;
// "tick done" of the j-th tick of a group: one named barrier each (NS arrivals can be pending)
__device__ __forceinline__ void rp_done_arrive(int j) {
    if (j == 0) asm volatile("bar.arrive 2, %0;" :: "n"(RF_THREADS) : "memory");
    else if (j == 1) asm volatile("bar.arrive 5, %0;" :: "n"(RF_THREADS) : "memory");
    else asm volatile("bar.arrive 6, %0;" :: "n"(RF_THREADS) : "memory");
}
__device__ __forceinline__ void rp_done_sync(int j) {
    if (j == 0) asm volatile("bar.sync 2, %0;" :: "n"(RF_THREADS) : "memory");
    else if (j == 1) asm volatile("bar.sync 5, %0;" :: "n"(RF_THREADS) : "memory");
    else asm volatile("bar.sync 6, %0;" :: "n"(RF_THREADS) : "memory");
}
// operand ring: the NS + 9 frames of a group's windows and the NS frames of the next group never share a slot
__host__ __device__ constexpr int rp_slots(int NS) { return RP_H + 2 * NS - 1; }

struct PairHook {
    static constexpr bool ACTIVE = true;
    uint8_t *xhi, *xlo;              // operand slot of this tick's frame + (tick warp) * TN_SBO
    float* stash;                    // [RP_STASH][32 envs] of this tick
    int nenv_w, tw;
    template <int FD> __device__ __forceinline__ void frame(const float* tile, int per_env, int keep, int lane) const {
        const int rr = lane & 7, kk = lane >> 3;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            const int k = kc * 4 + kk;
            const float xv = (rr < nenv_w && k < FD) ? tile[rr * per_env + keep + k] : 0.0f;
            uint32_t hi, lo;
            tf32_split(xv, hi, lo);
            const uint32_t off = kc * TN_X_LBO + rr * 16 + kk * 4;
            *reinterpret_cast<uint32_t*>(xhi + off) = hi;
            *reinterpret_cast<uint32_t*>(xlo + off) = lo;
        }
    }
    __device__ __forceinline__ void state(bool is_drone, bool is_ev, int slot, int env_w, const V3& p, const Q4& q, const V3& lv,
                                          const V3& tp, float progress, bool bdetect) const {
        float* s = stash + tw * ENVS_PER_WARP + env_w;         // column of this env; row r at s[r * 32]
        if (is_drone) {
            float* d = s + slot * 10 * TN_E;
            d[0] = p.x; d[TN_E] = p.y; d[2 * TN_E] = p.z;
            d[3 * TN_E] = q.w; d[4 * TN_E] = q.x; d[5 * TN_E] = q.y; d[6 * TN_E] = q.z;
            d[7 * TN_E] = lv.x; d[8 * TN_E] = lv.y; d[9 * TN_E] = lv.z;
        }
        if (is_ev) {
            float* d = s + 30 * TN_E;
            d[0] = tp.x; d[TN_E] = tp.y; d[2 * TN_E] = tp.z; d[3 * TN_E] = progress; d[4 * TN_E] = bdetect ? 1.0f : 0.0f;
        }
    }
};

template <int A>
__device__ __forceinline__ TnRowIn tn_row_from_stash(const float* stash) {
    TnRowIn R;
    R.p = mk(0.f, 0.f, 0.f); R.lv = R.p; R.tp = R.p; R.q.w = 1.f; R.q.x = R.q.y = R.q.z = 0.f; R.progress = 0.f; R.bdetect = false;
    const int tid = threadIdx.x;
    if (tid < TN_E * A) {
        const int slot = tid / TN_E, el = tid - slot * TN_E;
        const float* d = stash + slot * 10 * TN_E + el;
        R.p = mk(d[0], d[TN_E], d[2 * TN_E]);
        R.q.w = d[3 * TN_E]; R.q.x = d[4 * TN_E]; R.q.y = d[5 * TN_E]; R.q.z = d[6 * TN_E];
        R.lv = mk(d[7 * TN_E], d[8 * TN_E], d[9 * TN_E]);
        const float* ev = stash + 30 * TN_E + el;
        R.tp = mk(ev[0], ev[TN_E], ev[2 * TN_E]);
        R.progress = ev[3 * TN_E];
        R.bdetect = ev[4 * TN_E] != 0.0f;
    }
    return R;
}

template <int A, int CT, int NS>
__global__ void __launch_bounds__(RF_THREADS, 1)
hs_rollout_pair_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W, const __grid_constant__ RolloutParams RP) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int RING = 2 * NS;                       // depth of the table / stash rings: two groups
    constexpr int RP_SLOTS = rp_slots(NS);
    __shared__ hs_buffers sB[RING];                    // buffer table of tick t at sB[t % RING]
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    constexpr int NTH = RF_MAIN_THREADS;
    const int H = c.history_step;                      // == RP_H (checked by the host)
    const int F3 = 3 * c.future_step;
    const int E = c.num_envs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = RP.num_ticks;

    uint8_t* H0 = smem_raw;                                     // stream 0: hi, then lo
    float* fcw = reinterpret_cast<float*>(H0 + 2 * TN_H_BYTES);
    float* fcb = fcw + F3 * TP_HID;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(fcb + 32);    // d_ready[4], h_ready[4] (index = stream)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 8);
    uint8_t* Xhi = reinterpret_cast<uint8_t*>(mbar + 10);
    uint8_t* Xlo = Xhi + (size_t)RP_SLOTS * TN_X_STEP;
    float* preds = reinterpret_cast<float*>(Xlo + (size_t)RP_SLOTS * TN_X_STEP);
    float* rowbuf = preds + TN_E * 3 * FMAX;
    float* wst = rowbuf + TN_E * A * (20 + 3 * FMAX);           // weight staging tile; after the prologue:
    float* rowbuf2 = wst;                                       //   second row tile
    uint8_t* H1 = reinterpret_cast<uint8_t*>(rowbuf2 + TN_E * A * (20 + 3 * FMAX));      //   h (hi, lo) of streams 1 .. NS-1
    float* stash = reinterpret_cast<float*>(H1 + (NS - 1) * 2 * TN_H_BYTES);             //   RING x [RP_STASH][32]
    static_assert((size_t)TN_E * A * (20 + 3 * FMAX) * 4 + (size_t)(NS - 1) * 2 * TN_H_BYTES + (size_t)RING * RP_STASH * TN_E * 4
                  <= (size_t)256 * TN_WPITCH * 4, "the rings must fit the dead weight staging tile");
    auto Hbuf = [&](int k) -> uint8_t* { return k == 0 ? H0 : H1 + (size_t)(k - 1) * 2 * TN_H_BYTES; };
    auto dcol = [](int k) -> uint32_t { return k < 2 ? (uint32_t)(k * 2 * TN_E) : (uint32_t)(TN_COL_A + 320); };   // 0, 64, 448
    float* tick_mem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wst + 256 * TN_WPITCH) + 127) & ~(uintptr_t)127);
    auto slot_of = [](int f) { return f % RP_SLOTS; };          // frame produced by tick t: f = H + t

    if (warp >= NTH / 32) {
        // ================= tick warps: up to two ticks ahead of the predictor warps =================
        const int tw = warp - NTH / 32;
        const int ttid = tid - NTH;
        float* m = tick_mem + tw * FUSED_TICK_WORDS;
        const int64_t warp_g = (int64_t)blockIdx.x * FUSED_TICK_WARPS + tw;
        {   // chronological window before the first tick (for TP_input): from then on it lives in the warp's tile
            const int64_t ew = warp_g * ENVS_PER_WARP;
            const int nw = (int)max((int64_t)0, min((int64_t)ENVS_PER_WARP, E - ew)) * H * FD;
            float* tile = m + 2 * TICK_STAGE_WORDS;
            for (int i = lane; i < nw; i += 32) tile[i] = RP.first_tp_prev[ew * (H * FD) + i];
            __syncwarp();
        }
        constexpr int TBL_WORDS = (int)(sizeof(hs_buffers) / 4);
        static_assert(TBL_WORDS <= 32 * FUSED_TICK_WARPS, "one table word per tick thread");
        auto table_word = [&](int t) -> uint32_t {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(RP.sets + (RP.first_set + t) % RP.num_sets);
            return (ttid < TBL_WORDS && t < T) ? __ldg(src + ttid) : 0u;
        };
        uint32_t tbl = table_word(0);
        // The stash / table rings and the weight-staging region they alias become free when the prologue is over: the
        // predictor warps raise "group free" once for that, then at the top of every group but the last.
        asm volatile("bar.sync %0, %1;" :: "n"(RP_BAR_GROUP_FREE), "n"(RF_THREADS) : "memory");
        for (int t = 0, j = 0; t < T; ++t, j = (j + 1 == NS) ? 0 : j + 1) {
            // the ticks of group g+1 may start once the predictor warps have begun group g (= finished group g-1: its
            // stash, table and operand slots are free)
            if (t >= NS && j == 0) asm volatile("bar.sync %0, %1;" :: "n"(RP_BAR_GROUP_FREE), "n"(RF_THREADS) : "memory");
            if (ttid < TBL_WORDS) reinterpret_cast<uint32_t*>(&sB[t % RING])[ttid] = tbl;
            tbl = table_word(t + 1);
            asm volatile("bar.sync %0, %1;" :: "n"(RP_BAR_TICKW), "n"(32 * FUSED_TICK_WARPS) : "memory");
            const float* act = RP.action + (int64_t)t * RP.action_tick_stride;
            PairHook hook;
            hook.tw = tw;
            hook.nenv_w = (int)max((int64_t)0, min((int64_t)ENVS_PER_WARP, E - warp_g * ENVS_PER_WARP));
            hook.xhi = Xhi + (size_t)slot_of(H + t) * TN_X_STEP + (size_t)tw * TN_SBO;
            hook.xlo = Xlo + (size_t)slot_of(H + t) * TN_X_STEP + (size_t)tw * TN_SBO;
            hook.stash = stash + (t % RING) * (RP_STASH * TN_E);
            hs_tick_body<A, false, CT, true, PairHook>(P, sB[t % RING], act, warp_g, m, m + TICK_STAGE_WORDS, m + 2 * TICK_STAGE_WORDS,
                                                       m + 2 * TICK_STAGE_WORDS + ENVS_PER_WARP * TP_ENV_WORDS_MAX, hook);
            fence_async_smem();                          // operand ring: generic-proxy stores -> the MMAs' async proxy
            __threadfence_block();
            rp_done_arrive(j);
        }
        return;
    }

    // ================= predictor warps =================
    const int row = (warp & 3) * 32 + lane;            // TMEM lane = gate row of both M-tiles
    const int cg = warp >> 2;                          // 0..3: epilogue column group; 4: issuing warps
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar + k)) : "memory");        // d_ready: one commit per M-tile
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 4 + k)) : "memory");   // h_ready: 16 epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int64_t e0 = (int64_t)blockIdx.x * TN_E;
    const int nenv = (int)min((int64_t)TN_E, E - e0);
    // weights -> TMEM once per rollout; the window before the first tick -> operand slots 0..H-1
    for (int i = tid; i < F3 * TP_HID; i += NTH) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    tn_stage_weights_g2s<FD>(W, wst, tid, NTH);
    tn_stage_x_ptr<FD, NTH>(RP.first_tp_prev + e0 * (int64_t)(H * FD), (int64_t)(H * FD), nenv, H, Xhi, Xlo);
    tc_fence_before();
    tn_sync<RP_BAR_MAIN, NTH>();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    tn_stage_weights_s2t(wst, lane_base, row, cg);
    const TnLane L = tn_lane_consts(W, row);
    fence_async_smem();
    tc_fence_before();
    tn_sync<RP_BAR_MAIN, NTH>();
    tc_fence_after();
    // the weight staging tile is dead: the tick warps may use the stash that aliases it
    asm volatile("bar.arrive %0, %1;" :: "n"(RP_BAR_GROUP_FREE), "n"(RF_THREADS) : "memory");

    const uint32_t d_ready = smem_u32(mbar), h_ready = smem_u32(mbar + 4);
    uint32_t ph_d = 0u, ph_h = 0u;
    const uint32_t warp_u = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const bool issuer = warp_u >= TN_THREADS / 32;
    const uint32_t mytl = warp_u & 1u;                         // M-tile of an issuing warp (warps 16, 17)
    TnIssue I;
    I.aA_hi = tmem_u + TN_COL_A + 160 * mytl;
    I.aA_lo = I.aA_hi + 80;
    I.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN_E >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_mine = tmem_u + mytl * TN_E;              // stream k: + k * 2 * TN_E

    for (int t0 = 0; t0 < T; t0 += NS) {
        const int ns = min(NS, T - t0);                        // streams of this group: ticks t0 .. t0 + ns - 1
#pragma unroll
        for (int k = 0; k < NS; ++k)
            if (k < ns) rp_done_sync(k);
        // the tick warps may start the next group: everything it overwrites belongs to groups that are complete (raised
        // only now, with every "tick done" of this group consumed, so that no named barrier ever sees two pending phases)
        if (t0 + NS < T) asm volatile("bar.arrive %0, %1;" :: "n"(RP_BAR_GROUP_FREE), "n"(RF_THREADS) : "memory");
        tc_fence_before();
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                auto xdesc = [&](int k, int s, bool lo) {      // step s of tick t0 + k: frame (t0 + k + 1 + s)
                    return tc_desc(smem_u32(lo ? Xlo : Xhi) + (uint32_t)slot_of(t0 + k + 1 + s) * TN_X_STEP, TN_X_LBO, TN_SBO);
                };
                auto hdesc = [&](int k, bool lo) { return tc_desc(smem_u32(Hbuf(k) + (lo ? TN_H_BYTES : 0)), TN_H_LBO, TN_SBO); };
                for (int k = 0; k < ns; ++k) {
                    I.x_part(d_mine + dcol(k), xdesc(k, 0, false), xdesc(k, 0, true), 0u);
                    tc_commit(d_ready + 8u * (uint32_t)k);
                }
                for (int s = 0; s < H; ++s)
                    for (int k = 0; k < ns; ++k) {
                        mbar_wait_idx(h_ready, (uint32_t)k, ph_h);
                        if (s + 1 < H) {
                            tc_fence_after();
                            const uint32_t d = d_mine + dcol(k);
                            I.x_part(d, xdesc(k, s + 1, false), xdesc(k, s + 1, true), 0u);
                            I.h_part(d, hdesc(k, false), hdesc(k, true));
                            tc_commit(d_ready + 8u * (uint32_t)k);
                        }
                    }
            }
            __syncwarp();
        } else {
            float cst[NS][8];
#pragma unroll
            for (int k = 0; k < NS; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) cst[k][j] = 0.f;
            for (int s = 0; s < H; ++s) {
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    if (k < ns) {
                        mbar_wait_idx(d_ready, (uint32_t)k, ph_d);
                        tc_fence_after();
                        uint8_t* Hk = Hbuf(k);
                        tn_epilogue(lane_base + dcol(k), L, cg, cst[k], Hk, Hk + TN_H_BYTES);
                        fence_async_smem();                      // h (generic proxy) -> async proxy of the next MMAs
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(h_ready + 8u * (uint32_t)k);
                    }
                }
            }
        }
        tn_sync<RP_BAR_MAIN, NTH>();          // all h of the last step written; the issuing warps have consumed every arrival
        for (int k = 0; k < ns; ++k) {
            const int t = t0 + k;
            const TnRowIn RI = tn_row_from_stash<A>(stash + (t % RING) * (RP_STASH * TN_E));
            float* pred_out = (RP.pred_out != nullptr) ? RP.pred_out + (int64_t)t * RP.pred_tick_stride : nullptr;
            uint8_t* Hk = Hbuf(k);
            tn_fc_rows<A, NTH, RP_BAR_MAIN, true>(P, sB[t % RING].state_self, sB[t % RING].state_drones, pred_out, e0, nenv, Hk,
                                                  Hk + TN_H_BYTES, fcw, fcb, preds, rowbuf, RI, rowbuf2);
        }
        // FC of every stream has read h (barriers inside tn_fc_rows); the next group's epilogue may overwrite it
    }
    if (tid == 0) bulk_wait_read<0>();        // the last tick's row tiles are still being read by the bulk engine
    tc_fence_before();
    tn_sync<RP_BAR_MAIN, NTH>();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static size_t rollout_pair_smem_bytes(const hs_config& c, int NS) {
    return tp_fused_smem_bytes(c) + 2 * (size_t)(rp_slots(NS) - c.history_step) * TN_X_STEP + 32;
}

}  // namespace
