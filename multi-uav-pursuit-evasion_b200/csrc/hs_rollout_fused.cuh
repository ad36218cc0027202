// hs_rollout_fused.cuh -- hs_rollout_fused_kernel: T control ticks of one 32-env tile in ONE launch (hs_rollout_fused).
// Part of the single translation unit hs_kernels.cu.
//
// Envs are independent and the predictor's output only enters the observation, never the state: nothing tick t+1
// computes depends on the LSTM of tick t.  A CTA therefore keeps its tile for the whole rollout and runs the two
// halves of the tick as a two-stage pipeline over warp roles:
//   * warps 18-21 (the "tick warps", 8 envs each) run hs_tick_body for tick t+1 - CTBR/PID, rotors, integration, evader,
//     observation, reward, stats; the TP window stays in their shared tile and is shifted in place;
//   * warps 0-17 (16 epilogue warps + 2 MMA-issuing warps) run the tcgen05 predictor of tick t on the window the tick
//     warps built - the same two ping-ponging 16-env halves as hs_tick_tp_fused_kernel - then FC + tanh and the
//     prediction-dependent rows.  The predictor's weights are staged into TMEM ONCE per rollout.
// Hand-off: named barrier 2 "tick t done" (tick warps arrive, predictor warps wait) and named barrier 3 "tile free"
// (predictor warps arrive once they have copied the window into the MMA operand and loaded the new state, tick warps
// wait).  Every tick writes its own buffer table (`sets[(first + t) % num_sets]`: the engine's output sets or the rows
// of the time-major rollout storage), read from device memory into shared memory by the warp that needs it.
// Results are bit-identical to T calls of hs_step_fused (same device functions, same operation order).
#pragma once
#include "hs_predictor_tcgen05.cuh"

namespace {

constexpr int RF_MAIN_THREADS = TCW_THREADS;                    // 576: 16 epilogue warps + 2 issuing warps
constexpr int RF_THREADS = RF_MAIN_THREADS + 32 * FUSED_TICK_WARPS;   // + 4 tick warps = 704
constexpr int RF_BAR_MAIN = 1, RF_BAR_TICK_DONE = 2, RF_BAR_TILE_FREE = 3, RF_BAR_TICKW = 4;

// New TP frame of a tick warp's 8 envs -> one step slot of the predictor's B operand (tf32 hi / lo, the layout of
// tn_stage_x_smem: lane = (env & 7) + 8 * (k & 3) per core matrix), so that the predictor warps do not restage the window.
struct XRingHook {
    static constexpr bool ACTIVE = true;
    uint8_t *xhi, *xlo;              // slot base + (tick warp) * TN_SBO; nullptr: leave the staging to the predictor warps
    int nenv_w;                      // valid envs of this warp
    __device__ __forceinline__ void state(bool, bool, int, int, const V3&, const Q4&, const V3&, const V3&, float, bool) const {}
    template <int FD> __device__ __forceinline__ void frame(const float* tile, int per_env, int keep, int lane) const {
        if (xhi == nullptr) return;
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
};

struct RolloutParams {
    const hs_buffers* sets;          // device memory: [num_sets] buffer tables
    int num_sets, first_set, num_ticks;
    const float* first_tp_prev;      // previous TP window of the first tick (the later ticks keep it in shared memory)
    const float* action;             // [T or 1][E,A,4]
    int64_t action_tick_stride;      // floats between the actions of consecutive ticks (0: the same action every tick)
    float* pred_out;                 // [T or 1][E,3F] or nullptr
    int64_t pred_tick_stride;
};

template <int A, int CT>
__global__ void __launch_bounds__(RF_THREADS, 1)
hs_rollout_fused_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W, const __grid_constant__ RolloutParams RP) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ hs_buffers sB[2];                       // buffer table of tick t at sB[t & 1]
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    constexpr int NTH = RF_MAIN_THREADS;
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int E = c.num_envs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = RP.num_ticks;

    uint8_t* Hhi = smem_raw;
    uint8_t* Hlo = Hhi + TN_H_BYTES;
    float* fcw = reinterpret_cast<float*>(Hlo + TN_H_BYTES);
    float* fcb = fcw + F3 * TP_HID;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(fcb + 32);    // d_ready[2], h_ready[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 4);
    uint8_t* Xhi = reinterpret_cast<uint8_t*>(mbar + 6);
    uint8_t* Xlo = Xhi + (size_t)H * TN_X_STEP;
    float* preds = reinterpret_cast<float*>(Xlo + (size_t)H * TN_X_STEP);
    float* rowbuf = preds + TN_E * 3 * FMAX;
    float* wst = rowbuf + TN_E * A * (20 + 3 * FMAX);
    float* tick_mem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wst + 256 * TN_WPITCH) + 127) & ~(uintptr_t)127);

    if (warp >= NTH / 32) {
        // ================= tick warps: tick t while the predictor warps work on tick t-1 =================
        const int tw = warp - NTH / 32;
        const int ttid = tid - NTH;
        float* m = tick_mem + tw * FUSED_TICK_WORDS;
        const int64_t warp_g = (int64_t)blockIdx.x * FUSED_TICK_WARPS + tw;
        {   // the window before the first tick comes from global memory once; from then on it lives in the warp's tile
            const int64_t ew = warp_g * ENVS_PER_WARP;
            const int nw = (int)max((int64_t)0, min((int64_t)ENVS_PER_WARP, E - ew)) * H * FD;
            float* tile = m + 2 * TICK_STAGE_WORDS;
            for (int i = lane; i < nw; i += 32) tile[i] = RP.first_tp_prev[ew * (H * FD) + i];
            __syncwarp();
        }
        constexpr int TBL_WORDS = (int)(sizeof(hs_buffers) / 4);
        static_assert(TBL_WORDS <= 32 * FUSED_TICK_WARPS, "one table word per tick thread");
        auto table_word = [&](int t) -> uint32_t {                    // this thread's word of tick t's buffer table
            const uint32_t* src = reinterpret_cast<const uint32_t*>(RP.sets + (RP.first_set + t) % RP.num_sets);
            return (ttid < TBL_WORDS && t < T) ? __ldg(src + ttid) : 0u;
        };
        uint32_t tbl = table_word(0);
        for (int t = 0; t < T; ++t) {
            // sB[t & 1] was last read by the predictor warps for tick t-2, which ended before they released the tile of t-1
            if (t == 8) HS_TSTAMP_AT(10, NTH);
            if (t > 0) asm volatile("bar.sync %0, %1;" :: "n"(RF_BAR_TILE_FREE), "n"(RF_THREADS) : "memory");
            if (t == 8) HS_TSTAMP_AT(11, NTH);
            if (ttid < TBL_WORDS) reinterpret_cast<uint32_t*>(&sB[t & 1])[ttid] = tbl;
            tbl = table_word(t + 1);                                  // in flight during the tick: no global latency at the loop top
            asm volatile("bar.sync %0, %1;" :: "n"(RF_BAR_TICKW), "n"(32 * FUSED_TICK_WARPS) : "memory");
            const float* act = RP.action + (int64_t)t * RP.action_tick_stride;
            // from the second tick on the new frame goes straight into the operand ring: slot (t-1) % H held the oldest
            // frame of the previous window (the predictor warps release the tile only after the MMAs that read it)
            XRingHook hook;
            hook.nenv_w = (int)max((int64_t)0, min((int64_t)ENVS_PER_WARP, E - warp_g * ENVS_PER_WARP));
            hook.xhi = (t > 0) ? Xhi + (size_t)((t - 1) % H) * TN_X_STEP + (size_t)tw * TN_SBO : nullptr;
            hook.xlo = (t > 0) ? Xlo + (size_t)((t - 1) % H) * TN_X_STEP + (size_t)tw * TN_SBO : nullptr;
            hs_tick_body<A, false, CT, true, XRingHook>(P, sB[t & 1], act, warp_g, m, m + TICK_STAGE_WORDS, m + 2 * TICK_STAGE_WORDS,
                                                        m + 2 * TICK_STAGE_WORDS + ENVS_PER_WARP * TP_ENV_WORDS_MAX, hook);
            fence_async_smem();                          // operand ring: generic-proxy stores -> the MMAs' async proxy
            __threadfence_block();
            if (t == 8) HS_TSTAMP_AT(12, NTH);
            asm volatile("bar.arrive %0, %1;" :: "n"(RF_BAR_TICK_DONE), "n"(RF_THREADS) : "memory");
        }
        return;
    }

    // ================= predictor warps =================
    const int row = (warp & 3) * 32 + lane;            // TMEM lane = gate row of both M-tiles
    const int cg = warp >> 2;                          // 0..3: epilogue column group; 4: issuing warps
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar)) : "memory");        // d_ready: one commit per M-tile
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar + 1)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 2)) : "memory");   // h_ready: 16 epilogue warps
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 3)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weights -> TMEM once per rollout, in the shadow of the first tick
    for (int i = tid; i < F3 * TP_HID; i += NTH) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    tn_stage_weights_g2s<FD>(W, wst, tid, NTH);
    tc_fence_before();
    tn_sync<RF_BAR_MAIN, NTH>();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    tn_stage_weights_s2t(wst, lane_base, row, cg);
    const TnLane L = tn_lane_consts(W, row);
    tc_fence_before();
    tn_sync<RF_BAR_MAIN, NTH>();
    tc_fence_after();

    const uint32_t d_ready = smem_u32(mbar), h_ready = smem_u32(mbar + 2);
    uint32_t ph_d = 0u, ph_h = 0u;
    const uint32_t warp_u = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const bool issuer = warp_u >= TN_THREADS / 32;
    const uint32_t mytl = warp_u & 1u;                         // M-tile of an issuing warp (warps 16, 17)
    TnIssueHalf I;
    I.aA_hi = tmem_u + TN_COL_A + 160 * mytl;
    I.aA_lo = I.aA_hi + 80;
    I.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_mine = tmem_u + mytl * TN_E;
    const int64_t e0 = (int64_t)blockIdx.x * TN_E;
    const int nenv = (int)min((int64_t)TN_E, E - e0);

    for (int t = 0; t < T; ++t) {
        if (t == 8) HS_TSTAMP_AT(0, 0);
        if (t == 9) HS_TSTAMP_AT(6, 0);
        asm volatile("bar.sync %0, %1;" :: "n"(RF_BAR_TICK_DONE), "n"(RF_THREADS) : "memory");   // tick t is complete
        if (t == 8) HS_TSTAMP_AT(1, 0);
        const TnRowIn RI = tn_row_load<A>(P, e0, nenv);          // new state of the tile (written by the tick warps)
        // step s of this tick's window sits in operand slot (head + s) % H: tick 0 stages the whole window, afterwards the
        // tick warps replace the oldest slot by the new frame
        const int head = t % H;
        if (t == 0) tn_stage_x_smem<FD, NTH>(tick_mem, nenv, H, Xhi, Xlo);
        float* const state_self = sB[t & 1].state_self;
        float* const state_drones = sB[t & 1].state_drones;
        fence_async_smem();
        tc_fence_before();
        __threadfence_block();
        // "tile free" (the tick warps may start tick t+1): they overwrite the arena rows loaded above, the TP tile and the
        // operand slot of this window's OLDEST frame - the epilogue warps therefore arrive only once the step-0 MMAs of both
        // halves have completed (below); the issuing warps have nothing to protect
        if (issuer && t + 1 < T) asm volatile("bar.arrive %0, %1;" :: "n"(RF_BAR_TILE_FREE), "n"(RF_THREADS) : "memory");
        tn_sync<RF_BAR_MAIN, NTH>();
        if (t == 8) HS_TSTAMP_AT(2, 0);
        // ---- the recurrence (hs_tick_tp_fused_kernel): two 16-env halves ping-pong between tensor pipe and epilogue
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                auto xdesc = [&](int hf, int s, bool lo) {
                    const int sl = (head + s >= H) ? head + s - H : head + s;
                    return tc_desc(smem_u32(lo ? Xlo : Xhi) + (uint32_t)sl * TN_X_STEP + (uint32_t)hf * 2u * TN_SBO, TN_X_LBO, TN_SBO);
                };
                auto hdesc = [&](int hf, bool lo) { return tc_desc(smem_u32(lo ? Hlo : Hhi) + (uint32_t)hf * 2u * TN_SBO, TN_H_LBO, TN_SBO); };
                for (int hf = 0; hf < 2; ++hf) {
                    I.x_part(d_mine + 16u * (uint32_t)hf, xdesc(hf, 0, false), xdesc(hf, 0, true));
                    tc_commit(d_ready + 8u * (uint32_t)hf);
                }
                for (int s = 0; s < H; ++s)
                    for (int hf = 0; hf < 2; ++hf) {
                        mbar_wait_idx(h_ready, (uint32_t)hf, ph_h);
                        HS_TSTAMP_IF(16 + 4 * hf + 8 * (s - 3), t == 8 && (s == 3 || s == 4) && warp_u == 16u);
                        if (s + 1 < H) {
                            tc_fence_after();
                            const uint32_t d = d_mine + 16u * (uint32_t)hf;
                            I.x_part(d, xdesc(hf, s + 1, false), xdesc(hf, s + 1, true));
                            I.h_part(d, hdesc(hf, false), hdesc(hf, true));
                            tc_commit(d_ready + 8u * (uint32_t)hf);
                        }
                        HS_TSTAMP_IF(17 + 4 * hf, t == 8 && s == 3 && warp_u == 16u);
                    }
            }
            __syncwarp();
        } else {
            float cst[2][4];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int j = 0; j < 4; ++j) cst[hf][j] = 0.f;
            for (int s = 0; s < H; ++s) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    mbar_wait_idx(d_ready, (uint32_t)hf, ph_d);
                    HS_TSTAMP_IF(18 + 4 * hf, t == 8 && s == 4 && tid == 0);
                    tc_fence_after();
                    tn_epilogue4(lane_base, L, 16 * hf + 4 * cg, cst[hf], Hhi, Hlo);
                    fence_async_smem();                      // h (generic proxy) -> async proxy of the next MMAs
                    tc_fence_before();
                    __syncwarp();
                    HS_TSTAMP_IF(19 + 4 * hf, t == 8 && s == 4 && tid == 0);
                    if (lane == 0) mbar_arrive(h_ready + 8u * (uint32_t)hf);
                    __syncwarp();
                    if (s == 0 && hf == 1 && t + 1 < T)      // both halves' step-0 MMAs are complete: release the tile
                        asm volatile("bar.arrive %0, %1;" :: "n"(RF_BAR_TILE_FREE), "n"(RF_THREADS) : "memory");
                }
            }
        }
        tn_sync<RF_BAR_MAIN, NTH>();          // all h of the last step written; the issuing warps have consumed every arrival
        if (t == 8) HS_TSTAMP_AT(3, 0);
        float* pred_out = (RP.pred_out != nullptr) ? RP.pred_out + (int64_t)t * RP.pred_tick_stride : nullptr;
        tn_fc_rows<A, NTH, RF_BAR_MAIN, true>(P, state_self, state_drones, pred_out, e0, nenv, Hhi, Hlo, fcw, fcb, preds, rowbuf, RI,
                                        wst);   // second row tile: the weight staging tile is dead after the prologue
        // (no barrier here: the next one every predictor warp meets is the one before the next recurrence)
        if (t == 8) HS_TSTAMP_AT(4, 0);
    }
    if (tid == 0) bulk_wait_read<0>();        // the last tick's row tiles are still being read by the bulk engine
    tc_fence_before();
    tn_sync<RF_BAR_MAIN, NTH>();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static size_t rollout_fused_smem_bytes(const hs_config& c) { return tp_fused_smem_bytes(c); }

}  // namespace
