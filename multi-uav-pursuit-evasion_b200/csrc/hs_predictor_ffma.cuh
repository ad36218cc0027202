// hs_predictor_ffma.cuh -- fused TP_net predictor, fp32 FFMA kernels (16-env and 32-env tiles)
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"

namespace {

// =========================================================================================
// Fused trajectory predictor + second half of the observation.
//   pred = tanh(FC(LSTM_64(TP_input)))           omni_drones/learning/mappo.py:572-589
//   state_self / state_drones rows                omni_drones/envs/hide_and_seek/hideandseek.py:834-887
// The reference runs the predictor through cuDNN between two groups of eager ops; here one
// kernel keeps the whole recurrence on chip: a CTA owns TPB_E environments, the 80x256 gate
// matrix [W_ih | W_hh]^T lives in shared memory (80 KB, permuted so that a thread owns the
// i,f,g,o columns of two hidden units), x_t / h_t are broadcast reads, and each thread keeps
// a 4-env x 8-column fp32 accumulator tile in registers (SIMT FFMA; the 1e-4 fp32 parity bar
// rules out the TF32/BF16 tensor-core paths).  The epilogue applies the FC + tanh, forms the
// 35-wide rows and sends both row tiles out with TMA bulk stores.
// =========================================================================================
constexpr int TPB_E = 16;            // envs per tile
constexpr int TP_THREADS = 128;      // thread = (env group of 8, hidden unit j): warp w -> group w>>1, j = (w&1)*32 + lane
constexpr int TP_NE = 8;             // envs per thread -> 8 env x 4 gate accumulators, 32 FFMA per 3 LDS.128
constexpr int TP_HID = 64;
constexpr int TP_WS = 260;           // row pitch of the gate matrix in smem: 256 + 4 keeps float4 alignment and
                                     // spreads the (coalesced-read) staging stores over 8 banks instead of 1

struct TPParams {
    const float* w_ih;   // [256, FD]   gate order i,f,g,o (torch.nn.LSTM)
    const float* w_hh;   // [256, 64]
    const float* b_ih;   // [256]
    const float* b_hh;   // [256]
    const float* fc_w;   // [3F, 64]
    const float* fc_b;   // [3F]
    float* pred_out;     // [E, 3F] or null
};

// Shared-memory layouts of the predictor kernel (chosen for conflict-free 128-bit access):
//  * gate matrix row k: column of (gate g, hidden unit j) = j*4 + g, so a thread's four gate
//    weights are one float4 and a warp's LDS.128 is one contiguous 512 B span;
//  * h[j][e] (16 envs per row): row j is rotated by 4*j floats, so that the 32 lanes writing
//    consecutive rows spread over the banks while 8 consecutive envs stay two aligned float4.
__device__ __forceinline__ int tp_hoff(int j, int e) { return j * TPB_E + ((e + 4 * j) & (TPB_E - 1)); }

__device__ __forceinline__ float fex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoidf_(float x) { return frcp(1.0f + fex2(-1.4426950408889634f * x)); }
// One LSTM cell update from the four gate pre-activations with 7 SFU operations instead of 10:
// the four reciprocals 1/(1+e^-i), 1/(1+e^-f), 1/(e^2g+1), 1/(1+e^-o) share ONE rcp of the product of
// their denominators (inputs clamped to +-15, where sigmoid/tanh are saturated to 3e-7, so the
// product stays below 4e32).  Returns h; c is updated in place.
__device__ __forceinline__ float lstm_cell(float zi, float zf, float zg, float zo, float& c) {
    const float L2E = 1.4426950408889634f;
    zi = fminf(fmaxf(zi, -15.f), 15.f); zf = fminf(fmaxf(zf, -15.f), 15.f);
    zg = fminf(fmaxf(zg, -15.f), 15.f); zo = fminf(fmaxf(zo, -15.f), 15.f);
    const float di = 1.0f + fex2(-L2E * zi), df = 1.0f + fex2(-L2E * zf);
    const float dg = 1.0f + fex2(2.0f * L2E * zg), dO = 1.0f + fex2(-L2E * zo);
    const float p1 = di * df, p2 = dg * dO;
    const float r = frcp(p1 * p2);
    const float rp2 = r * p2, rp1 = r * p1;
    const float ig = rp2 * df, fg = rp2 * di;            // 1/di, 1/df
    const float gg = 1.0f - 2.0f * (rp1 * dO);           // tanh(zg) = 1 - 2/dg
    const float og = rp1 * dg;                           // 1/do
    c = fmaf(fg, c, ig * gg);
    const float cc = fminf(fmaxf(c, -15.f), 15.f);
    const float th = 1.0f - 2.0f * frcp(1.0f + fex2(2.0f * L2E * cc));
    return og * th;
}
// tanh(x) = 1 - 2/(exp(2x)+1): exact limits at +-inf, abs error ~1e-7 (h and c are O(1))
__device__ __forceinline__ float tanhf_(float x) { return 1.0f - 2.0f * frcp(fex2(2.8853900817779268f * x) + 1.0f); }

template <int A>
__global__ void __launch_bounds__(TP_THREADS, 2)
hs_tp_fill_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(128) float smem[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    constexpr int KTOT = FD + TP_HID;
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int D = 20 + F3;
    const int E = c.num_envs;
    const int tid = threadIdx.x;
    const int ntiles = (E + TPB_E - 1) / TPB_E;

    float* Wp = smem;                               // [KTOT][TP_WS] gate matrix, column j*4+g
    float* bias = Wp + KTOT * TP_WS;                // [256]
    float* fcw = bias + 256;                        // [F3][64]
    float* fcb = fcw + F3 * TP_HID;                 // [F3] (padded to 32)
    float* xs = fcb + 32;                           // [2][FD][TPB_E] double-buffered time step of the input window
    float* hs = xs + 2 * FD * TPB_E;                // [2][64][TPB_E] (rotated rows)
    float* preds = hs + 2 * TP_HID * TPB_E;         // [TPB_E][F3]
    float* rowbuf = xs;                             // [TPB_E*A][D] row staging, aliases xs+hs (dead after the FC)

    // ---- stage the weights once per CTA: linear (coalesced) global reads, transposing smem stores
    for (int i = tid; i < 256 * FD; i += TP_THREADS) {
        const int row = i / FD, k = i - row * FD;
        Wp[k * TP_WS + (row & 63) * 4 + (row >> 6)] = __ldg(W.w_ih + i);
    }
    for (int i = tid; i < 256 * TP_HID / 4; i += TP_THREADS) {       // 16 float4 per row of W_hh
        const int row = i >> 4, k = (i & 15) * 4;
        const float4 w = __ldg(reinterpret_cast<const float4*>(W.w_hh) + i);
        float* d = Wp + (FD + k) * TP_WS + (row & 63) * 4 + (row >> 6);
        d[0] = w.x; d[TP_WS] = w.y; d[2 * TP_WS] = w.z; d[3 * TP_WS] = w.w;
    }
    for (int row = tid; row < 256; row += TP_THREADS)
        bias[(row & 63) * 4 + (row >> 6)] = __ldg(W.b_ih + row) + __ldg(W.b_hh + row);
    for (int i = tid; i < F3 * TP_HID; i += TP_THREADS) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    __syncthreads();

    const int eg = tid >> 6;                                   // env group: envs eg*8 .. eg*8+7
    const int j = ((tid >> 5) & 1) * 32 + (tid & 31);          // hidden unit of this thread
    const float4 bv = *reinterpret_cast<const float4*>(bias + j * 4);

    // ---- persistent loop over 16-env tiles ----------------------------------------------------
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TPB_E;
        const int nenv = (int)min((int64_t)TPB_E, E - e0);
        int64_t xstride;
        const float* xin = tp_window_base(P, e0, H * FD, FD, xstride);
        float cst[TP_NE];
#pragma unroll
        for (int e = 0; e < TP_NE; ++e) cst[e] = 0.f;
        int cur = 0;
        // x_s tile [FD][16] (transposed) arrives by cp.async one time step ahead of its use
        auto fetch_x = [&](int s) {
            float* dst = xs + (s & 1) * FD * TPB_E;
            for (int i = tid; i < TPB_E * FD; i += TP_THREADS) {
                const int e = i / FD, k = i - e * FD;
                if (e < nenv) cp_async4(dst + k * TPB_E + e, xin + (int64_t)e * xstride + s * FD + k);
                else dst[k * TPB_E + e] = 0.0f;
            }
            cp_async_commit();
        };
        fetch_x(0);
        for (int s = 0; s < H; ++s) {
            cp_async_wait_all();
            __syncthreads();             // x_s visible to all; also orders the previous step's h writes
            if (s + 1 < H) fetch_x(s + 1);
            float acc[TP_NE][4];
#pragma unroll
            for (int e = 0; e < TP_NE; ++e) { acc[e][0] = bv.x; acc[e][1] = bv.y; acc[e][2] = bv.z; acc[e][3] = bv.w; }
            // operands of step k+1 are fetched while the 32 FFMAs of step k issue.
            // SWZ: the activation rows are the rotated h rows (tp_hoff); otherwise the plain x rows
            auto mac_block = [&](const float* abase, const float* wrow, int nk, bool swz) {
                auto aoff = [&](int k, int el) { return swz ? tp_hoff(k, el) : (k * TPB_E + el); };
                float4 a0 = *reinterpret_cast<const float4*>(abase + aoff(0, eg * TP_NE));
                float4 a1 = *reinterpret_cast<const float4*>(abase + aoff(0, eg * TP_NE + 4));
                float4 w0 = *reinterpret_cast<const float4*>(wrow);
#pragma unroll 4
                for (int k = 0; k < nk; ++k) {
                    const float ae[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float wq[4] = {w0.x, w0.y, w0.z, w0.w};
                    const int kn = (k + 1 < nk) ? (k + 1) : k;
                    a0 = *reinterpret_cast<const float4*>(abase + aoff(kn, eg * TP_NE));
                    a1 = *reinterpret_cast<const float4*>(abase + aoff(kn, eg * TP_NE + 4));
                    w0 = *reinterpret_cast<const float4*>(wrow + kn * TP_WS);
#pragma unroll
                    for (int e = 0; e < TP_NE; ++e)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[e][q] = fmaf(ae[e], wq[q], acc[e][q]);
                }
            };
            mac_block(xs + (s & 1) * FD * TPB_E, Wp + j * 4, FD, false);
            if (s > 0)                        // h_0 = 0
                mac_block(hs + cur * TP_HID * TPB_E, Wp + FD * TP_WS + j * 4, TP_HID, true);
            float* hnext = hs + (cur ^ 1) * TP_HID * TPB_E;
            float hv[TP_NE];
#pragma unroll
            for (int e = 0; e < TP_NE; ++e) {
                hv[e] = lstm_cell(acc[e][0], acc[e][1], acc[e][2], acc[e][3], cst[e]);
            }
            *reinterpret_cast<float4*>(hnext + tp_hoff(j, eg * TP_NE)) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4*>(hnext + tp_hoff(j, eg * TP_NE + 4)) = make_float4(hv[4], hv[5], hv[6], hv[7]);
            cur ^= 1;
        }
        __syncthreads();                 // h(cur) complete

        // ---- FC + tanh ---------------------------------------------------------------------
        {
            const float* hfin = hs + cur * TP_HID * TPB_E;
            for (int i = tid; i < TPB_E * F3; i += TP_THREADS) {
                const int o = i / TPB_E, e = i - o * TPB_E;
                float a = fcb[o];
#pragma unroll 8
                for (int jj = 0; jj < TP_HID; ++jj) a = fmaf(fcw[o * TP_HID + jj], hfin[tp_hoff(jj, e)], a);
                const float pv = tanhf(a);
                preds[e * F3 + o] = pv;
                if (W.pred_out != nullptr && e < nenv) W.pred_out[(e0 + e) * F3 + o] = pv;
            }
        }
        __syncthreads();

        // ---- rows: thread (a, e) with e fastest -> coalesced arena reads -------------------------
        // state_self and state_drones differ only in their first 3 words (masked / unmasked
        // evader offset): stage the row tile once, store it, patch the heads, store it again.
        V3 t_rpos = mk(0.f, 0.f, 0.f);
        float* r1 = nullptr;
        if (tid < TPB_E * A) {
            const int slot = tid / TPB_E, el = tid - slot * TPB_E;
            const bool valid = el < nenv;
            const int64_t e = valid ? (e0 + el) : (int64_t)(E - 1);
            const V3 p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
            Q4 q; q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
            const V3 lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
            const V3 tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
            const float progress = *EROW(E_PROGRESS);
            const bool bdetect = *EROW(E_BDETECT) != 0.0f;
            V3 heading, up;
            heading_up(q, heading, up);
            const float tfrac = fdiv(progress, (float)c.max_episode_length);
            t_rpos = p - tp;
            const float mv = c.mask_value;
            const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
            r1 = rowbuf + (el * A + slot) * D;
            r1[0] = head_m.x; r1[1] = head_m.y; r1[2] = head_m.z;
            const float* pr = preds + el * F3;
            for (int f = 0; f < c.future_step; ++f) {
                const float px = (pr[3 * f] * 0.5f) * c.arena_size;
                const float py = (pr[3 * f + 1] * 0.5f) * c.arena_size;
                const float pz = ((pr[3 * f + 2] + 1.0f) * 0.5f) * c.max_height;
                r1[3 + 3 * f] = p.x - px; r1[4 + 3 * f] = p.y - py; r1[5 + 3 * f] = p.z - pz;
            }
            const int o = 3 + F3;
            const float tail[17] = {q.w, q.x, q.y, q.z, lv.x, lv.y, lv.z, heading.x, heading.y, heading.z,
                                    up.x, up.y, up.z, tfrac, tfrac, tfrac, tfrac};
#pragma unroll
            for (int i = 0; i < 17; ++i) r1[o + i] = tail[i];
        }
        const int nwords = nenv * A * D;
        float* g1 = P.b.state_self + e0 * A * D;
        float* g2 = P.b.state_drones + e0 * A * D;
        const bool bulk = HS_USE_BULK_STORE && (nenv == TPB_E) && ((nwords & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(g1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g2) & 15) == 0);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            float* gdst = pass == 0 ? g1 : g2;
            if (pass == 1 && r1 != nullptr) { r1[0] = t_rpos.x; r1[1] = t_rpos.y; r1[2] = t_rpos.z; }
            if (bulk) {
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_store(gdst, rowbuf, (uint32_t)nwords * 4u);
                    bulk_commit();
                    bulk_wait_read<0>();     // the tile is patched / reused right after
                }
            } else {
                __syncthreads();
                for (int i = tid; i < nwords; i += TP_THREADS) gdst[i] = rowbuf[i];
            }
            __syncthreads();
        }
    }
}

// ---- large-batch variant: 32-env tiles, 8 env x 8 column register tile (64 FFMA per 4 LDS.128) ----
constexpr int TW_E = 32;            // envs per tile
constexpr int TW_THREADS = 128;      // 4 warps: warp w owns envs 8w..8w+7, lane t owns hidden units 2t, 2t+1
constexpr int TW_NE = 8;             // envs per thread -> 8 x 8 accumulator tile: 64 FFMA per 4 LDS.128


// Shared-memory layouts of the predictor kernel (both chosen for conflict-free 128-bit access):
//  * gate matrix row k: two planes of 128 floats; lane t owns floats [t*4, t*4+4) of each plane,
//    i.e. its 8 columns q = g*2 + u (gate g, hidden unit 2t+u) live at plane q>>2, slot q&3.
//    A warp's LDS.128 of one plane is one contiguous 512 B span.
//  * h[j][e]: row j is rotated by 4*(j>>1) floats, so that the 32 lanes writing rows 2t, 2t+1
//    spread over all banks while 8 consecutive envs stay two aligned float4.
__device__ __forceinline__ int tw_col(int g, int j) {
    const int q = g * 2 + (j & 1);
    return (q >> 2) * 128 + (j >> 1) * 4 + (q & 3);
}
__device__ __forceinline__ int tw_hoff(int j, int e) { return j * TW_E + ((e + 4 * (j >> 1)) & (TW_E - 1)); }


template <int A>
__global__ void __launch_bounds__(TW_THREADS, 2)
hs_tp_fill_wide_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(128) float smem[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    constexpr int KTOT = FD + TP_HID;
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int D = 20 + F3;
    const int E = c.num_envs;
    const int tid = threadIdx.x;
    const int ntiles = (E + TW_E - 1) / TW_E;

    float* Wp = smem;                               // [KTOT][TP_WS] permuted gate matrix
    float* bias = Wp + KTOT * TP_WS;                // [256]
    float* fcw = bias + 256;                        // [F3][64]
    float* fcb = fcw + F3 * TP_HID;                 // [F3] (padded to 32)
    float* xs = fcb + 32;                           // [2][FD][TW_E] double-buffered time step of the input window
    float* hs = xs + 2 * FD * TW_E;                // [2][64][TW_E]
    float* preds = hs + 2 * TP_HID * TW_E;         // [TW_E][F3]
    float* rowbuf = xs;                             // [TW_E*A][D] row staging, aliases xs+hs (dead after the FC)

    // ---- stage the weights once per CTA (conflict-free: consecutive threads -> consecutive smem).
    // column d of (gate g, hidden unit j): lane t = j/2 owns columns t*8 + g*2 + (j&1)
    // global reads are linear (coalesced); the transposing smem stores hit 8 banks (pitch 260)
    for (int i = tid; i < 256 * FD; i += TW_THREADS) {
        const int row = i / FD, k = i - row * FD;
        const int g = row >> 6, j = row & 63;
        Wp[k * TP_WS + tw_col(g, j)] = __ldg(W.w_ih + i);
    }
    for (int i = tid; i < 256 * TP_HID / 4; i += TW_THREADS) {       // 16 float4 per row of W_hh
        const int row = i >> 4, k = (i & 15) * 4;
        const int g = row >> 6, j = row & 63;
        const float4 w = __ldg(reinterpret_cast<const float4*>(W.w_hh) + i);
        float* d = Wp + (FD + k) * TP_WS + tw_col(g, j);
        d[0] = w.x; d[TP_WS] = w.y; d[2 * TP_WS] = w.z; d[3 * TP_WS] = w.w;
    }
    for (int row = tid; row < 256; row += TW_THREADS)
        bias[tw_col(row >> 6, row & 63)] = __ldg(W.b_ih + row) + __ldg(W.b_hh + row);
    for (int i = tid; i < F3 * TP_HID; i += TW_THREADS) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    __syncthreads();

    const int t = tid & 31;              // column group
    const int eg = tid >> 5;             // env group (= warp)
    float bv[8];
    {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + t * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + 128 + t * 4);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    }

    // ---- persistent loop over 32-env tiles ----------------------------------------------------
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TW_E;
        const int nenv = (int)min((int64_t)TW_E, E - e0);
        int64_t xstride;
        const float* xin = tp_window_base(P, e0, H * FD, FD, xstride);
        float cst[TW_NE][2];
#pragma unroll
        for (int e = 0; e < TW_NE; ++e) { cst[e][0] = 0.f; cst[e][1] = 0.f; }
        int cur = 0;
        // x_s tile [FD][32] (transposed) arrives by cp.async one time step ahead of its use
        auto fetch_x = [&](int s) {
            float* dst = xs + (s & 1) * FD * TW_E;
            for (int i = tid; i < TW_E * FD; i += TW_THREADS) {
                const int e = i / FD, k = i - e * FD;
                if (e < nenv) cp_async4(dst + k * TW_E + e, xin + (int64_t)e * xstride + s * FD + k);
                else dst[k * TW_E + e] = 0.0f;
            }
            cp_async_commit();
        };
        fetch_x(0);
        for (int s = 0; s < H; ++s) {
            cp_async_wait_all();
            __syncthreads();             // x_s visible to all; also orders the previous step's h writes
            if (s + 1 < H) fetch_x(s + 1);
            float acc[TW_NE][8];
#pragma unroll
            for (int e = 0; e < TW_NE; ++e)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[e][q] = bv[q];
            // operands of step k+1 are fetched while the 64 FFMAs of step k issue
            // SWZ: the activation rows are the rotated h rows (tw_hoff); otherwise the plain x rows
            auto mac_block = [&](const float* abase, const float* wrow, int nk, bool swz) {
                auto aoff = [&](int k, int e0) { return swz ? tw_hoff(k, e0) : (k * TW_E + e0); };
                float4 a0 = *reinterpret_cast<const float4*>(abase + aoff(0, eg * TW_NE));
                float4 a1 = *reinterpret_cast<const float4*>(abase + aoff(0, eg * TW_NE + 4));
                float4 w0 = *reinterpret_cast<const float4*>(wrow);
                float4 w1 = *reinterpret_cast<const float4*>(wrow + 128);
#pragma unroll 4
                for (int k = 0; k < nk; ++k) {
                    const float ae[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float wq[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    const int kn = (k + 1 < nk) ? (k + 1) : k;
                    a0 = *reinterpret_cast<const float4*>(abase + aoff(kn, eg * TW_NE));
                    a1 = *reinterpret_cast<const float4*>(abase + aoff(kn, eg * TW_NE + 4));
                    w0 = *reinterpret_cast<const float4*>(wrow + kn * TP_WS);
                    w1 = *reinterpret_cast<const float4*>(wrow + kn * TP_WS + 128);
#pragma unroll
                    for (int e = 0; e < TW_NE; ++e)
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[e][q] = fmaf(ae[e], wq[q], acc[e][q]);
                }
            };
            mac_block(xs + (s & 1) * FD * TW_E, Wp + t * 4, FD, false);
            if (s > 0)                        // h_0 = 0
                mac_block(hs + cur * TP_HID * TW_E, Wp + FD * TP_WS + t * 4, TP_HID, true);
            float* hnext = hs + (cur ^ 1) * TP_HID * TW_E;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float hv[TW_NE];
#pragma unroll
                for (int e = 0; e < TW_NE; ++e) {
                    hv[e] = lstm_cell(acc[e][0 + u], acc[e][2 + u], acc[e][4 + u], acc[e][6 + u], cst[e][u]);
                }
                *reinterpret_cast<float4*>(hnext + tw_hoff(2 * t + u, eg * TW_NE)) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(hnext + tw_hoff(2 * t + u, eg * TW_NE + 4)) = make_float4(hv[4], hv[5], hv[6], hv[7]);
            }
            cur ^= 1;
        }
        __syncthreads();                 // h(cur) complete

        // ---- FC + tanh ---------------------------------------------------------------------
        {
            const float* hfin = hs + cur * TP_HID * TW_E;
            for (int i = tid; i < TW_E * F3; i += TW_THREADS) {
                const int o = i / TW_E, e = i - o * TW_E;
                float a = fcb[o];
#pragma unroll 8
                for (int j = 0; j < TP_HID; ++j) a = fmaf(fcw[o * TP_HID + j], hfin[tw_hoff(j, e)], a);
                const float pv = tanhf(a);
                preds[e * F3 + o] = pv;
                if (W.pred_out != nullptr && e < nenv) W.pred_out[(e0 + e) * F3 + o] = pv;
            }
        }
        __syncthreads();

        // ---- rows: thread (a, e) with e fastest -> coalesced arena reads -------------------------
        // state_self and state_drones differ only in their first 3 words (masked / unmasked
        // evader offset): stage the row tile once, store it, patch the heads, store it again.
        V3 t_rpos = mk(0.f, 0.f, 0.f);
        float* r1 = nullptr;
        if (tid < TW_E * A) {
            const int slot = tid / TW_E, el = tid - slot * TW_E;
            const bool valid = el < nenv;
            const int64_t e = valid ? (e0 + el) : (int64_t)(E - 1);
            const V3 p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
            Q4 q; q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
            const V3 lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
            const V3 tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
            const float progress = *EROW(E_PROGRESS);
            const bool bdetect = *EROW(E_BDETECT) != 0.0f;
            V3 heading, up;
            heading_up(q, heading, up);
            const float tfrac = fdiv(progress, (float)c.max_episode_length);
            t_rpos = p - tp;
            const float mv = c.mask_value;
            const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
            r1 = rowbuf + (el * A + slot) * D;
            r1[0] = head_m.x; r1[1] = head_m.y; r1[2] = head_m.z;
            const float* pr = preds + el * F3;
            for (int f = 0; f < c.future_step; ++f) {
                const float px = (pr[3 * f] * 0.5f) * c.arena_size;
                const float py = (pr[3 * f + 1] * 0.5f) * c.arena_size;
                const float pz = ((pr[3 * f + 2] + 1.0f) * 0.5f) * c.max_height;
                r1[3 + 3 * f] = p.x - px; r1[4 + 3 * f] = p.y - py; r1[5 + 3 * f] = p.z - pz;
            }
            const int o = 3 + F3;
            const float tail[17] = {q.w, q.x, q.y, q.z, lv.x, lv.y, lv.z, heading.x, heading.y, heading.z,
                                    up.x, up.y, up.z, tfrac, tfrac, tfrac, tfrac};
#pragma unroll
            for (int i = 0; i < 17; ++i) r1[o + i] = tail[i];
        }
        const int nwords = nenv * A * D;
        float* g1 = P.b.state_self + e0 * A * D;
        float* g2 = P.b.state_drones + e0 * A * D;
        const bool bulk = HS_USE_BULK_STORE && (nenv == TW_E) && ((nwords & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(g1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g2) & 15) == 0);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            float* gdst = pass == 0 ? g1 : g2;
            if (pass == 1 && r1 != nullptr) { r1[0] = t_rpos.x; r1[1] = t_rpos.y; r1[2] = t_rpos.z; }
            if (bulk) {
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_store(gdst, rowbuf, (uint32_t)nwords * 4u);
                    bulk_commit();
                    bulk_wait_read<0>();     // the tile is patched / reused right after
                }
            } else {
                __syncthreads();
                for (int i = tid; i < nwords; i += TW_THREADS) gdst[i] = rowbuf[i];
            }
            __syncthreads();
        }
    }
}

static size_t tp_wide_smem_bytes(const hs_config& c) {
    const int FD = 7 + 3 * c.num_agents, KT = FD + TP_HID, F3 = 3 * c.future_step;
    size_t words = (size_t)KT * TP_WS + 256 + (size_t)F3 * TP_HID + 32 + 2 * (size_t)FD * TW_E + 2 * TP_HID * TW_E +
                   (size_t)TW_E * 3 * FMAX;
    return words * sizeof(float);
}


static size_t tp_smem_bytes(const hs_config& c) {
    const int FD = 7 + 3 * c.num_agents, KT = FD + TP_HID, F3 = 3 * c.future_step;
    size_t words = (size_t)KT * TP_WS + 256 + (size_t)F3 * TP_HID + 32 + 2 * (size_t)FD * TPB_E + 2 * TP_HID * TPB_E +
                   (size_t)TPB_E * 3 * FMAX;      // the row staging tile aliases the x/h region
    return words * sizeof(float);
}


}  // namespace
