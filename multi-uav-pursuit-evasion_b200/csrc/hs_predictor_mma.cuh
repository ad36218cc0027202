// hs_predictor_mma.cuh -- fused TP_net predictor, warp-level mma.sync 3xTF32 kernel
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"
#include "hs_predictor_ffma.cuh"

namespace {

// =========================================================================================
// Tensor-core variant of the fused predictor: the two GEMMs of every LSTM step
// ([envs x 80] x [80 x 256]) run on the tensor pipe as error-compensated TF32 ("3xTF32":
// a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with fp32 accumulation), which keeps fp32-level
// accuracy (the 1e-4 parity bar) at a fraction of the issue slots of the FFMA version.
// Warp-level mma.sync.m16n8k8: operands are staged in shared memory already in FRAGMENT
// order, so every operand fetch is one conflict-free LDS.64/LDS.128:
//   * weights  Wf[kstep][ntile][lane][2]      (b0,b1 of the col-major 8x8 B fragment)
//   * h        Ah{hi,lo}[mtile][kstep][lane][4] (a0..a3 of the 16x8 A fragment), written by the
//              cell-update epilogue directly in fragment order and pre-split into hi/lo
//   * x_t      Ax[mtile][kstep][lane][4] raw fp32 (cp.async, split on the fly)
// Column permutation: n-tile (warp w, pair p, half h) holds, at column 2t+b, gate 2h+b of hidden
// unit w*16+p*4+t, so the thread that owns accumulator pair (c0,c1) of an env row owns i,f (h=0)
// and g,o (h=1) of the same (env, unit): the LSTM cell update is thread-local.
// =========================================================================================
constexpr int TM_THREADS = 128;
constexpr int TM_XK = 2;             // k-steps (of 8) covering the padded input width 16
constexpr int TM_HK = TP_HID / 8;    // 8 k-steps for the hidden part
constexpr int TM_KS = TM_XK + TM_HK;

__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_hi(x);
    lo = tf32_hi(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// position of element (row r of the tile, column c of k-step ks) in an A-fragment array
__device__ __forceinline__ int tm_aidx(int nks, int r, int ks, int c) {
    const int m = r >> 4, rr = r & 15;
    const int lane = (rr & 7) * 4 + (c & 3), elem = (rr >> 3) + 2 * (c >> 2);
    return ((m * nks + ks) * 32 + lane) * 4 + elem;
}

template <int A, int MT>
__global__ void __launch_bounds__(TM_THREADS, 2)
hs_tp_fill_mma_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(128) float smem[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    constexpr int TE = 16 * MT;                     // envs per tile
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int D = 20 + F3;
    const int E = c.num_envs;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntiles = (E + TE - 1) / TE;

    float* Wf = smem;                               // [TM_KS][32 ntiles][32 lanes][2]
    float* bias = Wf + TM_KS * 32 * 64;             // [256] indexed gate*64+unit
    float* fcw = bias + 256;                        // [F3][64]
    float* fcb = fcw + F3 * TP_HID;                 // [32]
    float* Ahi = fcb + 32;                          // [MT][TM_HK][32][4]
    float* Alo = Ahi + MT * TM_HK * 128;
    float* Ax = Alo + MT * TM_HK * 128;             // [2][MT][TM_XK][32][4]
    float* preds = Ax + 2 * MT * TM_XK * 128;       // [TE][F3]
    float* rowbuf = Ahi;                            // [TE*A][D] row staging aliases the (dead) A fragments
    static_assert(2 * MT * TM_HK * 128 + 2 * MT * TM_XK * 128 >= 16 * MT * A * (20 + 3 * FMAX), "row staging must fit");

    // ---- stage weights in fragment order (once per CTA) ---------------------------------------
    for (int i = tid; i < TM_KS * 32 * 64; i += TM_THREADS) Wf[i] = 0.0f;
    __syncthreads();
    auto wf_index = [&](int k, int gate, int unit) {
        const int ks = k >> 3, tt = k & 3, jj = (k & 7) >> 2;
        const int ww = unit >> 4, pp = (unit & 15) >> 2, gg = 2 * (unit & 3) + (gate & 1), hh = gate >> 1;
        const int nt = (ww * 4 + pp) * 2 + hh;
        return ((ks * 32 + nt) * 32 + gg * 4 + tt) * 2 + jj;
    };
    for (int i = tid; i < 256 * FD; i += TM_THREADS) {
        const int row = i / FD, k = i - row * FD;
        Wf[wf_index(k, row >> 6, row & 63)] = __ldg(W.w_ih + i);
    }
    for (int i = tid; i < 256 * TP_HID; i += TM_THREADS) {
        const int row = i >> 6, k = i & 63;
        Wf[wf_index(8 * TM_XK + k, row >> 6, row & 63)] = __ldg(W.w_hh + i);
    }
    for (int row = tid; row < 256; row += TM_THREADS) bias[row] = __ldg(W.b_ih + row) + __ldg(W.b_hh + row);
    for (int i = tid; i < F3 * TP_HID; i += TM_THREADS) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    __syncthreads();

    // bias of this thread's accumulator pairs: pair p -> unit w*16+p*4+t, gates (i,f) and (g,o)
    float bi[4], bf[4], bg[4], bo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int u = w * 16 + p * 4 + t;
        bi[p] = bias[u]; bf[p] = bias[64 + u]; bg[p] = bias[128 + u]; bo[p] = bias[192 + u];
    }

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TE;
        const int nenv = (int)min((int64_t)TE, E - e0);
        int64_t xstride;
        const float* xin = tp_window_base(P, e0, H * FD, FD, xstride);
        float cst[MT][4][2];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int p = 0; p < 4; ++p) { cst[m][p][0] = 0.f; cst[m][p][1] = 0.f; }

        // x_s arrives by cp.async (4 B granules, scattered into fragment order) one step ahead
        auto fetch_x = [&](int s) {
            float* dst = Ax + (s & 1) * MT * TM_XK * 128;
            for (int i = tid; i < TE * 8 * TM_XK; i += TM_THREADS) {
                const int r = i / (8 * TM_XK), k = i - r * (8 * TM_XK);
                float* d = dst + tm_aidx(TM_XK, r, k >> 3, k & 7);
                if (r < nenv && k < FD) cp_async4(d, xin + (int64_t)r * xstride + s * FD + k);
                else *d = 0.0f;
            }
            cp_async_commit();
        };
        fetch_x(0);
        for (int s = 0; s < H; ++s) {
            cp_async_wait_all();
            __syncthreads();                 // x_s landed; h_{s-1} (written below) visible
            if (s + 1 < H) fetch_x(s + 1);
            float acc[MT][8][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[m][2 * p][0] = bi[p]; acc[m][2 * p][1] = bf[p]; acc[m][2 * p][2] = bi[p]; acc[m][2 * p][3] = bf[p];
                    acc[m][2 * p + 1][0] = bg[p]; acc[m][2 * p + 1][1] = bo[p]; acc[m][2 * p + 1][2] = bg[p]; acc[m][2 * p + 1][3] = bo[p];
                }
            const float* ax = Ax + (s & 1) * MT * TM_XK * 128;
            const int nks = (s > 0) ? TM_KS : TM_XK;     // h_0 = 0: skip the hidden part on the first step
#pragma unroll 1
            for (int ks = 0; ks < nks; ++ks) {
                uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    if (ks < TM_XK) {
                        const float4 v = *reinterpret_cast<const float4*>(ax + ((m * TM_XK + ks) * 32 + lane) * 4);
                        tf32_split(v.x, ahi[m][0], alo[m][0]); tf32_split(v.y, ahi[m][1], alo[m][1]);
                        tf32_split(v.z, ahi[m][2], alo[m][2]); tf32_split(v.w, ahi[m][3], alo[m][3]);
                    } else {
                        const int o = ((m * TM_HK + (ks - TM_XK)) * 32 + lane) * 4;
                        const float4 vh = *reinterpret_cast<const float4*>(Ahi + o);
                        const float4 vl = *reinterpret_cast<const float4*>(Alo + o);
                        ahi[m][0] = __float_as_uint(vh.x); ahi[m][1] = __float_as_uint(vh.y);
                        ahi[m][2] = __float_as_uint(vh.z); ahi[m][3] = __float_as_uint(vh.w);
                        alo[m][0] = __float_as_uint(vl.x); alo[m][1] = __float_as_uint(vl.y);
                        alo[m][2] = __float_as_uint(vl.z); alo[m][3] = __float_as_uint(vl.w);
                    }
                }
                const float2* wrow = reinterpret_cast<const float2*>(Wf) + (ks * 32 + w * 8) * 32 + lane;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float2 wv = wrow[n * 32];
                    uint32_t bh0, bl0, bh1, bl1;
                    tf32_split(wv.x, bh0, bl0);
                    tf32_split(wv.y, bh1, bl1);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        mma_tf32(acc[m][n], alo[m], bh0, bh1);      // small terms first
                        mma_tf32(acc[m][n], ahi[m], bl0, bl1);
                        mma_tf32(acc[m][n], ahi[m], bh0, bh1);
                    }
                }
            }
            __syncthreads();                 // every warp has read h_{s-1}: safe to overwrite
            // ---- cell update: thread owns (env rows g, g+8) x (unit w*16+p*4+t) per m-tile ----------
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const float hval = lstm_cell(acc[m][2 * p][2 * rh], acc[m][2 * p][2 * rh + 1], acc[m][2 * p + 1][2 * rh],
                                                     acc[m][2 * p + 1][2 * rh + 1], cst[m][p][rh]);
                        const int u = w * 16 + p * 4 + t;            // hidden unit = k column of the next step
                        const int idx = tm_aidx(TM_HK, m * 16 + g + 8 * rh, u >> 3, u & 7);
                        uint32_t hh, hl;
                        tf32_split(hval, hh, hl);
                        Ahi[idx] = __uint_as_float(hh);
                        Alo[idx] = __uint_as_float(hl);
                    }
        }
        __syncthreads();                     // h_H complete

        // ---- FC + tanh (h = hi + lo) -----------------------------------------------------------
        for (int i = tid; i < TE * F3; i += TM_THREADS) {
            const int o = i / TE, e = i - o * TE;
            float a = fcb[o];
#pragma unroll 8
            for (int jj = 0; jj < TP_HID; ++jj) {
                const int idx = tm_aidx(TM_HK, e, jj >> 3, jj & 7);
                a = fmaf(fcw[o * TP_HID + jj], Ahi[idx] + Alo[idx], a);
            }
            const float pv = tanhf(a);
            preds[e * F3 + o] = pv;
            if (W.pred_out != nullptr && e < nenv) W.pred_out[(e0 + e) * F3 + o] = pv;
        }
        __syncthreads();

        // ---- rows (same as the FFMA variants) -------------------------------------------------------
        V3 t_rpos = mk(0.f, 0.f, 0.f);
        float* r1 = nullptr;
        if (tid < TE * A) {
            const int slot = tid / TE, el = tid - slot * TE;
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
        const bool bulk = HS_USE_BULK_STORE && (nenv == TE) && ((nwords & 3) == 0) &&
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
                    bulk_wait_read<0>();
                }
            } else {
                __syncthreads();
                for (int i = tid; i < nwords; i += TM_THREADS) gdst[i] = rowbuf[i];
            }
            __syncthreads();
        }
    }
}

static size_t tp_mma_smem_bytes(const hs_config& c, int MT) {
    const int F3 = 3 * c.future_step, TE = 16 * MT;
    size_t words = (size_t)TM_KS * 32 * 64 + 256 + (size_t)F3 * TP_HID + 32 + 2 * (size_t)MT * TM_HK * 128 +
                   2 * (size_t)MT * TM_XK * 128 + (size_t)TE * 3 * FMAX;
    return words * sizeof(float);
}


}  // namespace
