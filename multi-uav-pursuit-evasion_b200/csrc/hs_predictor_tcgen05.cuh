// hs_predictor_tcgen05.cuh -- fused TP_net predictor on the 5th-gen tensor cores: tcgen05.mma + TMEM (128-env tiles; 32-env tiles with the gates on M, single and ping-pong)
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"
#include "hs_predictor_ffma.cuh"
#include "hs_predictor_mma.cuh"

namespace {

// =========================================================================================
// tcgen05 variant of the fused predictor (Blackwell 5th-gen tensor cores, TMEM accumulators).
// One CTA = 128 envs.  Per LSTM step the gate pre-activations D[128 x 256] live in TMEM and are
// produced by tcgen05.mma.kind::tf32 (M=128, N=256, K=8 per instruction) issued by ONE thread:
//   * B = [W_ih | W_hh]^T as tf32 hi/lo pairs in shared memory (canonical K-major core-matrix
//     layout, no swizzle: 8 rows x 16 B per core matrix, SBO between 8-column groups, LBO
//     between 16 B K-chunks), split once per CTA;
//   * A = [x_t | h_{t-1}] ALSO lives in TMEM (the "TS" form of tcgen05.mma): every thread owns
//     one env = one TMEM lane and writes its row (already split into tf32 hi/lo) with tcgen05.st,
//     so the recurrent operand never touches shared memory;
//   * error-compensated 3xTF32: D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32 accumulation,
//     which keeps the fp32 parity bar;
//   * completion is signalled by tcgen05.commit on an mbarrier; the epilogue (tcgen05.ld, cell
//     update, tcgen05.st of h_t) is thread-local because column n = 4*unit + gate.
// TMEM columns: D [0,256), A_hi [256,336), A_lo [336,416) -> 512 allocated (1 CTA per SM).
// =========================================================================================
constexpr int TC_M = 128;
constexpr int TC_THREADS = 256;                    // 2 threads per env row: each updates half of the hidden units
constexpr int TC_K = 16 + TP_HID;                   // 80, input width padded to 16
constexpr int TC_COL_AHI = 256, TC_COL_ALO = 256 + TC_K;
constexpr uint32_t TC_LBO = 4096, TC_SBO = 128;     // bytes: K-chunk stride / 8-column-group stride
constexpr uint32_t TC_B_BYTES = (TC_K / 4) * TC_LBO; // 81920 per hi or lo copy

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in the barrier unit until the phase completes (or the hint expires)
// instead of polling - in the rollout kernel the polling of 18 waiting warps was 15 % of all issued instructions, taken
// from the tick warps that share the SM (profiles/r2_ncu_rollout_fused_4k.txt, per-line table).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t tc_bdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) |
           (1ull << 46);                              // version 1 (sm_100), no swizzle, base offset 0
}

template <int A>
__global__ void __launch_bounds__(TC_THREADS, 1)
hs_tp_fill_tc_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int D = 20 + F3;
    const int E = c.num_envs;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = (warp & 3) * 32 + (tid & 31);      // env row of the tile = TMEM lane
    const int hf = warp >> 2;                          // which half of the hidden units this thread updates
    const int ntiles = (E + TC_M - 1) / TC_M;

    uint8_t* Bhi = smem_raw;                                   // [K/4][32][8][4] tf32
    uint8_t* Blo = Bhi + TC_B_BYTES;
    float* bias = reinterpret_cast<float*>(Blo + TC_B_BYTES); // [256], column n = unit*4 + gate
    float* fcw = bias + 256;                                   // [F3][64]
    float* fcb = fcw + F3 * TP_HID;                            // [32]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(fcb + 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    float* part = reinterpret_cast<float*>(mbar + 2);          // [2][128][3*FMAX] partial FC sums of the two halves
    float* rowbuf = part + 2 * TC_M * 3 * FMAX;                // [128*A][D]

    // ---- one-time setup: TMEM, barrier, B operand (tf32 hi/lo split, canonical layout) ------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto b_off = [&](int n, int k) { return (uint32_t)((k >> 2) * TC_LBO + (n >> 3) * TC_SBO + (n & 7) * 16 + (k & 3) * 4); };
    for (int i = tid; i < 256 * 16; i += TC_THREADS) {           // input part, zero padded to 16
        const int r = i >> 4, k = i & 15;
        const float wv = (k < FD) ? __ldg(W.w_ih + r * FD + k) : 0.0f;
        const int n = (r & 63) * 4 + (r >> 6);
        uint32_t hi, lo;
        tf32_split(wv, hi, lo);
        *reinterpret_cast<uint32_t*>(Bhi + b_off(n, k)) = hi;
        *reinterpret_cast<uint32_t*>(Blo + b_off(n, k)) = lo;
    }
    for (int i = tid; i < 256 * TP_HID; i += TC_THREADS) {
        const int r = i >> 6, k = 16 + (i & 63);
        const int n = (r & 63) * 4 + (r >> 6);
        uint32_t hi, lo;
        tf32_split(__ldg(W.w_hh + i), hi, lo);
        *reinterpret_cast<uint32_t*>(Bhi + b_off(n, k)) = hi;
        *reinterpret_cast<uint32_t*>(Blo + b_off(n, k)) = lo;
    }
    for (int r = tid; r < 256; r += TC_THREADS)
        bias[(r & 63) * 4 + (r >> 6)] = __ldg(W.b_ih + r) + __ldg(W.b_hh + r);
    for (int i = tid; i < F3 * TP_HID; i += TC_THREADS) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    fence_async_smem();                       // B was written through the generic proxy, the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);     // this warp's 32 TMEM lanes
    const uint32_t bar = smem_u32(mbar);
    const uint64_t dhi = tc_bdesc(smem_u32(Bhi)), dlo = tc_bdesc(smem_u32(Blo));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TC_M;
        const int nenv = (int)min((int64_t)TC_M, E - e0);
        const bool valid = row < nenv;
        const int64_t e = valid ? (e0 + row) : (int64_t)(E - 1);
        float cst[32], hreg[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { cst[j] = 0.f; hreg[j] = 0.f; }
        int64_t xstride;
        const float* xin = tp_window_base(P, e0, H * FD, FD, xstride);
        xin += (e - e0) * xstride;
        float xf[16];
        auto load_x = [&](int s) {
#pragma unroll
            for (int k = 0; k < 16; ++k) xf[k] = (valid && k < FD) ? __ldg(xin + s * FD + k) : 0.0f;
        };
        load_x(0);
        for (int s = 0; s < H; ++s) {
            // ---- A[:, 0:16] <- x_s: the hf=0 thread of the row writes the hi words, its partner the lo words
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t vv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t hi, lo;
                    tf32_split(xf[q * 8 + k], hi, lo);
                    vv[k] = hf ? lo : hi;
                }
                tc_st8(lane_base + (hf ? TC_COL_ALO : TC_COL_AHI) + q * 8, vv);
            }
            tc_wait_st();
            tc_fence_before();
            __syncthreads();
            if (s + 1 < H) load_x(s + 1);                // global latency hides behind the MMAs
            // ---- D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, issued by one thread -----------------------
            if (tid == 0) {
                tc_fence_after();
                const int nk = (s > 0) ? (TC_K / 8) : 2;         // h_0 = 0: input part only on the first step
                uint32_t acc = 0;
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t acol = (pass == 0) ? TC_COL_ALO : TC_COL_AHI;
                    const uint64_t bd = (pass == 1) ? dlo : dhi;
                    for (int j = 0; j < nk; ++j) {
                        tc_mma_ts(tmem, tmem + acol + 8 * j, bd + (uint64_t)((2 * j * TC_LBO) >> 4), idesc, acc);
                        acc = 1;
                    }
                }
                tc_commit(bar);
            }
            {   // wait for the accumulator (bounded spin: a wrong descriptor must not hang the box)
                uint32_t spins = 0;
                while (!mbar_try_wait(bar, phase)) { if (++spins > (1u << 16)) __trap(); }
                phase ^= 1;
            }
            tc_fence_after();
            // ---- epilogue: this thread owns hidden units hf*32 .. hf*32+31 of its env ------------------
            const bool last = (s + 1 == H);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {                 // fully unrolled: cst[] stays in registers
                const int ch = hf * 4 + cc;
                uint32_t v[32];
                tc_ld32(lane_base + ch * 32, v);
                uint32_t hh[8], hl[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias + (ch * 8 + u) * 4);
                    const float hval = lstm_cell(__uint_as_float(v[4 * u]) + b4.x, __uint_as_float(v[4 * u + 1]) + b4.y,
                                                 __uint_as_float(v[4 * u + 2]) + b4.z, __uint_as_float(v[4 * u + 3]) + b4.w,
                                                 cst[cc * 8 + u]);
                    tf32_split(hval, hh[u], hl[u]);
                    hreg[cc * 8 + u] = hval;
                }
                if (!last) {
                    tc_st8(lane_base + TC_COL_AHI + 16 + ch * 8, hh);
                    tc_st8(lane_base + TC_COL_ALO + 16 + ch * 8, hl);
                }
            }
        }
        // ---- FC: each half sums over its 32 hidden units, halves are combined through smem ----------
#pragma unroll 1
        for (int o = 0; o < F3; ++o) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) a = fmaf(fcw[o * TP_HID + hf * 32 + j], hreg[j], a);
            part[(hf * TC_M + row) * (3 * FMAX) + o] = a;
        }
        tc_fence_before();
        __syncthreads();
        // ---- rows: thread (row, hf) builds drone slots hf, hf+2 of its env; the tile leaves in two
        // halves of 64 envs (the staging buffer holds 64 envs) --------------------------------------
        const V3 tpv = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
        const float progress = *EROW(E_PROGRESS);
        const bool bdetect = *EROW(E_BDETECT) != 0.0f;
        const float tfrac = fdiv(progress, (float)c.max_episode_length);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const bool mine = (row >> 6) == half;
            const int rl = row & 63;
            V3 trp[2];
#pragma unroll
            for (int si = 0; si < 2; ++si) {
                const int slot = hf + 2 * si;
                trp[si] = mk(0.f, 0.f, 0.f);
                if (mine && slot < A) {
                    const V3 p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
                    Q4 q; q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
                    const V3 lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
                    V3 heading, up;
                    heading_up(q, heading, up);
                    trp[si] = p - tpv;
                    const float mv = c.mask_value;
                    const V3 head_m = bdetect ? trp[si] : mk(mv, mv, mv);
                    float* r1 = rowbuf + (rl * A + slot) * D;
                    r1[0] = head_m.x; r1[1] = head_m.y; r1[2] = head_m.z;
                    for (int f = 0; f < c.future_step; ++f) {
                        float pr[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int o = 3 * f + k;
                            pr[k] = tanhf(fcb[o] + part[row * (3 * FMAX) + o] + part[(TC_M + row) * (3 * FMAX) + o]);
                            if (W.pred_out != nullptr && valid && slot == 0) W.pred_out[e * F3 + o] = pr[k];
                        }
                        const float px = (pr[0] * 0.5f) * c.arena_size;
                        const float py = (pr[1] * 0.5f) * c.arena_size;
                        const float pz = ((pr[2] + 1.0f) * 0.5f) * c.max_height;
                        r1[3 + 3 * f] = p.x - px; r1[4 + 3 * f] = p.y - py; r1[5 + 3 * f] = p.z - pz;
                    }
                    const int o = 3 + F3;
                    const float tail[17] = {q.w, q.x, q.y, q.z, lv.x, lv.y, lv.z, heading.x, heading.y, heading.z,
                                            up.x, up.y, up.z, tfrac, tfrac, tfrac, tfrac};
#pragma unroll
                    for (int i = 0; i < 17; ++i) r1[o + i] = tail[i];
                }
            }
            const int nen = max(0, min(64, nenv - half * 64));
            const int nwords = nen * A * D;
            float* g1 = P.b.state_self + (e0 + half * 64) * A * D;
            float* g2 = P.b.state_drones + (e0 + half * 64) * A * D;
            const bool bulk = HS_USE_BULK_STORE && (nen == 64) && ((nwords & 3) == 0) &&
                              ((reinterpret_cast<uintptr_t>(g1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g2) & 15) == 0);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float* gdst = pass == 0 ? g1 : g2;
                if (pass == 1 && mine) {
#pragma unroll
                    for (int si = 0; si < 2; ++si) {
                        const int slot = hf + 2 * si;
                        if (slot < A) {
                            float* r1 = rowbuf + (rl * A + slot) * D;
                            r1[0] = trp[si].x; r1[1] = trp[si].y; r1[2] = trp[si].z;
                        }
                    }
                }
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
                    for (int i = tid; i < nwords; i += TC_THREADS) gdst[i] = rowbuf[i];
                }
                __syncthreads();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static size_t tp_tc_smem_bytes(const hs_config& c) {
    const int F3 = 3 * c.future_step;
    return 2 * (size_t)TC_B_BYTES + (256 + (size_t)F3 * TP_HID + 32) * sizeof(float) + 16 +
           (2 * (size_t)TC_M * 3 * FMAX + (size_t)(TC_M / 2) * c.num_agents * (20 + 3 * FMAX)) * sizeof(float);
}

// =========================================================================================
// tcgen05 variant for SMALL batches ("gates on M"): the 128-env tile above leaves most SMs idle
// when a launch has only a few thousand envs (4096 envs = 32 tiles on 148 SMs).  Here the
// product is transposed: D^T[gate row, env] = W[gate row, k] * [x_t | h_{t-1}]^T[k, env], so the
// MMA's M dimension (fixed at 128) carries the 256 gate rows as two M-tiles and the N dimension
// carries the envs - N = 32 envs per CTA, 128 CTAs at 4096 envs.
//   * A = the weights, tf32 hi/lo, constant for the whole launch and RESIDENT IN TMEM (TS form):
//     2 M-tiles x (hi, lo) x 80 k-columns = 320 TMEM columns, written once per CTA by tcgen05.st.
//     (A first version kept them in shared memory: every step then streamed 240 KB of weights
//     through the 128 B/clk shared-memory port, ~1 us per step - see profiles/.)
//     Row l of M-tile 0 is gate (l odd ? f : i) of hidden unit l/2, row l of M-tile 1 is gate
//     (l odd ? o : g) of that unit, so TMEM lanes l, l^1 hold the four gates of one cell and the
//     cell update needs only warp shuffles between neighbouring lanes;
//   * B = [x_t | h_{t-1}] per env, K-major core matrices in shared memory, tf32 hi/lo (1 KB per
//     MMA); x of all H steps is staged once per tile, h is rewritten by the epilogue each step;
//   * D double buffered in TMEM (2 x 2 x 32 columns): the input half of step t+1 (independent of
//     h_t) is issued right after the recurrent half of step t and runs under epilogue t;
//   * two issuing threads, one per M-tile (independent accumulators), descriptors in uniform registers;
//   * same error-compensated 3xTF32 as above (fp32-level results).
// TMEM columns: D [0,128), A(tile, hi|lo) at 128 + 80 * (2 * tile + lo) -> 448 used, 512 allocated.
// =========================================================================================
constexpr int TN_E = 32;                              // envs per tile = MMA N
constexpr int TN_THREADS = 512;                       // 16 warps: 4 per TMEM lane quarter, 8 env columns each
constexpr uint32_t TN_SBO = 128;
constexpr uint32_t TN_X_LBO = 512, TN_X_STEP = 4 * TN_X_LBO;   // x_t: 32 rows x 16 k = 2048 B per (step, hi|lo)
constexpr uint32_t TN_H_LBO = 528;                    // h: K-chunk stride padded by 16 B -> conflict-free epilogue stores
constexpr uint32_t TN_H_BYTES = (TP_HID / 4) * TN_H_LBO;       // 8448 B per hi|lo
constexpr uint32_t TN_COL_A = 4 * TN_E;               // first weight column in TMEM
constexpr int TN_WPITCH = 81;                         // words per row of the weight staging tile

__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ bool elect_one() {          // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- pieces shared by the single-tile kernel (hs_tp_fill_tcn_kernel) and the ping-pong kernel
// (hs_tp_fill_tcw_kernel) -------------------------------------------------------------------------
struct TnLane {                 // per-thread constants of the epilogue
    float bias0, bias1;         // exponent-argument biases of the two gate rows behind this TMEM lane
    float sa, sb;               // second gate = sa + sb / d1: tanh(g) on even lanes (1, -2), sigmoid(o) on odd lanes (0, 1)
    int unit;
    bool odd;
};

// Weights -> TMEM, once per CTA.  (1) coalesced global reads into a staging tile whose row index is
// already the TMEM lane: row (tile*128 + l) = gate (tile ? (l&1 ? o : g) : (l&1 ? f : i)) of unit l/2, pitch
// 81 words (odd -> the row-per-lane reads below are conflict-free);  (2) warp (quarter, part cg) writes
// the 80 k-columns of (M-tile cg>>1, hi|lo = cg&1) of its 32 lanes with tcgen05.st.  The rows are
// PRE-SCALED by the constant of their activation (-log2 e for the sigmoid gates, +2 log2 e for the tanh
// gate), so the accumulator already holds the argument of ex2 in the cell update.
// (1) global -> staging tile, by the `nt` threads numbered t = 0..nt-1
template <int FD>
__device__ __forceinline__ void tn_stage_weights_g2s(const TPParams& W, float* wst, int t, int nt) {
    auto lane_of = [](int wr) { const int g = wr >> 6, u = wr & 63; return (g >> 1) * 128 + 2 * u + (g & 1); };
#pragma unroll 8
    for (int i = t; i < 256 * TP_HID; i += nt) {
        const int wr = i >> 6, k = i & 63;
        wst[lane_of(wr) * TN_WPITCH + 16 + k] = __ldg(W.w_hh + i);
    }
#pragma unroll 8
    for (int i = t; i < 256 * 16; i += nt) {
        const int wr = i >> 4, k = i & 15;
        wst[lane_of(wr) * TN_WPITCH + k] = (k < FD) ? __ldg(W.w_ih + wr * FD + k) : 0.0f;
    }
}
// (2) staging tile -> TMEM, by the 16 warps (quarter = warp & 3, part cg = warp >> 2 < 4)
__device__ __forceinline__ void tn_stage_weights_s2t(const float* wst, uint32_t lane_base, int row, int cg) {
    if (cg >= 4) return;                                    // (a dedicated issuing warp only helps with the copy above)
    const int tl = cg >> 1, want_lo = cg & 1;
    const float L2E = 1.4426950408889634f;
    const float scale = (tl == 1 && !(row & 1)) ? 2.0f * L2E : -L2E;         // tile 1, even lane = gate g (tanh)
    const float* src = wst + (tl * 128 + row) * TN_WPITCH;
    const uint32_t col0 = TN_COL_A + (uint32_t)(80 * cg);
#pragma unroll
    for (int ch = 0; ch < 5; ++ch) {                        // 16 k-columns per tcgen05.st
        uint32_t vv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t hi, lo;
            tf32_split(src[16 * ch + k] * scale, hi, lo);
            vv[k] = want_lo ? lo : hi;
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     :: "r"(lane_base + col0 + 16 * ch), "r"(vv[0]), "r"(vv[1]), "r"(vv[2]), "r"(vv[3]), "r"(vv[4]), "r"(vv[5]),
                        "r"(vv[6]), "r"(vv[7]), "r"(vv[8]), "r"(vv[9]), "r"(vv[10]), "r"(vv[11]), "r"(vv[12]), "r"(vv[13]),
                        "r"(vv[14]), "r"(vv[15]) : "memory");
    }
    tc_wait_st();
}
template <int FD, int NTHREADS = TN_THREADS>
__device__ __forceinline__ void tn_stage_weights(const TPParams& W, float* wst, uint32_t lane_base, int row, int cg) {
    tn_stage_weights_g2s<FD>(W, wst, threadIdx.x, NTHREADS);
    __syncthreads();
    tn_stage_weights_s2t(wst, lane_base, row, cg);
}

__device__ __forceinline__ TnLane tn_lane_consts(const TPParams& W, int row) {
    TnLane L;
    L.unit = row >> 1;
    L.odd = (row & 1) != 0;
    const float L2E = 1.4426950408889634f;
    const int wr0 = (L.odd ? 64 : 0) + L.unit, wr1 = (L.odd ? 192 : 128) + L.unit;
    L.bias0 = -L2E * (__ldg(W.b_ih + wr0) + __ldg(W.b_hh + wr0));
    L.bias1 = (L.odd ? -L2E : 2.0f * L2E) * (__ldg(W.b_ih + wr1) + __ldg(W.b_hh + wr1));
    L.sa = L.odd ? 0.0f : 1.0f;
    L.sb = L.odd ? 1.0f : -2.0f;
    return L;
}

// x of all H steps of one 32-env tile -> B operand (tf32 hi/lo): lane = (env & 7) + 8 * (k & 3) per core
// matrix, so the 32 stores of a warp cover 128 contiguous bytes; loads are issued ten at a time.
template <int FD, int NTHREADS = TN_THREADS>
__device__ __forceinline__ void tn_stage_x_ptr(const float* __restrict__ win, int64_t xstride, int nenv, int H, uint8_t* Xhi, uint8_t* Xlo);
template <int FD, int NTHREADS = TN_THREADS>
__device__ __forceinline__ void tn_stage_x(const KParams& P, int64_t e0, int nenv, int H, uint8_t* Xhi, uint8_t* Xlo) {
    int64_t xstride;
    const float* win = tp_window_base(P, e0, H * FD, FD, xstride);
    tn_stage_x_ptr<FD, NTHREADS>(win, xstride, nenv, H, Xhi, Xlo);
}
// `win`: chronological window of the tile's first env, `xstride` floats between consecutive envs
template <int FD, int NTHREADS>
__device__ __forceinline__ void tn_stage_x_ptr(const float* __restrict__ win, int64_t xstride, int nenv, int H, uint8_t* Xhi, uint8_t* Xlo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = lane & 7, kk = lane >> 3;
    constexpr int NW = NTHREADS / 32, BATCH = 10;
    for (int b0 = warp; b0 < H * 16; b0 += NW * BATCH) {
        float xv[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {                  // DRAM latency paid once per batch
            const int cm = b0 + u * NW;
            const int s = cm >> 4, kc = (cm >> 2) & 3, ng = cm & 3;
            const int n = ng * 8 + rr, k = kc * 4 + kk;
            xv[u] = 0.0f;
            if (cm < H * 16 && n < nenv && k < FD) xv[u] = __ldg(win + n * xstride + s * FD + k);
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int cm = b0 + u * NW;
            if (cm < H * 16) {
                const int s = cm >> 4, kc = (cm >> 2) & 3, ng = cm & 3;
                uint32_t hi, lo;
                tf32_split(xv[u], hi, lo);
                const uint32_t off = s * TN_X_STEP + kc * TN_X_LBO + ng * TN_SBO + rr * 16 + kk * 4;
                *reinterpret_cast<uint32_t*>(Xhi + off) = hi;
                *reinterpret_cast<uint32_t*>(Xlo + off) = lo;
            }
        }
    }
}

// Cell update of one LSTM step for this thread's 8 env columns.  Lane pair (l, l^1) = one hidden unit:
// the even lane holds the ex2 arguments of gates i, g, the odd lane those of f, o and the cell state.
// Per column and lane: 2 ex2 + 1 shared rcp for the two gates; tanh(c) of two columns is split between
// the two lanes (ex2 + rcp each).  h goes to the B operand buffer as tf32 hi/lo.
__device__ __forceinline__ void tn_epilogue(uint32_t d_taddr, const TnLane& L, int cg, float (&cst)[8], uint8_t* Hhi, uint8_t* Hlo) {
    uint32_t v0[8], v1[8];
    tc_ld8_nowait(d_taddr + (uint32_t)(cg * 8), v0);
    tc_ld8_nowait(d_taddr + (uint32_t)(TN_E + cg * 8), v1);
    tc_wait_ld();
    const float T2 = 2.8853900817779268f;              // 2 log2 e
#pragma unroll
    for (int np = 0; np < 4; ++np) {
        float gb[2], cc[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int n = 2 * np + q;
            const float a0 = fminf(__uint_as_float(v0[n]) + L.bias0, 60.f);      // upper clamp only: ex2(-inf) = 0 is fine
            const float a1 = fminf(__uint_as_float(v1[n]) + L.bias1, 60.f);
            const float d0 = 1.0f + fex2(a0), d1 = 1.0f + fex2(a1);
            const float r = frcp(d0 * d1);                                        // <= 2^120: no overflow
            const float ga = r * d1;                       // sigmoid(i) | sigmoid(f)
            gb[q] = fmaf(L.sb, r * d0, L.sa);              // tanh(g) = 1 - 2/d1 | sigmoid(o) = 1/d1
            const float ig = __shfl_xor_sync(0xffffffffu, ga * gb[q], 1);          // even lane: sigmoid(i) * tanh(g)
            cst[n] = fmaf(ga, cst[n], ig);                 // (odd lanes) c = f*c + i*g
            cc[q] = cst[n];
        }
        // tanh(c) of the two columns: the odd lane keeps column 2np, its even partner takes column 2np+1
        const float other = __shfl_xor_sync(0xffffffffu, cc[1], 1);
        const float tin = L.odd ? cc[0] : other;
        const float th = 1.0f - 2.0f * frcp(1.0f + fex2(T2 * tin));               // |c| <= H: no overflow
        const float thb = __shfl_xor_sync(0xffffffffu, th, 1);
        if (L.odd) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float hval = gb[q] * (q == 0 ? th : thb);
                const int n = cg * 8 + 2 * np + q;
                uint32_t hh, hl;
                tf32_split(hval, hh, hl);
                const uint32_t off = (L.unit >> 2) * TN_H_LBO + (n >> 3) * TN_SBO + (n & 7) * 16 + (L.unit & 3) * 4;
                *reinterpret_cast<uint32_t*>(Hhi + off) = hh;
                *reinterpret_cast<uint32_t*>(Hlo + off) = hl;
            }
        }
    }
}

// FC + tanh from the final h (hi + lo in shared memory), then the state_self / state_drones rows of the tile.
#ifdef HS_FUSED_TIMING
__device__ unsigned long long hs_dbg_times[32];
#define HS_TSTAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == (HS_TSTAMP_TID)) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); hs_dbg_times[i] = t_; } } while (0)
#define HS_TSTAMP_TID 0
#define HS_TSTAMP_AT(i, tid_) do { if (blockIdx.x == 0 && threadIdx.x == (tid_)) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); hs_dbg_times[i] = t_; } } while (0)
#define HS_TSTAMP_IF(i, cond_) do { if (blockIdx.x == 0 && (cond_)) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); hs_dbg_times[i] = t_; } } while (0)
#else
#define HS_TSTAMP(i) do {} while (0)
#define HS_TSTAMP_AT(i, tid_) do {} while (0)
#define HS_TSTAMP_IF(i, cond_) do {} while (0)
#endif
// What a row thread (tid < 32 A: env el = tid % 32, pursuer slot = tid / 32) needs from the arena.  The fused kernel
// loads it right after the tick phase so that the L2 round trip is not exposed after the recurrence.
struct TnRowIn {
    V3 p, lv, tp;
    Q4 q;
    float progress;
    bool bdetect;
};
template <int A>
__device__ __forceinline__ TnRowIn tn_row_load(const KParams& P, int64_t e0, int nenv) {
    TnRowIn R;
    R.p = mk(0.f, 0.f, 0.f); R.lv = R.p; R.tp = R.p; R.q.w = 1.f; R.q.x = R.q.y = R.q.z = 0.f; R.progress = 0.f; R.bdetect = false;
    const int tid = threadIdx.x;
    if (tid < TN_E * A) {
        const int slot = tid / TN_E, el = tid - slot * TN_E;
        const bool valid = el < nenv;
        const int64_t e = valid ? (e0 + el) : (int64_t)(P.c.num_envs - 1);
        R.p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
        R.q.w = *DROW(D_ROT); R.q.x = *DROW(D_ROT + 1); R.q.y = *DROW(D_ROT + 2); R.q.z = *DROW(D_ROT + 3);
        R.lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
        R.tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
        R.progress = *EROW(E_PROGRESS);
        R.bdetect = *EROW(E_BDETECT) != 0.0f;
    }
    return R;
}

// SYNC_ID 0: the NTHREADS threads are the whole CTA (__syncthreads); otherwise they meet at named barrier SYNC_ID (the
// rollout kernel, where the tick warps run ahead beside them).
template <int SYNC_ID, int NTHREADS>
__device__ __forceinline__ void tn_sync() {
    if (SYNC_ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "n"(SYNC_ID), "n"(NTHREADS) : "memory");
}
// DEFER (rollout kernel): the bulk stores of the two row tiles are left in flight - the wait for their shared-memory reads
// moves to the next call (before the tiles are rewritten) and, after the last tick, to the caller.
template <int A, int NTHREADS = TN_THREADS, int SYNC_ID = 0, bool DEFER = false>
__device__ __forceinline__ void tn_fc_rows(const KParams& P, float* state_self, float* state_drones, float* pred_out,
                                           int64_t e0, int nenv, const uint8_t* Hhi,
                                           const uint8_t* Hlo, const float* fcw, const float* fcb, float* preds, float* rowbuf,
                                           const TnRowIn& RI, float* rowbuf2 = nullptr) {
    const hs_config& c = P.c;
    const int F3 = 3 * c.future_step, D = 20 + F3;
    const int tid = threadIdx.x;
    {
        const int n = tid & 31;
        for (int og = tid >> 5; og < F3; og += NTHREADS / 32) {
            float a0 = fcb[og];
            const float* w0 = fcw + og * TP_HID;
#pragma unroll
            for (int kc = 0; kc < TP_HID / 4; ++kc) {          // fully unrolled: the 32 LDS.128 are issued ahead of the FMA chain
                const float4 hh = *reinterpret_cast<const float4*>(Hhi + kc * TN_H_LBO + n * 16);
                const float4 hl = *reinterpret_cast<const float4*>(Hlo + kc * TN_H_LBO + n * 16);
                const float4 ww = *reinterpret_cast<const float4*>(w0 + 4 * kc);
                a0 = fmaf(ww.x, hh.x + hl.x, a0); a0 = fmaf(ww.y, hh.y + hl.y, a0);
                a0 = fmaf(ww.z, hh.z + hl.z, a0); a0 = fmaf(ww.w, hh.w + hl.w, a0);
            }
            const float pv = tanhf(a0);
            preds[n * F3 + og] = pv;
            if (pred_out != nullptr && n < nenv) pred_out[(e0 + n) * F3 + og] = pv;
        }
    }
    if (DEFER && tid == 0) bulk_wait_read<0>();               // the previous call's row tiles have left shared memory
    tn_sync<SYNC_ID, NTHREADS>();
    HS_TSTAMP(21);
    const int nwords = nenv * A * D;
    float* g1 = state_self + e0 * A * D;
    float* g2 = state_drones + e0 * A * D;
    const bool bulk = HS_USE_BULK_STORE && (nenv == TN_E) && ((nwords & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(g1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g2) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(rowbuf) & 15) == 0);
    // both tensors in one pass: state_drones = the same rows with the unmasked target offset, built in a second tile
    const bool both = rowbuf2 != nullptr && bulk && ((reinterpret_cast<uintptr_t>(rowbuf2) & 15) == 0);
    V3 t_rpos = mk(0.f, 0.f, 0.f);
    float* r1 = nullptr;
    if (tid < TN_E * A) {
        const int slot = tid / TN_E, el = tid - slot * TN_E;
        const V3 p = RI.p, lv = RI.lv, tp = RI.tp;
        const Q4 q = RI.q;
        const float progress = RI.progress;
        const bool bdetect = RI.bdetect;
        V3 heading, up;
        heading_up(q, heading, up);
        const float tfrac = fdiv(progress, (float)c.max_episode_length);
        t_rpos = p - tp;
        const float mv = c.mask_value;
        const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
        r1 = rowbuf + (el * A + slot) * D;
        float* r2 = both ? rowbuf2 + (el * A + slot) * D : r1;           // (!both: every store below lands in r1 twice)
        r1[0] = head_m.x; r1[1] = head_m.y; r1[2] = head_m.z;
        if (both) { r2[0] = t_rpos.x; r2[1] = t_rpos.y; r2[2] = t_rpos.z; }
        const float* pr = preds + el * F3;
        for (int f = 0; f < c.future_step; ++f) {
            const float px = (pr[3 * f] * 0.5f) * c.arena_size;
            const float py = (pr[3 * f + 1] * 0.5f) * c.arena_size;
            const float pz = ((pr[3 * f + 2] + 1.0f) * 0.5f) * c.max_height;
            const float dx = p.x - px, dy = p.y - py, dz = p.z - pz;
            r1[3 + 3 * f] = dx; r1[4 + 3 * f] = dy; r1[5 + 3 * f] = dz;
            r2[3 + 3 * f] = dx; r2[4 + 3 * f] = dy; r2[5 + 3 * f] = dz;
        }
        const int o = 3 + F3;
        const float tail[17] = {q.w, q.x, q.y, q.z, lv.x, lv.y, lv.z, heading.x, heading.y, heading.z,
                                up.x, up.y, up.z, tfrac, tfrac, tfrac, tfrac};
#pragma unroll
        for (int i = 0; i < 17; ++i) { r1[o + i] = tail[i]; r2[o + i] = tail[i]; }
    }
    HS_TSTAMP(22);
    if (both) {
        fence_async_smem();
        tn_sync<SYNC_ID, NTHREADS>();
        if (tid == 0) {
            bulk_store(g1, rowbuf, (uint32_t)nwords * 4u);
            bulk_store(g2, rowbuf2, (uint32_t)nwords * 4u);
            bulk_commit();
            if (!DEFER) bulk_wait_read<0>();
        }
        if (!DEFER) tn_sync<SYNC_ID, NTHREADS>();
        return;
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float* gdst = pass == 0 ? g1 : g2;
        if (pass == 1 && r1 != nullptr) { r1[0] = t_rpos.x; r1[1] = t_rpos.y; r1[2] = t_rpos.z; }
        if (bulk) {
            fence_async_smem();
            tn_sync<SYNC_ID, NTHREADS>();
            if (tid == 0) {
                bulk_store(gdst, rowbuf, (uint32_t)nwords * 4u);
                bulk_commit();
                bulk_wait_read<0>();
            }
        } else {
            tn_sync<SYNC_ID, NTHREADS>();
            for (int i = tid; i < nwords; i += NTHREADS) gdst[i] = rowbuf[i];
        }
        tn_sync<SYNC_ID, NTHREADS>();
    }
}

// MMA issue helpers: one elected thread per M-tile; all operands warp-uniform, every descriptor is base + immediate.
struct TnIssue {
    uint32_t aA_hi, aA_lo;      // TMEM column addresses of this issuer's weight tile (hi, lo)
    uint32_t idesc;
    __device__ __forceinline__ void x_part(uint32_t d, uint64_t dX_hi, uint64_t dX_lo, uint32_t first_acc) const {   // 6 MMAs
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)                // small terms first: A_lo*B_hi, A_hi*B_lo, A_hi*B_hi
#pragma unroll
            for (int j = 0; j < 2; ++j)
                tc_mma_ts(d, ((pass == 0) ? aA_lo : aA_hi) + 8 * j,
                          ((pass == 1) ? dX_lo : dX_hi) + (uint64_t)((2 * j * TN_X_LBO) >> 4), idesc, (pass | j) ? 1u : first_acc);
    }
    __device__ __forceinline__ void h_part(uint32_t d, uint64_t dH_hi, uint64_t dH_lo) const {                         // 24 MMAs
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int j = 0; j < TP_HID / 8; ++j)
                tc_mma_ts(d, ((pass == 0) ? aA_lo : aA_hi) + 16 + 8 * j,
                          ((pass == 1) ? dH_lo : dH_hi) + (uint64_t)((2 * j * TN_H_LBO) >> 4), idesc, 1u);
    }
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t& phase) {
    uint32_t spins = 0;                                     // bounded: a wrong descriptor must not hang the box
    while (!mbar_try_wait(bar, phase)) { if (++spins > (1u << 16)) __trap(); }
    phase ^= 1;
}
// barrier `idx` of an array of mbarriers; `bits` holds one phase bit per barrier (no dynamically indexed registers)
__device__ __forceinline__ void mbar_wait_idx(uint32_t bar0, uint32_t idx, uint32_t& bits) {
    uint32_t spins = 0;
    // (polling is 17 % of the rollout kernels' issued instructions, profiles/r2_ncu_rollout_pair_4k.txt, but it is not what
    // bounds them: a __nanosleep(40) back-off here changed nothing - measured, 15.7 us per tick either way)
    while (!mbar_try_wait(bar0 + 8u * idx, (bits >> idx) & 1u)) { if (++spins > (1u << 22)) __trap(); }
    bits ^= 1u << idx;
}

constexpr int TCW_THREADS = TN_THREADS + 64;            // 16 epilogue warps + 2 issuing warps (one per M-tile)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

template <int A>
__global__ void __launch_bounds__(TN_THREADS, 1)
hs_tp_fill_tcn_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int E = c.num_envs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;            // TMEM lane = gate row of both M-tiles
    const int cg = warp >> 2;                          // env columns [8*cg, 8*cg+8) of the tile
    const int ntiles = (E + TN_E - 1) / TN_E;

    uint8_t* Hhi = smem_raw;                                   // [16 K chunks (528 B)][4][8][4]
    uint8_t* Hlo = Hhi + TN_H_BYTES;
    float* fcw = reinterpret_cast<float*>(Hlo + TN_H_BYTES);   // [F3][64]
    float* fcb = fcw + F3 * TP_HID;                            // [32]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(fcb + 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    uint8_t* Xhi = reinterpret_cast<uint8_t*>(mbar + 2);       // [H][4 K chunks][4][8][4]
    uint8_t* Xlo = Xhi + (size_t)H * TN_X_STEP;
    float* preds = reinterpret_cast<float*>(Xlo + (size_t)H * TN_X_STEP);   // [32][3F]
    float* rowbuf = preds + TN_E * 3 * FMAX;                   // [32*A][D]
    float* wst = rowbuf + TN_E * A * (20 + 3 * FMAX);          // [256][81] weight staging (prologue only)

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar)) : "memory");    // two issuing threads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < F3 * TP_HID; i += TN_THREADS) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    const TnLane L = tn_lane_consts(W, row);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    tn_stage_weights<FD>(W, wst, lane_base, row, cg);
    const uint32_t bar = smem_u32(mbar);
    uint32_t phase = 0;

    // (warp index and TMEM base are made provably warp-uniform so that the descriptors live in uniform registers)
    const uint32_t warp_u = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const bool issue_warp = warp_u < 2;
    const uint32_t mytl = warp_u & 1u;
    TnIssue I;
    I.aA_hi = tmem_u + TN_COL_A + 160 * mytl;
    I.aA_lo = I.aA_hi + 80;
    I.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN_E >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t dH_hi = tc_desc(smem_u32(Hhi), TN_H_LBO, TN_SBO), dH_lo = tc_desc(smem_u32(Hlo), TN_H_LBO, TN_SBO);
    const uint64_t dX_hi = tc_desc(smem_u32(Xhi), TN_X_LBO, TN_SBO), dX_lo = tc_desc(smem_u32(Xlo), TN_X_LBO, TN_SBO);
    const uint32_t d_mine = tmem_u + mytl * TN_E;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TN_E;
        const int nenv = (int)min((int64_t)TN_E, E - e0);
        tn_stage_x<FD>(P, e0, nenv, H, Xhi, Xlo);
        float cst[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cst[j] = 0.f;
        fence_async_smem();                   // generic-proxy writes (x) -> visible to the MMA's async proxy
        tc_fence_before();                    // (first tile: also orders the tcgen05.st of the weights)
        __syncthreads();
        if (issue_warp && elect_one()) {
            tc_fence_after();
            I.x_part(d_mine, dX_hi, dX_lo, 0u);
        }
        for (int s = 0; s < H; ++s) {
            const int dbuf = s & 1;
            if (issue_warp && elect_one()) {
                if (s > 0) {
                    tc_fence_after();
                    I.h_part(d_mine + (uint32_t)(dbuf * 2 * TN_E), dH_hi, dH_lo);   // += W_hh * h_{s-1}
                }
                tc_commit(bar);
                if (s + 1 < H) {                         // input half of the next step, under this epilogue
                    const uint64_t xo = (uint64_t)(((uint32_t)(s + 1) * TN_X_STEP) >> 4);
                    I.x_part(d_mine + (uint32_t)((dbuf ^ 1) * 2 * TN_E), dX_hi + xo, dX_lo + xo, 0u);
                }
            }
            mbar_wait(bar, phase);
            tc_fence_after();
            tn_epilogue(lane_base + (uint32_t)(dbuf * 2 * TN_E), L, cg, cst, Hhi, Hlo);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
        tn_fc_rows<A>(P, P.b.state_self, P.b.state_drones, W.pred_out, e0, nenv, Hhi, Hlo, fcw, fcb, preds, rowbuf, tn_row_load<A>(P, e0, nenv));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

constexpr size_t HS_MAX_DYN_SMEM = 232448;           // 227 KB per CTA on sm_100
static size_t tp_tcn_smem_bytes(const hs_config& c) {
    const int F3 = 3 * c.future_step;
    return 2 * (size_t)TN_H_BYTES + ((size_t)F3 * TP_HID + 32) * sizeof(float) + 16 + 2 * (size_t)c.history_step * TN_X_STEP +
           ((size_t)TN_E * 3 * FMAX + (size_t)TN_E * c.num_agents * (20 + 3 * FMAX) + (size_t)256 * TN_WPITCH) * sizeof(float);
}

// =========================================================================================
// Fused tick + predictor for batches of at most one 32-env tile per SM (the BASELINE workload: 4096 envs = 128
// CTAs): ONE launch per control tick.  Warps 0-3 of the CTA run the control tick of the tile's 32 envs
// (hs_tick_body, 8 envs per warp) while warps 4-15 bring the predictor's weights from global memory into the
// staging tile; after one block barrier the CTA is the small-batch tcgen05 predictor above, with the LSTM input
// taken from the TP_input tiles the tick warps just built in shared memory instead of from global memory.
// Against the two-kernel sequence this removes a launch boundary (~3 us of drain + ramp at this size), hides the
// weight prologue behind the tick, and skips the global round trip of the window.  Results are bit-identical to
// hs_tick_kernel followed by hs_tp_fill_tcn_kernel (same device functions, same operation order).
// =========================================================================================
constexpr int FUSED_TICK_WARPS = TN_E / ENVS_PER_WARP;                       // 4

constexpr int FUSED_TICK_WORDS = 2 * TICK_STAGE_WORDS + ENVS_PER_WARP * TP_ENV_WORDS_MAX + TICK_STAT_WORDS;   // per tick warp

// x of all H steps from the tick warps' shared TP_input tiles ([8 envs][H][FD] per warp) -> B operand (tf32 hi/lo)
template <int FD, int NTHREADS>
__device__ __forceinline__ void tn_stage_x_smem(const float* tick_mem, int nenv, int H, uint8_t* Xhi, uint8_t* Xlo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = lane & 7, kk = lane >> 3;
    constexpr int NW = NTHREADS / 32;
    for (int cm = warp; cm < H * 16; cm += NW) {
        const int s = cm >> 4, kc = (cm >> 2) & 3, ng = cm & 3;
        const int n = ng * 8 + rr, k = kc * 4 + kk;
        // env n lives in tick warp n / 8 = ng, row rr of its tile
        const float* tile = tick_mem + ng * FUSED_TICK_WORDS + 2 * TICK_STAGE_WORDS;
        const float xv = (n < nenv && k < FD) ? tile[rr * (H * FD) + s * FD + k] : 0.0f;
        uint32_t hi, lo;
        tf32_split(xv, hi, lo);
        const uint32_t off = s * TN_X_STEP + kc * TN_X_LBO + ng * TN_SBO + rr * 16 + kk * 4;
        *reinterpret_cast<uint32_t*>(Xhi + off) = hi;
        *reinterpret_cast<uint32_t*>(Xlo + off) = lo;
    }
}

// Cell update for FOUR env columns [n0, n0+4) of the tile (the fused kernel splits the 32-env tile into two 16-env
// halves and spreads each half over all 16 epilogue warps); same arithmetic, lane pairing and h layout as tn_epilogue.
__device__ __forceinline__ void tn_epilogue4(uint32_t d_taddr, const TnLane& L, int n0, float (&cst)[4], uint8_t* Hhi, uint8_t* Hlo) {
    uint32_t v0[4], v1[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]) : "r"(d_taddr + (uint32_t)n0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3]) : "r"(d_taddr + (uint32_t)(TN_E + n0)));
    tc_wait_ld();
    const float T2 = 2.8853900817779268f;              // 2 log2 e
#pragma unroll
    for (int np = 0; np < 2; ++np) {
        float gb[2], cc[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int n = 2 * np + q;
            const float a0 = fminf(__uint_as_float(v0[n]) + L.bias0, 60.f);
            const float a1 = fminf(__uint_as_float(v1[n]) + L.bias1, 60.f);
            const float d0 = 1.0f + fex2(a0), d1 = 1.0f + fex2(a1);
            const float r = frcp(d0 * d1);
            const float ga = r * d1;
            gb[q] = fmaf(L.sb, r * d0, L.sa);
            const float ig = __shfl_xor_sync(0xffffffffu, ga * gb[q], 1);
            cst[n] = fmaf(ga, cst[n], ig);
            cc[q] = cst[n];
        }
        const float other = __shfl_xor_sync(0xffffffffu, cc[1], 1);
        const float tin = L.odd ? cc[0] : other;
        const float th = 1.0f - 2.0f * frcp(1.0f + fex2(T2 * tin));
        const float thb = __shfl_xor_sync(0xffffffffu, th, 1);
        if (L.odd) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float hval = gb[q] * (q == 0 ? th : thb);
                const int n = n0 + 2 * np + q;
                uint32_t hh, hl;
                tf32_split(hval, hh, hl);
                const uint32_t off = (L.unit >> 2) * TN_H_LBO + (n >> 3) * TN_SBO + (n & 7) * 16 + (L.unit & 3) * 4;
                *reinterpret_cast<uint32_t*>(Hhi + off) = hh;
                *reinterpret_cast<uint32_t*>(Hlo + off) = hl;
            }
        }
    }
}

// MMA issue for one 16-env half: same operand layouts as the 32-env tile (a half is two of its four 8-env core-matrix
// groups, i.e. the descriptors move by 2 * SBO and the accumulator by 16 columns), N = 16 in the instruction descriptor.
struct TnIssueHalf {
    uint32_t aA_hi, aA_lo, idesc;
    __device__ __forceinline__ void x_part(uint32_t d, uint64_t dX_hi, uint64_t dX_lo) const {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                tc_mma_ts(d, ((pass == 0) ? aA_lo : aA_hi) + 8 * j,
                          ((pass == 1) ? dX_lo : dX_hi) + (uint64_t)((2 * j * TN_X_LBO) >> 4), idesc, (pass | j) ? 1u : 0u);
    }
    __device__ __forceinline__ void h_part(uint32_t d, uint64_t dH_hi, uint64_t dH_lo) const {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int j = 0; j < TP_HID / 8; ++j)
                tc_mma_ts(d, ((pass == 0) ? aA_lo : aA_hi) + 16 + 8 * j,
                          ((pass == 1) ? dH_lo : dH_hi) + (uint64_t)((2 * j * TN_H_LBO) >> 4), idesc, 1u);
    }
};

// WITH_TICK = false: the same predictor as a kernel of its own (hs_step_post_tp variant 5, the default for small
// batches): all warps stage the weights, the LSTM input comes from TP_input in global memory.
template <int A, int CT, bool WITH_TICK = true>
__global__ void __launch_bounds__(TCW_THREADS, 1)
hs_tick_tp_fused_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    constexpr int NTH = TCW_THREADS;                   // 16 tick/epilogue warps + 2 issuing warps
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int E = c.num_envs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;            // TMEM lane = gate row of both M-tiles
    const int cg = warp >> 2;                          // 0..3: epilogue column group; 4: issuing warps

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
    // tick pieces of the four tick warps, 128-byte aligned
    float* tick_mem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wst + 256 * TN_WPITCH) + 127) & ~(uintptr_t)127);

    HS_TSTAMP(0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar)) : "memory");        // d_ready: one commit per M-tile
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar + 1)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 2)) : "memory");   // h_ready: 16 epilogue warps
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 3)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!WITH_TICK) {
        if (warp == FUSED_TICK_WARPS) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        for (int i = tid; i < F3 * TP_HID; i += NTH) fcw[i] = __ldg(W.fc_w + i);
        if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
        tn_stage_weights_g2s<FD>(W, wst, tid, NTH);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tn_stage_weights_s2t(wst, *tmem_slot + ((uint32_t)((warp & 3) * 32) << 16), row, cg);
    } else if (warp < FUSED_TICK_WARPS) {
        // ---- phase 1a: the control tick of this tile's envs, one warp per 8 envs
        float* m = tick_mem + warp * FUSED_TICK_WORDS;
        hs_tick_body<A, false, CT>(P, P.b, P.action, (int64_t)blockIdx.x * FUSED_TICK_WARPS + warp, m, m + TICK_STAGE_WORDS,
                                   m + 2 * TICK_STAGE_WORDS, m + 2 * TICK_STAGE_WORDS + ENVS_PER_WARP * TP_ENV_WORDS_MAX);
    } else {
        // ---- phase 1b, in the shadow of the tick: TMEM allocation (warp 4), predictor constants and weights
        // global -> shared -> TMEM.  These warps meet at named barrier 1; a warp can only write the TMEM lane
        // quarter (warp & 3), so the part the tick warps would own (cg 0) is taken by warps 4-7 as a second pass.
        if (warp == FUSED_TICK_WARPS) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        const int t = tid - 32 * FUSED_TICK_WARPS;
        constexpr int nt = NTH - 32 * FUSED_TICK_WARPS;
        for (int i = t; i < F3 * TP_HID; i += nt) fcw[i] = __ldg(W.fc_w + i);
        if (t < F3) fcb[t] = __ldg(W.fc_b + t);
        tn_stage_weights_g2s<FD>(W, wst, t, nt);
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" :: "n"(nt) : "memory");
        tc_fence_after();
        const uint32_t lane_base_w = *tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
        tn_stage_weights_s2t(wst, lane_base_w, row, cg);
        if (cg == 1) tn_stage_weights_s2t(wst, lane_base_w, row, 0);
    }
    HS_TSTAMP(1);
    const TnLane L = tn_lane_consts(W, row);
    tc_fence_before();
    __syncthreads();
    HS_TSTAMP(2);
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
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
    const TnRowIn RI = tn_row_load<A>(P, e0, nenv);          // new state of the tile (written by the tick warps above)
    if (WITH_TICK) tn_stage_x_smem<FD, NTH>(tick_mem, nenv, H, Xhi, Xlo);
    else tn_stage_x<FD, NTH>(P, e0, nenv, H, Xhi, Xlo);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    HS_TSTAMP(4);
    // ---- the recurrence: the tile's two 16-env halves ping-pong - while the tensor pipe works on one half the 16
    // epilogue warps update the other (4 env columns per thread and half); mbarrier hand-off, no block barrier
    if (issuer) {
        if (elect_one()) {
            tc_fence_after();
            auto xdesc = [&](int hf, int s, bool lo) {
                return tc_desc(smem_u32(lo ? Xlo : Xhi) + (uint32_t)s * TN_X_STEP + (uint32_t)hf * 2u * TN_SBO, TN_X_LBO, TN_SBO);
            };
            auto hdesc = [&](int hf, bool lo) { return tc_desc(smem_u32(lo ? Hlo : Hhi) + (uint32_t)hf * 2u * TN_SBO, TN_H_LBO, TN_SBO); };
            for (int hf = 0; hf < 2; ++hf) {
                I.x_part(d_mine + 16u * (uint32_t)hf, xdesc(hf, 0, false), xdesc(hf, 0, true));
                tc_commit(d_ready + 8u * (uint32_t)hf);
            }
            for (int s = 0; s < H; ++s)
                for (int hf = 0; hf < 2; ++hf) {
                    mbar_wait_idx(h_ready, (uint32_t)hf, ph_h);
                    if (s + 1 < H) {
                        tc_fence_after();
                        const uint32_t d = d_mine + 16u * (uint32_t)hf;
                        I.x_part(d, xdesc(hf, s + 1, false), xdesc(hf, s + 1, true));
                        I.h_part(d, hdesc(hf, false), hdesc(hf, true));
                        tc_commit(d_ready + 8u * (uint32_t)hf);
                    }
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
                tc_fence_after();
                tn_epilogue4(lane_base, L, 16 * hf + 4 * cg, cst[hf], Hhi, Hlo);
                fence_async_smem();                      // h (generic proxy) -> async proxy of the next MMAs
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(h_ready + 8u * (uint32_t)hf);
            }
            HS_TSTAMP(5 + s);
        }
    }
    __syncthreads();                      // all h of the last step written; the issuing warps have consumed every arrival
    tn_fc_rows<A, NTH>(P, P.b.state_self, P.b.state_drones, W.pred_out, e0, nenv, Hhi, Hlo, fcw, fcb, preds, rowbuf, RI, reinterpret_cast<float*>(Xhi));   // x is dead
    HS_TSTAMP(20);
    tc_fence_before();
    __syncthreads();
    if (warp == FUSED_TICK_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static size_t tp_half_smem_bytes(const hs_config& c) { return tp_tcn_smem_bytes(c) + 32 + 128; }     // WITH_TICK = false
static size_t tp_fused_smem_bytes(const hs_config& c) {
    return tp_tcn_smem_bytes(c) + 32 + 128 + (size_t)FUSED_TICK_WARPS * FUSED_TICK_WORDS * sizeof(float);
}

// =========================================================================================
// Ping-pong, warp-specialised version for batches with more than one 32-env tile per SM (the default
// there): a CTA advances TWO tiles, one accumulator slot (2 M-tiles x 32 columns) each, sharing the weights
// in TMEM; 16 epilogue warps + 2 issuing warps (one per M-tile).  The epilogue warps never meet at a block
// barrier inside the recurrence: a warp waits for an accumulator (mbarrier d_ready[t], armed by
// tcgen05.commit, count 2), updates its 8 env columns x 32 gate rows, publishes h and arrives on
// h_ready[t] (count 16); an issuing warp waits for h_ready[t], issues the 30 MMAs of the next step of its
// M-tile and commits.  While the tensor pipe works on tile 0 the epilogue warps update tile 1 and vice
// versa.  (For ONE tile per CTA this hand-off is slower than the block barrier of the kernel above -
// 29.2 vs 26.6 us at 4096 envs - so small batches keep hs_tp_fill_tcn_kernel.)
// TMEM: D slot t at columns 64*t, weights at 128..447.
// =========================================================================================

template <int A>
__global__ void __launch_bounds__(TCW_THREADS, 1)
hs_tp_fill_tcw_kernel(const __grid_constant__ KParams P, const __grid_constant__ TPParams W) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const hs_config& c = P.c;
    constexpr int FD = 7 + 3 * A;
    const int H = c.history_step;
    const int F3 = 3 * c.future_step;
    const int E = c.num_envs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;
    const int cg = warp >> 2;                                  // 0..3 epilogue column groups, 4 = issuing warp
    const int ntiles = (E + TN_E - 1) / TN_E;
    constexpr int NT = 2;
    const int ngroups = (ntiles + NT - 1) / NT;

    uint8_t* Hb = smem_raw;                                    // slot t: hi at t*2*TN_H_BYTES, lo right after
    float* fcw = reinterpret_cast<float*>(Hb + NT * 2 * TN_H_BYTES);
    float* fcb = fcw + F3 * TP_HID;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(fcb + 32);    // d_ready[2], h_ready[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 4);
    uint8_t* Xb = reinterpret_cast<uint8_t*>(mbar + 6);        // slot t: hi at t*xslot, lo right after
    const size_t xslot = 2 * (size_t)H * TN_X_STEP;
    float* preds = reinterpret_cast<float*>(Xb + NT * xslot);
    float* rowbuf = preds + TN_E * 3 * FMAX;
    float* wst = reinterpret_cast<float*>(Xb);                 // prologue only: aliases the x / preds / row region

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar)) : "memory");        // d_ready: one commit per M-tile
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_u32(mbar + 1)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 2)) : "memory");   // h_ready: 16 epilogue warps
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 16;" :: "r"(smem_u32(mbar + 3)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < F3 * TP_HID; i += TCW_THREADS) fcw[i] = __ldg(W.fc_w + i);
    if (tid < F3) fcb[tid] = __ldg(W.fc_b + tid);
    const TnLane L = tn_lane_consts(W, row);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    tn_stage_weights<FD, TCW_THREADS>(W, wst, lane_base, row, cg);
    tc_fence_before();
    __syncthreads();                                           // weights are in TMEM; the staging tile may be overwritten
    const uint32_t d_ready = smem_u32(mbar), h_ready = smem_u32(mbar + 2);
    uint32_t ph_d = 0u, ph_h = 0u;                             // phase bits; each role tracks only the barriers it waits on

    const uint32_t warp_u = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const bool issuer = warp_u >= TN_THREADS / 32;
    const uint32_t mytl = warp_u & 1u;                         // M-tile of an issuing warp (warps 16, 17)
    TnIssue I;
    I.aA_hi = tmem_u + TN_COL_A + 160 * mytl;
    I.aA_lo = I.aA_hi + 80;
    I.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN_E >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_mine = tmem_u + mytl * TN_E;

    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int nslots = (NT * grp + 1 >= ntiles) ? 1 : NT;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (t < nslots) {
                const int64_t e0 = (int64_t)(NT * grp + t) * TN_E;
                tn_stage_x<FD, TCW_THREADS>(P, e0, (int)min((int64_t)TN_E, E - e0), H, Xb + t * xslot,
                                           Xb + t * xslot + (size_t)H * TN_X_STEP);
            }
        fence_async_smem();                   // generic-proxy writes (x) -> visible to the MMA's async proxy
        tc_fence_before();
        __syncthreads();
        if (issuer) {
            // ------------------------------------------------------------------ issuing warp
            if (elect_one()) {
                tc_fence_after();
                auto xdesc = [&](int t, int s, bool lo) {
                    return tc_desc(smem_u32(Xb + t * xslot) + (uint32_t)(lo ? H : 0) * TN_X_STEP + (uint32_t)s * TN_X_STEP, TN_X_LBO, TN_SBO);
                };
                auto hdesc = [&](int t, bool lo) { return tc_desc(smem_u32(Hb + t * 2 * TN_H_BYTES + (lo ? TN_H_BYTES : 0)), TN_H_LBO, TN_SBO); };
                {
                    for (int t = 0; t < nslots; ++t) {
                        I.x_part(d_mine + (uint32_t)(t * 2 * TN_E), xdesc(t, 0, false), xdesc(t, 0, true), 0u);
                        tc_commit(d_ready + 8u * (uint32_t)t);
                    }
                    for (int s = 0; s < H; ++s)
                        for (int t = 0; t < nslots; ++t) {
                            mbar_wait_idx(h_ready, (uint32_t)t, ph_h);
                            if (s + 1 < H) {
                                tc_fence_after();
                                const uint32_t d = d_mine + (uint32_t)(t * 2 * TN_E);
                                I.x_part(d, xdesc(t, s + 1, false), xdesc(t, s + 1, true), 0u);
                                I.h_part(d, hdesc(t, false), hdesc(t, true));
                                tc_commit(d_ready + 8u * (uint32_t)t);
                            }
                        }
                }
            }
            __syncwarp();
        } else {
            // ------------------------------------------------------------------ epilogue warps
            float cst[NT][8];
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) cst[t][j] = 0.f;
            for (int s = 0; s < H; ++s) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    if (t < nslots) {
                        const int b = t;                                // accumulator slot = barrier index
                        mbar_wait_idx(d_ready, (uint32_t)b, ph_d);
                        tc_fence_after();
                        uint8_t* Hhi = Hb + t * 2 * TN_H_BYTES;
                        tn_epilogue(lane_base + (uint32_t)(b * 2 * TN_E), L, cg, cst[t], Hhi, Hhi + TN_H_BYTES);
                        fence_async_smem();                      // h (generic proxy) -> async proxy of the next MMAs
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(h_ready + 8u * (uint32_t)t);
                    }
                }
            }
        }
        __syncthreads();                      // all h of the last step written; the issuing warp has consumed every arrival
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (t < nslots) {
                const int64_t e0 = (int64_t)(NT * grp + t) * TN_E;
                const int nenv_t = (int)min((int64_t)TN_E, E - e0);
                tn_fc_rows<A, TCW_THREADS>(P, P.b.state_self, P.b.state_drones, W.pred_out, e0, nenv_t, Hb + t * 2 * TN_H_BYTES,
                                          Hb + t * 2 * TN_H_BYTES + TN_H_BYTES, fcw, fcb, preds, rowbuf, tn_row_load<A>(P, e0, nenv_t));
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static size_t tp_tcw_smem_bytes(const hs_config& c) {
    const int NT = 2;
    const int F3 = 3 * c.future_step;
    const size_t region = 2 * (size_t)NT * c.history_step * TN_X_STEP +
                          ((size_t)TN_E * 3 * FMAX + (size_t)TN_E * c.num_agents * (20 + 3 * FMAX)) * sizeof(float);
    const size_t wst = (size_t)256 * TN_WPITCH * sizeof(float);
    return 2 * (size_t)NT * TN_H_BYTES + ((size_t)F3 * TP_HID + 32) * sizeof(float) + 48 + (region > wst ? region : wst);
}


}  // namespace
