// hs_policy_tc.cuh -- the policy network of hs_policy.cuh on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Same network, same folded algebra (W_kq, W_ov), fp32-level results through error-compensated 3xTF32
// (D = A_lo B_hi + A_hi B_lo + A_hi B_hi, fp32 accumulation).  A CTA (16 warps) owns a tile of 128 rows = the 128 TMEM
// lanes; the five dense layers are tcgen05.mma.kind::tf32 with M = 128 (rows), N = 128 (output features), K = 8 per
// instruction, issued by one thread:
//   * A = the layer's input activations, tf32 hi/lo, IN TMEM (TS form): thread (row, part) owns 32 features of its row
//     and writes them with tcgen05.st after its epilogue, so activations never touch shared memory;
//   * B = the layer's weights, tf32 hi/lo, in shared memory in the canonical K-major no-swizzle core-matrix layout
//     (8 features x 16 B, SBO 128 B between 8-feature groups, LBO 2 KB between 16 B K-chunks).  hs_policy_prepare
//     writes every layer ALREADY split and laid out like that, so a layer arrives with two cp.async.bulk copies (TMA,
//     mbarrier complete_tx) that run under the previous layer's epilogue;
//   * D[128 x 128] fp32 in TMEM; the epilogues (bias, residual, LayerNorm, GELU, head) read it with tcgen05.ld, 32
//     columns per thread; row-wide sums are combined across a row's 4 threads through a small shared exchange buffer;
//   * the attention over the 6 tokens is thread-local per (row, 32 features) with the same exchange.
// TMEM columns: D [0,128), A_hi [128,256), A_lo [256,384), fp32 stash (x0, then y1: the residual inputs) [384,512).
// Part of the single translation unit hs_kernels.cu.
#pragma once

namespace {

constexpr int PT_THREADS = 512;
constexpr int PT_M = 128;
constexpr uint32_t PT_LBO = 2048, PT_SBO = 128;
constexpr uint32_t PT_LAYER_BYTES = (PL_E / 4) * PT_LBO;          // 65536: one 128 x 128 matrix, hi or lo
constexpr int PT_COL_D = 0, PT_COL_AHI = 128, PT_COL_ALO = 256, PT_COL_STASH = 384;
constexpr int PT_XW = 24;                                          // exchange floats per (row, part)

struct PolicyTcBlob {          // byte offsets into the tensor-core weight image (hi then lo per layer)
    uint32_t L0, L1, L2, L3, L4, total, k0;                        // k0 = padded K of the embedding layer (multiple of 8)
};
__host__ __device__ inline PolicyTcBlob policy_tc_layout(int self_dim) {
    PolicyTcBlob T;
    T.k0 = (uint32_t)((self_dim + 7) & ~7);
    const uint32_t l0 = (T.k0 / 4) * PT_LBO;
    T.L0 = 0; T.L1 = 2 * l0; T.L2 = T.L1 + 2 * PT_LAYER_BYTES; T.L3 = T.L2 + 2 * PT_LAYER_BYTES; T.L4 = T.L3 + 2 * PT_LAYER_BYTES;
    T.total = T.L4 + 2 * PT_LAYER_BYTES;
    return T;
}

// K-major fp32 matrices of the FFMA blob -> tf32 hi/lo images in the canonical core-matrix layout
__global__ void __launch_bounds__(128)
hs_policy_prepare_tc_kernel(const float* __restrict__ blob, uint8_t* __restrict__ img, int self_dim) {
    const PolicyBlob L = policy_blob_layout(self_dim);
    const PolicyTcBlob T = policy_tc_layout(self_dim);
    const int n = threadIdx.x;              // output feature
    const int k = blockIdx.x;               // input feature
    const int layer = blockIdx.y;
    const int srcs[5] = {L.We0t, L.Wkqt, L.Wovt, L.W1t, L.W2t};
    const uint32_t dsts[5] = {T.L0, T.L1, T.L2, T.L3, T.L4};
    const int K = layer == 0 ? (int)T.k0 : PL_E;
    if (k >= K) return;
    const float w = (layer == 0 && k >= L.Dpad) ? 0.0f : blob[srcs[layer] + k * PL_E + n];
    uint32_t hi, lo;
    tf32_split(w, hi, lo);
    const uint32_t off = (uint32_t)((k >> 2) * PT_LBO + (n >> 3) * PT_SBO + (n & 7) * 16 + (k & 3) * 4);
    const uint32_t lbytes = (uint32_t)(K / 4) * PT_LBO;
    *reinterpret_cast<uint32_t*>(img + dsts[layer] + off) = hi;
    *reinterpret_cast<uint32_t*>(img + dsts[layer] + lbytes + off) = lo;
}

__device__ __forceinline__ uint64_t pt_bdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(PT_LBO >> 4) << 16) | ((uint64_t)(PT_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void pt_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                    "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26) on SFU ex2/rcp: the GELU of the feed-forward block
__device__ __forceinline__ float pt_erf(float x) {
    const float ax = fabsf(x);
    const float t = frcp(fmaf(0.3275911f, ax, 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float r = 1.0f - poly * fex2(-1.4426950408889634f * ax * ax);
    return copysignf(r, x);
}

// 32 consecutive floats of a staged vector (this thread's features) as 8 broadcast LDS.128 - the epilogues are bound by
// the shared-memory pipe (every lane of a warp reads the same address), so scalar loads cost four times as much
__device__ __forceinline__ void pt_ld32f(const float* p, float (&o)[32]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(p + i);
        o[i] = t.x; o[i + 1] = t.y; o[i + 2] = t.z; o[i + 3] = t.w;
    }
}

// 32-term dot product / sums with four independent accumulators (the epilogues run 4 warps per scheduler: a 32-long
// dependent FMA chain per value would leave the pipe idle three cycles out of four)
__device__ __forceinline__ float pt_dot32(const float (&a)[32], const float (&b)[32]) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        s0 = fmaf(a[i], b[i], s0); s1 = fmaf(a[i + 1], b[i + 1], s1);
        s2 = fmaf(a[i + 2], b[i + 2], s2); s3 = fmaf(a[i + 3], b[i + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float pt_sum32(const float (&a)[32]) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) { s0 += a[i]; s1 += a[i + 1]; s2 += a[i + 2]; s3 += a[i + 3]; }
    return (s0 + s1) + (s2 + s3);
}

// sum over the 4 threads of a row of N partial values (two block barriers: the buffer is reused)
template <int N>
__device__ __forceinline__ void pt_row_exchange(float* xch, int row, int part, const float (&mine)[N], float (&total)[N]) {
    // layout [part][value][row]: the 32 lanes of a warp (32 consecutive rows) hit 32 different banks
#pragma unroll
    for (int i = 0; i < N; ++i) xch[(part * PT_XW + i) * PT_M + row] = mine[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i)
        total[i] = (xch[i * PT_M + row] + xch[(PT_XW + i) * PT_M + row]) + (xch[(2 * PT_XW + i) * PT_M + row] + xch[(3 * PT_XW + i) * PT_M + row]);
    __syncthreads();
}

template <int DUMMY = 0>
__global__ void __launch_bounds__(PT_THREADS, 1)
hs_policy_forward_tc_kernel(const PolicyArgs A, const uint8_t* __restrict__ img) {
    extern __shared__ __align__(1024) uint8_t pt_smem[];
    const PolicyBlob L = policy_blob_layout(A.D);
    const PolicyTcBlob T = policy_tc_layout(A.D);
    const float* __restrict__ blob = A.blob;
    uint8_t* Bsm = pt_smem;                                            // [hi 64 KB | lo 64 KB]
    float* xch = reinterpret_cast<float*>(Bsm + 2 * PT_LAYER_BYTES);   // [128 rows][4 parts][PT_XW]
    float* oc = xch + PT_M * 4 * PT_XW;                                // [128][PL_MAX_TOK_IN]
    float* vec = oc + PT_M * PL_MAX_TOK_IN;                            // staged vectors, see V_* below
    enum { V_BE0 = 0, V_LNE_W = 128, V_LNE_B = 256, V_BKQ = 384, V_BOV = 512, V_LN1_W = 640, V_LN1_B = 768, V_B1 = 896,
           V_B2 = 1024, V_LN2_W = 1152, V_LN2_B = 1280, V_BEO = 1408, V_BEC = 1536, V_WEO = 1664, V_WEC = 2048,
           V_WH = 2688, V_BH = 3712, V_LS = 3720, V_GRAM = 3728, V_TOTAL = 3792 };
    uint64_t* mbar = reinterpret_cast<uint64_t*>(vec + V_TOTAL);       // [0] MMA done, [1] weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);
    __shared__ unsigned long long rng_sh[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;             // TMEM lane = row of the tile
    const int part = warp >> 2;                         // features [32 part, 32 part + 32)
    const int f0 = 32 * part;
    const int no3 = A.n_others * 3, nc5 = A.n_cyl * 5, tok_in = no3 + nc5, nx = A.n_others + A.n_cyl;
    const int64_t ntiles = (A.R + PT_M - 1) / PT_M;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar + 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (A.rng != nullptr) {             // see hs_policy_forward_kernel: read {seed, step}, sign in, last CTA advances
            rng_sh[0] = A.rng[0];
            rng_sh[1] = *reinterpret_cast<volatile unsigned long long*>(A.rng + 1);
            __threadfence();
            const unsigned long long seen = atomicAdd(reinterpret_cast<unsigned long long*>(A.rng + 2), 1ull);
            if (seen == (unsigned long long)gridDim.x - 1ull) { A.rng[2] = 0ull; A.rng[1] = rng_sh[1] + 1ull; }
        }
    }
    // vectors the epilogues need, once per CTA
    {
        const int src[13] = {L.be0, L.lnE_w, L.lnE_b, L.bkq, L.bov, L.ln1_w, L.ln1_b, L.b1, L.b2, L.ln2_w, L.ln2_b, L.beo, L.bec};
        for (int i = tid; i < 13 * PL_E; i += PT_THREADS) vec[i] = __ldg(blob + src[i >> 7] + (i & 127));
        for (int i = tid; i < 3 * PL_E; i += PT_THREADS) vec[V_WEO + i] = __ldg(blob + L.Weo + i);
        for (int i = tid; i < 5 * PL_E; i += PT_THREADS) vec[V_WEC + i] = __ldg(blob + L.Wec + i);
        for (int i = tid; i < PL_HEAD_MAX * PL_E; i += PT_THREADS) vec[V_WH + i] = __ldg(blob + L.Wh + i);
        if (tid < PL_HEAD_MAX) { vec[V_BH + tid] = __ldg(blob + L.bh + tid); vec[V_LS + tid] = __ldg(blob + L.log_std + tid); }
        if (tid < 64) vec[V_GRAM + tid] = __ldg(blob + L.gram + tid);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t bar_mma = smem_u32(mbar), bar_w = smem_u32(mbar + 1);
    uint32_t ph_mma = 0, ph_w = 0;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t dhi = pt_bdesc(smem_u32(Bsm));
    const uint32_t k0 = T.k0;

    // weights of one layer: two bulk copies (hi, lo) into the B image, completion on bar_w   (thread 0)
    auto load_weights = [&](uint32_t off, uint32_t kdim) {
        const uint32_t bytes = (kdim / 4) * PT_LBO;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_w), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(Bsm)), "l"(img + off), "r"(bytes), "r"(bar_w) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(Bsm) + PT_LAYER_BYTES), "l"(img + off + bytes), "r"(bytes), "r"(bar_w) : "memory");
    };
    // D = A x W^T over kdim, 3xTF32, then commit; the NEXT layer's weights are requested as soon as these MMAs are done
    auto gemm = [&](uint32_t kdim, uint32_t next_off, uint32_t next_k) {
        tc_wait_st();
        tc_fence_before();
        __syncthreads();                       // A (tcgen05.st of every thread) is in TMEM
        if (tid == 0) {
            mbar_wait(bar_w, ph_w);            // this layer's weights have landed
            tc_fence_after();
            const uint64_t dlo = dhi + (uint64_t)(PT_LAYER_BYTES >> 4);
            uint32_t acc = 0;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t acol = tmem + ((pass == 0) ? PT_COL_ALO : PT_COL_AHI);
                const uint64_t bd = (pass == 1) ? dlo : dhi;
                for (uint32_t j = 0; j < kdim / 8; ++j) {
                    tc_mma_ts(tmem + PT_COL_D, acol + 8 * j, bd + (uint64_t)((2 * j * PT_LBO) >> 4), idesc, acc);
                    acc = 1;
                }
            }
            tc_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        tc_fence_after();
        if (tid == 0 && next_k != 0) load_weights(next_off, next_k);   // B is free again: runs under the epilogue
    };
    auto store_A = [&](const float (&v)[32]) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tf32_split(v[i], hi[i], lo[i]);
        pt_st32(lane_base + PT_COL_AHI + f0, hi);
        pt_st32(lane_base + PT_COL_ALO + f0, lo);
    };
    auto layernorm = [&](float (&v)[32], int vw, int vb) {
        float p[2] = {pt_sum32(v), pt_dot32(v, v)}, t[2];
        pt_row_exchange<2>(xch, row, part, p, t);
        const float mean = t[0] * (1.0f / PL_E);
        const float rstd = rsqrtf(fmaxf(t[1] * (1.0f / PL_E) - mean * mean, 0.0f) + 1e-5f);
        float g[32], b[32];
        pt_ld32f(vec + vw + f0, g);
        pt_ld32f(vec + vb + f0, b);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf((v[i] - mean) * rstd, g[i], b[i]);
    };
    auto add_vec = [&](float (&v)[32], int off) {
        float b[32];
        pt_ld32f(vec + off + f0, b);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += b[i];
    };

    if (tid == 0) load_weights(T.L0, k0);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * PT_M;
        const int nrow = (int)min((int64_t)PT_M, A.R - row0);
        const bool valid = row < nrow;
        const int64_t r = row0 + (valid ? row : 0);

        // ---- inputs: state_self -> A (K = k0 columns), the other tokens' raw inputs -> shared
        {
            uint32_t hi[32], lo[32];
            if ((uint32_t)f0 < k0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int k = f0 + i;
                    const float x = (valid && k < A.D) ? __ldg(A.state_self + r * A.D + k) : 0.0f;
                    tf32_split(x, hi[i], lo[i]);
                }
                pt_st32(lane_base + PT_COL_AHI + f0, hi);
                pt_st32(lane_base + PT_COL_ALO + f0, lo);
            }
            for (int i = part; i < tok_in; i += 4) {
                float v = 0.f;
                if (valid) v = i < no3 ? __ldg(A.state_others + r * no3 + i) : __ldg(A.cylinders + r * nc5 + (i - no3));
                oc[i * PT_M + row] = v;
            }
        }
        float v[32];
        uint32_t raw[32];
        auto load_D = [&]() {
            tc_ld32(lane_base + PT_COL_D + f0, raw);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        };

        HS_TSTAMP(0);
        // ---- L0: x0 = LN(We0 s + be0)
        gemm(k0, T.L1, PL_E);
        HS_TSTAMP(1);
        load_D();
        add_vec(v, V_BE0);
        layernorm(v, V_LNE_W, V_LNE_B);
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(v[i]);
        pt_st32(lane_base + PT_COL_STASH + f0, raw);          // x0 (fp32) for the attention and the residual
        store_A(v);

        HS_TSTAMP(2);
        // ---- L1: q' = W_kq x0 + b_kq (kept in registers)
        gemm(PL_E, T.L2, PL_E);
        HS_TSTAMP(3);
        float q[32];
        tc_ld32(lane_base + PT_COL_D + f0, raw);
#pragma unroll
        for (int i = 0; i < 32; ++i) q[i] = __uint_as_float(raw[i]);
        add_vec(q, V_BKQ);

        // ---- attention over the agent's own token and the nx other tokens (networks.py:296-306).
        // A token embedding is affine in the token's 3 or 5 raw inputs, y_f = sum_a in~_a W~_af (in~ = inputs and 1,
        // W~ = weights and bias), so the LayerNorm statistics of every token follow from the network constants
        // c_a = sum_f W~_af and M_ab = sum_f W~_af W~_bf (hs_policy_prepare), the scores from the ten row sums
        // G_a = sum_f q_f lw_f W~_af, and the weighted token average from U_a = sum_j p_j rstd_j in~_ja: no token is
        // ever embedded feature by feature.
        {
            constexpr int MT = 6;
            HS_TSTAMP(11);
            tc_ld32(lane_base + PT_COL_STASH + f0, raw);      // x0
            HS_TSTAMP(12);
            float ps[13], tot[13];
#pragma unroll
            for (int a = 0; a < 13; ++a) ps[a] = 0.f;
            float qlw[32];
            {
                float t[32];
                pt_ld32f(vec + V_LNE_W + f0, t);
#pragma unroll
                for (int i = 0; i < 32; ++i) qlw[i] = q[i] * t[i];
                ps[1] = pt_sum32(qlw);
                pt_ld32f(vec + V_LNE_B + f0, t);
                ps[2] = pt_dot32(q, t);
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = __uint_as_float(raw[i]);
                ps[0] = pt_dot32(q, t);
                // G_a: rows of the augmented embedding matrices (3 weights + bias, 5 weights + bias)
                const int rows[10] = {V_WEO, V_WEO + PL_E, V_WEO + 2 * PL_E, V_BEO, V_WEC, V_WEC + PL_E, V_WEC + 2 * PL_E,
                                      V_WEC + 3 * PL_E, V_WEC + 4 * PL_E, V_BEC};
#pragma unroll
                for (int a = 0; a < 10; ++a) {
                    pt_ld32f(vec + rows[a] + f0, t);
                    ps[3 + a] = pt_dot32(qlw, t);
                }
            }
            HS_TSTAMP(13);
            pt_row_exchange<13>(xch, row, part, ps, tot);
            HS_TSTAMP(14);
            auto in = [&](int i) { return oc[i * PT_M + row]; };        // [input][row]: conflict-free
            const float* gr = vec + V_GRAM;
            float mean[MT], rstd[MT], sc[MT], m = tot[0];
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                float s1 = 0.f, s2 = 0.f, s3 = 0.f;
                if (j < A.n_others) {
                    const float t[4] = {in(j * 3), in(j * 3 + 1), in(j * 3 + 2), 1.0f};
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const float4 mrow = *reinterpret_cast<const float4*>(gr + 4 + 4 * a);
                        s1 = fmaf(t[a], gr[a], s1);
                        s3 = fmaf(t[a], tot[3 + a], s3);
                        s2 = fmaf(t[a], fmaf(t[3], mrow.w, fmaf(t[2], mrow.z, fmaf(t[1], mrow.y, t[0] * mrow.x))), s2);
                    }
                } else if (j < nx) {
                    const int tb = no3 + (j - A.n_others) * 5;
                    const float t[6] = {in(tb), in(tb + 1), in(tb + 2), in(tb + 3), in(tb + 4), 1.0f};
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const float2 m0 = *reinterpret_cast<const float2*>(gr + 26 + 6 * a);
                        const float2 m1 = *reinterpret_cast<const float2*>(gr + 28 + 6 * a);
                        const float2 m2 = *reinterpret_cast<const float2*>(gr + 30 + 6 * a);
                        s1 = fmaf(t[a], gr[20 + a], s1);
                        s3 = fmaf(t[a], tot[7 + a], s3);
                        s2 = fmaf(t[a], fmaf(t[5], m2.y, fmaf(t[4], m2.x, fmaf(t[3], m1.y, fmaf(t[2], m1.x, fmaf(t[1], m0.y, t[0] * m0.x))))), s2);
                    }
                }
                mean[j] = s1 * (1.0f / PL_E);
                rstd[j] = rsqrtf(fmaxf(s2 * (1.0f / PL_E) - mean[j] * mean[j], 0.0f) + 1e-5f);
                sc[j] = j < nx ? fmaf(rstd[j], s3 - mean[j] * tot[1], tot[2]) : -INFINITY;
                m = fmaxf(m, sc[j]);
            }
            HS_TSTAMP(15);
            const float p0 = expf(tot[0] - m);
            float l = p0, pl = 0.f, pm = 0.f, uo[4] = {0.f, 0.f, 0.f, 0.f}, uc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const float pj = expf(sc[j] - m);               // exp(-inf) = 0 for absent tokens
                l += pj; pl += pj;
                const float pr = pj * rstd[j];
                pm = fmaf(pr, mean[j], pm);
                if (j < A.n_others) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) uo[a] = fmaf(pr, in(j * 3 + a), uo[a]);
                    uo[3] += pr;
                } else if (j < nx) {
                    const int tb = no3 + (j - A.n_others) * 5;
#pragma unroll
                    for (int a = 0; a < 5; ++a) uc[a] = fmaf(pr, in(tb + a), uc[a]);
                    uc[5] += pr;
                }
            }
            const float inv = 1.0f / l;
            HS_TSTAMP(16);
            {
                float y[32], t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = -pm;
                const int rows[10] = {V_WEO, V_WEO + PL_E, V_WEO + 2 * PL_E, V_BEO, V_WEC, V_WEC + PL_E, V_WEC + 2 * PL_E,
                                      V_WEC + 3 * PL_E, V_WEC + 4 * PL_E, V_BEC};
                const float u[10] = {uo[0], uo[1], uo[2], uo[3], uc[0], uc[1], uc[2], uc[3], uc[4], uc[5]};
#pragma unroll
                for (int a = 0; a < 10; ++a) {
                    pt_ld32f(vec + rows[a] + f0, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[i] = fmaf(u[a], t[i], y[i]);
                }
                float lwv[32];
                pt_ld32f(vec + V_LNE_W + f0, lwv);
                pt_ld32f(vec + V_LNE_B + f0, t);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(p0, __uint_as_float(raw[i]), fmaf(y[i], lwv[i], pl * t[i])) * inv;
            }
            HS_TSTAMP(17);
            store_A(v);                                   // xbar
            HS_TSTAMP(18);
        }

        HS_TSTAMP(4);
        // ---- L2: y1 = LN1(x0 + W_ov xbar + b_ov)
        gemm(PL_E, T.L3, PL_E);
        HS_TSTAMP(5);
        load_D();
        {
            uint32_t x0r[32];
            tc_ld32(lane_base + PT_COL_STASH + f0, x0r);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(x0r[i]);
            add_vec(v, V_BOV);
        }
        layernorm(v, V_LN1_W, V_LN1_B);
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(v[i]);
        pt_st32(lane_base + PT_COL_STASH + f0, raw);          // y1 replaces x0
        store_A(v);

        HS_TSTAMP(6);
        // ---- L3: h = gelu(W1 y1 + b1)
        gemm(PL_E, T.L4, PL_E);
        HS_TSTAMP(7);
        load_D();
        add_vec(v, V_B1);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.0f + pt_erf(v[i] * 0.70710678118654752f));
        store_A(v);

        // ---- L4: y2 = LN2(y1 + W2 h + b2); the next tile's first layer is requested behind it
        const bool more = tile + gridDim.x < ntiles;
        HS_TSTAMP(8);
        gemm(PL_E, T.L0, more ? k0 : 0u);
        HS_TSTAMP(9);
        load_D();
        {
            uint32_t y1r[32];
            tc_ld32(lane_base + PT_COL_STASH + f0, y1r);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(y1r[i]);
            add_vec(v, V_B2);
        }
        layernorm(v, V_LN2_W, V_LN2_B);
        if (A.feat_out != nullptr && valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(A.feat_out + r * PL_E + f0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        // ---- head + sample + log-prob
        {
            float ph[PL_HEAD_MAX], th[PL_HEAD_MAX];
#pragma unroll
            for (int h = 0; h < PL_HEAD_MAX; ++h) {
                float s = 0.f;
                if (h < A.head_dim) {
                    float t[32];
                    pt_ld32f(vec + V_WH + h * PL_E + f0, t);
                    s = pt_dot32(v, t);
                }
                ph[h] = s;
            }
            pt_row_exchange<PL_HEAD_MAX>(xch, row, part, ph, th);
            if (part == 0 && valid) {
                float z[PL_HEAD_MAX];
#pragma unroll
                for (int h = 0; h < PL_HEAD_MAX; ++h) z[h] = 0.f;
                if (A.rng != nullptr) {
#pragma unroll
                    for (int blk = 0; blk < PL_HEAD_MAX / 4; ++blk) {
                        if (4 * blk < A.head_dim) {
                            const unsigned long long seed = rng_sh[0], step = rng_sh[1];
                            const uint4 u = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)((unsigned long long)r >> 32), (uint32_t)step,
                                                                     ((uint32_t)(step >> 32) << 1) | (uint32_t)blk),
                                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                            const float u0 = ((float)(u.x >> 8) + 1.0f) * 5.9604644775390625e-08f, u1 = (float)(u.y >> 8) * 5.9604644775390625e-08f;
                            const float u2 = ((float)(u.z >> 8) + 1.0f) * 5.9604644775390625e-08f, u3 = (float)(u.w >> 8) * 5.9604644775390625e-08f;
                            const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
                            float sn0, cs0, sn1, cs1;
                            sincosf(6.283185307179586f * u1, &sn0, &cs0);
                            sincosf(6.283185307179586f * u3, &sn1, &cs1);
                            z[4 * blk] = r0 * cs0; z[4 * blk + 1] = r0 * sn0; z[4 * blk + 2] = r1 * cs1; z[4 * blk + 3] = r1 * sn1;
                        }
                    }
                }
                float lp = 0.f;
#pragma unroll
                for (int h = 0; h < PL_HEAD_MAX; ++h) {
                    if (h < A.head_dim) {
                        const float mean = th[h] + vec[V_BH + h];
                        A.head_out[r * A.head_dim + h] = mean;
                        if (A.action != nullptr || A.logp != nullptr) {
                            const float ls = vec[V_LS + h];
                            const float sd = expf(ls);
                            const float noise = A.eps ? __ldg(A.eps + r * A.head_dim + h) : z[h];
                            if (A.eps_out) A.eps_out[r * A.head_dim + h] = noise;
                            const float act = (A.eps || A.rng) ? fmaf(sd, noise, mean) : mean;
                            if (A.action) A.action[r * A.head_dim + h] = act;
                            const float d = act - mean;
                            lp += -(d * d) / (2.0f * sd * sd) - ls - 0.91893853320467274f;
                        }
                    }
                }
                if (A.logp) A.logp[r] = lp;
            }
        }
    }
    HS_TSTAMP(10);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

static size_t policy_tc_smem_bytes() {
    return 2 * (size_t)PT_LAYER_BYTES + ((size_t)PT_M * 4 * PT_XW + (size_t)PT_M * PL_MAX_TOK_IN + 3792) * sizeof(float) + 64;
}

}  // namespace
