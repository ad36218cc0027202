// hs_policy.cuh -- inference of the MAPPO actor / critic that runs next to the tick:
// PartialAttentionEncoder (omni_drones/learning/modules/networks.py:249-314: SplitEmbedding -> LayerNorm -> one-head
// attention with the agent's own token as the query -> post-norm feed-forward block) followed by DiagGaussian
// (modules/distributions.py:66-82; Actor.forward, mappo.py:614-635) or the value head (Critic.forward, mappo.py:652-668),
// as ONE kernel per network instead of ~25 eager launches, so that actor -> tick -> predictor replay as one CUDA graph.
//
// Algebra (exact in real arithmetic, fp32 rounding differs at the 1e-6 level):
//   * scores: q.k_j = x_j^T Wk^T (Wq x_0 + bq) + q.bk.  The last term is the same for every key and cancels in the
//     softmax; W_kq = Wk^T Wq / sqrt(d) and b_kq = Wk^T bq / sqrt(d) are formed once per weight update (hs_policy_prepare),
//     so the keys are never projected: s_j = x_j . (W_kq x_0 + b_kq).
//   * values: sum_j p_j (Wv x_j + bv) = Wv (sum_j p_j x_j) + bv, and the output projection folds in:
//     attn = W_ov xbar + b_ov with W_ov = Wo Wv, b_ov = Wo bv + bo.
//   A row therefore costs four 128x128 matrix-vector products and one D x128 instead of fifteen: 73 k MAC instead of 270 k.
//
// Mapping: a CTA owns RT rows (one row = one agent of one env).  The dense layers are [RT x K] x [K x 128] tiles on the
// FFMA pipe: 4 rows x 8 columns per thread, activations K-major in shared memory (one LDS.128 = 4 rows), weights streamed
// K-major from the prepared blob through a cp.async double buffer (2 LDS.128 = 8 columns) -> 32 FFMA per 3 LDS.  LayerNorm
// statistics and the head are reduced with shuffles across the 16 lanes that share a row group; the attention (6 tokens)
// is one warp per row with an online softmax.  No tensor cores: fp32 parity (1e-4) rules out single-pass TF32 and the
// weights of a 3xTF32 scheme (4 x 128 KB hi/lo) do not fit TMEM; see DESIGN.md.
// Part of the single translation unit hs_kernels.cu.
#pragma once

namespace {

constexpr int PL_E = 128;                 // embed_dim = dim_feedforward = 128 (networks.py:256-259 defaults)
constexpr int PL_HEAD_MAX = 8;
constexpr int PL_KC = 16;                 // weight rows per cp.async chunk
constexpr int PL_STAGES = 3;              // cp.async ring depth
constexpr int PL_MAX_TOK_IN = 2 * 3 + 4 * 5;   // others (<= 2 x 3) + cylinders (<= 4 x 5) floats per row

struct PolicyBlob {                        // offsets (floats) into the prepared parameter blob
    int We0t, be0, Weo, beo, Wec, bec, lnE_w, lnE_b, Wkqt, bkq, Wovt, bov, ln1_w, ln1_b, W1t, b1, W2t, b2, ln2_w, ln2_b,
        Wh, bh, log_std, gram, total, Dpad;
};
// gram: column sums and Gram matrices of the two token embeddings over the 128 features (constants of the network):
// [0..3] c^o_a = sum_f W^o_af (a = 3: bias), [4..19] M^o_ab = sum_f W^o_af W^o_bf, [20..25] c^c_a (a = 5: bias), [26..61] M^c_ab.
// With them the LayerNorm statistics of a token embedding follow from the token's 3 or 5 raw inputs alone.
__host__ __device__ inline PolicyBlob policy_blob_layout(int self_dim) {
    PolicyBlob L;
    int o = 0;
    L.Dpad = (self_dim + 3) & ~3;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    L.We0t = take(L.Dpad * PL_E); L.be0 = take(PL_E);
    L.Weo = take(3 * PL_E); L.beo = take(PL_E);
    L.Wec = take(5 * PL_E); L.bec = take(PL_E);
    L.lnE_w = take(PL_E); L.lnE_b = take(PL_E);
    L.Wkqt = take(PL_E * PL_E); L.bkq = take(PL_E);
    L.Wovt = take(PL_E * PL_E); L.bov = take(PL_E);
    L.ln1_w = take(PL_E); L.ln1_b = take(PL_E);
    L.W1t = take(PL_E * PL_E); L.b1 = take(PL_E);
    L.W2t = take(PL_E * PL_E); L.b2 = take(PL_E);
    L.ln2_w = take(PL_E); L.ln2_b = take(PL_E);
    L.Wh = take(PL_HEAD_MAX * PL_E); L.bh = take(PL_HEAD_MAX); L.log_std = take(PL_HEAD_MAX);
    L.gram = take(64);
    L.total = o;
    return L;
}

// ---- prepare: raw nn.Module parameters -> K-major blob with the folded products (once per weight update) ----------
__global__ void __launch_bounds__(128)
hs_policy_prepare_kernel(const hs_policy_weights w, float* __restrict__ blob) {
    const PolicyBlob L = policy_blob_layout(w.self_dim);
    const int a = threadIdx.x;              // output feature
    const int b = blockIdx.x;               // input feature / row of the K-major matrices
    const float* wq = w.attn_in_w;
    const float* wk = w.attn_in_w + PL_E * PL_E;
    const float* wv = w.attn_in_w + 2 * PL_E * PL_E;
    const float scale = rsqrtf((float)PL_E);
    if (b < PL_E) {
        float kq = 0.f, ov = 0.f;
        for (int c = 0; c < PL_E; ++c) {
            kq = fmaf(wk[c * PL_E + a], wq[c * PL_E + b], kq);           // W_kq[a][b] = sum_c Wk[c][a] Wq[c][b]
            ov = fmaf(w.attn_out_w[a * PL_E + c], wv[c * PL_E + b], ov); // W_ov[a][b] = sum_c Wo[a][c] Wv[c][b]
        }
        blob[L.Wkqt + b * PL_E + a] = kq * scale;
        blob[L.Wovt + b * PL_E + a] = ov;
        blob[L.W1t + b * PL_E + a] = w.lin1_w[a * PL_E + b];
        blob[L.W2t + b * PL_E + a] = w.lin2_w[a * PL_E + b];
    }
    if (b < L.Dpad) blob[L.We0t + b * PL_E + a] = b < w.self_dim ? w.embed_self_w[a * w.self_dim + b] : 0.f;
    if (b < 3) blob[L.Weo + b * PL_E + a] = w.embed_others_w ? w.embed_others_w[a * 3 + b] : 0.f;
    if (b < 5) blob[L.Wec + b * PL_E + a] = w.embed_cyl_w ? w.embed_cyl_w[a * 5 + b] : 0.f;
    if (b < PL_HEAD_MAX) blob[L.Wh + b * PL_E + a] = b < w.head_dim ? w.head_w[b * PL_E + a] : 0.f;
    if (b == 0) {
        float bkq = 0.f, bov = w.attn_out_b[a];
        const float* bq = w.attn_in_b;
        const float* bv = w.attn_in_b + 2 * PL_E;
        for (int c = 0; c < PL_E; ++c) {
            bkq = fmaf(wk[c * PL_E + a], bq[c], bkq);
            bov = fmaf(w.attn_out_w[a * PL_E + c], bv[c], bov);
        }
        blob[L.bkq + a] = bkq * scale;
        blob[L.bov + a] = bov;
        blob[L.be0 + a] = w.embed_self_b[a];
        blob[L.beo + a] = w.embed_others_b ? w.embed_others_b[a] : 0.f;
        blob[L.bec + a] = w.embed_cyl_b ? w.embed_cyl_b[a] : 0.f;
        blob[L.lnE_w + a] = w.embed_ln_w[a]; blob[L.lnE_b + a] = w.embed_ln_b[a];
        blob[L.ln1_w + a] = w.norm1_w[a]; blob[L.ln1_b + a] = w.norm1_b[a];
        blob[L.ln2_w + a] = w.norm2_w[a]; blob[L.ln2_b + a] = w.norm2_b[a];
        blob[L.b1 + a] = w.lin1_b[a]; blob[L.b2 + a] = w.lin2_b[a];
        if (a < PL_HEAD_MAX) {
            blob[L.bh + a] = a < w.head_dim ? w.head_b[a] : 0.f;
            blob[L.log_std + a] = (w.log_std && a < w.head_dim) ? w.log_std[a] : 0.f;
        }
    }
    if (b == 1 && a < 62) {
        // augmented embedding rows: others (W_0, W_1, W_2, bias), cylinders (W_0..W_4, bias)
        auto wo = [&](int r, int f) { return !w.embed_others_w ? 0.f : (r < 3 ? w.embed_others_w[f * 3 + r] : w.embed_others_b[f]); };
        auto wc = [&](int r, int f) { return !w.embed_cyl_w ? 0.f : (r < 5 ? w.embed_cyl_w[f * 5 + r] : w.embed_cyl_b[f]); };
        float acc = 0.f;
        for (int f = 0; f < PL_E; ++f) {
            if (a < 4) acc += wo(a, f);
            else if (a < 20) acc = fmaf(wo((a - 4) >> 2, f), wo((a - 4) & 3, f), acc);
            else if (a < 26) acc += wc(a - 20, f);
            else acc = fmaf(wc((a - 26) / 6, f), wc((a - 26) % 6, f), acc);
        }
        blob[L.gram + a] = acc;
    }
}

// ---- forward -------------------------------------------------------------------------------------------------------
struct PolicyArgs {
    const float* blob;
    const float* state_self;      // [R, D]
    const float* state_others;    // [R, n_others, 3] or nullptr
    const float* cylinders;       // [R, n_cyl, 5] or nullptr
    const float* eps;             // [R, head_dim] standard-normal noise, or nullptr
    uint64_t* rng;                // {seed, step, arrivals, -}: in-kernel noise when eps == nullptr (nullptr too: mode)
    float* eps_out;               // [R, head_dim] or nullptr
    float* head_out;              // [R, head_dim]  action mean | state value
    float* action;                // [R, head_dim] or nullptr
    float* logp;                  // [R] or nullptr
    float* feat_out;              // [R, 128] or nullptr (tests)
    int64_t R;
    int D, n_others, n_cyl, head_dim;
};

// acc[4][8] += actT[K][P] (K-major activations, this thread's 4 rows) x Wg[K][128] (this thread's 8 columns)
template <int RT>
__device__ __forceinline__ void pl_gemm(const float* __restrict__ Wg, int K, const float* __restrict__ actT, float* wbuf,
                                        int rg, int cg, float (&acc)[4][8]) {
    constexpr int NT = RT * 4, P = RT + 4;
    const int tid = threadIdx.x;
    const int nchunk = (K + PL_KC - 1) / PL_KC;
    auto prefetch = [&](int ch) {
        if (ch < nchunk) {
            const int k0 = ch * PL_KC, rows = min(PL_KC, K - k0);
            const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)k0 * PL_E);
            float4* dst = reinterpret_cast<float4*>(wbuf + (ch % PL_STAGES) * PL_KC * PL_E);
            for (int i = tid; i < rows * (PL_E / 4); i += NT) cp_async16(dst + i, src + i);
        }
        cp_async_commit();                   // (empty groups keep the wait_group arithmetic uniform)
    };
    auto fma_row = [&](const float4& a, const float4& w0, const float4& w1) {
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], wv[c], acc[r][c]);
    };
    // three-stage ring, ONE barrier per chunk: the barrier that publishes chunk ch also proves that every thread is
    // done with chunk ch-1, whose buffer is the target of the prefetch of chunk ch+2 issued right after it
    prefetch(0);
    prefetch(1);
    for (int ch = 0; ch < nchunk; ++ch) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        prefetch(ch + 2);
        const int k0 = ch * PL_KC, rows = min(PL_KC, K - k0);
        const float* wb = wbuf + (ch % PL_STAGES) * PL_KC * PL_E + 4 * cg;
        const float* ap = actT + k0 * P + 4 * rg;
        // software pipeline: the operands of row kk+1 are loaded before the 32 FMAs of row kk
        float4 a = *reinterpret_cast<const float4*>(ap);
        float4 w0 = *reinterpret_cast<const float4*>(wb);
        float4 w1 = *reinterpret_cast<const float4*>(wb + 64);
        if (rows == PL_KC) {
#pragma unroll
            for (int kk = 0; kk < PL_KC; ++kk) {
                constexpr int last = PL_KC - 1;
                const int kn = kk < last ? kk + 1 : last;                       // compile-time after unrolling
                const float4 an = *reinterpret_cast<const float4*>(ap + kn * P);
                const float4 w0n = *reinterpret_cast<const float4*>(wb + kn * PL_E);
                const float4 w1n = *reinterpret_cast<const float4*>(wb + kn * PL_E + 64);
                fma_row(a, w0, w1);
                a = an; w0 = w0n; w1 = w1n;
            }
        } else {
            for (int kk = 0; kk < rows; ++kk) {
                const int kn = min(kk + 1, rows - 1);
                const float4 an = *reinterpret_cast<const float4*>(ap + kn * P);
                const float4 w0n = *reinterpret_cast<const float4*>(wb + kn * PL_E);
                const float4 w1n = *reinterpret_cast<const float4*>(wb + kn * PL_E + 64);
                fma_row(a, w0, w1);
                a = an; w0 = w0n; w1 = w1n;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();                         // the ring is reused by the next layer
}

// column index of this thread's c-th accumulator column
__device__ __forceinline__ int pl_col(int cg, int c) { return (c < 4 ? 0 : 64) + 4 * cg + (c & 3); }

// sum over the 16 lanes that share a row group (lanes 0-15 / 16-31 of a warp)
__device__ __forceinline__ float pl_rowsum(float v) {
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 8);
    return v;
}

// LayerNorm over the 128 features of each of this thread's 4 rows (values spread over 16 lanes x 8 columns), eps 1e-5
__device__ __forceinline__ void pl_layernorm(float (&v)[4][8], const float* __restrict__ g, const float* __restrict__ b, int cg) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) s += v[r][c];
        const float mean = pl_rowsum(s) * (1.0f / PL_E);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float d = v[r][c] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(pl_rowsum(q) * (1.0f / PL_E) + 1e-5f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int col = pl_col(cg, c);
            v[r][c] = (v[r][c] - mean) * rstd * __ldg(g + col) + __ldg(b + col);
        }
    }
}

template <int RT>
__device__ __forceinline__ void pl_store(float* bufT, const float (&v)[4][8], int rg, int cg) {
    constexpr int P = RT + 4;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(bufT + pl_col(cg, c) * P + 4 * rg) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
}

template <int RT>
__global__ void __launch_bounds__(RT * 4)
hs_policy_forward_kernel(const PolicyArgs A) {
    constexpr int NT = RT * 4, P = RT + 4, NW = NT / 32;
    extern __shared__ __align__(16) float pl_smem[];
    const PolicyBlob L = policy_blob_layout(A.D);
    const float* __restrict__ blob = A.blob;
    float* bufA = pl_smem;                           // [128][P]   x0 -> y1
    float* bufB = bufA + PL_E * P;                   // [128][P]   q' -> xbar -> gelu(ff1)
    float* inT = bufB + PL_E * P;                    // [Dpad][P]  state_self, K-major
    float* oc = inT + L.Dpad * P;                    // [RT][tok_in] other-agent and cylinder rows
    float* wbuf = oc + RT * PL_MAX_TOK_IN;           // [PL_STAGES][16][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = tid >> 4, cg = tid & 15;
    const int64_t row0 = (int64_t)blockIdx.x * RT;
    const int nrow = (int)min((int64_t)RT, A.R - row0);
    const int D = A.D, no3 = A.n_others * 3, nc5 = A.n_cyl * 5, tok_in = no3 + nc5;

    // in-kernel noise: every CTA reads {seed, step} BEFORE it signs in; the CTA that signs in last advances the step
    // for the next launch, so no launch is spent on the counter and no CTA can see the new value
    __shared__ unsigned long long rng_sh[2];
    if (A.rng != nullptr && tid == 0) {
        rng_sh[0] = A.rng[0];
        rng_sh[1] = *reinterpret_cast<volatile unsigned long long*>(A.rng + 1);
        __threadfence();
        const unsigned long long seen = atomicAdd(reinterpret_cast<unsigned long long*>(A.rng + 2), 1ull);
        if (seen == (unsigned long long)gridDim.x - 1ull) {
            A.rng[2] = 0ull;
            A.rng[1] = rng_sh[1] + 1ull;
        }
    }

    // ---- stage the observation rows of the tile
    for (int i = tid; i < RT * L.Dpad; i += NT) {
        const int r = i / L.Dpad, k = i - r * L.Dpad;
        inT[k * P + r] = (r < nrow && k < D) ? __ldg(A.state_self + (row0 + r) * D + k) : 0.f;
    }
    for (int i = tid; i < RT * tok_in; i += NT) {
        const int r = i / tok_in, k = i - r * tok_in;
        float v = 0.f;
        if (r < nrow) v = k < no3 ? __ldg(A.state_others + (row0 + r) * no3 + k) : __ldg(A.cylinders + (row0 + r) * nc5 + (k - no3));
        oc[r * PL_MAX_TOK_IN + k] = v;
    }
    __syncthreads();

    float acc[4][8];
    auto zero = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    };
    auto add_bias = [&](int off) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float bb = __ldg(blob + off + pl_col(cg, c));
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][c] += bb;
        }
    };

    // ---- token 0: x0 = LN(We0 s + be0)                               networks.py:153-161
    zero();
    pl_gemm<RT>(blob + L.We0t, L.Dpad, inT, wbuf, rg, cg, acc);
    add_bias(L.be0);
    pl_layernorm(acc, blob + L.lnE_w, blob + L.lnE_b, cg);
    pl_store<RT>(bufA, acc, rg, cg);
    __syncthreads();

    // ---- q' = W_kq x0 + b_kq
    zero();
    pl_gemm<RT>(blob + L.Wkqt, PL_E, bufA, wbuf, rg, cg, acc);
    add_bias(L.bkq);
    pl_store<RT>(bufB, acc, rg, cg);
    __syncthreads();

    // ---- attention: one warp per row, lane owns features lane + 32 i         networks.py:296-306
    {
        float weo[3][4], wec[5][4], beo[4], bec[4], lw[4], lb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = lane + 32 * i;
#pragma unroll
            for (int d = 0; d < 3; ++d) weo[d][i] = __ldg(blob + L.Weo + d * PL_E + f);
#pragma unroll
            for (int d = 0; d < 5; ++d) wec[d][i] = __ldg(blob + L.Wec + d * PL_E + f);
            beo[i] = __ldg(blob + L.beo + f); bec[i] = __ldg(blob + L.bec + f);
            lw[i] = __ldg(blob + L.lnE_w + f); lb[i] = __ldg(blob + L.lnE_b + f);
        }
        // All warp reductions of a row are issued as independent butterfly chains (two rounds: the token means, then
        // the centred second moments and the q-weighted sums), so their shuffle latencies overlap.
        constexpr int MT = 6;                                   // tokens besides the agent's own: n_others + n_cyl <= 6
        const int nx = A.n_others + A.n_cyl;
        for (int r = warp; r < RT; r += NW) {
            float q[4], x0[4], qlw[4], y[MT][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                q[i] = bufB[(lane + 32 * i) * P + r];
                x0[i] = bufA[(lane + 32 * i) * P + r];
                qlw[i] = q[i] * lw[i];
            }
            const float* in = oc + r * PL_MAX_TOK_IN;
            float red[3 + MT];                                  // round 1: q.x0, sum q lw, sum q lb, token sums
            red[0] = fmaf(q[3], x0[3], fmaf(q[2], x0[2], fmaf(q[1], x0[1], q[0] * x0[0])));
            red[1] = (qlw[0] + qlw[1]) + (qlw[2] + qlw[3]);
            red[2] = fmaf(q[3], lb[3], fmaf(q[2], lb[2], fmaf(q[1], lb[1], q[0] * lb[0])));
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (j < nx) {
                    if (j < A.n_others) {
                        const float* t = in + j * 3;
#pragma unroll
                        for (int i = 0; i < 4; ++i) y[j][i] = fmaf(t[2], weo[2][i], fmaf(t[1], weo[1][i], fmaf(t[0], weo[0][i], beo[i])));
                    } else {
                        const float* t = in + no3 + (j - A.n_others) * 5;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            y[j][i] = fmaf(t[4], wec[4][i], fmaf(t[3], wec[3][i], fmaf(t[2], wec[2][i], fmaf(t[1], wec[1][i], fmaf(t[0], wec[0][i], bec[i])))));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) y[j][i] = 0.f;
                }
                red[3 + j] = (y[j][0] + y[j][1]) + (y[j][2] + y[j][3]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int k = 0; k < 3 + MT; ++k) red[k] += __shfl_xor_sync(FULL, red[k], o);
            float var[MT], dot[MT];                             // round 2 (centred): sum (y - mean)^2, sum q lw (y - mean)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const float mean = red[3 + j] * (1.0f / PL_E);
                float v = 0.f, d = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) { y[j][i] -= mean; v = fmaf(y[j][i], y[j][i], v); d = fmaf(qlw[i], y[j][i], d); }
                var[j] = v; dot[j] = d;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                    var[j] += __shfl_xor_sync(FULL, var[j], o);
                    dot[j] += __shfl_xor_sync(FULL, dot[j], o);
                }
            // scores: s_0 = q.x0, s_j = q.LN(y_j) = rstd_j sum q lw (y_j - mean_j) + sum q lb; softmax over 1 + nx tokens
            float sc[MT], rstd[MT], m = red[0];
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                rstd[j] = rsqrtf(var[j] * (1.0f / PL_E) + 1e-5f);
                sc[j] = j < nx ? fmaf(rstd[j], dot[j], red[2]) : -INFINITY;
                m = fmaxf(m, sc[j]);
            }
            const float p0 = expf(red[0] - m);
            float l = p0, pl = 0.f, xb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xb[i] = 0.f;
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const float pj = expf(sc[j] - m);               // exp(-inf) = 0 for absent tokens
                l += pj; pl += pj;
                const float pr = pj * rstd[j];
#pragma unroll
                for (int i = 0; i < 4; ++i) xb[i] = fmaf(pr, y[j][i], xb[i]);
            }
            const float inv = 1.0f / l;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                bufB[(lane + 32 * i) * P + r] = (fmaf(p0, x0[i], fmaf(xb[i], lw[i], pl * lb[i]))) * inv;
        }
    }
    __syncthreads();

    // ---- y1 = LN1(x0 + W_ov xbar + b_ov)                              networks.py:300
    zero();
    pl_gemm<RT>(blob + L.Wovt, PL_E, bufB, wbuf, rg, cg, acc);
    add_bias(L.bov);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 x0 = *reinterpret_cast<const float4*>(bufA + pl_col(cg, c) * P + 4 * rg);
        acc[0][c] += x0.x; acc[1][c] += x0.y; acc[2][c] += x0.z; acc[3][c] += x0.w;
    }
    pl_layernorm(acc, blob + L.ln1_w, blob + L.ln1_b, cg);
    pl_store<RT>(bufA, acc, rg, cg);      // own elements only: safe while other threads still read their x0
    __syncthreads();

    // ---- h = gelu(W1 y1 + b1)                                         networks.py:308-310
    zero();
    pl_gemm<RT>(blob + L.W1t, PL_E, bufA, wbuf, rg, cg, acc);
    add_bias(L.b1);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.5f * acc[r][c] * (1.0f + erff(acc[r][c] * 0.70710678118654752f));
    pl_store<RT>(bufB, acc, rg, cg);
    __syncthreads();

    // ---- y2 = LN2(y1 + W2 h + b2); features = y2 (mean over the single query token)   networks.py:301-302
    zero();
    pl_gemm<RT>(blob + L.W2t, PL_E, bufB, wbuf, rg, cg, acc);
    add_bias(L.b2);
#pragma unroll
    for (int c = 0; c < 8; ++c) {          // residual y1: this thread's own elements, still in bufA
        const float4 y1 = *reinterpret_cast<const float4*>(bufA + pl_col(cg, c) * P + 4 * rg);
        acc[0][c] += y1.x; acc[1][c] += y1.y; acc[2][c] += y1.z; acc[3][c] += y1.w;
    }
    pl_layernorm(acc, blob + L.ln2_w, blob + L.ln2_b, cg);
    if (A.feat_out != nullptr) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (4 * rg + r < nrow)
#pragma unroll
                for (int c = 0; c < 8; ++c) A.feat_out[(row0 + 4 * rg + r) * PL_E + pl_col(cg, c)] = acc[r][c];
    }

    // ---- head: fc_mean / v_out, then sample + log-prob                 distributions.py:78-82, mappo.py:614-635
    float hv[4][PL_HEAD_MAX];
#pragma unroll
    for (int h = 0; h < PL_HEAD_MAX; ++h) {
        if (h < A.head_dim) {
            float wh[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) wh[c] = __ldg(blob + L.Wh + h * PL_E + pl_col(cg, c));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) s = fmaf(acc[r][c], wh[c], s);
                hv[r][h] = pl_rowsum(s) + __ldg(blob + L.bh + h);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) hv[r][h] = 0.f;
        }
    }
    if (cg < 4 && 4 * rg + cg < nrow) {            // lane cg of the row group finishes row 4 rg + cg
        const int64_t row = row0 + 4 * rg + cg;
        float lp = 0.f;
        float z[PL_HEAD_MAX];
#pragma unroll
        for (int h = 0; h < PL_HEAD_MAX; ++h) z[h] = 0.f;
        if (A.rng != nullptr) {
            // Philox4x32-10, key = seed, counter = (row, step, block of four head columns); Box-Muller on (0,1] x [0,1)
#pragma unroll
            for (int blk = 0; blk < PL_HEAD_MAX / 4; ++blk) {
                if (4 * blk < A.head_dim) {
                    const unsigned long long seed = rng_sh[0], step = rng_sh[1];
                    const uint4 u = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)((unsigned long long)row >> 32), (uint32_t)step,
                                                             ((uint32_t)(step >> 32) << 1) | (uint32_t)blk),
                                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                    const float u0 = ((float)(u.x >> 8) + 1.0f) * 5.9604644775390625e-08f, u1 = (float)(u.y >> 8) * 5.9604644775390625e-08f;
                    const float u2 = ((float)(u.z >> 8) + 1.0f) * 5.9604644775390625e-08f, u3 = (float)(u.w >> 8) * 5.9604644775390625e-08f;
                    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
                    float s0, c0, s1, c1;
                    sincosf(6.283185307179586f * u1, &s0, &c0);
                    sincosf(6.283185307179586f * u3, &s1, &c1);
                    z[4 * blk] = r0 * c0; z[4 * blk + 1] = r0 * s0; z[4 * blk + 2] = r1 * c1; z[4 * blk + 3] = r1 * s1;
                }
            }
        }
#pragma unroll
        for (int h = 0; h < PL_HEAD_MAX; ++h) {
            if (h < A.head_dim) {
                float mean = hv[0][h];
                if (cg == 1) mean = hv[1][h];
                if (cg == 2) mean = hv[2][h];
                if (cg == 3) mean = hv[3][h];
                A.head_out[row * A.head_dim + h] = mean;
                if (A.action != nullptr || A.logp != nullptr) {
                    const float ls = __ldg(blob + L.log_std + h);
                    const float sd = expf(ls);
                    const float noise = A.eps ? __ldg(A.eps + row * A.head_dim + h) : z[h];
                    if (A.eps_out) A.eps_out[row * A.head_dim + h] = noise;
                    const float act = (A.eps || A.rng) ? fmaf(sd, noise, mean) : mean;
                    if (A.action) A.action[row * A.head_dim + h] = act;
                    const float d = act - mean;
                    lp += -(d * d) / (2.0f * sd * sd) - ls - 0.91893853320467274f;      // Normal.log_prob
                }
            }
        }
        if (A.logp) A.logp[row] = lp;
    }
}

template <int RT>
static size_t policy_smem_bytes(int self_dim) {
    const int Dpad = (self_dim + 3) & ~3;
    return ((size_t)(2 * PL_E + Dpad) * (RT + 4) + (size_t)RT * PL_MAX_TOK_IN + PL_STAGES * PL_KC * PL_E) * sizeof(float);
}

}  // namespace
