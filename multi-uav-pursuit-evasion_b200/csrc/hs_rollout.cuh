// hs_rollout.cuh -- what the caller does with a finished rollout: generalised advantage
// estimation as one backward scan per (env, agent) column, fused with the batch moments of the advantages, and the
// batch-level advantage normalisation.  Reference: omni_drones/learning/utils/gae.py:27-51 (compute_gae) as called by
// MAPPOPolicy.train_op, omni_drones/learning/mappo.py:381-397.
// Part of the single translation unit hs_kernels.cu.
#pragma once

namespace {

// One thread per (env, agent) column; the T steps are walked backwards with the reference's operation order and
// explicit round-to-nearest intrinsics (no FMA contraction), so advantages / returns are bit-identical to the eager
// fp32 loop.  Columns are independent, consecutive threads own consecutive agents of consecutive envs: in the
// time-major rollout layout [T, E, A] every step is one coalesced row; in the reference's [E, T, A] layout a warp
// touches 32/A env rows per step and the following T-1 steps hit the same 128 B lines in L1/L2.
// moments[0..1] += (sum, sum of squares) of the advantages in double precision (one atomic pair per block).
__global__ void __launch_bounds__(256)
hs_gae_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ done, const float* __restrict__ value,
              const float* __restrict__ next_value, float* __restrict__ adv, float* __restrict__ ret,
              double* __restrict__ moments, int64_t ncol, int T, int A, int64_t se, int64_t st, int64_t dse, int64_t dst,
              float gamma, float gl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (i < ncol) {
        const int64_t e = i / A;
        const int a = (int)(i - e * A);
        const int64_t base = e * se + a;
        const uint8_t* dcol = done + e * dse;
        float nv = next_value[i];
        float gae = 0.0f;
        // software pipeline: the loads of step t-1 are issued before the arithmetic of step t
        float r = reward[base + (int64_t)(T - 1) * st], v = value[base + (int64_t)(T - 1) * st];
        uint8_t d = dcol[(int64_t)(T - 1) * dst];
        for (int t = T - 1; t >= 0; --t) {
            float rn = 0.f, vn = 0.f;
            uint8_t dn = 0;
            if (t > 0) {
                rn = reward[base + (int64_t)(t - 1) * st];
                vn = value[base + (int64_t)(t - 1) * st];
                dn = dcol[(int64_t)(t - 1) * dst];
            }
            const float nd = __fsub_rn(1.0f, d ? 1.0f : 0.0f);
            // delta = reward + gamma * next_value * not_done - value          gae.py:41-45
            const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, nv), nd)), v);
            // gae = delta + gamma * lmbda * not_done * gae                    gae.py:46
            gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd), gae));
            adv[base + (int64_t)t * st] = gae;
            ret[base + (int64_t)t * st] = __fadd_rn(gae, v);                   // gae.py:49
            s1 += (double)gae;
            s2 += (double)gae * (double)gae;
            nv = v;
            r = rn; v = vn; d = dn;
        }
    }
    if (moments != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(FULL, s1, o);
            s2 += __shfl_xor_sync(FULL, s2, o);
        }
        __shared__ double w1[8], w2[8];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) { w1[wid] = s1; w2[wid] = s2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a1 = 0.0, a2 = 0.0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { a1 += w1[k]; a2 += w2[k]; }
            atomicAdd(moments, a1);
            atomicAdd(moments + 1, a2);
        }
    }
}

// advantages <- (advantages - mean) / (std + 1e-8), std with Bessel's correction like torch.Tensor.std()
// (mappo.py:391-396).  mean/std come from the moments the scan accumulated; stats_out = {mean, std}.
__global__ void __launch_bounds__(256)
hs_adv_normalize_kernel(float* __restrict__ adv, const double* __restrict__ moments, float* __restrict__ stats_out,
                        int64_t ncol, int T, int A, int64_t se, int64_t st, int do_normalize) {
    const double n = (double)ncol * (double)T;
    const double mean = moments[0] / n;
    const double var = n > 1.0 ? fmax((moments[1] - n * mean * mean) / (n - 1.0), 0.0) : 0.0;
    const float meanf = (float)mean, stdf = (float)sqrt(var);
    const float denom = __fadd_rn(stdf, 1e-8f);
    const int64_t total = do_normalize ? ncol * T : 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
        // j enumerates (t, column) so that consecutive threads stay on consecutive columns
        const int64_t t = j / ncol, i = j - t * ncol;
        const int64_t e = i / A;
        const int64_t idx = e * se + (i - e * A) + t * st;
        adv[idx] = __fdiv_rn(__fsub_rn(adv[idx], meanf), denom);
    }
    if (stats_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { stats_out[0] = meanf; stats_out[1] = stdf; }
}

// ---------------------------------------------------------------------------------------------------------------
// PPO minibatches straight from the rollout: what make_dataset_naive
// (omni_drones/learning/mappo.py:493-513) yields - `tensordict.reshape(-1)[indices]` for every key, with the flat
// sample index n = env * T + step - gathered from tensors that are strided over (env, step): the engine's time-major
// [T, E, ...] rollout storage or the reference's [E, T, ...] batch, without materialising the flattened copy first.
// One launch gathers every key: blockIdx.y = tensor, a warp per output row; rows move as 16 / 4 / 1 byte words,
// whichever the row size and the addresses allow.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GATHER_MAX_TENSORS = 24;
struct GatherDesc {
    const uint8_t* src;
    uint8_t* dst;
    int64_t stride_env, stride_step;     // bytes between consecutive envs / steps of the source
    int32_t row_bytes, pad;
};
struct GatherParams {
    GatherDesc d[GATHER_MAX_TENSORS];
    const int64_t* indices;              // [num_rows] flat sample ids n = env * T + step
    int64_t num_rows;
    int32_t T, num_tensors;
};

__global__ void __launch_bounds__(256)
hs_gather_rows_kernel(const __grid_constant__ GatherParams G) {
    const GatherDesc& D = G.d[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int rb = D.row_bytes;
    const bool v16 = (rb & 15) == 0 && ((reinterpret_cast<uintptr_t>(D.src) | reinterpret_cast<uintptr_t>(D.dst) |
                                          (uintptr_t)D.stride_env | (uintptr_t)D.stride_step) & 15) == 0;
    const bool v4 = (rb & 3) == 0 && ((reinterpret_cast<uintptr_t>(D.src) | reinterpret_cast<uintptr_t>(D.dst) |
                                        (uintptr_t)D.stride_env | (uintptr_t)D.stride_step) & 3) == 0;
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < G.num_rows; i += warps) {
        const int64_t n = __ldg(G.indices + i);
        const int64_t e = n / G.T, t = n - e * G.T;
        const uint8_t* s = D.src + e * D.stride_env + t * D.stride_step;
        uint8_t* d = D.dst + i * (int64_t)rb;
        if (v16) {
            for (int k = lane; k < (rb >> 4); k += 32) reinterpret_cast<uint4*>(d)[k] = __ldg(reinterpret_cast<const uint4*>(s) + k);
        } else if (v4) {
            for (int k = lane; k < (rb >> 2); k += 32) reinterpret_cast<uint32_t*>(d)[k] = __ldg(reinterpret_cast<const uint32_t*>(s) + k);
        } else {
            for (int k = lane; k < rb; k += 32) d[k] = __ldg(s + k);
        }
    }
}

}  // namespace
