// hs_samplers.cuh -- counter-based Philox, device-side reset sampler, HideAndSeek_envgen control-plane kernels (archive perturbation sampler, farthest point sampling)
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once
#include "hs_common.cuh"

namespace {


// ---------------------------------------------------------------------------------------------
// Device-side reset sampler.  One thread per env; counter-based Philox4x32-10
// (Salmon et al., SC'11) so a draw depends only on (seed, epoch, global env index).  Restated on
// the CPU in oracle/reset_sampler.py (bit-exact bar for everything but sinf/cosf).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

struct PhiloxStream {
    uint4 buf;
    uint4 ctr;
    uint2 key;
    int used;
    __device__ __forceinline__ uint32_t next() {
        if (used == 4) {
            buf = philox4x32_10(ctr, key);
            ctr.y += 1;
            used = 0;
        }
        const uint32_t v = used == 0 ? buf.x : used == 1 ? buf.y : used == 2 ? buf.z : buf.w;
        ++used;
        return v;
    }
    __device__ __forceinline__ float uniform(float lo, float hi) {
        const float u = __fmul_rn((float)(next() >> 8), 5.9604644775390625e-08f);     // 2^-24, exact
        return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));
    }
};

constexpr int RS_MAX_GRID = 11;                    // num_grid^2 <= 121 bits
constexpr int RS_WORDS = 4;

__global__ void __launch_bounds__(128) hs_reset_sample_kernel(hs_reset_dist d, int E, int A, int C, uint64_t epoch,
                                                              float* __restrict__ drone_pos, float* __restrict__ drone_rot,
                                                              float* __restrict__ target_pos, float* __restrict__ cyl_pos,
                                                              float* __restrict__ n_active_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    PhiloxStream rng;
    rng.ctr = make_uint4((uint32_t)((uint64_t)d.env_offset + (uint64_t)e), 0u, (uint32_t)epoch, (uint32_t)(epoch >> 32));
    rng.key = make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32));
    rng.used = 4;

    const int ng = d.num_grid, half = ng / 2;
    // occupancy bits (1 = free): inside the circle of radius num_grid/2 cells, hideandseek.py:168-181
    uint32_t freew[RS_WORDS] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < ng; ++i)
        for (int j = 0; j < ng; ++j)
            if ((i - half) * (i - half) + (j - half) * (j - half) < half * half) {
                const int b = i * ng + j;
                freew[b >> 5] |= 1u << (b & 31);
            }
    auto occupy = [&](float x, float y) {          // continuous_to_grid, hideandseek.py:144-166
        int gx = (int)rintf(__fdiv_rn(x, d.grid_size)) + half;
        int gy = (int)rintf(__fdiv_rn(y, d.grid_size)) + half;
        gx = min(max(gx, 0), ng - 1);
        gy = min(max(gy, 0), ng - 1);
        const int b = gx * ng + gy;
        freew[b >> 5] &= ~(1u << (b & 31));
    };

    float dxy[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (a < A) {
            dxy[a][0] = rng.uniform(d.drone_lo[0], d.drone_hi[0]);
            dxy[a][1] = rng.uniform(d.drone_lo[1], d.drone_hi[1]);
        }
    float tx = rng.uniform(d.target_lo[0], d.target_hi[0]);
    float ty = rng.uniform(d.target_lo[1], d.target_hi[1]);
    if (d.fixed_xy) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (a < A) { dxy[a][0] = d.fixed_drone_xy[a][0]; dxy[a][1] = d.fixed_drone_xy[a][1]; }
        tx = d.fixed_target_xy[0];
        ty = d.fixed_target_xy[1];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (a < A) {
            float* p = drone_pos + ((size_t)e * A + a) * 3;
            p[0] = dxy[a][0];
            p[1] = dxy[a][1];
            p[2] = rng.uniform(d.z_lo, d.z_hi);
            occupy(dxy[a][0], dxy[a][1]);
        }
    target_pos[(size_t)e * 3 + 0] = tx;
    target_pos[(size_t)e * 3 + 1] = ty;
    target_pos[(size_t)e * 3 + 2] = rng.uniform(d.z_lo, d.z_hi);
    occupy(tx, ty);

    const uint32_t wn = rng.next();
    const int n_active = d.fixed_num >= 0 ? d.fixed_num : d.min_cylinders + (int)__umulhi(wn, (uint32_t)(C + 1 - d.min_cylinders));
    if (n_active_out) n_active_out[e] = (float)n_active;

    int nfree = 0;
#pragma unroll
    for (int w = 0; w < RS_WORDS; ++w) nfree += __popc(freew[w]);
    // max_num distinct free cells, uniformly without replacement (hideandseek.py:106-119):
    // the k-th draw takes the r-th still-free cell in ascending cell index, r uniform in [0, free-k)
    for (int k = 0; k < C; ++k) {
        int r = (int)__umulhi(rng.next(), (uint32_t)(nfree - k));
        int cellidx = 0;
#pragma unroll
        for (int w = 0; w < RS_WORDS; ++w) {
            const int c = __popc(freew[w]);
            if (r >= 0 && r < c) {
                const int bit = (int)__fns(freew[w], 0, r + 1);
                cellidx = w * 32 + bit;
                freew[w] &= ~(1u << bit);
                r = -1;
            } else if (r >= 0) {
                r -= c;
            }
        }
        const int gx = cellidx / ng, gy = cellidx - gx * ng;
        float x = __fmul_rn((float)(gx - half), d.grid_size);      // grid_to_continuous, hideandseek.py:120-142
        float y = __fmul_rn((float)(gy - half), d.grid_size);
        x = fminf(fmaxf(x, -d.boundary), d.boundary);
        y = fminf(fmaxf(y, -d.boundary), d.boundary);
        float* p = cyl_pos + ((size_t)e * C + k) * 3;
        p[0] = x;
        p[1] = y;
        p[2] = k >= n_active ? d.cyl_z_inactive : d.cyl_z_active;
    }

#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (a < A) {
            const float hr = 0.5f * rng.uniform(d.rpy_lo[0], d.rpy_hi[0]);
            const float hp = 0.5f * rng.uniform(d.rpy_lo[1], d.rpy_hi[1]);
            const float hy = 0.5f * rng.uniform(d.rpy_lo[2], d.rpy_hi[2]);
            float sr, cr, sp, cp, sy, cy;                       // euler_to_quaternion (utils/torch.py), wxyz
            sincosf(hr, &sr, &cr);
            sincosf(hp, &sp, &cp);
            sincosf(hy, &sy, &cy);
            float4 q = make_float4(cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                                   cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy);
            *reinterpret_cast<float4*>(drone_rot + ((size_t)e * A + a) * 4) = q;
        }
}


// ---------------------------------------------------------------------------------------------
// HideAndSeek_envgen control plane: archive perturbation sampler and
// farthest point sampling.  Restated on the CPU in oracle/envgen_oracle.py (bit-exact bar).
// ---------------------------------------------------------------------------------------------
constexpr int GEN_MAX_DIM = 3 * 3 + 3 + 3 * CMAX;

struct GenBounds { float lo[GEN_MAX_DIM], hi[GEN_MAX_DIM]; };

__global__ void __launch_bounds__(128) hs_gen_sample_nearby_kernel(hs_gen_params g, GenBounds B, const float* __restrict__ history,
                                                                   int64_t n_history, int64_t num_tasks, uint64_t epoch,
                                                                   float* __restrict__ tasks_out, uint8_t* __restrict__ valid_out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tasks) return;
    const int A = g.num_agents, C = g.num_cylinders;
    const int nb = 3 * A + 3, dim = nb + 3 * C;
    const int ng = g.num_grid, half = ng / 2;
    const uint64_t key64 = g.seed ^ 0x9E3779B97F4A7C15ull;
    const uint2 key = make_uint2((uint32_t)key64, (uint32_t)(key64 >> 32));
    uint32_t inside[RS_WORDS] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < ng; ++i)
        for (int j = 0; j < ng; ++j)
            if ((i - half) * (i - half) + (j - half) * (j - half) < half * half) {
                const int b = i * ng + j;
                inside[b >> 5] |= 1u << (b & 31);
            }
    const uint32_t tg = (uint32_t)((uint64_t)g.task_offset + (uint64_t)t);       // global task index = counter word 0
    const uint4 w0 = philox4x32_10(make_uint4(tg, 0xFFFF0000u, (uint32_t)epoch, (uint32_t)(epoch >> 32)), key);
    const int64_t idx = (int64_t)__umulhi(w0.x, (uint32_t)n_history);
    float origin[GEN_MAX_DIM], cand[GEN_MAX_DIM];
    for (int j = 0; j < dim; ++j) origin[j] = history[idx * dim + j];
    bool ok = false;
    for (int attempt = 0; attempt < 10 && !ok; ++attempt) {
        PhiloxStream rng;
        rng.ctr = make_uint4(tg, (uint32_t)(64 * attempt), (uint32_t)epoch, (uint32_t)(epoch >> 32));
        rng.key = key;
        rng.used = 4;
        for (int j = 0; j < nb; ++j) {
            const float u = __fmul_rn((float)(rng.next() >> 8), 5.9604644775390625e-08f);
            const float noise = __fmul_rn(__fadd_rn(-1.0f, __fmul_rn(2.0f, u)), g.expand_step);
            cand[j] = __fadd_rn(origin[j], noise);
        }
        for (int c = 0; c < C; ++c) {
            for (int a = 0; a < 2; ++a) {
                const int s = (int)__umulhi(rng.next(), 3u) - 1;
                cand[nb + 3 * c + a] = g.expand_cylinders ? __fadd_rn(origin[nb + 3 * c + a], __fmul_rn((float)s, g.grid_size))
                                                          : origin[nb + 3 * c + a];
            }
            cand[nb + 3 * c + 2] = origin[nb + 3 * c + 2];
        }
        for (int j = 0; j < dim; ++j) cand[j] = fminf(fmaxf(cand[j], B.lo[j]), B.hi[j]);
        // sanity_check: every object on its own free cell
        uint32_t freew[RS_WORDS] = {inside[0], inside[1], inside[2], inside[3]};
        ok = true;
        for (int o = 0; o < A + 1 + C; ++o) {
            const int base = 3 * o;
            int gx = (int)rintf(__fdiv_rn(cand[base], g.grid_size)) + half;
            int gy = (int)rintf(__fdiv_rn(cand[base + 1], g.grid_size)) + half;
            gx = min(max(gx, 0), ng - 1);
            gy = min(max(gy, 0), ng - 1);
            const int b = gx * ng + gy;
            const uint32_t bit = 1u << (b & 31);
            uint32_t wsel = 0u;
#pragma unroll
            for (int w = 0; w < RS_WORDS; ++w) if (w == (b >> 5)) wsel = freew[w];
            if (!(wsel & bit)) { ok = false; break; }
#pragma unroll
            for (int w = 0; w < RS_WORDS; ++w) if (w == (b >> 5)) freew[w] &= ~bit;
        }
    }
    for (int j = 0; j < dim; ++j) tasks_out[t * dim + j] = cand[j];
    valid_out[t] = ok ? 1 : 0;
}

// Farthest point sampling: every CTA owns a contiguous chunk of the points (cached in shared memory
// when it fits), keeps their running minimum distance in global scratch, and proposes its local
// argmax; one grid barrier per selected point, then every CTA reduces the proposals redundantly.
// Key = (float bits of the distance << 32) | ~index: the maximum key is the maximum distance and, among
// equal distances, the LOWEST index - numpy's argmax.
constexpr int FPS_THREADS = 256;
__device__ __forceinline__ unsigned long long fps_block_max(unsigned long long v, unsigned long long* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    v = sh[0];
#pragma unroll
    for (int w = 1; w < FPS_THREADS / 32; ++w) v = sh[w] > v ? sh[w] : v;
    return v;
}

__global__ void __launch_bounds__(FPS_THREADS) hs_fps_kernel(const float* __restrict__ pts, int n, int dim, int k, int start,
                                                             int chunk, int cache_pts, int32_t* __restrict__ out_idx,
                                                             float* __restrict__ mind, unsigned long long* slots,
                                                             unsigned int* bar) {
    extern __shared__ __align__(16) float fps_smem[];
    __shared__ unsigned long long red[FPS_THREADS / 32];
    __shared__ float q[64];
    const int tid = threadIdx.x, G = gridDim.x, c = blockIdx.x;
    const int lo = c * chunk, hi = min(n, lo + chunk);
    if (cache_pts)
        for (int i = tid; i < (hi - lo) * dim; i += FPS_THREADS) fps_smem[i] = pts[(size_t)lo * dim + i];
    for (int p = lo + tid; p < hi; p += FPS_THREADS) mind[p] = __int_as_float(0x7f800000);
    __syncthreads();
    int cur = start;
    for (int it = 0; it < k; ++it) {
        if (c == 0 && tid == 0) out_idx[it] = cur;
        if (it + 1 == k) break;
        if (tid < dim) q[tid] = pts[(size_t)cur * dim + tid];
        __syncthreads();
        unsigned long long best = 0ull;
        for (int p = lo + tid; p < hi; p += FPS_THREADS) {
            const float* x = cache_pts ? (fps_smem + (size_t)(p - lo) * dim) : (pts + (size_t)p * dim);
            float d = 0.f;
            for (int j = 0; j < dim; ++j) {
                const float df = __fsub_rn(x[j], q[j]);
                d = __fadd_rn(d, __fmul_rn(df, df));
            }
            const float m = fminf(mind[p], d);
            mind[p] = m;
            const unsigned long long key = ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)p);
            best = key > best ? key : best;
        }
        best = fps_block_max(best, red);
        unsigned long long* sl = slots + (size_t)(it & 1) * G;
        if (tid == 0) {
            *reinterpret_cast<volatile unsigned long long*>(sl + c) = best;
            __threadfence();
            // grid barrier (all CTAs are co-resident: cooperative launch)
            const unsigned int gen = *reinterpret_cast<volatile unsigned int*>(bar + 1);
            if (atomicAdd(bar, 1u) == (unsigned)(G - 1)) {
                *reinterpret_cast<volatile unsigned int*>(bar) = 0u;
                __threadfence();
                atomicAdd(bar + 1, 1u);
            } else {
                while (*reinterpret_cast<volatile unsigned int*>(bar + 1) == gen) { }
            }
            __threadfence();
        }
        __syncthreads();
        unsigned long long v = 0ull;
        for (int i = tid; i < G; i += FPS_THREADS) {
            const unsigned long long o = *reinterpret_cast<volatile unsigned long long*>(sl + i);
            v = o > v ? o : v;
        }
        v = fps_block_max(v, red);
        cur = (int)(0xFFFFFFFFu - (unsigned)(v & 0xFFFFFFFFull));
    }
}

}  // namespace
