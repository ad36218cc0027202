// hs_kernels.cu -- sm_100a kernels + C ABI for the HideAndSeek environment tick.
//
// One fused kernel per control tick (hs_tick_kernel) covers what the reference does with
// ~300 eager torch launches plus a PhysX step (see include/hs_b200.h for the file:line map).
//
// Work decomposition (B200: 148 SMs, 32-wide warps):
//   * a GROUP of G=4 adjacent lanes owns one environment: lanes 0..A-1 are the pursuers,
//     lane A is the evader.  A warp therefore advances 8 environments; all cross-agent
//     terms (downwash all-pairs, evader repulsion sum, capture/detect "any", per-env
//     means for the stats) are warp-shuffle exchanges inside the group -- no shared
//     memory round trip and no atomics.
//   * state lives in a private SoA arena [row][E] so that the 8 envs of a warp are 8
//     consecutive floats (one 32 B sector) per row and slot.
//   * reference-facing outputs are AoS ([E,A,W] row-major).  A warp's 8 envs are ONE
//     contiguous span of every such tensor, so wide rows (W >= 6) are staged in shared
//     memory and leave through a single TMA bulk store (cp.async.bulk.global.shared::cta,
//     SASS UBLKCP) per tensor per warp; narrow rows (W <= 4) are written directly
//     (float4 / scalar), which is already sector-exact.
//   * no tensor cores: there is no dense contraction on this path.
//
// Numerics: fp32 throughout.  Products/sums may contract to FFMA; divisions and square roots
// use the SFU approximations (rcp/sqrt.approx, <= 2 ulp) because the tick is instruction-bound,
// not bandwidth-bound (profiles/): that keeps every output within ~1e-6 relative of the eager
// reference, two orders of magnitude inside the 1e-4 parity bar.  tanh/sin/cos stay accurate.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <new>

#include "hs_b200.h"

#ifndef HS_USE_BULK_STORE
#define HS_USE_BULK_STORE 1
#endif

namespace {

constexpr int G = 4;                 // lanes per environment
constexpr int ENVS_PER_WARP = 32 / G;
constexpr unsigned FULL = 0xffffffffu;
constexpr int CMAX = HS_MAX_CYLINDERS;
constexpr int KMAX = HS_MAX_OBS_CYLINDERS;
constexpr int FMAX = HS_MAX_FUTURE;
constexpr int ND = 23;               // per-drone arena scalars
// per-drone scalar ids
enum { D_POS = 0, D_ROT = 3, D_LIN = 7, D_ANG = 10, D_THR = 13, D_INT = 17, D_LAST = 20 };
// per-env rows that follow the 23*A drone rows
enum { E_TPOS = 0, E_TVEL = 3, E_PROGRESS = 6, E_BDETECT = 7, E_CYL = 8 };

struct KParams {
    hs_config c;
    hs_buffers b;
    int64_t Ep;                      // arena row pitch (E rounded up to 32)
    const float* action;
    const uint8_t* reset_pid;
    const uint8_t* env_mask;
    const float* init_drone_pos;
    const float* init_drone_rot;
    const float* init_target_pos;
    const float* init_cyl_pos;
    const float* tp_pred;
    int action_is_raw;
    int tp_init;                     // 1: first frame ever -> fill all H history rows
};

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { const float r = frcp(s); return mk(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float norm3(V3 a) { return fsqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// torch.clamp semantics (NaN propagates)
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// omni_drones/utils/torch.py:182-201 -- a (+/-) b + c with the same grouping
template <bool INV>
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    const V3 u = mk(q.x, q.y, q.z);
    const float s = 2.0f * (q.w * q.w) - 1.0f;
    const V3 a = v * s;
    const V3 b = (cross3(u, v) * q.w) * 2.0f;
    const V3 c = (u * dot3(u, v)) * 2.0f;
    return INV ? ((a - b) + c) : ((a + b) + c);
}
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    Q4 r;
    r.w = ((a.w * b.w - a.x * b.x) - a.y * b.y) - a.z * b.z;
    r.x = ((a.w * b.x + a.x * b.w) + a.y * b.z) - a.z * b.y;
    r.y = ((a.w * b.y - a.x * b.z) + a.y * b.w) + a.z * b.x;
    r.z = ((a.w * b.z + a.x * b.y) - a.y * b.x) + a.z * b.w;
    return r;
}

// Round-to-nearest primitives that the compiler may not contract or approximate.  Used where
// the reference's arithmetic has catastrophic cancellation that amplifies 1-ulp differences
// (the D term of the rate PID: (rate - last_rate)/dt * kd, gain ~1.4e4 on the body rate).
namespace ex {
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
}  // namespace ex
// omni_drones/utils/torch.py:193-201 with the reference's exact operation order and rounding
__device__ __forceinline__ V3 qrot_inv_exact(Q4 q, V3 v) {
    using namespace ex;
    const float s = sub(mul(2.0f, mul(q.w, q.w)), 1.0f);
    const V3 a = mk(mul(v.x, s), mul(v.y, s), mul(v.z, s));
    const V3 cr = mk(sub(mul(q.y, v.z), mul(q.z, v.y)), sub(mul(q.z, v.x), mul(q.x, v.z)), sub(mul(q.x, v.y), mul(q.y, v.x)));
    const V3 b = mk(mul(mul(cr.x, q.w), 2.0f), mul(mul(cr.y, q.w), 2.0f), mul(mul(cr.z, q.w), 2.0f));
    const float d = add(add(mul(q.x, v.x), mul(q.y, v.y)), mul(q.z, v.z));
    const V3 c = mk(mul(mul(q.x, d), 2.0f), mul(mul(q.y, d), 2.0f), mul(mul(q.z, d), 2.0f));
    return mk(add(sub(a.x, b.x), c.x), add(sub(a.y, b.y), c.y), add(sub(a.z, b.z), c.z));
}

// Ampere-style async copies global -> shared (SASS LDGSTS): prefetch without holding registers
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(sdst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(sdst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float gshfl(float v, int src_lane) { return __shfl_sync(FULL, v, src_lane); }
__device__ __forceinline__ V3 gshfl3(V3 v, int src_lane) {
    return mk(gshfl(v.x, src_lane), gshfl(v.y, src_lane), gshfl(v.z, src_lane));
}

// ---- shared-memory staging + TMA bulk store --------------------------------------------
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

// Per-warp staging: two buffers used alternately so that filling tile n+1 overlaps the
// bulk store of tile n.
struct Stager {
    float* buf[2];
    int cur;
    int lane;
    __device__ __forceinline__ float* begin() {
        // the store issued two flushes ago read from buf[cur]; wait until it has
        if (HS_USE_BULK_STORE) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
        }
        return buf[cur];
    }
    // all lanes have written their part of buf[cur]; send nwords to gdst
    __device__ __forceinline__ void flush(float* gdst, int nwords, bool full_tile) {
        float* s = buf[cur];
        const bool bulk = HS_USE_BULK_STORE && full_tile && ((nwords & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0);
        if (bulk) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store(gdst, s, static_cast<uint32_t>(nwords) * 4u);
                bulk_commit();
            }
        } else {
            __syncwarp();
            for (int i = lane; i < nwords; i += 32) gdst[i] = s[i];
            __syncwarp();
            if (HS_USE_BULK_STORE && lane == 0) bulk_commit();   // keep group parity
        }
        cur ^= 1;
    }
    __device__ __forceinline__ void finish() {
        if (HS_USE_BULK_STORE) {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
        }
    }
};

constexpr int FILL_STAGE_WORDS = ENVS_PER_WARP * HS_MAX_AGENTS * (20 + 3 * FMAX);   // 8*3*44 = 1056
constexpr int TICK_STAGE_WORDS = ENVS_PER_WARP * HS_MAX_AGENTS * 20;                // widest tick tile: [24][20]
constexpr int TP_ENV_WORDS_MAX = 192;                                               // history_step * (7+3A) <= 192

// ---- line of sight, hideandseek.py:47-103 ------------------------------------------------
template <int CT>
__device__ __forceinline__ bool los_blocked(const V3 p, const V3 t, const float (&cx)[CT],
                                            const float (&cy)[CT], const float (&cz)[CT],
                                            int C, float size) {
    const float ddx = p.x - t.x, ddy = p.y - t.y;
    // dist/(seg+eps) <= size  and  0 <= num/(den+eps) <= 1  with the (positive) denominators
    // multiplied out: no division per cylinder; underground (inactive) cylinders are skipped
    const float seg_sz = (fsqrt(ddx * ddx + ddy * ddy) + 1e-5f) * size;
    const float dx = t.x - p.x, dy = t.y - p.y;
    const float den = (dx * dx + dy * dy) + 1e-5f;
    bool blocked = false;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        if (c < C && cz[c] > 0.0f) {
            const float ccx = cx[c] - t.x, ccy = cy[c] - t.y;
            const float cr = fabsf(ddx * ccy - ddy * ccx);
            const float num = (cx[c] - p.x) * dx + (cy[c] - p.y) * dy;
            blocked = blocked || ((cr <= seg_sz) && (num >= 0.0f) && (num <= den));
        }
    }
    return blocked;
}

// heading = R x, up = R z (utils/torch.py:221-225 evaluated on a basis vector)
__device__ __forceinline__ void heading_up(Q4 q, V3& heading, V3& up) {
    heading = qrot<false>(q, mk(1.0f, 0.0f, 0.0f));
    up = qrot<false>(q, mk(0.0f, 0.0f, 1.0f));
}

// Writes one [*, D] row: [head3, (p - pred_f) x F, quat4, linvel3, heading3, up3, t x4]
__device__ __forceinline__ void write_self_row(float* row, V3 head, int F3, const float* rp,
                                               Q4 q, V3 v, V3 heading, V3 up, float t) {
    row[0] = head.x; row[1] = head.y; row[2] = head.z;
    int o = 3;
    for (int i = 0; i < F3; ++i) row[o + i] = rp[i];
    o += F3;
    row[o + 0] = q.w; row[o + 1] = q.x; row[o + 2] = q.y; row[o + 3] = q.z;
    row[o + 4] = v.x; row[o + 5] = v.y; row[o + 6] = v.z;
    row[o + 7] = heading.x; row[o + 8] = heading.y; row[o + 9] = heading.z;
    row[o + 10] = up.x; row[o + 11] = up.y; row[o + 12] = up.z;
    row[o + 13] = t; row[o + 14] = t; row[o + 15] = t; row[o + 16] = t;
}

#define AROW(r) (P.b.arena + (int64_t)(r) * P.Ep + e)
#define DROW(k) AROW((k) * A + slot)
#define EROW(k) AROW(ND * A + (k))

// =========================================================================================
// The tick.  RESET=false: full control tick.  RESET=true: the unforced physics tick + obs
// that closes a reset (hideandseek.py:722-723, isaac_env.py:220-224).
// =========================================================================================
template <int A, bool RESET, int CT>
__global__ void __launch_bounds__(128, 5)
hs_tick_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(128) float stage_mem[4][2][TICK_STAGE_WORDS];
    __shared__ __align__(128) float tp_mem[4][ENVS_PER_WARP * TP_ENV_WORDS_MAX];   // TP_input tile of the warp
    __shared__ __align__(16) float stat_mem[4][ENVS_PER_WARP][HS_NUM_STATS];

    const hs_config& c = P.c;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slot = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const int64_t e0 = warp_g * ENVS_PER_WARP;           // first env of this warp
    const int E = c.num_envs;
    if (e0 >= E) return;                                 // whole warp out of range
    const int64_t e_raw = e0 + (lane >> 2);
    const bool valid = e_raw < E;
    const int64_t e = valid ? e_raw : (E - 1);           // clamp: idle lanes shadow the last env (no stores)
    const bool is_drone = slot < A;
    const bool is_ev = slot == A;
    const int nenv = (int)min((int64_t)ENVS_PER_WARP, E - e0);
    const bool full_tile = nenv == ENVS_PER_WARP;
    const int C = c.num_cylinders, K = c.obs_max_cylinder;
    const int FD = 7 + 3 * A;
    const int H = c.history_step;
    const float dt = c.dt;
    // arena offsets fit 32 bits (checked at hs_create): one IMAD + one wide add per access
    const uint32_t Ep32 = (uint32_t)P.Ep;
    const uint32_t o_drone = (uint32_t)slot * Ep32 + (uint32_t)e;     // + k * (A*Ep32)
    const uint32_t o_env = (uint32_t)(ND * A) * Ep32 + (uint32_t)e;   // + k * Ep32
    float* const arena = P.b.arena;
#undef DROW
#undef EROW
#define DROW(k) (arena + (o_drone + (uint32_t)(k) * ((uint32_t)A * Ep32)))
#define EROW(k) (arena + (o_env + (uint32_t)(k) * Ep32))

    Stager st;
    st.buf[0] = stage_mem[wib][0];
    st.buf[1] = stage_mem[wib][1];
    st.cur = 0;
    st.lane = lane;

    // ---- prefetch (no registers held): the previous TP_input rows 1..H-1 land in the warp's
    // shared tile already shifted to rows 0..H-2, and the env's stats row lands in stat_mem.
    float* tp_tile = tp_mem[wib];
    const int per_env = H * FD, keep = (H - 1) * FD;
    if (c.use_tp_net && !P.tp_init) {
        const float* src = P.b.tp_input_prev + e0 * per_env;
        if ((FD & 3) == 0 && H == 10 && full_tile) {
            // common shape (A=3, H=10): 8 envs x 36 float4 = 9 per lane, all indices compile-time
            constexpr int fd4 = FD / 4, pe4 = 10 * fd4, keep4 = 9 * fd4;
#pragma unroll
            for (int it = 0; it < (ENVS_PER_WARP * keep4 + 31) / 32; ++it) {
                const int i = it * 32 + lane;
                const int env = i / keep4, j = i - env * keep4;      // division by a constant
                if (i < ENVS_PER_WARP * keep4)
                    cp_async16(reinterpret_cast<float4*>(tp_tile) + env * pe4 + j,
                               reinterpret_cast<const float4*>(src) + env * pe4 + j + fd4);
            }
        } else if ((FD & 3) == 0) {
            const int pe4 = per_env >> 2, keep4 = keep >> 2, fd4 = FD >> 2;
            const int total = nenv * keep4;
            int env = 0, j = lane;                       // i = env*keep4 + j, kept incrementally
            for (int i = lane; i < total; i += 32, j += 32) {
                while (j >= keep4) { j -= keep4; ++env; }
                cp_async16(reinterpret_cast<float4*>(tp_tile) + env * pe4 + j,
                           reinterpret_cast<const float4*>(src) + env * pe4 + j + fd4);
            }
        } else {
            const int total = nenv * keep;
            int env = 0, j = lane;
            for (int i = lane; i < total; i += 32, j += 32) {
                while (j >= keep) { j -= keep; ++env; }
                cp_async4(tp_tile + env * per_env + j, src + env * per_env + j + FD);
            }
        }
    }
    if (!RESET && valid && is_ev) {
#pragma unroll
        for (int k = 0; k < HS_NUM_STATS; ++k) cp_async4(&stat_mem[wib][lane >> 2][k], P.b.stats + (int64_t)k * E + e);
    }
    cp_async_commit();

    // ---- load state ------------------------------------------------------------------
    V3 p = mk(0, 0, 0), lv = mk(0, 0, 0), av = mk(0, 0, 0);
    Q4 q; q.w = 1.f; q.x = q.y = q.z = 0.f;
    float thr[4] = {0, 0, 0, 0};
    V3 integ = mk(0, 0, 0), last = mk(0, 0, 0);
    if (is_drone) {
        p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
        q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
        lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
        av = mk(*DROW(D_ANG), *DROW(D_ANG + 1), *DROW(D_ANG + 2));
        if (!RESET) {
#pragma unroll
            for (int k = 0; k < 4; ++k) thr[k] = *DROW(D_THR + k);
            integ = mk(*DROW(D_INT), *DROW(D_INT + 1), *DROW(D_INT + 2));
            last = mk(*DROW(D_LAST), *DROW(D_LAST + 1), *DROW(D_LAST + 2));
        }
    }
    // per-env scalars: every lane of the group reads the same address (one broadcast sector)
    V3 tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
    V3 tv = mk(*EROW(E_TVEL), *EROW(E_TVEL + 1), *EROW(E_TVEL + 2));
    float progress = *EROW(E_PROGRESS);
    float4 act4 = make_float4(0.f, 0.f, 0.f, 0.f), prev4 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool pid_reset = false;
    float v_prey = 0.f;
    if (!RESET) {
        if (is_drone) {
            const int64_t row = e * A + slot;
            act4 = __ldg(reinterpret_cast<const float4*>(P.action) + row);
            if (P.action_is_raw) {
                prev4 = *(reinterpret_cast<const float4*>(P.b.prev_action) + row);
                pid_reset = (P.reset_pid != nullptr) && (P.reset_pid[e] != 0);
            }
        }
        if (is_ev) v_prey = __ldg(P.b.v_prey);
    }
    float cx[CT], cy[CT], cz[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        if (k < C) {
            cx[k] = __ldg(EROW(E_CYL + 3 * k));
            cy[k] = __ldg(EROW(E_CYL + 3 * k + 1));
            cz[k] = __ldg(EROW(E_CYL + 3 * k + 2));
        } else { cx[k] = 0.f; cy[k] = 0.f; cz[k] = -20.f; }
    }

    float action_err = 0.f, throttle_diff = 0.f;
    float T[4] = {0, 0, 0, 0};
    float yaw_torque = 0.f;
    V3 ext = mk(0, 0, 0);
    bool out_of_arena = false;

    if (!RESET) {
        // ---- CTBR transform + body-rate PID (transforms.py:425-459, lee_position_controller.py:476-550)
        float cmd[4] = {0, 0, 0, 0};
        if (is_drone) {
            const int64_t row = e * A + slot;
            const float4 act = act4;
            if (P.action_is_raw) {
                using namespace ex;
                const float4 prev = prev4;
                const float a0 = tanhf(act.x), a1 = tanhf(act.y), a3 = tanhf(act.w);
                float a2 = tanhf(act.z);
                const float thrust = clampf(mul(add(a3, 1.0f), 0.5f), 0.0f, c.max_thrust_ratio);
                if (c.fixed_yaw) a2 = 0.0f;
                const float d0 = sub(a0, prev.x), d1 = sub(a1, prev.y), d2 = sub(a2, prev.z), d3 = sub(thrust, prev.w);
                action_err = __fsqrt_rn(add(add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)), mul(d3, d3)));
                if (valid) *(reinterpret_cast<float4*>(P.b.prev_action) + row) = make_float4(a0, a1, a2, thrust);
                const V3 trate = mk(mul(mul(a0, 180.0f), c.target_clip), mul(mul(a1, 180.0f), c.target_clip),
                                    mul(mul(a2, 180.0f), c.target_clip));
                const float tthrust = mul(thrust, 65536.0f);
                if (pid_reset) { integ = mk(0, 0, 0); last = mk(0, 0, 0); }
                const V3 br0 = qrot_inv_exact(q, av);
                const float pi_f = 3.14159265358979323846f;
                const V3 br = mk(div(mul(br0.x, 180.0f), pi_f), div(mul(br0.y, 180.0f), pi_f), div(mul(br0.z, 180.0f), pi_f));
                float o[3];
                const float errv[3] = {sub(trate.x, br.x), sub(trate.y, br.y), sub(trate.z, br.z)};
                const float brv[3] = {br.x, br.y, br.z};
                const float lastv[3] = {last.x, last.y, last.z};
                float integv[3] = {integ.x, integ.y, integ.z};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float outP = mul(errv[k], c.pid_kp[k]);
                    float deriv = div(-sub(brv[k], lastv[k]), dt);
                    if (isnan(deriv)) deriv = 0.0f;
                    const float outD = mul(deriv, c.pid_kd[k]);
                    integv[k] = clampf(add(integv[k], mul(errv[k], dt)), -c.pid_ilimit[k], c.pid_ilimit[k]);
                    const float outI = mul(integv[k], c.pid_ki[k]);
                    float out = add(add(outP, outD), outI);
                    if (isnan(out)) out = 0.0f;
                    o[k] = clampf(out, -c.pid_out_limit, c.pid_out_limit);
                }
                integ = mk(integv[0], integv[1], integv[2]);
                last = br;
                const float r = o[0] * 0.5f, pp = o[1] * 0.5f, y = o[2];
                const float m[4] = {add(sub(add(tthrust, r), pp), y), sub(add(add(tthrust, r), pp), y),
                                    add(add(sub(tthrust, r), pp), y), sub(sub(sub(tthrust, r), pp), y)};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v = sub(mul(mul(m[k], 1.0f / 65536.0f), 2.0f), c.max_thrust_ratio);
                    if (isnan(v)) v = 0.0f;                       // torch.nan_to_num_(cmds, 0.)
                    else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
                    cmd[k] = v;
                }
                if (valid) {
                    *(reinterpret_cast<float4*>(P.b.rotor_cmds) + row) = make_float4(cmd[0], cmd[1], cmd[2], cmd[3]);
                    *(reinterpret_cast<float4*>(P.b.ctbr) + row) = make_float4(r, pp, y, tthrust);
                    P.b.target_rate[row * 3 + 0] = trate.x;
                    P.b.target_rate[row * 3 + 1] = trate.y;
                    P.b.target_rate[row * 3 + 2] = trate.z;
                    P.b.action_error[row] = action_err;
                }
            } else {
                cmd[0] = act.x; cmd[1] = act.y; cmd[2] = act.z; cmd[3] = act.w;
                action_err = P.b.action_error[row];
            }
            // ---- rotor model, rotor_group.py:55-71
            float dsq = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float target = fsqrt(clampf((cmd[k] + 1.0f) / 2.0f, 0.0f, 1.0f));
                const float nt = thr[k] + c.rotor_alpha * (target - thr[k]);
                const float dth = nt - thr[k];
                dsq = (k == 0) ? dth * dth : dsq + dth * dth;
                thr[k] = nt;
                const float t = clampf(nt * nt + 0.0f, 0.0f, 1.0f);
                T[k] = t * c.kf;
                const float mom = (t * c.km) * (-c.rotor_dirs[k]);
                yaw_torque = (k == 0) ? mom : yaw_torque + mom;
            }
            throttle_diff = fsqrt(dsq);
        }
        // ---- downwash all-pairs, multirotor.py:488-494, 724-753
        const float total_thrust = ((T[0] + T[1]) + T[2]) + T[3];
        const V3 Fw = qrot<false>(q, mk(0.f, 0.f, total_thrust));
        V3 dw = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const V3 Fj = gshfl3(Fw, gbase + j);
            const V3 pj = gshfl3(p, gbase + j);
            if (is_drone && j != slot) {
                const V3 d = Fj / (norm3(Fj) + 1e-6f);
                const V3 rel = pj - p;
                const float zd = dot3(rel, d);
                const float rr = norm3(rel - d * zd);
                const float z = zd < 0.0f ? 0.0f : zd;
                const float qq = fdiv(c.downwash_kr * rr, z);
                const float den = 1.0f + c.downwash_kz * z;
                const float v = fdiv(__expf(-0.5f * (qq * qq)), den * den);
                dw = dw + neg(Fj) * v;
            }
        }
        ext = dw + lv * c.drag_coef_times_mass;

        // ---- evader, hideandseek.py:1067-1141 + 737-744
        V3 fp = mk(0.f, 0.f, 0.f);
        if (is_drone) {
            const V3 rel = p - tp;
            const float dist = norm3(rel);
            const bool blocked = los_blocked(p, tp, cx, cy, cz, C, c.cylinder_size);
            const float active = ((dist < c.target_detect_radius) && !blocked) ? 1.0f : 0.0f;
            const float inv_d = frcp(dist + 1e-5f);
            fp = (neg(rel) * (inv_d * inv_d)) * active;
        }
        V3 force = gshfl3(fp, gbase);
#pragma unroll
        for (int j = 1; j < A; ++j) force = force + gshfl3(fp, gbase + j);
        if (is_ev) {
            force = mk(0.f, 0.f, 0.f) + force;
            const float rho = fsqrt(tp.x * tp.x + tp.y * tp.y);
            const float inv_rho = frcp(rho + 1e-5f);
            const float inx = -tp.x * inv_rho, iny = -tp.y * inv_rho;
            out_of_arena = (tp.x * tp.x + tp.y * tp.y) > c.arena_size_sq;
            const float o = out_of_arena ? 1.0f : 0.0f, no = out_of_arena ? 0.0f : 1.0f;
            const float wall = frcp((c.arena_size - rho) + 1e-5f);
            V3 fr;
            fr.x = (o * inx) * 1e5f + (no * inx) * wall;
            fr.y = (o * iny) * 1e5f + (no * iny) * wall;
            const bool hi = tp.z > c.max_height;
            const float hz = c.max_height - tp.z;
            fr.z = hi ? -1e5f : fdiv(-hz, hz * hz + 1e-5f);
            const bool lo = tp.z < 0.0f;
            const float lz = 0.0f - tp.z;
            fr.z = fr.z + (lo ? 1e5f : fdiv(-lz, lz * lz + 1e-5f));
            force = force + fr;
            float fcx = 0.f, fcy = 0.f;
#pragma unroll
            for (int k = 0; k < CT; ++k) {
                if (k < C && !(cz[k] < 0.0f)) {
                    const float tx = tp.x - cx[k], ty = tp.y - cy[k];
                    const float dxy = fsqrt(tx * tx + ty * ty);
                    if (dxy < c.target_detect_radius) {
                        const float sc = frcp(dxy + 1e-5f) * frcp((dxy - c.cylinder_size) + 1e-5f);
                        fcx = fcx + tx * sc;
                        fcy = fcy + ty * sc;
                    }
                }
            }
            force = force + mk(fcx, fcy, 0.f);
            const float vp = v_prey;
            tv = mk(fdiv(vp * force.x, fabsf(force.x) + 1e-5f), fdiv(vp * force.y, fabsf(force.y) + 1e-5f),
                    fdiv(vp * force.z, fabsf(force.z) + 1e-5f));
        }
    }

    // ---- rigid-body integration (PhysX stand-in; oracle/hs_oracle.py rigid_body_step) ----
    if (is_drone) {
        V3 force = mk(0.f, 0.f, 0.f), tau = mk(0.f, 0.f, 0.f);
        if (!RESET) {
            const float total_thrust = ((T[0] + T[1]) + T[2]) + T[3];
            force = qrot<false>(q, mk(0.f, 0.f, total_thrust));
            tau.x = ((c.rotor_y[0] * T[0] + c.rotor_y[1] * T[1]) + c.rotor_y[2] * T[2]) + c.rotor_y[3] * T[3];
            tau.y = (((-c.rotor_x[0]) * T[0] + (-c.rotor_x[1]) * T[1]) + (-c.rotor_x[2]) * T[2]) + (-c.rotor_x[3]) * T[3];
            tau.z = yaw_torque;
            force = force + ext;
        }
        V3 acc = force / c.total_mass;
        acc.z = acc.z - c.gravity;
        V3 v = lv + acc * dt;
        const V3 I = mk(c.inertia[0], c.inertia[1], c.inertia[2]);
        V3 wb = qrot<true>(q, av);
        const V3 gyro = cross3(wb, mk(I.x * wb.x, I.y * wb.y, I.z * wb.z));
        const V3 tg = tau - gyro;
        wb = wb + mk(tg.x * c.inv_inertia[0], tg.y * c.inv_inertia[1], tg.z * c.inv_inertia[2]) * dt;
        V3 w = qrot<false>(q, wb);
        v = v * c.lin_damp_factor;
        w = w * c.ang_damp_factor;
        const float vn = norm3(v);
        if (vn > c.max_linear_velocity) v = v * fdiv(c.vmax_clamped, vn);
        float wn = norm3(w);
        if (wn > c.max_angular_velocity) w = w * fdiv(c.max_angular_velocity, wn);
        p = p + v * dt;
        wn = norm3(w);
        const float half = (0.5f * dt) * wn;
        const bool small = wn < 1e-6f;
        float sh, ch;
        sincosf(half, &sh, &ch);
        const float kk = small ? (0.5f * dt) : fdiv(sh, fmaxf(wn, 1e-6f));
        Q4 dq; dq.w = small ? 1.0f : ch; dq.x = w.x * kk; dq.y = w.y * kk; dq.z = w.z * kk;
        Q4 qn = qmul(dq, q);
        const float qinv = rsqrtf(((qn.w * qn.w + qn.x * qn.x) + qn.y * qn.y) + qn.z * qn.z);
        q.w = qn.w * qinv; q.x = qn.x * qinv; q.y = qn.y * qinv; q.z = qn.z * qinv;
        if (c.ground_clamp && p.z < c.ground_z) {
            p.z = c.ground_z;
            if (v.z < 0.0f) v.z = 0.0f;
        }
        lv = v; av = w;
    }
    if (is_ev) tp = tp + tv * dt;
    // everyone needs the evader's new position/velocity
    tp = gshfl3(tp, gbase + A);
    tv = gshfl3(tv, gbase + A);
    if (!RESET) progress = progress + 1.0f;
    else if (P.env_mask == nullptr || P.env_mask[e]) progress = 0.0f;

    // ---- write back state ------------------------------------------------------------
    if (valid && is_drone) {
        *DROW(D_POS) = p.x; *DROW(D_POS + 1) = p.y; *DROW(D_POS + 2) = p.z;
        *DROW(D_ROT) = q.w; *DROW(D_ROT + 1) = q.x; *DROW(D_ROT + 2) = q.y; *DROW(D_ROT + 3) = q.z;
        *DROW(D_LIN) = lv.x; *DROW(D_LIN + 1) = lv.y; *DROW(D_LIN + 2) = lv.z;
        *DROW(D_ANG) = av.x; *DROW(D_ANG + 1) = av.y; *DROW(D_ANG + 2) = av.z;
        if (!RESET) {
#pragma unroll
            for (int k = 0; k < 4; ++k) *DROW(D_THR + k) = thr[k];
            if (P.action_is_raw) {
                *DROW(D_INT) = integ.x; *DROW(D_INT + 1) = integ.y; *DROW(D_INT + 2) = integ.z;
                *DROW(D_LAST) = last.x; *DROW(D_LAST + 1) = last.y; *DROW(D_LAST + 2) = last.z;
            }
        }
    }
    if (valid && is_ev) {
        *EROW(E_TPOS) = tp.x; *EROW(E_TPOS + 1) = tp.y; *EROW(E_TPOS + 2) = tp.z;
        if (!RESET) { *EROW(E_TVEL) = tv.x; *EROW(E_TVEL + 1) = tv.y; *EROW(E_TVEL + 2) = tv.z; }
        *EROW(E_PROGRESS) = progress;
    }

    // ---- observation, hideandseek.py:746-917 ------------------------------------------
    const int row_l = (lane >> 2) * A + slot;            // row of this lane inside the warp tile
    const int64_t tile_row0 = e0 * A;                    // first [E*A] row of the warp
    V3 heading, up;
    heading_up(q, heading, up);

    // info.drone_state [E,A,13]
    {
        float* s = st.begin();
        if (is_drone) {
            float* r = s + row_l * 13;
            r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = q.w; r[4] = q.x; r[5] = q.y; r[6] = q.z;
            r[7] = lv.x; r[8] = lv.y; r[9] = lv.z; r[10] = av.x; r[11] = av.y; r[12] = av.z;
        }
        st.flush(P.b.drone_state + tile_row0 * 13, nenv * A * 13, full_tile);
    }
    // state_others [E,A,A-1,3] = p_a - p_j, j != a ascending; also drone-drone collisions
    float hit_drone = 0.f;
    if (A > 1) {
        float* s = st.begin();
        int o = 0;
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const V3 pj = gshfl3(p, gbase + j);
            if (is_drone && j != slot) {
                const V3 d = p - pj;
                float* r = s + row_l * ((A - 1) * 3) + o * 3;
                r[0] = d.x; r[1] = d.y; r[2] = d.z;
                hit_drone = hit_drone + ((norm3(d) < c.coll_radius_x2) ? 1.0f : 0.0f);
                ++o;
            }
        }
        st.flush(P.b.state_others + tile_row0 * ((A - 1) * 3), nenv * A * (A - 1) * 3, full_tile);
    }
    // k nearest cylinders [E,A,K,5]; lowest index wins ties
    float hit_cyl = 0.f;
    if (K > 0) {
        float* s = st.begin();
        if (is_drone) {
            float key[CT];
#pragma unroll
            for (int k = 0; k < CT; ++k)
                key[k] = (k < C) ? (norm3(mk(p.x - cx[k], p.y - cy[k], p.z - cz[k])) - c.cylinder_size) : INFINITY;
            unsigned taken = 0u;
            float* r = s + row_l * (K * 5);
#pragma unroll
            for (int n = 0; n < KMAX; ++n) {
                if (n < K) {
                    int best = 0; float bk = INFINITY; bool found = false;
#pragma unroll
                    for (int k = 0; k < CT; ++k) {
                        const bool cand = (k < C) && !((taken >> k) & 1u);
                        if (cand && (!found || key[k] < bk)) { best = k; bk = key[k]; found = true; }
                    }
                    taken |= 1u << best;
                    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
                    for (int k = 0; k < CT; ++k) if (k == best) { bx = cx[k]; by = cy[k]; bz = cz[k]; }
                    const bool inactive = bz < 0.0f;
                    const float rx = p.x - bx, ry = p.y - by, rz = p.z - bz;
                    const float mv = c.mask_value;
                    r[n * 5 + 0] = inactive ? mv : rx;
                    r[n * 5 + 1] = inactive ? mv : ry;
                    r[n * 5 + 2] = inactive ? mv : rz;
                    r[n * 5 + 3] = inactive ? mv : c.max_height;
                    r[n * 5 + 4] = inactive ? mv : c.cylinder_size;
                    const float dxy = fsqrt(rx * rx + ry * ry);
                    const float hit = ((dxy - c.cylinder_size) < c.collision_radius) ? 1.0f : 0.0f;
                    hit_cyl = hit_cyl + (inactive ? 0.0f : hit);
                }
            }
        }
        st.flush(P.b.obs_cylinders + tile_row0 * (K * 5), nenv * A * K * 5, full_tile);
    }
    // target visibility
    const V3 t_rpos = p - tp;
    bool blocked = false, detect = false;
    if (is_drone) {
        blocked = los_blocked(p, tp, cx, cy, cz, C, c.cylinder_size);
        detect = (norm3(t_rpos) < c.drone_detect_radius) && !blocked;
    }
    const unsigned gmask = ((1u << A) - 1u) << gbase;
    const unsigned det_ballot = __ballot_sync(FULL, detect);
    const bool bdetect = (det_ballot & gmask) != 0u;
    const float mv = c.mask_value;
    const float tfrac = fdiv(progress, (float)c.max_episode_length);

    if (c.use_tp_net) {
        // new TP frame [progress, tpos_masked3, tvel_masked3, p_0..p_{A-1}] = last row of the tile
        cp_async_wait_all();
        float* fr = tp_tile + (lane >> 2) * per_env + keep;
        if (is_drone) { fr[7 + 3 * slot] = p.x; fr[8 + 3 * slot] = p.y; fr[9 + 3 * slot] = p.z; }
        if (is_ev) {
            fr[0] = progress;
            fr[1] = bdetect ? tp.x : mv; fr[2] = bdetect ? tp.y : mv; fr[3] = bdetect ? tp.z : mv;
            fr[4] = bdetect ? tv.x : mv; fr[5] = bdetect ? tv.y : mv; fr[6] = bdetect ? tv.z : mv;
        }
        if (P.tp_init) {                                 // very first frame: every history row = this frame
            __syncwarp();
            float* row0 = tp_tile + (lane >> 2) * per_env;
            int k = slot;
            for (int i = slot; i < keep; i += G, k += G) {
                while (k >= FD) k -= FD;
                row0[i] = fr[k];
            }
        }
        {
            float* gdst = P.b.tp_input + e0 * per_env;
            const int nwords = nenv * per_env;
            const bool bulk = HS_USE_BULK_STORE && full_tile && ((nwords & 3) == 0) &&
                              ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0);
            if (bulk) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { bulk_store(gdst, tp_tile, (uint32_t)nwords * 4u); bulk_commit(); }
            } else {
                __syncwarp();
                for (int i = lane; i < nwords; i += 32) gdst[i] = tp_tile[i];
            }
        }
        if (valid && is_ev) {
            const float inv_ha = frcp(c.half_arena);
            P.b.tp_groundtruth[e * 3 + 0] = tp.x * inv_ha;
            P.b.tp_groundtruth[e * 3 + 1] = tp.y * inv_ha;
            P.b.tp_groundtruth[e * 3 + 2] = fdiv(tp.z, c.max_height) * 2.0f - 1.0f;
            P.b.tp_done[e] = (progress <= (float)(c.max_episode_length - c.future_step)) ? 1 : 0;
            *EROW(E_BDETECT) = bdetect ? 1.0f : 0.0f;
        }
    } else {
        // no predictor: the rows are complete now (width 20)
        const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
        float* s = st.begin();
        if (is_drone) write_self_row(s + row_l * 20, head_m, 0, nullptr, q, lv, heading, up, tfrac);
        st.flush(P.b.state_self + tile_row0 * 20, nenv * A * 20, full_tile);
        s = st.begin();
        if (is_drone) write_self_row(s + row_l * 20, t_rpos, 0, nullptr, q, lv, heading, up, tfrac);
        st.flush(P.b.state_drones + tile_row0 * 20, nenv * A * 20, full_tile);
    }

    if (RESET) {
        if (valid && is_ev && P.b.truncated != nullptr)
            P.b.truncated[e] = (progress > (float)c.max_episode_length) ? 1 : 0;
        st.finish();
        return;
    }

    // ---- reward / done / stats, hideandseek.py:919-1065 --------------------------------
    float r_dist = 0.f, r_speed = 0.f, r_coll = 0.f, r_smooth = 0.f, hit_wall = 0.f;
    bool seen_capture = false;
    if (is_drone) {
        const float dist = norm3(tp - p);
        r_dist = (-c.dist_reward_coef * dist) * ((dist > c.catch_radius) ? 1.0f : 0.0f);
        seen_capture = (dist < c.catch_radius) && !blocked;
        r_speed = -c.speed_coef * ((norm3(lv) > c.v_drone) ? 1.0f : 0.0f);
        r_coll = -c.collision_coef * hit_cyl;
        r_coll = r_coll + (-c.collision_coef * hit_drone);
        hit_wall = ((p.z > c.max_height) ? 1.0f : 0.0f) +
                   (((p.x * p.x + p.y * p.y) > c.arena_size_sq) ? 1.0f : 0.0f);
        r_coll = r_coll + (-c.collision_coef * hit_wall);
        r_smooth = c.smoothness_gated ? 0.0f : c.smoothness_coef * __expf(-action_err);
    }
    const bool any_capture = (__ballot_sync(FULL, seen_capture) & gmask) != 0u;
    const bool all_blocked = (__ballot_sync(FULL, blocked) & gmask) == gmask;
    const bool any_coll = (__ballot_sync(FULL, is_drone && (r_coll < 0.0f)) & gmask) != 0u;
    const float r_detect = c.detect_reward_coef * (bdetect ? 1.0f : 0.0f);
    const float r_catch = c.catch_reward_coef * (any_capture ? 1.0f : 0.0f);
    const float reward = ((((r_dist + r_detect) + r_catch) + r_coll) + r_speed) + r_smooth;
    if (valid && is_drone) P.b.reward[e * A + slot] = reward;

    // per-env means over the A pursuers (sum in agent order, then / A like torch.mean)
    // xor-butterfly over the 4 lanes of the group; non-pursuer lanes contribute the neutral
    // element, so for A=3 the sum is ((x0+x1)+(x2+0)) = the reference's left-to-right order
    const float inv_A = 1.0f / (float)A;
    auto gmean = [&](float x) {
        float s = is_drone ? x : 0.0f;
        s = s + __shfl_xor_sync(FULL, s, 1);
        s = s + __shfl_xor_sync(FULL, s, 2);
        return s * inv_A;
    };
    auto gmax = [&](float x) {
        float s = is_drone ? x : -INFINITY;
        s = fmaxf(s, __shfl_xor_sync(FULL, s, 1));
        s = fmaxf(s, __shfl_xor_sync(FULL, s, 2));
        return s;
    };
    const float m_ae = gmean(action_err), m_dist = gmean(r_dist), m_detect = gmean(r_detect),
                m_catch = gmean(r_catch), m_speed = gmean(r_speed), m_hcyl = gmean(hit_cyl),
                m_hdrone = gmean(hit_drone), m_hwall = gmean(hit_wall), m_coll = gmean(r_coll),
                m_smooth = gmean(r_smooth), m_tdiff = gmean(throttle_diff), m_reward = gmean(reward),
                x_tdiff = gmax(throttle_diff);

    if (valid && is_ev) {
        const bool done = progress >= (float)c.max_episode_length;
        P.b.done[e] = done ? 1 : 0;
        const float inv_len = done ? frcp(progress) : 1.0f;
        float* S = P.b.stats + e;
        const int64_t Es = E;
        cp_async_wait_all();
        const float* SO = stat_mem[wib][lane >> 2];     // values prefetched at kernel entry
#define ST(k) S[(int64_t)(k) * Es]
#define OLD(k) SO[k]
        // accumulators that are divided by the episode length on the done tick
        ST(HS_STAT_ACTION_ERROR_MEAN) = (OLD(HS_STAT_ACTION_ERROR_MEAN) + m_ae) * inv_len;
        ST(HS_STAT_ACTION_ERROR_MAX) = fmaxf(OLD(HS_STAT_ACTION_ERROR_MAX), m_ae);
        ST(HS_STAT_OUT_OF_ARENA) = ((OLD(HS_STAT_OUT_OF_ARENA) != 0.0f) || out_of_arena) ? 1.0f : 0.0f;
        ST(HS_STAT_DISTANCE_REWARD) = (OLD(HS_STAT_DISTANCE_REWARD) + m_dist) * inv_len;
        ST(HS_STAT_SUM_DETECT_STEP) = OLD(HS_STAT_SUM_DETECT_STEP) + 1.0f * (bdetect ? 1.0f : 0.0f);
        ST(HS_STAT_DETECT_REWARD) = (OLD(HS_STAT_DETECT_REWARD) + m_detect) * inv_len;
        ST(HS_STAT_BLOCKED) = OLD(HS_STAT_BLOCKED) + (all_blocked ? 1.0f : 0.0f);
        const bool capture_flag = r_catch != 0.0f;
        ST(HS_STAT_SUCCESS) = (capture_flag || (OLD(HS_STAT_SUCCESS) != 0.0f)) ? 1.0f : 0.0f;
        const float step_now = (capture_flag ? 1.0f : 0.0f) * progress +
                               (capture_flag ? 0.0f : 1.0f) * (float)c.max_episode_length;
        ST(HS_STAT_FIRST_CAPTURE_STEP) = fminf(OLD(HS_STAT_FIRST_CAPTURE_STEP), step_now);
        ST(HS_STAT_CATCH_REWARD) = (OLD(HS_STAT_CATCH_REWARD) + m_catch) * inv_len;
        ST(HS_STAT_SPEED_REWARD) = (OLD(HS_STAT_SPEED_REWARD) + m_speed) * inv_len;
        ST(HS_STAT_COLLISION_CYLINDER) = (OLD(HS_STAT_COLLISION_CYLINDER) + m_hcyl) * inv_len;
        ST(HS_STAT_COLLISION_DRONE) = (OLD(HS_STAT_COLLISION_DRONE) + m_hdrone) * inv_len;
        ST(HS_STAT_COLLISION) = (OLD(HS_STAT_COLLISION) + (any_coll ? 1.0f : 0.0f)) * inv_len;
        ST(HS_STAT_COLLISION_WALL) = (OLD(HS_STAT_COLLISION_WALL) + m_hwall) * inv_len;
        ST(HS_STAT_COLLISION_REWARD) = (OLD(HS_STAT_COLLISION_REWARD) + m_coll) * inv_len;
        if (c.write_smoothness_coef_stat) ST(HS_STAT_SMOOTHNESS_COEF) = c.smoothness_coef;
        ST(HS_STAT_SMOOTHNESS_REWARD) = (OLD(HS_STAT_SMOOTHNESS_REWARD) + m_smooth) * inv_len;
        ST(HS_STAT_SMOOTHNESS_MEAN) = (OLD(HS_STAT_SMOOTHNESS_MEAN) + m_tdiff) * inv_len;
        ST(HS_STAT_SMOOTHNESS_MAX) = fmaxf(x_tdiff, OLD(HS_STAT_SMOOTHNESS_MAX));
        ST(HS_STAT_RETURN) = OLD(HS_STAT_RETURN) + m_reward;
        // target_predicted_error is only ever divided (stays 0); distance_predicted_reward and
        // distance_threshold_L are never written (hideandseek.py:1023-1025).
#undef ST
#undef OLD
    }
    st.finish();
}

#undef DROW
#undef EROW
#define DROW(k) AROW((k) * A + slot)
#define EROW(k) AROW(ND * A + (k))

// =========================================================================================
// Second half with the trajectory predictor: state_self / state_drones rows (width 20+3F).
// hideandseek.py:834-887
// =========================================================================================
template <int A>
__global__ void __launch_bounds__(128)
hs_fill_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(128) float stage_mem[4][2][FILL_STAGE_WORDS];
    const hs_config& c = P.c;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slot = lane & (G - 1);
    const int64_t e0 = warp_g * ENVS_PER_WARP;
    const int E = c.num_envs;
    if (e0 >= E) return;
    const int64_t e_raw = e0 + (lane >> 2);
    const bool valid = e_raw < E;
    const int64_t e = valid ? e_raw : (E - 1);
    const bool is_drone = slot < A;
    const int nenv = (int)min((int64_t)ENVS_PER_WARP, E - e0);
    const bool full_tile = nenv == ENVS_PER_WARP;
    const int F = c.future_step, F3 = 3 * F, D = 20 + F3;

    Stager st;
    st.buf[0] = stage_mem[wib][0];
    st.buf[1] = stage_mem[wib][1];
    st.cur = 0;
    st.lane = lane;

    V3 p = mk(0, 0, 0), lv = mk(0, 0, 0);
    Q4 q; q.w = 1.f; q.x = q.y = q.z = 0.f;
    if (is_drone) {
        p = mk(*DROW(D_POS), *DROW(D_POS + 1), *DROW(D_POS + 2));
        q.w = *DROW(D_ROT); q.x = *DROW(D_ROT + 1); q.y = *DROW(D_ROT + 2); q.z = *DROW(D_ROT + 3);
        lv = mk(*DROW(D_LIN), *DROW(D_LIN + 1), *DROW(D_LIN + 2));
    }
    const V3 tp = mk(*EROW(E_TPOS), *EROW(E_TPOS + 1), *EROW(E_TPOS + 2));
    const float progress = *EROW(E_PROGRESS);
    const bool bdetect = *EROW(E_BDETECT) != 0.0f;
    float rp[3 * FMAX];
    const float* pr = P.tp_pred + e * F3;
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
        if (f < F) {
            const float px = (__ldg(pr + 3 * f) * 0.5f) * c.arena_size;
            const float py = (__ldg(pr + 3 * f + 1) * 0.5f) * c.arena_size;
            const float pz = ((__ldg(pr + 3 * f + 2) + 1.0f) / 2.0f) * c.max_height;
            rp[3 * f] = p.x - px; rp[3 * f + 1] = p.y - py; rp[3 * f + 2] = p.z - pz;
        } else { rp[3 * f] = rp[3 * f + 1] = rp[3 * f + 2] = 0.f; }
    }
    V3 heading, up;
    heading_up(q, heading, up);
    const float tfrac = fdiv(progress, (float)c.max_episode_length);
    const V3 t_rpos = p - tp;
    const float mv = c.mask_value;
    const V3 head_m = bdetect ? t_rpos : mk(mv, mv, mv);
    const int row_l = (lane >> 2) * A + slot;
    const int64_t tile_row0 = e0 * A;

    float* s = st.begin();
    if (is_drone) write_self_row(s + row_l * D, head_m, F3, rp, q, lv, heading, up, tfrac);
    st.flush(P.b.state_self + tile_row0 * D, nenv * A * D, full_tile);
    s = st.begin();
    if (is_drone) write_self_row(s + row_l * D, t_rpos, F3, rp, q, lv, heading, up, tfrac);
    st.flush(P.b.state_drones + tile_row0 * D, nenv * A * D, full_tile);
    st.finish();
}

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
        const float* xin = P.b.tp_input + e0 * (int64_t)(H * FD);
        float cst[TP_NE];
#pragma unroll
        for (int e = 0; e < TP_NE; ++e) cst[e] = 0.f;
        int cur = 0;
        // x_s tile [FD][16] (transposed) arrives by cp.async one time step ahead of its use
        auto fetch_x = [&](int s) {
            float* dst = xs + (s & 1) * FD * TPB_E;
            for (int i = tid; i < TPB_E * FD; i += TP_THREADS) {
                const int e = i / FD, k = i - e * FD;
                if (e < nenv) cp_async4(dst + k * TPB_E + e, xin + (int64_t)e * (H * FD) + s * FD + k);
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
        const float* xin = P.b.tp_input + e0 * (int64_t)(H * FD);
        float cst[TW_NE][2];
#pragma unroll
        for (int e = 0; e < TW_NE; ++e) { cst[e][0] = 0.f; cst[e][1] = 0.f; }
        int cur = 0;
        // x_s tile [FD][32] (transposed) arrives by cp.async one time step ahead of its use
        auto fetch_x = [&](int s) {
            float* dst = xs + (s & 1) * FD * TW_E;
            for (int i = tid; i < TW_E * FD; i += TW_THREADS) {
                const int e = i / FD, k = i - e * FD;
                if (e < nenv) cp_async4(dst + k * TW_E + e, xin + (int64_t)e * (H * FD) + s * FD + k);
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
        const float* xin = P.b.tp_input + e0 * (int64_t)(H * FD);
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
                if (r < nenv && k < FD) cp_async4(d, xin + (int64_t)r * (H * FD) + s * FD + k);
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
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
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
        const float* xin = P.b.tp_input + e * (int64_t)(H * FD);
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
                while (!mbar_try_wait(bar, phase)) { if (++spins > (1u << 24)) __trap(); }
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
template <int FD, int NTHREADS = TN_THREADS>
__device__ __forceinline__ void tn_stage_weights(const TPParams& W, float* wst, uint32_t lane_base, int row, int cg) {
    const int tid = threadIdx.x;
    auto lane_of = [](int wr) { const int g = wr >> 6, u = wr & 63; return (g >> 1) * 128 + 2 * u + (g & 1); };
#pragma unroll 8
    for (int i = tid; i < 256 * TP_HID; i += NTHREADS) {
        const int wr = i >> 6, k = i & 63;
        wst[lane_of(wr) * TN_WPITCH + 16 + k] = __ldg(W.w_hh + i);
    }
#pragma unroll 8
    for (int i = tid; i < 256 * 16; i += NTHREADS) {
        const int wr = i >> 4, k = i & 15;
        wst[lane_of(wr) * TN_WPITCH + k] = (k < FD) ? __ldg(W.w_ih + wr * FD + k) : 0.0f;
    }
    __syncthreads();
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
__device__ __forceinline__ void tn_stage_x(const float* __restrict__ tp_input, int64_t e0, int nenv, int H, uint8_t* Xhi, uint8_t* Xlo) {
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
            if (cm < H * 16 && n < nenv && k < FD) xv[u] = __ldg(tp_input + (e0 + n) * (int64_t)(H * FD) + s * FD + k);
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
template <int A, int NTHREADS = TN_THREADS>
__device__ __forceinline__ void tn_fc_rows(const KParams& P, const TPParams& W, int64_t e0, int nenv, const uint8_t* Hhi,
                                           const uint8_t* Hlo, const float* fcw, const float* fcb, float* preds, float* rowbuf) {
    const hs_config& c = P.c;
    const int F3 = 3 * c.future_step, D = 20 + F3, E = c.num_envs;
    const int tid = threadIdx.x;
    {
        const int n = tid & 31;
        for (int og = tid >> 5; og < F3; og += NTHREADS / 32) {
            float a0 = fcb[og];
            const float* w0 = fcw + og * TP_HID;
#pragma unroll 4
            for (int kc = 0; kc < TP_HID / 4; ++kc) {
                const float4 hh = *reinterpret_cast<const float4*>(Hhi + kc * TN_H_LBO + n * 16);
                const float4 hl = *reinterpret_cast<const float4*>(Hlo + kc * TN_H_LBO + n * 16);
                a0 = fmaf(w0[4 * kc], hh.x + hl.x, a0); a0 = fmaf(w0[4 * kc + 1], hh.y + hl.y, a0);
                a0 = fmaf(w0[4 * kc + 2], hh.z + hl.z, a0); a0 = fmaf(w0[4 * kc + 3], hh.w + hl.w, a0);
            }
            const float pv = tanhf(a0);
            preds[n * F3 + og] = pv;
            if (W.pred_out != nullptr && n < nenv) W.pred_out[(e0 + n) * F3 + og] = pv;
        }
    }
    __syncthreads();
    V3 t_rpos = mk(0.f, 0.f, 0.f);
    float* r1 = nullptr;
    if (tid < TN_E * A) {
        const int slot = tid / TN_E, el = tid - slot * TN_E;
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
    const bool bulk = HS_USE_BULK_STORE && (nenv == TN_E) && ((nwords & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(g1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g2) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(rowbuf) & 15) == 0);
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
            for (int i = tid; i < nwords; i += NTHREADS) gdst[i] = rowbuf[i];
        }
        __syncthreads();
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
    while (!mbar_try_wait(bar, phase)) { if (++spins > (1u << 24)) __trap(); }
    phase ^= 1;
}
// barrier `idx` of an array of mbarriers; `bits` holds one phase bit per barrier (no dynamically indexed registers)
__device__ __forceinline__ void mbar_wait_idx(uint32_t bar0, uint32_t idx, uint32_t& bits) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar0 + 8u * idx, (bits >> idx) & 1u)) { if (++spins > (1u << 24)) __trap(); }
    bits ^= 1u << idx;
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
        tn_stage_x<FD>(P.b.tp_input, e0, nenv, H, Xhi, Xlo);
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
        tn_fc_rows<A>(P, W, e0, nenv, Hhi, Hlo, fcw, fcb, preds, rowbuf);
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
constexpr int TCW_THREADS = TN_THREADS + 64;            // 16 epilogue warps + 2 issuing warps (one per M-tile)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

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
                tn_stage_x<FD, TCW_THREADS>(P.b.tp_input, e0, (int)min((int64_t)TN_E, E - e0), H, Xb + t * xslot,
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
                tn_fc_rows<A, TCW_THREADS>(P, W, e0, (int)min((int64_t)TN_E, E - e0), Hb + t * 2 * TN_H_BYTES,
                                          Hb + t * 2 * TN_H_BYTES + TN_H_BYTES, fcw, fcb, preds, rowbuf);
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

static size_t tp_smem_bytes(const hs_config& c) {
    const int FD = 7 + 3 * c.num_agents, KT = FD + TP_HID, F3 = 3 * c.future_step;
    size_t words = (size_t)KT * TP_WS + 256 + (size_t)F3 * TP_HID + 32 + 2 * (size_t)FD * TPB_E + 2 * TP_HID * TPB_E +
                   (size_t)TPB_E * 3 * FMAX;      // the row staging tile aliases the x/h region
    return words * sizeof(float);
}

// =========================================================================================
// Reset scatter: hideandseek.py:698-717, multirotor.py:635-650
// =========================================================================================
template <int A>
__global__ void __launch_bounds__(128)
hs_reset_scatter_kernel(const __grid_constant__ KParams P) {
    const hs_config& c = P.c;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = (int)(gt & (G - 1));
    const int64_t e = gt >> 2;
    const int E = c.num_envs;
    if (e >= E) return;
    const bool masked = (P.env_mask == nullptr) || (P.env_mask[e] != 0);
    const int C = c.num_cylinders;
    if (slot < A && masked) {
        const int64_t row = e * A + slot;
        for (int k = 0; k < 3; ++k) *DROW(D_POS + k) = P.init_drone_pos[row * 3 + k];
        for (int k = 0; k < 4; ++k) *DROW(D_ROT + k) = P.init_drone_rot[row * 4 + k];
        for (int k = 0; k < 3; ++k) { *DROW(D_LIN + k) = 0.f; *DROW(D_ANG + k) = 0.f; }
        const float h = c.hover_throttle;
        for (int k = 0; k < 4; ++k) *DROW(D_THR + k) = h;
        const float cmd_init = 2.0f * (h * h) - 1.0f;
        P.b.prev_action[row * 4 + 3] = 0.5f * (c.max_thrust_ratio + cmd_init);
    }
    if (slot == A) {
        if (masked) {
            for (int k = 0; k < 3; ++k) *EROW(E_TPOS + k) = P.init_target_pos[e * 3 + k];
            for (int k = 0; k < 3 * C; ++k) *EROW(E_CYL + k) = P.init_cyl_pos[e * 3 * C + k];
            for (int k = 0; k < HS_NUM_STATS; ++k) P.b.stats[(int64_t)k * E + e] = 0.f;
        }
        // every env, not just the masked ones (hideandseek.py:712)
        P.b.stats[(int64_t)HS_STAT_FIRST_CAPTURE_STEP * E + e] = (float)c.max_episode_length;
    }
}

// ---- AoS <-> arena field copies (views/* replacement, used by tests and tools) -----------
__global__ void hs_field_copy_kernel(float* arena, int64_t Ep, int row0, int n_slots, int width,
                                     int row_stride_slot, int row_stride_comp, int E, float* aos, int to_aos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_env = (int64_t)n_slots * width;
    if (i >= per_env * E) return;
    const int64_t e = i / per_env;
    const int r = (int)(i - e * per_env);
    const int a = r / width, k = r - a * width;
    float* ap = arena + ((int64_t)row0 + (int64_t)k * row_stride_comp + (int64_t)a * row_stride_slot) * Ep + e;
    if (to_aos) aos[i] = *ap; else *ap = aos[i];
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
struct hs_handle {
    hs_config cfg;
    hs_buffers bufs;
    bool bound;
    int device;
    int64_t Ep;
    int64_t launches;
    int tp_frames;               // number of TP frames written so far (0 -> next one initialises history)
    int block;                   // threads per block for the tick kernels
    int num_sms;
    int tp_variant;              // -1 auto, 0 fp32 FFMA, 1 3xTF32 mma.sync, 2/3 3xTF32 tcgen05 (128-/32-env tiles) (hs_set_option)
};

static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CUDA_OK(call)                                                        \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return set_err(HS_ERR_CUDA, #call ": %s", cudaGetErrorString(_e)); \
    } while (0)

template <bool RESET>
static cudaError_t launch_tick(const hs_handle* h, const KParams& P, cudaStream_t s) {
    const int64_t warps = ((int64_t)h->cfg.num_envs + ENVS_PER_WARP - 1) / ENVS_PER_WARP;
    const int wpb = h->block / 32;
    const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
    const bool small_c = h->cfg.num_cylinders <= 5;      // compile-time cylinder capacity 5 or 8
    switch (h->cfg.num_agents) {
        case 1: if (small_c) hs_tick_kernel<1, RESET, 5><<<grid, h->block, 0, s>>>(P); else hs_tick_kernel<1, RESET, CMAX><<<grid, h->block, 0, s>>>(P); break;
        case 2: if (small_c) hs_tick_kernel<2, RESET, 5><<<grid, h->block, 0, s>>>(P); else hs_tick_kernel<2, RESET, CMAX><<<grid, h->block, 0, s>>>(P); break;
        default: if (small_c) hs_tick_kernel<3, RESET, 5><<<grid, h->block, 0, s>>>(P); else hs_tick_kernel<3, RESET, CMAX><<<grid, h->block, 0, s>>>(P); break;
    }
    return cudaGetLastError();
}


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
    const uint4 w0 = philox4x32_10(make_uint4((uint32_t)t, 0xFFFF0000u, (uint32_t)epoch, (uint32_t)(epoch >> 32)), key);
    const int64_t idx = (int64_t)__umulhi(w0.x, (uint32_t)n_history);
    float origin[GEN_MAX_DIM], cand[GEN_MAX_DIM];
    for (int j = 0; j < dim; ++j) origin[j] = history[idx * dim + j];
    bool ok = false;
    for (int attempt = 0; attempt < 10 && !ok; ++attempt) {
        PhiloxStream rng;
        rng.ctr = make_uint4((uint32_t)t, (uint32_t)(64 * attempt), (uint32_t)epoch, (uint32_t)(epoch >> 32));
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

extern "C" {

int hs_abi_version(void) { return HS_ABI_VERSION; }
const char* hs_last_error(void) { return g_err; }

int hs_default_config(hs_config* c, int32_t num_envs) {
    if (!c) return set_err(HS_ERR_INVALID, "hs_default_config: null cfg%s");
    memset(c, 0, sizeof(*c));
    c->abi_version = HS_ABI_VERSION;
    c->num_envs = num_envs;
    c->num_agents = 3; c->num_cylinders = 5; c->obs_max_cylinder = 3;
    c->future_step = 5; c->history_step = 10; c->max_episode_length = 800;
    c->use_tp_net = 1; c->smoothness_gated = 1; c->write_smoothness_coef_stat = 1;
    c->fixed_yaw = 0; c->ground_clamp = 1;
    c->dt = 0.01f;
    c->arena_size = 0.9f; c->max_height = 1.2f; c->cylinder_size = 0.1f;
    c->catch_radius = 0.3f; c->collision_radius = 0.07f;
    c->drone_detect_radius = 100.0f; c->target_detect_radius = 100.0f;
    c->v_drone = 1.0f; c->mask_value = -5.0f;
    c->dist_reward_coef = 1.0f; c->catch_reward_coef = 20.0f; c->detect_reward_coef = 0.0f;
    c->collision_coef = 100.0f; c->speed_coef = 10.0f; c->smoothness_coef = 0.0f;
    c->target_clip = 1.0f; c->max_thrust_ratio = 0.9f;
    const float kp[3] = {250.f, 250.f, 120.f}, ki[3] = {500.f, 500.f, 16.7f}, kd[3] = {2.5f, 2.5f, 0.f},
                il[3] = {33.3f, 33.3f, 166.7f};
    for (int i = 0; i < 3; ++i) { c->pid_kp[i] = kp[i]; c->pid_ki[i] = ki[i]; c->pid_kd[i] = kd[i]; c->pid_ilimit[i] = il[i]; }
    c->pid_out_limit = 32767.0f;
    const float wmax = 2315.0f;
    c->kf = (wmax * wmax) * 2.350347298350041e-08f;
    c->km = (wmax * wmax) * 7.24e-10f;
    c->rotor_alpha = 0.01f / 0.025f;
    const float dirs[4] = {-1.f, 1.f, -1.f, 1.f};
    const float rx[4] = {0.028f, -0.028f, -0.028f, 0.028f}, ry[4] = {0.028f, 0.028f, -0.028f, -0.028f};
    for (int i = 0; i < 4; ++i) { c->rotor_dirs[i] = dirs[i]; c->rotor_x[i] = rx[i]; c->rotor_y[i] = ry[i]; }
    c->drag_coef_times_mass = 0.0f;
    c->downwash_kr = 2.0f; c->downwash_kz = 0.3f;
    const double m = 0.0321 + 4 * 1.0e-4;
    c->total_mass = (float)m;
    const double s = 4 * 1.0e-4 * 0.028 * 0.028;
    c->inertia[0] = (float)(1.4e-5 + s); c->inertia[1] = (float)(1.4e-5 + s); c->inertia[2] = (float)(2.17e-5 + 2 * s);
    c->gravity = 9.81f;
    c->lin_damp_factor = (float)(1.0 - 0.01 * 0.2); c->ang_damp_factor = (float)(1.0 - 0.01 * 0.2);
    c->max_linear_velocity = 1.0f; c->max_angular_velocity = 1000.0f;
    c->ground_z = 0.0125f;
    for (int i = 0; i < 3; ++i) c->inv_inertia[i] = 1.0f / c->inertia[i];
    c->hover_throttle = sqrtf((c->total_mass * 9.81f) / (4.0f * c->kf));
    c->arena_size_sq = (float)(0.9 * 0.9);
    c->half_arena = (float)(0.5 * 0.9);
    c->coll_radius_x2 = (float)(2.0 * 0.07);
    c->vmax_clamped = (float)(1.0 * (1.0 - 1e-6));
    return HS_OK;
}

static int check_cfg(const hs_config* c) {
    if (!c) return set_err(HS_ERR_INVALID, "null config%s");
    if (c->abi_version != HS_ABI_VERSION) return set_err(HS_ERR_INVALID, "config abi_version mismatch%s");
    if (c->num_envs <= 0) return set_err(HS_ERR_INVALID, "num_envs must be > 0%s");
    if (c->num_agents < 1 || c->num_agents > HS_MAX_AGENTS) return set_err(HS_ERR_INVALID, "num_agents must be 1..3%s");
    if (c->num_cylinders < 0 || c->num_cylinders > HS_MAX_CYLINDERS) return set_err(HS_ERR_INVALID, "num_cylinders must be 0..8%s");
    if (c->obs_max_cylinder < 0 || c->obs_max_cylinder > HS_MAX_OBS_CYLINDERS || c->obs_max_cylinder > c->num_cylinders)
        return set_err(HS_ERR_INVALID, "obs_max_cylinder must be <= min(num_cylinders, 4)%s");
    if (c->future_step < 0 || c->future_step > HS_MAX_FUTURE) return set_err(HS_ERR_INVALID, "future_step must be 0..8%s");
    if (c->history_step < 1) return set_err(HS_ERR_INVALID, "history_step must be >= 1%s");
    if (((int64_t)ND * c->num_agents + E_CYL + 3 * (int64_t)c->num_cylinders) * (((int64_t)c->num_envs + 31) & ~(int64_t)31) >= ((int64_t)1 << 31))
        return set_err(HS_ERR_INVALID, "num_envs too large: the state arena must stay below 2^31 words%s");
    if (c->use_tp_net && c->history_step * (7 + 3 * c->num_agents) > TP_ENV_WORDS_MAX)
        return set_err(HS_ERR_INVALID, "history_step * (7 + 3*num_agents) must be <= 192%s");
    return HS_OK;
}

int64_t hs_arena_floats(const hs_config* c) {
    if (check_cfg(c) != HS_OK) return -1;
    const int64_t Ep = ((int64_t)c->num_envs + 31) & ~(int64_t)31;
    return ((int64_t)ND * c->num_agents + E_CYL + 3 * (int64_t)c->num_cylinders) * Ep;
}

int hs_create(const hs_config* cfg, hs_handle** out) {
    if (!out) return set_err(HS_ERR_INVALID, "hs_create: null out%s");
    *out = nullptr;
    int rc = check_cfg(cfg);
    if (rc != HS_OK) return rc;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_err(HS_ERR_NO_DEVICE, "no CUDA device visible (%s); this library has no CPU path",
                       ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    }
    hs_handle* h = new (std::nothrow) hs_handle();
    if (!h) return set_err(HS_ERR_INVALID, "out of host memory%s");
    h->cfg = *cfg;
    memset(&h->bufs, 0, sizeof(h->bufs));
    h->bound = false;
    CUDA_OK(cudaGetDevice(&h->device));
    h->Ep = ((int64_t)cfg->num_envs + 31) & ~(int64_t)31;
    h->launches = 0;
    h->tp_frames = 0;
    h->tp_variant = -1;
    h->num_sms = 148;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    // small batches: smaller blocks spread the warps over more SMs (latency bound regime)
    const int64_t warps = ((int64_t)cfg->num_envs + ENVS_PER_WARP - 1) / ENVS_PER_WARP;
    h->block = (warps >= 4 * 148 * 4) ? 128 : (warps >= 2 * 148 * 2 ? 64 : 32);
    {
        // 5 CTAs x 43 KB of static shared memory per SM: ask for the largest shared carve-out
        cudaError_t e = cudaSuccess;
        const int co = cudaSharedmemCarveoutMaxShared;
        const bool small_c = cfg->num_cylinders <= 5;
#define HS_CARVE(AA) (small_c ? cudaFuncSetAttribute(hs_tick_kernel<AA, false, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, co) \
                              : cudaFuncSetAttribute(hs_tick_kernel<AA, false, CMAX>, cudaFuncAttributePreferredSharedMemoryCarveout, co))
        switch (cfg->num_agents) {
            case 1: e = HS_CARVE(1); break;
            case 2: e = HS_CARVE(2); break;
            default: e = HS_CARVE(3); break;
        }
#undef HS_CARVE
        if (e != cudaSuccess) { delete h; return set_err(HS_ERR_CUDA, "cudaFuncSetAttribute(carveout): %s", cudaGetErrorString(e)); }
    }
    if (cfg->use_tp_net) {
        // opt in to > 48 KB dynamic shared memory once (not a stream operation: keeps the
        // step entry points legal inside CUDA-graph capture)
        const int smem = (int)tp_smem_bytes(*cfg), smem_w = (int)tp_wide_smem_bytes(*cfg);
        cudaError_t e = cudaSuccess;
        const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        switch (cfg->num_agents) {
            case 1: e = cudaFuncSetAttribute(hs_tp_fill_kernel<1>, attr, smem);
                    if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tp_fill_wide_kernel<1>, attr, smem_w); break;
            case 2: e = cudaFuncSetAttribute(hs_tp_fill_kernel<2>, attr, smem);
                    if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tp_fill_wide_kernel<2>, attr, smem_w); break;
            default: e = cudaFuncSetAttribute(hs_tp_fill_kernel<3>, attr, smem);
                    if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tp_fill_wide_kernel<3>, attr, smem_w); break;
        }
        if (e == cudaSuccess) {
            const int m1 = (int)tp_mma_smem_bytes(*cfg, 1), m2 = (int)tp_mma_smem_bytes(*cfg, 2);
            switch (cfg->num_agents) {
                case 1: e = cudaFuncSetAttribute(hs_tp_fill_mma_kernel<1, 1>, attr, m1);
                        if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tp_fill_mma_kernel<1, 2>, attr, m2); break;
                case 2: e = cudaFuncSetAttribute(hs_tp_fill_mma_kernel<2, 1>, attr, m1);
                        if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tp_fill_mma_kernel<2, 2>, attr, m2); break;
                default: e = cudaFuncSetAttribute(hs_tp_fill_mma_kernel<3, 1>, attr, m1);
                        if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tp_fill_mma_kernel<3, 2>, attr, m2); break;
            }
        }
        if (e == cudaSuccess) {
            const int t = (int)tp_tc_smem_bytes(*cfg);
            switch (cfg->num_agents) {
                case 1: e = cudaFuncSetAttribute(hs_tp_fill_tc_kernel<1>, attr, t); break;
                case 2: e = cudaFuncSetAttribute(hs_tp_fill_tc_kernel<2>, attr, t); break;
                default: e = cudaFuncSetAttribute(hs_tp_fill_tc_kernel<3>, attr, t); break;
            }
        }
        if (e == cudaSuccess && tp_tcn_smem_bytes(*cfg) <= HS_MAX_DYN_SMEM) {
            const int t = (int)tp_tcn_smem_bytes(*cfg);
            switch (cfg->num_agents) {
                case 1: e = cudaFuncSetAttribute(hs_tp_fill_tcn_kernel<1>, attr, t); break;
                case 2: e = cudaFuncSetAttribute(hs_tp_fill_tcn_kernel<2>, attr, t); break;
                default: e = cudaFuncSetAttribute(hs_tp_fill_tcn_kernel<3>, attr, t); break;
            }
        }
        if (e == cudaSuccess && tp_tcw_smem_bytes(*cfg) <= HS_MAX_DYN_SMEM) {
            const int t = (int)tp_tcw_smem_bytes(*cfg);
            switch (cfg->num_agents) {
                case 1: e = cudaFuncSetAttribute(hs_tp_fill_tcw_kernel<1>, attr, t); break;
                case 2: e = cudaFuncSetAttribute(hs_tp_fill_tcw_kernel<2>, attr, t); break;
                default: e = cudaFuncSetAttribute(hs_tp_fill_tcw_kernel<3>, attr, t); break;
            }
        }
        if (e != cudaSuccess) { delete h; return set_err(HS_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem): %s", cudaGetErrorString(e)); }
    }
    *out = h;
    return HS_OK;
}

int hs_destroy(hs_handle* h) {
    delete h;
    return HS_OK;
}

int hs_bind_buffers(hs_handle* h, const hs_buffers* b) {
    if (!h || !b) return set_err(HS_ERR_INVALID, "hs_bind_buffers: null argument%s");
    const hs_config& c = h->cfg;
    if (!b->arena || !b->stats || (c.obs_max_cylinder > 0 && !b->obs_cylinders) || !b->state_self || !b->state_drones || !b->reward ||
        !b->done || !b->drone_state || !b->prev_action || !b->rotor_cmds || !b->ctbr || !b->target_rate ||
        !b->action_error || !b->v_prey)
        return set_err(HS_ERR_INVALID, "hs_bind_buffers: a required buffer is NULL%s");
    if (c.num_agents > 1 && !b->state_others) return set_err(HS_ERR_INVALID, "state_others is NULL%s");
    if (c.use_tp_net && (!b->tp_input || !b->tp_input_prev || !b->tp_groundtruth || !b->tp_done))
        return set_err(HS_ERR_INVALID, "use_tp_net needs tp_input/tp_input_prev/tp_groundtruth/tp_done%s");
    h->bufs = *b;
    h->bound = true;
    return HS_OK;
}

static KParams make_params(const hs_handle* h) {
    KParams P;
    memset(&P, 0, sizeof(P));
    P.c = h->cfg;
    P.b = h->bufs;
    P.Ep = h->Ep;
    return P;
}

int hs_step_pre(hs_handle* h, const float* action, int action_is_raw, const uint8_t* reset_pid, void* stream) {
    if (!h || !action) return set_err(HS_ERR_INVALID, "hs_step_pre: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_pre: call hs_bind_buffers first%s");
    KParams P = make_params(h);
    P.action = action;
    P.action_is_raw = action_is_raw;
    P.reset_pid = reset_pid;
    P.tp_init = (h->cfg.use_tp_net && h->tp_frames == 0) ? 1 : 0;
    CUDA_OK(launch_tick<false>(h, P, (cudaStream_t)stream));
    h->launches += 1;
    if (h->cfg.use_tp_net) h->tp_frames += 1;
    return HS_OK;
}

int hs_step_post(hs_handle* h, const float* tp_pred, void* stream) {
    if (!h || !tp_pred) return set_err(HS_ERR_INVALID, "hs_step_post: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_post: call hs_bind_buffers first%s");
    if (!h->cfg.use_tp_net) return set_err(HS_ERR_INVALID, "hs_step_post: config has use_tp_net == 0%s");
    KParams P = make_params(h);
    P.tp_pred = tp_pred;
    const int64_t warps = ((int64_t)h->cfg.num_envs + ENVS_PER_WARP - 1) / ENVS_PER_WARP;
    const int wpb = h->block / 32;
    const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
    cudaStream_t s = (cudaStream_t)stream;
    switch (h->cfg.num_agents) {
        case 1: hs_fill_kernel<1><<<grid, h->block, 0, s>>>(P); break;
        case 2: hs_fill_kernel<2><<<grid, h->block, 0, s>>>(P); break;
        default: hs_fill_kernel<3><<<grid, h->block, 0, s>>>(P); break;
    }
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    return HS_OK;
}

int hs_step_post_tp(hs_handle* h, const hs_tp_weights* w, float* tp_pred_out, void* stream) {
    if (!h || !w) return set_err(HS_ERR_INVALID, "hs_step_post_tp: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_post_tp: call hs_bind_buffers first%s");
    if (!h->cfg.use_tp_net) return set_err(HS_ERR_INVALID, "hs_step_post_tp: config has use_tp_net == 0%s");
    if (!w->weight_ih || !w->weight_hh || !w->bias_ih || !w->bias_hh || !w->fc_weight || !w->fc_bias)
        return set_err(HS_ERR_INVALID, "hs_step_post_tp: a weight pointer is NULL%s");
    if (w->hidden_size != TP_HID || w->input_size != 7 + 3 * h->cfg.num_agents || w->output_size != 3 * h->cfg.future_step)
        return set_err(HS_ERR_INVALID, "hs_step_post_tp: predictor shape must be LSTM(7+3A -> 64) + Linear(64 -> 3F)%s");
    KParams P = make_params(h);
    TPParams W;
    W.w_ih = w->weight_ih; W.w_hh = w->weight_hh; W.b_ih = w->bias_ih; W.b_hh = w->bias_hh;
    W.fc_w = w->fc_weight; W.fc_b = w->fc_bias; W.pred_out = tp_pred_out;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t wide_tiles = ((int64_t)h->cfg.num_envs + TW_E - 1) / TW_E;
    // auto (-1): the FFMA kernels win while a batch is a single wave of small tiles (one dependent
    // chain per tile, measured crossover ~6k envs); above that the tcgen05 kernel is 2-2.5x faster
    // auto (-1): 32-env tcgen05 tiles with the gates on M; one tile per CTA while the batch has at most one
    // tile per SM (variant 3), two tiles ping-ponging per CTA above that (variant 4).  The 128-env tcgen05
    // tile (2) and the FFMA kernels (0) remain as options and as the fallback when history_step makes the
    // 32-env tile's shared memory exceed 227 KB.
    const bool tcn_fits = tp_tcn_smem_bytes(h->cfg) <= HS_MAX_DYN_SMEM, tcw_fits = tp_tcw_smem_bytes(h->cfg) <= HS_MAX_DYN_SMEM;
    const int64_t tiles32 = ((int64_t)h->cfg.num_envs + TN_E - 1) / TN_E;
    int variant = (h->tp_variant >= 0) ? h->tp_variant : ((tiles32 > h->num_sms) ? 4 : 3);
    if ((variant == 3 && !tcn_fits) || (variant == 4 && !tcw_fits)) {
        if (h->tp_variant >= 3) return set_err(HS_ERR_INVALID, "predictor variant 3/4: history_step too large for the 32-env tile%s");
        variant = (variant == 4 && tcn_fits) ? 3 : ((h->cfg.num_envs >= 6144) ? 2 : 0);
    }
    if (variant == 4) {
        const size_t smem = tp_tcw_smem_bytes(h->cfg);
        const unsigned grid = (unsigned)min((tiles32 + 1) / 2, (int64_t)h->num_sms);   // persistent over tile pairs
        switch (h->cfg.num_agents) {
            case 1: hs_tp_fill_tcw_kernel<1><<<grid, TCW_THREADS, smem, s>>>(P, W); break;
            case 2: hs_tp_fill_tcw_kernel<2><<<grid, TCW_THREADS, smem, s>>>(P, W); break;
            default: hs_tp_fill_tcw_kernel<3><<<grid, TCW_THREADS, smem, s>>>(P, W); break;
        }
    } else
    if (variant == 3) {
        const size_t smem = tp_tcn_smem_bytes(h->cfg);
        const unsigned grid = (unsigned)min(((int64_t)h->cfg.num_envs + TN_E - 1) / TN_E, (int64_t)h->num_sms);   // persistent
        switch (h->cfg.num_agents) {
            case 1: hs_tp_fill_tcn_kernel<1><<<grid, TN_THREADS, smem, s>>>(P, W); break;
            case 2: hs_tp_fill_tcn_kernel<2><<<grid, TN_THREADS, smem, s>>>(P, W); break;
            default: hs_tp_fill_tcn_kernel<3><<<grid, TN_THREADS, smem, s>>>(P, W); break;
        }
    } else if (variant == 2) {
        // tcgen05 / TMEM variant: 128-env tiles, one CTA per tile
        const size_t smem = tp_tc_smem_bytes(h->cfg);
        const unsigned grid = (unsigned)min(((int64_t)h->cfg.num_envs + TC_M - 1) / TC_M, (int64_t)h->num_sms);   // persistent
        switch (h->cfg.num_agents) {
            case 1: hs_tp_fill_tc_kernel<1><<<grid, TC_THREADS, smem, s>>>(P, W); break;
            case 2: hs_tp_fill_tc_kernel<2><<<grid, TC_THREADS, smem, s>>>(P, W); break;
            default: hs_tp_fill_tc_kernel<3><<<grid, TC_THREADS, smem, s>>>(P, W); break;
        }
    } else if (variant == 1) {
        // tensor-core (3xTF32 mma.sync) variant: 32-env tiles when they fill the machine, else 16-env tiles
        const bool big = wide_tiles >= (int64_t)2 * h->num_sms;
        const int MT = big ? 2 : 1;
        const size_t smem = tp_mma_smem_bytes(h->cfg, MT);
        const int64_t ntiles = ((int64_t)h->cfg.num_envs + 16 * MT - 1) / (16 * MT);
        const unsigned grid = (unsigned)min(ntiles, (int64_t)2 * h->num_sms);
#define HS_MMA(AA) do { if (big) hs_tp_fill_mma_kernel<AA, 2><<<grid, TM_THREADS, smem, s>>>(P, W); \
                        else hs_tp_fill_mma_kernel<AA, 1><<<grid, TM_THREADS, smem, s>>>(P, W); } while (0)
        switch (h->cfg.num_agents) {
            case 1: HS_MMA(1); break;
            case 2: HS_MMA(2); break;
            default: HS_MMA(3); break;
        }
#undef HS_MMA
    } else if (wide_tiles >= (int64_t)2 * h->num_sms) {
        // enough 32-env tiles to give every SM two CTAs: the 8x8 register tile has the better FFMA:LDS ratio
        const size_t smem = tp_wide_smem_bytes(h->cfg);
        const unsigned grid = (unsigned)min(wide_tiles, (int64_t)2 * h->num_sms);
        switch (h->cfg.num_agents) {
            case 1: hs_tp_fill_wide_kernel<1><<<grid, TW_THREADS, smem, s>>>(P, W); break;
            case 2: hs_tp_fill_wide_kernel<2><<<grid, TW_THREADS, smem, s>>>(P, W); break;
            default: hs_tp_fill_wide_kernel<3><<<grid, TW_THREADS, smem, s>>>(P, W); break;
        }
    } else {
        // small batches: 16-env tiles double the number of CTAs so that SMs hold 8 warps
        const size_t smem = tp_smem_bytes(h->cfg);
        const int64_t ntiles = ((int64_t)h->cfg.num_envs + TPB_E - 1) / TPB_E;
        const unsigned grid = (unsigned)min(ntiles, (int64_t)2 * h->num_sms);
        switch (h->cfg.num_agents) {
            case 1: hs_tp_fill_kernel<1><<<grid, TP_THREADS, smem, s>>>(P, W); break;
            case 2: hs_tp_fill_kernel<2><<<grid, TP_THREADS, smem, s>>>(P, W); break;
            default: hs_tp_fill_kernel<3><<<grid, TP_THREADS, smem, s>>>(P, W); break;
        }
    }
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    return HS_OK;
}

int hs_reset(hs_handle* h, const uint8_t* env_mask, const float* drone_pos, const float* drone_rot,
             const float* target_pos, const float* cyl_pos, void* stream) {
    if (!h || !drone_pos || !drone_rot || !target_pos) return set_err(HS_ERR_INVALID, "hs_reset: null argument%s");
    if (h->cfg.num_cylinders > 0 && !cyl_pos) return set_err(HS_ERR_INVALID, "hs_reset: cyl_pos is NULL%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_reset: call hs_bind_buffers first%s");
    KParams P = make_params(h);
    P.env_mask = env_mask;
    P.init_drone_pos = drone_pos; P.init_drone_rot = drone_rot;
    P.init_target_pos = target_pos; P.init_cyl_pos = cyl_pos;
    P.tp_init = (h->cfg.use_tp_net && h->tp_frames == 0) ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t threads = (int64_t)h->cfg.num_envs * G;
    const unsigned grid = (unsigned)((threads + 127) / 128);
    switch (h->cfg.num_agents) {
        case 1: hs_reset_scatter_kernel<1><<<grid, 128, 0, s>>>(P); break;
        case 2: hs_reset_scatter_kernel<2><<<grid, 128, 0, s>>>(P); break;
        default: hs_reset_scatter_kernel<3><<<grid, 128, 0, s>>>(P); break;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(launch_tick<true>(h, P, s));
    h->launches += 2;
    if (h->cfg.use_tp_net) h->tp_frames += 1;
    return HS_OK;
}

int hs_sample_reset(hs_handle* h, const hs_reset_dist* dist, uint64_t epoch, float* drone_pos, float* drone_rot,
                    float* target_pos, float* cyl_pos, float* n_active, void* stream) {
    if (!h || !dist || !drone_pos || !drone_rot || !target_pos) return set_err(HS_ERR_INVALID, "hs_sample_reset: null argument%s");
    const int A = h->cfg.num_agents, C = h->cfg.num_cylinders, ng = dist->num_grid;
    if (C > 0 && !cyl_pos) return set_err(HS_ERR_INVALID, "hs_sample_reset: cyl_pos is NULL%s");
    if (ng < 1 || ng > RS_MAX_GRID || !(dist->grid_size > 0.f))
        return set_err(HS_ERR_INVALID, "hs_sample_reset: num_grid must be in [1, 11] and grid_size > 0%s");
    if (dist->fixed_num > C || (dist->fixed_num < 0 && (dist->min_cylinders < 0 || dist->min_cylinders > C)))
        return set_err(HS_ERR_INVALID, "hs_sample_reset: active-cylinder range outside [0, num_cylinders]%s");
    int inside = 0;
    const int half = ng / 2;
    for (int i = 0; i < ng; ++i)
        for (int j = 0; j < ng; ++j)
            inside += ((i - half) * (i - half) + (j - half) * (j - half) < half * half) ? 1 : 0;
    if (inside - (A + 1) < C)      // worst case: every body on its own free cell (reference: ValueError, :111-112)
        return set_err(HS_ERR_INVALID, "hs_sample_reset: not enough available grid cells for num_cylinders%s");
    const int E = h->cfg.num_envs;
    hs_reset_sample_kernel<<<(E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*dist, E, A, C, epoch, drone_pos, drone_rot,
                                                                             target_pos, cyl_pos, n_active);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    return HS_OK;
}

int hs_gen_sample_nearby(const hs_gen_params* p, const float* history, int64_t n_history, int64_t num_tasks, uint64_t epoch,
                         float* tasks_out, uint8_t* valid_out, void* stream) {
    if (!p || !history || !tasks_out || !valid_out) return set_err(HS_ERR_INVALID, "hs_gen_sample_nearby: null argument%s");
    if (p->num_agents < 1 || p->num_agents > 3 || p->num_cylinders < 0 || p->num_cylinders > CMAX)
        return set_err(HS_ERR_INVALID, "hs_gen_sample_nearby: unsupported task shape%s");
    if (p->num_grid < 1 || p->num_grid > RS_MAX_GRID || !(p->grid_size > 0.f))
        return set_err(HS_ERR_INVALID, "hs_gen_sample_nearby: num_grid must be in [1, 11] and grid_size > 0%s");
    if (n_history < 1 || n_history > 0xFFFFFFFFll) return set_err(HS_ERR_INVALID, "hs_gen_sample_nearby: empty archive%s");
    if (num_tasks <= 0) return HS_OK;
    // task bounds, hideandseek_envgen.py:327-340 (double arithmetic, rounded to fp32 once)
    GenBounds B;
    const double cb = (double)(int)((double)p->arena_size / (double)p->grid_size) * (double)p->grid_size;
    const double bxy = (double)p->arena_size / sqrt(2.0) - 0.1;
    const int A = p->num_agents, C = p->num_cylinders;
    int j = 0;
    for (int o = 0; o < A + 1; ++o) {
        B.lo[j] = (float)-bxy; B.hi[j++] = (float)bxy;
        B.lo[j] = (float)-bxy; B.hi[j++] = (float)bxy;
        B.lo[j] = (float)((double)p->max_height - 0.1); B.hi[j++] = (float)((double)p->max_height + 0.1);
    }
    for (int o = 0; o < C; ++o) {
        B.lo[j] = (float)-cb; B.hi[j++] = (float)cb;
        B.lo[j] = (float)-cb; B.hi[j++] = (float)cb;
        B.lo[j] = -20.0f; B.hi[j++] = (float)((double)p->max_height / 2.0);
    }
    const unsigned grid = (unsigned)((num_tasks + 127) / 128);
    hs_gen_sample_nearby_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*p, B, history, n_history, num_tasks, epoch, tasks_out, valid_out);
    CUDA_OK(cudaGetLastError());
    return HS_OK;
}

int64_t hs_fps_scratch_bytes(int64_t n) { return n * 4 + 8192; }

int hs_fps(const float* points, int64_t n, int32_t dim, int32_t k, int32_t start, int32_t* idx_out, void* scratch, void* stream) {
    if (!points || !idx_out || !scratch) return set_err(HS_ERR_INVALID, "hs_fps: null argument%s");
    if (n < 1 || n > 0x7FFFFFFFll || dim < 1 || dim > 64 || k < 1 || k > n || start < 0 || start >= n)
        return set_err(HS_ERR_INVALID, "hs_fps: need 1 <= k <= n, 1 <= dim <= 64, 0 <= start < n%s");
    int dev = 0, sms = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int G = (int)min((int64_t)sms, (n + FPS_THREADS - 1) / FPS_THREADS);
    if (G > 256) G = 256;                                    // slots: 2 x 256 x 8 B of the scratch tail
    int chunk = (int)((n + G - 1) / G);
    size_t smem = (size_t)chunk * dim * sizeof(float);
    int cache = 1;
    if (smem > 200 * 1024) { smem = 0; cache = 0; }
    CUDA_OK(cudaFuncSetAttribute(hs_fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaStream_t s = (cudaStream_t)stream;
    float* mind = reinterpret_cast<float*>(scratch);
    uint8_t* tail = reinterpret_cast<uint8_t*>(scratch) + ((n * 4 + 15) / 16) * 16;
    unsigned int* bar = reinterpret_cast<unsigned int*>(tail);
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(tail + 64);
    CUDA_OK(cudaMemsetAsync(tail, 0, 64, s));
    int ni = (int)n;
    void* args[] = {(void*)&points, (void*)&ni, (void*)&dim, (void*)&k, (void*)&start, (void*)&chunk, (void*)&cache,
                    (void*)&idx_out, (void*)&mind, (void*)&slots, (void*)&bar};
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)hs_fps_kernel, dim3(G), dim3(FPS_THREADS), args, smem, s));
    return HS_OK;
}

int hs_step_host(hs_handle* h, const float* action_host, int action_is_raw, float* reward_host,
                 uint8_t* done_host, float* staging_dev, void* stream) {
    if (!h || !action_host || !reward_host || !done_host || !staging_dev)
        return set_err(HS_ERR_INVALID, "hs_step_host: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_host: call hs_bind_buffers first%s");
    if (h->cfg.use_tp_net) return set_err(HS_ERR_INVALID, "hs_step_host: only for use_tp_net == 0%s");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.num_envs * h->cfg.num_agents;
    CUDA_OK(cudaMemcpyAsync(staging_dev, action_host, n * 4 * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = hs_step_pre(h, staging_dev, action_is_raw, nullptr, stream);
    if (rc != HS_OK) return rc;
    CUDA_OK(cudaMemcpyAsync(reward_host, h->bufs.reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(done_host, h->bufs.done, (size_t)h->cfg.num_envs, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    return HS_OK;
}

static int field_desc(const hs_handle* h, int field, int* row0, int* n_slots, int* width, int* stride_slot,
                      int* stride_comp) {
    const int A = h->cfg.num_agents, C = h->cfg.num_cylinders;
    const int ebase = ND * A;
    switch (field) {
        case HS_FIELD_DRONE_POS:     *row0 = D_POS * A;  *n_slots = A; *width = 3; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_DRONE_ROT:     *row0 = D_ROT * A;  *n_slots = A; *width = 4; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_DRONE_LINVEL:  *row0 = D_LIN * A;  *n_slots = A; *width = 3; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_DRONE_ANGVEL:  *row0 = D_ANG * A;  *n_slots = A; *width = 3; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_THROTTLE:      *row0 = D_THR * A;  *n_slots = A; *width = 4; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_PID_INTEG:     *row0 = D_INT * A;  *n_slots = A; *width = 3; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_PID_LAST_RATE: *row0 = D_LAST * A; *n_slots = A; *width = 3; *stride_slot = 1; *stride_comp = A; break;
        case HS_FIELD_TARGET_POS:    *row0 = ebase + E_TPOS; *n_slots = 1; *width = 3; *stride_slot = 0; *stride_comp = 1; break;
        case HS_FIELD_TARGET_VEL:    *row0 = ebase + E_TVEL; *n_slots = 1; *width = 3; *stride_slot = 0; *stride_comp = 1; break;
        case HS_FIELD_CYL_POS:       *row0 = ebase + E_CYL;  *n_slots = C; *width = 3; *stride_slot = 3; *stride_comp = 1; break;
        case HS_FIELD_PROGRESS:      *row0 = ebase + E_PROGRESS; *n_slots = 1; *width = 1; *stride_slot = 0; *stride_comp = 1; break;
        default: return set_err(HS_ERR_INVALID, "unknown field id%s");
    }
    return HS_OK;
}

static int field_copy(hs_handle* h, int field, float* aos, int to_aos, void* stream) {
    if (!h || !aos) return set_err(HS_ERR_INVALID, "hs_state_get/set: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_state_get/set: call hs_bind_buffers first%s");
    int row0, n_slots, width, ss, sc;
    int rc = field_desc(h, field, &row0, &n_slots, &width, &ss, &sc);
    if (rc != HS_OK) return rc;
    const int64_t n = (int64_t)h->cfg.num_envs * n_slots * width;
    if (n == 0) return HS_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    hs_field_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h->bufs.arena, h->Ep, row0, n_slots, width, ss, sc,
                                                                 h->cfg.num_envs, aos, to_aos);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    return HS_OK;
}

int hs_state_get(hs_handle* h, int field, float* dst, void* stream) { return field_copy(h, field, dst, 1, stream); }
int hs_state_set(hs_handle* h, int field, const float* src, void* stream) {
    return field_copy(h, field, const_cast<float*>(src), 0, stream);
}
int64_t hs_launch_count(const hs_handle* h) { return h ? h->launches : -1; }

int hs_set_option(hs_handle* h, int option, int value) {
    if (!h) return set_err(HS_ERR_INVALID, "hs_set_option: null handle%s");
    switch (option) {
        case HS_OPT_PREDICTOR_VARIANT:
            if (value < -1 || value > 4) return set_err(HS_ERR_INVALID, "predictor variant must be -1 (auto), 0 (fp32 FFMA), 1 (3xTF32 mma.sync), 2 (tcgen05, 128-env tiles), 3 (tcgen05, 32-env tiles) or 4 (tcgen05, 2 x 32-env tiles ping-pong)%s");
            h->tp_variant = value;
            return HS_OK;
        default:
            return set_err(HS_ERR_INVALID, "unknown option%s");
    }
}

}  // extern "C"
