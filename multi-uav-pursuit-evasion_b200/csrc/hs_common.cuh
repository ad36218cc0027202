// hs_common.cuh -- constants, kernel parameter block, vector/quaternion helpers, SFU wrappers, cp.async / TMA bulk-store helpers, arena row macros
// Part of the single translation unit hs_kernels.cu (unity build: everything lives in one anonymous
// namespace so that nvcc can inline across the pieces; -lineinfo still maps SASS to this file).
#pragma once

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
    int64_t Ep;                      // E rounded up to 32 (the arena holds Ep / 32 tiles)
    int32_t R;                       // arena rows per tile: 23 A + 8 + 3 C
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

// Chronological TP window of the tile that starts at env e0 (a multiple of 32): base pointer + distance between
// consecutive envs.  Plain mode: bufs.tp_input [E,H,FD].  Ring mode (bufs.tp_ring, include/hs_b200.h): the span that
// starts at the tile's advanced ring position; the position is the same for every tile of a batch.
__device__ __forceinline__ const float* tp_window_base(const KParams& P, int64_t e0, int HFD, int FD, int64_t& env_stride) {
    if (P.b.tp_ring == nullptr) { env_stride = HFD; return P.b.tp_input + e0 * (int64_t)HFD; }
    env_stride = 2 * (int64_t)HFD;
    return P.b.tp_ring + e0 * env_stride + (int64_t)P.b.tp_ring_pos[e0 >> 5] * FD;
}

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
// Division / square root / exp: the product build uses the SFU approximations (rcp/sqrt/ex2.approx, <= 2 ulp; the
// tick is instruction-bound, profiles/).  HS_EXACT_MATH=1 (csrc/hs_tick_exact.cu, compiled with -fmad=false; selected by
// HS_OPT_EXACT_MATH) swaps in IEEE round-to-nearest operations and follows the reference's operation order in the
// ill-conditioned stages, so that parity tests can show that every element outside the tolerance of the fast build
// sits on a discontinuity / cancellation and not on a defect.
#ifndef HS_EXACT_MATH
#define HS_EXACT_MATH 0
#endif
#if HS_EXACT_MATH
__device__ __forceinline__ float frcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float fsqrt(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
__device__ __forceinline__ float fexp(float x) { return expf(x); }
__device__ __forceinline__ float frsqrt(float x) { return __frcp_rn(__fsqrt_rn(x)); }
#else
__device__ __forceinline__ float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { const float r = frcp(s); return mk(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ float frsqrt(float x) { return rsqrtf(x); }
#endif
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

// Cold paths of the tick (ragged last tile, unaligned tensors) are kept OUT of the instruction stream of the hot
// path: the tick is bound by instruction delivery at small batches (profiles/), and these loops were ~15 % of its SASS.
__device__ __noinline__ void warp_copy_slow(float* gdst, const float* ssrc, int nwords, int lane) {
    for (int i = lane; i < nwords; i += 32) gdst[i] = ssrc[i];
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
            warp_copy_slow(gdst, s, nwords, lane);
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

constexpr int NARROW_MAX_AGENTS = 3;                 // the 4-lanes-per-env mapping: lanes 0..2 pursuers, lane A the evader
constexpr int FILL_STAGE_WORDS = ENVS_PER_WARP * NARROW_MAX_AGENTS * (20 + 3 * FMAX);   // 8*3*44 = 1056
constexpr int TICK_STAGE_WORDS = ENVS_PER_WARP * NARROW_MAX_AGENTS * 20;                // widest tick tile: [24][20]
constexpr int TP_ENV_WORDS_MAX = 192;                                               // history_step * (7+3A) <= 192

// ---- line of sight, hideandseek.py:47-103 ------------------------------------------------
template <int CT>
__device__ __forceinline__ bool los_blocked(const V3 p, const V3 t, const float (&cx)[CT],
                                            const float (&cy)[CT], const float (&cz)[CT],
                                            int C, float size) {
    const float ddx = p.x - t.x, ddy = p.y - t.y;
    const float dx = t.x - p.x, dy = t.y - p.y;
    const float den = (dx * dx + dy * dy) + 1e-5f;
    bool blocked = false;
#if HS_EXACT_MATH
    // the reference's own form: |cross| / (|seg| + 1e-5) <= size  and  0 <= num / (den + 1e-5) <= 1
    const float seg1 = __fsqrt_rn(ddx * ddx + ddy * ddy) + 1e-5f;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        if (c < C && cz[c] > 0.0f) {
            const float ccx = cx[c] - t.x, ccy = cy[c] - t.y;
            const float dl = __fdiv_rn(fabsf(ddx * ccy - ddy * ccx), seg1);
            const float tt = __fdiv_rn((cx[c] - p.x) * dx + (cy[c] - p.y) * dy, den);
            blocked = blocked || ((dl <= size) && (tt >= 0.0f) && (tt <= 1.0f));
        }
    }
#else
    // dist/(seg+eps) <= size  and  0 <= num/(den+eps) <= 1  with the (positive) denominators
    // multiplied out: no division per cylinder; underground (inactive) cylinders are skipped
    const float seg_sz = (fsqrt(ddx * ddx + ddy * ddy) + 1e-5f) * size;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        if (c < C && cz[c] > 0.0f) {
            const float ccx = cx[c] - t.x, ccy = cy[c] - t.y;
            const float cr = fabsf(ddx * ccy - ddy * ccx);
            const float num = (cx[c] - p.x) * dx + (cy[c] - p.y) * dy;
            blocked = blocked || ((cr <= seg_sz) && (num >= 0.0f) && (num <= den));
        }
    }
#endif
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

// State arena = tile-blocked SoA: tile t = envs [32 t, 32 t + 32) is ONE contiguous block of R rows x 32 floats, row r of
// env e at ((e >> 5) * R + r) * 32 + (e & 31).  A tile is what a CTA of the one-lane mapping moves with a single TMA box
// and what four warps of the 4-lane mapping share sector by sector; nothing of a tick is more than 12 KB away from the
// rest of its env (DRAM page locality; the plain [row][E] layout put the 92 rows of an env 4 MB apart at 1 Mi envs).
#define AROW(r) (P.b.arena + (((int64_t)(e) >> 5) * P.R + (r)) * 32 + ((e) & 31))
#define DROW(k) AROW((k) * A + slot)
#define EROW(k) AROW(ND * A + (k))


}  // namespace
