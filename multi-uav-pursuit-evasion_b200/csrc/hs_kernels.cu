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

#include "hs_common.cuh"
#include "hs_stages.cuh"
#include "hs_tick.cuh"
#include "hs_tick_wide.cuh"
#include "hs_predictor_ffma.cuh"
#include "hs_predictor_mma.cuh"
#include "hs_predictor_tcgen05.cuh"
#include "hs_rollout_fused.cuh"
#include "hs_rollout_pair.cuh"
#ifndef HS_ROLLOUT_AUTO_STREAMS
#define HS_ROLLOUT_AUTO_STREAMS 2     /* measured: 1 -> 17.96, 2 -> 15.35, 3 -> 15.35 us per tick at 4096 envs (the tick warps bound 2 and 3) */
#endif
#include "hs_reset.cuh"
#include "hs_hover.cuh"
#include "hs_samplers.cuh"
#include "hs_rollout.cuh"
#include "hs_policy.cuh"
#include "hs_policy_tc.cuh"

// the tick kernel built with IEEE arithmetic (csrc/hs_tick_exact.cu, HS_OPT_EXACT_MATH)
cudaError_t hs_launch_tick_exact(const void* kparams, size_t bytes, int num_agents, int reset, int small_c, unsigned grid,
                                 unsigned block, cudaStream_t s);
cudaError_t hs_launch_tick_wide_exact(const void* kparams, size_t bytes, const void* maps3, int num_agents, int reset, int small_c,
                                      unsigned grid, size_t smem, cudaStream_t s);
cudaError_t hs_wide_exact_set_smem(int num_agents, int small_c, int bytes);

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
    // hs_step_host_io: side stream that carries the tick's own outputs to the host while the predictor runs
    cudaStream_t io_stream = nullptr;
    cudaEvent_t io_tick_done = nullptr, io_copy_done = nullptr;
    // ... and the whole host-to-host tick as ONE graph launch (memcpy + kernel nodes), cached per pointer set
    struct IoKey {
        hs_host_io io; hs_buffers bufs; hs_tp_weights w;
        const void* staging; const void* reset_pid;
        int raw, tp_init, variant, has_w;
    };
    struct IoGraph { IoKey key; cudaGraphExec_t exec; uint64_t last_use; };
    static constexpr int IO_GRAPHS = 8;
    IoGraph io_graphs[IO_GRAPHS] = {};
    cudaStream_t io_capture = nullptr;
    uint64_t io_clock = 0;
    int fused_tick = 1;          // HS_OPT_FUSED_TICK: hs_step_fused may use the one-launch kernel
    int io_fused = 0;            // hs_step_host_io: two kernels + overlapped copy (0) or the one-launch kernel (1); see DESIGN.md
    int io_graph_mode = 1;       // HS_OPT_HOST_IO_GRAPH: 1 = graph launch (default), 0 = stream API calls
    int io_zero_copy_action = 1; // HS_OPT_HOST_IO_ZERO_COPY_ACTION: pinned host actions are read in place by the tick kernel
    int exact_math = 0;          // HS_OPT_EXACT_MATH: the tick runs the IEEE-arithmetic build of hs_tick_kernel (parity evidence)
    bool rollout_ready = false;  // hs_rollout_fused_kernel's shared-memory attribute set
    int rollout_variant = 0;     // HS_OPT_ROLLOUT_VARIANT: 0 auto, else ticks per predictor pass (1, 2, 3)
    int tick_mapping = 0;        // HS_OPT_TICK_MAPPING: 0 auto, 1 four lanes per env, 2 one lane per env (hs_tick_wide_kernel)
    // TMA tensor maps of the one-lane mapping (state tile load / store, stats tile), valid for tm_arena / tm_stats
    CUtensorMap tm[3];
    const void* tm_arena = nullptr;
    const void* tm_stats = nullptr;
    bool wide_ready = false;     // shared-memory opt-in done for this handle's kernels
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

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*hs_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static hs_encode_tiled_fn tensor_map_encoder() {
    static hs_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<hs_encode_tiled_fn>(p);
    }
    return fn;
}
// 2-D fp32 tensor [rows][cols] with row pitch `pitch_floats`, box = [box_rows][32 columns]
static bool encode_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_floats, uint32_t box_rows) {
    hs_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch_floats * sizeof(float)};
    const cuuint32_t box[2] = {32u, box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one lane per env? (HS_OPT_TICK_MAPPING; the wide kernel needs a 16-byte stats row pitch and a TP window that is not shifted in place)
static bool wide_possible(const hs_handle* h) {
    const hs_config& c = h->cfg;
    if ((c.num_envs & 3) != 0 || c.num_agents < 3) return false;
    if (c.use_tp_net && !h->bufs.tp_ring && h->bufs.tp_input == h->bufs.tp_input_prev) return false;
    return true;
}
static bool tp_ring_mode(const hs_handle* h) { return h->cfg.use_tp_net && h->bufs.tp_ring != nullptr; }
static bool use_wide(const hs_handle* h) {
    if (h->cfg.num_agents > NARROW_MAX_AGENTS || (h->cfg.use_obstacles && h->cfg.use_tp_net) || tp_ring_mode(h)) return true;
    if (h->tick_mapping == 1) return false;
    if (h->tick_mapping == 2) return wide_possible(h);
    return h->cfg.num_envs >= 32768 && wide_possible(h);
}

template <int A, int CT, bool RESET>
static cudaError_t launch_wide_one(const KParams& P, const CUtensorMap* tm, unsigned grid, size_t smem, cudaStream_t s) {
    hs_tick_wide_kernel<A, CT, RESET><<<grid, (A + 1) * 32, smem, s>>>(P, tm[0], tm[1]);
    return cudaGetLastError();
}
template <int A, int CT>
static cudaError_t wide_set_smem(int bytes) {
    cudaError_t e = cudaFuncSetAttribute(hs_tick_wide_kernel<A, CT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hs_tick_wide_kernel<A, CT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}
#define HS_WIDE_DISPATCH(CALL)                                                                        \
    switch (A) {                                                                                      \
        case 3: return small_c ? CALL(3, 5) : CALL(3, CMAX);                                          \
        case 4: return small_c ? CALL(4, 5) : CALL(4, CMAX);                                          \
        case 5: return small_c ? CALL(5, 5) : CALL(5, CMAX);                                          \
        case 6: return small_c ? CALL(6, 5) : CALL(6, CMAX);                                          \
        default: return cudaErrorInvalidValue;                                                        \
    }
static cudaError_t wide_set_smem_dispatch(int A, bool small_c, int bytes) {
#define HS_CALL(AA, CC) wide_set_smem<AA, CC>(bytes)
    HS_WIDE_DISPATCH(HS_CALL)
#undef HS_CALL
}
template <bool RESET>
static cudaError_t launch_wide_dispatch(int A, bool small_c, const KParams& P, const CUtensorMap* tm, unsigned grid, size_t smem, cudaStream_t s) {
#define HS_CALL(AA, CC) launch_wide_one<AA, CC, RESET>(P, tm, grid, smem, s)
    HS_WIDE_DISPATCH(HS_CALL)
#undef HS_CALL
}

template <bool RESET>
static cudaError_t launch_tick(hs_handle* h, const KParams& P, cudaStream_t s) {
    const bool small_c = h->cfg.num_cylinders <= 5;      // compile-time cylinder capacity 5 or 8
    if (use_wide(h)) {
        const hs_config& c = h->cfg;
        if (!wide_possible(h)) return cudaErrorInvalidConfiguration;
        const WidePlan w = wide_plan(c.num_agents, c.num_cylinders, c.obs_max_cylinder, c.use_tp_net != 0);
        const size_t smem = (size_t)WIDE_WARPS * w.total * sizeof(float);
        if (!h->wide_ready) {
            cudaError_t e = wide_set_smem_dispatch(c.num_agents, small_c, (int)smem);
            if (e == cudaSuccess) e = hs_wide_exact_set_smem(c.num_agents, small_c ? 1 : 0, (int)smem);
            if (e != cudaSuccess) return e;
            h->wide_ready = true;
        }
        if (h->tm_arena != h->bufs.arena) {
            // the tile-blocked arena as a 2-D tensor [tiles x R rows][32 floats]: a tile is the box of its R consecutive rows
            const uint64_t rows = (uint64_t)(h->Ep / 32) * (uint64_t)w.rows_all;
            if (!encode_map(&h->tm[0], h->bufs.arena, 32, rows, 32, (uint32_t)w.rows_all) ||
                !encode_map(&h->tm[1], h->bufs.arena, 32, rows, 32, (uint32_t)w.rows_rw))
                return cudaErrorInvalidValue;
            h->tm_arena = h->bufs.arena;
            h->tm_stats = h->bufs.stats;
        }
        const int64_t tiles = ((int64_t)c.num_envs + 31) / 32;
        const unsigned grid = (unsigned)((tiles + WIDE_WARPS - 1) / WIDE_WARPS);
        if (h->exact_math)
            return hs_launch_tick_wide_exact(&P, sizeof(P), h->tm, c.num_agents, RESET ? 1 : 0, small_c ? 1 : 0, grid, smem, s);
        return launch_wide_dispatch<RESET>(c.num_agents, small_c, P, h->tm, grid, smem, s);
    }
    const int64_t warps = ((int64_t)h->cfg.num_envs + ENVS_PER_WARP - 1) / ENVS_PER_WARP;
    const int wpb = h->block / 32;
    const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
    if (h->exact_math)
        return hs_launch_tick_exact(&P, sizeof(P), h->cfg.num_agents, RESET ? 1 : 0, small_c ? 1 : 0, grid, (unsigned)h->block, s);
    switch (h->cfg.num_agents) {
        case 1: if (small_c) hs_tick_kernel<1, RESET, 5><<<grid, h->block, 0, s>>>(P); else hs_tick_kernel<1, RESET, CMAX><<<grid, h->block, 0, s>>>(P); break;
        case 2: if (small_c) hs_tick_kernel<2, RESET, 5><<<grid, h->block, 0, s>>>(P); else hs_tick_kernel<2, RESET, CMAX><<<grid, h->block, 0, s>>>(P); break;
        default: if (small_c) hs_tick_kernel<3, RESET, 5><<<grid, h->block, 0, s>>>(P); else hs_tick_kernel<3, RESET, CMAX><<<grid, h->block, 0, s>>>(P); break;
    }
    return cudaGetLastError();
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
    c->drone_radius = 0.06f; c->evader_radius = 0.05f;
    return HS_OK;
}

static int check_cfg(const hs_config* c) {
    if (!c) return set_err(HS_ERR_INVALID, "null config%s");
    if (c->abi_version != HS_ABI_VERSION) return set_err(HS_ERR_INVALID, "config abi_version mismatch%s");
    if (c->num_envs <= 0) return set_err(HS_ERR_INVALID, "num_envs must be > 0%s");
    if (c->num_agents < 1 || c->num_agents > HS_MAX_AGENTS) return set_err(HS_ERR_INVALID, "num_agents must be 1..6%s");
    if (c->num_agents > NARROW_MAX_AGENTS && (c->num_envs & 3) != 0)
        return set_err(HS_ERR_INVALID, "num_agents > 3 runs on the one-lane-per-env kernel, which needs num_envs % 4 == 0%s");
    if (c->num_cylinders < 0 || c->num_cylinders > HS_MAX_CYLINDERS) return set_err(HS_ERR_INVALID, "num_cylinders must be 0..8%s");
    if (c->obs_max_cylinder < 0 || c->obs_max_cylinder > HS_MAX_OBS_CYLINDERS || c->obs_max_cylinder > c->num_cylinders)
        return set_err(HS_ERR_INVALID, "obs_max_cylinder must be <= min(num_cylinders, 4)%s");
    if (c->future_step < 0 || c->future_step > HS_MAX_FUTURE) return set_err(HS_ERR_INVALID, "future_step must be 0..8%s");
    if (c->history_step < 1) return set_err(HS_ERR_INVALID, "history_step must be >= 1%s");
    if (((int64_t)ND * c->num_agents + E_CYL + 3 * (int64_t)c->num_cylinders) * (((int64_t)c->num_envs + 31) & ~(int64_t)31) >= ((int64_t)1 << 31))
        return set_err(HS_ERR_INVALID, "num_envs too large: the state arena must stay below 2^31 words%s");
    if (c->use_obstacles && c->use_tp_net && (c->num_agents < 3 || (c->num_envs & 3) != 0))
        return set_err(HS_ERR_INVALID, "use_obstacles runs on the one-lane-per-env tick: needs num_agents >= 3 and num_envs % 4 == 0%s");
    if (c->use_tp_net && !c->use_obstacles && c->num_agents <= NARROW_MAX_AGENTS && c->history_step * (7 + 3 * c->num_agents) > TP_ENV_WORDS_MAX)
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
    if (cfg->use_tp_net && cfg->num_agents <= NARROW_MAX_AGENTS) {
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
        if (e == cudaSuccess && tp_fused_smem_bytes(*cfg) <= HS_MAX_DYN_SMEM) {
            const int t = (int)tp_fused_smem_bytes(*cfg);
            const bool small_c = cfg->num_cylinders <= 5;
#define HS_FATTR(AA) (small_c ? cudaFuncSetAttribute(hs_tick_tp_fused_kernel<AA, 5>, attr, t) : cudaFuncSetAttribute(hs_tick_tp_fused_kernel<AA, 8>, attr, t))
            switch (cfg->num_agents) {
                case 1: e = HS_FATTR(1); break;
                case 2: e = HS_FATTR(2); break;
                default: e = HS_FATTR(3); break;
            }
#undef HS_FATTR
        }
        if (e == cudaSuccess && tp_half_smem_bytes(*cfg) <= HS_MAX_DYN_SMEM) {
            const int t = (int)tp_half_smem_bytes(*cfg);
            switch (cfg->num_agents) {
                case 1: e = cudaFuncSetAttribute(hs_tick_tp_fused_kernel<1, 5, false>, attr, t); break;
                case 2: e = cudaFuncSetAttribute(hs_tick_tp_fused_kernel<2, 5, false>, attr, t); break;
                default: e = cudaFuncSetAttribute(hs_tick_tp_fused_kernel<3, 5, false>, attr, t); break;
            }
        }
        if (e != cudaSuccess) { delete h; return set_err(HS_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem): %s", cudaGetErrorString(e)); }
    }
    *out = h;
    return HS_OK;
}

int hs_destroy(hs_handle* h) {
    if (h) {
        if (h->io_tick_done) cudaEventDestroy(h->io_tick_done);
        if (h->io_copy_done) cudaEventDestroy(h->io_copy_done);
        if (h->io_stream) cudaStreamDestroy(h->io_stream);
        for (auto& g : h->io_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        if (h->io_capture) cudaStreamDestroy(h->io_capture);
    }
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
    if ((b->tp_ring != nullptr) != (b->tp_ring_pos != nullptr))
        return set_err(HS_ERR_INVALID, "tp_ring and tp_ring_pos come together%s");
    if (c.use_tp_net && b->tp_ring && (c.num_agents < 3 || (c.num_envs & 3) != 0))
        return set_err(HS_ERR_INVALID, "tp_ring needs the lane-per-env tick mapping: num_agents >= 3 and num_envs a multiple of 4%s");
    if (c.use_tp_net && (!(b->tp_ring || (b->tp_input && b->tp_input_prev)) || !b->tp_groundtruth || !b->tp_done))
        return set_err(HS_ERR_INVALID, "use_tp_net needs tp_input/tp_input_prev (or tp_ring/tp_ring_pos) and tp_groundtruth/tp_done%s");
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
    P.R = ND * h->cfg.num_agents + E_CYL + 3 * h->cfg.num_cylinders;
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

// One control tick including the predictor.  Small batches (at most one 32-env tile per SM, auto predictor policy):
// ONE launch of hs_tick_tp_fused_kernel; otherwise hs_step_pre followed by hs_step_post_tp.  Same results either way.
int hs_step_fused(hs_handle* h, const float* action, int action_is_raw, const uint8_t* reset_pid, const hs_tp_weights* w,
                  float* tp_pred_out, void* stream) {
    if (!h || !action || !w) return set_err(HS_ERR_INVALID, "hs_step_fused: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_fused: call hs_bind_buffers first%s");
    if (!h->cfg.use_tp_net) return set_err(HS_ERR_INVALID, "hs_step_fused: config has use_tp_net == 0 (use hs_step_pre)%s");
    if (h->cfg.num_agents > NARROW_MAX_AGENTS || h->cfg.use_obstacles)
        return set_err(HS_ERR_INVALID, "hs_step_fused: the fused predictor kernels cover num_agents <= 3 without use_obstacles; use hs_step_pre + the module + hs_step_post%s");
    const int64_t tiles32 = ((int64_t)h->cfg.num_envs + TN_E - 1) / TN_E;
    const bool one_launch = h->fused_tick && !h->exact_math && h->tick_mapping != 2 && !tp_ring_mode(h) && (h->tp_variant < 0 || h->tp_variant == 5) &&
                            tiles32 <= h->num_sms && tp_fused_smem_bytes(h->cfg) <= HS_MAX_DYN_SMEM && h->cfg.num_agents <= 3;
    if (!one_launch) {
        const int rc = hs_step_pre(h, action, action_is_raw, reset_pid, stream);
        return rc != HS_OK ? rc : hs_step_post_tp(h, w, tp_pred_out, stream);
    }
    if (!w->weight_ih || !w->weight_hh || !w->bias_ih || !w->bias_hh || !w->fc_weight || !w->fc_bias)
        return set_err(HS_ERR_INVALID, "hs_step_fused: a weight pointer is NULL%s");
    if (w->hidden_size != TP_HID || w->input_size != 7 + 3 * h->cfg.num_agents || w->output_size != 3 * h->cfg.future_step)
        return set_err(HS_ERR_INVALID, "hs_step_fused: predictor shape must be LSTM(7+3A -> 64) + Linear(64 -> 3F)%s");
    KParams P = make_params(h);
    P.action = action;
    P.action_is_raw = action_is_raw;
    P.reset_pid = reset_pid;
    P.tp_init = (h->tp_frames == 0) ? 1 : 0;
    TPParams W;
    W.w_ih = w->weight_ih; W.w_hh = w->weight_hh; W.b_ih = w->bias_ih; W.b_hh = w->bias_hh;
    W.fc_w = w->fc_weight; W.fc_b = w->fc_bias; W.pred_out = tp_pred_out;
    const size_t smem = tp_fused_smem_bytes(h->cfg);
    const unsigned grid = (unsigned)tiles32;
    cudaStream_t s = (cudaStream_t)stream;
    const bool small_c = h->cfg.num_cylinders <= 5;
#define HS_FUSED(AA) do { if (small_c) hs_tick_tp_fused_kernel<AA, 5><<<grid, TCW_THREADS, smem, s>>>(P, W); \
                          else hs_tick_tp_fused_kernel<AA, 8><<<grid, TCW_THREADS, smem, s>>>(P, W); } while (0)
    switch (h->cfg.num_agents) {
        case 1: HS_FUSED(1); break;
        case 2: HS_FUSED(2); break;
        default: HS_FUSED(3); break;
    }
#undef HS_FUSED
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    h->tp_frames += 1;
    return HS_OK;
}

int hs_gather_rows(const hs_gather_tensor* tensors, int num_tensors, const int64_t* indices_device, int64_t num_rows,
                   int num_steps, void* stream) {
    if (!tensors || !indices_device) return set_err(HS_ERR_INVALID, "hs_gather_rows: null argument%s");
    if (num_tensors < 1 || num_tensors > HS_GATHER_MAX_TENSORS || num_rows < 0 || num_steps < 1)
        return set_err(HS_ERR_INVALID, "hs_gather_rows: 1 <= num_tensors <= HS_GATHER_MAX_TENSORS, num_rows >= 0, num_steps >= 1%s");
    static_assert(GATHER_MAX_TENSORS == HS_GATHER_MAX_TENSORS, "header and kernel disagree");
    if (num_rows == 0) return HS_OK;
    GatherParams G;
    memset(&G, 0, sizeof(G));
    for (int k = 0; k < num_tensors; ++k) {
        const hs_gather_tensor& t = tensors[k];
        if (!t.src || !t.dst || t.row_bytes < 1) return set_err(HS_ERR_INVALID, "hs_gather_rows: a tensor has a null pointer or an empty row%s");
        G.d[k].src = static_cast<const uint8_t*>(t.src); G.d[k].dst = static_cast<uint8_t*>(t.dst);
        G.d[k].stride_env = t.stride_env; G.d[k].stride_step = t.stride_step; G.d[k].row_bytes = t.row_bytes;
    }
    G.indices = indices_device; G.num_rows = num_rows; G.T = num_steps; G.num_tensors = num_tensors;
    const int64_t blocks = (num_rows + 7) / 8;                       // 8 warps per block, one row per warp and pass
    dim3 grid((unsigned)(blocks < 4096 ? blocks : 4096), (unsigned)num_tensors);
    hs_gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(G);
    CUDA_OK(cudaGetLastError());
    return HS_OK;
}

// T control ticks (tick + predictor each) in ONE launch: hs_rollout_fused_kernel keeps a 32-env tile per CTA for the whole
// rollout (csrc/hs_rollout_fused.cuh).  Same results as T calls of hs_step_fused.
int hs_rollout_fused(hs_handle* h, const hs_buffers* sets_device, int num_sets, int first_set, const float* first_tp_prev,
                     const float* action, int64_t action_tick_stride, int action_is_raw, int num_ticks, const hs_tp_weights* w,
                     float* tp_pred_out, int64_t pred_tick_stride, void* stream) {
    if (!h || !sets_device || !first_tp_prev || !action || !w) return set_err(HS_ERR_INVALID, "hs_rollout_fused: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_rollout_fused: call hs_bind_buffers first%s");
    const hs_config& c = h->cfg;
    if (num_sets < 1 || first_set < 0 || first_set >= num_sets || num_ticks < 1)
        return set_err(HS_ERR_INVALID, "hs_rollout_fused: num_sets >= 1, 0 <= first_set < num_sets, num_ticks >= 1%s");
    const int64_t tiles32 = ((int64_t)c.num_envs + TN_E - 1) / TN_E;
    static_assert(RP_H == 10, "hs_rollout_pair_kernel is built for history_step 10");
    if (!c.use_tp_net || c.num_agents != 3 || c.history_step != 10 || c.use_obstacles || h->exact_math || tp_ring_mode(h) ||
        rollout_fused_smem_bytes(c) > HS_MAX_DYN_SMEM)
        return set_err(HS_ERR_INVALID, "hs_rollout_fused covers the reference's shape (3 pursuers, history_step 10, use_tp_net, no use_obstacles, "
                                       "fast-math build, plain TP window); use hs_step_fused per tick otherwise%s");
    if (h->tp_frames == 0) return set_err(HS_ERR_INVALID, "hs_rollout_fused: run hs_reset first (the first frame fills the TP window)%s");
    if (!w->weight_ih || !w->weight_hh || !w->bias_ih || !w->bias_hh || !w->fc_weight || !w->fc_bias)
        return set_err(HS_ERR_INVALID, "hs_rollout_fused: a weight pointer is NULL%s");
    if (w->hidden_size != TP_HID || w->input_size != 7 + 3 * c.num_agents || w->output_size != 3 * c.future_step)
        return set_err(HS_ERR_INVALID, "hs_rollout_fused: predictor shape must be LSTM(7+3A -> 64) + Linear(64 -> 3F)%s");
    KParams P = make_params(h);
    P.action = action;
    P.action_is_raw = action_is_raw;
    TPParams W;
    W.w_ih = w->weight_ih; W.w_hh = w->weight_hh; W.b_ih = w->bias_ih; W.b_hh = w->bias_hh;
    W.fc_w = w->fc_weight; W.fc_b = w->fc_bias; W.pred_out = nullptr;
    RolloutParams RP;
    RP.sets = sets_device; RP.num_sets = num_sets; RP.first_set = first_set; RP.num_ticks = num_ticks;
    RP.first_tp_prev = first_tp_prev;
    RP.action = action; RP.action_tick_stride = action_tick_stride;
    RP.pred_out = tp_pred_out; RP.pred_tick_stride = pred_tick_stride;
    // streams per predictor pass: 1 = hs_rollout_fused_kernel; 2, 3 = hs_rollout_pair_kernel<NS>
    int ns = (h->rollout_variant == 0) ? HS_ROLLOUT_AUTO_STREAMS : h->rollout_variant;
    while (ns > 1 && rollout_pair_smem_bytes(c, ns) > HS_MAX_DYN_SMEM) --ns;
    const size_t smem = ns > 1 ? rollout_pair_smem_bytes(c, ns) : rollout_fused_smem_bytes(c);
    if (!h->rollout_ready) {
        const cudaFuncAttribute attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        CUDA_OK(cudaFuncSetAttribute(hs_rollout_fused_kernel<3, 5>, attr, (int)rollout_fused_smem_bytes(c)));
        CUDA_OK(cudaFuncSetAttribute(hs_rollout_fused_kernel<3, CMAX>, attr, (int)rollout_fused_smem_bytes(c)));
        if (rollout_pair_smem_bytes(c, 2) <= HS_MAX_DYN_SMEM) {
            CUDA_OK(cudaFuncSetAttribute(hs_rollout_pair_kernel<3, 5, 2>, attr, (int)rollout_pair_smem_bytes(c, 2)));
            CUDA_OK(cudaFuncSetAttribute(hs_rollout_pair_kernel<3, CMAX, 2>, attr, (int)rollout_pair_smem_bytes(c, 2)));
        }
        if (rollout_pair_smem_bytes(c, 3) <= HS_MAX_DYN_SMEM) {
            CUDA_OK(cudaFuncSetAttribute(hs_rollout_pair_kernel<3, 5, 3>, attr, (int)rollout_pair_smem_bytes(c, 3)));
            CUDA_OK(cudaFuncSetAttribute(hs_rollout_pair_kernel<3, CMAX, 3>, attr, (int)rollout_pair_smem_bytes(c, 3)));
        }
        h->rollout_ready = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const bool small_c = c.num_cylinders <= 5;
    const unsigned grid = (unsigned)tiles32;
    if (ns == 3) {
        if (small_c) hs_rollout_pair_kernel<3, 5, 3><<<grid, RF_THREADS, smem, s>>>(P, W, RP);
        else hs_rollout_pair_kernel<3, CMAX, 3><<<grid, RF_THREADS, smem, s>>>(P, W, RP);
    } else if (ns == 2) {
        if (small_c) hs_rollout_pair_kernel<3, 5, 2><<<grid, RF_THREADS, smem, s>>>(P, W, RP);
        else hs_rollout_pair_kernel<3, CMAX, 2><<<grid, RF_THREADS, smem, s>>>(P, W, RP);
    } else {
        if (small_c) hs_rollout_fused_kernel<3, 5><<<grid, RF_THREADS, smem, s>>>(P, W, RP);
        else hs_rollout_fused_kernel<3, CMAX><<<grid, RF_THREADS, smem, s>>>(P, W, RP);
    }
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    h->tp_frames += num_ticks;
    return HS_OK;
}

#ifdef HS_FUSED_TIMING
int hs_debug_times(unsigned long long* out32) {          // debug builds only (tools/fused_phases.py)
    return cudaMemcpyFromSymbol(out32, hs_dbg_times, 32 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
#endif

int hs_step_post(hs_handle* h, const float* tp_pred, void* stream) {
    if (!h || !tp_pred) return set_err(HS_ERR_INVALID, "hs_step_post: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_post: call hs_bind_buffers first%s");
    if (!h->cfg.use_tp_net) return set_err(HS_ERR_INVALID, "hs_step_post: config has use_tp_net == 0%s");
    KParams P = make_params(h);
    P.tp_pred = tp_pred;
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cfg.num_agents > NARROW_MAX_AGENTS) {
        const unsigned grid = (unsigned)(((int64_t)h->cfg.num_envs + 127) / 128);
        switch (h->cfg.num_agents) {
            case 4: hs_fill_wide_kernel<4><<<grid, 128, 0, s>>>(P); break;
            case 5: hs_fill_wide_kernel<5><<<grid, 128, 0, s>>>(P); break;
            default: hs_fill_wide_kernel<6><<<grid, 128, 0, s>>>(P); break;
        }
    } else {
        const int64_t warps = ((int64_t)h->cfg.num_envs + ENVS_PER_WARP - 1) / ENVS_PER_WARP;
        const int wpb = h->block / 32;
        const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
        switch (h->cfg.num_agents) {
            case 1: hs_fill_kernel<1><<<grid, h->block, 0, s>>>(P); break;
            case 2: hs_fill_kernel<2><<<grid, h->block, 0, s>>>(P); break;
            default: hs_fill_kernel<3><<<grid, h->block, 0, s>>>(P); break;
        }
    }
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    return HS_OK;
}

int hs_step_post_tp(hs_handle* h, const hs_tp_weights* w, float* tp_pred_out, void* stream) {
    if (!h || !w) return set_err(HS_ERR_INVALID, "hs_step_post_tp: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_post_tp: call hs_bind_buffers first%s");
    if (!h->cfg.use_tp_net) return set_err(HS_ERR_INVALID, "hs_step_post_tp: config has use_tp_net == 0%s");
    if (h->cfg.num_agents > NARROW_MAX_AGENTS || h->cfg.use_obstacles)
        return set_err(HS_ERR_INVALID, "hs_step_post_tp: the fused predictor kernels cover num_agents <= 3 without use_obstacles; use the module + hs_step_post%s");
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
    const bool half_fits = tp_half_smem_bytes(h->cfg) <= HS_MAX_DYN_SMEM;
    int variant = (h->tp_variant >= 0) ? h->tp_variant : ((tiles32 > h->num_sms) ? 4 : 5);
    if (variant == 5 && !half_fits && h->tp_variant < 0) variant = 3;
    if ((variant == 3 && !tcn_fits) || (variant == 4 && !tcw_fits) || (variant == 5 && !half_fits)) {
        if (h->tp_variant >= 3) return set_err(HS_ERR_INVALID, "predictor variant 3/4/5: history_step too large for the 32-env tile%s");
        variant = (variant == 4 && tcn_fits) ? 3 : ((h->cfg.num_envs >= 6144) ? 2 : 0);
    }
    if (variant == 5) {
        // one 32-env tile per CTA, its two 16-env halves ping-pong (the predictor half of hs_tick_tp_fused_kernel)
        const size_t smem = tp_half_smem_bytes(h->cfg);
        const unsigned grid = (unsigned)tiles32;
        switch (h->cfg.num_agents) {
            case 1: hs_tick_tp_fused_kernel<1, 5, false><<<grid, TCW_THREADS, smem, s>>>(P, W); break;
            case 2: hs_tick_tp_fused_kernel<2, 5, false><<<grid, TCW_THREADS, smem, s>>>(P, W); break;
            default: hs_tick_tp_fused_kernel<3, 5, false><<<grid, TCW_THREADS, smem, s>>>(P, W); break;
        }
    } else
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
    const int64_t threads = (int64_t)h->cfg.num_envs * (h->cfg.num_agents + 1);
    const unsigned grid = (unsigned)((threads + 127) / 128);
    switch (h->cfg.num_agents) {
        case 1: hs_reset_scatter_kernel<1><<<grid, 128, 0, s>>>(P); break;
        case 2: hs_reset_scatter_kernel<2><<<grid, 128, 0, s>>>(P); break;
        case 3: hs_reset_scatter_kernel<3><<<grid, 128, 0, s>>>(P); break;
        case 4: hs_reset_scatter_kernel<4><<<grid, 128, 0, s>>>(P); break;
        case 5: hs_reset_scatter_kernel<5><<<grid, 128, 0, s>>>(P); break;
        default: hs_reset_scatter_kernel<6><<<grid, 128, 0, s>>>(P); break;
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

int64_t hs_policy_blob_floats(int32_t self_dim) {
    if (self_dim < 1 || self_dim > PL_E) return 0;
    // fp32 part rounded up to 1 KB so that the tensor-core image that follows is bulk-copy aligned
    return (((int64_t)policy_blob_layout(self_dim).total + 255) & ~(int64_t)255) + policy_tc_layout(self_dim).total / 4;
}

int hs_policy_prepare(const hs_policy_weights* w, float* blob, void* stream) {
    if (!w || !blob) return set_err(HS_ERR_INVALID, "hs_policy_prepare: null argument%s");
    if (w->self_dim < 1 || w->self_dim > PL_E || w->head_dim < 1 || w->head_dim > PL_HEAD_MAX)
        return set_err(HS_ERR_INVALID, "hs_policy_prepare: need 1 <= self_dim <= 128 and 1 <= head_dim <= 8%s");
    if (!w->embed_self_w || !w->embed_self_b || !w->embed_ln_w || !w->embed_ln_b || !w->attn_in_w || !w->attn_in_b ||
        !w->attn_out_w || !w->attn_out_b || !w->lin1_w || !w->lin1_b || !w->lin2_w || !w->lin2_b || !w->norm1_w ||
        !w->norm1_b || !w->norm2_w || !w->norm2_b || !w->head_w || !w->head_b)
        return set_err(HS_ERR_INVALID, "hs_policy_prepare: a required parameter pointer is NULL%s");
    hs_policy_prepare_kernel<<<PL_E, PL_E, 0, (cudaStream_t)stream>>>(*w, blob);
    CUDA_OK(cudaGetLastError());
    uint8_t* img = reinterpret_cast<uint8_t*>(blob + (((int64_t)policy_blob_layout(w->self_dim).total + 255) & ~(int64_t)255));
    hs_policy_prepare_tc_kernel<<<dim3(PL_E, 5), PL_E, 0, (cudaStream_t)stream>>>(blob, img, w->self_dim);
    CUDA_OK(cudaGetLastError());
    return HS_OK;
}

int hs_policy_forward(const float* blob, int32_t self_dim, int32_t head_dim, const hs_policy_io* io, void* stream) {
    if (!blob || !io || !io->state_self || !io->head_out) return set_err(HS_ERR_INVALID, "hs_policy_forward: null argument%s");
    const int64_t num_rows = io->num_rows;
    if (self_dim < 1 || self_dim > PL_E || head_dim < 1 || head_dim > PL_HEAD_MAX || io->n_others < 0 || io->n_others > 2 ||
        io->n_cyl < 0 || io->n_cyl > 4 || num_rows < 1)
        return set_err(HS_ERR_INVALID, "hs_policy_forward: need self_dim <= 128, head_dim <= 8, n_others <= 2, n_cyl <= 4%s");
    if ((io->n_others > 0 && !io->state_others) || (io->n_cyl > 0 && !io->cylinders))
        return set_err(HS_ERR_INVALID, "hs_policy_forward: state_others / cylinders missing%s");
    PolicyArgs A;
    A.blob = blob; A.state_self = io->state_self; A.state_others = io->state_others; A.cylinders = io->cylinders;
    A.eps = io->eps; A.rng = io->eps ? nullptr : io->rng_state;
    A.head_out = io->head_out; A.action = io->action; A.logp = io->logp; A.eps_out = io->eps_out; A.feat_out = io->feat_out;
    A.R = num_rows; A.D = self_dim; A.n_others = io->n_others; A.n_cyl = io->n_cyl; A.head_dim = head_dim;
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    static int sm_count[64] = {};
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (sms == 0) {
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_count[dev] = sms;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int impl = io->impl ? io->impl : (num_rows >= 1024 ? 2 : 1);
    if (impl == 2) {
        // tcgen05 kernel: 128-row tiles, persistent CTAs (one per SM)
        const uint8_t* img = reinterpret_cast<const uint8_t*>(blob + (((int64_t)policy_blob_layout(self_dim).total + 255) & ~(int64_t)255));
        const size_t smem = policy_tc_smem_bytes();
        static bool attr_set[64] = {};                       // per device, once (the call costs ~10 us of host time)
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            CUDA_OK(cudaFuncSetAttribute(hs_policy_forward_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        const int64_t ntiles = (num_rows + PT_M - 1) / PT_M;
        hs_policy_forward_tc_kernel<0><<<(unsigned)min(ntiles, (int64_t)sms), PT_THREADS, smem, s>>>(A, img);
        CUDA_OK(cudaGetLastError());
        return HS_OK;
    }
    // 64-row tiles halve the weight traffic per row; 32-row tiles fill the machine at small batches
    const bool big = num_rows >= (int64_t)sms * 64 * 2;
    if (big) {
        const size_t smem = policy_smem_bytes<64>(self_dim);
        CUDA_OK(cudaFuncSetAttribute(hs_policy_forward_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hs_policy_forward_kernel<64><<<(unsigned)((num_rows + 63) / 64), 256, smem, s>>>(A);
    } else {
        const size_t smem = policy_smem_bytes<32>(self_dim);
        CUDA_OK(cudaFuncSetAttribute(hs_policy_forward_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hs_policy_forward_kernel<32><<<(unsigned)((num_rows + 31) / 32), 128, smem, s>>>(A);
    }
    CUDA_OK(cudaGetLastError());
    return HS_OK;
}

int hs_gae(const hs_gae_params* p, const float* reward, const uint8_t* done, const float* value, const float* next_value,
           float* advantages, float* returns, void* scratch, float* stats_out, void* stream) {
    if (!p || !reward || !done || !value || !next_value || !advantages || !returns)
        return set_err(HS_ERR_INVALID, "hs_gae: null argument%s");
    if (p->num_envs < 1 || p->num_steps < 1 || p->num_agents < 1)
        return set_err(HS_ERR_INVALID, "hs_gae: num_envs, num_steps and num_agents must be >= 1%s");
    if ((p->normalize || stats_out) && !scratch)
        return set_err(HS_ERR_INVALID, "hs_gae: normalize / stats_out need the 16-byte scratch buffer%s");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ncol = p->num_envs * p->num_agents;
    double* moments = reinterpret_cast<double*>(scratch);
    if (moments) CUDA_OK(cudaMemsetAsync(moments, 0, 2 * sizeof(double), s));
    const float gamma = (float)p->gamma, gl = (float)(p->gamma * p->lmbda);
    const unsigned grid = (unsigned)((ncol + 255) / 256);
    hs_gae_kernel<<<grid, 256, 0, s>>>(reward, done, value, next_value, advantages, returns, moments, ncol, p->num_steps,
                                       p->num_agents, p->stride_env, p->stride_step, p->done_stride_env,
                                       p->done_stride_step, gamma, gl);
    CUDA_OK(cudaGetLastError());
    if (p->normalize || stats_out) {
        const int64_t total = ncol * p->num_steps;
        int dev = 0, sms = 0;
        CUDA_OK(cudaGetDevice(&dev));
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        // without normalize the launch only publishes {mean, std}
        const unsigned g2 = p->normalize ? (unsigned)min((int64_t)sms * 8, (total + 255) / 256) : 1u;
        hs_adv_normalize_kernel<<<g2, 256, 0, s>>>(advantages, moments, stats_out, ncol, p->num_steps, p->num_agents,
                                                   p->stride_env, p->stride_step, p->normalize ? 1 : 0);
        CUDA_OK(cudaGetLastError());
    }
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

// Enqueues H2D action -> tick -> {D2H of the tick's own outputs on the side stream || predictor -> D2H state_self} on `s`
// (no synchronisation).  Runs either directly or under stream capture.
static int host_io_enqueue(hs_handle* h, const hs_host_io* io, int action_is_raw, const uint8_t* reset_pid, const hs_tp_weights* w,
                           float* staging_dev, void* stream) {
    const hs_config& c = h->cfg;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t EA = (size_t)c.num_envs * c.num_agents;
    // Pinned (page-locked, UVA-mapped) action buffers are read by the tick kernel straight over PCIe - 48 B per
    // pursuer, prefetched at kernel entry - which removes the H2D copy node and its dependency from the critical
    // path; pageable memory goes through the staging buffer.
    const float* action_dev = staging_dev;
    {
        cudaPointerAttributes pa;
        const cudaError_t e = cudaPointerGetAttributes(&pa, io->action);
        if (e == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr && h->io_zero_copy_action)
            action_dev = static_cast<const float*>(pa.devicePointer);
        else
            (void)cudaGetLastError();
    }
    if (action_dev == staging_dev)
        CUDA_OK(cudaMemcpyAsync(staging_dev, io->action, EA * 4 * sizeof(float), cudaMemcpyHostToDevice, s));
    // one-launch tick + predictor when the batch qualifies: nothing is complete before that kernel ends, so the
    // whole observation follows it in one copy (the side-branch overlap below is for the two-kernel sequence)
    const int64_t tiles32 = ((int64_t)c.num_envs + TN_E - 1) / TN_E;
    const bool one_launch = c.use_tp_net && h->fused_tick && !h->exact_math && (h->tp_variant < 0 || h->tp_variant == 5) &&
                            tiles32 <= h->num_sms && tp_fused_smem_bytes(c) <= HS_MAX_DYN_SMEM && h->io_fused;
    int rc = one_launch ? hs_step_fused(h, action_dev, action_is_raw, reset_pid, w, nullptr, stream)
                        : hs_step_pre(h, action_dev, action_is_raw, reset_pid, stream);
    if (rc != HS_OK) return rc;
    const size_t D = 20 + (c.use_tp_net ? 3 * (size_t)c.future_step : 0);
    // seg[0] is written by the predictor kernel (use_tp_net) -- the rest is complete after the tick
    struct Seg { const char* dev; char* host; size_t bytes; } seg[5] = {
        {(const char*)h->bufs.state_self, (char*)io->state_self, EA * D * sizeof(float)},
        {(const char*)h->bufs.state_others, (char*)io->state_others, EA * (size_t)(c.num_agents - 1) * 3 * sizeof(float)},
        {(const char*)h->bufs.obs_cylinders, (char*)io->obs_cylinders, EA * (size_t)c.obs_max_cylinder * 5 * sizeof(float)},
        {(const char*)h->bufs.reward, (char*)io->reward, EA * sizeof(float)},
        {(const char*)h->bufs.done, (char*)io->done, (size_t)c.num_envs}};
    auto copy_segments = [&](int first, int last, cudaStream_t cs) -> cudaError_t {
        int i = first;
        while (i < last) {
            if (!seg[i].host || !seg[i].dev || seg[i].bytes == 0) { ++i; continue; }
            const char* d0 = seg[i].dev;
            char* h0 = seg[i].host;
            size_t len = seg[i].bytes;
            int j = i + 1;
            // merge neighbours: same spacing on both sides, gap (alignment padding) of at most 4 KB
            while (j < last && seg[j].host && seg[j].dev && seg[j].bytes > 0 && seg[j].dev >= d0 + len &&
                   (size_t)(seg[j].dev - (d0 + len)) <= 4096 && (seg[j].dev - d0) == (seg[j].host - h0)) {
                len = (size_t)(seg[j].dev - d0) + seg[j].bytes;
                ++j;
            }
            const cudaError_t e = cudaMemcpyAsync(h0, d0, len, cudaMemcpyDeviceToHost, cs);
            if (e != cudaSuccess) return e;
            i = j;
        }
        return cudaSuccess;
    };
    if (one_launch) {
        CUDA_OK(copy_segments(0, 5, s));
    } else if (c.use_tp_net) {
        // the rows the tick itself completed travel on a side stream while the predictor kernel runs
        if (!h->io_stream) {
            CUDA_OK(cudaStreamCreateWithFlags(&h->io_stream, cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&h->io_tick_done, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&h->io_copy_done, cudaEventDisableTiming));
        }
        CUDA_OK(cudaEventRecord(h->io_tick_done, s));
        CUDA_OK(cudaStreamWaitEvent(h->io_stream, h->io_tick_done, 0));
        CUDA_OK(copy_segments(1, 5, h->io_stream));
        CUDA_OK(cudaEventRecord(h->io_copy_done, h->io_stream));
        rc = hs_step_post_tp(h, w, nullptr, stream);
        if (rc != HS_OK) return rc;
        CUDA_OK(copy_segments(0, 1, s));
        CUDA_OK(cudaStreamWaitEvent(s, h->io_copy_done, 0));
    } else {
        CUDA_OK(copy_segments(0, 5, s));
    }
    return HS_OK;
}

int hs_step_host_io(hs_handle* h, const hs_host_io* io, int action_is_raw, const uint8_t* reset_pid, const hs_tp_weights* w,
                    float* staging_dev, void* stream) {
    const int rc = hs_step_host_io_async(h, io, action_is_raw, reset_pid, w, staging_dev, stream);
    if (rc != HS_OK) return rc;
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return HS_OK;
}

int hs_host_io_wait(hs_handle* h, void* stream) {
    if (!h) return set_err(HS_ERR_INVALID, "hs_host_io_wait: null handle%s");
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return HS_OK;
}

int hs_step_host_io_async(hs_handle* h, const hs_host_io* io, int action_is_raw, const uint8_t* reset_pid, const hs_tp_weights* w,
                          float* staging_dev, void* stream) {
    if (!h || !io || !io->action || !staging_dev) return set_err(HS_ERR_INVALID, "hs_step_host_io: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_step_host_io: call hs_bind_buffers first%s");
    if (h->cfg.use_tp_net && !w) return set_err(HS_ERR_INVALID, "hs_step_host_io: use_tp_net == 1 needs the predictor weights%s");
    cudaStream_t s = (cudaStream_t)stream;
    if (!h->io_graph_mode) {
        return host_io_enqueue(h, io, action_is_raw, reset_pid, w, staging_dev, stream);
    }
    // one graph launch per tick: the seven stream calls above cost ~25 us of host time per tick at 4096 envs.
    // A graph is valid for one exact set of pointers (host buffers, bound output set, weights) -> small LRU cache.
    hs_handle::IoKey key;
    memset(&key, 0, sizeof(key));
    key.io = *io; key.bufs = h->bufs;
    if (w) { key.w = *w; key.has_w = 1; }
    key.staging = staging_dev; key.reset_pid = reset_pid; key.raw = action_is_raw;
    key.tp_init = (h->cfg.use_tp_net && h->tp_frames == 0) ? 1 : 0;
    key.variant = h->tp_variant;
    hs_handle::IoGraph* slot = nullptr;
    for (auto& g : h->io_graphs)
        if (g.exec && memcmp(&g.key, &key, sizeof(key)) == 0) { slot = &g; break; }
    const int kernels = h->cfg.use_tp_net ? 2 : 1;
    if (slot) {
        h->launches += kernels;                              // the captured calls counted themselves once, at capture
        if (h->cfg.use_tp_net) h->tp_frames += 1;
    } else {
        slot = &h->io_graphs[0];
        for (auto& g : h->io_graphs) {
            if (!g.exec) { slot = &g; break; }
            if (g.last_use < slot->last_use) slot = &g;
        }
        if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
        if (!h->io_capture) CUDA_OK(cudaStreamCreateWithFlags(&h->io_capture, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamBeginCapture(h->io_capture, cudaStreamCaptureModeThreadLocal));
        const int rc = host_io_enqueue(h, io, action_is_raw, reset_pid, w, staging_dev, h->io_capture);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(h->io_capture, &graph);
        if (rc != HS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess || !graph) return set_err(HS_ERR_CUDA, "hs_step_host_io: stream capture failed: %s", cudaGetErrorString(ce));
        const cudaError_t ie = cudaGraphInstantiate(&slot->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { slot->exec = nullptr; return set_err(HS_ERR_CUDA, "hs_step_host_io: cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
        slot->key = key;
    }
    slot->last_use = ++h->io_clock;
    CUDA_OK(cudaGraphLaunch(slot->exec, s));
    return HS_OK;
}

int hs_step_host_io_many(hs_host_batch* batches, int32_t num_batches, int32_t num_ticks, int32_t in_flight,
                         int action_is_raw, const hs_tp_weights* w, hs_obs_callback on_obs, void* user) {
    if (!batches || num_batches < 1 || num_ticks < 0 || in_flight < 1 || in_flight > num_batches)
        return set_err(HS_ERR_INVALID, "hs_step_host_io_many: need 1 <= in_flight <= num_batches and num_ticks >= 0%s");
    for (int32_t b = 0; b < num_batches; ++b)
        if (!batches[b].h || !batches[b].sets || !batches[b].ios || batches[b].num_sets < 1 || batches[b].next_set < 0 ||
            batches[b].next_set >= batches[b].num_sets)
            return set_err(HS_ERR_INVALID, "hs_step_host_io_many: incomplete batch descriptor%s");
    for (int32_t i = 0; i < num_ticks + in_flight - 1; ++i) {
        if (i < num_ticks) {
            hs_host_batch& B = batches[i % num_batches];
            int rc = hs_bind_buffers(B.h, &B.sets[B.next_set]);
            if (rc != HS_OK) return rc;
            rc = hs_step_host_io_async(B.h, &B.ios[B.next_set], action_is_raw, nullptr, w, B.staging_dev, B.stream);
            if (rc != HS_OK) return rc;
            B.next_set = (B.next_set + 1) % B.num_sets;
        }
        const int32_t j = i - (in_flight - 1);                   // the tick whose results the host needs now
        if (j >= 0) {
            hs_host_batch& W = batches[j % num_batches];
            CUDA_OK(cudaStreamSynchronize((cudaStream_t)W.stream));
            if (on_obs) on_obs(user, j % num_batches);
        }
    }
    return HS_OK;
}

int hs_hover_post(hs_handle* h, const hs_hover_params* p, const hs_hover_io* io, void* stream) {
    if (!h || !p || !io || !io->observation || !io->stats || !io->state || !io->target_heading)
        return set_err(HS_ERR_INVALID, "hs_hover_post: null argument%s");
    if (!h->bound) return set_err(HS_ERR_UNBOUND, "hs_hover_post: call hs_bind_buffers first%s");
    if (h->cfg.num_agents != 1) return set_err(HS_ERR_INVALID, "hs_hover_post: the handle must have num_agents == 1%s");
    if (p->with_reward && (!io->reward || !io->done)) return set_err(HS_ERR_INVALID, "hs_hover_post: reward / done missing%s");
    KParams P = make_params(h);
    const unsigned grid = (unsigned)(((int64_t)h->cfg.num_envs + 127) / 128);
    hs_hover_post_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, *p, *io);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
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
    hs_field_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h->bufs.arena, ND * h->cfg.num_agents + E_CYL + 3 * h->cfg.num_cylinders, row0, n_slots, width, ss, sc,
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
            if (value < -1 || value > 5) return set_err(HS_ERR_INVALID, "predictor variant must be -1 (auto), 0 (fp32 FFMA), 1 (3xTF32 mma.sync), 2 (tcgen05, 128-env tiles), 3 (tcgen05, 32-env tiles), 4 (tcgen05, 2 x 32-env tiles ping-pong) or 5 (tcgen05, 32-env tile as 2 x 16-env halves ping-pong)%s");
            h->tp_variant = value;
            return HS_OK;
        case HS_OPT_FUSED_TICK:
            if (value != 0 && value != 1) return set_err(HS_ERR_INVALID, "HS_OPT_FUSED_TICK must be 0 or 1%s");
            h->fused_tick = value;
            for (auto& g : h->io_graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            return HS_OK;
        case HS_OPT_HOST_IO_ZERO_COPY_ACTION:
            if (value != 0 && value != 1) return set_err(HS_ERR_INVALID, "HS_OPT_HOST_IO_ZERO_COPY_ACTION must be 0 or 1%s");
            h->io_zero_copy_action = value;
            for (auto& g : h->io_graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            return HS_OK;
        case HS_OPT_HOST_IO_GRAPH:
            if (value != 0 && value != 1) return set_err(HS_ERR_INVALID, "HS_OPT_HOST_IO_GRAPH must be 0 or 1%s");
            h->io_graph_mode = value;
            return HS_OK;
        case HS_OPT_ROLLOUT_VARIANT:
            if (value < 0 || value > 3) return set_err(HS_ERR_INVALID, "HS_OPT_ROLLOUT_VARIANT must be 0 (auto) or the number of ticks per predictor pass: 1, 2, 3%s");
            h->rollout_variant = value;
            return HS_OK;
        case HS_OPT_TICK_MAPPING:
            if (value < 0 || value > 2) return set_err(HS_ERR_INVALID, "HS_OPT_TICK_MAPPING must be 0 (auto), 1 (4 lanes per env) or 2 (one lane per env)%s");
            if (value == 1 && (h->cfg.num_agents > NARROW_MAX_AGENTS || (h->cfg.use_obstacles && h->cfg.use_tp_net)))
                return set_err(HS_ERR_INVALID, "HS_OPT_TICK_MAPPING = 1: the 4-lane mapping covers num_agents <= 3%s");
            if (value == 2 && (h->cfg.num_agents < 3 || (h->cfg.num_envs & 3) != 0))
                return set_err(HS_ERR_INVALID, "HS_OPT_TICK_MAPPING = 2: the one-lane mapping needs num_agents >= 3 and num_envs % 4 == 0%s");
            h->tick_mapping = value;
            for (auto& g : h->io_graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            return HS_OK;
        case HS_OPT_EXACT_MATH:
            if (value != 0 && value != 1) return set_err(HS_ERR_INVALID, "HS_OPT_EXACT_MATH must be 0 or 1%s");
            h->exact_math = value;
            for (auto& g : h->io_graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            return HS_OK;
        default:
            return set_err(HS_ERR_INVALID, "unknown option%s");
    }
}

}  // extern "C"

