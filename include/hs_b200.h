/*
 * hs_b200.h -- C ABI of the B200-native HideAndSeek environment step.
 *
 * This library replaces, for ONE path, what the reference reaches through Isaac Sim /
 * PhysX plus ~300 eager torch launches per tick (all paths relative to the reference
 * tree thu-uav/Multi-UAV-pursuit-evasion @ ba29682):
 *
 *   hs_step_pre   <- PIDRateController._inv_call      omni_drones/utils/torchrl/transforms.py:425-459
 *                    PIDRateController.forward        omni_drones/controllers/lee_position_controller.py:476-550
 *                    HideAndSeek._pre_sim_step        omni_drones/envs/hide_and_seek/hideandseek.py:725-744
 *                    MultirotorBase.apply_action      omni_drones/robots/drone/multirotor.py:466-508
 *                    RotorGroup.forward               omni_drones/actuators/rotor_group.py:55-71
 *                    _get_dummy_policy_prey           omni_drones/envs/hide_and_seek/hideandseek.py:1067-1141
 *                    SimulationContext.step (PhysX)   omni_drones/envs/isaac_env.py:233-234
 *                    _compute_state_and_obs           omni_drones/envs/hide_and_seek/hideandseek.py:746-917
 *                    _compute_reward_and_done         omni_drones/envs/hide_and_seek/hideandseek.py:919-1065
 *   hs_step_post  <- the part of _compute_state_and_obs that depends on the trajectory
 *                    predictor output                 omni_drones/envs/hide_and_seek/hideandseek.py:834-887
 *   hs_reset      <- IsaacEnv._reset + _reset_idx     omni_drones/envs/isaac_env.py:210-225,
 *                                                     omni_drones/envs/hide_and_seek/hideandseek.py:698-723,
 *                                                     omni_drones/robots/drone/multirotor.py:635-650
 *   views         <- omni.physics.tensors get/set     omni_drones/views/articulation_view.py:211-251,
 *   (hs_state_*)                                      omni_drones/views/rigid_prim_view.py:61-198
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types.
 *   - every `float*` / `uint8_t*` below is a DEVICE pointer unless the name ends in `_host`.
 *     The library never allocates or frees them; the caller (PyTorch on the Python side)
 *     owns all buffers and keeps them alive while the handle uses them.
 *   - every entry point returns 0 on success or a negative hs_status; hs_last_error()
 *     gives a message for the calling thread.  Nothing throws across the boundary.
 *   - kernels are launched asynchronously on the `stream` argument (a cudaStream_t passed
 *     as void*; NULL = legacy default stream).  No entry point synchronises the device
 *     except hs_step_host (which must, it returns host data).
 *   - one host thread per handle; a handle is bound to the CUDA device current at create.
 *   - there is NO CPU fallback: without a CUDA device hs_create fails with HS_ERR_NO_DEVICE.
 *
 * Layouts (E envs, A pursuers, C cylinders, K observed cylinders, F predicted steps,
 * H history frames, D = 20 + 3F if use_tp_net else 20).  "AoS" tensors are exactly the
 * reference's tensordict entries, row-major, contiguous.
 */
#ifndef HS_B200_H
#define HS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_ABI_VERSION 3
#define HS_NUM_STATS 24
#define HS_MAX_AGENTS 6          /* 1..3: both tick mappings and every fused kernel; 4..6: one-lane-per-env tick only */
#define HS_MAX_CYLINDERS 8
#define HS_MAX_OBS_CYLINDERS 4
#define HS_MAX_FUTURE 8

typedef enum hs_status {
    HS_OK = 0,
    HS_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    HS_ERR_NO_DEVICE = -2,   /* no CUDA device: this library has no CPU path */
    HS_ERR_CUDA = -3,        /* a CUDA runtime call failed; see hs_last_error() */
    HS_ERR_UNBOUND = -4      /* hs_bind_buffers has not been called */
} hs_status;

/* Stats slot order = declaration order of the reference's stats spec
 * (hideandseek.py:400-425).  stats buffer layout is [HS_NUM_STATS][E] so that each key
 * is a contiguous [E] vector (viewed as [E,1] by the host). */
enum {
    HS_STAT_SUCCESS = 0, HS_STAT_COLLISION, HS_STAT_BLOCKED, HS_STAT_DISTANCE_REWARD,
    HS_STAT_DISTANCE_PREDICTED_REWARD, HS_STAT_SPEED_REWARD, HS_STAT_COLLISION_REWARD,
    HS_STAT_COLLISION_WALL, HS_STAT_COLLISION_CYLINDER, HS_STAT_COLLISION_DRONE,
    HS_STAT_DETECT_REWARD, HS_STAT_CATCH_REWARD, HS_STAT_SMOOTHNESS_REWARD,
    HS_STAT_SMOOTHNESS_MEAN, HS_STAT_SMOOTHNESS_MAX, HS_STAT_FIRST_CAPTURE_STEP,
    HS_STAT_SUM_DETECT_STEP, HS_STAT_RETURN, HS_STAT_ACTION_ERROR_MEAN,
    HS_STAT_ACTION_ERROR_MAX, HS_STAT_TARGET_PREDICTED_ERROR, HS_STAT_DISTANCE_THRESHOLD_L,
    HS_STAT_OUT_OF_ARENA, HS_STAT_SMOOTHNESS_COEF
};

/* Task + vehicle constants.  Defaults (hs_default_config) = cfg/task/HideAndSeek.yaml,
 * omni_drones/robots/assets/usd/crazyflie.yaml and the USD rigid-body constants. */
typedef struct hs_config {
    int32_t abi_version;        /* must be HS_ABI_VERSION */
    int32_t num_envs;           /* E */
    int32_t num_agents;         /* A, 1..HS_MAX_AGENTS */
    int32_t num_cylinders;      /* C = cylinder.max_num, 0..HS_MAX_CYLINDERS */
    int32_t obs_max_cylinder;   /* K <= min(C, HS_MAX_OBS_CYLINDERS) */
    int32_t future_step;        /* F */
    int32_t history_step;       /* H */
    int32_t max_episode_length;
    int32_t use_tp_net;         /* 1: state_self/state_drones carry 3F prediction slots, written by hs_step_post */
    int32_t smoothness_gated;   /* 1: smoothness reward zeroed (HideAndSeek with use_deployment=0) */
    int32_t write_smoothness_coef_stat; /* 1 for HideAndSeek, 0 for the envgen variant */
    int32_t fixed_yaw;
    int32_t ground_clamp;
    int32_t use_obstacles;      /* 1: every TP frame also carries [x, y, size] of the C cylinders (hideandseek.py:808-817);
                                   frame width 7 + 3A + 3C.  Runs on the one-lane-per-env tick (num_agents >= 3) with the
                                   predictor as a module between hs_step_pre and hs_step_post */
    int32_t contact_mode;       /* 0 (default): ground clamp only.  1: + analytic inelastic contacts of every pursuer with the
                                   standing cylinders (2-D, radius cylinder_size + drone_radius, below the cylinder top) and
                                   with the evader's sphere (at its position when the tick starts): projection out of the
                                   overlap, inward normal velocity removed.  PhysX resolves these contacts in the reference
                                   (robot.py:125-135 colliders); like the integrator itself this response is a documented
                                   stand-in, PARITY UNPINNED.  Pursuer-pursuer contacts are not modelled. */
    int32_t reserved_i[1];
    float dt;
    float arena_size, max_height, cylinder_size, catch_radius, collision_radius;
    float drone_detect_radius, target_detect_radius, v_drone, mask_value;
    float dist_reward_coef, catch_reward_coef, detect_reward_coef, collision_coef, speed_coef;
    float smoothness_coef;
    float target_clip, max_thrust_ratio;
    float pid_kp[3], pid_ki[3], pid_kd[3], pid_ilimit[3], pid_out_limit;
    float kf, km, rotor_alpha;  /* KF, KM per rotor; alpha = dt / clamp(tau,0,1) */
    float rotor_dirs[4];
    float rotor_x[4], rotor_y[4];
    float drag_coef_times_mass; /* drag_coef * base-link mass (multirotor.py:495) */
    float downwash_kr, downwash_kz;
    float total_mass, inertia[3], gravity;
    float lin_damp_factor, ang_damp_factor;   /* max(0, 1 - dt*c) */
    float max_linear_velocity, max_angular_velocity;
    float ground_z;
    float hover_throttle;       /* sqrt(total_mass*9.81 / (4 KF)), multirotor.py:647-648 */
    /* Derived constants.  The reference forms these from Python doubles and only then
     * rounds to fp32 (e.g. `self.arena_size**2`), so the host computes them in double. */
    float arena_size_sq;        /* arena_size^2            hideandseek.py:980,1096 */
    float half_arena;           /* 0.5*arena_size          hideandseek.py:835,841 */
    float coll_radius_x2;       /* 2*collision_radius      hideandseek.py:975 */
    float vmax_clamped;         /* max_linear_velocity*(1-1e-6): see DESIGN.md "Integrator" */
    float inv_inertia[3];       /* 1/inertia */
    float drone_radius;         /* 0.06: base_link collider (USD cylinder r = 0.06, h = 0.025) */
    float evader_radius;        /* 0.05: the evader's sphere (hideandseek.py:544-551) */
    float reserved_f[1];
} hs_config;

/* Device buffers.  All owned by the caller.
 * State arena (private to the library; SoA so that consecutive envs are consecutive
 * addresses): `arena` holds hs_arena_floats(cfg) floats.  Everything else is a
 * reference-facing tensor in the reference's own AoS layout. */
typedef struct hs_buffers {
    float* arena;               /* state arena, see hs_arena_floats / hs_state_get */
    float* stats;               /* [HS_NUM_STATS][E] */
    /* outputs of hs_step_pre / hs_reset (observation side) */
    float* state_self;          /* [E,A,1,D]  ("agents","observation","state_self") */
    float* state_others;        /* [E,A,A-1,3] (may be NULL when A == 1) */
    float* obs_cylinders;       /* [E,A,K,5]  (also ("agents","state","cylinders")) */
    float* state_drones;        /* [E,A,D]    ("agents","state","state_drones") */
    float* tp_input;            /* [E,H,7+3A] written: previous frames shifted + new frame */
    const float* tp_input_prev; /* [E,H,7+3A] read; may equal tp_input (in-place shift) */
    float* tp_groundtruth;      /* [E,3] */
    uint8_t* tp_done;           /* [E,1] bool */
    float* reward;              /* [E,A,1] */
    uint8_t* done;              /* [E,1] bool */
    uint8_t* truncated;         /* [E,1] bool, written by hs_reset only (may be NULL) */
    float* drone_state;         /* [E,A,13] ("info","drone_state") */
    float* prev_action;         /* [E,A,4]  ("info","prev_action"); also the PID transform's memory */
    /* by-products of the fused action transform (keys the reference's transform writes) */
    float* rotor_cmds;          /* [E,A,4]  post-transform ("agents","action") */
    float* ctbr;                /* [E,A,4]  'ctbr' */
    float* target_rate;         /* [E,A,3]  'target_rate' */
    float* action_error;        /* [E,A]    ("stats","action_error_order1") */
    const float* v_prey;        /* device scalar: current evader speed (curriculum, hideandseek.py:1012-1015) */
    const float* smoothness_coef; /* device scalar or NULL (= cfg.smoothness_coef): min(max_smoothness_coef,
                                   init_smoothness_coef + smooth_lr * update_epoch), recomputed by the host whenever
                                   `update_epoch` changes (hideandseek.py:988-991; scripts/train_deploy.py writes
                                   base_env.update_epoch every iteration).  Read by every tick, so captured CUDA
                                   graphs follow the curriculum without re-capture. */
    float* throttle_diff;       /* [E,A] or NULL: |throttle_t - throttle_{t-1}|_2 per pursuer = drone.throttle_difference
                                   (multirotor.py:480-484), an output for callers that log it (Hover's action_smoothness) */
    /* Ring form of the TP window (optional, both or neither; wide tick mapping, num_agents >= 3).  When set, the tick
     * does NOT shift a chronological [E,H,FD] tensor every tick (576 B read + 640 B written per env for the reference's
     * shape, hideandseek.py:819-831 re-stacks its deque); it writes the new frame twice, at slots p and p + H of
     * tp_ring [E, 2H, FD], and advances p = tp_ring_pos[e / 32].  The chronological window of env e after the tick is
     * the CONTIGUOUS span tp_ring[e, p' : p' + H, :] with p' the advanced position - a strided [E,H,FD] view with env
     * stride 2*H*FD, valid until the next tick; hs_step_post_tp reads it in place.  tp_input / tp_input_prev are
     * ignored (may be NULL). */
    float* tp_ring;             /* [E, 2H, FD] or NULL */
    int32_t* tp_ring_pos;       /* [ceil(E/32)] zero-initialised by the caller, or NULL */
} hs_buffers;

typedef struct hs_handle hs_handle;

/* ---- lifecycle ---------------------------------------------------------------------- */
int hs_abi_version(void);
const char* hs_last_error(void);
/* Fill `cfg` with the reference defaults for E envs. */
int hs_default_config(hs_config* cfg, int32_t num_envs);
/* Number of floats the state arena needs for this config. */
int64_t hs_arena_floats(const hs_config* cfg);
int hs_create(const hs_config* cfg, hs_handle** out);
int hs_destroy(hs_handle* h);
/* (Re)bind device buffers.  Cheap: may be called before every step to rotate output sets. */
int hs_bind_buffers(hs_handle* h, const hs_buffers* bufs);

/* ---- the hot path ------------------------------------------------------------------- */
/* One control tick for every env.
 *   action        [E,A,4]  device.  action_is_raw != 0: policy output before tanh; the CTBR
 *                 transform + rate PID run inside the kernel (reference path with
 *                 action_transform: PIDrate).  action_is_raw == 0: rotor commands in [-1,1]
 *                 applied directly (base env stepped without the transform); then
 *                 action_error must be supplied by the caller in bufs.action_error.
 *   reset_pid     [E] bool device or NULL: the `done` entry of the tensordict handed to
 *                 step (transforms.py:453); true rows zero the PID memory first. */
int hs_step_pre(hs_handle* h, const float* action, int action_is_raw,
                const uint8_t* reset_pid, void* stream);
/* Second half, only when use_tp_net: tp_pred [E,3F] is TP_net's tanh output for
 * bufs.tp_input; writes state_self / state_drones (hideandseek.py:834-887). */
int hs_step_post(hs_handle* h, const float* tp_pred, void* stream);
/* Same second half, but with the trajectory predictor fused in: the kernel evaluates
 * TP_net (one-layer LSTM, hidden 64, zero initial state, last step -> Linear -> tanh;
 * omni_drones/learning/mappo.py:572-589) on bufs.tp_input straight from the module's live
 * parameter tensors and writes state_self / state_drones.  Replaces the cuDNN LSTM call the
 * reference makes inside _compute_state_and_obs (hideandseek.py:834).  tp_pred_out [E,3F]
 * receives the raw prediction when non-NULL. */
typedef struct hs_tp_weights {
    const float* weight_ih;     /* [4*64, 7+3A]  lstm.weight_ih_l0, gate order i,f,g,o */
    const float* weight_hh;     /* [4*64, 64]    lstm.weight_hh_l0 */
    const float* bias_ih;       /* [4*64]        lstm.bias_ih_l0 */
    const float* bias_hh;       /* [4*64]        lstm.bias_hh_l0 */
    const float* fc_weight;     /* [3F, 64]      fc.weight */
    const float* fc_bias;       /* [3F]          fc.bias */
    int32_t input_size, hidden_size, output_size;
    int32_t reserved;
} hs_tp_weights;
int hs_step_post_tp(hs_handle* h, const hs_tp_weights* w, float* tp_pred_out, void* stream);
/* hs_step_pre + hs_step_post_tp as one call - and, when the batch has at most one 32-env tile per
 * SM (<= 4736 envs on a B200), the predictor policy is auto or 5, and A <= 3, as ONE kernel launch
 * (hs_tick_tp_fused_kernel: four warps of each CTA run the tick of the tile's envs while the others
 * stage the predictor's weights, then the CTA runs the tcgen05 predictor on the TP_input tiles still
 * in shared memory).  Results are identical to the two-call sequence.  Arguments as in hs_step_pre
 * and hs_step_post_tp. */
int hs_step_fused(hs_handle* h, const float* action, int action_is_raw, const uint8_t* reset_pid,
                  const hs_tp_weights* w, float* tp_pred_out, void* stream);
/* Partial reset.  env_mask [E] bool (NULL = all).  Initial poses are injected (sampling
 * stays on the host side so that it can follow the reference's RNG streams):
 *   drone_pos [E,A,3], drone_rot [E,A,4] (wxyz), target_pos [E,3], cyl_pos [E,C,3];
 *   rows outside the mask are ignored.  Performs the reference's extra unforced physics
 *   tick for ALL envs (hideandseek.py:722-723), zeroes progress/stats of the masked
 *   envs and writes the observation tensors (state_self/state_drones too when
 *   use_tp_net == 0; otherwise call hs_step_post afterwards). */
int hs_reset(hs_handle* h, const uint8_t* env_mask, const float* drone_pos,
             const float* drone_rot, const float* target_pos, const float* cyl_pos,
             void* stream);

/* ---- convenience: host-buffer tick (the e2e path) ------------------------------------ */
/* action_host [E,A,4] host (pinned for full speed) -> H2D -> hs_step_pre -> D2H of
 * reward [E,A] and done [E].  Only valid when use_tp_net == 0 (otherwise the predictor
 * runs between the halves on the caller's side).  Synchronises `stream`. */
int hs_step_host(hs_handle* h, const float* action_host, int action_is_raw,
                 float* reward_host, uint8_t* done_host, float* staging_dev, void* stream);

/* Host-buffer tick INCLUDING the predictor (the end-to-end path of bench.py): H2D of `action`
 * -> hs_step_pre -> hs_step_post_tp (when use_tp_net; `w` may be NULL otherwise) -> D2H of every
 * non-NULL output below -> stream synchronise.  Outputs that are neighbours on the device AND in
 * host memory with the same spacing (e.g. a host mirror of the caller's output slab) leave in a
 * single copy.  `reset_pid` is a DEVICE pointer like in hs_step_pre (or NULL). */
typedef struct hs_host_io {
    const float* action;      /* in : [E,A,4] host, pinned for full speed */
    float* state_self;        /* out: [E,A,D]      D = 20 (+3F with use_tp_net)   (any of these may be NULL) */
    float* state_others;      /* out: [E,A,A-1,3] */
    float* obs_cylinders;     /* out: [E,A,K,5] */
    float* reward;            /* out: [E,A] */
    uint8_t* done;            /* out: [E] */
} hs_host_io;
int hs_step_host_io(hs_handle* h, const hs_host_io* io, int action_is_raw, const uint8_t* reset_pid,
                    const hs_tp_weights* w, float* staging_dev, void* stream);
/* The same call without the final synchronisation, and the wait that completes it: lets a host that
 * drives several env batches (one handle and one stream each) keep one batch's copies on the PCIe
 * link while another batch computes.  The host buffers of `io` are valid after hs_host_io_wait. */
int hs_step_host_io_async(hs_handle* h, const hs_host_io* io, int action_is_raw, const uint8_t* reset_pid,
                          const hs_tp_weights* w, float* staging_dev, void* stream);
int hs_host_io_wait(hs_handle* h, void* stream);
/* K host-buffer ticks driven from ONE call: the loop an actor process runs around hs_step_host_io_async /
 * hs_host_io_wait - rotate over `num_batches` independent env batches, keep `in_flight` of them between issue and wait
 * (one batch's observation is on the PCIe link while the next one computes), rotate each batch's output sets - without a
 * round trip through the host language per tick (the reference's collector is such a loop in Python,
 * omni_drones/utils/torchrl/collector.py:33-38).  Tick i runs batch i % num_batches: bind sets[next_set], issue
 * hs_step_host_io_async with ios[next_set] on the batch's stream, then wait for the batch issued `in_flight` ticks earlier
 * and call on_obs(user, that batch index) - the place where a host-side policy reads the observation now in host memory
 * and writes the batch's next action (NULL: the action buffers are left as they are). */
typedef struct hs_host_batch {
    hs_handle* h;
    const hs_buffers* sets;          /* [num_sets] complete buffer tables, set s reading tp_input_prev from set s-1 */
    const hs_host_io* ios;           /* [num_sets] host destinations of each set (and the batch's action buffer) */
    int32_t num_sets;
    int32_t next_set;                /* in/out: the set the batch's next tick writes */
    float* staging_dev;              /* [E,A,4] device scratch (used when the action buffer is pageable) */
    void* stream;                    /* the batch's stream */
} hs_host_batch;
typedef void (*hs_obs_callback)(void* user, int32_t batch);
int hs_step_host_io_many(hs_host_batch* batches, int32_t num_batches, int32_t num_ticks, int32_t in_flight,
                         int action_is_raw, const hs_tp_weights* w, hs_obs_callback on_obs, void* user);

/* ---- state views (what the get/set methods of omni_drones/views did) ------------------- */
enum {
    HS_FIELD_DRONE_POS = 0,   /* [E,A,3] */
    HS_FIELD_DRONE_ROT,       /* [E,A,4] wxyz */
    HS_FIELD_DRONE_LINVEL,    /* [E,A,3] world */
    HS_FIELD_DRONE_ANGVEL,    /* [E,A,3] world */
    HS_FIELD_THROTTLE,        /* [E,A,4] */
    HS_FIELD_PID_INTEG,       /* [E,A,3] */
    HS_FIELD_PID_LAST_RATE,   /* [E,A,3] */
    HS_FIELD_TARGET_POS,      /* [E,3] */
    HS_FIELD_TARGET_VEL,      /* [E,3] */
    HS_FIELD_CYL_POS,         /* [E,C,3] */
    HS_FIELD_PROGRESS,        /* [E] */
    HS_FIELD_COUNT
};
/* Gather an arena field into a contiguous AoS device buffer / scatter it back. */
int hs_state_get(hs_handle* h, int field, float* dst, void* stream);
int hs_state_set(hs_handle* h, int field, const float* src, void* stream);
/* Number of kernels this handle has launched so far (bench `gpu_launches`). */
int64_t hs_launch_count(const hs_handle* h);
/* PPO minibatch gather: make_dataset_naive (omni_drones/learning/mappo.py:493-513) yields, per minibatch,
 * tensordict.reshape(-1)[indices] for every key of the [E, T] rollout batch, i.e. rows addressed by the flat sample id
 * n = env * T + step.  hs_gather_rows produces those rows for up to HS_GATHER_MAX_TENSORS keys in ONE launch from tensors
 * that are merely strided over (env, step) - the time-major [T, E, ...] rollout storage the ticks wrote, or the reference's
 * own [E, T, ...] layout - so the flattened copy the reference makes first never exists.  Pure data movement: the result
 * is bit-identical to the reference's indexing for the same `indices` (the caller draws the permutation, torch.randperm). */
#define HS_GATHER_MAX_TENSORS 24
typedef struct hs_gather_tensor {
    const void* src;            /* element (env 0, step 0) */
    void* dst;                  /* [num_rows][row_bytes] contiguous */
    int64_t stride_env;         /* bytes between consecutive envs of src */
    int64_t stride_step;        /* bytes between consecutive steps of src */
    int32_t row_bytes;          /* bytes of one (env, step) row: contiguous in src */
    int32_t reserved;
} hs_gather_tensor;
int hs_gather_rows(const hs_gather_tensor* tensors, int num_tensors, const int64_t* indices_device, int64_t num_rows,
                   int num_steps, void* stream);

/* A whole rollout of `num_ticks` control ticks (tick + predictor, as hs_step_fused) in ONE kernel launch: what the
 * collector's loop `for t in range(T): td = env.step(td)` (omni_drones/utils/torchrl/collector.py:34-66) does when the
 * actions of the T ticks are already on the device (open-loop action sequences, replayed rollouts, benchmarks) - with a
 * policy in the loop use hs_step_fused per tick.  Every CTA keeps its 32-env tile for the whole rollout: the predictor's
 * weights go to tensor memory once, the TP window stays in shared memory, and the tick of step t+1 runs on dedicated
 * warps while the tensor cores work on the predictor of step t.
 *   sets_device [num_sets]: DEVICE array of buffer tables; tick t writes sets[(first_set + t) % num_sets] (e.g. the rows of
 *     a time-major rollout storage).  The handle must be bound (hs_bind_buffers) to a table with the same arena / stats /
 *     prev_action; tp_input_prev of the tables is ignored: the window before the first tick is first_tp_prev [E,H,7+3A].
 *   action [T][E,A,4] with action_tick_stride floats between ticks (0 = the same action every tick).
 *   tp_pred_out (or NULL) [T][E,3F] with pred_tick_stride floats between ticks.
 * Same results, bit for bit, as num_ticks calls of hs_step_fused.  HS_ERR_INVALID unless num_agents == 3, history_step == 10,
 * use_tp_net, no use_obstacles, fast-math build, plain TP window, and hs_reset has run. */
int hs_rollout_fused(hs_handle* h, const hs_buffers* sets_device, int num_sets, int first_set, const float* first_tp_prev,
                     const float* action, int64_t action_tick_stride, int action_is_raw, int num_ticks, const hs_tp_weights* w,
                     float* tp_pred_out, int64_t pred_tick_stride, void* stream);

/* Tunables.  HS_OPT_PREDICTOR_VARIANT: -1 = auto (default: variant 5 while a launch has at most one
 * 32-env tile per SM, variant 4 above), 0 = fp32 FFMA predictor kernel,
 * 1 = tensor-core predictor, warp-level mma.sync (error-compensated 3xTF32, fp32-level results),
 * 2 = tensor-core predictor, tcgen05.mma with TMEM accumulators and TMEM-resident recurrent operand,
 *     128 envs per CTA (envs on the MMA's M dimension),
 * 3 = tcgen05.mma with the gates on M, weights resident in TMEM, 32 envs per CTA on N (fills the SMs
 *     at small batches),
 * 4 = as 3 with two 32-env tiles ping-ponging per CTA (one tile's MMAs run under the other's cell update),
 * 5 = as 3 with the 32-env tile split into two 16-env halves that ping-pong (N = 16 MMAs, 4 env columns per
 *     epilogue thread and half, mbarrier hand-off): the small-batch default and the predictor half of the
 *     one-launch tick (hs_step_fused).
 * HS_OPT_HOST_IO_GRAPH: 1 (default) = hs_step_host_io replays its copies and kernels as ONE CUDA graph
 * launch, cached per set of pointers (host buffers, bound outputs, weights); 0 = plain stream calls.
 * HS_OPT_HOST_IO_ZERO_COPY_ACTION: 1 (default) = a page-locked io->action is read in place by the tick
 * kernel (UVA), no H2D copy; 0 = always copy into the staging buffer first.
 * HS_OPT_FUSED_TICK: 1 (default) = hs_step_fused may use the one-launch kernel; 0 = always two launches.
 * HS_OPT_EXACT_MATH: 0 (default) = the product tick kernel (rcp/sqrt/ex2.approx, FMA contraction); 1 = the same
 * source built with IEEE round-to-nearest division / square root / exp, no FMA contraction, and the reference's
 * operation order in the ill-conditioned stages (csrc/hs_tick_exact.cu).  About 2x slower; exists so that parity
 * tests can separate rounding amplified by the task's discontinuities from defects. */
enum { HS_OPT_PREDICTOR_VARIANT = 1, HS_OPT_HOST_IO_GRAPH = 2, HS_OPT_HOST_IO_ZERO_COPY_ACTION = 3, HS_OPT_FUSED_TICK = 4,
       HS_OPT_EXACT_MATH = 5, HS_OPT_TICK_MAPPING = 6, HS_OPT_ROLLOUT_VARIANT = 7 };
/* HS_OPT_ROLLOUT_VARIANT: which kernel runs a fused rollout = how many ticks the predictor warps advance per pass.
 * 0 (default) = auto; 2, 3 = the full 32-env MMA tiles of 2 or 3 consecutive ticks ping-pong on the tensor pipe (half
 * the tensor-pipe instructions of variant 1); 1 = one tick at a time as two 16-env halves.  Same results. */
/* HS_OPT_TICK_MAPPING: which work decomposition hs_step_pre / hs_reset use for the tick kernel.
 *   0 (default) = auto: 4 lanes per env (hs_tick_kernel, latency-bound small batches) below 32768 envs, one lane per
 *       env (hs_tick_wide_kernel: TMA tensor loads of the SoA state tile, bulk stores of the outputs; bandwidth-bound
 *       large batches) from 32768 envs on and always for num_agents > 3;
 *   1 = always 4 lanes per env (num_agents <= 3 only);  2 = always one lane per env.
 * Both give bit-identical results.  The one-lane mapping needs num_envs % 4 == 0 (16-byte row pitch of the stats
 * tensor map) and distinct tp_input / tp_input_prev buffers; otherwise auto stays with the 4-lane kernel. */
int hs_set_option(hs_handle* h, int option, int value);

/* ---- Hover (BASELINE config 1): observation / reward / stats post-kernel ------------------ */
/* omni_drones/envs/single/hover.py:334-523 after a tick of a handle created with num_agents == 1, num_cylinders == 0:
 * _pre_sim_step's logging stats (:334-359), _compute_state_and_obs (:361-437) and _compute_reward_and_done (:439-523) in
 * one launch, one thread per env.  Reads what the tick left in the handle's bound buffers (drone_state, rotor_cmds, ctbr,
 * target_rate, throttle_diff) and the arena (progress, throttle).  Stats slot order = the declaration order of the
 * reference's stats spec (:239-279), stats layout [HS_HOVER_NUM_STATS][E]. */
#define HS_HOVER_NUM_STATS 39
#define HS_HOVER_NUM_STATE 12
typedef struct hs_hover_params {
    float reward_distance_scale, reward_v_scale, reward_acc_scale, reward_jerk_scale;
    float linear_vel_max, linear_acc_max;
    float alpha;                        /* 0.8: stats.lerp_(x, 1 - alpha)  hover.py:498-501 */
    float target_pos[3];                /* (0, 0, 1)  hover.py:149 */
    int32_t time_encoding;              /* append progress / max_episode_length x 4 */
    int32_t omega, motor;               /* append the world angular velocity / throttle * 2 - 1 (cfg.task.omega / motor) */
    int32_t with_reward;                /* 0: the observation half only (what _reset runs), 1: the whole step */
} hs_hover_params;
typedef struct hs_hover_io {
    float* observation;                 /* [E,1,16 (+3 omega) (+4 motor) (+4 time)] */
    float* reward;                      /* [E,1,1] */
    uint8_t* done;                      /* [E,1] */
    float* stats;                       /* [HS_HOVER_NUM_STATS][E] read-modify-write */
    float* state;                       /* [HS_HOVER_NUM_STATE][E]: last linear/angular v, a, jerk; the six episode sums */
    const float* target_heading;        /* [E,3] */
} hs_hover_io;
int hs_hover_post(hs_handle* h, const hs_hover_params* p, const hs_hover_io* io, void* stream);

/* ---- device-side reset sampler --------------------------- */
/* Replaces the sampling half of HideAndSeek._reset_idx for use_random_cylinder == 1
 * (hideandseek.py:609-697): the uniform pose draws (:283-303, 616-629, 696-697) and
 * rejection_sampling_random_cylinder + select_unoccupied_positions (:576-607, 106-119 - a
 * per-env host loop over torch.randperm in the reference).  Same distribution; the random
 * stream is a counter-based Philox4x32-10: words = philox(counter = (global env index, block,
 * epoch lo, epoch hi), key = seed), so a draw depends only on (seed, epoch, env_offset + e) -
 * not on the batch size, the shard or the reset mask.  Draw order: oracle/reset_sampler.py. */
typedef struct hs_reset_dist {
    float drone_lo[2], drone_hi[2];     /* init_drone_pos_dist   hideandseek.py:283-286 */
    float target_lo[2], target_hi[2];   /* init_target_pos_dist  hideandseek.py:287-290 */
    float z_lo, z_hi;                   /* init_*_pos_dist_z     hideandseek.py:291-298 */
    float rpy_lo[3], rpy_hi[3];         /* init_rpy_dist         hideandseek.py:300-309 */
    float grid_size;                    /* 2 * cylinder.size     hideandseek.py:578 */
    int32_t num_grid;                   /* int(arena_size * 2 / grid_size), <= 11  :579 */
    float boundary;                     /* clamp of the cylinder xy, grid_to_continuous :139 */
    float cyl_z_active, cyl_z_inactive; /* 0.5 * cylinder.height / invalid_z       :685-689 */
    int32_t min_cylinders;              /* active count ~ U{min_cylinders..num_cylinders} :598 */
    int32_t fixed_num;                  /* >= 0: use_fixed_num (:595-596); -1: random */
    int32_t fixed_xy;                   /* use_eval: xy from fixed_* instead of the draws (:618-627) */
    float fixed_drone_xy[3][2];
    float fixed_target_xy[2];
    int64_t env_offset;                 /* global index of this handle's env 0 (sharded jobs) */
    uint64_t seed;
} hs_reset_dist;
/* Writes drone_pos [E,A,3], drone_rot [E,A,4] wxyz, target_pos [E,3], cyl_pos [E,C,3] and
 * n_active [E] (float, `active_cylinders`) - the arguments hs_reset takes - for ALL envs.
 * One launch, asynchronous on `stream`.  HS_ERR_INVALID when the grid cannot hold
 * num_cylinders free cells for every env (the reference raises ValueError, :111-112). */
int hs_sample_reset(hs_handle* h, const hs_reset_dist* dist, uint64_t epoch, float* drone_pos,
                    float* drone_rot, float* target_pos, float* cyl_pos, float* n_active,
                    void* stream);

/* ---- HideAndSeek_envgen control plane on the device ---------- */
/* The two loops of GenBuffer that are O(archive) / O(E) host work in the reference
 * (omni_drones/envs/hide_and_seek/hideandseek_envgen.py): no env handle is involved. */
typedef struct hs_gen_params {
    int32_t num_agents, num_cylinders;  /* task = [pursuer xyz * A, evader xyz, cylinder xyz * C] */
    float arena_size;                   /* GenBuffer.arena_size      :226 */
    float grid_size;                    /* 2 * cylinder_size         :228 */
    float max_height;                   /* GenBuffer.max_height      :229 */
    int32_t num_grid;                   /* int(arena_size * 2 / grid_size), <= 11 :230 */
    int32_t expand_cylinders;           /* task.expand_cylinders */
    float expand_step;                  /* task.expand_step */
    uint64_t seed;
    int64_t task_offset;                /* global index of this call's task 0 (sharded jobs: ranks draw disjoint streams) */
} hs_gen_params;
/* GenBuffer.samplenearby (:322-372): for each of num_tasks outputs pick an archive row of
 * `history` [n_history, 3A+3+3C] uniformly, perturb, clip to the task bounds, accept when
 * sanity_check (:185-207) passes, at most 10 attempts.  tasks_out [num_tasks, dim];
 * valid_out[i] = 0 when all ten attempts failed (the caller re-draws those rows from the
 * valid ones, :361-366).  Counter-based Philox stream (task, attempt, epoch), see
 * oracle/envgen_oracle.py.  One launch, asynchronous. */
int hs_gen_sample_nearby(const hs_gen_params* p, const float* history, int64_t n_history,
                         int64_t num_tasks, uint64_t epoch, float* tasks_out, uint8_t* valid_out,
                         void* stream);
/* Greedy farthest point sampling of k of the n points [n, dim] (row-major device fp32),
 * starting from `start` - what dgl.geometry.farthest_point_sampler does for
 * GenBuffer.insert_history (:300-314).  idx_out [k] int32.  scratch: at least
 * hs_fps_scratch_bytes(n) bytes of device memory.  One cooperative launch (all SMs, a grid
 * barrier per selected point), asynchronous on `stream`. */
int64_t hs_fps_scratch_bytes(int64_t n);
int hs_fps(const float* points, int64_t n, int32_t dim, int32_t k, int32_t start, int32_t* idx_out,
           void* scratch, void* stream);

/* ---- the caller's side of a rollout: advantages -------------- */
/* compute_gae (omni_drones/learning/utils/gae.py:27-51) as MAPPOPolicy.train_op calls it
 * (omni_drones/learning/mappo.py:381-397): one backward scan over the T steps per
 * (env, agent) column, bit-identical to the reference's eager fp32 loop, optionally followed
 * by the batch-level advantage normalisation (adv - mean) / (std + 1e-8) of mappo.py:391-396.
 * reward / value / advantages / returns: fp32 element [e][t][a] at e*stride_env +
 * t*stride_step + a (the reference's [N,T,k] tensors have stride_env = T*k, stride_step = k;
 * the time-major rollout storage has stride_env = k, stride_step = N*k).  done: one byte per
 * (env, step) at e*done_stride_env + t*done_stride_step (the reference's [N,T,1] bool
 * broadcast over the agents).  next_value [N,k] contiguous.  gamma/lmbda are doubles because
 * the reference rounds the Python product gamma*lmbda to fp32 once. */
typedef struct hs_gae_params {
    int64_t num_envs;
    int32_t num_steps, num_agents;
    int64_t stride_env, stride_step;
    int64_t done_stride_env, done_stride_step;
    double gamma, lmbda;
    int32_t normalize;                  /* 1: advantages are normalised in place after the scan */
    int32_t reserved;
} hs_gae_params;
/* scratch: 16 bytes of device memory (two doubles: sum and sum of squares of the advantages;
 * zeroed by the call).  stats_out: NULL or 2 device floats that receive {mean, std} of the
 * un-normalised advantages (train_info["advantages_mean"/"advantages_std"]).  One launch
 * (two with normalize), asynchronous on `stream`. */
int hs_gae(const hs_gae_params* p, const float* reward, const uint8_t* done, const float* value,
           const float* next_value, float* advantages, float* returns, void* scratch,
           float* stats_out, void* stream);

/* ---- policy inference next to the tick ----------------------- */
/* The MAPPO actor / critic of the reference for this task: PartialAttentionEncoder
 * (omni_drones/learning/modules/networks.py:249-314; embed_dim = dim_feedforward = 128, one
 * head, query = the agent's own token, post-norm) over the observation
 * {state_self [1,D], state_others [n_others,3], cylinders [n_cyl,5]} of one agent, followed by
 * a linear head: DiagGaussian.fc_mean + log_std (modules/distributions.py:66-82, sampled as in
 * Actor.forward, omni_drones/learning/mappo.py:614-635) or Critic.v_out (mappo.py:652-668).
 * All pointers are the DEVICE pointers of the live nn.Module parameters (fp32, contiguous,
 * PyTorch layouts: Linear weight [out,in], MultiheadAttention in_proj_weight [3*128,128]). */
typedef struct hs_policy_weights {
    const float *embed_self_w, *embed_self_b;       /* split_embed.embed.state_self   [128,D], [128] */
    const float *embed_others_w, *embed_others_b;   /* split_embed.embed.state_others [128,3] (NULL when n_others == 0) */
    const float *embed_cyl_w, *embed_cyl_b;         /* split_embed.embed.cylinders    [128,5] (NULL when n_cyl == 0) */
    const float *embed_ln_w, *embed_ln_b;           /* split_embed.layer_norm */
    const float *attn_in_w, *attn_in_b;             /* attn.in_proj_weight / in_proj_bias */
    const float *attn_out_w, *attn_out_b;           /* attn.out_proj */
    const float *lin1_w, *lin1_b, *lin2_w, *lin2_b; /* linear1, linear2 */
    const float *norm1_w, *norm1_b, *norm2_w, *norm2_b;
    const float *head_w, *head_b;                   /* fc_mean [head_dim,128] | v_out [1,128] */
    const float *log_std;                           /* [head_dim] (actor) or NULL (critic) */
    int32_t self_dim, head_dim;                     /* D <= 128, head_dim <= 8 */
} hs_policy_weights;
/* Size (floats) of the prepared parameter blob for a given state_self width (K-major fp32 matrices
 * and vectors for the FFMA kernel, followed by the tf32 hi/lo weight images of the tcgen05 kernel). */
int64_t hs_policy_blob_floats(int32_t self_dim);
/* Packs the parameters K-major and folds Wk^T Wq and Wo Wv (see csrc/hs_policy.cuh); call it
 * after every optimiser step that touched the module.  One launch, asynchronous. */
int hs_policy_prepare(const hs_policy_weights* w, float* blob, void* stream);
/* Inputs and outputs of one forward pass over num_rows = E * A observation rows. */
typedef struct hs_policy_io {
    int64_t num_rows;
    int32_t n_others, n_cyl;        /* tokens besides the agent's own: n_others <= 2, n_cyl <= 4 */
    const float* state_self;        /* [num_rows, D]                                            */
    const float* state_others;      /* [num_rows, n_others, 3] (NULL when n_others == 0)        */
    const float* cylinders;         /* [num_rows, n_cyl, 5]    (NULL when n_cyl == 0)           */
    float* head_out;                /* [num_rows, head_dim]: action mean (actor) | state value (critic) */
    /* actor extras, each may be NULL */
    const float* eps;               /* [num_rows, head_dim] caller-supplied standard-normal noise */
    uint64_t* rng_state;            /* device {seed, step, 0, 0}: when eps is NULL and this is set, the kernel
                                       draws the noise itself (Philox4x32-10 keyed by seed, counter = (row, step),
                                       Box-Muller) and advances step by one per launch - the rollout needs no
                                       separate noise launch.  Both NULL: the mode (deterministic=True). */
    float* action;                  /* [num_rows, head_dim] = mean + exp(log_std) * noise       */
    float* logp;                    /* [num_rows] log-probability of that action                */
    float* eps_out;                 /* [num_rows, head_dim] the noise that was used             */
    float* feat_out;                /* [num_rows, 128] encoder features                         */
    int32_t impl;                   /* 0 = auto (tensor cores from 1024 rows on), 1 = fp32 FFMA kernel,
                                       2 = tcgen05 kernel (3xTF32: fp32-level results)          */
    int32_t reserved;
} hs_policy_io;
/* One launch.  Asynchronous on `stream`; io->action can be handed to hs_step_pre as the raw
 * action of the same tick. */
int hs_policy_forward(const float* blob, int32_t self_dim, int32_t head_dim, const hs_policy_io* io, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HS_B200_H */
