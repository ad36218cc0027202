"""ctypes binding of include/hs_b200.h.

There is deliberately no fallback: if ``libhs_b200.so`` is missing or an entry point is
absent, import raises; if no CUDA device is visible, ``hs_create`` fails with
HS_ERR_NO_DEVICE and :class:`HsError` is raised.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HS_B200_LIB: an instrumented build for the tools/ phase timers (never set in tests, smoke or bench)
LIB_PATH = os.environ.get("HS_B200_LIB") or os.path.join(PKG_DIR, "libhs_b200.so")

HS_ABI_VERSION = 3
HS_NUM_STATS = 24
HS_OPT_PREDICTOR_VARIANT = 1
HS_OPT_HOST_IO_GRAPH = 2
HS_OPT_HOST_IO_ZERO_COPY_ACTION = 3
HS_OPT_FUSED_TICK = 4
HS_OPT_EXACT_MATH = 5
HS_OPT_TICK_MAPPING = 6
HS_OPT_ROLLOUT_VARIANT = 7

# field ids, include/hs_b200.h
(FIELD_DRONE_POS, FIELD_DRONE_ROT, FIELD_DRONE_LINVEL, FIELD_DRONE_ANGVEL, FIELD_THROTTLE,
 FIELD_PID_INTEG, FIELD_PID_LAST_RATE, FIELD_TARGET_POS, FIELD_TARGET_VEL, FIELD_CYL_POS,
 FIELD_PROGRESS) = range(11)


class HsError(RuntimeError):
    pass


class hs_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("num_envs", C.c_int32), ("num_agents", C.c_int32),
        ("num_cylinders", C.c_int32), ("obs_max_cylinder", C.c_int32), ("future_step", C.c_int32),
        ("history_step", C.c_int32), ("max_episode_length", C.c_int32), ("use_tp_net", C.c_int32),
        ("smoothness_gated", C.c_int32), ("write_smoothness_coef_stat", C.c_int32),
        ("fixed_yaw", C.c_int32), ("ground_clamp", C.c_int32), ("use_obstacles", C.c_int32), ("contact_mode", C.c_int32), ("reserved_i", C.c_int32 * 1),
        ("dt", C.c_float),
        ("arena_size", C.c_float), ("max_height", C.c_float), ("cylinder_size", C.c_float),
        ("catch_radius", C.c_float), ("collision_radius", C.c_float),
        ("drone_detect_radius", C.c_float), ("target_detect_radius", C.c_float),
        ("v_drone", C.c_float), ("mask_value", C.c_float),
        ("dist_reward_coef", C.c_float), ("catch_reward_coef", C.c_float),
        ("detect_reward_coef", C.c_float), ("collision_coef", C.c_float), ("speed_coef", C.c_float),
        ("smoothness_coef", C.c_float),
        ("target_clip", C.c_float), ("max_thrust_ratio", C.c_float),
        ("pid_kp", C.c_float * 3), ("pid_ki", C.c_float * 3), ("pid_kd", C.c_float * 3),
        ("pid_ilimit", C.c_float * 3), ("pid_out_limit", C.c_float),
        ("kf", C.c_float), ("km", C.c_float), ("rotor_alpha", C.c_float),
        ("rotor_dirs", C.c_float * 4), ("rotor_x", C.c_float * 4), ("rotor_y", C.c_float * 4),
        ("drag_coef_times_mass", C.c_float),
        ("downwash_kr", C.c_float), ("downwash_kz", C.c_float),
        ("total_mass", C.c_float), ("inertia", C.c_float * 3), ("gravity", C.c_float),
        ("lin_damp_factor", C.c_float), ("ang_damp_factor", C.c_float),
        ("max_linear_velocity", C.c_float), ("max_angular_velocity", C.c_float),
        ("ground_z", C.c_float), ("hover_throttle", C.c_float),
        ("arena_size_sq", C.c_float), ("half_arena", C.c_float), ("coll_radius_x2", C.c_float),
        ("vmax_clamped", C.c_float), ("inv_inertia", C.c_float * 3), ("drone_radius", C.c_float), ("evader_radius", C.c_float),
        ("reserved_f", C.c_float * 1),
    ]


_BUF_FIELDS = [
    "arena", "stats", "state_self", "state_others", "obs_cylinders", "state_drones", "tp_input",
    "tp_input_prev", "tp_groundtruth", "tp_done", "reward", "done", "truncated", "drone_state",
    "prev_action", "rotor_cmds", "ctbr", "target_rate", "action_error", "v_prey", "smoothness_coef", "throttle_diff",
    "tp_ring", "tp_ring_pos",
]


class hs_buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _BUF_FIELDS]


class hs_gather_tensor(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("stride_env", C.c_int64), ("stride_step", C.c_int64),
                ("row_bytes", C.c_int32), ("reserved", C.c_int32)]


HS_GATHER_MAX_TENSORS = 24


class hs_tp_weights(C.Structure):
    _fields_ = [("weight_ih", C.c_void_p), ("weight_hh", C.c_void_p), ("bias_ih", C.c_void_p),
                ("bias_hh", C.c_void_p), ("fc_weight", C.c_void_p), ("fc_bias", C.c_void_p),
                ("input_size", C.c_int32), ("hidden_size", C.c_int32), ("output_size", C.c_int32),
                ("reserved", C.c_int32)]


class hs_reset_dist(C.Structure):
    """include/hs_b200.h::hs_reset_dist (device-side reset sampler)."""
    _fields_ = [("drone_lo", C.c_float * 2), ("drone_hi", C.c_float * 2),
                ("target_lo", C.c_float * 2), ("target_hi", C.c_float * 2),
                ("z_lo", C.c_float), ("z_hi", C.c_float),
                ("rpy_lo", C.c_float * 3), ("rpy_hi", C.c_float * 3),
                ("grid_size", C.c_float), ("num_grid", C.c_int32), ("boundary", C.c_float),
                ("cyl_z_active", C.c_float), ("cyl_z_inactive", C.c_float),
                ("min_cylinders", C.c_int32), ("fixed_num", C.c_int32), ("fixed_xy", C.c_int32),
                ("fixed_drone_xy", (C.c_float * 2) * 3), ("fixed_target_xy", C.c_float * 2),
                ("env_offset", C.c_int64), ("seed", C.c_uint64)]


class hs_host_io(C.Structure):
    """include/hs_b200.h::hs_host_io (host-buffer tick)."""
    _fields_ = [("action", C.c_void_p), ("state_self", C.c_void_p), ("state_others", C.c_void_p),
                ("obs_cylinders", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p)]


HS_HOVER_NUM_STATS, HS_HOVER_NUM_STATE = 39, 12


class hs_hover_params(C.Structure):
    """include/hs_b200.h::hs_hover_params."""
    _fields_ = [("reward_distance_scale", C.c_float), ("reward_v_scale", C.c_float), ("reward_acc_scale", C.c_float),
                ("reward_jerk_scale", C.c_float), ("linear_vel_max", C.c_float), ("linear_acc_max", C.c_float), ("alpha", C.c_float),
                ("target_pos", C.c_float * 3), ("time_encoding", C.c_int32), ("omega", C.c_int32), ("motor", C.c_int32),
                ("with_reward", C.c_int32)]


class hs_hover_io(C.Structure):
    _fields_ = [("observation", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p), ("stats", C.c_void_p),
                ("state", C.c_void_p), ("target_heading", C.c_void_p)]


class hs_host_batch(C.Structure):
    """include/hs_b200.h::hs_host_batch (hs_step_host_io_many)."""
    _fields_ = [("h", C.c_void_p), ("sets", C.POINTER(hs_buffers)), ("ios", C.POINTER(hs_host_io)), ("num_sets", C.c_int32),
                ("next_set", C.c_int32), ("staging_dev", C.c_void_p), ("stream", C.c_void_p)]


HS_OBS_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32)


class hs_gen_params(C.Structure):
    """include/hs_b200.h::hs_gen_params (HideAndSeek_envgen control plane)."""
    _fields_ = [("num_agents", C.c_int32), ("num_cylinders", C.c_int32), ("arena_size", C.c_float),
                ("grid_size", C.c_float), ("max_height", C.c_float), ("num_grid", C.c_int32),
                ("expand_cylinders", C.c_int32), ("expand_step", C.c_float), ("seed", C.c_uint64),
                ("task_offset", C.c_int64)]


_POLICY_PTRS = ["embed_self_w", "embed_self_b", "embed_others_w", "embed_others_b", "embed_cyl_w", "embed_cyl_b",
                "embed_ln_w", "embed_ln_b", "attn_in_w", "attn_in_b", "attn_out_w", "attn_out_b", "lin1_w", "lin1_b",
                "lin2_w", "lin2_b", "norm1_w", "norm1_b", "norm2_w", "norm2_b", "head_w", "head_b", "log_std"]


class hs_policy_weights(C.Structure):
    """include/hs_b200.h::hs_policy_weights (actor / critic parameters)."""
    _fields_ = [(n, C.c_void_p) for n in _POLICY_PTRS] + [("self_dim", C.c_int32), ("head_dim", C.c_int32)]


class hs_policy_io(C.Structure):
    """include/hs_b200.h::hs_policy_io."""
    _fields_ = [("num_rows", C.c_int64), ("n_others", C.c_int32), ("n_cyl", C.c_int32),
                ("state_self", C.c_void_p), ("state_others", C.c_void_p), ("cylinders", C.c_void_p),
                ("head_out", C.c_void_p), ("eps", C.c_void_p), ("rng_state", C.c_void_p), ("action", C.c_void_p),
                ("logp", C.c_void_p), ("eps_out", C.c_void_p), ("feat_out", C.c_void_p),
                ("impl", C.c_int32), ("reserved", C.c_int32)]


class hs_gae_params(C.Structure):
    """include/hs_b200.h::hs_gae_params (advantage scan over a rollout)."""
    _fields_ = [("num_envs", C.c_int64), ("num_steps", C.c_int32), ("num_agents", C.c_int32),
                ("stride_env", C.c_int64), ("stride_step", C.c_int64),
                ("done_stride_env", C.c_int64), ("done_stride_step", C.c_int64),
                ("gamma", C.c_double), ("lmbda", C.c_double), ("normalize", C.c_int32), ("reserved", C.c_int32)]


_EXPORTS = {
    "hs_abi_version": (C.c_int, []),
    "hs_last_error": (C.c_char_p, []),
    "hs_default_config": (C.c_int, [C.POINTER(hs_config), C.c_int32]),
    "hs_arena_floats": (C.c_int64, [C.POINTER(hs_config)]),
    "hs_create": (C.c_int, [C.POINTER(hs_config), C.POINTER(C.c_void_p)]),
    "hs_destroy": (C.c_int, [C.c_void_p]),
    "hs_bind_buffers": (C.c_int, [C.c_void_p, C.POINTER(hs_buffers)]),
    "hs_step_pre": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hs_step_post": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hs_step_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(hs_tp_weights), C.c_void_p, C.c_void_p]),
    "hs_gather_rows": (C.c_int, [C.POINTER(hs_gather_tensor), C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "hs_rollout_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.POINTER(hs_tp_weights), C.c_void_p, C.c_int64, C.c_void_p]),
    "hs_step_post_tp": (C.c_int, [C.c_void_p, C.POINTER(hs_tp_weights), C.c_void_p, C.c_void_p]),
    "hs_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hs_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hs_state_get": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hs_state_set": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hs_launch_count": (C.c_int64, [C.c_void_p]),
    "hs_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "hs_step_host_io": (C.c_int, [C.c_void_p, C.POINTER(hs_host_io), C.c_int, C.c_void_p, C.POINTER(hs_tp_weights),
                                  C.c_void_p, C.c_void_p]),
    "hs_step_host_io_async": (C.c_int, [C.c_void_p, C.POINTER(hs_host_io), C.c_int, C.c_void_p, C.POINTER(hs_tp_weights),
                                        C.c_void_p, C.c_void_p]),
    "hs_host_io_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hs_hover_post": (C.c_int, [C.c_void_p, C.POINTER(hs_hover_params), C.POINTER(hs_hover_io), C.c_void_p]),
    "hs_step_host_io_many": (C.c_int, [C.POINTER(hs_host_batch), C.c_int32, C.c_int32, C.c_int32, C.c_int, C.POINTER(hs_tp_weights),
                                       C.c_void_p, C.c_void_p]),
    "hs_gen_sample_nearby": (C.c_int, [C.POINTER(hs_gen_params), C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "hs_fps_scratch_bytes": (C.c_int64, [C.c_int64]),
    "hs_fps": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hs_policy_blob_floats": (C.c_int64, [C.c_int32]),
    "hs_policy_prepare": (C.c_int, [C.POINTER(hs_policy_weights), C.c_void_p, C.c_void_p]),
    "hs_policy_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(hs_policy_io), C.c_void_p]),
    "hs_gae": (C.c_int, [C.POINTER(hs_gae_params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "hs_sample_reset": (C.c_int, [C.c_void_p, C.POINTER(hs_reset_dist), C.c_uint64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_EXPORTS)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  There is no CPU/eager fallback for the environment step.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _EXPORTS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.hs_abi_version() != HS_ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.hs_abi_version()} != {HS_ABI_VERSION}; rebuild it")
    return lib


lib = _load()


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.hs_last_error()
        raise HsError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")


def default_config(num_envs: int) -> hs_config:
    cfg = hs_config()
    check(lib.hs_default_config(C.byref(cfg), int(num_envs)), "hs_default_config")
    return cfg
