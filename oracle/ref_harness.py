"""Runs the reference's OWN source for the HideAndSeek step on CPU (build container only).

TEST INFRASTRUCTURE.  This module needs /root/reference and therefore only runs in
the build container; it is used by ``oracle/gen_golden.py`` to (a) validate the
restatement in ``oracle/hs_oracle.py`` and (b) write the fixtures under
``tests/golden/``.  Nothing here is copied from the reference: the reference files
are parsed with ``ast`` where they lie and the wanted function bodies are exec'd
against a fake ``self`` (recipe: SURVEY.md Appendix C).

The one piece the reference does not contain -- the PhysX step -- is supplied by
``hs_oracle.rigid_body_step`` (parity unpinned, see that module's header).
"""
from __future__ import annotations

import ast
import collections
import importlib.util
import math
import sys
import textwrap
import types
from pathlib import Path

import numpy as np
import torch
import yaml

from . import hs_oracle as O

# The reference tree: where it lies in the build container, or the copy `pip install --target baseline/_ref` made of its
# Python package (git-ignored, travels to the GPU box; see __graft_entry__.build and DESIGN.md).  Nothing is ever copied
# into the tracked tree.
_REPO = Path(__file__).resolve().parent.parent
REF = next((p for p in (Path("/root/reference"), _REPO / "baseline" / "_ref") if (p / "omni_drones" / "envs").is_dir()),
           Path("/root/reference"))


def available() -> bool:
    return (REF / "omni_drones" / "envs" / "hide_and_seek" / "hideandseek.py").is_file()


def _drone_yaml():
    """crazyflie.yaml of the reference (the pip-installed copy carries only .py files: then the package's own copy of the
    vehicle parameters, which is configuration data)."""
    p = REF / "omni_drones/robots/assets/usd/crazyflie.yaml"
    if not p.is_file():
        p = _REPO / "multi-uav-pursuit-evasion_b200" / "assets" / "crazyflie.yaml"
    return yaml.safe_load(p.read_text())


# ----------------------------------------------------------------------------
# a dict that behaves enough like tensordict.TensorDict for the extracted methods
# ----------------------------------------------------------------------------
class TD(dict):
    def __init__(self, source=None, batch_size=None, device=None):
        super().__init__()
        self.batch_size = list(batch_size) if batch_size is not None else []
        for k, v in (source or {}).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(key, tuple) and all(isinstance(k, str) for k in key):
            if len(key) == 1:
                return self.__setitem__(key[0], value)
            if key[0] not in self:
                dict.__setitem__(self, key[0], TD({}, self.batch_size))
            self[key[0]][key[1:]] = value
            return
        if isinstance(key, str):
            if isinstance(value, dict) and not isinstance(value, TD):
                value = TD(value, self.batch_size)
            dict.__setitem__(self, key, value)
            return
        # index assignment: td[env_ids] = scalar
        for v in self.values():
            if isinstance(v, TD):
                v[key] = value
            else:
                v[key] = value

    def __getitem__(self, key):
        if isinstance(key, tuple) and all(isinstance(k, str) for k in key):
            out = self
            for k in key:
                out = dict.__getitem__(out, k)
            return out
        return dict.__getitem__(self, key)

    def set(self, key, value):
        self[key] = value
        return self

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def update(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and k in self and isinstance(self[k], TD):
                self[k].update(v)
            else:
                self[k] = v
        return self

    def clone(self):
        return TD({k: v.clone() for k, v in self.items()}, self.batch_size)

    def flat(self, prefix=()):
        out = {}
        for k, v in self.items():
            if isinstance(v, TD):
                out.update(v.flat(prefix + (k,)))
            else:
                out[prefix + (k,)] = v
        return out


# ----------------------------------------------------------------------------
# module loading / AST extraction
# ----------------------------------------------------------------------------
def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, REF / relpath)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_LOADED = {}


def load_reference():
    """Path-loads the importable reference modules and extracts the hot methods."""
    if _LOADED:
        return _LOADED
    _stub("omni_drones")
    _stub("omni_drones.utils")
    _stub("omni_drones.controllers")
    _stub("omni_drones.actuators")
    _stub("tensordict", TensorDict=TD)
    ut = _load("omni_drones.utils.torch", "omni_drones/utils/torch.py")
    rg = _load("omni_drones.actuators.rotor_group", "omni_drones/actuators/rotor_group.py")
    lc = _load("omni_drones.controllers.lee_position_controller",
               "omni_drones/controllers/lee_position_controller.py")

    ns = dict(torch=torch, np=np, math=math, collections=collections, vmap=torch.vmap,
              TensorDict=TD, TensorDictBase=TD, D=torch.distributions,
              cpos=ut.cpos, off_diag=ut.off_diag, quat_axis=ut.quat_axis, others=ut.others,
              quat_rotate=ut.quat_rotate, quat_rotate_inverse=ut.quat_rotate_inverse,
              normalize=ut.normalize, euler_to_quaternion=ut.euler_to_quaternion,
              symlog=ut.symlog, Optional=None)

    def extract(relpath, names, cls=None, strip_decorators=True):
        src = (REF / relpath).read_text()
        tree = ast.parse(src)
        body = tree.body
        if cls is not None:
            body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
        out = {}
        for node in body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                if strip_decorators:
                    node.decorator_list = []
                mod = ast.Module(body=[node], type_ignores=[])
                code = compile(ast.fix_missing_locations(mod), f"<ref:{relpath}:{node.name}>", "exec")
                local = {}
                exec(code, ns, local)
                out[node.name] = local[node.name]
        missing = set(names) - set(out)
        assert not missing, f"not found in {relpath}: {missing}"
        return out

    hs = "omni_drones/envs/hide_and_seek/hideandseek.py"
    free = extract(hs, ["is_perpendicular_line_intersecting_segment", "is_line_blocked_by_cylinder",
                        "select_unoccupied_positions", "grid_to_continuous", "continuous_to_grid",
                        "set_outside_circle_to_one"])
    ns.update(free)
    env_m = extract(hs, ["_pre_sim_step", "_compute_state_and_obs", "_compute_reward_and_done",
                         "_get_dummy_policy_prey", "_reset_idx", "rejection_sampling_random_cylinder"],
                    cls="HideAndSeek")
    mr = "omni_drones/robots/drone/multirotor.py"
    ns.update(extract(mr, ["separation"]))
    drone_m = extract(mr, ["apply_action", "get_state", "_reset_idx", "downwash"], cls="MultirotorBase")
    tr = extract("omni_drones/utils/torchrl/transforms.py", ["_inv_call"], cls="PIDRateController")
    ie = extract("omni_drones/envs/isaac_env.py", ["_reset", "_step", "get_env_poses"], cls="IsaacEnv")
    hover_m = extract("omni_drones/envs/single/hover.py",
                      ["_reset_idx", "_pre_sim_step", "_compute_state_and_obs", "_compute_reward_and_done"], cls="Hover")
    tp = ast.parse((REF / "omni_drones/learning/mappo.py").read_text())
    node = next(n for n in tp.body if isinstance(n, ast.ClassDef) and n.name == "TP_net")
    tp_ns = dict(torch=torch, nn=torch.nn)
    exec(compile(ast.Module(body=[node], type_ignores=[]), "<ref:TP_net>", "exec"), tp_ns)
    _LOADED.update(ut=ut, rg=rg, lc=lc, env=env_m, drone=drone_m, transform=tr, isaac=ie, hover=hover_m,
                   TP_net=tp_ns["TP_net"], ns=ns, extract=extract)
    return _LOADED


def load_envgen():
    """HideAndSeek_envgen's own source (omni_drones/envs/hide_and_seek/hideandseek_envgen.py): the env methods, the free
    function `sanity_check` and the `GenBuffer` class.  `dgl.geometry.farthest_point_sampler` (dgl is not in this image) is
    replaced by the numpy restatement of farthest point sampling in oracle/envgen_oracle.py with start index 0 (dgl draws a
    random start): the archive bookkeeping around it is the reference's."""
    R = load_reference()
    if "envgen" in R:
        return R
    import copy as _copy
    from . import envgen_oracle as EO
    eg = "omni_drones/envs/hide_and_seek/hideandseek_envgen.py"
    ns, extract = R["ns"], R["extract"]

    def farthest_point_sampler(points, npoints, start_idx=None):
        pts = points[0].numpy() if isinstance(points, torch.Tensor) else np.asarray(points)[0]
        return torch.from_numpy(EO.fps(pts.astype(np.float32), int(npoints), start=0).astype(np.int64)).unsqueeze(0)
    ns.update(copy=_copy, deque=collections.deque, farthest_point_sampler=farthest_point_sampler)
    ns.update(extract(eg, ["sanity_check"]))
    tree = ast.parse((REF / eg).read_text())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GenBuffer")
    exec(compile(ast.Module(body=[node], type_ignores=[]), "<ref:GenBuffer>", "exec"), ns)
    env_m = extract(eg, ["_pre_sim_step", "_compute_state_and_obs", "_compute_reward_and_done", "_get_dummy_policy_prey",
                         "_reset_idx", "uniform_sampling", "rejection_sampling_random_cylinder"], cls="HideAndSeek_envgen")
    R.update(envgen=env_m, GenBuffer=ns["GenBuffer"])
    return R


# ----------------------------------------------------------------------------
# fakes for the simulator views
# ----------------------------------------------------------------------------
class _View:
    """pos/rot/vel accessors over tensors held in a dict (stands in for omni_drones/views/*)."""

    def __init__(self, store, pos, rot=None, vel=None):
        self.s, self.kp, self.kr, self.kv = store, pos, rot, vel

    def get_world_poses(self, clone=False):
        p = self.s[self.kp]
        r = self.s[self.kr] if self.kr else torch.zeros(*p.shape[:-1], 4)
        return (p.clone(), r.clone()) if clone else (p, r)

    def set_world_poses(self, positions=None, orientations=None, env_indices=None):
        if positions is not None:
            self.s[self.kp][env_indices] = positions
        if orientations is not None and self.kr:
            self.s[self.kr][env_indices] = orientations

    def get_velocities(self, clone=False):
        v = self.s[self.kv]
        return v.clone() if clone else v

    def set_velocities(self, velocities, env_indices=None):
        full = self.s[self.kv]
        if env_indices is None:
            full[:] = velocities
        elif velocities.shape[0] == full.shape[0]:
            full[env_indices] = velocities[env_indices]
        else:
            full[env_indices] = velocities


class _Wrench:
    def __init__(self):
        self.forces = None
        self.torques = None

    def apply_forces_and_torques_at_pos(self, forces=None, torques=None, positions=None, is_global=True):
        self.forces = None if forces is None else forces.clone()
        self.torques = None if torques is None else torques.clone()


class _Obj:
    pass


class RefEnv:
    """The reference's HideAndSeek step, driven from its extracted source."""

    def __init__(self, P: O.HSParams, E: int, use_random_cylinder=True, scenario_flag="empty",
                 min_cylinders=4, tp_state_dict=None):
        R = load_reference()
        self.R, self.P, self.E = R, P, E
        A, C = P.num_agents, P.num_cylinders
        dev = torch.device("cpu")
        store = dict(dpos=torch.zeros(E, A, 3), drot=torch.zeros(E, A, 4), dvel=torch.zeros(E, A, 6),
                     tpos=torch.zeros(E, 1, 3), tvel=torch.zeros(E, 1, 6), cpos=torch.zeros(E, C, 3))
        store["drot"][..., 0] = 1.0
        store["cpos"][..., 2] = -20.0
        self.store = store

        params = _drone_yaml()
        # ---- drone (MultirotorBase stand-in, attributes as multirotor.py:159-263 sets them)
        d = _Obj()
        d.shape, d.n, d.num_rotors, d.device, d.dt = (E, A), A, 4, dev, P.dt
        d.params, d.is_articulation, d.rotor_joint_indices = params, True, None
        d.use_force_sensor, d.randomization = False, {}
        rotors_mod = R["rg"].RotorGroup(params["rotor_configuration"], dt=P.dt)
        d.rotors_module = rotors_mod
        d.rotors = lambda cmds, prm: torch.func.functional_call(rotors_mod, prm, (cmds,))
        d.rotors.f_inv = torch.sqrt
        d.rotor_params = {k: v.detach().expand(E, A, *v.shape).clone() for k, v in rotors_mod.named_parameters()}
        d.throttle = d.rotor_params["throttle"]
        d.directions, d.KF, d.KM = d.rotor_params["directions"], d.rotor_params["KF"], d.rotor_params["KM"]
        d.MAX_ROT_VEL = torch.as_tensor(params["rotor_configuration"]["max_rotation_velocities"]).float()
        d.thrusts, d.torques, d.forces = torch.zeros(E, A, 4, 3), torch.zeros(E, A, 3), torch.zeros(E, A, 3)
        d.pos, d.rot = store["dpos"].clone(), store["drot"].clone()
        d.throttle_difference = torch.zeros(E, A)
        d.heading, d.up = torch.zeros(E, A, 3), torch.zeros(E, A, 3)
        d.vel = d.vel_w = torch.zeros(E, A, 6)
        d.vel_b = torch.zeros(E, A, 6)
        d.acc = d.acc_w = torch.zeros(E, A, 6)
        d.jerk = torch.zeros(E, A, 6)
        d.masses = torch.ones(E, A, 1) * torch.tensor(params["mass"])
        d.gravity = torch.ones(E, A, 1) * torch.tensor(P.total_mass) * 9.81       # sum of body masses
        d.drag_coef = torch.zeros(E, A, 1) * torch.tensor(params["drag_coef"])
        d.rotor_pos_offset = torch.zeros(E, A, 4, 3)
        d._envs_positions = torch.zeros(E, 1, 3)
        dv = _View(store, "dpos", "drot", "dvel")
        d.get_world_poses, d.set_world_poses = dv.get_world_poses, dv.set_world_poses
        d.get_velocities, d.set_velocities = dv.get_velocities, dv.set_velocities
        d.rotors_view = _Obj()
        d.rotors_view.get_world_poses = lambda clone=False: (
            store["dpos"].unsqueeze(2).expand(E, A, 4, 3), store["drot"].unsqueeze(2).expand(E, A, 4, 4))
        self.rotor_wrench, self.base_wrench = _Wrench(), _Wrench()
        d.rotors_view.apply_forces_and_torques_at_pos = self.rotor_wrench.apply_forces_and_torques_at_pos
        d.base_link = self.base_wrench
        d.downwash = R["drone"]["downwash"]
        for name in ("apply_action", "get_state", "_reset_idx"):
            setattr(d, name, types.MethodType(R["drone"][name], d))
        d.rotors_f_inv = torch.sqrt
        # MultirotorBase._reset_idx uses self.rotors.f_inv
        self.drone = d

        # ---- env (HideAndSeek stand-in; attribute list: SURVEY.md Appendix C item 5)
        e = _Obj()
        e.num_envs, e.num_agents, e.num_cylinders, e.device, e.batch_size = E, A, C, dev, [E]
        e.drone = d
        e.target = _View(store, "tpos", None, "tvel")
        e.cylinders = _View(store, "cpos")
        e.envs_positions = torch.zeros(E, 3)
        e.cylinder_height, e.cylinder_size = P.max_height, P.cylinder_size
        e.obs_max_cylinder, e.mask_value = P.obs_max_cylinder, P.mask_value
        e.drone_detect_radius, e.target_detect_radius = P.drone_detect_radius, P.target_detect_radius
        e.progress_buf = torch.zeros(E)
        e.max_episode_length = P.max_episode_length
        e.use_TP_net, e.use_obstacles = P.use_tp_net, int(P.use_obstacles)
        e.history_step = P.history_step
        e.history_data = collections.deque(maxlen=P.history_step)
        e.future_predcition_step, e.arena_size, e.max_height = P.future_step, P.arena_size, P.max_height
        e.time_encoding_dim = 4
        e._should_render = lambda substep: False
        e.use_eval = 0
        e.catch_radius, e.collision_radius = P.catch_radius, P.collision_radius
        e.dist_reward_coef, e.detect_reward_coef = P.dist_reward_coef, P.detect_reward_coef
        e.catch_reward_coef, e.speed_coef, e.collision_coef = P.catch_reward_coef, P.speed_coef, P.collision_coef
        e.init_smoothness_coef, e.smooth_lr, e.update_epoch, e.max_smoothness_coef = P.smoothness_coef, P.smooth_lr, 0, P.max_smoothness_coef
        e.use_deployment = P.use_deployment
        e.cfg = _Obj()
        e.cfg.task = _Obj()
        e.cfg.task.v_drone = P.v_drone
        e.v_prey = P.v_prey
        e.env_ids = torch.arange(E)
        e.stats = TD({k: torch.zeros(E, 1) for k in O.STAT_KEYS}, [E])
        e.info = TD({"drone_state": torch.zeros(E, A, 13), "prev_action": torch.zeros(E, A, 4)}, [E])
        e.prev_actions = torch.zeros(E, A, 4)
        e.use_random_cylinder, e.scenario_flag = use_random_cylinder, scenario_flag
        e.max_cylinders, e.min_cylinders = C, min_cylinders
        e.use_fixed_num, e.fixed_num = False, None
        e.invalid_z, e.boundary = -20.0, P.arena_size - 0.1
        a = P.arena_size / math.sqrt(2.0)
        U = torch.distributions.Uniform
        e.init_drone_pos_dist = U(torch.tensor([0.1, -a + 0.1]), torch.tensor([a - 0.1, a - 0.1]))
        e.init_target_pos_dist = U(torch.tensor([-a + 0.1, -a + 0.1]), torch.tensor([-0.1, a - 0.1]))
        e.init_drone_pos_dist_z = U(torch.tensor([P.max_height / 2 - 0.1]), torch.tensor([P.max_height / 2 + 0.1]))
        e.init_target_pos_dist_z = U(torch.tensor([P.max_height / 2 - 0.1]), torch.tensor([P.max_height / 2 + 0.1]))
        e.init_rpy_dist = U(torch.tensor([-0.2, -0.2, 0.0]) * torch.pi, torch.tensor([0.2, 0.2, 0.2]) * torch.pi)
        e.active_cylinders = torch.zeros(E, 1)
        if P.use_tp_net:
            e.TP = R["TP_net"](input_dim=P.tp_frame_dim, output_dim=3 * P.future_step,
                               future_predcition_step=P.future_step, window_step=1)
            if tp_state_dict is not None:
                e.TP.load_state_dict(tp_state_dict)
            e.TP.requires_grad_(False)
        e.sim = _Obj()
        e.sim.step = self._sim_step
        e.sim._physics_sim_view = _Obj()
        e.sim._physics_sim_view.flush = lambda: None
        e._post_sim_step = lambda td: None
        for name, fn in R["env"].items():
            setattr(e, name, types.MethodType(fn, e))
        for name, fn in R["isaac"].items():
            setattr(e, name, types.MethodType(fn, e))
        self.env = e

        # ---- the PIDrate transform (transforms.py:404-459)
        t = _Obj()
        t.controller = R["lc"].PIDRateController(P.dt, 9.81, params)
        t.controller.requires_grad_(False)
        t.action_key = ("agents", "action")
        t.target_clip, t.max_thrust_ratio, t.fixed_yaw = params["target_clip"], params["max_thrust_ratio"], params["fixed_yaw"]
        t._inv_call = types.MethodType(R["transform"]["_inv_call"], t)
        self.transform = t
        self.td = None

    # PhysX stand-in: consumes the wrench recorded by apply_action, then clears it
    def _sim_step(self, render=False):
        P, s = self.P, self.store
        q = s["drot"]
        if self.rotor_wrench.forces is not None:
            thrusts = self.rotor_wrench.forces.reshape(self.E, P.num_agents, 4, 3)[..., 2]
            tau_w = self.base_wrench.torques.reshape(self.E, P.num_agents, 3)
            yaw = O.quat_apply_inverse(q, tau_w)[..., 2]
            ext = self.base_wrench.forces.reshape(self.E, P.num_agents, 3)
        else:
            thrusts = yaw = ext = None
        p, q2, v, w = O.rigid_body_step(P, s["dpos"], q, s["dvel"][..., :3], s["dvel"][..., 3:], thrusts, yaw, ext)
        s["dpos"][:], s["drot"][:] = p, q2
        s["dvel"][..., :3], s["dvel"][..., 3:] = v, w
        s["tpos"][:] = s["tpos"] + P.dt * s["tvel"][..., :3]
        self.rotor_wrench.forces = self.base_wrench.forces = self.base_wrench.torques = None

    # -- driving ---------------------------------------------------------------
    def reset_with(self, mask, init):
        """Reset through IsaacEnv._reset, but with the initial poses injected instead of sampled."""
        e = self.env
        orig = e._reset_idx

        def injected(env_ids):
            # run the reference _reset_idx for its bookkeeping, then overwrite the sampled poses
            orig(env_ids)
        # Patch the pose setters so that the *sampled* poses are replaced by `init`
        d = self.drone
        real_set = d.set_world_poses
        real_tset = e.target.set_world_poses
        real_cset = e.cylinders.set_world_poses

        def dset(positions=None, orientations=None, env_indices=None):
            real_set(init["drone_pos"][env_indices], init["drone_rot"][env_indices], env_indices)

        def tset(positions=None, orientations=None, env_indices=None):
            real_tset(positions=init["target_pos"][env_indices].unsqueeze(1), env_indices=env_indices)

        def cset(positions=None, orientations=None, env_indices=None):
            real_cset(positions=init["cyl_pos"][env_indices], env_indices=env_indices)

        d.set_world_poses, e.target.set_world_poses, e.cylinders.set_world_poses = dset, tset, cset
        if not e.use_random_cylinder:
            # fixed scenarios never call cylinders.set_world_poses in _reset_idx
            ids = mask.nonzero().squeeze(-1)
            real_cset(positions=init["cyl_pos"][ids], env_indices=ids)
            e.active_cylinders = (init["cyl_pos"][..., 2] > 0).sum(-1, keepdim=True).float()
        try:
            td_in = TD({"_reset": mask.clone()}, [self.E])
            out = e._reset(td_in)
        finally:
            d.set_world_poses, e.target.set_world_poses, e.cylinders.set_world_poses = real_set, real_tset, real_cset
        self.td = out
        return out

    def step(self, raw_action, done_prev):
        e = self.env
        td = TD({"agents": {"action": raw_action.clone()},
                 "info": {"drone_state": e.info["drone_state"].clone(), "prev_action": e.info["prev_action"].clone()},
                 "stats": TD({}, [self.E]),
                 "done": done_prev.reshape(self.E, 1).clone()}, [self.E])
        td = self.transform._inv_call(td)
        aux = dict(cmds=td[("agents", "action")].clone(), ctbr=td["ctbr"].clone(),
                   target_rate=td["target_rate"].clone(),
                   action_error=td[("stats", "action_error_order1")].clone())
        out = e._step(td)
        return out["next"], aux


class RefEnvgen(RefEnv):
    """The reference's HideAndSeek_envgen driven from its own source: reset with the particle generator
    (hideandseek_envgen.py:875-1013 - uniform / archive split, GenBuffer.samplenearby, insert), the tick, and
    `_compute_reward_and_done` with the archive bookkeeping of an episode end (:1241-1333).  Only PhysX is ours."""

    EXTRA = ("success_buffer", "success_unif", "history_buffer", "add_history", "ratio_unif")

    def __init__(self, P: O.HSParams, E: int, eval_iter=2, ratio_unif=0.3, R_min=0.5, R_max=0.9, success_threshold=0.98,
                 expand_cylinders=True, expand_step=0.05, buffer_length=5000, min_cylinders=4, tp_state_dict=None):
        super().__init__(P, E, use_random_cylinder=True, scenario_flag="empty", min_cylinders=min_cylinders,
                         tp_state_dict=tp_state_dict)
        R = load_envgen()
        e = self.env
        e.smoothness_coef = P.smoothness_coef
        e.use_particle_generator, e.update_iter, e.eval_iter = True, 0, eval_iter
        e.ratio_unif, e.R_min, e.R_max, e.success_threshold = ratio_unif, R_min, R_max, success_threshold
        e.expand_cylinders, e.expand_step = expand_cylinders, expand_step
        e.num_unif = E
        e.gen_buffer = R["GenBuffer"](P.num_agents, P.num_cylinders, e.device)
        e.gen_buffer.buffer_length = buffer_length
        # GenBuffer.insert_weights keeps `weights.to('cpu').numpy()`: on the reference's CUDA device that is a copy; on this
        # CPU harness it would alias stats["success"], which the next reset zeroes in place - hand it a copy as CUDA does
        _iw = e.gen_buffer.insert_weights
        e.gen_buffer.insert_weights = lambda w: _iw(w.clone())
        keys = [k for k in O.STAT_KEYS if k != "smoothness_coef"] + list(self.EXTRA)
        for i in range(P.num_cylinders + 1):
            keys += [f"ratio_cylinders_{i}", f"success_cylinders_{i}"]
        self.stat_keys = keys
        e.stats = TD({k: torch.zeros(E, 1) for k in keys}, [E])
        for name, fn in R["envgen"].items():
            setattr(e, name, types.MethodType(fn, e))

    def reset_all(self):
        """IsaacEnv._reset of every env with the reference's own sampling (torch / numpy global RNG state)."""
        self.td = self.env._reset(TD({"_reset": torch.ones(self.E, dtype=torch.bool)}, [self.E]))
        return self.td


class RefHover:
    """The reference's Hover task (omni_drones/envs/single/hover.py:296-523) driven from its extracted source: `_reset_idx`,
    `_pre_sim_step`, `_compute_state_and_obs`, `_compute_reward_and_done` bound to a fake self, the drone and the PIDrate
    transform exactly as in RefEnv (one pursuer, no evader, no cylinders), hs_oracle.rigid_body_step for PhysX."""

    STAT_KEYS = ("return", "pos_bonus", "head_bonus", "reward_pos", "reward_up", "reward_vel", "reward_acc", "reward_jerk",
                 "episode_len", "pos_error", "heading_alignment", "uprightness", "action_smoothness",
                 "linear_v_max", "angular_v_max", "linear_a_max", "angular_a_max", "linear_jerk_max", "angular_jerk_max",
                 "linear_v_mean", "angular_v_mean", "linear_a_mean", "angular_a_mean", "linear_jerk_mean", "angular_jerk_mean",
                 "motor1", "motor2", "motor3", "motor4", "cmd_r", "cmd_p", "cmd_y", "cmd_thrust",
                 "target_r_rate", "target_p_rate", "target_y_rate", "real_r_rate", "real_p_rate", "real_y_rate")

    def __init__(self, E: int, max_episode_length: int = 500, task=None):
        P = O.HSParams(num_agents=1, num_cylinders=0, obs_max_cylinder=0, use_tp_net=False,
                       max_episode_length=max_episode_length, max_linear_velocity=1000.0)
        # reuse RefEnv for the drone / transform / PhysX stand-in, then swap the task methods
        self.base = RefEnv(P, E, use_random_cylinder=False, scenario_flag="empty")
        self.P, self.E = P, E
        R, d, store = self.base.R, self.base.drone, self.base.store
        store["tpos"][:] = torch.tensor([50.0, 50.0, 0.5])          # the unused evader slot, far away
        d.acc, d.jerk = torch.zeros(E, 1, 6), torch.zeros(E, 1, 6)
        d.intrinsics = TD({"mass": torch.zeros(E, 1, 1)}, [E, 1])
        t = dict(reward_action_smoothness_weight=0.0, reward_distance_scale=10.0, reward_v_scale=0.0, reward_acc_scale=0.0,
                 reward_jerk_scale=0.0, linear_vel_max=3.0, linear_acc_max=10.0, omega=False, motor=False, time_encoding=True,
                 add_noise=False, action_noise=False, latency=False)
        t.update(task or {})
        e = _Obj()
        e.num_envs, e.device, e.batch_size, e.drone, e.training = E, torch.device("cpu"), [E], d, True
        e.cfg = _Obj()
        e.cfg.task = _Obj()
        for k, v in t.items():
            setattr(e.cfg.task, k, v)
        e.reward_distance_scale, e.reward_v_scale = t["reward_distance_scale"], t["reward_v_scale"]
        e.reward_acc_scale, e.reward_jerk_scale = t["reward_acc_scale"], t["reward_jerk_scale"]
        e.linear_vel_max, e.linear_acc_max = t["linear_vel_max"], t["linear_acc_max"]
        e.time_encoding, e.time_encoding_dim, e.latency, e.has_payload = t["time_encoding"], 4, t["latency"], False
        e.dt, e.max_episode_length, e.alpha = P.dt, max_episode_length, 0.8
        e.envs_positions = torch.zeros(E, 3)
        e.progress_buf = torch.zeros(E)
        U = torch.distributions.Uniform
        e.init_pos_dist = U(torch.tensor([-1., -1., 0.05]), torch.tensor([1., 1., 2.0]))
        e.init_rpy_dist = U(torch.tensor([-0.2, -0.2, 0.0]) * torch.pi, torch.tensor([0.2, 0.2, 0.5]) * torch.pi)
        # (degenerate Uniform(0, 0) like the reference's: modern torch validates low < high, the old pinned one did not)
        e.target_rpy_dist = U(torch.tensor([0., 0., 0.]) * torch.pi, torch.tensor([0., 0., 0.]) * torch.pi, validate_args=False)
        e.target_pos = torch.tensor([[0.0, 0.0, 1.0]])
        e.target_heading = torch.zeros(E, 1, 3)
        e.target_vis = _Obj()
        e.target_vis.set_world_poses = lambda orientations=None, env_indices=None: None
        e.init_vels = torch.zeros(E, 1, 6)
        for k in ("last_linear_v", "last_angular_v", "last_linear_a", "last_angular_a", "last_linear_jerk", "last_angular_jerk"):
            setattr(e, k, torch.zeros(E, 1))
        e.stats = TD({k: torch.zeros(E, 1) for k in self.STAT_KEYS}, [E])
        e.info = TD({"drone_state": torch.zeros(E, 1, 13), "prev_action": torch.zeros(E, 1, 4)}, [E])
        for name, fn in R["hover"].items():
            setattr(e, name, types.MethodType(fn, e))
        self.env = e

    def reset_with(self, mask, pos, rot):
        """hover.py:296-332 with the sampled pose replaced by (pos [E,1,3], rot [E,1,4]); then the observation half that
        IsaacEnv._reset runs (isaac_env.py:217-224; Hover's _reset_idx does not step the simulator)."""
        e, d = self.env, self.base.drone
        env_ids = mask.nonzero().squeeze(-1)
        real_set = d.set_world_poses
        d.set_world_poses = lambda positions=None, orientations=None, env_indices=None: real_set(pos[env_indices], rot[env_indices], env_indices)
        try:
            e._reset_idx(env_ids)
        finally:
            d.set_world_poses = real_set
        e.progress_buf[env_ids] = 0.0
        return e._compute_state_and_obs()

    def step(self, raw_action, done_prev):
        e, b = self.env, self.base
        td = TD({"agents": {"action": raw_action.clone()},
                 "info": {"drone_state": e.info["drone_state"].clone(), "prev_action": e.info["prev_action"].clone()},
                 "stats": TD({}, [self.E]), "done": done_prev.reshape(self.E, 1).clone()}, [self.E])
        td = b.transform._inv_call(td)
        e.info["prev_action"] = td[("info", "prev_action")]
        e._pre_sim_step(td)
        b._sim_step()
        e.progress_buf += 1
        obs = e._compute_state_and_obs()
        rd = e._compute_reward_and_done()
        aux = dict(cmds=td[("agents", "action")].clone(), ctbr=td["ctbr"].clone(), target_rate=td["target_rate"].clone())
        return obs, rd, aux

    def stats_matrix(self):
        return torch.cat([self.env.stats[k].reshape(self.E, 1) for k in self.STAT_KEYS], dim=-1).detach().clone().float()
