"""CPU oracle for the HideAndSeek vectorised environment step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or as the timed CPU arm -- never as a fallback for the CUDA path.

What it is: a restatement, in plain fp32 torch-CPU tensor arithmetic, of the
reference's algorithm for one control tick of the 3-v-1 pursuit-evasion task
(all paths relative to /root/reference):

  * CTBR action transform          omni_drones/utils/torchrl/transforms.py:425-459
  * body-rate PID                  omni_drones/controllers/lee_position_controller.py:476-550
  * first-order rotor model        omni_drones/actuators/rotor_group.py:55-71
  * wrench assembly + downwash     omni_drones/robots/drone/multirotor.py:466-508, 724-753
  * potential-field evader         omni_drones/envs/hide_and_seek/hideandseek.py:1067-1141, 725-744
  * line-of-sight test             omni_drones/envs/hide_and_seek/hideandseek.py:47-103
  * observation / TP frames        omni_drones/envs/hide_and_seek/hideandseek.py:746-917
  * reward / done / stats          omni_drones/envs/hide_and_seek/hideandseek.py:919-1065
  * reset bookkeeping              omni_drones/envs/hide_and_seek/hideandseek.py:609-723,
                                   omni_drones/robots/drone/multirotor.py:635-650
  * step / reset sequencing        omni_drones/envs/isaac_env.py:210-240

PARITY STATUS
  * Everything listed above is pinned: ``oracle/gen_golden.py`` executes the
    reference's own source (AST-extracted, run on CPU in the build container)
    on seeded inputs, checks this restatement against it and writes the
    fixtures in ``tests/golden/``.
  * The rigid-body integrator (`rigid_body_step`) is PARITY UNPINNED: in the
    reference it is closed-source PhysX inside Isaac Sim 2022.2.0
    (omni_drones/envs/isaac_env.py:233-234), which is absent from the reference
    tree and from this image.  The scheme below is our documented stand-in
    (DESIGN.md "Integrator"); the CUDA kernel is checked against *this* code.

Layout conventions: E envs, A pursuers, C cylinders, K observed cylinders,
F predicted steps, H history frames.  Quaternions are (w, x, y, z).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

F32 = torch.float32

# stats slot order follows the spec declaration order, hideandseek.py:400-425
STAT_KEYS = (
    "success", "collision", "blocked", "distance_reward", "distance_predicted_reward",
    "speed_reward", "collision_reward", "collision_wall", "collision_cylinder",
    "collision_drone", "detect_reward", "catch_reward", "smoothness_reward",
    "smoothness_mean", "smoothness_max", "first_capture_step", "sum_detect_step",
    "return", "action_error_order1_mean", "action_error_order1_max",
    "target_predicted_error", "distance_threshold_L", "out_of_arena", "smoothness_coef",
)
S = {k: i for i, k in enumerate(STAT_KEYS)}
# stats that are divided by the episode length on the `done` tick, hideandseek.py:1017-1056
STATS_DIV_ON_DONE = (
    "collision", "action_error_order1_mean", "target_predicted_error", "smoothness_mean",
    "smoothness_reward", "distance_reward", "detect_reward", "catch_reward",
    "collision_reward", "collision_wall", "collision_drone", "collision_cylinder",
    "speed_reward",
)


@dataclass
class HSParams:
    """Constants of the task.  Defaults = cfg/task/HideAndSeek.yaml + crazyflie.yaml
    + the USD-derived rigid-body constants of SURVEY.md Appendix B."""
    num_agents: int = 3
    num_cylinders: int = 5            # cylinder.max_num
    obs_max_cylinder: int = 3
    future_step: int = 5              # future_predcition_step
    history_step: int = 10
    use_tp_net: bool = True
    use_obstacles: bool = False       # TP frame also carries [x, y, size] per cylinder (hideandseek.py:808-817)
    max_episode_length: int = 800
    dt: float = 0.01
    # arena / task
    arena_size: float = 0.9
    max_height: float = 1.2
    cylinder_size: float = 0.1
    catch_radius: float = 0.3
    collision_radius: float = 0.07
    drone_detect_radius: float = 100.0
    target_detect_radius: float = 100.0
    v_drone: float = 1.0
    v_prey: float = 1.3               # cfg v_prey * v_drone, hideandseek.py:263
    mask_value: float = -5.0
    dist_reward_coef: float = 1.0
    catch_reward_coef: float = 20.0
    detect_reward_coef: float = 0.0
    collision_coef: float = 100.0
    speed_coef: float = 10.0
    smoothness_coef: float = 0.0      # init_smoothness_coef; the coefficient used is
    smooth_lr: float = 0.0            #   min(max_smoothness_coef, init + smooth_lr * update_epoch), recomputed at
    max_smoothness_coef: float = 5.0  #   every reward call (hideandseek.py:988-989; envgen: the constant, :465)
    use_deployment: bool = False      # HideAndSeek gates smoothness on this; envgen does not
    envgen_variant: bool = False
    # controller (crazyflie.yaml:4-6, lee_position_controller.py:448-454)
    target_clip: float = 1.0
    max_thrust_ratio: float = 0.9
    fixed_yaw: bool = False
    pid_kp: tuple = (250.0, 250.0, 120.0)
    pid_ki: tuple = (500.0, 500.0, 16.7)
    pid_kd: tuple = (2.5, 2.5, 0.0)
    pid_ilimit: tuple = (33.3, 33.3, 166.7)
    pid_out_limit: float = 2.0 ** 15 - 1.0
    # rotors (rotor_group.py:29-53, crazyflie.yaml:17-52)
    force_constant: float = 2.350347298350041e-08
    moment_constant: float = 7.24e-10
    max_rot_vel: float = 2315.0
    time_constant: float = 0.025
    rotor_dirs: tuple = (-1.0, 1.0, -1.0, 1.0)
    rotor_xy: tuple = ((0.028, 0.028), (-0.028, 0.028), (-0.028, -0.028), (0.028, -0.028))
    drag_coef: float = 0.0
    downwash_kr: float = 2.0
    downwash_kz: float = 0.3
    # rigid body (PhysX stand-in)
    base_mass: float = 0.0321
    rotor_mass: float = 1.0e-4
    base_inertia: tuple = (1.4e-5, 1.4e-5, 2.17e-5)
    gravity: float = 9.81
    linear_damping: float = 0.2
    angular_damping: float = 0.2
    max_linear_velocity: float = 1.0
    max_angular_velocity: float = 1000.0
    ground_clamp: bool = True
    ground_z: float = 0.0125          # collider half height, USD cylinder h=0.025
    contact_mode: int = 0             # 1: analytic cylinder / evader contacts after the integration (unpinned stand-in)
    drone_radius: float = 0.06        # base_link collider radius (USD)
    evader_radius: float = 0.05       # hideandseek.py:544-551

    # derived -----------------------------------------------------------------
    @property
    def total_mass(self) -> float:
        return self.base_mass + 4 * self.rotor_mass

    @property
    def inertia(self) -> tuple:
        # composite of base link + 4 point masses at rotor_xy
        sx = sum(self.rotor_mass * y * y for _, y in self.rotor_xy)
        sy = sum(self.rotor_mass * x * x for x, _ in self.rotor_xy)
        return (self.base_inertia[0] + sx, self.base_inertia[1] + sy,
                self.base_inertia[2] + sx + sy)

    @property
    def kf(self) -> float:
        # float32 like the reference parameter tensors (rotor_group.py:42)
        return float(torch.tensor(self.max_rot_vel, dtype=F32).square()
                     * torch.tensor(self.force_constant, dtype=F32))

    @property
    def km(self) -> float:
        return float(torch.tensor(self.max_rot_vel, dtype=F32).square()
                     * torch.tensor(self.moment_constant, dtype=F32))

    @property
    def tp_frame_dim(self) -> int:
        return 7 + 3 * self.num_agents + (3 * self.num_cylinders if self.use_obstacles else 0)

    @property
    def self_dim(self) -> int:
        return 3 + (3 * self.future_step if self.use_tp_net else 0) + 4 + 13


# ----------------------------------------------------------------------------
# quaternion helpers (omni_drones/utils/torch.py:182-201, 221-225)
# ----------------------------------------------------------------------------
def _qsplit(q):
    return q[..., 0:1], q[..., 1:4]


def quat_apply(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Rotate v by q: a + b + c with the reference's operation order."""
    w, u = _qsplit(q)
    a = v * (2.0 * w ** 2 - 1.0)
    b = torch.linalg.cross(u, v, dim=-1) * w * 2.0
    c = u * (u * v).sum(-1, keepdim=True) * 2.0
    return a + b + c


def quat_apply_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    w, u = _qsplit(q)
    a = v * (2.0 * w ** 2 - 1.0)
    b = torch.linalg.cross(u, v, dim=-1) * w * 2.0
    c = u * (u * v).sum(-1, keepdim=True) * 2.0
    return a - b + c


def quat_basis(q: torch.Tensor, axis: int) -> torch.Tensor:
    e = torch.zeros(*q.shape[:-1], 3, dtype=q.dtype)
    e[..., axis] = 1.0
    return quat_apply(q, e)


def quat_product(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a (x) b (used only by the integrator stand-in)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ], dim=-1)


def euler_to_quat(rpy: torch.Tensor) -> torch.Tensor:
    """omni_drones/utils/torch.py:110-127."""
    r, p, y = rpy.unbind(-1)
    cy, sy = torch.cos(y * 0.5), torch.sin(y * 0.5)
    cp, sp = torch.cos(p * 0.5), torch.sin(p * 0.5)
    cr, sr = torch.cos(r * 0.5), torch.sin(r * 0.5)
    return torch.stack([
        cr * cp * cy + sr * sp * sy,
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
    ], dim=-1)


# ----------------------------------------------------------------------------
# stage 1: CTBR transform + body-rate PID
# ----------------------------------------------------------------------------
def ctbr_pid(P: HSParams, raw_action, quat, angvel, prev_action, integ, last_rate, reset_pid):
    """transforms.py:425-459 feeding lee_position_controller.py:476-550.

    raw_action [E,A,4]; quat [E,A,4]; angvel (world) [E,A,3]; prev_action [E,A,4];
    integ / last_rate [E,A,3] (persistent, returned updated); reset_pid bool [E].
    Returns dict(cmds, ctbr, target_rate, action_error, prev_action, integ, last_rate).
    """
    a = torch.tanh(raw_action)
    rate = a[..., :3].clone()
    thrust = torch.clamp((a[..., 3:4] + 1) / 2, min=0.0, max=P.max_thrust_ratio)
    if P.fixed_yaw:
        rate[..., 2] = 0.0
    ctbr_action = torch.cat([rate, thrust], dim=-1)
    action_error = torch.linalg.vector_norm(ctbr_action - prev_action, dim=-1)
    target_rate = rate * 180.0 * P.target_clip                # deg/s
    target_thrust = thrust * 2 ** 16                          # 16-bit units

    integ = integ.clone()
    last_rate = last_rate.clone()
    m = reset_pid.reshape(-1, 1, 1).expand_as(integ)
    integ[m] = 0.0
    last_rate[m] = 0.0

    dt = torch.tensor(P.dt, dtype=F32)
    kp = torch.tensor(P.pid_kp, dtype=F32)
    ki = torch.tensor(P.pid_ki, dtype=F32)
    kd = torch.tensor(P.pid_kd, dtype=F32)
    ilim = torch.tensor(P.pid_ilimit, dtype=F32)
    olim = torch.tensor(P.pid_out_limit, dtype=F32)

    body_rate = quat_apply_inverse(quat, angvel) * 180.0 / torch.pi
    err = target_rate - body_rate
    out_p = err * kp
    deriv = -(body_rate - last_rate) / dt
    deriv = torch.where(torch.isnan(deriv), torch.zeros_like(deriv), deriv)
    out_d = deriv * kd
    integ = integ + err * dt
    integ = torch.clip(integ, -ilim, ilim)
    out_i = integ * ki
    out_ff = target_rate * 0.0
    out = out_p + out_d + out_i + out_ff
    out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
    out = torch.clip(out, -olim, olim)
    last_rate = body_rate.clone()

    r = out[..., 0:1] / 2.0
    p = out[..., 1:2] / 2.0
    y = out[..., 2:3]
    T = target_thrust
    motors = torch.cat([T + r - p + y, T + r + p - y, T - r + p + y, T - r - p - y], dim=-1)
    cmds = motors / 2 ** 16 * 2 - P.max_thrust_ratio
    cmds = torch.nan_to_num(cmds, 0.0)
    ctbr = torch.cat([r, p, y, T], dim=-1)
    return dict(cmds=cmds, ctbr=ctbr, target_rate=target_rate, action_error=action_error,
                prev_action=ctbr_action, integ=integ, last_rate=last_rate)


# ----------------------------------------------------------------------------
# stage 2: rotor model
# ----------------------------------------------------------------------------
def rotor_model(P: HSParams, cmds, throttle):
    """rotor_group.py:55-71.  Returns thrusts, moments [E,A,4] and the new throttle."""
    target = torch.sqrt(torch.clamp((cmds + 1) / 2, 0, 1))
    tau = torch.full_like(throttle, P.time_constant)           # tau_up == tau_down
    tau = torch.clamp(tau, 0, 1)
    alpha = P.dt / tau
    throttle = throttle + alpha * (target - throttle)
    t = torch.clamp(throttle.square() + torch.zeros_like(throttle), 0.0, 1.0)
    kf = torch.tensor(P.kf, dtype=F32)
    km = torch.tensor(P.km, dtype=F32)
    dirs = torch.tensor(P.rotor_dirs, dtype=F32)
    thrusts = t * kf
    moments = (t * km) * -dirs
    return thrusts, moments, throttle


# ----------------------------------------------------------------------------
# stage 3: downwash (multirotor.py:724-753)
# ----------------------------------------------------------------------------
def downwash_force(P: HSParams, pos, thrust_world):
    """Force on drone i from every other drone j's wake.  pos, thrust_world [E,A,3]."""
    A = pos.shape[1]
    if A < 2:
        return torch.zeros_like(pos)
    d = thrust_world / (torch.linalg.vector_norm(thrust_world, dim=-1, keepdim=True) + 1e-6)
    rel = pos.unsqueeze(1) - pos.unsqueeze(2)                  # [E,i,j,3] = p_j - p_i
    dj = d.unsqueeze(1)                                        # direction of the *source* j
    z_dist = (rel * dj).sum(-1, keepdim=True)
    r_dist = torch.linalg.vector_norm(rel - z_dist * dj, dim=-1, keepdim=True)
    z = torch.clip(z_dist, 0)
    v = torch.exp(-0.5 * torch.square(P.downwash_kr * r_dist / z)) / (1 + P.downwash_kz * z) ** 2
    contrib = v * -thrust_world.unsqueeze(1)                   # [E,i,j,3]
    off = ~torch.eye(A, dtype=torch.bool)
    contrib = torch.where(off.reshape(1, A, A, 1), contrib, torch.zeros_like(contrib))
    # the reference sums the A-1 off-diagonal terms in increasing j
    out = torch.zeros_like(pos)
    for j in range(A):
        out = out + contrib[:, :, j]
    return out


# ----------------------------------------------------------------------------
# line of sight (hideandseek.py:47-103)
# ----------------------------------------------------------------------------
def los_blocked(P: HSParams, drone_pos, target_pos, cyl_pos):
    """bool [E,A]: is the xy segment drone->target cut by an above-ground cylinder."""
    d = drone_pos - target_pos.unsqueeze(1)                    # [E,A,3]
    c = cyl_pos - target_pos.unsqueeze(1)                      # [E,C,3]
    cross = torch.abs(d[..., 0:1] * c[..., 1].unsqueeze(1) - d[..., 1:2] * c[..., 0].unsqueeze(1))
    seg = torch.sqrt(d[..., 0:1] ** 2 + d[..., 1:2] ** 2)
    near = cross / (seg + 1e-5) <= P.cylinder_size            # [E,A,C]
    dx = target_pos[:, None, 0] - drone_pos[..., 0]            # [E,A]
    dy = target_pos[:, None, 1] - drone_pos[..., 1]
    num = (cyl_pos[:, None, :, 0] - drone_pos[..., 0:1]) * dx.unsqueeze(2) \
        + (cyl_pos[:, None, :, 1] - drone_pos[..., 1:2]) * dy.unsqueeze(2)
    den = dx.unsqueeze(2) ** 2 + dy.unsqueeze(2) ** 2
    t = num / (den + 1e-5)
    between = (t >= 0) & (t <= 1)
    standing = (cyl_pos[..., 2] > 0.0).unsqueeze(1)
    return (near & between & standing).any(-1)


# ----------------------------------------------------------------------------
# stage 4: potential-field evader (hideandseek.py:1067-1141 and 737-744)
# ----------------------------------------------------------------------------
def evader_force(P: HSParams, drone_pos, target_pos, cyl_pos, cyl_inactive):
    """hideandseek.py:1067-1141.  Returns (force [E,3], magnitude [E,3] = sum of |terms| per component, i.e. the
    scale of the rounding error the cancellation in `force` is exposed to, out_of_arena bool [E])."""
    E = drone_pos.shape[0]
    rel = drone_pos - target_pos.unsqueeze(1)                  # drone - evader [E,A,3]
    dist = torch.linalg.vector_norm(rel, dim=-1, keepdim=True)  # [E,A,1]
    blocked = los_blocked(P, drone_pos, target_pos, cyl_pos)
    active = (dist < P.target_detect_radius) & (~blocked).unsqueeze(-1)
    away = -rel / (dist + 1e-5)
    f_p = away * (1 / (dist + 1e-5)) * active
    force = torch.zeros(E, 3, dtype=F32)
    A = drone_pos.shape[1]
    acc = f_p[:, 0]
    for a in range(1, A):
        acc = acc + f_p[:, a]
    force = force + acc
    mag = f_p.abs().sum(1)

    # arena wall, ceiling, floor
    rho = torch.linalg.vector_norm(target_pos[:, :2], dim=-1)  # [E]
    inward = -target_pos[:, :2] / (rho.unsqueeze(-1) + 1e-5)
    outside = target_pos[:, 0] ** 2 + target_pos[:, 1] ** 2 > P.arena_size ** 2
    o = outside.float()
    no = (~outside).float()
    f_r = torch.zeros(E, 3, dtype=F32)
    for ax in range(2):
        f_r[:, ax] = o * inward[:, ax] * (1 / 1e-5) \
            + no * inward[:, ax] * (1 / ((P.arena_size - rho) + 1e-5))
    z = target_pos[:, 2]
    hi = z > P.max_height
    up = hi.float() * (-1 / 1e-5) \
        + (~hi).float() * -(P.max_height - z) / ((P.max_height - z) ** 2 + 1e-5)
    lo = z < 0.0
    down = (lo.float() * (1 / 1e-5)
            + (~lo).float() * -(0.0 - z) / ((0.0 - z) ** 2 + 1e-5))
    f_r[:, 2] = up
    f_r[:, 2] += down
    force = force + f_r
    mag = mag + torch.stack([f_r[:, 0].abs(), f_r[:, 1].abs(), up.abs() + down.abs()], dim=-1)

    # cylinders: xy repulsion from every active cylinder
    tc = target_pos.unsqueeze(1) - cyl_pos                     # [E,C,3]
    d_xy = torch.linalg.vector_norm(tc[..., :2], dim=-1)       # [E,C]
    gap = d_xy - P.cylinder_size
    act = ((~cyl_inactive) & (d_xy < P.target_detect_radius)).float()
    dir_xy = tc[..., :2] / (d_xy + 1e-5).unsqueeze(-1)
    terms = act.unsqueeze(-1) * dir_xy * (1 / (gap.unsqueeze(-1) + 1e-5))
    f_c = torch.zeros(E, 3, dtype=F32)
    f_c[:, :2] = terms.sum(1)
    force = force + f_c
    mag[:, :2] = mag[:, :2] + terms.abs().sum(1)
    return force.to(F32), mag, outside


def evader_velocity(P: HSParams, v_prey: float, drone_pos, target_pos, cyl_pos, cyl_inactive):
    """Returns (new evader velocity [E,3], out_of_arena bool [E])."""
    force, _, outside = evader_force(P, drone_pos, target_pos, cyl_pos, cyl_inactive)
    # per-component normalisation (torch.norm over the size-1 agent dim), hideandseek.py:741
    vel = v_prey * force / (torch.abs(force) + 1e-5)
    return vel.to(F32), outside


# ----------------------------------------------------------------------------
# stage 5: rigid-body integration -- PhysX stand-in, PARITY UNPINNED
# ----------------------------------------------------------------------------
def rigid_body_step(P: HSParams, pos, quat, linvel, angvel, thrusts, yaw_torque, ext_force):
    """One semi-implicit Euler step of every drone.

    thrusts [E,A,4] per-rotor body-z thrust and yaw_torque [E,A] = sum of the rotor
    reaction moments (body z), or None for the unforced step inside reset
    (hideandseek.py:722-723); ext_force world [E,A,3] or None.
    Order: wrench -> v += dt(F/m+g), w_b += dt I^-1(tau_b - w_b x I w_b) -> damping
    v*=max(0,1-dt c) -> clamp |v|, |w| -> p += dt v -> q <- normalize(dq(w dt) (x) q)
    -> optional ground clamp.
    """
    dt = P.dt
    m = P.total_mass
    Ix, Iy, Iz = P.inertia
    I = torch.tensor([Ix, Iy, Iz], dtype=F32)
    if thrusts is not None:
        total = thrusts.sum(-1)
        fb = torch.zeros_like(pos)
        fb[..., 2] = total
        force = quat_apply(quat, fb)
        rx = torch.tensor([xy[0] for xy in P.rotor_xy], dtype=F32)
        ry = torch.tensor([xy[1] for xy in P.rotor_xy], dtype=F32)
        tau_b = torch.stack([(ry * thrusts).sum(-1), (-rx * thrusts).sum(-1), yaw_torque], dim=-1)
    else:
        force = torch.zeros_like(pos)
        tau_b = torch.zeros_like(pos)
    if ext_force is not None:
        force = force + ext_force
    acc = force / m
    acc[..., 2] = acc[..., 2] - P.gravity
    v = linvel + dt * acc
    wb = quat_apply_inverse(quat, angvel)
    gyro = torch.linalg.cross(wb, I * wb, dim=-1)
    wb = wb + dt * ((tau_b - gyro) / I)
    w = quat_apply(quat, wb)
    v = v * max(0.0, 1.0 - dt * P.linear_damping)
    w = w * max(0.0, 1.0 - dt * P.angular_damping)
    vn = torch.linalg.vector_norm(v, dim=-1, keepdim=True)
    # The clamp lands a hair *below* the limit (factor 1-1e-6): the task penalises
    # |v| > v_drone with v_drone == max_linear_velocity (hideandseek.py:539, 954-957), so a
    # clamp that lands exactly on the limit would make that penalty rounding noise.
    v = torch.where(vn > P.max_linear_velocity, v * (P.max_linear_velocity * (1.0 - 1e-6) / vn), v)
    wn = torch.linalg.vector_norm(w, dim=-1, keepdim=True)
    w = torch.where(wn > P.max_angular_velocity, w * (P.max_angular_velocity / wn), w)
    p = pos + dt * v
    wn = torch.linalg.vector_norm(w, dim=-1, keepdim=True)     # |w| after the clamp
    half = 0.5 * dt * wn
    small = wn < 1e-6
    k = torch.where(small, torch.full_like(wn, 0.5 * dt), torch.sin(half) / wn.clamp(min=1e-6))
    wq = torch.where(small, torch.ones_like(wn), torch.cos(half))
    dq = torch.cat([wq, w * k], dim=-1)
    q = quat_product(dq, quat)
    q = q / torch.linalg.vector_norm(q, dim=-1, keepdim=True)
    if P.ground_clamp:
        under = p[..., 2] < P.ground_z
        p[..., 2] = torch.where(under, torch.full_like(p[..., 2], P.ground_z), p[..., 2])
        v[..., 2] = torch.where(under & (v[..., 2] < 0), torch.zeros_like(v[..., 2]), v[..., 2])
    return p, q, v, w


def apply_contacts(P: HSParams, p, v, tpos_old, cyl):
    """contact_mode = 1 (PhysX stand-in, PARITY UNPINNED; the CUDA kernel's stage_contacts is the same arithmetic): project
    every pursuer out of the standing cylinders it penetrates (2-D, below the cylinder top, in cylinder order) and out of the
    evader's sphere (evader position at the start of the tick); the inward normal velocity is removed."""
    if not P.contact_mode:
        return p, v
    p, v = p.clone(), v.clone()
    Rc = P.cylinder_size + P.drone_radius
    for k in range(cyl.shape[1]):
        c = cyl[:, k].unsqueeze(1)                                # [E,1,3]
        dx, dy = p[..., 0] - c[..., 0], p[..., 1] - c[..., 1]
        d = torch.sqrt(dx * dx + dy * dy)
        hit = (c[..., 2] > 0.0) & (p[..., 2] < 2.0 * c[..., 2]) & (d < Rc)
        inv = 1.0 / d.clamp(min=1e-6)
        nx, ny = dx * inv, dy * inv
        p[..., 0] = torch.where(hit, c[..., 0] + nx * Rc, p[..., 0])
        p[..., 1] = torch.where(hit, c[..., 1] + ny * Rc, p[..., 1])
        vn = v[..., 0] * nx + v[..., 1] * ny
        rem = hit & (vn < 0)
        v[..., 0] = torch.where(rem, v[..., 0] - vn * nx, v[..., 0])
        v[..., 1] = torch.where(rem, v[..., 1] - vn * ny, v[..., 1])
    Re = P.evader_radius + P.drone_radius
    rel = p - tpos_old.unsqueeze(1)
    d = torch.linalg.vector_norm(rel, dim=-1, keepdim=True)
    hit = d < Re
    n = rel / d.clamp(min=1e-6)
    p = torch.where(hit, tpos_old.unsqueeze(1) + n * Re, p)
    vn = (v * n).sum(-1, keepdim=True)
    v = torch.where(hit & (vn < 0), v - n * vn, v)
    return p, v


# ----------------------------------------------------------------------------
# stage 6: observation (hideandseek.py:746-917)
# ----------------------------------------------------------------------------
def k_nearest_cylinders(P: HSParams, drone_pos, cyl_pos):
    """Returns (features [E,A,K,5] unmasked, masked copy, inactive mask [E,A,K], inactive [E,C]).
    Ties in the sort key are broken towards the lowest cylinder index (stable sort)."""
    E, A, _ = drone_pos.shape
    C = cyl_pos.shape[1]
    inactive = cyl_pos[..., 2] < 0.0
    rpos = drone_pos.unsqueeze(2) - cyl_pos.unsqueeze(1)       # [E,A,C,3]
    feat = torch.cat([rpos,
                      torch.full((E, A, C, 1), P.max_height, dtype=F32),
                      torch.full((E, A, C, 1), P.cylinder_size, dtype=F32)], dim=-1)
    key = torch.linalg.vector_norm(rpos, dim=-1) - P.cylinder_size
    order = torch.sort(key, dim=-1, stable=True).indices[..., :P.obs_max_cylinder]
    near = feat.gather(2, order.unsqueeze(-1).expand(-1, -1, -1, 5))
    near_inactive = inactive.unsqueeze(1).expand(-1, A, -1).gather(2, order)
    masked = torch.where(near_inactive.unsqueeze(-1), torch.full_like(near, P.mask_value), near)
    return near, masked, near_inactive, inactive


def observe(P: HSParams, st: Dict[str, torch.Tensor], tp_pred: Optional[torch.Tensor]):
    """Builds every observation tensor from the current state.  `tp_pred` is the raw
    TP_net output [E, 3F] in (-1,1) or None (slots left as zeros / absent)."""
    pos, quat, linvel, angvel = st["pos"], st["quat"], st["linvel"], st["angvel"]
    tpos, tvel, cyl, progress = st["tpos"], st["tvel"], st["cyl"], st["progress"]
    E, A, _ = pos.shape
    heading = quat_basis(quat, 0)
    up = quat_basis(quat, 2)
    drone_state13 = torch.cat([pos, quat, linvel, angvel], dim=-1)

    # others: p_a - p_j for j != a in increasing j
    rel = pos.unsqueeze(2) - pos.unsqueeze(1)                  # [E,a,j,3]
    idx = torch.tensor([[j for j in range(A) if j != a] for a in range(A)], dtype=torch.long)
    others = rel.gather(2, idx.reshape(1, A, A - 1, 1).expand(E, -1, -1, 3)) if A > 1 else None

    near, near_masked, near_inactive, inactive = k_nearest_cylinders(P, pos, cyl)

    t_rpos = pos - tpos.unsqueeze(1)                           # [E,A,3]
    blocked = los_blocked(P, pos, tpos, cyl)
    detect = (torch.linalg.vector_norm(t_rpos, dim=-1) < P.drone_detect_radius) & (~blocked)
    bdetect = detect.any(dim=1)                                # [E]
    hidden = ~bdetect
    mv = P.mask_value
    t_rpos_masked = torch.where(hidden.reshape(E, 1, 1), torch.full_like(t_rpos, mv), t_rpos)
    tpos_masked = torch.where(hidden.unsqueeze(-1), torch.full_like(tpos, mv), tpos)
    tvel_masked = torch.where(hidden.unsqueeze(-1), torch.full_like(tvel, mv), tvel)
    t = (progress / P.max_episode_length).reshape(E, 1, 1).expand(E, A, 4)

    out = dict(drone_state=drone_state13, others=others, cylinders=near_masked,
               blocked=blocked, bdetect=bdetect, near=near, near_inactive=near_inactive,
               cyl_inactive=inactive, heading=heading, up=up)

    if P.use_tp_net:
        frame = torch.cat([progress.unsqueeze(-1), tpos_masked, tvel_masked, pos.reshape(E, -1)], dim=-1)
        if P.use_obstacles:
            frame = torch.cat([frame, torch.cat([cyl[..., :2], torch.full((E, cyl.shape[1], 1), P.cylinder_size, dtype=F32)],
                                                dim=-1).reshape(E, -1)], dim=-1)
        out["tp_frame"] = frame
        out["tp_done"] = (progress <= (P.max_episode_length - P.future_step)).unsqueeze(-1)
        gt = tpos.clone()
        gt[:, :2] = gt[:, :2] / (0.5 * P.arena_size)
        gt[:, 2] = gt[:, 2] / P.max_height * 2.0 - 1.0
        out["tp_groundtruth"] = gt
        if tp_pred is not None:
            pred = tp_pred.reshape(E, P.future_step, 3).clone()
            pred[..., :2] = pred[..., :2] * 0.5 * P.arena_size
            pred[..., 2] = (pred[..., 2] + 1.0) / 2.0 * P.max_height
            rp = (pos.unsqueeze(2) - pred.unsqueeze(1)).reshape(E, A, -1)
        else:
            rp = torch.zeros(E, A, 3 * P.future_step, dtype=F32)
        tail = [rp, quat, linvel, heading, up, t]
    else:
        tail = [quat, linvel, heading, up, t]
    out["state_self"] = torch.cat([t_rpos_masked] + tail, dim=-1).unsqueeze(2)
    out["state_drones"] = torch.cat([t_rpos] + tail, dim=-1)
    return out


# ----------------------------------------------------------------------------
# stage 7: reward / done / stats (hideandseek.py:919-1065)
# ----------------------------------------------------------------------------
def reward_done(P: HSParams, st, obs, action_error, throttle_diff, stats, smoothness_coef=None):
    """Returns reward [E,A,1], done [E,1]; updates `stats` [E,24] in place.  `smoothness_coef`: the value
    hideandseek.py:988-989 recomputes from update_epoch (default: P.smoothness_coef, i.e. update_epoch = 0)."""
    coef = P.smoothness_coef if smoothness_coef is None else smoothness_coef
    pos, linvel, tpos, progress = st["pos"], st["linvel"], st["tpos"], st["progress"]
    E, A, _ = pos.shape
    blocked, bdetect = obs["blocked"], obs["bdetect"]

    def add(key, per_agent):
        stats[:, S[key]] += per_agent.mean(-1)

    dist = torch.linalg.vector_norm(tpos.unsqueeze(1) - pos, dim=-1)       # [E,A]
    r_dist = -P.dist_reward_coef * dist * (dist > P.catch_radius).float()
    add("distance_reward", r_dist)

    r_detect = P.detect_reward_coef * bdetect.unsqueeze(-1).expand(E, A)
    stats[:, S["sum_detect_step"]] += 1.0 * bdetect
    add("detect_reward", r_detect.float())

    capture = dist < P.catch_radius
    seen_capture = capture * (~blocked).float()
    any_capture = torch.any(seen_capture, dim=-1)                          # [E]
    r_catch = P.catch_reward_coef * any_capture.unsqueeze(-1).expand(E, A)
    capture_flag = torch.any(r_catch, dim=1)
    stats[:, S["blocked"]] += torch.all(blocked, dim=-1)
    stats[:, S["success"]] = torch.logical_or(capture_flag, stats[:, S["success"]]).float()
    step_now = capture_flag.float() * progress + (~capture_flag).float() * P.max_episode_length
    stats[:, S["first_capture_step"]] = torch.min(stats[:, S["first_capture_step"]], step_now)
    add("catch_reward", r_catch)

    speed = torch.linalg.vector_norm(linvel, dim=-1)
    r_speed = -P.speed_coef * (speed > P.v_drone)
    add("speed_reward", r_speed.float())

    near_xy = torch.linalg.vector_norm(obs["near"][..., :2], dim=-1)       # [E,A,K]
    hit_cyl = (near_xy - P.cylinder_size < P.collision_radius).float()
    hit_cyl = torch.where(obs["near_inactive"], torch.zeros_like(hit_cyl), hit_cyl).sum(-1)
    r_coll = -P.collision_coef * hit_cyl
    add("collision_cylinder", hit_cyl)
    if A > 1:
        dd = torch.linalg.vector_norm(obs["others"], dim=-1)
        hit_drone = (dd < 2.0 * P.collision_radius).float().sum(-1)
    else:
        hit_drone = torch.zeros(E, A, dtype=F32)
    r_coll = r_coll + -P.collision_coef * hit_drone
    add("collision_drone", hit_drone)
    hit_wall = (pos[..., 2] > P.max_height).float() \
        + ((pos[..., 0] ** 2 + pos[..., 1] ** 2) > P.arena_size ** 2).float()
    r_coll = r_coll + -P.collision_coef * hit_wall
    stats[:, S["collision"]] += torch.any(r_coll < 0, dim=1)
    add("collision_wall", hit_wall)
    add("collision_reward", r_coll)

    if not P.envgen_variant:
        stats[:, S["smoothness_coef"]] = coef
    r_smooth = coef * torch.exp(-action_error)
    if (not P.envgen_variant) and (not P.use_deployment):
        r_smooth = torch.zeros_like(r_smooth)
    add("smoothness_reward", r_smooth)
    add("smoothness_mean", throttle_diff)
    stats[:, S["smoothness_max"]] = torch.max(throttle_diff.max(-1).values, stats[:, S["smoothness_max"]])

    reward = r_dist + r_detect + r_catch + r_coll + r_speed + r_smooth
    done = progress >= P.max_episode_length
    ep_len = torch.where(done, progress, torch.ones_like(progress))
    for k in STATS_DIV_ON_DONE:
        stats[:, S[k]] /= ep_len
    stats[:, S["return"]] += reward.mean(-1)
    return reward.unsqueeze(-1), done.unsqueeze(-1)


# ----------------------------------------------------------------------------
# the environment
# ----------------------------------------------------------------------------
class HideAndSeekOracle:
    """Holds the per-env state and sequences reset()/step() like IsaacEnv._reset/_step."""

    def __init__(self, params: HSParams, num_envs: int):
        self.P = params
        self.E = num_envs
        P, E, A, C = params, num_envs, params.num_agents, params.num_cylinders
        z = lambda *s: torch.zeros(*s, dtype=F32)
        self.st = dict(pos=z(E, A, 3), quat=z(E, A, 4), linvel=z(E, A, 3), angvel=z(E, A, 3),
                       tpos=z(E, 3), tvel=z(E, 3), cyl=z(E, C, 3), progress=z(E))
        self.st["quat"][..., 0] = 1.0
        self.st["cyl"][..., 2] = -20.0
        self.throttle = z(E, A, 4)
        self.integ = z(E, A, 3)
        self.last_rate = z(E, A, 3)
        self.prev_action = z(E, A, 4)
        self.stats = z(E, len(STAT_KEYS))
        self.tp_hist = None                 # [E,H,frame] once the first frame arrives
        self.v_prey = float(params.v_prey)
        self.update_epoch = 0               # scripts/train_deploy.py writes base_env.update_epoch = i
        self.last = {}

    def smoothness_coef(self) -> float:
        P = self.P
        if P.envgen_variant:
            return P.smoothness_coef
        return min(P.max_smoothness_coef, P.smoothness_coef + P.smooth_lr * self.update_epoch)

    # -- helpers --------------------------------------------------------------
    def hover_throttle(self) -> float:
        P = self.P
        g = torch.tensor(P.total_mass, dtype=F32) * 9.81
        return float(torch.sqrt(g / (4 * torch.tensor(P.kf, dtype=F32))))

    def _push_frame(self, frame):
        H = self.P.history_step
        if self.tp_hist is None:
            self.tp_hist = frame.unsqueeze(1).repeat(1, H, 1)
        else:
            self.tp_hist = torch.cat([self.tp_hist[:, 1:], frame.unsqueeze(1)], dim=1)
        return self.tp_hist.clone()

    def _observe(self, tp_fn):
        P = self.P
        obs = observe(P, self.st, None)
        if P.use_tp_net:
            tp_input = self._push_frame(obs["tp_frame"])
            obs["tp_input"] = tp_input
            if tp_fn is not None:
                pred = tp_fn(tp_input)
                obs2 = observe(P, self.st, pred)
                obs["state_self"], obs["state_drones"] = obs2["state_self"], obs2["state_drones"]
                obs["tp_pred"] = pred
        self.last = obs
        return obs

    # -- API ------------------------------------------------------------------
    def reset(self, mask: torch.Tensor, init: Dict[str, torch.Tensor], tp_fn=None):
        """mask bool [E]; init has drone_pos [E,A,3], drone_rot [E,A,4], target_pos [E,3],
        cyl_pos [E,C,3] (rows outside the mask ignored).  hideandseek.py:609-723 +
        isaac_env.py:210-225."""
        P, st = self.P, self.st
        last_stats = self.stats.clone()
        m = mask.bool()
        st["pos"][m] = init["drone_pos"][m]
        st["quat"][m] = init["drone_rot"][m]
        st["linvel"][m] = 0.0
        st["angvel"][m] = 0.0
        st["tpos"][m] = init["target_pos"][m]          # evader velocity is NOT reset
        st["cyl"][m] = init["cyl_pos"][m]
        self.throttle[m] = self.hover_throttle()
        self.stats[m] = 0.0
        self.stats[:, S["first_capture_step"]] = float(P.max_episode_length)   # all envs
        cmd_init = 2.0 * self.throttle[m] ** 2 - 1.0
        self.prev_action[m, :, 3] = (0.5 * (P.max_thrust_ratio + cmd_init)).mean(-1)
        # one unforced physics tick for every env (hideandseek.py:722-723)
        p, q, v, w = rigid_body_step(P, st["pos"], st["quat"], st["linvel"], st["angvel"], None, None, None)
        p, v = apply_contacts(P, p, v, st["tpos"], st["cyl"])
        st["pos"], st["quat"], st["linvel"], st["angvel"] = p, q, v, w
        st["tpos"] = st["tpos"] + P.dt * st["tvel"]
        st["progress"][m] = 0.0
        obs = self._observe(tp_fn)
        obs["last_stats"] = last_stats
        obs["truncated"] = (st["progress"] > P.max_episode_length).unsqueeze(1)
        return obs

    def step(self, raw_action: torch.Tensor, done_prev: torch.Tensor, tp_fn=None):
        """raw_action [E,A,4] (policy output before tanh); done_prev bool [E] is the
        `done` entry of the tensordict handed to step (PID reset mask)."""
        P, st = self.P, self.st
        c = ctbr_pid(P, raw_action, st["quat"], st["angvel"], self.prev_action,
                     self.integ, self.last_rate, done_prev.bool())
        self.integ, self.last_rate, self.prev_action = c["integ"], c["last_rate"], c["prev_action"]
        ae = c["action_error"]
        self.stats[:, S["action_error_order1_mean"]] += ae.mean(-1)
        self.stats[:, S["action_error_order1_max"]] = torch.max(
            self.stats[:, S["action_error_order1_max"]], ae.mean(-1))
        old_throttle = self.throttle
        thrusts, moments, self.throttle = rotor_model(P, c["cmds"], self.throttle)
        fb = torch.zeros_like(st["pos"])
        fb[..., 2] = thrusts.sum(-1)
        ext = downwash_force(P, st["pos"], quat_apply(st["quat"], fb))
        ext = ext + (P.drag_coef * P.base_mass) * st["linvel"]
        throttle_diff = torch.linalg.vector_norm(self.throttle - old_throttle, dim=-1)
        inactive = st["cyl"][..., 2] < 0.0
        tvel, outside = evader_velocity(P, self.v_prey, st["pos"], st["tpos"], st["cyl"], inactive)
        self.stats[:, S["out_of_arena"]] = torch.logical_or(self.stats[:, S["out_of_arena"]].bool(), outside).float()
        st["tvel"] = tvel
        p, q, v, w = rigid_body_step(P, st["pos"], st["quat"], st["linvel"], st["angvel"], thrusts, moments.sum(-1), ext)
        p, v = apply_contacts(P, p, v, st["tpos"], st["cyl"])
        st["pos"], st["quat"], st["linvel"], st["angvel"] = p, q, v, w
        st["tpos"] = st["tpos"] + P.dt * st["tvel"]
        st["progress"] = st["progress"] + 1
        obs = self._observe(tp_fn)
        reward, done = reward_done(P, st, obs, ae, throttle_diff, self.stats, self.smoothness_coef())
        if bool(done.any()) and float(self.stats[:, S["success"]].mean()) >= 0.98 and not P.envgen_variant:
            self.v_prey = min(1.3, self.v_prey + 0.05)
        obs.update(reward=reward, done=done, cmds=c["cmds"], ctbr=c["ctbr"],
                   target_rate=c["target_rate"], action_error=ae, throttle_diff=throttle_diff,
                   prev_action=self.prev_action.clone(), stats=self.stats.clone())
        return obs


# ----------------------------------------------------------------------------
# reset sampling (hideandseek.py:283-309, 576-607, 609-697) -- CPU generator
# ----------------------------------------------------------------------------
def sample_reset(P: HSParams, E: int, gen: torch.Generator, scenario: str = "random_cylinders",
                 min_cylinders: int = 4) -> Dict[str, torch.Tensor]:
    """Draws an initial configuration.  `scenario`: 'random_cylinders' (use_random_cylinder=1)
    or one of the fixed layouts 'empty', 'wall', 'narrow_gap', 'random', 'passage'."""
    A, C = P.num_agents, P.num_cylinders
    a = P.arena_size / math.sqrt(2.0)
    u = lambda lo, hi, *shape: lo + (hi - lo) * torch.rand(*shape, generator=gen, dtype=F32)
    rpy = torch.stack([u(-0.2 * math.pi, 0.2 * math.pi, E, A), u(-0.2 * math.pi, 0.2 * math.pi, E, A),
                       u(0.0, 0.2 * math.pi, E, A)], dim=-1)
    rot = euler_to_quat(rpy)
    cs, ch = P.cylinder_size, P.max_height
    cyl = torch.zeros(E, C, 3, dtype=F32)
    cyl[..., 0] = torch.arange(C, dtype=F32) * 2 * cs
    cyl[..., 2] = -20.0
    if scenario == "random_cylinders":
        dxy = torch.stack([u(0.1, a - 0.1, E, A), u(-a + 0.1, a - 0.1, E, A)], dim=-1)
        txy = torch.stack([u(-a + 0.1, -0.1, E), u(-a + 0.1, a - 0.1, E)], dim=-1)
        dpos = torch.cat([dxy, u(ch / 2 - 0.1, ch / 2 + 0.1, E, A, 1)], dim=-1)
        tpos = torch.cat([txy, u(ch / 2 - 0.1, ch / 2 + 0.1, E, 1)], dim=-1)
        n = int(round(P.arena_size * 2 / (2 * cs)))            # 9
        half = n // 2
        ii, jj = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
        occ0 = (torch.sqrt(((ii - half) ** 2 + (jj - half) ** 2).float()) >= half)
        occ = occ0.unsqueeze(0).repeat(E, 1, 1)
        to_cell = lambda xy: torch.clamp(torch.round(xy / (2 * cs)).int() + half, 0, n - 1).long()
        dc, tc = to_cell(dxy), to_cell(txy)
        ar = torch.arange(E)
        for k in range(A):
            occ[ar, dc[:, k, 0], dc[:, k, 1]] = True
        occ[ar, tc[:, 0], tc[:, 1]] = True
        n_active = torch.randint(min_cylinders, C + 1, (E,), generator=gen)
        score = torch.rand(E, n * n, generator=gen)
        score[occ.reshape(E, -1)] = 2.0
        pick = torch.argsort(score, dim=-1)[:, :C]
        gx, gy = pick // n, pick % n
        xy = torch.stack([gx, gy], dim=-1).float()
        xy = torch.clamp((xy - half) * (2 * cs), -(P.arena_size - 0.1), P.arena_size - 0.1)
        cyl[..., :2] = xy
        cyl[..., 2] = torch.where(torch.arange(C).unsqueeze(0) >= n_active.unsqueeze(1),
                                  torch.tensor(-20.0), torch.tensor(0.5 * ch))
    else:
        starts = {
            "empty": ([[0.6, 0.0, 0.5], [0.8, 0.0, 0.5], [0.8, -0.2, 0.5], [0.8, 0.2, 0.5]], [-0.8, 0.0, 0.5]),
            "wall": ([[0.6, 0.4, 0.5], [0.6, 0.0, 0.5], [0.6, -0.4, 0.5], [0.8, 0.2, 0.5]], [-0.8, 0.0, 0.5]),
            "narrow_gap": ([[0.0, 0.7, 0.5], [0.2, 0.7, 0.5], [-0.2, 0.7, 0.5], [0.8, 0.2, 0.5]], [-0.5, 0.2, 0.5]),
            "random": ([[0.6, 0.0, 0.5], [0.8, 0.0, 0.5], [0.8, -0.2, 0.5], [0.8, 0.2, 0.5]], [-0.8, 0.0, 0.5]),
            "passage": ([[0.6, 0.0, 0.5], [0.8, 0.2, 0.5], [0.8, -0.2, 0.5], [0.8, 0.2, 0.5]], [0.0, 0.6, 0.5]),
        }
        layouts = {
            "empty": [],
            "wall": [[0.0, 1.5], [0.0, -1.5], [0.0, 4.5], [0.0, -4.5]],
            "narrow_gap": [[3, -3], [3, 3], [-3, 3], [-3, -3], [0, 3]],
            "random": [[6, 4], [-6, 4], [-2, 4], [0, 2], [-2, -4], [0, -2]],
            "passage": [[0, 3], [-2, 3], [2, 3], [2, -2], [-2, -2], [0, -2]],
        }
        d, t = starts[scenario]
        dpos = torch.tensor(d[:A], dtype=F32).unsqueeze(0).repeat(E, 1, 1)
        tpos = torch.tensor(t, dtype=F32).unsqueeze(0).repeat(E, 1)
        lay = layouts[scenario]
        assert len(lay) <= C, "cylinder.max_num too small for this scenario"
        for k, (x, y) in enumerate(lay):
            cyl[:, k, 0], cyl[:, k, 1], cyl[:, k, 2] = x * cs, y * cs, 0.5 * ch
    return dict(drone_pos=dpos, drone_rot=rot, target_pos=tpos, cyl_pos=cyl)
