"""Conditioning of one HideAndSeek tick: WHERE may two correct fp32 implementations disagree by more than 1e-4?

TEST INFRASTRUCTURE ONLY (see oracle/hs_oracle.py header).  The parity tests compare the CUDA kernels with the oracle /
the reference-run fixtures element by element at 1e-4 relative.  Two mechanisms of the TASK ITSELF amplify rounding-level
(1e-7) differences beyond that bar; this module computes, from the oracle's own numbers, exactly which environments they
touch and by how much, so that the tests need no blanket "fraction of elements may be wrong" budget:

1. Evader velocity.  v = v_prey * f / (|f| + 1e-5) per component (reference quirk, hideandseek.py:741) has slope
   v_prey * 1e-5 / (|f| + 1e-5)^2 -- up to 1.3e5 at f = 0 -- and f is a sum of repulsion terms that cancel
   (hideandseek.py:1067-1141).  With M = sum of |terms| and eps the relative rounding level of the implementation, the
   error of f is bounded by eps * M and the error of v by  dv = v_prey * 1e-5 * eps * M / (|f| + 1e-5)^2  (capped at
   2 v_prey).  `evader_bound` returns dv per env and component; every tensor that carries the evader's velocity or
   position gets this much extra absolute tolerance FOR THAT ENV (positions: dt * dv), nothing else does.

2. Rate-PID derivative.  ctbr[0:2] = out/2 with out = P + D + I and D = -(rate - last_rate) / dt * kd in deg/s
   (lee_position_controller.py:505-515): a rounding difference d in the body rate (the reference forms it with a bmm
   dot product whose accumulation order / FMA use is BLAS-internal) appears as d * 180/pi * kd / dt / 2 ~ 7e3 * d in
   ctbr.  `pid_bound` returns ulps(body rate) times that gain per (env, pursuer, component); only `ctbr` gets it (the
   rotor commands divide it by 2^15 again).

3. Indicators.  Rewards, masks and the k-nearest selection compare a continuous quantity with a threshold
   (capture |p - t| < 0.3, collisions, line of sight, arena wall, sort order of the cylinder distances ...).  An env is
   an "edge" env when one of these quantities lies within `margin` (relative) of its threshold in the oracle; only such
   envs may show a flipped indicator, and the tests bound how many there are.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import hs_oracle as O


def evader_bound(P: O.HSParams, v_prey: float, pos, tpos, cyl, eps: float) -> torch.Tensor:
    """[E,3]: bound on |delta tvel| caused by relative rounding `eps` in the terms of the evader's force.
    pos/tpos/cyl: the PRE-tick positions the force is computed from (hideandseek.py:737-744)."""
    inactive = cyl[..., 2] < 0.0
    force, mag, _ = O.evader_force(P, pos, tpos, cyl, inactive)
    slope = v_prey * 1e-5 / (force.abs() + 1e-5) ** 2
    return torch.clamp(slope * eps * mag, max=2.0 * v_prey)


def pid_bound(P: O.HSParams, angvel, ulps_rate: float = 8.0, ulps_tanh: float = 4.0) -> torch.Tensor:
    """[E,A,4]: bound on |delta ctbr| (r, p, y, thrust).  out = P + D + I in 16-bit motor units per deg/s:
    * the target rate is tanh(raw) * 180 * target_clip: CUDA's tanhf (<= 2 ulp) and torch's (Sleef, <= 1 ulp) differ by a
      few ulp of |tanh| <= 1 -> d_target = ulps_tanh * 6e-8 * 180 deg/s, amplified by kp (250, 250, 120);
    * the body rate a - b + c (terms up to 3 |w|_inf) carries ulps_rate ulp of rounding (the reference's dot product is a
      bmm whose accumulation order is BLAS-internal), amplified by kp and by kd / dt (250 / 0.01 s).
    r = out0 / 2, p = out1 / 2, y = out2; the thrust slot is exact.  angvel: PRE-tick world angular velocity [E,A,3]."""
    w = angvel.abs().max(-1, keepdim=True).values
    d_rate = ulps_rate * 6e-8 * 3.0 * w * (180.0 / torch.pi)                    # deg/s
    d_target = ulps_tanh * 6e-8 * 180.0 * P.target_clip
    kp = torch.tensor(P.pid_kp, dtype=torch.float32)
    kd = torch.tensor(P.pid_kd, dtype=torch.float32)
    out = (d_rate + d_target) * kp + d_rate * kd / P.dt                          # [E,A,3]
    scale = torch.tensor([0.5, 0.5, 1.0])
    return torch.cat([out * scale, torch.zeros_like(w)], dim=-1)


def _near(x, thr, margin):
    return (x - thr).abs() <= margin * max(abs(thr), 1.0) if not torch.is_tensor(thr) else \
        (x - thr).abs() <= margin * torch.clamp(thr.abs(), min=1.0)


def los_edge(P: O.HSParams, pos, tpos, cyl, margin: float) -> torch.Tensor:
    """[E] bool: some (pursuer, standing cylinder) pair has one of the three predicates of
    is_line_blocked_by_cylinder (hideandseek.py:47-103) within `margin` of flipping."""
    d = pos - tpos.unsqueeze(1)
    c = cyl - tpos.unsqueeze(1)
    cross = torch.abs(d[..., 0:1] * c[..., 1].unsqueeze(1) - d[..., 1:2] * c[..., 0].unsqueeze(1))
    seg = torch.sqrt(d[..., 0:1] ** 2 + d[..., 1:2] ** 2)
    dist_line = cross / (seg + 1e-5)
    dx = tpos[:, None, 0] - pos[..., 0]
    dy = tpos[:, None, 1] - pos[..., 1]
    num = (cyl[:, None, :, 0] - pos[..., 0:1]) * dx.unsqueeze(2) + (cyl[:, None, :, 1] - pos[..., 1:2]) * dy.unsqueeze(2)
    den = dx.unsqueeze(2) ** 2 + dy.unsqueeze(2) ** 2
    t = num / (den + 1e-5)
    standing = (cyl[..., 2] > 0.0).unsqueeze(1)
    edge = ((dist_line - P.cylinder_size).abs() <= margin) | (t.abs() <= margin) | ((t - 1.0).abs() <= margin)
    return (edge & standing).flatten(1).any(-1)


def indicator_edges(P: O.HSParams, pre: Dict[str, torch.Tensor], post: Dict[str, torch.Tensor], margin: float = 2e-5
                    ) -> Dict[str, torch.Tensor]:
    """name -> [E] bool for every thresholded quantity of the tick.  pre/post: the oracle's state dicts (pos, linvel,
    tpos, cyl) before / after the tick: the evader's policy reads the PRE state, observation and reward the POST state."""
    out = {}
    E, A, _ = post["pos"].shape
    pos, tpos, cyl, lv = post["pos"], post["tpos"], post["cyl"], post["linvel"]
    out["los_prey"] = los_edge(P, pre["pos"], pre["tpos"], pre["cyl"], margin)
    out["los_obs"] = los_edge(P, pos, tpos, cyl, margin)
    dist = torch.linalg.vector_norm(tpos.unsqueeze(1) - pos, dim=-1)
    out["capture"] = _near(dist, P.catch_radius, margin).any(-1)
    # the integrator clamps |v| to v_max * (1 - 1e-6), a hair below the penalty threshold (DESIGN.md "Integrator"):
    # only speeds within 4e-7 of the threshold can flip
    out["speed"] = ((torch.linalg.vector_norm(lv, dim=-1) - P.v_drone).abs() <= 4e-7).any(-1)
    # cylinders: sort key (3-D distance to the centre) gaps and the 2-D collision test, hideandseek.py:757-778, 961-968
    rpos = pos.unsqueeze(2) - cyl.unsqueeze(1)
    key = torch.linalg.vector_norm(rpos, dim=-1) - P.cylinder_size
    C = cyl.shape[1]
    if C > 1 and P.obs_max_cylinder > 0:
        sk, order = torch.sort(key, dim=-1)
        k = min(P.obs_max_cylinder, C - 1)
        gaps = (sk[..., 1:k + 1] - sk[..., :k]).abs()
        down = (cyl[..., 2] < 0.0).unsqueeze(1).expand(-1, A, -1).gather(2, order)          # inactive, in sorted order
        both_down = down[..., 1:k + 1] & down[..., :k]      # two inactive neighbours: both rows are masked to -5 either way
        out["knearest_order"] = ((gaps <= margin * torch.clamp(sk[..., :k].abs(), min=1.0)) & ~both_down).flatten(1).any(-1)
    if C > 0:
        dxy = torch.linalg.vector_norm(rpos[..., :2], dim=-1)
        standing = (cyl[..., 2] >= 0.0).unsqueeze(1)
        out["hit_cylinder"] = (_near(dxy - P.cylinder_size, P.collision_radius, margin) & standing).flatten(1).any(-1)
    if A > 1:
        dd = torch.linalg.vector_norm(pos.unsqueeze(2) - pos.unsqueeze(1), dim=-1)
        off = ~torch.eye(A, dtype=torch.bool)
        out["hit_drone"] = (_near(dd, 2.0 * P.collision_radius, margin) & off).flatten(1).any(-1)
    out["wall"] = (_near(pos[..., 2], P.max_height, margin) |
                   _near(pos[..., 0] ** 2 + pos[..., 1] ** 2, P.arena_size ** 2, margin)).any(-1)
    out["ground"] = (pos[..., 2] <= P.ground_z + margin).any(-1) if P.ground_clamp else torch.zeros(E, dtype=torch.bool)
    pt = pre["tpos"]
    out["evader_bounds"] = (_near(pt[:, 0] ** 2 + pt[:, 1] ** 2, P.arena_size ** 2, margin) |
                            _near(pt[:, 2], P.max_height, margin) | (pt[:, 2].abs() <= margin))
    return out


# how much of the evader-velocity bound a tensor's elements inherit: "vel" = the velocity itself (tvel, the TP frame),
# "pos" = a position advanced by dt * v (tpos, relative positions in state_self / state_drones, the TP ground truth =
# position / (0.5 arena) resp. * 2 / max_height, the distance terms of reward / return), None = nothing
def kind_scale(P: O.HSParams, kind) -> float:
    if kind == "vel":
        return 1.0
    if kind == "pos":
        return P.dt * max(1.0, 1.0 / (0.5 * P.arena_size), 2.0 / P.max_height)
    return 0.0


KIND = {"tvel": "vel", "tp_input": "vel", "tpos": "pos", "state_self": "pos", "state_drones": "pos", "tp_groundtruth": "pos",
        "reward": "pos", "stats": "pos", "distance_reward": "pos", "return": "pos"}


class TickConditioning:
    """Per-env extra tolerance / edge flags of ONE tick, computed from the oracle's pre and post state."""

    def __init__(self, P: O.HSParams, v_prey: float, pre, post, eps: float = 4e-6, margin: float = 2e-5, safety: float = 4.0):
        self.P = P
        self.dv = safety * evader_bound(P, v_prey, pre["pos"], pre["tpos"], pre["cyl"], eps).max(-1).values   # [E]
        self.pid = pid_bound(P, pre["angvel"]) if "angvel" in pre else None                                    # [E,A,4]
        self.edges = indicator_edges(P, pre, post, margin)
        self.edge = torch.stack(list(self.edges.values()), 0).any(0)                                          # [E]

    def check(self, name, got, want, rtol=1e-4, atol=1e-5, cap=1e3, kind="auto"):
        """Elementwise |got - want| <= atol + rtol |want| + scale(kind) * dv[env]; elements of edge envs are only
        held to the sanity cap.  got/want: [E, ...]; kind: "vel" / "pos" / None, default by the last component of
        `name` (KIND).  Returns (elements that needed the edge exemption, elements that needed the dv allowance)."""
        if kind == "auto":
            kind = KIND.get(name.split("/")[-1])
        got = got.detach().cpu().float()
        want = want.detach().cpu().float().reshape(got.shape)
        E = got.shape[0]
        shape = (E,) + (1,) * (got.dim() - 1)
        base = atol + rtol * want.abs()
        tol = base + kind_scale(self.P, kind) * self.dv.reshape(shape)
        if name.split("/")[-1] == "ctbr" and self.pid is not None:
            tol = tol + self.pid.reshape(got.shape)
        err = (got - want).abs()
        bad = err > tol
        hard = bad & ~self.edge.reshape(shape)
        if hard.any():
            i = int(torch.argmax((err - tol).masked_fill(~hard, -1.0).flatten()))
            e = i // max(1, got[0].numel())
            raise AssertionError(
                f"{name}: {int(hard.sum())}/{bad.numel()} elements outside tolerance in well-conditioned envs; worst env {e} "
                f"|err|={err.flatten()[i].item():.3e} tol={tol.flatten()[i].item():.3e} got={got.flatten()[i].item():.6e} "
                f"want={want.flatten()[i].item():.6e} (dv bound of that env {self.dv[e].item():.2e})")
        if (err.masked_fill(~bad, 0.0) > cap).any() or torch.isnan(got).any() or torch.isnan(want).any():
            i = int(torch.argmax(torch.nan_to_num(err, nan=float("inf")).flatten()))
            raise AssertionError(f"{name}: an edge-env element is beyond the sanity cap {cap} or NaN: flat index {i} of shape "
                                 f"{tuple(got.shape)} got={got.flatten()[i].item()} want={want.flatten()[i].item()} "
                                 f"(NaNs: got {int(torch.isnan(got).sum())}, want {int(torch.isnan(want).sum())})")
        n_edge, n_dv = int(bad.sum()), int(((err > base) & ~bad).sum())
        if n_edge or n_dv:
            self.used = getattr(self, "used", [])
            self.used.append((name, n_edge, n_dv))
        self.n_edge_exempt = getattr(self, "n_edge_exempt", 0) + n_edge
        self.n_dv_needed = getattr(self, "n_dv_needed", 0) + n_dv
        return n_edge, n_dv


class TrajectoryConditioning:
    """Free-running comparison (no teacher forcing).  Every env carries the running sum of its per-tick evader-velocity
    bounds (`cum_dv`); its elements get that much extra absolute tolerance (scaled by kind like in TickConditioning).
    An env stays CLEAN until an indicator sits at an edge or cum_dv exceeds `dv_budget`; from then on the two
    implementations may legitimately follow different trajectories (the per-component sign normalisation makes the task
    chaotic at those points) and the env is only counted.  Clean envs must agree within the (slowly growing) tolerance."""

    def __init__(self, P: O.HSParams, E: int, eps: float = 1e-6, margin: float = 2e-5, dv_budget: float = 1e-2):
        self.P, self.eps, self.margin, self.dv_budget = P, eps, margin, dv_budget
        self.clean = torch.ones(E, dtype=torch.bool)
        self.cum_dv = torch.zeros(E)
        self.ticks = 0
        self.first_unclean_tick = None

    def update(self, v_prey, pre, post):
        c = TickConditioning(self.P, v_prey, pre, post, eps=self.eps, margin=self.margin)
        self.cum_dv += c.dv
        self.clean &= ~c.edge & (self.cum_dv <= self.dv_budget)
        self.ticks += 1
        if self.first_unclean_tick is None and not bool(self.clean.all()):
            self.first_unclean_tick = self.ticks - 1
        return c

    def check(self, name, got, want, rtol=1e-4, atol=1e-5, growth=0.25, kind="auto"):
        """Comparison on the clean envs; atol and rtol grow by `growth` per elapsed tick (accumulating rounding of a
        free-running fp32 integration), plus the env's cum_dv allowance for evader-derived tensors."""
        if kind == "auto":
            kind = KIND.get(name.split("/")[-1].lower())
        got = got.detach().cpu().float()
        want = want.detach().cpu().float().reshape(got.shape)
        m = self.clean
        if not bool(m.any()):
            return 0
        g = 1.0 + growth * self.ticks
        shape = (int(m.sum()),) + (1,) * (got.dim() - 1)
        err = (got[m] - want[m]).abs()
        tol = atol * g + rtol * g * want[m].abs() + kind_scale(self.P, kind) * self.cum_dv[m].reshape(shape)
        if (err > tol).any():
            i = int(torch.argmax((err - tol).flatten()))
            raise AssertionError(f"{name}: {int((err > tol).sum())}/{err.numel()} elements of CLEAN envs outside tolerance after "
                                 f"{self.ticks} free-running ticks; worst |err|={err.flatten()[i].item():.3e} "
                                 f"tol={tol.flatten()[i].item():.3e}")
        return int(m.sum())
