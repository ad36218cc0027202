"""CPU restatement of the Hover task's observation / reward / stats (omni_drones/envs/single/hover.py:334-523).

TEST INFRASTRUCTURE ONLY (see oracle/hs_oracle.py).  The vehicle tick is hs_oracle's (one pursuer); this module restates
what hs_hover_post computes from the post-tick drone state.  Pinned by tests/golden/hover.npz, which
oracle/gen_hover_golden.py produces by executing the reference's own Hover methods (oracle/ref_harness.RefHover).
"""
import torch

from . import hs_oracle as O

STAT_KEYS = ("return", "pos_bonus", "head_bonus", "reward_pos", "reward_up", "reward_vel", "reward_acc", "reward_jerk",
             "episode_len", "pos_error", "heading_alignment", "uprightness", "action_smoothness",
             "linear_v_max", "angular_v_max", "linear_a_max", "angular_a_max", "linear_jerk_max", "angular_jerk_max",
             "linear_v_mean", "angular_v_mean", "linear_a_mean", "angular_a_mean", "linear_jerk_mean", "angular_jerk_mean",
             "motor1", "motor2", "motor3", "motor4", "cmd_r", "cmd_p", "cmd_y", "cmd_thrust",
             "target_r_rate", "target_p_rate", "target_y_rate", "real_r_rate", "real_p_rate", "real_y_rate")
S = {k: i for i, k in enumerate(STAT_KEYS)}


def hover_post(drone_state, progress, stats, state, target_heading, cmds=None, ctbr=None, target_rate=None, throttle_diff=None,
               dt=0.01, max_episode_length=500, with_reward=True, reward_distance_scale=10.0, reward_v_scale=0.0,
               reward_acc_scale=0.0, reward_jerk_scale=0.0, linear_vel_max=3.0, linear_acc_max=10.0, alpha=0.8):
    """drone_state [E,13] (post tick), progress [E], stats [E,39] and state [E,12] (last lin/ang v, a, jerk; six episode sums)
    are updated in place; returns (observation [E,20], reward [E], done [E])."""
    p, q, lv, av = drone_state[:, :3], drone_state[:, 3:7], drone_state[:, 7:10], drone_state[:, 10:13]
    if with_reward:                                            # _pre_sim_step logging, hover.py:334-359
        stats[:, S["motor1"]:S["motor4"] + 1] = cmds
        stats[:, S["cmd_r"]:S["cmd_thrust"] + 1] = ctbr
        stats[:, S["target_r_rate"]:S["target_y_rate"] + 1] = target_rate
    stats[:, S["real_r_rate"]:S["real_y_rate"] + 1] = O.quat_apply_inverse(q, av) * 180.0 / torch.pi
    heading, up = O.quat_basis(q, 0), O.quat_basis(q, 2)
    rpos = torch.tensor([0.0, 0.0, 1.0]) - p
    rheading = target_heading - heading
    t = (progress / max_episode_length).unsqueeze(-1).expand(-1, 4)
    obs = torch.cat([rpos, q, lv, heading, up, t], dim=-1)
    lin_v, ang_v = torch.linalg.vector_norm(lv, dim=-1), torch.linalg.vector_norm(av, dim=-1)
    lin_a, ang_a = (lin_v - state[:, 0]).abs() / dt, (ang_v - state[:, 1]).abs() / dt
    lin_j, ang_j = (lin_a - state[:, 2]).abs() / dt, (ang_a - state[:, 3]).abs() / dt
    for i, v in enumerate((lin_v, ang_v, lin_a, ang_a, lin_j, ang_j)):
        stats[:, S["linear_v_max"] + i] = torch.max(stats[:, S["linear_v_max"] + i], v.abs())
        state[:, 6 + i] += v.abs()
        stats[:, S["linear_v_mean"] + i] = state[:, 6 + i] / (progress + 1.0)
        state[:, i] = v
    if not with_reward:
        return obs, None, None
    pos_error, head_error = torch.linalg.vector_norm(rpos, dim=-1), torch.linalg.vector_norm(rheading, dim=-1)
    reward_pos = -pos_error * reward_distance_scale
    bonus = (pos_error <= 0.02).float() * 10
    near = (bonus > 0).float()
    reward_head = -head_error * near
    head_bonus = (head_error <= 0.02).float() * 10 * near
    reward_up = torch.square((up[:, 2] + 1) / 2)
    reward_v = reward_v_scale * near * (lin_v < linear_vel_max).float()
    reward_acc = reward_acc_scale * near * (lin_a < linear_acc_max).float()
    reward_jerk = reward_jerk_scale * near * (-lin_j)
    reward = reward_pos + bonus + reward_head + head_bonus + reward_up + reward_v + reward_acc + reward_jerk
    w = 1 - alpha
    for k, x in (("pos_error", pos_error), ("heading_alignment", (heading * target_heading).sum(-1)), ("uprightness", up[:, 2]),
                 ("action_smoothness", -throttle_diff)):
        stats[:, S[k]] = stats[:, S[k]] + w * (x - stats[:, S[k]])
    stats[:, S["return"]] += reward
    for k, x in (("reward_pos", reward_pos), ("pos_bonus", bonus), ("head_bonus", head_bonus), ("reward_vel", reward_v),
                 ("reward_acc", reward_acc), ("reward_jerk", reward_jerk), ("episode_len", progress)):
        stats[:, S[k]] = x
    return obs, reward, progress >= max_episode_length
