"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the device-side reset sampler
`hs_sample_reset`.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu legs may import this file; the product path never does.

What it restates
----------------
The reference draws the initial configuration of a random-cylinder episode in
`HideAndSeek._reset_idx` (omni_drones/envs/hide_and_seek/hideandseek.py:609-697):

* pursuer xy ~ U([0.1, a-0.1] x [-a+0.1, a-0.1]), evader xy ~ U([-a+0.1, -0.1] x [-a+0.1, a-0.1]),
  a = arena_size / sqrt(2)                                  (hideandseek.py:283-290, 616-617)
* z ~ U(max_height/2 -+ 0.1) for every body                 (hideandseek.py:291-298, 628-629)
* rpy ~ U([-.2pi,-.2pi,0], [.2pi,.2pi,.2pi]) -> quaternion   (hideandseek.py:300-303, 696-697)
* cylinders: a num_grid x num_grid occupancy grid of cell 2*cylinder.size; cells at integer
  distance >= num_grid//2 from the centre and the cells of the pursuers / the evader are
  occupied (hideandseek.py:576-592, 168-181, 144-166); `max_num` DISTINCT free cells are drawn
  uniformly without replacement (hideandseek.py:106-119: randperm prefix, one env at a time
  on the host); the number of active cylinders ~ U{min_cylinders..max_num} (hideandseek.py:595-598);
  inactive ones are parked at z = invalid_z (hideandseek.py:687-689); cell -> metres by
  grid_to_continuous (hideandseek.py:120-142).

The DISTRIBUTION above is the contract.  The random STREAM is new (the reference consumes
torch's global generator through a per-env host loop, which cannot be reproduced on the
device): a counter-based Philox4x32-10 generator keyed by `seed`, counter
(global env index, draw block, epoch lo, epoch hi), so that a draw depends only on
(seed, epoch, global env index) - independent of the batch size, the shard and the mask.
Parity bar: the CUDA kernel is bit-exact against this file for positions, cells and counts
(integer work + correctly rounded fp32 mul/add), 1e-6 absolute for the quaternion (sinf/cosf).

Draw order per env (32-bit words of the Philox stream):
  [0, 2A)        pursuer x, y (agent-major)
  [2A, 2A+2)     evader x, y
  [2A+2, 3A+2)   pursuer z
  [3A+2]         evader z
  [3A+3]         number of active cylinders
  [3A+4, 3A+4+C) cylinder cells (k-th draw picks the r-th still-free cell, r = mulhi(u32, free-k))
  [3A+4+C, 6A+4+C) roll, pitch, yaw (agent-major)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; Random123).  ctr [...,4] uint32, key [...,2] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


@dataclass
class ResetDist:
    """Mirror of include/hs_b200.h::hs_reset_dist (same field meaning)."""
    num_agents: int = 3
    num_cylinders: int = 5
    drone_lo: List[float] = field(default_factory=lambda: [0.0, 0.0])
    drone_hi: List[float] = field(default_factory=lambda: [0.0, 0.0])
    target_lo: List[float] = field(default_factory=lambda: [0.0, 0.0])
    target_hi: List[float] = field(default_factory=lambda: [0.0, 0.0])
    z_lo: float = 0.4
    z_hi: float = 0.6
    rpy_lo: List[float] = field(default_factory=lambda: [-0.2 * math.pi, -0.2 * math.pi, 0.0])
    rpy_hi: List[float] = field(default_factory=lambda: [0.2 * math.pi, 0.2 * math.pi, 0.2 * math.pi])
    grid_size: float = 0.2
    num_grid: int = 9
    boundary: float = 0.8
    cyl_z_active: float = 0.5
    cyl_z_inactive: float = -20.0
    min_cylinders: int = 0
    fixed_num: int = -1
    fixed_xy: int = 0
    fixed_drone_xy: Optional[np.ndarray] = None      # [A,2] when fixed_xy
    fixed_target_xy: Optional[np.ndarray] = None     # [2]
    env_offset: int = 0
    seed: int = 0

    @staticmethod
    def for_task(arena_size=0.9, cylinder_size=0.1, max_height=1.0, cylinder_height=1.0, num_agents=3,
                 num_cylinders=5, min_cylinders=0, fixed_num=-1, invalid_z=-20.0, seed=0, env_offset=0,
                 use_eval=False) -> "ResetDist":
        a = arena_size / math.sqrt(2.0)
        gs = 2 * cylinder_size
        d = ResetDist(num_agents=num_agents, num_cylinders=num_cylinders,
                      drone_lo=[0.1, -a + 0.1], drone_hi=[a - 0.1, a - 0.1],
                      target_lo=[-a + 0.1, -a + 0.1], target_hi=[-0.1, a - 0.1],
                      z_lo=max_height / 2 - 0.1, z_hi=max_height / 2 + 0.1,
                      grid_size=gs, num_grid=int(arena_size * 2 / gs), boundary=arena_size - 0.1,
                      cyl_z_active=0.5 * cylinder_height, cyl_z_inactive=invalid_z,
                      min_cylinders=min_cylinders, fixed_num=fixed_num, seed=seed, env_offset=env_offset)
        if use_eval:
            d.fixed_xy = 1
            d.fixed_drone_xy = np.array([[0.6, 0.0], [0.8, 0.0], [0.8, -0.2], [0.8, 0.2]], np.float32)[:num_agents]
            d.fixed_target_xy = np.array([-0.8, 0.0], np.float32)
            d.rpy_lo = [0.0, 0.0, 0.0]
            d.rpy_hi = [0.0, 0.0, 0.0]
        return d


def free_cells_upper_bound(num_grid: int) -> int:
    """Cells inside the circle (hideandseek.py:168-181)."""
    h = num_grid // 2
    return sum(1 for i in range(num_grid) for j in range(num_grid) if (i - h) ** 2 + (j - h) ** 2 < h * h)


def _u01(x: np.ndarray) -> np.ndarray:
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def _uniform(x: np.ndarray, lo: float, hi: float) -> np.ndarray:
    lo32, hi32 = np.float32(lo), np.float32(hi)
    return (lo32 + (hi32 - lo32) * _u01(x)).astype(np.float32)      # fp32 sub, mul, add: each rounded once


def occupancy(d: ResetDist, dxy: np.ndarray, txy: np.ndarray) -> np.ndarray:
    """grid_map of rejection_sampling_random_cylinder (hideandseek.py:576-592): True = occupied.
    dxy [E,A,2], txy [E,2]."""
    E, ng = dxy.shape[0], d.num_grid
    half = ng // 2
    ii, jj = np.meshgrid(np.arange(ng), np.arange(ng), indexing="ij")
    occ = np.broadcast_to(((ii - half) ** 2 + (jj - half) ** 2 >= half * half)[None], (E, ng, ng)).copy()

    def cell(xy):      # continuous_to_grid, hideandseek.py:144-166 (round half to even, IEEE divide)
        q = np.rint((xy.astype(np.float32) / np.float32(d.grid_size)).astype(np.float32)).astype(np.int64) + half
        return np.clip(q, 0, ng - 1)
    ar = np.arange(E)
    dc, tc = cell(dxy), cell(txy)
    for a in range(dxy.shape[1]):
        occ[ar, dc[:, a, 0], dc[:, a, 1]] = True
    occ[ar, tc[:, 0], tc[:, 1]] = True
    return occ


def cell_to_xy(d: ResetDist, cells: np.ndarray) -> np.ndarray:
    """grid_to_continuous (hideandseek.py:120-142) of flat cell indices x * num_grid + y."""
    ng = d.num_grid
    half = ng // 2
    cxy = np.stack([cells // ng, cells % ng], -1).astype(np.float32)
    cxy = ((cxy - np.float32(half)) * np.float32(d.grid_size)).astype(np.float32)
    return np.clip(cxy, np.float32(-d.boundary), np.float32(d.boundary))


def sample_reset(d: ResetDist, num_envs: int, epoch: int) -> Dict[str, np.ndarray]:
    A, C, E, ng = d.num_agents, d.num_cylinders, num_envs, d.num_grid
    ndraw = 6 * A + 4 + C
    nblk = (ndraw + 3) // 4
    env = (np.arange(E, dtype=np.uint64) + np.uint64(d.env_offset)).astype(np.uint32)
    ctr = np.zeros((E, nblk, 4), np.uint32)
    ctr[..., 0] = env[:, None]
    ctr[..., 1] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 2] = np.uint32(epoch & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((epoch >> 32) & 0xFFFFFFFF)
    key = np.zeros((E, nblk, 2), np.uint32)
    key[..., 0] = np.uint32(d.seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((d.seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, key).reshape(E, nblk * 4)

    o = 0
    dxy = np.zeros((E, A, 2), np.float32)
    for a in range(A):
        for c in range(2):
            dxy[:, a, c] = _uniform(w[:, o], d.drone_lo[c], d.drone_hi[c]); o += 1
    txy = np.zeros((E, 2), np.float32)
    for c in range(2):
        txy[:, c] = _uniform(w[:, o], d.target_lo[c], d.target_hi[c]); o += 1
    if d.fixed_xy:
        dxy[:] = np.asarray(d.fixed_drone_xy, np.float32)[None]
        txy[:] = np.asarray(d.fixed_target_xy, np.float32)[None]
    dz = np.zeros((E, A), np.float32)
    for a in range(A):
        dz[:, a] = _uniform(w[:, o], d.z_lo, d.z_hi); o += 1
    tz = _uniform(w[:, o], d.z_lo, d.z_hi); o += 1
    if d.fixed_num >= 0:
        n_active = np.full(E, d.fixed_num, np.int64)
    else:
        span = np.uint64(C + 1 - d.min_cylinders)
        n_active = d.min_cylinders + ((w[:, o].astype(np.uint64) * span) >> np.uint64(32)).astype(np.int64)
    o += 1

    half = ng // 2
    ar = np.arange(E)
    occ = occupancy(d, dxy, txy)
    free = ~occ.reshape(E, ng * ng)
    nfree = free.sum(-1)
    assert (nfree >= C).all(), "Not enough available coordinates (hideandseek.py:111-112)"
    picks = np.zeros((E, C), np.int64)
    for k in range(C):
        r = ((w[:, o].astype(np.uint64) * (nfree - k).astype(np.uint64)) >> np.uint64(32)).astype(np.int64); o += 1
        rank = np.cumsum(free, axis=-1) - 1                          # rank of each free cell, ascending index
        hit = free & (rank == r[:, None])
        idx = hit.argmax(-1)
        picks[:, k] = idx
        free[ar, idx] = False
    cxy = cell_to_xy(d, picks)
    cz = np.where(np.arange(C)[None] >= n_active[:, None], np.float32(d.cyl_z_inactive), np.float32(d.cyl_z_active))
    cyl = np.concatenate([cxy, cz[..., None].astype(np.float32)], -1)

    rpy = np.zeros((E, A, 3), np.float32)
    for a in range(A):
        for c in range(3):
            rpy[:, a, c] = _uniform(w[:, o], d.rpy_lo[c], d.rpy_hi[c]); o += 1
    assert o == ndraw
    h = rpy.astype(np.float64) * 0.5      # euler_to_quaternion, omni_drones/utils/torch.py (wxyz)
    cr, sr, cp, sp, cy, sy = np.cos(h[..., 0]), np.sin(h[..., 0]), np.cos(h[..., 1]), np.sin(h[..., 1]), \
        np.cos(h[..., 2]), np.sin(h[..., 2])
    rot = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                    cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], -1).astype(np.float32)
    return dict(drone_pos=np.concatenate([dxy, dz[..., None]], -1), drone_rot=rot,
                target_pos=np.concatenate([txy, tz[:, None]], -1), cyl_pos=cyl,
                active_cylinders=n_active.astype(np.float32)[:, None], cells=picks, rpy=rpy, occ=occ)
