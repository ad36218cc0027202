"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the two device kernels of the
HideAndSeek_envgen control plane.  Only tests/ may import this file.

What it restates
----------------
* `fps(points, k, start)` - greedy farthest point sampling, the recurrence behind
  `dgl.geometry.farthest_point_sampler` that GenBuffer.insert_history calls to cap the archive at
  5000 tasks (omni_drones/envs/hide_and_seek/hideandseek_envgen.py:300-314; DGL is a third-party
  dependency absent from /root/reference and from this image).  Squared distances are accumulated
  dimension by dimension in fp32 with separately rounded multiply and add, the running minimum is
  kept per point, and the next point is the FIRST index attaining the maximum - the CUDA kernel
  hs_fps follows exactly this order, so the selected indices are compared bit for bit.
  (DGL starts from a random point unless start_idx is given; here the start is an argument.)
* `sample_nearby(...)` - GenBuffer.samplenearby (hideandseek_envgen.py:322-372): pick an archive
  task uniformly, add U(-1,1)*expand_step noise to the pursuer/evader coordinates and, with
  expand_cylinders, one grid step of {-1,0,1} to every cylinder's x and y, clip to the task
  bounds, accept when every object sits on its own free cell of the occupancy grid
  (sanity_check :185-207), retry up to 10 times.  The reference does this in a Python loop over
  tasks with numpy's global generator in float64; the device version is fp32 with the same
  counter-based Philox stream as the reset sampler (counter = (task, block + 64*attempt, epoch),
  key = seed ^ stream tag), so only the distribution is the contract.  A task that fails ten times
  is flagged invalid (the caller re-draws it from the valid ones, :361-366).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from oracle.reset_sampler import philox4x32_10

GEN_STREAM_TAG = 0x9E3779B97F4A7C15        # xor-ed into the seed: keeps this stream apart from the reset sampler's


def fps(points: np.ndarray, k: int, start: int = 0) -> np.ndarray:
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, dim = pts.shape
    k = min(k, n)
    out = np.zeros(k, np.int64)
    mind = np.full(n, np.inf, np.float32)
    cur = start
    for i in range(k):
        out[i] = cur
        d = np.zeros(n, np.float32)
        q = pts[cur]
        for j in range(dim):
            diff = (pts[:, j] - q[j]).astype(np.float32)
            d = (d + (diff * diff).astype(np.float32)).astype(np.float32)
        mind = np.minimum(mind, d)
        cur = int(np.argmax(mind))           # first index of the maximum
    return out


def task_bounds(A: int, C: int, arena_size: float, grid_size: float, max_height: float) -> np.ndarray:
    """hideandseek_envgen.py:327-340 (the z range of the bodies is [max_height-0.1, max_height+0.1] there - sic)."""
    cb = int(arena_size / grid_size) * grid_size
    bxy = arena_size / math.sqrt(2.0) - 0.1
    drone_b = [[-bxy, bxy], [-bxy, bxy], [max_height - 0.1, max_height + 0.1]]
    cyl_b = [[-cb, cb], [-cb, cb], [-20.0, max_height / 2]]
    return np.array(drone_b * (A + 1) + cyl_b * C, np.float32)


def inside_mask(ng: int) -> np.ndarray:
    half = ng // 2
    ii, jj = np.meshgrid(np.arange(ng), np.arange(ng), indexing="ij")
    return (ii - half) ** 2 + (jj - half) ** 2 < half * half


def task_cells(task: np.ndarray, A: int, C: int, grid, ng: int) -> np.ndarray:
    """continuous_to_grid (hideandseek_envgen.py:140-163) of every object's xy: [A+1+C, 2]."""
    nb = 3 * A + 3
    t = np.asarray(task, np.float32)
    xy = np.concatenate([t[:nb].reshape(-1, 3)[:, :2], t[nb:].reshape(-1, 3)[:, :2]])
    return np.clip(np.rint((xy / np.float32(grid)).astype(np.float32)).astype(np.int64) + ng // 2, 0, ng - 1)


def sanity_ok(task: np.ndarray, A: int, C: int, grid, ng: int, inside: np.ndarray) -> bool:
    """sanity_check (hideandseek_envgen.py:185-207): every object claims its own free cell."""
    free = inside.copy()
    for cx, cy in task_cells(task, A, C, grid, ng):
        if not free[cx, cy]:
            return False
        free[cx, cy] = False
    return True


def sample_nearby(history: np.ndarray, num_tasks: int, A: int, C: int, arena_size: float, cylinder_size: float,
                  max_height: float, expand_cylinders: bool, expand_step: float, seed: int, epoch: int,
                  task_offset: int = 0) -> Dict[str, np.ndarray]:
    hist = np.ascontiguousarray(history, np.float32)
    n_hist, dim = hist.shape
    assert dim == 3 * A + 3 + 3 * C
    grid = np.float32(2 * cylinder_size)
    ng = int(arena_size * 2 / (2 * cylinder_size))
    half = ng // 2
    bounds = task_bounds(A, C, arena_size, 2 * cylinder_size, max_height)
    nb = 3 * A + 3
    words_per_attempt = nb + 2 * C
    blocks = (words_per_attempt + 3) // 4
    key64 = (seed ^ GEN_STREAM_TAG) & (2 ** 64 - 1)
    key = np.array([key64 & 0xFFFFFFFF, key64 >> 32], np.uint32)
    inside = inside_mask(ng)
    out = np.zeros((num_tasks, dim), np.float32)
    valid = np.zeros(num_tasks, np.uint8)
    origin_idx = np.zeros(num_tasks, np.int64)
    step32 = np.float32(expand_step)
    for t in range(num_tasks):
        def words(block0, nblk):
            ctr = np.zeros((nblk, 4), np.uint32)
            ctr[:, 0] = np.uint32((t + task_offset) & 0xFFFFFFFF)
            ctr[:, 1] = np.arange(block0, block0 + nblk, dtype=np.uint32)
            ctr[:, 2] = np.uint32(epoch & 0xFFFFFFFF)
            ctr[:, 3] = np.uint32((epoch >> 32) & 0xFFFFFFFF)
            return philox4x32_10(ctr, np.broadcast_to(key, (nblk, 2))).reshape(-1)
        w0 = words(0xFFFF0000, 1)                       # the archive pick has its own block
        idx = int((int(w0[0]) * n_hist) >> 32)
        origin_idx[t] = idx
        origin = hist[idx]
        cand = origin.copy()
        for attempt in range(10):
            w = words(64 * attempt, blocks)
            cand = origin.copy()
            o = 0
            for j in range(nb):
                u = np.float32(w[o] >> np.uint32(8)) * np.float32(2.0 ** -24); o += 1
                noise = (np.float32(-1.0) + np.float32(2.0) * u) * step32
                cand[j] = origin[j] + np.float32(noise)
            for c in range(C):
                for a in range(2):
                    s = int((int(w[o]) * 3) >> 32) - 1; o += 1
                    if expand_cylinders:
                        cand[nb + 3 * c + a] = origin[nb + 3 * c + a] + np.float32(s) * grid
            cand = np.minimum(np.maximum(cand, bounds[:, 0]), bounds[:, 1]).astype(np.float32)
            ok = sanity_ok(cand, A, C, grid, ng, inside)
            if ok:
                valid[t] = 1
                break
        out[t] = cand
    return dict(tasks=out, valid=valid, origin=origin_idx)
