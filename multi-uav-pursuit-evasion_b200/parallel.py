"""Env-sharded data parallelism: one process per GPU, each owning a contiguous slice of the
global env index space with its own arena, TP history and RNG offset.  The tick itself needs
no exchange.  Collectives (NCCL on GPUs, gloo in the CPU tests) happen only at
  * rollout boundaries  -- all_gather of per-env episode returns / success for logging
                           (north_star: "a single NCCL all-gather of episode returns per rollout");
  * episode boundaries  -- all_reduce of {sum(success), count} so that the evader-speed
                           curriculum gate uses the mean over ALL envs of the job, as the
                           reference's single process does (hideandseek.py:1012-1015).
The reference has no distributed code at all; this is new.
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    global_envs: int

    @property
    def bounds(self) -> Tuple[int, int]:
        """[lo, hi) of this rank's envs: sizes differ by at most one, low ranks take the remainder."""
        base, rem = divmod(self.global_envs, self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)

    @property
    def local_envs(self) -> int:
        lo, hi = self.bounds
        return hi - lo

    def seed(self, base_seed: int) -> int:
        return base_seed + 1000003 * self.rank


def current_shard(global_envs: int) -> Shard:
    if dist.is_available() and dist.is_initialized():
        return Shard(dist.get_rank(), dist.get_world_size(), global_envs)
    return Shard(0, 1, global_envs)


def gather_env_vector(local: torch.Tensor, shard: Optional[Shard] = None, group=None) -> torch.Tensor:
    """all_gather of a per-env vector [E_local] -> [E_global] in global env order (every rank
    gets the full vector).  Uneven shards are padded to the largest one for the collective."""
    local = local.reshape(-1).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.clone()
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of a [n_local, d] matrix whose row count differs per rank -> [sum n, d] in rank order on
    every rank (the envgen archive is replicated: every rank inserts the tasks ALL ranks evaluated).
    One all_gather of the counts (host sync: episode boundary only) and one of the padded rows."""
    local = local.contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.clone()
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    m = max(sizes)
    if m == 0:
        return local.clone()
    pad = torch.zeros(m, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:k] for b, k in zip(bufs, sizes)])


def global_mean(local: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over the envs of ALL ranks, computed on device without a host sync:
    all_reduce(SUM) of [sum, count]."""
    local = local.reshape(-1).float()
    acc = torch.stack([local.sum(), torch.tensor(float(local.numel()), device=local.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc[0] / acc[1]


def global_sum(values: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise sum over all ranks of a small device vector (one all_reduce, no host sync)."""
    out = values.clone().float()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def global_any(flag: torch.Tensor, group=None) -> torch.Tensor:
    f = flag.reshape(-1).any().float().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)
    return f[0] > 0


def curriculum_step(v_prey: torch.Tensor, done: torch.Tensor, success: torch.Tensor, group=None,
                    threshold: float = 0.98, step: float = 0.05, v_max: float = 1.3) -> torch.Tensor:
    """hideandseek.py:1012-1015 for a sharded job: when any env of the job finished an episode
    and the job-wide success rate is >= threshold, the evader speeds up.  Updates `v_prey`
    (a 1-element device tensor the kernels read) in place; returns it."""
    ok = global_any(done, group) & (global_mean(success, group) >= threshold)
    v_prey.copy_(torch.where(ok, torch.clamp(v_prey + step, max=v_max), v_prey))
    return v_prey
