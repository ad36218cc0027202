"""The caller's side of a rollout.

* :func:`compute_gae` - same name, argument order and result as the reference's
  ``omni_drones/learning/utils/gae.py::compute_gae`` (``[N, T, k]`` rewards/values, ``[N, T, 1]``
  dones, ``[N, k]`` bootstrap value), executed by ``hs_gae`` (one backward-scan kernel);
  ``normalize=True`` adds the batch normalisation of ``MAPPOPolicy.train_op``
  (``mappo.py:391-396``) in a second launch.  Strided ``[N, T]`` views of time-major
  ``[T, N, ...]`` storage are taken as they are (no copy).
* :class:`RolloutStorage` - preallocated time-major ``[T, E, ...]`` tensors the tick kernels
  write into directly (every tick's output set IS slot ``t`` of the rollout), replacing the
  collector's per-step clone + final stack (``omni_drones/utils/torchrl/collector.py:33-37`` on
  top of torchrl's ``SyncDataCollector.rollout``).  ``batch()`` hands the learner ``[E, T, ...]``
  views.

No CPU path: tensors must live on a CUDA device.
"""
import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib


def _col_view(x: torch.Tensor, name: str) -> Tuple[torch.Tensor, int, int, int]:
    """[N, T, *k] with contiguous trailing dims -> (tensor, k, stride_env, stride_step)."""
    if x.dim() < 2:
        raise _lib.HsError(f"compute_gae: {name} must be [N, T, ...], got {tuple(x.shape)}")
    k, inner = 1, 1
    for d in range(x.dim() - 1, 1, -1):
        if x.shape[d] != 1 and x.stride(d) != inner:
            raise _lib.HsError(f"compute_gae: the trailing dims of {name} must be contiguous (strides {x.stride()})")
        inner *= x.shape[d]
        k *= x.shape[d]
    return x, k, x.stride(0), x.stride(1)


def compute_gae(reward: torch.Tensor, done: torch.Tensor, value: torch.Tensor, next_value: torch.Tensor,
                gamma: float = 0.99, lmbda: float = 0.95, normalize: bool = False, return_stats: bool = False):
    """gae.py:27-51.  Returns (advantages, returns) shaped and strided like ``reward``; with
    ``return_stats`` also a 2-element device tensor {mean, std} of the un-normalised advantages."""
    if reward.device.type != "cuda":
        raise _lib.HsError("compute_gae runs on the GPU only (hs_gae); got tensors on " + str(reward.device))
    if reward.shape != value.shape:
        raise _lib.HsError(f"compute_gae: reward {tuple(reward.shape)} and value {tuple(value.shape)} differ")
    if reward.dtype != torch.float32 or value.dtype != torch.float32:
        raise _lib.HsError("compute_gae: reward and value must be float32")
    _, k, se, st = _col_view(reward, "reward")
    if value.stride() != reward.stride():
        value = value.contiguous() if reward.is_contiguous() else value.clone(memory_format=torch.preserve_format)
        if value.stride() != reward.stride():
            raise _lib.HsError("compute_gae: reward and value must share one memory layout")
    N, T = reward.shape[:2]
    # done: [N, T, 1] or [N, T, k, 1]-style broadcasts of the env-level flag (MAPPOPolicy._get_dones)
    d = done
    while d.dim() > 2:
        d = d.select(-1, 0) if d.dim() > 2 else d
    if d.shape != (N, T):
        raise _lib.HsError(f"compute_gae: done must broadcast from [N, T, 1], got {tuple(done.shape)}")
    if d.dtype == torch.bool:
        d = d.view(torch.uint8)
    elif d.dtype != torch.uint8:
        d = (d != 0).view(torch.uint8)
    nv = next_value.reshape(N, k).to(torch.float32).contiguous()
    adv = torch.empty_strided(reward.shape, reward.stride(), dtype=torch.float32, device=reward.device)
    ret = torch.empty_strided(reward.shape, reward.stride(), dtype=torch.float32, device=reward.device)
    scratch = torch.empty(2, dtype=torch.float64, device=reward.device)
    stats = torch.empty(2, dtype=torch.float32, device=reward.device) if (return_stats or normalize) else None
    p = _lib.hs_gae_params()
    p.num_envs, p.num_steps, p.num_agents = N, T, k
    p.stride_env, p.stride_step = se, st
    p.done_stride_env, p.done_stride_step = d.stride(0), d.stride(1)
    p.gamma, p.lmbda, p.normalize = float(gamma), float(lmbda), 1 if normalize else 0
    stream = torch.cuda.current_stream(reward.device).cuda_stream
    with torch.cuda.device(reward.device):
        check(lib.hs_gae(C.byref(p), reward.data_ptr(), d.data_ptr(), value.data_ptr(), nv.data_ptr(), adv.data_ptr(),
                         ret.data_ptr(), scratch.data_ptr(), stats.data_ptr() if stats is not None else None, stream),
              "hs_gae")
    return (adv, ret, stats) if return_stats else (adv, ret)
