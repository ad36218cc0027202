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


def _flat_items(td, prefix=()):
    """(key tuple, tensor) of every leaf of a (nested) dict / tensordict."""
    keys = td.keys() if hasattr(td, "keys") else []
    for k in keys:
        v = td[k] if not hasattr(td, "get") else td.get(k)
        kk = prefix + (k if isinstance(k, tuple) else (k,))
        if torch.is_tensor(v):
            yield kk, v
        else:
            yield from _flat_items(v, kk)


def gather_minibatch(batch, indices: torch.Tensor, num_steps: Optional[int] = None) -> Dict[tuple, torch.Tensor]:
    """``tensordict.reshape(-1)[indices]`` of make_dataset_naive (omni_drones/learning/mappo.py:493-513) for every leaf of
    ``batch`` - tensors shaped ``[E, T, ...]`` in ANY memory layout whose (env, step) rows are contiguous, in particular
    the ``[E, T]`` views of the engine's time-major rollout storage - in one kernel launch per 24 keys (hs_gather_rows);
    the flattened copy the reference makes first never exists.  ``indices``: int64 flat sample ids n = env * T + step.
    Returns {key tuple: [len(indices), ...] contiguous tensor}; bit-identical to the reference's indexing."""
    items = list(_flat_items(batch))
    if not items:
        return {}
    dev = items[0][1].device
    if dev.type != "cuda":
        raise _lib.HsError("gather_minibatch runs on the GPU only (hs_gather_rows); got tensors on " + str(dev))
    idx = indices.to(device=dev, dtype=torch.int64).contiguous()
    n = idx.numel()
    out, descs, keep = {}, [], []
    for key, v in items:
        if v.dim() < 2:
            raise _lib.HsError(f"gather_minibatch: {key} must be shaped [E, T, ...]")
        T = v.shape[1]
        if num_steps is None:
            num_steps = T
        if T != num_steps:
            raise _lib.HsError(f"gather_minibatch: {key} has {T} steps, expected {num_steps}")
        row = v[0, 0]
        if not row.is_contiguous():
            v = v.contiguous()                       # rows themselves strided: rare (none of the engine's tensors)
            row = v[0, 0]
        keep.append(v)
        src = v.view(torch.uint8) if v.dtype == torch.bool else v
        dst = torch.empty((n,) + tuple(v.shape[2:]), dtype=v.dtype, device=dev)
        out[key] = dst
        d = _lib.hs_gather_tensor()
        d.src, d.dst = src.data_ptr(), dst.data_ptr()
        d.stride_env, d.stride_step = v.stride(0) * v.element_size(), v.stride(1) * v.element_size()
        d.row_bytes = max(1, row.numel()) * v.element_size()
        descs.append(d)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        for i in range(0, len(descs), _lib.HS_GATHER_MAX_TENSORS):
            chunk = descs[i:i + _lib.HS_GATHER_MAX_TENSORS]
            arr = (_lib.hs_gather_tensor * len(chunk))(*chunk)
            check(lib.hs_gather_rows(arr, len(chunk), idx.data_ptr(), n, int(num_steps), stream), "hs_gather_rows")
    return out


def make_dataset_naive(batch, num_minibatches: int = 4, seq_len: int = 1, perm: Optional[torch.Tensor] = None):
    """omni_drones/learning/mappo.py:493-513: a random permutation of the first (N // M) * M samples, split into M
    minibatches, each gathered with :func:`gather_minibatch`.  seq_len == 1 (the reference's default: MAPPOPolicy has no
    ``minibatch_seq_len``): N = E * T flat samples, minibatch tensors ``[N / M, ...]``.  seq_len = L > 1: the steps are cut
    into T // L chunks, N = E * (T // L) samples of L consecutive steps, minibatch tensors ``[N / M, L, ...]``.  ``perm``
    (optional) supplies the permutation, e.g. the reference's own ``torch.randperm`` draw; yields {key tuple: tensor}."""
    first = next(_flat_items(batch))[1]
    E, T = first.shape[:2]
    L = int(seq_len)
    chunks = T // L if L > 1 else T
    total = (E * chunks // num_minibatches) * num_minibatches
    if perm is None:
        perm = torch.randperm(total, device=first.device)
    perm = perm.to(first.device).reshape(num_minibatches, -1)
    steps = torch.arange(L, device=first.device) if L > 1 else None
    for indices in perm:
        if L == 1:
            yield gather_minibatch(batch, indices, T)
        else:
            # sample s = env * chunks + chunk  ->  flat rows env * T + chunk * L + (0 .. L-1)
            env, chunk = indices // chunks, indices % chunks
            rows = ((env * T + chunk * L).unsqueeze(1) + steps).reshape(-1)
            mb = gather_minibatch(batch, rows, T)
            yield {k: v.view(indices.numel(), L, *v.shape[1:]) for k, v in mb.items()}
