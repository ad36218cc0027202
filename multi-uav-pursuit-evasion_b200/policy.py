"""Policy inference next to the tick: the reference's MAPPO actor /
critic for this task - ``PartialAttentionEncoder`` + ``DiagGaussian`` / ``v_out``
(omni_drones/learning/modules/networks.py:249-314, modules/distributions.py:66-82,
omni_drones/learning/mappo.py:575-668) - evaluated by ONE kernel per network
(``hs_policy_forward``) from the module's live parameters.  The learner itself (optimisers,
losses) stays where it is; after an optimiser step call :meth:`FusedPolicy.refresh`.

Parameter names are the reference's ``state_dict`` names relative to the encoder
(``split_embed.embed.state_self.weight`` ... ``norm2.bias``); longer prefixed names
(``module.encoder.split_embed...`` as in ``make_functional(actor)``) are matched by suffix.  The
head is ``head.weight`` / ``head.bias`` (aliases: ``act_dist.fc_mean.*``, ``v_out.*``) and, for an
actor, ``log_std`` (alias ``act_dist.log_std``).  No CPU path.
"""
import ctypes as C
from typing import Dict, Mapping, Optional

import torch

from . import _lib
from ._lib import check, lib

_ENC = {
    "embed_self_w": "split_embed.embed.state_self.weight", "embed_self_b": "split_embed.embed.state_self.bias",
    "embed_others_w": "split_embed.embed.state_others.weight", "embed_others_b": "split_embed.embed.state_others.bias",
    "embed_cyl_w": "split_embed.embed.cylinders.weight", "embed_cyl_b": "split_embed.embed.cylinders.bias",
    "embed_ln_w": "split_embed.layer_norm.weight", "embed_ln_b": "split_embed.layer_norm.bias",
    "attn_in_w": "attn.in_proj_weight", "attn_in_b": "attn.in_proj_bias",
    "attn_out_w": "attn.out_proj.weight", "attn_out_b": "attn.out_proj.bias",
    "lin1_w": "linear1.weight", "lin1_b": "linear1.bias", "lin2_w": "linear2.weight", "lin2_b": "linear2.bias",
    "norm1_w": "norm1.weight", "norm1_b": "norm1.bias", "norm2_w": "norm2.weight", "norm2_b": "norm2.bias",
}
_HEAD = {"head_w": ("head.weight", "fc_mean.weight", "v_out.weight"), "head_b": ("head.bias", "fc_mean.bias", "v_out.bias"),
         "log_std": ("log_std",)}


def _find(params: Mapping[str, torch.Tensor], names, required=True):
    names = (names,) if isinstance(names, str) else names
    for n in names:
        for k, v in params.items():
            ks = ".".join(k) if isinstance(k, tuple) else k
            if ks == n or ks.endswith("." + n):
                return v
    if required:
        raise _lib.HsError(f"FusedPolicy: parameter {names[0]} not found")
    return None


def init_params(self_dim: int, n_others: int, n_cyl: int, head_dim: int, actor: bool, device="cuda:0",
                gain: float = 0.01) -> Dict[str, torch.Tensor]:
    """Fresh parameters with the reference's shapes and initialisers (torch defaults for the encoder,
    networks.py:126-151, 249-279; orthogonal(gain) head with zero bias and log_std = 0, distributions.py:66-76,
    mappo.py:596-603) under the reference's state_dict names - for benchmarks and tests, and as the layout a
    learner's own modules must have."""
    import torch.nn as nn
    mods = {"split_embed.embed.state_self": nn.Linear(self_dim, 128), "split_embed.layer_norm": nn.LayerNorm(128),
            "attn": nn.MultiheadAttention(128, 1, batch_first=True), "linear1": nn.Linear(128, 128),
            "linear2": nn.Linear(128, 128), "norm1": nn.LayerNorm(128), "norm2": nn.LayerNorm(128)}
    if n_others:
        mods["split_embed.embed.state_others"] = nn.Linear(3, 128)
    if n_cyl:
        mods["split_embed.embed.cylinders"] = nn.Linear(5, 128)
    head = nn.Linear(128, head_dim)
    nn.init.orthogonal_(head.weight, gain)
    nn.init.zeros_(head.bias)
    out = {}
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            out[f"{name}.{k}"] = v.detach().to(device).contiguous()
    out["head.weight"], out["head.bias"] = head.weight.detach().to(device).contiguous(), head.bias.detach().to(device)
    if actor:
        out["log_std"] = torch.zeros(head_dim, device=device)
    return out


class FusedPolicy:
    """One network (actor or critic) bound to its live parameters."""

    def __init__(self, params: Mapping[str, torch.Tensor], n_others: int, n_cyl: int, device="cuda:0"):
        device = torch.device(device)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.HsError("FusedPolicy needs a CUDA device: policy inference has no CPU path")
        self.device, self.n_others, self.n_cyl = device, int(n_others), int(n_cyl)
        self.impl = 0                                           # kernel choice of forward(): 0 auto, 1 FFMA, 2 tcgen05
        self._p: Dict[str, Optional[torch.Tensor]] = {}
        for f, n in _ENC.items():
            optional = (f.startswith("embed_others") and n_others == 0) or (f.startswith("embed_cyl") and n_cyl == 0)
            self._p[f] = _find(params, n, required=not optional)
        for f, n in _HEAD.items():
            self._p[f] = _find(params, n, required=f != "log_std")
        for f, t in self._p.items():
            if t is not None and (t.dtype != torch.float32 or t.device != device or not t.is_contiguous()):
                raise _lib.HsError(f"FusedPolicy: {f} must be a contiguous float32 tensor on {device}")
        self.self_dim = int(self._p["embed_self_w"].shape[1])
        self.head_dim = int(self._p["head_w"].shape[0])
        self.is_actor = self._p["log_std"] is not None
        n = int(lib.hs_policy_blob_floats(self.self_dim))
        if n <= 0:
            raise _lib.HsError(f"FusedPolicy: unsupported state_self width {self.self_dim}")
        self.blob = torch.zeros(n, dtype=torch.float32, device=device)
        self.refresh()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def refresh(self):
        """Re-reads the live parameters (hs_policy_prepare); call after an optimiser step."""
        w = _lib.hs_policy_weights()
        for f, t in self._p.items():
            setattr(w, f, None if t is None else t.data_ptr())
        w.self_dim, w.head_dim = self.self_dim, self.head_dim
        with torch.cuda.device(self.device):
            check(lib.hs_policy_prepare(C.byref(w), self.blob.data_ptr(), self._stream()), "hs_policy_prepare")
        return self

    def seed(self, seed: int, step: int = 0):
        """(Re)creates the device RNG state {seed, step, 0, 0} used when ``forward(sample=True)`` draws the noise
        in the kernel (counter-based Philox: a draw depends only on (seed, step, row))."""
        self.rng_state = torch.tensor([seed & (2 ** 63 - 1), step, 0, 0], dtype=torch.int64, device=self.device)
        return self

    def forward(self, state_self: torch.Tensor, state_others: Optional[torch.Tensor], cylinders: Optional[torch.Tensor],
                eps: Optional[torch.Tensor] = None, sample: bool = False, out: Optional[Dict[str, torch.Tensor]] = None,
                want_features: bool = False, want_eps: bool = False, impl: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """state_self [..., 1, D], state_others [..., n_others, 3], cylinders [..., n_cyl, 5] (the
        ``("agents", "observation")`` entries, any leading batch dims).  Returns ``head`` [..., head_dim]
        (action mean or state value) and, for an actor, ``action`` [..., head_dim] and ``logp`` [..., 1].
        Noise: ``eps`` (caller-supplied standard normal), else ``sample=True`` (drawn in the kernel from the
        state set by :meth:`seed`), else none - the mode, as ``deterministic=True``.  ``out`` may hold
        preallocated result tensors (static buffers for CUDA-graph capture)."""
        D = self.self_dim
        if state_self.shape[-1] != D or state_self.dtype != torch.float32 or not state_self.is_contiguous():
            raise _lib.HsError(f"FusedPolicy: state_self must be contiguous float32 [..., 1, {D}]")
        lead = tuple(state_self.shape[:-2])
        R = state_self.numel() // D
        chk = lambda t, n, d: t is not None and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == R * n * d
        if self.n_others and not chk(state_others, self.n_others, 3):
            raise _lib.HsError("FusedPolicy: state_others must be contiguous float32 [..., n_others, 3]")
        if self.n_cyl and not chk(cylinders, self.n_cyl, 5):
            raise _lib.HsError("FusedPolicy: cylinders must be contiguous float32 [..., n_cyl, 5]")
        out = {} if out is None else out
        mk = lambda k, w: out.setdefault(k, torch.empty(lead + (w,), dtype=torch.float32, device=self.device))
        io = _lib.hs_policy_io()
        io.num_rows, io.n_others, io.n_cyl = R, self.n_others, self.n_cyl
        io.impl = int(self.impl if impl is None else impl)      # 0 auto, 1 fp32 FFMA kernel, 2 tcgen05 (3xTF32) kernel
        io.state_self = state_self.data_ptr()
        io.state_others = state_others.data_ptr() if self.n_others else None
        io.cylinders = cylinders.data_ptr() if self.n_cyl else None
        io.head_out = mk("head", self.head_dim).data_ptr()
        if self.is_actor:
            io.action, io.logp = mk("action", self.head_dim).data_ptr(), mk("logp", 1).data_ptr()
            if eps is not None:
                if eps.numel() != R * self.head_dim or not eps.is_contiguous() or eps.dtype != torch.float32:
                    raise _lib.HsError("FusedPolicy: eps must be contiguous float32 [..., head_dim]")
                io.eps = eps.data_ptr()
            elif sample:
                if getattr(self, "rng_state", None) is None:
                    self.seed(0)
                io.rng_state = self.rng_state.data_ptr()
            if want_eps:
                io.eps_out = mk("eps", self.head_dim).data_ptr()
        if want_features:
            io.feat_out = mk("features", 128).data_ptr()
        with torch.cuda.device(self.device):
            check(lib.hs_policy_forward(self.blob.data_ptr(), D, self.head_dim, C.byref(io), self._stream()),
                  "hs_policy_forward")
        return out

    __call__ = forward


class MAPPOActorCritic:
    """What ``MAPPOPolicy.__call__`` does during a rollout (omni_drones/learning/mappo.py:235-251: shared actor over
    ``("agents", "observation")`` -> ``("agents", "action")`` + ``"drone.action_logp"``, then ``value_op`` with
    ``critic_input: obs`` -> ``"state_value"``), as two kernel launches.  A drop-in for the ``policy`` argument of the
    collector (``SyncDataCollector(env, policy=MAPPOActorCritic(actor, critic))``); the learner keeps its own modules and
    calls :meth:`refresh` after its optimiser steps."""

    def __init__(self, actor: FusedPolicy, critic: Optional[FusedPolicy] = None, agent_name: str = "drone",
                 obs_key=("agents", "observation"), action_key=("agents", "action"), keep_noise: bool = False):
        if not actor.is_actor:
            raise _lib.HsError("MAPPOActorCritic: `actor` needs a log_std parameter (DiagGaussian head)")
        self.actor, self.critic = actor, critic
        self.obs_key, self.action_key = tuple(obs_key), tuple(action_key)
        self.logp_key, self.keep_noise = f"{agent_name}.action_logp", keep_noise
        if getattr(actor, "rng_state", None) is None:
            actor.seed(0)

    def refresh(self):
        self.actor.refresh()
        if self.critic is not None:
            self.critic.refresh()
        return self

    def __call__(self, tensordict, deterministic: bool = False):
        obs = tensordict.get(self.obs_key)
        s = obs.get("state_self")
        o = obs.get("state_others") if self.actor.n_others else None
        c = obs.get("cylinders") if self.actor.n_cyl else None
        # fresh result tensors every call: the collector keeps references to the previous step's
        a = self.actor.forward(s, o, c, sample=not deterministic, want_eps=self.keep_noise)
        tensordict.set(self.action_key, a["action"])
        tensordict.set(self.logp_key, a["logp"])
        if self.keep_noise and "eps" in a:
            tensordict.set("action_noise", a["eps"])
        if self.critic is not None:
            tensordict.set("state_value", self.critic.forward(s, o, c)["head"])
        return tensordict
