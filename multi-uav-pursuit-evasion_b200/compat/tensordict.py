"""Minimal stand-in for the slice of ``tensordict`` (pinned fork btx0424/tensordict@6d8119c)
that the environment, the collector and scripts/train.py exercise.

Used only when the real package is not importable (it is absent from this image).  It is a
container: nested dict of tensors sharing leading batch dimensions.  No arithmetic of the
hot path lives here.
"""
from typing import Any, Dict, Iterable, List, Sequence, Tuple, Union

import torch

NestedKey = Union[str, Tuple[str, ...]]


def _norm_key(key) -> Tuple[str, ...]:
    if isinstance(key, str):
        return (key,)
    if isinstance(key, tuple) and all(isinstance(k, str) for k in key):
        return key
    raise KeyError(f"not a tensordict key: {key!r}")


def _is_key(key) -> bool:
    return isinstance(key, str) or (isinstance(key, tuple) and len(key) > 0 and all(isinstance(k, str) for k in key))


class TensorDict:
    def __init__(self, source: Dict[str, Any] = None, batch_size: Sequence[int] = (), device=None):
        self._d: Dict[str, Any] = {}
        self._batch_size = torch.Size(batch_size)
        self._device = torch.device(device) if device is not None else None
        for k, v in (source or {}).items():
            self.set(k, v)

    # ---- meta ------------------------------------------------------------------------
    @property
    def batch_size(self) -> torch.Size:
        return self._batch_size

    @batch_size.setter
    def batch_size(self, value):
        self._batch_size = torch.Size(value)
        for v in self._d.values():
            if isinstance(v, TensorDict):
                v.batch_size = torch.Size(value) + v.batch_size[len(value):] if len(v.batch_size) >= len(value) else torch.Size(value)

    @property
    def shape(self):
        return self._batch_size

    @property
    def device(self):
        if self._device is not None:
            return self._device
        for v in self._d.values():
            d = v.device
            if d is not None:
                return d
        return None

    def numel(self) -> int:
        n = 1
        for s in self._batch_size:
            n *= s
        return n

    def batch_dims(self) -> int:
        return len(self._batch_size)

    # ---- element access ----------------------------------------------------------------
    def _wrap(self, v):
        if isinstance(v, dict):
            return TensorDict(v, self._batch_size, self._device)
        if isinstance(v, (int, float, bool)):
            return torch.full(tuple(self._batch_size), v, device=self.device)
        return v

    def set(self, key: NestedKey, value, inplace: bool = False):
        key = _norm_key(key)
        if len(key) > 1:
            sub = self._d.get(key[0])
            if not isinstance(sub, TensorDict):
                sub = TensorDict({}, self._batch_size, self._device)
                self._d[key[0]] = sub
            sub.set(key[1:], value, inplace)
            return self
        value = self._wrap(value)
        if inplace and key[0] in self._d and isinstance(self._d[key[0]], torch.Tensor):
            self._d[key[0]].copy_(value)
        else:
            self._d[key[0]] = value
        return self

    def set_(self, key, value):
        return self.set(key, value, inplace=True)

    def get(self, key: NestedKey, default=...):
        try:
            cur = self
            for k in _norm_key(key):
                cur = cur._d[k] if isinstance(cur, TensorDict) else cur[k]
            return cur
        except KeyError:
            if default is ...:
                raise
            return default

    def __contains__(self, key):
        try:
            self.get(key)
            return True
        except KeyError:
            return False

    def __getitem__(self, idx):
        if _is_key(idx):
            return self.get(idx)
        return self._index(idx)

    def _index(self, idx):
        out = {}
        new_bs = None
        for k, v in self._d.items():
            out[k] = v._index(idx) if isinstance(v, TensorDict) else v[idx]
        probe = torch.empty(tuple(self._batch_size), device="meta")[idx]
        new_bs = probe.shape
        td = TensorDict({}, new_bs, self._device)
        td._d = out
        return td

    def __setitem__(self, idx, value):
        if _is_key(idx):
            self.set(idx, value)
            return
        for k, v in self._d.items():
            src = value.get(k) if isinstance(value, TensorDict) else value
            if isinstance(v, TensorDict):
                v[idx] = src
            else:
                v[idx] = src

    def pop(self, key, default=...):
        key = _norm_key(key)
        parent = self if len(key) == 1 else self.get(key[:-1])
        if key[-1] in parent._d:
            return parent._d.pop(key[-1])
        if default is ...:
            raise KeyError(key)
        return default

    # ---- iteration ---------------------------------------------------------------------
    def keys(self, include_nested: bool = False, leaves_only: bool = False) -> List:
        out = []
        for k, v in self._d.items():
            if isinstance(v, TensorDict):
                if not leaves_only:
                    out.append(k)
                if include_nested:
                    out.extend((k,) + (s if isinstance(s, tuple) else (s,)) for s in v.keys(True, leaves_only))
            else:
                out.append(k)
        return out

    def items(self, include_nested: bool = False, leaves_only: bool = False):
        return [(k, self.get(k)) for k in self.keys(include_nested, leaves_only)]

    def values(self, include_nested: bool = False, leaves_only: bool = False):
        return [self.get(k) for k in self.keys(include_nested, leaves_only)]

    # ---- whole-dict ops -------------------------------------------------------------------
    def apply(self, fn, batch_size=None) -> "TensorDict":
        td = TensorDict({}, self._batch_size if batch_size is None else batch_size, self._device)
        for k, v in self._d.items():
            td._d[k] = v.apply(fn, batch_size) if isinstance(v, TensorDict) else fn(v)
        return td

    def _shallow(self) -> "TensorDict":
        td = TensorDict.__new__(TensorDict)
        td._batch_size, td._device = self._batch_size, self._device
        td._d = {k: (v._shallow() if type(v) is TensorDict else v) for k, v in self._d.items()}
        return td

    def clone(self, recurse: bool = True) -> "TensorDict":
        return self.apply(lambda t: t.clone()) if recurse else self._shallow()

    def to_tensordict(self):
        return self.clone()

    def contiguous(self):
        return self.apply(lambda t: t.contiguous())

    def detach(self):
        return self.apply(lambda t: t.detach())

    def to(self, device):
        td = self.apply(lambda t: t.to(device))
        td._device = torch.device(device)
        return td

    def cpu(self):
        return self.to("cpu")

    def update(self, other, inplace: bool = False):
        items = other._d.items() if isinstance(other, TensorDict) else other.items()
        for k, v in items:
            cur = self._d.get(k) if isinstance(k, str) else None
            if isinstance(v, (TensorDict, dict)) and isinstance(cur, TensorDict):
                cur.update(v, inplace)
            elif isinstance(v, TensorDict) and isinstance(k, str):
                self._d[k] = v.clone(False)            # never alias the other tensordict's containers
            else:
                self.set(k, v, inplace)
        return self

    def update_(self, other):
        return self.update(other, inplace=True)

    def select(self, *keys, strict: bool = True, inplace: bool = False) -> "TensorDict":
        td = TensorDict({}, self._batch_size, self._device)
        for k in keys:
            try:
                td.set(k, self.get(k))
            except KeyError:
                if strict:
                    raise
        if inplace:
            self._d = td._d
            return self
        return td

    def exclude(self, *keys, inplace: bool = False) -> "TensorDict":
        td = self if inplace else self.clone(False)
        for k in keys:
            td.pop(k, None)
        return td

    def empty(self):
        return TensorDict({}, self._batch_size, self._device)

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        nb = len(self._batch_size)
        probe = torch.empty(tuple(self._batch_size), device="meta").reshape(*shape)
        return self.apply(lambda t: t.reshape(*probe.shape, *t.shape[nb:]), batch_size=probe.shape)

    def view(self, *shape):
        return self.reshape(*shape)

    def unsqueeze(self, dim):
        nb = len(self._batch_size)
        dim = dim if dim >= 0 else nb + 1 + dim
        probe = torch.empty(tuple(self._batch_size), device="meta").unsqueeze(dim)
        return self.apply(lambda t: t.unsqueeze(dim), batch_size=probe.shape)

    def squeeze(self, dim):
        probe = torch.empty(tuple(self._batch_size), device="meta").squeeze(dim)
        return self.apply(lambda t: t.squeeze(dim), batch_size=probe.shape)

    def expand(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        nb = len(self._batch_size)
        return self.apply(lambda t: t.expand(*shape, *t.shape[nb:]), batch_size=torch.Size(shape))

    def unbind(self, dim: int):
        n = self._batch_size[dim]
        idx = lambda i: tuple([slice(None)] * dim + [i])
        return tuple(self._index(idx(i)) for i in range(n))

    def mean(self):  # convenience used by logging code
        return self.apply(lambda t: t.float().mean(), batch_size=())

    @staticmethod
    def stack(tds: Iterable["TensorDict"], dim: int = 0) -> "TensorDict":
        tds = list(tds)
        first = tds[0]
        bs = list(first.batch_size)
        d = dim if dim >= 0 else len(bs) + 1 + dim
        bs.insert(d, len(tds))
        out = TensorDict({}, bs, first._device)
        for k, v in first._d.items():
            if isinstance(v, TensorDict):
                out._d[k] = TensorDict.stack([t._d[k] for t in tds], d)
            else:
                out._d[k] = torch.stack([t._d[k] for t in tds], d)
        return out

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.stack:
            return cls.stack(*args, **kwargs)
        raise NotImplementedError(f"{func} is not supported by the TensorDict stand-in")

    def __len__(self):
        return self._batch_size[0] if len(self._batch_size) else 0

    def __repr__(self):
        def fmt(v):
            return repr(v) if isinstance(v, TensorDict) else f"Tensor{tuple(v.shape)} {str(v.dtype).replace('torch.', '')}"
        body = ", ".join(f"{k}: {fmt(v)}" for k, v in self._d.items())
        return f"TensorDict({{{body}}}, batch_size={list(self._batch_size)})"


TensorDictBase = TensorDict
