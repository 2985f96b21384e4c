"""Host-side mirror of the reference's flow model API.

`RQSFlow(event_shape, num_layers, hidden_sizes, num_bins, ...)` keeps the
signature of /root/reference/cnf_ot/models/flows.py:178-228 and returns an
object with the interface the reference gets from
`hk.without_apply_rng(hk.multi_transform(RQSFlow(...)))`
(/root/reference/cnf_ot/mfc/solvers.py:41-54):

  params = model.init(rng, x, cond)
  model.apply.log_prob(params, value, cond=...)
  model.apply.sample(params, *, cond, seed, sample_shape)
  model.apply.sample_and_log_prob(params, *, cond, seed, sample_shape)
  model.apply.forward(params, x, c) / model.apply.inverse(params, y, c)

`params` is the haiku-shaped two-level dict of SURVEY.md A.3; its leaves are
views into ONE device buffer (the blob the kernels read), so packing is free.
All arithmetic runs in libcnfot.so; torch only owns the memory.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import Dict, Optional, Sequence

import torch

from . import ops, random
from .layout import FlowShape, pack

Flow = namedtuple("Flow", [
  "log_prob", "sample", "sample_and_log_prob", "forward", "inverse", "forward_jac",
  "inverse_jac", "gauge_potential"
])


class ParamTree(dict):
  """haiku-style params whose leaves alias one contiguous fp32 device blob."""

  def __init__(self, shape: FlowShape, blob: torch.Tensor):
    super().__init__()
    if blob.numel() != shape.blob_size or blob.dtype != torch.float32:
      raise ValueError("blob does not match the flow shape")
    self.shape = shape
    self.blob = blob
    for mod, leaf, shp, off, stride in shape.leaves():
      rows = 1
      for s in shp[:-1]:
        rows *= s
      view = torch.as_strided(blob, (rows, shp[-1]), (stride, 1), off).view(*shp) if len(shp) == 2 \
        else torch.as_strided(blob, (shp[-1], ), (1, ), off)
      self.setdefault(mod, {})[leaf] = view

  def like(self, blob: torch.Tensor) -> "ParamTree":
    return ParamTree(self.shape, blob)

  def clone(self) -> "ParamTree":
    return ParamTree(self.shape, self.blob.clone())

  # -- parameter I/O (the reference itself never saves its parameters): one array per
  # haiku leaf under the key "<module>/<leaf>" (e.g. "mlp_layer0_d1/~/linear_0/w", "~/first"), float32 -- the file a
  # `np.savez(path, **flatten(params))` of the reference's pytree would give -- plus optional extras: the flow shape
  # ("__flow_shape__", inferred from the leaves when absent) and whatever the caller adds (solvers.main: "opt/mu",
  # "opt/nu" in blob layout, "opt/count", "rng/train_key").
  @staticmethod
  def _npz_path(path: str) -> str:
    return path if str(path).endswith(".npz") else str(path) + ".npz"   # np.savez appends the suffix itself

  def save(self, path: str, extras: Optional[Dict] = None) -> None:
    import numpy as np
    out = {f"{mod}/{leaf}": view.detach().cpu().numpy() for mod, leaves in self.items() for leaf, view in leaves.items()}
    sh = self.shape
    out["__flow_shape__"] = np.array([sh.dim, sh.num_layers, sh.mlp_layers, sh.hidden, sh.num_bins, int(sh.conditional)])
    for k, v in (extras or {}).items():
      out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.uint64 if k.startswith("rng/") else None)
    np.savez(ParamTree._npz_path(path), **out)

  @staticmethod
  def _infer_shape(files, z) -> FlowShape:
    """FlowShape from the leaf names and shapes of a plain `np.savez(path, **flatten(params))` file."""
    import re
    layers, dims, mlps = set(), set(), set()
    for f in files:
      m = re.match(r"mlp_layer(\d+)_d(\d+)/~/linear_(\d+)/w$", f)
      if m:
        layers.add(int(m.group(1))); dims.add(int(m.group(2))); mlps.add(int(m.group(3)))
    if not layers:
      raise ValueError("cannot infer the flow shape: no mlp_layer*_d*/~/linear_*/w leaves")
    dim, L, M = max(dims) + 1, max(layers) + 1, max(mlps) + 1
    w0 = z["mlp_layer0_d1/~/linear_0/w"]
    hidden = int(w0.shape[1])
    P = int(np_shape(z["~/first"])[-1])
    if (P - 1) % 3:
      raise ValueError("~/first has an unexpected size")
    return FlowShape(dim, L, M, hidden, (P - 1) // 3, conditional=int(w0.shape[0]) == 2)

  @staticmethod
  def load(path: str, device=None, with_extras: bool = False):
    import numpy as np
    z = np.load(ParamTree._npz_path(path) if not __import__("os").path.exists(path) else path)
    extra_keys = {f for f in z.files if f == "__flow_shape__" or f.startswith(("opt/", "rng/"))}
    if "__flow_shape__" in z.files:
      d, n_layers, m, h, k, cond = (int(v) for v in z["__flow_shape__"])
      shape = FlowShape(d, n_layers, m, h, k, conditional=bool(cond))
    else:
      shape = ParamTree._infer_shape(set(z.files) - extra_keys, z)
    tree = ParamTree(shape, torch.zeros(shape.blob_size, dtype=torch.float32))
    want = {f"{mod}/{leaf}" for mod, leaves in tree.items() for leaf in leaves}
    have = set(z.files) - extra_keys
    if want != have:
      raise ValueError(f"parameter file does not match the flow shape: {sorted(want ^ have)[:4]} ...")
    for mod, leaves in tree.items():
      for leaf, view in leaves.items():
        arr = torch.from_numpy(np.asarray(z[f"{mod}/{leaf}"], dtype=np.float32))
        if tuple(arr.shape) != tuple(view.shape):
          raise ValueError(f"{mod}/{leaf}: shape {tuple(arr.shape)} != {tuple(view.shape)}")
        view.copy_(arr)
    tree = tree if device is None else ParamTree(shape, tree.blob.to(device))
    if not with_extras:
      return tree
    extras = {}
    for k in extra_keys - {"__flow_shape__"}:
      a = np.asarray(z[k])
      extras[k] = int(a) if a.ndim == 0 else torch.from_numpy(a.astype(np.float32))
    return tree, extras


def np_shape(a):
  return tuple(a.shape)


def _blob_of(shape: FlowShape, params, device) -> torch.Tensor:
  if isinstance(params, ParamTree):
    return params.blob
  return pack(shape, params).to(device)  # plain dict: pack on the fly (slow path)


def _trunc_normal(key, shape, std, device):
  # hk.Linear default w_init: TruncatedNormal(stddev = 1/sqrt(fan_in)), cut at +-2 sigma
  out = torch.empty(shape, dtype=torch.float32, device=device)
  torch.nn.init.trunc_normal_(out, mean=0.0, std=1.0, a=-2.0, b=2.0,
                              generator=random._generator(random.as_key(key), shape, device, 7))
  return out * std


class _Apply:
  def __init__(self, model: "FlowModel"):
    self._m = model
    self.log_prob = model._log_prob
    self.sample = model._sample
    self.sample_and_log_prob = model._sample_and_log_prob
    self.forward = model._forward
    self.inverse = model._inverse
    self.forward_jac = model._unsupported("forward_jac")
    self.inverse_jac = model._unsupported("inverse_jac")
    self.gauge_potential = model._unsupported("gauge_potential")


class FlowModel:
  """What `hk.without_apply_rng(hk.multi_transform(RQSFlow(...)))` is to the solvers."""

  def __init__(self, shape: FlowShape, device=None):
    self.shape = shape
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    self.apply = _Apply(self)

  # -- model.init(rng, x, cond): reference initialisation = identity flow (flows.py:48-55,65-81)
  def init(self, rng, x=None, cond=None) -> ParamTree:
    s = self.shape
    if x is not None and tuple(x.shape[-1:]) != (s.dim, ):
      raise ValueError(f"init: expected x of shape (..., {s.dim})")
    params = ParamTree(s, torch.zeros(s.blob_size, dtype=torch.float32, device=self.device))
    keys = random.split(random.as_key(rng), s.num_layers * max(s.dim - 1, 1) * s.mlp_layers + 1)
    k = 0
    for l in range(s.num_layers):
      for d in range(1, s.dim):
        fan_in = d + 1 if s.conditional else d
        for m in range(s.mlp_layers):
          w = params[f"mlp_layer{l}_d{d}/~/linear_{m}"]["w"]
          w.copy_(_trunc_normal(keys[k], tuple(w.shape), 1.0 / math.sqrt(fan_in), self.device))
          k += 1
          fan_in = s.hidden
    return params

  def _unsupported(self, name):
    def fn(*a, **k):
      raise NotImplementedError(
        f"{name} is never called by the MFC solvers (SURVEY.md §8b) and is not part of the B200 path")
    return fn

  def _latent(self, seed, sample_shape, latent):
    if latent is not None:
      return latent
    n = 1
    for v in (sample_shape if isinstance(sample_shape, (tuple, list)) else (sample_shape, )):
      n *= int(v)
    return random.normal(seed, (n, self.shape.dim), device=self.device)

  def _log_prob(self, params, value, cond=None):
    """ConditionalTransformed.log_prob (conditional.py:316-321); cond: (1,) broadcast."""
    W = _blob_of(self.shape, params, self.device)
    _, lp = ops.flow_eval(self.shape, W, value, cond, inverse=True, add_base=True)
    return lp.reshape(value.shape[:-1])

  def _sample(self, params, *, cond, seed=None, sample_shape=(), latent=None):
    """ConditionalTransformed.sample (conditional.py:323-351); cond: (N, 1).
    `latent=` (an explicit N(0,I) draw) may replace `seed`."""
    W = _blob_of(self.shape, params, self.device)
    x = self._latent(seed, sample_shape, latent)
    y, _ = ops.flow_eval(self.shape, W, x, cond, inverse=False, want_logdet=False)
    return y

  def _sample_and_log_prob(self, params, *, cond, seed=None, sample_shape=(), latent=None):
    """ConditionalTransformed.sample_and_log_prob (conditional.py:353-402)."""
    W = _blob_of(self.shape, params, self.device)
    x = self._latent(seed, sample_shape, latent)
    return ops.flow_eval(self.shape, W, x, cond, inverse=False, add_base=True)

  def _cond(self, c):
    # unconditional flows (cond_shape=(0,), dr/trainers.py:41-68) run the same kernels with t = 0 and zero t-weights
    if not self.shape.conditional:
      if c is not None:
        raise TypeError("this flow is unconditional (cond_shape=(0,)): no condition argument")
      return torch.zeros(1)
    return c

  def _forward(self, params, x, c=None):
    """flow.bijector.forward: latent -> physical."""
    W = _blob_of(self.shape, params, self.device)
    return ops.flow_eval(self.shape, W, x, self._cond(c), inverse=False, want_logdet=False)[0]

  def _inverse(self, params, y, c=None):
    """flow.bijector.inverse: physical -> latent."""
    W = _blob_of(self.shape, params, self.device)
    return ops.flow_eval(self.shape, W, y, self._cond(c), inverse=True, want_logdet=False)[0]


def RQSFlow(
  event_shape: Sequence[int],
  num_layers: int,
  hidden_sizes: Sequence[int],
  num_bins: int,
  periodized: bool = False,
  cond_shape=(1, ),
  base_range=(0, 2 * math.pi),
  device=None,
) -> FlowModel:
  """Same arguments as cnf_ot.models.flows.RQSFlow (flows.py:178-186)."""
  if periodized:
    raise NotImplementedError("periodized flows are outside the MFC hot path (solvers.py:46 passes False)")
  if tuple(cond_shape) not in ((1, ), (0, )):
    raise NotImplementedError("cond_shape must be (1,) (time-conditioned, mfc/) or (0,) (unconditional, dr/)")
  if len(event_shape) != 1:
    raise ValueError("event_shape must be (dim,)")
  hs = [int(h) for h in hidden_sizes]
  if not hs or any(h != hs[0] for h in hs):
    raise ValueError("hidden_sizes must be [hidden_size] * mlp_num_layers (solvers.py:44)")
  return FlowModel(FlowShape(int(event_shape[0]), int(num_layers), len(hs), hs[0], int(num_bins),
                             conditional=tuple(cond_shape) == (1, )), device)


# haiku spellings used at solvers.py:48 -- the model above is already "transformed"
def multi_transform(model):
  return model


def without_apply_rng(model):
  return model
