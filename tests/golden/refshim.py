"""Stand-ins for the `jax` / `haiku` / `distrax` / `jaxtyping` APIs, backed by torch float64 + autograd, so that the
reference's OWN source files run unmodified in this container (which has no JAX wheels and no network).

TEST INFRASTRUCTURE ONLY: used by `tests/golden/make_reference_golden.py`, which imports
`/root/reference/cnf_ot/models/{flows,autoregressive,conditional}.py` and `/root/reference/cnf_ot/mfc/applications.py`
as they lie and runs them to generate `tests/golden/ref_*.npz`.  Nothing else imports this module.

What runs from the reference: the flow construction (`RQSFlow`, `make_flow_model`, `make_conditioner`), the
autoregressive layers and their permutations, `ConditionalChain / ConditionalInverse / ConditionalTransformed`
(log_prob, sample, sample_and_log_prob), and every loss function of `applications.py`, including the finite
differences, the shared PRNG key, and the scalings.  What is stood in:
  * array API (`jax.numpy`, `.at[].set`, `jax.vmap` as a loop over the leading axis, `jax.random` as a programmable
    source: the draws are set by the caller and handed to the oracle / the kernels as explicit inputs; same key and
    same shape -> same values, a (b, D) draw is the first b rows of the (B, D) draw);
  * haiku's parameter plumbing (`multi_transform`, `get_parameter`, `Linear`, `nets.MLP`, `Flatten`, `Reshape`) with
    haiku's module naming rules (children made in `__init__` are `parent/~/child`; top-level parameters live in `~`);
  * distrax's `Bijector` / `Transformed` / `Normal` / `Independent` base classes, and
    `distrax.RationalQuadraticSpline`, which delegates to `oracle/rqs.py`: distrax is a third-party dependency that is
    not vendored in the reference, so the spline stays a restatement of the published algorithm (pinned by the
    reference's own invariants, `tests/test_oracle_rqs.py`).
"""
import math
import re
import sys
import types
from collections import namedtuple

import numpy as np
import torch

F64 = torch.float64


# ---------------------------------------------------------------------------------------------- array layer
def _as(v, dtype=None):
  if isinstance(v, torch.Tensor):
    return v if dtype is None else v.to(dtype)
  if isinstance(v, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in v):
    t = torch.stack([_as(e) for e in v])
  else:   # through numpy: python floats become float64 (torch.as_tensor(0.005) would round to float32 first)
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(v)))
  if dtype is not None:
    return t.to(dtype)
  return t.to(F64) if t.is_floating_point() else t


def _dt(dtype):
  return F64 if dtype is None else dtype


def _shape(s):
  return (int(s), ) if isinstance(s, (int, np.integer)) else tuple(int(v) for v in s)


class _At:
  """`x.at[idx].set(v)`: functional update (clone + assignment; autograd flows through both)."""

  def __init__(self, t):
    self.t = t
    self.idx = None

  def __getitem__(self, idx):
    self.idx = idx
    return self

  def set(self, v):
    out = self.t.clone()
    out[_norm_index(self.idx)] = _as(v).to(out.dtype)
    return out


def _norm_index(idx):
  """numpy index arrays (possibly negative-stride views such as p[::-1][:d]) -> lists."""
  if isinstance(idx, tuple):
    return tuple(_norm_index(i) for i in idx)
  if isinstance(idx, np.ndarray):
    return idx.tolist()
  return idx


def patch_tensor():
  if getattr(torch.Tensor, "_refshim", False):
    return
  orig_get = torch.Tensor.__getitem__
  torch.Tensor.__getitem__ = lambda self, idx: orig_get(self, _norm_index(idx))
  torch.Tensor.at = property(lambda self: _At(self))
  torch.Tensor._refshim = True


def _mod(name):
  m = types.ModuleType(name)
  sys.modules[name] = m
  if "." in name:
    parent, child = name.rsplit(".", 1)
    setattr(sys.modules[parent], child, m)
  return m


# ---------------------------------------------------------------------------------------------- programmable PRNG
class Draws:
  """The values every `jax.random.*` call returns (one key per loss call in the reference: solvers.py:94)."""
  normal = None    # (B, D)
  uniform = None   # (Tn,)
  choice = None    # (B,) int64
  log = []

  @classmethod
  def set(cls, normal=None, uniform=None, choice=None):
    cls.normal, cls.uniform, cls.choice, cls.log = normal, uniform, choice, []


def _build_jax():
  jax = _mod("jax")
  jnp = _mod("jax.numpy")
  jnp.ndarray = torch.Tensor
  jnp.float32, jnp.float64 = torch.float32, F64
  jnp.pi = math.pi
  jnp.ones = lambda shape, dtype=None: torch.ones(_shape(shape), dtype=_dt(dtype))
  jnp.zeros = lambda shape, dtype=None: torch.zeros(_shape(shape), dtype=_dt(dtype))
  jnp.zeros_like = lambda x: torch.zeros_like(x)
  jnp.array = lambda v, dtype=None: _as(v, dtype)
  jnp.asarray = jnp.array
  jnp.arange = lambda *a: torch.from_numpy(np.arange(*a))   # numpy semantics: an empty range is allowed
  jnp.eye = lambda n: torch.eye(int(n), dtype=F64)
  jnp.exp, jnp.log, jnp.sin, jnp.cos = torch.exp, torch.log, torch.sin, torch.cos
  jnp.subtract = lambda a, b: a - b
  jnp.abs = torch.abs
  jnp.max = lambda x, axis=None: torch.max(x) if axis is None else torch.max(x, dim=axis).values
  jnp.tile = lambda x, reps: x.repeat(*reps)
  jnp.dot = lambda a, b: _as(a).to(F64) @ _as(b).to(F64)
  # numpy, so that `np.ones(...) * (t - dt / 2)` (cnf_ot/utils.py:329) works as it does with a jax scalar
  jnp.linspace = lambda a, b, n: np.linspace(a, b, n)
  jnp.broadcast_to = lambda x, shape: torch.broadcast_to(_as(x), _shape(shape))
  jnp.concatenate = lambda xs, axis=0: torch.cat([_as(x).to(F64) for x in xs], dim=axis)
  jnp.concat = jnp.concatenate
  jnp.hstack = lambda xs: torch.cat([_as(x).to(F64) for x in xs], dim=1)   # cnf_ot/utils.py:588,618: (n, 1) columns

  def _reduce(fn):
    def f(x, axis=None, keepdims=False):
      return fn(x) if axis is None else fn(x, dim=axis, keepdim=keepdims)
    return f
  jnp.sum, jnp.mean = _reduce(torch.sum), _reduce(torch.mean)
  linalg = _mod("jax.numpy.linalg")
  linalg.norm = lambda x, axis=None: torch.linalg.norm(x) if axis is None else torch.linalg.norm(x, dim=axis)
  linalg.cholesky = torch.linalg.cholesky
  linalg.det = torch.linalg.det

  nn = _mod("jax.nn")
  nn.relu, nn.tanh = torch.relu, torch.tanh

  rnd = _mod("jax.random")

  class PRNGKey(int):
    pass
  rnd.PRNGKey = PRNGKey
  rnd.split = lambda key, num=2: tuple(PRNGKey(int(key) * 1000003 + i + 1) for i in range(num))

  def normal(key, shape=(), dtype=None):
    shape = _shape(shape)
    src = Draws.normal
    assert src is not None and shape[1:] == tuple(src.shape[1:]) and shape[0] <= src.shape[0], (shape, src.shape)
    Draws.log.append(("normal", shape))
    return src[:shape[0]].clone()

  def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    shape = _shape(shape)
    if Draws.uniform is None:   # not programmed (the reference's own unit tests): a seeded stream per key
      g = torch.Generator().manual_seed(int(key) % (2**63 - 1))
      return minval + (maxval - minval) * torch.rand(shape, generator=g, dtype=F64)
    assert len(shape) == 1 and shape[0] <= Draws.uniform.shape[0] and minval == 0.0 and maxval == 1.0
    Draws.log.append(("uniform", shape))
    return Draws.uniform[:shape[0]].clone()

  def choice(key, a, shape=(), p=None):
    shape = _shape(shape)
    assert int(a) == 8 and shape[0] <= Draws.choice.shape[0]
    Draws.log.append(("choice", shape))
    return Draws.choice[:shape[0]].clone()
  rnd.normal, rnd.uniform, rnd.choice = normal, uniform, choice

  tu = _mod("jax.tree_util")

  def tree_map(fn, tree):
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
      return type(tree)(*(tree_map(fn, t) for t in tree))
    if isinstance(tree, (tuple, list)):
      return type(tree)(tree_map(fn, t) for t in tree)
    if isinstance(tree, dict):
      return {k: tree_map(fn, v) for k, v in tree.items()}
    return fn(tree)
  tu.tree_map = tree_map

  def vmap(fn, in_axes=0):
    """Loop over the leading axis of every argument, stack the results (tuples leaf by leaf)."""
    def mapped(*args):
      n = args[0].shape[0]
      assert all(a.shape[0] == n for a in args)
      outs = [fn(*(a[i] for a in args)) for i in range(n)]
      if isinstance(outs[0], tuple):
        return tuple(torch.stack([o[j] for o in outs]) for j in range(len(outs[0])))
      return torch.stack(outs)
    return mapped
  jax.vmap = vmap
  jax.jit = lambda fn=None, **kw: fn if fn is not None else (lambda f: f)

  def _not_needed(*a, **k):
    raise NotImplementedError("not on the train-step path")
  jax.jacfwd = lambda fn: (lambda x: torch.autograd.functional.jacobian(fn, x))
  jax.eval_shape = _not_needed

  cfg = types.SimpleNamespace(update=lambda *a, **k: None)
  jax.config = cfg

  _mod("jax.scipy")
  stats = _mod("jax.scipy.stats")
  mvn = _mod("jax.scipy.stats.multivariate_normal")

  def pdf(x, mean, cov):
    d = mean.shape[-1]
    diff = (x - mean).to(F64)
    sol = torch.linalg.solve(cov, diff.unsqueeze(-1)).squeeze(-1)
    return torch.exp(-0.5 * (diff * sol).sum(-1)) / torch.sqrt((2.0 * math.pi)**d * torch.linalg.det(cov))
  mvn.pdf = pdf
  stats.multivariate_normal = mvn
  return jax


# ---------------------------------------------------------------------------------------------- haiku
class _Frame:
  params = None
  init = False
  scope = []


def _build_haiku():
  hk = _mod("haiku")
  hk.Params = dict

  def get_parameter(name, shape, dtype=torch.float32, init=None):
    mod = "/".join(_Frame.scope) if _Frame.scope else "~"
    bucket = _Frame.params.setdefault(mod, {}) if _Frame.init else _Frame.params[mod]
    if name not in bucket:
      if not _Frame.init:
        raise KeyError(f"missing parameter {mod}/{name}")
      bucket[name] = init(_shape(shape), dtype)
    p = bucket[name]
    assert tuple(p.shape) == _shape(shape), (mod, name, tuple(p.shape), _shape(shape))
    return p
  hk.get_parameter = get_parameter

  init_mod = _mod("haiku.initializers")

  class RandomNormal:
    def __init__(self, stddev=1.0, mean=0.0):
      self.stddev, self.mean = stddev, mean

    def __call__(self, shape, dtype):
      return torch.randn(shape, dtype=F64).to(dtype) * self.stddev + self.mean

  class TruncatedNormal(RandomNormal):
    def __call__(self, shape, dtype):
      return torch.nn.init.trunc_normal_(torch.empty(shape, dtype=F64), 0.0, 1.0, -2.0, 2.0).to(dtype) * self.stddev
  init_mod.RandomNormal, init_mod.TruncatedNormal = RandomNormal, TruncatedNormal

  class Module:
    def __init__(self, name=None):
      base = name or re.sub(r"(?<!^)(?=[A-Z])", "_", type(self).__name__).lower()
      # a child made while the parent is being constructed is `parent/~/child` (haiku's naming rule)
      self.module_name = "/".join(_Frame.scope + [base])

    def _enter(self):
      self._saved = _Frame.scope
      _Frame.scope = [self.module_name]

    def _exit(self):
      _Frame.scope = self._saved
  hk.Module = Module

  class Linear(Module):
    def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
      super().__init__(name or "linear")
      self.output_size, self.w_init, self.b_init = int(output_size), w_init, b_init

    def __call__(self, x):
      self._enter()
      try:
        fan_in = x.shape[-1]
        w_init = self.w_init or TruncatedNormal(stddev=1.0 / math.sqrt(fan_in))
        b_init = self.b_init or (lambda shape, dtype: torch.zeros(shape, dtype=dtype))
        w = get_parameter("w", (fan_in, self.output_size), x.dtype, init=w_init)
        b = get_parameter("b", (self.output_size, ), x.dtype, init=b_init)
      finally:
        self._exit()
      return x @ w.to(x.dtype) + b.to(x.dtype)
  hk.Linear = Linear

  nets = _mod("haiku.nets")

  class MLP(Module):
    def __init__(self, output_sizes, activation=torch.relu, activate_final=False, name=None):
      super().__init__(name or "mlp")
      saved = _Frame.scope
      _Frame.scope = [self.module_name, "~"]
      self.layers = [Linear(s, name=f"linear_{i}") for i, s in enumerate(output_sizes)]
      _Frame.scope = saved
      self.activation, self.activate_final = activation, activate_final

    def __call__(self, x):
      for i, layer in enumerate(self.layers):
        x = layer(x)
        if i < len(self.layers) - 1 or self.activate_final:
          x = self.activation(x)
      return x
  nets.MLP = MLP

  class Flatten:
    def __init__(self, preserve_dims=1):
      self.p = preserve_dims

    def __call__(self, x):
      keep = x.ndim + self.p if self.p < 0 else self.p   # negative: flatten that many trailing dims
      return x.reshape(tuple(x.shape[:keep]) + (-1, ))

  class Reshape:
    def __init__(self, output_shape, preserve_dims=1):
      self.out, self.p = tuple(int(v) for v in output_shape), preserve_dims

    def __call__(self, x):
      keep = x.ndim + self.p if self.p < 0 else self.p
      return x.reshape(tuple(x.shape[:keep]) + self.out)
  hk.Flatten, hk.Reshape = Flatten, Reshape
  hk.Sequential = type("Sequential", (), {})

  MultiTransformed = namedtuple("MultiTransformed", ["init", "apply"])

  def multi_transform(f):
    def run(fn_getter, params, init, args, kwargs):
      saved = (_Frame.params, _Frame.init, _Frame.scope)
      _Frame.params, _Frame.init, _Frame.scope = params, init, []
      try:
        template, fns = f()
        return fn_getter(template, fns)(*args, **kwargs)
      finally:
        _Frame.params, _Frame.init, _Frame.scope = saved

    def init(rng, *args, **kwargs):
      params = {}
      run(lambda template, fns: template, params, True, args, kwargs)
      if Trainer.init_queue:   # same leaves (names, shapes), caller-chosen values
        preset = Trainer.init_queue.pop(0)
        assert {m: {k: tuple(v.shape) for k, v in lv.items()} for m, lv in preset.items()} == \
               {m: {k: tuple(v.shape) for k, v in lv.items()} for m, lv in params.items()}
        return preset
      return params

    _Frame.params, _Frame.init, _Frame.scope = {}, True, []
    _, fns0 = f()   # structure of the apply namedtuple
    _Frame.params = None

    def make(field):
      def apply(params, rng, *args, **kwargs):
        return run(lambda template, fns: getattr(fns, field), params, False, args, kwargs)
      return apply
    return MultiTransformed(init, type(fns0)(*(make(fld) for fld in fns0._fields)))
  hk.multi_transform = multi_transform

  def without_apply_rng(mt):
    def strip(fn):
      return lambda params, *a, **k: fn(params, None, *a, **k)
    return type(mt)(mt.init, type(mt.apply)(*(strip(fn) for fn in mt.apply)))
  hk.without_apply_rng = without_apply_rng
  return hk


# ---------------------------------------------------------------------------------------------- distrax
def _build_distrax(jax):
  from oracle import rqs as orqs
  distrax = _mod("distrax")
  _mod("distrax._src")
  _mod("distrax._src.bijectors")
  _mod("distrax._src.distributions")
  _mod("distrax._src.utils")
  base = _mod("distrax._src.bijectors.bijector")
  dist_base = _mod("distrax._src.distributions.distribution")
  transformed = _mod("distrax._src.distributions.transformed")
  conversion = _mod("distrax._src.utils.conversion")

  class Bijector:
    def __init__(self, event_ndims_in, event_ndims_out=None, is_constant_jacobian=False, is_constant_log_det=None):
      self._event_ndims_in = event_ndims_in
      self._event_ndims_out = event_ndims_in if event_ndims_out is None else event_ndims_out
      self._is_constant_jacobian = is_constant_jacobian
      self._is_constant_log_det = is_constant_jacobian if is_constant_log_det is None else is_constant_log_det

    event_ndims_in = property(lambda self: self._event_ndims_in)
    event_ndims_out = property(lambda self: self._event_ndims_out)
    is_constant_jacobian = property(lambda self: self._is_constant_jacobian)
    is_constant_log_det = property(lambda self: self._is_constant_log_det)
    name = property(lambda self: type(self).__name__)

    def forward(self, x):
      return self.forward_and_log_det(x)[0]

    def inverse(self, y):
      return self.inverse_and_log_det(y)[0]

    def _check_forward_input_shape(self, x):
      assert x.ndim >= self._event_ndims_in

    def _check_inverse_input_shape(self, y):
      assert y.ndim >= self._event_ndims_out

    def same_as(self, other):
      return other is self
  base.Bijector = Bijector
  base.Array = torch.Tensor
  base.BijectorLike = object
  base.BijectorT = object

  def as_bijector(obj):
    assert isinstance(obj, Bijector), type(obj)
    return obj
  conversion.as_bijector = as_bijector

  dist_base.PRNGKey = int
  dist_base.Array = torch.Tensor
  dist_base.EventT = object
  dist_base.ShapeT = tuple
  dist_base.IntLike = int

  def convert_seed_and_sample_shape(seed, sample_shape):
    return seed, _shape(sample_shape)
  dist_base.convert_seed_and_sample_shape = convert_seed_and_sample_shape

  class Normal:
    def __init__(self, loc, scale):
      self.loc, self.scale = loc, scale
    batch_shape = property(lambda self: tuple(self.loc.shape))
    event_shape = ()
    dtype = F64

    def _sample_n(self, key, n):
      return self.loc + self.scale * jax.random.normal(key, (n, ) + self.batch_shape)

    def log_prob(self, v):
      z = (v - self.loc) / self.scale
      return -0.5 * z * z - 0.5 * math.log(2.0 * math.pi) - torch.log(self.scale)

  class Independent:
    def __init__(self, distribution, reinterpreted_batch_ndims):
      self.d, self.k = distribution, reinterpreted_batch_ndims
    dtype = F64

    def log_prob(self, v):
      lp = self.d.log_prob(v)
      return lp.sum(dim=tuple(range(lp.ndim - self.k, lp.ndim)))

    def sample(self, *, seed, sample_shape=()):
      shape = _shape(sample_shape)
      x = self.d._sample_n(seed, int(np.prod(shape)))
      return x.reshape(shape + tuple(x.shape[1:]))

    def sample_and_log_prob(self, *, seed, sample_shape=()):
      x = self.sample(seed=seed, sample_shape=sample_shape)
      return x, self.log_prob(x)

  class Transformed:
    def __init__(self, distribution, bijector):
      self._distribution, self._bijector = distribution, as_bijector(bijector)
    distribution = property(lambda self: self._distribution)
    bijector = property(lambda self: self._bijector)
  transformed.Transformed = Transformed

  class RationalQuadraticSpline(Bijector):
    """Scalar spline bijector; the algorithm is oracle/rqs.py's restatement of distrax's."""

    def __init__(self, params, range_min, range_max, boundary_slopes="unconstrained", min_bin_size=1e-4,
                 min_knot_slope=1e-4):
      super().__init__(event_ndims_in=0)
      assert boundary_slopes == "unconstrained"
      self.params = params
      self.kw = dict(range_min=float(range_min), range_max=float(range_max), min_bin_size=min_bin_size,
                     min_knot_slope=min_knot_slope)

    def forward_and_log_det(self, x):
      y, ld, _ = orqs.rqs_forward(x, self.params, **self.kw)
      return y, ld

    def inverse_and_log_det(self, y):
      x, ld, _ = orqs.rqs_inverse(y, self.params, **self.kw)
      return x, ld

  distrax.Normal, distrax.Independent, distrax.Transformed = Normal, Independent, Transformed
  distrax.RationalQuadraticSpline = RationalQuadraticSpline
  distrax.Bijector = Bijector
  distrax.Uniform = distrax.UnconstrainedAffine = None
  return distrax


def install():
  """Register the stand-in modules; afterwards `import cnf_ot...` (with /root/reference on sys.path) works."""
  for name in ("jax", "haiku", "distrax", "jaxtyping"):
    assert name not in sys.modules, f"{name} already imported"
  patch_tensor()
  jax = _build_jax()
  _build_haiku()
  _build_distrax(jax)
  jt = _mod("jaxtyping")
  jt.Array = torch.Tensor

  class _Sub:
    def __class_getitem__(cls, item):
      return torch.Tensor
  jt.Float = jt.Int = jt.Bool = _Sub


def stub_plotting():
  """`cnf_ot/utils.py` imports matplotlib at module level for its plotting helpers (not installed here, not needed by
  the two energy functions)."""
  if "matplotlib" in sys.modules:
    return
  mpl = _mod("matplotlib")
  for sub in ("colors", "cm", "pyplot"):
    _mod("matplotlib." + sub)
  mpl.colors.LinearSegmentedColormap = object
  mpl.pyplot.quiver = None   # a bare `plt.quiver` expression statement sits in calc_score_kinetic_energy (utils.py:386)


class Plots:
  """What the reference's plotting helpers hand to matplotlib (cnf_ot/utils.py:572-642), recorded instead of drawn."""
  images = []     # every imshow(array)
  scatters = []   # every scatter(x, y)


def record_plots():
  """A recording `matplotlib.pyplot`: `plot_density_snapshot` / `plot_density_and_trajectory` run unmodified and leave
  the arrays they would draw in `Plots`."""
  stub_plotting()
  plt = sys.modules["matplotlib.pyplot"]
  Plots.images, Plots.scatters = [], []

  class Axes:
    def imshow(self, a, **k):
      Plots.images.append(torch.as_tensor(a).detach().clone())

    def scatter(self, x, y, **k):
      Plots.scatters.append((torch.as_tensor(x).detach().clone(), torch.as_tensor(y).detach().clone()))

    def __getattr__(self, name):   # axis, set_xlabel, set_title, ...
      return lambda *a, **k: None

  class AxesArray(list):
    def flatten(self):
      return self

  class Figure:
    def __getattr__(self, name):
      return lambda *a, **k: None
  ax = Axes()
  plt.imshow, plt.scatter = ax.imshow, ax.scatter
  plt.subplots = lambda r, c, **k: (Figure(), AxesArray([Axes() for _ in range(int(r) * int(c))]))
  for name in ("clf", "figure", "subplot", "axis", "title", "savefig", "legend", "colorbar", "tight_layout"):
    setattr(plt, name, lambda *a, **k: None)


# ---------------------------------------------------------------------------------------------- dr/trainers.py
class Trainer:
  """State shared with the stand-ins that `cnf_ot/dr/trainers.py:train` needs beyond the model code."""
  init_queue = []     # parameter trees handed out by successive `model.init` calls (instead of random draws)
  captured = []       # (loss, grads) of every jax.value_and_grad(loss_fn)(...) call


def stub_trainer():
  """optax / box / ml_collections / flax stand-ins + `jax.value_and_grad`, `jax.tree.leaves`, so that
  `cnf_ot.dr.trainers.train(..., epochs=1)` runs its own `loss_fn` once on parameters we choose: the optimiser
  stand-in applies no update, `model.init` pops the trees queued in `Trainer.init_queue`."""
  stub_plotting()
  import jax
  plt = sys.modules["matplotlib.pyplot"]
  for name in ("plot", "yscale", "savefig", "clf"):
    setattr(plt, name, lambda *a, **k: None)
  _mod("ml_collections").ConfigDict = dict

  class Box(dict):
    def __getattr__(self, k):
      v = self[k]
      return Box(v) if isinstance(v, dict) else v
  _mod("box").Box = Box
  _mod("flax")
  tr = _mod("flax.traverse_util")

  def flatten_dict(d, prefix=()):
    out = {}
    for k, v in d.items():
      out.update(flatten_dict(v, prefix + (k, )) if isinstance(v, dict) else {prefix + (k, ): v})
    return out
  tr.flatten_dict = flatten_dict

  optax = _mod("optax")
  optax.piecewise_constant_schedule = lambda init_value, boundaries_and_scales=None: (lambda step: init_value)
  Opt = namedtuple("GradientTransformation", ["init", "update"])
  optax.adam = lambda lr: Opt(lambda params: None, lambda grads, state, params=None: (None, state))
  optax.apply_updates = lambda params, updates: params

  def value_and_grad(fn):
    def run(params, *args):
      p = jax.tree_util.tree_map(lambda v: v.detach().clone().requires_grad_(True), params)
      loss = fn(p, *args)
      loss.backward()
      grads = jax.tree_util.tree_map(lambda v: v.grad if v.grad is not None else torch.zeros_like(v), p)
      Trainer.captured.append((loss.detach(), grads))
      return loss.detach(), grads
    return run
  jax.value_and_grad = value_and_grad
  tree = _mod("jax.tree")
  Leaf = namedtuple("Leaf", ["size"])
  tree.leaves = lambda t: [Leaf(v.numel()) for v in flatten_dict(t).values()]
