"""JAX side of the XLA FFI shim (cnf_ot_b200/csrc/xla_ffi_shim.cc): registration of the custom-call targets and the
`jax.custom_vjp` wrappers a maintainer substitutes at the reference's seams:

  seam 1  the `Flow` namedtuple of `RQSFlow(...)`          /root/reference/cnf_ot/models/flows.py:213-226
  seam 3  `jax.value_and_grad(loss_fn)` + Adam in `update`  /root/reference/cnf_ot/mfc/solvers.py:90-97

JAX is not installable in the build image (no wheels, no network), so nothing here runs in this repository's tests: the
tested binding is the ctypes one (`_lib.py`, `ops.py`), which calls the same C symbols with the same arguments, and the
shim is compiled against a stand-in of the FFI header (tests/test_ffi_shim.py).  Importing this module without jax
raises ImportError; the rest of the package never imports it.

    from cnf_ot_b200 import jax_ffi
    jax_ffi.register("cnf_ot_b200/libcnfot_xla.so")
    flow = jax_ffi.flow_api(static)          # drop-in for the namedtuple RQSFlow returns
    loss, grad_blob = jax_ffi.value_and_grad(problem, static)(blob, key, step, _lambda, batch_size)
"""
from __future__ import annotations

import ctypes
from functools import partial

import jax            # noqa: F401  (ImportError here is the documented behaviour without jax)
import jax.numpy as jnp
import numpy as np

HANDLERS = ("CnfotRqsForward", "CnfotRqsInverse", "CnfotRqsForwardVjp", "CnfotRqsInverseVjp", "CnfotFlowForward",
            "CnfotFlowInverse", "CnfotFlowForwardVjp", "CnfotFlowInverseVjp", "CnfotMfcStep", "CnfotMfcStepRng",
            "CnfotMfcUpdate", "CnfotAdam", "CnfotKineticEnergy", "CnfotDensityGrid", "CnfotDensityMc")
_lib = None
_abi = None


def register(shim_path: str, abi_path: str = None):
  """Load libcnfot_xla.so (the shim linked with libcnfot.so) and register every handler for the CUDA platform."""
  global _lib, _abi
  _lib = ctypes.CDLL(shim_path)
  _abi = ctypes.CDLL(abi_path) if abi_path else _lib
  for name in HANDLERS:
    jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(_lib, name)), platform="CUDA")


def _static(dim, num_layers, mlp_layers, hidden, num_bins):
  return dict(dim=np.int64(dim), num_layers=np.int64(num_layers), mlp_layers=np.int64(mlp_layers), hidden=np.int64(hidden),
              num_bins=np.int64(num_bins))


def _scratch(nbytes):
  return jax.ShapeDtypeStruct((max(int(nbytes), 1), ), jnp.uint8)


def _flow_desc(static):
  from ._lib import FlowDesc
  return FlowDesc(int(static["dim"]), int(static["num_layers"]), int(static["mlp_layers"]), int(static["hidden"]),
                  int(static["num_bins"]), -10.0, 10.0, 1e-4, 1e-4)


def _ws(fn_name, static, *args):
  fn = getattr(_abi, fn_name)
  fn.restype = ctypes.c_int64
  return fn(ctypes.byref(_flow_desc(static)), *args)


# ------------------------------------------------------------------ seam 1: flow.forward / inverse with custom VJPs
def _flow_call(name, static, add_base, blob, x, cond):
  rows = x.shape[0]
  ws = _ws("cnfot_flow_workspace_bytes", static, ctypes.c_int64(rows))
  y, ld, _ = jax.ffi.ffi_call(name, (jax.ShapeDtypeStruct(x.shape, jnp.float32), jax.ShapeDtypeStruct((rows, ), jnp.float32),
                                     _scratch(ws)), vmap_method="broadcast_all")(
      blob, x.astype(jnp.float32), jnp.asarray(cond, jnp.float32).reshape(-1), add_base=np.int64(add_base), **static)
  return y, ld


def _flow_vjp_call(name, static, add_base, blob, x, cond, g_y, g_ld):
  rows = x.shape[0]
  ws = _ws("cnfot_flow_vjp_workspace_bytes", static, ctypes.c_int64(rows))
  g_x, g_blob, _ = jax.ffi.ffi_call(name, (jax.ShapeDtypeStruct(x.shape, jnp.float32),
                                           jax.ShapeDtypeStruct(blob.shape, jnp.float32), _scratch(ws)))(
      blob, x.astype(jnp.float32), jnp.asarray(cond, jnp.float32).reshape(-1), g_y.astype(jnp.float32),
      g_ld.astype(jnp.float32), add_base=np.int64(add_base), **static)
  return g_blob, g_x


def _make_flow_fn(direction: str, static, add_base: int):
  fwd_name, vjp_name = f"CnfotFlow{direction}", f"CnfotFlow{direction}Vjp"

  @jax.custom_vjp
  def fn(blob, x, cond):
    return _flow_call(fwd_name, static, add_base, blob, x, cond)

  def fwd(blob, x, cond):
    return fn(blob, x, cond), (blob, x, cond)

  def bwd(res, g):
    blob, x, cond = res
    g_blob, g_x = _flow_vjp_call(vjp_name, static, add_base, blob, x, cond, g[0], g[1])
    return g_blob, g_x, jnp.zeros_like(jnp.asarray(cond, jnp.float32))   # d/dt is not needed by the MFC losses

  fn.defvjp(fwd, bwd)
  return fn


def flow_api(static):
  """The callables `model.apply.*` resolves to (flows.py:213-226), on the parameter BLOB (layout: cnf_ot_b200/layout.py,
  `cnfot_offset_linear`).  static = dict(dim=, num_layers=, mlp_layers=, hidden=, num_bins=)."""
  static = _static(**{k: int(v) for k, v in static.items()})
  forward = _make_flow_fn("Forward", static, 0)
  inverse = _make_flow_fn("Inverse", static, 0)
  forward_lp = _make_flow_fn("Forward", static, 1)
  inverse_lp = _make_flow_fn("Inverse", static, 1)
  dim = int(static["dim"])

  def sample(blob, *, cond, seed, sample_shape):
    z = jax.random.normal(seed, sample_shape + (dim, ), jnp.float32)
    return forward(blob, z, cond)[0]

  def sample_and_log_prob(blob, *, cond, seed, sample_shape):
    z = jax.random.normal(seed, sample_shape + (dim, ), jnp.float32)
    return forward_lp(blob, z, cond)

  return dict(log_prob=lambda blob, value, cond: inverse_lp(blob, value, cond)[1],
              sample=sample, sample_and_log_prob=sample_and_log_prob,
              forward=lambda blob, x, c: forward(blob, x, c)[0], inverse=lambda blob, y, c: inverse(blob, y, c)[0])


# ------------------------------------------------------------------ seam 3: the train step
def _problem(problem):
  return dict(type=np.int64(problem["type"]), subtype=np.int64(problem["subtype"]), T=np.float32(problem["T"]),
              beta=np.float32(problem["beta"]), a=np.float32(problem["a"]), sigma=np.float32(problem["sigma"]),
              dt=np.float32(problem["dt"]), dx=np.float32(problem["dx"]))


def value_and_grad(problem, static, n_t: int = 1):
  """jax.value_and_grad(loss_fn) of solvers.py:94 as ONE custom call with the draws made on chip from (key, step):
  f(blob, key, step, _lambda, batch_size) -> (loss, gradient blob).  key / step are Python ints (static under jit)."""
  static = _static(**{k: int(v) for k, v in static.items()})
  prob = _problem(problem)

  def f(blob, key: int, step: int, _lambda: float, batch_size: int):
    B, b = int(batch_size), int(batch_size) // 32
    ws = _ws("cnfot_mfc_step_workspace_bytes", static, ctypes.c_int64(B), ctypes.c_int64(b), ctypes.c_int32(n_t))
    n = blob.shape[0]
    out, _ = jax.ffi.ffi_call("CnfotMfcStepRng", (jax.ShapeDtypeStruct((n + 8, ), jnp.float32), _scratch(ws)))(
        blob, key=np.int64(key), step=np.int64(step), n_t=np.int64(n_t), row0_B=np.int64(0), rows_B=np.int64(B),
        row0_b=np.int64(0), rows_b=np.int64(b), global_B=np.int64(B), global_b=np.int64(b), **static, **prob,
        **{"lambda": np.float32(_lambda)})
    return out[n], out[:n]

  return f


def train_state_init(static, key: int, step: int = 0):
  """Device train state of `update` (cnfot_train_state_init): returns a uint8 jax array to thread through `update`."""
  nbytes = _ws("cnfot_train_state_bytes", static)
  state = jnp.zeros((nbytes, ), jnp.uint8)
  state.block_until_ready()
  fn = _abi.cnfot_train_state_init
  fn.restype = ctypes.c_int32
  rc = fn(None, ctypes.byref(_flow_desc(static)), ctypes.c_void_p(state.unsafe_buffer_pointer()), ctypes.c_int64(nbytes),
          ctypes.c_uint64(key), ctypes.c_uint64(step), ctypes.c_uint32(1))
  if rc != 0:
    raise RuntimeError("cnfot_train_state_init failed")
  return state


def update(problem, static, lr: float, n_t: int = 1, b1=0.9, b2=0.999, eps=1e-8):
  """`update` of solvers.py:90-97 as ONE kernel launch: (state, blob, mu, nu) -> (state, blob, mu, nu, out), all four
  updated in place (input_output_aliases); out = [gradient | loss slots] of the step."""
  static = _static(**{k: int(v) for k, v in static.items()})
  prob = _problem(problem)

  @partial(jax.jit, static_argnums=(4, 5), donate_argnums=(0, 1, 2, 3))
  def f(state, blob, mu, nu, _lambda: float, batch_size: int):
    B, b = int(batch_size), int(batch_size) // 32
    n = blob.shape[0]
    shapes = (jax.ShapeDtypeStruct(state.shape, jnp.uint8), jax.ShapeDtypeStruct(blob.shape, jnp.float32),
              jax.ShapeDtypeStruct(mu.shape, jnp.float32), jax.ShapeDtypeStruct(nu.shape, jnp.float32),
              jax.ShapeDtypeStruct((n + 8, ), jnp.float32))
    return jax.ffi.ffi_call("CnfotMfcUpdate", shapes, input_output_aliases={0: 0, 1: 1, 2: 2, 3: 3})(
        state, blob, mu, nu, n_t=np.int64(n_t), row0_B=np.int64(0), rows_B=np.int64(B), row0_b=np.int64(0),
        rows_b=np.int64(b), global_B=np.int64(B), global_b=np.int64(b), lr=np.float32(lr), b1=np.float32(b1),
        b2=np.float32(b2), eps=np.float32(eps), **static, **prob, **{"lambda": np.float32(_lambda)})

  return f
