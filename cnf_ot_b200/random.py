"""Counter-style PRNG keys with JAX-like value semantics.

The reference threads `jax.random.PRNGKey`s through its losses and relies on
"equal key => equal draw" (every sampler of one loss call receives the same
`rng`, /root/reference/cnf_ot/mfc/applications.py:81-82,233-239,253-263).  A
`Key` is an immutable 64-bit value.  Draws on a CUDA device are the library's
counter-based Philox draws (cnf_ot_b200/csrc/philox.cuh, `cnfot_philox_*`): a pure
function of (key, kind and leading size of the array, row, column) -- the SAME numbers
the fused step kernel generates on chip for that key, so `loss_fn(params, rng, ...)`
evaluated term by term and `value_and_grad(loss_fn)(params, rng, ...)` see identical
samples.  Like in jax.random, arrays of different shapes drawn from one key are
unrelated.  The streams are NOT the ones jax.random would produce -- parity tests
pass explicit arrays instead.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Tuple

import torch

_MASK = (1 << 64) - 1


def _mix(z: int) -> int:  # splitmix64 finaliser
  z = (z + 0x9E3779B97F4A7C15) & _MASK
  z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
  z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
  return z ^ (z >> 31)


@dataclass(frozen=True)
class Key:
  value: int


def PRNGKey(seed: int) -> Key:
  return Key(_mix(int(seed) & _MASK))


def split(key: Key, num: int = 2) -> Tuple[Key, ...]:
  return tuple(Key(_mix(key.value ^ _mix(i + 1))) for i in range(num))


def as_key(seed) -> Key:
  return seed if isinstance(seed, Key) else PRNGKey(int(seed))


def _generator(key: Key, shape: Sequence[int], device, salt: int) -> torch.Generator:
  g = torch.Generator(device=device)
  s = key.value ^ _mix(salt)
  for d in shape:
    s = _mix(s ^ int(d))
  g.manual_seed(s & ((1 << 63) - 1))
  return g


def _philox_ok(shape, device) -> bool:
  return len(shape) == 2 and 1 <= shape[1] <= 32 and torch.device(device).type == "cuda"


def normal(key, shape, device="cuda", dtype=torch.float32) -> torch.Tensor:
  key = as_key(key)
  if _philox_ok(shape, device) and dtype == torch.float32:
    from . import _lib, ops
    return ops.philox_rows(key.value, 0, _lib.ROWS_NORMAL, int(shape[0]), int(shape[1]), torch.device(device))
  return torch.randn(tuple(shape), generator=_generator(key, shape, device, 1), device=device, dtype=dtype)


def ot_source(key, shape, device="cuda") -> torch.Tensor:
  """The source batch of kl_loss_fn (applications.py:28-71) for `key`: z + mixture centre (dim 2) or z - 3,
  z = normal(key, shape)."""
  from . import _lib, ops
  key = as_key(key)
  return ops.philox_rows(key.value, 0, _lib.ROWS_OT_SOURCE, int(shape[0]), int(shape[1]), torch.device(device))


def uniform(key, shape, device="cpu", dtype=torch.float32) -> torch.Tensor:
  key = as_key(key)
  if len(shape) == 1 and dtype == torch.float32:
    # the draws the step kernel makes for its times (host-side restatement of the same Philox stream)
    from . import ops
    return torch.tensor(ops.philox_times(key.value, 0, int(shape[0]), 1.0), dtype=dtype).to(device)
  return torch.rand(tuple(shape), generator=_generator(key, shape, device, 2), device=device, dtype=dtype)


def randint(key, shape, high: int, device="cuda") -> torch.Tensor:
  key = as_key(key)
  return torch.randint(0, high, tuple(shape), generator=_generator(key, shape, device, 3), device=device)
