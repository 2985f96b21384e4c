"""Counter-style PRNG keys with JAX-like value semantics.

The reference threads `jax.random.PRNGKey`s through its losses and relies on
"equal key => equal draw" (every sampler of one loss call receives the same
`rng`, /root/reference/cnf_ot/mfc/applications.py:81-82,233-239,253-263).  A
`Key` is an immutable 64-bit value; draws are made on the target device by a
torch Philox generator seeded from the key and the requested shape, so the same
(key, shape, device type) always yields the same numbers.  The streams are NOT
the ones jax.random would produce -- parity tests pass explicit arrays instead.
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


def normal(key, shape, device="cuda", dtype=torch.float32) -> torch.Tensor:
  key = as_key(key)
  return torch.randn(tuple(shape), generator=_generator(key, shape, device, 1), device=device, dtype=dtype)


def uniform(key, shape, device="cpu", dtype=torch.float32) -> torch.Tensor:
  key = as_key(key)
  return torch.rand(tuple(shape), generator=_generator(key, shape, device, 2), device=device, dtype=dtype)


def randint(key, shape, high: int, device="cuda") -> torch.Tensor:
  key = as_key(key)
  return torch.randint(0, high, tuple(shape), generator=_generator(key, shape, device, 3), device=device)
