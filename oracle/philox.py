"""CPU restatement (numpy) of the library's counter-based draws (`cnf_ot_b200/csrc/philox.cuh`).

TEST INFRASTRUCTURE ONLY.  Philox4x32-10 is the published generator of Salmon, Moraes, Dror, Shaw,
"Parallel random numbers: as easy as 1, 2, 3" (SC'11); it is pinned here by the known-answer vectors of the
Random123 distribution (tests/test_philox.py).  The reference itself draws with jax.random (threefry), which is
not reproduced: parity tests feed the SAME arrays to the oracle and to the kernels.  What this file pins is that
the arrays `cnfot_philox_rows` / the step kernel produce are exactly this documented function of
(key, step, kind, leading size, row, column).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = (1 << 64) - 1
DRAW_NORMAL, DRAW_UNIFORM, DRAW_CATEGORICAL = 1, 2, 3
ROWS_NORMAL, ROWS_OT_SOURCE = 1, 3


def philox4x32_10(ctr, key):
  """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
  c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
  k = [np.asarray(key[..., i], dtype=np.uint32).copy() for i in range(2)]
  with np.errstate(over="ignore"):
    for _ in range(10):
      p0 = M0 * c[0].astype(np.uint64)
      p1 = M1 * c[2].astype(np.uint64)
      hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
      c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
      k = [k[0] + W0, k[1] + W1]
  return np.stack(c, axis=-1)


def splitmix64(z):
  z = (z + 0x9E3779B97F4A7C15) & MASK
  z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
  z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
  return z ^ (z >> 31)


def salt(kind, n):
  return splitmix64(splitmix64(kind) ^ n)


def draw(key, step, rows, block):
  """Words of counter (row_lo, row_hi, block, step) under `key` for an array of global row indices."""
  rows = np.asarray(rows, dtype=np.uint64)
  ctr = np.stack([(rows & np.uint64(0xFFFFFFFF)).astype(np.uint32), (rows >> np.uint64(32)).astype(np.uint32),
                  np.full(rows.shape, block, dtype=np.uint32), np.full(rows.shape, step, dtype=np.uint32)], axis=-1)
  k = np.broadcast_to(np.array([key & 0xFFFFFFFF, key >> 32], dtype=np.uint32), rows.shape + (2, ))
  return philox4x32_10(ctr, k)


def uniform_open(w):
  return ((w >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def uniform(w):
  return (w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def normal2(wa, wb):
  r = np.sqrt(np.float32(-2.0) * np.log(uniform_open(wa))).astype(np.float32)
  th = (np.float32(2.0) * uniform(wb)).astype(np.float64) * np.pi   # sincospi(2 u)
  return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


CENTRES = np.array([[0, 5], [5, 0], [0, -5], [-5, 0], [3, 4], [3, -4], [-3, -4], [-3, 4]], dtype=np.float32)


def rows(key, step, source, global_rows, dim, row0=0, n=None):
  """The (n, dim) float32 block [row0, row0 + n) of the (global_rows, dim) draw."""
  n = global_rows - row0 if n is None else n
  idx = np.arange(row0, row0 + n, dtype=np.uint64)
  kn = key ^ salt(DRAW_NORMAL, global_rows)
  out = np.zeros((n, dim), dtype=np.float32)
  for cb in range((dim + 3) // 4):
    w = draw(kn, step, idx, cb)
    z0, z1 = normal2(w[:, 0], w[:, 1])
    z2, z3 = normal2(w[:, 2], w[:, 3])
    for j, z in enumerate((z0, z1, z2, z3)):
      if cb * 4 + j < dim:
        out[:, cb * 4 + j] = z
  if source == ROWS_OT_SOURCE:
    if dim == 2:
      kc = key ^ salt(DRAW_CATEGORICAL, global_rows)
      out += CENTRES[draw(kc, step, idx, 0)[:, 0] >> np.uint32(29)]
    else:
      out -= np.float32(3.0)
  return out


def times(key, step, n_t, horizon):
  kt = key ^ salt(DRAW_UNIFORM, n_t)
  return uniform(draw(kt, step, np.arange(n_t, dtype=np.uint64), 0)[:, 0]) * np.float32(horizon)
