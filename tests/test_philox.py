"""The counter-based draws of the step (cnf_ot_b200/csrc/philox.cuh), CPU side: the generator against the
published Philox4x32-10 known-answer vectors (Random123 kat_vectors), the numpy restatement (oracle/philox.py)
against the header compiled for the host (tests/hostsim), and the host entry of the C ABI."""
import numpy as np
import torch

import hostsim as hs
from oracle import philox as op

# Random123 known-answer tests for philox4x32-10: (counter, key) -> output
KAT = [
  ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
  ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
  ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_known_answer_vectors():
  for ctr, key, want in KAT:
    got_np = op.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
    got_hs = hs.philox_words(ctr, key)
    assert [int(v) for v in got_np] == want, [hex(int(v)) for v in got_np]
    assert [int(v) for v in got_hs] == want, [hex(int(v)) for v in got_hs]


def test_rows_numpy_restatement_matches_the_header():
  key = 0x1234_5678_9ABC_DEF0
  for source, dim, n in ((op.ROWS_NORMAL, 2, 1000), (op.ROWS_OT_SOURCE, 2, 1000), (op.ROWS_NORMAL, 10, 257),
                         (op.ROWS_OT_SOURCE, 3, 64), (op.ROWS_NORMAL, 32, 33)):
    a = op.rows(key, 5, source, n, dim)
    b = hs.philox_rows(key, 5, source, n, dim).numpy()
    assert a.shape == b.shape and float(np.abs(a - b).max()) < 5e-6   # libm vs numpy log / sincos
    # a shard is the corresponding block of the whole draw; another leading size is another draw
    c = hs.philox_rows(key, 5, source, n, dim, row0=17, rows=12).numpy()
    assert np.array_equal(c, b[17:29])
    d = hs.philox_rows(key, 5, source, n + 1, dim, rows=n).numpy()
    assert float(np.abs(d - b).max()) > 0.1
  # moments of the normal draw
  z = hs.philox_rows(99, 0, op.ROWS_NORMAL, 200000, 2).double()
  assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1) < 0.01 and abs(float((z[:, 0] * z[:, 1]).mean())) < 0.01
  # mixture: z + centre with the SAME z as the normal draw of that shape (applications.py:81-82)
  s = hs.philox_rows(99, 0, op.ROWS_OT_SOURCE, 200000, 2)
  cen = (s - z.float()).round()
  assert set(map(tuple, cen.int().tolist())) == {(0, 5), (5, 0), (0, -5), (-5, 0), (3, 4), (3, -4), (-3, -4), (-3, 4)}
  counts = torch.unique(cen, dim=0, return_counts=True)[1].double() / 200000
  assert float((counts - 0.125).abs().max()) < 0.005


def test_times_and_c_abi_host_entry():
  from cnf_ot_b200 import ops
  key = 424242
  for n_t, hor in ((1, 1.0), (2, 2.0), (7, 0.5)):
    a = op.times(key, 3, n_t, hor)
    b = hs.philox_times(key, 3, n_t, hor).numpy()
    c = np.array(ops.philox_times(key, 3, n_t, hor), dtype=np.float32)
    assert np.array_equal(a, b) and np.array_equal(b, c)
    assert (a >= 0).all() and (a < hor).all()


def test_python_key_mixer_is_the_library_salt():
  """cnf_ot_b200.random's splitmix64 finaliser and the library's philox_salt are the same function."""
  import ctypes
  from cnf_ot_b200 import random as crandom
  f = hs.lib("philox").hs_philox_salt
  f.restype = ctypes.c_uint64
  f.argtypes = [ctypes.c_int, ctypes.c_uint64]
  for kind, n in ((1, 4096), (2, 1), (3, 262144)):
    assert f(kind, n) == op.salt(kind, n) == crandom._mix(crandom._mix(kind) ^ n)
