"""GPU parity: cnfot_mfc_step (C ABI) vs the oracle's value_and_grad of
ot_loss_fn / rwpo_loss_fn / fp_loss_fn (applications.py:377-441)."""
import pytest
import torch

from cnf_ot_b200 import ops
from cnf_ot_b200.layout import pack
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of

pytestmark = pytest.mark.gpu

# float32 tolerances (stated per north_star): loss relative 2e-5; gradient max-abs error relative
# to the largest gradient entry 5e-5.  The finite-difference terms (1/dt = 1/dx = 100) set the floor.
TOL_LOSS, TOL_GRAD = 2e-5, 5e-5

CASES = [
  ("ot", "free", {}), ("ot", "obstacle", {}), ("rwpo", "quadratic", {}), ("rwpo", "double_well", {}),
  ("fp", "gradient", {}), ("fp", "nongradient", {}),
  ("fp", "lorenz", dict(dim=3, L=3, sigma=0.1)),
  ("fp", "nongradient", dict(dim=10, sigma=0.05, B=320)),
  ("rwpo", "double_well", dict(dim=4, H=32, K=8, sigma=0.1, B=320)),
  ("ot", "obstacle", dict(M=1)), ("rwpo", "quadratic", dict(M=3)),
  ("ot", "obstacle", dict(dim=3, H=32, sigma=0.1, B=320)), ("fp", "lorenz", dict(dim=3, H=32, sigma=0.1, B=320)),
  ("ot", "free", dict(dim=2, H=64, sigma=0.1, B=320)),   # hidden 64 fits the fused kernels at small dim
]


def run_gpu(cfg, shape, params, inputs, lam, rows=None, sub_rows=None):
  B = cfg["train"]["batch_size"]
  b = B // 32
  W = pack(shape, params).cuda()
  f = lambda t: t.float().cuda()
  rs = slice(0, B) if rows is None else rows
  ss = slice(0, b) if sub_rows is None else sub_rows
  typ = cfg["general"]["type"]
  out = ops.mfc_step(shape, ops.problem_desc(cfg), W,
                     None if typ == "ot" else f(inputs["latent"][rs]), f(inputs["latent"][:b][ss]),
                     f(inputs["src"][rs]) if typ == "ot" else None,
                     f(inputs["tgt"][rs]) if typ == "ot" else None,
                     inputs["t_batch"].tolist(), lam, B, b)
  return out.cpu().double()


@pytest.mark.parametrize("typ,sub,kw", CASES)
def test_loss_and_gradient(typ, sub, kw, engine):
  from cnf_ot_b200 import _lib
  kw = dict(kw)
  sigma = kw.pop("sigma", 0.3)
  cfg = make_cfg(typ, sub, Tn=2, lam=500.0, **({"B": 1024 + 64} | kw))  # ragged tiles
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  inputs = make_inputs(cfg)
  loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
  Gor = pack(shape, grads, torch.float64)
  out = run_gpu(cfg, shape, params, inputs, 500.0)
  ran = _lib.last_launch_info()["engine"]
  wide16 = shape.hidden == 16 and shape.num_bins == 5  # larger flows stream weights + fragments
  assert ran == (engine if wide16 else "cuda"), ran
  G, slots = out[:shape.blob_size], out[shape.blob_size:]
  assert abs(float(slots[0]) - float(loss)) <= TOL_LOSS * abs(float(loss)), (float(slots[0]), float(loss))
  assert abs(float(slots[1:5].sum()) - float(slots[0])) <= 1e-5 * abs(float(slots[0]))
  assert float((G - Gor).abs().max() / Gor.abs().max()) <= TOL_GRAD


@pytest.mark.parametrize("typ,sub,kw", [("ot", "free", {}), ("ot", "obstacle", {}), ("rwpo", "double_well", {}),
                                        ("fp", "nongradient", {}), ("fp", "lorenz", dict(dim=3, L=3, sigma=0.1)),
                                        ("fp", "nongradient", dict(dim=10, sigma=0.05, B=320)),
                                        ("rwpo", "quadratic", dict(dim=4, H=32, K=8, sigma=0.1, B=320))])
def test_kinetic_row_forms_agree(typ, sub, kw, monkeypatch):
  """The two forms of the kinetic rows -- one thread per row (row_kinetic: what large batches run) and the passes of a
  row spread over a lane group (row_kinetic_split: what small steps run) -- against the oracle and against each other."""
  kw = dict(kw)
  sigma = kw.pop("sigma", 0.3)
  cfg = make_cfg(typ, sub, Tn=2, lam=500.0, **({"B": 1024 + 64} | kw))
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  inputs = make_inputs(cfg)
  loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
  Gor = pack(shape, grads, torch.float64)
  outs = {}
  for form in ("0", "1"):
    monkeypatch.setenv("CNFOT_KINETIC_SPLIT", form)
    out = run_gpu(cfg, shape, params, inputs, 500.0)
    G, slots = out[:shape.blob_size], out[shape.blob_size:]
    assert abs(float(slots[0]) - float(loss)) <= TOL_LOSS * abs(float(loss)), (form, float(slots[0]), float(loss))
    assert float((G - Gor).abs().max() / Gor.abs().max()) <= TOL_GRAD, form
    outs[form] = out
  d = (outs["0"] - outs["1"])[:shape.blob_size].abs().max() / Gor.abs().max()
  assert float(d) <= TOL_GRAD


def test_shards_sum_to_whole_batch():
  """Data-parallel contract: out buffers of row shards add up to the
  whole-batch loss and gradient."""
  cfg = make_cfg("rwpo", "double_well", B=4096, lam=100.0)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  inputs = make_inputs(cfg)
  whole = run_gpu(cfg, shape, params, inputs, 100.0)
  B, b = 4096, 128
  parts = [run_gpu(cfg, shape, params, inputs, 100.0, rows=slice(i * B // 4, (i + 1) * B // 4),
                   sub_rows=slice(i * b // 4, (i + 1) * b // 4)) for i in range(4)]
  tot = sum(parts)
  assert float((tot - whole).abs().max() / whole.abs().max()) < 2e-6


def test_persistent_workspace_and_plain_workspace_agree(monkeypatch):
  """C ABI: a workspace registered with cnfot_workspace_register is left clean by every step (no memset between two
  calls, back-to-back launches with programmatic stream serialisation); an unregistered one is zeroed per call.  Same
  results either way, with and without CNFOT_PDL, for consecutive calls of different batch sizes on the same buffer."""
  import ctypes
  from cnf_ot_b200 import _lib
  lib = _lib.load()
  cfg = make_cfg("ot", "obstacle", B=4096)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  inputs = make_inputs(cfg)
  desc, prob = _lib.flow_desc(shape), ops.problem_desc(cfg)
  W = pack(shape, params).cuda()
  f = lambda t: t.float().cuda().contiguous()
  sub, src, tgt = f(inputs["latent"][:128]), f(inputs["src"]), f(inputs["tgt"])
  tb = torch.tensor([0.3], dtype=torch.float32)
  nbytes = lib.cnfot_mfc_step_workspace_bytes(desc, 4096, 128, 1)
  stream = torch.cuda.current_stream().cuda_stream

  def call(ws, rows, out):
    _lib.check(lib.cnfot_mfc_step(stream, desc, prob, W.data_ptr(), None, sub[:rows // 32].data_ptr(), src[:rows].data_ptr(),
                                  tgt[:rows].data_ptr(), tb.data_ptr(), 1, rows, rows // 32, rows, rows // 32, 5000.0,
                                  out.data_ptr(), ws.data_ptr(), ws.numel()))

  plain = torch.full((nbytes,), 0x5A, dtype=torch.uint8, device="cuda")   # garbage: the call must clear what it uses
  ref = {}
  for rows in (4096, 1024):
    ref[rows] = torch.empty(shape.blob_size + 8, device="cuda")
    call(plain, rows, ref[rows])
  torch.cuda.synchronize()
  ws = torch.full((nbytes,), 0x5A, dtype=torch.uint8, device="cuda")
  _lib.check(lib.cnfot_workspace_register(stream, desc, ws.data_ptr(), ws.numel()))
  try:
    for pdl in ("1", "0"):
      monkeypatch.setenv("CNFOT_PDL", pdl)   # read once per process: this only documents that both paths are legal
      for rows in (4096, 1024, 4096, 4096, 1024):
        out = torch.empty(shape.blob_size + 8, device="cuda")
        call(ws, rows, out)
        torch.cuda.synchronize()
        assert float((out - ref[rows]).abs().max() / ref[rows].abs().max()) < 2e-6, (pdl, rows)
  finally:
    lib.cnfot_workspace_release(ws.data_ptr())
  out = torch.empty(shape.blob_size + 8, device="cuda")
  call(ws, 4096, out)   # released: zeroed per call again
  torch.cuda.synchronize()
  assert float((out - ref[4096]).abs().max() / ref[4096].abs().max()) < 2e-6


def test_host_entry_matches_device_entry():
  cfg = make_cfg("ot", "obstacle", B=2048)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  inputs = make_inputs(cfg)
  dev = run_gpu(cfg, shape, params, inputs, 5000.0)
  b = 2048 // 32
  pin = lambda t: t.float().contiguous().pin_memory()
  out = torch.empty(shape.blob_size + 8, dtype=torch.float32).pin_memory()
  ops.mfc_step_host(shape, ops.problem_desc(cfg), pin(pack(shape, params)), None,
                    pin(inputs["latent"][:b]), pin(inputs["src"]), pin(inputs["tgt"]),
                    inputs["t_batch"].tolist(), 5000.0, 2048, b, out)
  # same kernels behind both entries; the CTAs' partial sums may be combined in a different order
  assert float((out.double() - dev).abs().max() / dev.abs().max()) < 2e-6


def test_host_entry_pageable_and_pinned_rows_on_device_entry(monkeypatch):
  """Pageable host rows take the staged-copy path of the host entry (pinned ones are read in place);
  the device entry accepts pinned host rows directly.  All three agree with the device-resident run."""
  cfg = make_cfg("rwpo", "double_well", B=2048)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  inputs = make_inputs(cfg)
  dev = run_gpu(cfg, shape, params, inputs, 50.0)
  b = 2048 // 32
  host = lambda t: t.float().contiguous().clone()
  W = pack(shape, params)
  lat, sub = host(inputs["latent"]), host(inputs["latent"][:b])
  tol = lambda o: float((o.double().cpu() - dev).abs().max() / dev.abs().max())
  out = torch.empty(shape.blob_size + 8, dtype=torch.float32)
  ops.mfc_step_host(shape, ops.problem_desc(cfg), W, lat, sub, None, None, inputs["t_batch"].tolist(), 50.0,
                    2048, b, out)
  assert tol(out) < 2e-6
  monkeypatch.setenv("CNFOT_HOST_ZEROCOPY", "0")
  out2 = torch.empty(shape.blob_size + 8, dtype=torch.float32).pin_memory()
  ops.mfc_step_host(shape, ops.problem_desc(cfg), W.pin_memory(), lat.pin_memory(), sub.pin_memory(), None, None,
                    inputs["t_batch"].tolist(), 50.0, 2048, b, out2)
  assert tol(out2) < 2e-6
  out3 = ops.mfc_step(shape, ops.problem_desc(cfg), W.cuda(), lat.pin_memory(), sub.pin_memory(), None, None,
                      inputs["t_batch"].tolist(), 50.0, 2048, b)
  assert tol(out3) < 2e-6


def test_deterministic_and_zero_rows():
  cfg = make_cfg("fp", "nongradient", B=512)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  inputs = make_inputs(cfg)
  a = run_gpu(cfg, shape, params, inputs, 10.0)
  assert bool(torch.isfinite(a).all())
  # an empty shard contributes exactly zero
  z = run_gpu(cfg, shape, params, inputs, 10.0, rows=slice(0, 0), sub_rows=slice(0, 0))
  assert float(z.abs().max()) == 0.0


def test_adam_matches_optax_formula():
  g = torch.Generator().manual_seed(0)
  n = 5000
  p = torch.randn(n, generator=g)
  m = torch.zeros(n)
  v = torch.zeros(n)
  pd, md, vd = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
  p64, m64, v64 = p.double(), m.double(), v.double()
  lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
  for step in range(1, 6):
    grad = torch.randn(n, generator=g)
    ops.adam_update(pd, grad.cuda(), md, vd, lr, step)
    g64 = grad.double()
    m64 = b1 * m64 + (1 - b1) * g64
    v64 = b2 * v64 + (1 - b2) * g64 * g64
    p64 = p64 - lr * (m64 / (1 - b1**step)) / (torch.sqrt(v64 / (1 - b2**step)) + eps)
  assert float((pd.cpu().double() - p64).abs().max()) < 1e-6


@pytest.mark.parametrize("typ,sub,kw", [("ot", "obstacle", {}), ("rwpo", "double_well", {}),
                                        ("fp", "nongradient", {}), ("fp", "gradient", {}),
                                        ("ot", "free", dict(M=3)), ("ot", "obstacle", dict(M=1))])
def test_tcgen05_engine_variant(typ, sub, kw, monkeypatch):
  """The opt-in variant whose hidden/output linears run on tcgen05 (3xTF32, TMEM accumulators)
  meets the same tolerances as the CUDA-core layers."""
  from cnf_ot_b200 import _lib
  monkeypatch.setenv("CNFOT_ENGINE", "tc")
  kw = dict(kw)
  sigma = kw.pop("sigma", 0.3)
  cfg = make_cfg(typ, sub, Tn=2, lam=500.0, **({"B": 1024 + 64} | kw))
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  inputs = make_inputs(cfg)
  loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
  Gor = pack(shape, grads, torch.float64)
  out = run_gpu(cfg, shape, params, inputs, 500.0)
  assert _lib.last_launch_info()["engine"] == "tc"
  G, slots = out[:shape.blob_size], out[shape.blob_size:]
  assert abs(float(slots[0]) - float(loss)) <= TOL_LOSS * abs(float(loss))
  assert float((G - Gor).abs().max() / Gor.abs().max()) <= TOL_GRAD
