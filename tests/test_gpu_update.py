"""GPU: on-chip draws, the folded step kernel and the device-resident update (cnfot_mfc_step_rng,
cnfot_mfc_update) -- SURVEY.md section 8f row 1: `update` of cnf_ot/mfc/solvers.py:90-106.

  * cnfot_philox_rows on the device == the same header compiled for the host (tests/hostsim) == oracle/philox.py
  * the step with on-chip draws == the step on the exported arrays (same kernels, same numbers)
  * K fused updates (value_and_grad + Adam in one launch) follow the trajectory of oracle + float64 Adam on the
    exported draws: loss curve and final parameters, for ot / rwpo / fp
  * a CUDA graph of updates replays to the same trajectory; solvers.main trains device-resident
"""
import math

import numpy as np
import pytest
import torch

import hostsim as hs
from cnf_ot_b200 import _lib, applications, ops, random, solvers
from cnf_ot_b200.layout import pack, unpack
from oracle import losses as olosses
from oracle import philox as op
from util import make_cfg, make_params, shape_of

pytestmark = pytest.mark.gpu


def test_device_rows_match_host_header_and_numpy():
  key = 0xDEADBEEF12345678
  for source, dim, n in ((_lib.ROWS_NORMAL, 2, 4097), (_lib.ROWS_OT_SOURCE, 2, 4097), (_lib.ROWS_NORMAL, 10, 513),
                         (_lib.ROWS_OT_SOURCE, 5, 100)):
    d = ops.philox_rows(key, 9, source, n, dim, "cuda").cpu()
    h = hs.philox_rows(key, 9, source, n, dim)
    a = torch.from_numpy(op.rows(key, 9, source, n, dim))
    assert float((d - h).abs().max()) < 5e-6 and float((d - a).abs().max()) < 5e-6
    part = ops.philox_rows(key, 9, source, n, dim, "cuda", rows=slice(33, 97)).cpu()
    assert torch.equal(part, d[33:97])
  # the Python key API draws the same numbers (random.normal / ot_source / uniform)
  k = random.PRNGKey(5)
  assert torch.equal(random.normal(k, (777, 2)), ops.philox_rows(k.value, 0, _lib.ROWS_NORMAL, 777, 2, "cuda"))
  assert torch.equal(applications.sample_source_fn(k, 777, 2, "cuda") - random.normal(k, (777, 2)),
                     (applications.sample_source_fn(k, 777, 2, "cuda") - applications.sample_target_fn(k, 777, 2, "cuda")))
  assert random.uniform(k, (3, )).tolist() == ops.philox_times(k.value, 0, 3, 1.0)


def _explicit_inputs(cfg, shape, key, step, B, rows_B=None, rows_b=None):
  b, D = B // 32, shape.dim
  typ = cfg["general"]["type"]
  horizon = 1.0 if typ == "ot" else float(cfg[typ]["T"])
  inp = {"t_batch": ops.philox_times(key, step, cfg["general"]["t_batch_size"], horizon),
         "latent_sub": ops.philox_rows(key, step, _lib.ROWS_NORMAL, b, D, "cuda", rows=rows_b)}
  if typ == "ot":
    inp["src"] = ops.philox_rows(key, step, _lib.ROWS_OT_SOURCE, B, D, "cuda", rows=rows_B)
    inp["tgt"] = ops.philox_rows(key, step, _lib.ROWS_NORMAL, B, D, "cuda", rows=rows_B)
  else:
    inp["latent"] = ops.philox_rows(key, step, _lib.ROWS_NORMAL, B, D, "cuda", rows=rows_B)
  return inp


CASES = [("ot", "obstacle", {}), ("rwpo", "double_well", {}), ("fp", "nongradient", {}),
         ("fp", "nongradient", dict(dim=10, sigma=0.05)), ("ot", "free", dict(dim=3, H=32, sigma=0.1))]


@pytest.mark.parametrize("typ,sub,kw", CASES)
def test_step_with_on_chip_draws_equals_step_on_exported_arrays(typ, sub, kw):
  kw = dict(kw)
  sigma = kw.pop("sigma", 0.3)
  B = 2048 + 64
  cfg = make_cfg(typ, sub, Tn=2, lam=500.0, B=B, **kw)
  shape = shape_of(cfg)
  _, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  pd = ops.problem_desc(cfg)
  key, step = 0xABCDEF0123456789, 17
  inp = _explicit_inputs(cfg, shape, key, step, B)
  a = ops.mfc_step(shape, pd, W, inp.get("latent"), inp["latent_sub"], inp.get("src"), inp.get("tgt"), inp["t_batch"],
                   500.0, B, B // 32).clone()
  r = ops.mfc_step_rng(shape, pd, W, key, step, 2, 500.0, B, B // 32).clone()
  # same kernel, same numbers; only the order in which the CTAs' sums are combined differs
  assert float((a - r).abs().max() / a.abs().max()) < 2e-6
  # shards of the draw add up to the whole batch (global row index, not shard-local)
  b = B // 32
  parts = [ops.mfc_step_rng(shape, pd, W, key, step, 2, 500.0, B, b, rows_B=slice(i * B // 3, (i + 1) * B // 3),
                            rows_b=slice(i * b // 3, (i + 1) * b // 3)).clone() for i in range(3)]
  assert float((sum(parts) - a).abs().max() / a.abs().max()) < 2e-6
  # host entry: weights in, [gradient | loss] out, the key is the only other input
  out = torch.empty(shape.blob_size + 8, dtype=torch.float32).pin_memory()
  ops.mfc_step_rng_host(shape, pd, W.cpu().pin_memory(), key, step, 2, 500.0, B, b, out)
  assert float((out.cuda() - a).abs().max() / a.abs().max()) < 2e-6


def _oracle_adam_trajectory(cfg, shape, spec, params0, key, steps, lam, lr, B):
  """oracle value_and_grad (float64) + optax-style Adam in float64 on the exported draws of (key, step k)."""
  blob = pack(shape, params0, torch.float64)
  like = params0
  m, v = torch.zeros_like(blob), torch.zeros_like(blob)
  losses = []
  for k in range(steps):
    inp = _explicit_inputs(cfg, shape, key, k, B)
    o_in = {kk: (vv.double().cpu() if torch.is_tensor(vv) else torch.tensor(vv, dtype=torch.float64)) for kk, vv in inp.items()}
    if "latent" not in o_in:
      o_in["latent"] = o_in["latent_sub"]
    p = unpack(shape, blob, like)
    p = {mod: {n: t.double() for n, t in lv.items()} for mod, lv in p.items()}
    loss, grads = olosses.value_and_grad(cfg, spec, p, o_in, lam)
    g = pack(shape, grads, torch.float64)
    losses.append(float(loss))
    m = 0.9 * m + 0.1 * g
    v = 0.999 * v + 0.001 * g * g
    mh, vh = m / (1 - 0.9**(k + 1)), v / (1 - 0.999**(k + 1))
    blob = blob - lr * mh / (vh.sqrt() + 1e-8)
  return torch.tensor(losses, dtype=torch.float64), blob


@pytest.mark.parametrize("typ,sub", [("ot", "obstacle"), ("rwpo", "double_well"), ("fp", "nongradient")])
def test_fused_updates_follow_the_oracle_adam_trajectory(typ, sub):
  steps, B, lam, lr = 50, 1024, 50.0, 2e-3
  cfg = make_cfg(typ, sub, Tn=1, lam=lam, B=B)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, 0.3)
  W = pack(shape, params).cuda()
  pd = ops.problem_desc(cfg)
  key = 0x5EED5EED5EED
  state = ops.TrainState(shape, W, key)
  hist = torch.zeros(steps, device="cuda")
  out = torch.empty(shape.blob_size + 8, device="cuda")
  for k in range(steps):
    ops.mfc_update(shape, pd, state, W, 1, lam, B, B // 32, lr, out=out, loss_hist=hist)
  assert state.step_count() == steps and state.status() == 0
  l_or, w_or = _oracle_adam_trajectory(cfg, shape, spec, params, key, steps, lam, lr, B)
  rel = (hist.cpu().double() - l_or).abs() / l_or.abs()
  dw = (W.cpu().double() - w_or).abs()
  moved = float((w_or - pack(shape, params, torch.float64)).abs().max())
  print(f"loss curve rel err: first 5 steps {float(rel[:5].max()):.2e}, all {steps} steps {float(rel.max()):.2e}; weights moved "
        f"{moved:.3e}, |dw| median {float(dw.median()):.2e} q99 {float(dw.quantile(0.99)):.2e} max {float(dw.max()):.2e}")
  # float32 kernels vs float64 oracle.  The first steps are pure per-step parity; over 50 steps the two trajectories
  # separate slowly: Adam normalises every entry by its own magnitude (update = lr * m / sqrt(v)), so an entry whose
  # gradient sits at the float32 noise floor can move by a fraction of lr per step in either direction
  assert float(rel[:5].max()) < 5e-5, rel[:5]
  assert float(rel.max()) < 5e-2, float(rel.max())   # fp: the score finite differences make the noisiest gradients
  assert float(dw.median()) < 0.01 * moved and float(dw.quantile(0.99)) < 0.2 * moved, (float(dw.quantile(0.99)), moved)
  assert abs(float(out[shape.blob_size]) - float(hist[-1])) == 0.0


def test_graph_replay_and_main_train_device_resident():
  steps, B, lam, lr = 24, 2048, 50.0, 5e-3
  cfg = make_cfg("ot", "free", lam=lam, B=B)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  pd = ops.problem_desc(cfg)

  def run(graphed):
    W = pack(shape, params).cuda()
    state = ops.TrainState(shape, W, 77)
    hist = torch.zeros(steps, device="cuda")
    one = lambda: ops.mfc_update(shape, pd, state, W, 1, lam, B, B // 32, lr, loss_hist=hist)
    if not graphed:
      for _ in range(steps):
        one()
    else:
      one()   # warm-up outside the capture
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        for _ in range(4):
          one()
      for _ in range((steps - 1) // 4):
        g.replay()
      for _ in range((steps - 1) % 4):
        one()
    torch.cuda.synchronize()
    assert state.step_count() == steps
    return hist.cpu(), W.cpu()

  h0, w0 = run(False)
  h1, w1 = run(True)
  # same kernels, same draws; the order of the float atomics differs from run to run, training amplifies it slowly
  assert float(((h0 - h1).abs() / h0.abs())[:8].max()) < 1e-5
  assert float(((h0 - h1).abs() / h0.abs()).max()) < 1e-2 and float((w0 - w1).abs().max()) < 2e-2
  assert float(h0[-4:].mean()) < float(h0[:4].mean())
  # solvers.main: device-resident loop (eager head + replayed graphs), loss history from the kernel
  cfg["train"].update(epochs=130, lr=5e-3, eval_frequency=50)
  p, hist = solvers.main(cfg)
  h = torch.stack(hist).cpu()
  assert h.shape == (130, ) and bool(torch.isfinite(h).all()) and float(h[-10:].mean()) < 0.7 * float(h[:5].mean())
  # graphs of 50 updates and graphs of 1 update walk the same trajectory (same key, same step counter)
  p2, hist2 = solvers.main(cfg, graph_steps=1)
  h2 = torch.stack(hist2).cpu()
  assert float(((h - h2).abs() / h.abs())[:8].max()) < 1e-5 and abs(float(h[-10:].mean()) / float(h2[-10:].mean()) - 1) < 0.1
