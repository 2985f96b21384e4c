"""The kernels' per-row math (cnf_ot_b200/csrc/*_math.cuh) compiled for the HOST by
tests/hostsim and checked against the autograd oracle, in float64 (derivation
must be exact) and float32 (what the GPU computes).  Catches math / adjoint bugs
without a GPU; the GPU parity tests (-m gpu) then check the kernels themselves."""
import pytest
import torch

import hostsim as hs
from cnf_ot_b200.layout import pack
from oracle import flow as oflow
from oracle import losses as olosses
from oracle import rqs
from util import make_cfg, make_inputs, make_params, rel_err, shape_of


@pytest.mark.parametrize("K", [3, 5, 8, 10])
@pytest.mark.parametrize("direction", [0, 1])
def test_spline_and_adjoints_f64(K, direction):
  g = torch.Generator().manual_seed(K * 10 + direction)
  n = 3000
  theta = torch.randn(n, 3 * K + 1, generator=g, dtype=torch.float64) * 1.5
  v = torch.randn(n, generator=g, dtype=torch.float64) * 5.0  # ~5% of rows in the tails
  gout = torch.randn(n, generator=g, dtype=torch.float64)
  gld = torch.randn(n, generator=g, dtype=torch.float64)
  vv, th = v.clone().requires_grad_(True), theta.clone().requires_grad_(True)
  f = rqs.rqs_forward if direction == 0 else rqs.rqs_inverse
  o, l, idx = f(vv, th)
  (o * gout + l * gld).sum().backward()
  out, ld, bins, gin, gth = hs.rqs(K, direction, v, theta, gout, gld)
  assert torch.equal(bins.long(), idx)
  assert (out - o.detach()).abs().max() < 1e-10
  assert (ld - l.detach()).abs().max() < 1e-10
  assert (gin - vv.grad).abs().max() < 1e-8
  assert (gth - th.grad).abs().max() < 1e-8


@pytest.mark.parametrize("D,L,M,H,K,sigma", [(2, 2, 2, 16, 5, 0.3), (3, 3, 1, 8, 3, 0.3),
                                             (10, 2, 2, 16, 5, 0.05), (4, 3, 2, 32, 8, 0.2)])
def test_flow_pass_and_vjp_f64(D, L, M, H, K, sigma):
  cfg = make_cfg(dim=D, L=L, M=M, H=H, K=K)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  params["~"]["first"] = params["~"]["first"].double()  # exactness check: no f32 leaf
  W = pack(shape, params, torch.float64)
  g = torch.Generator().manual_seed(5)
  x = torch.randn(96, D, generator=g, dtype=torch.float64)
  for direction in (0, 1):
    for per_row in (False, True):
      cond = torch.rand(96 if per_row else 1, generator=g, dtype=torch.float64)
      c_or = cond.reshape(-1, 1) if per_row else cond
      xx = x.clone().requires_grad_(True)
      p = oflow.clone_params(params, True)
      fn = oflow.flow_forward_and_log_det if direction == 0 else oflow.flow_inverse_and_log_det
      o, l = fn(spec, p, xx, c_or)
      gout = torch.randn(96, D, generator=g, dtype=torch.float64)
      gld = torch.randn(96, generator=g, dtype=torch.float64)
      ((o * gout).sum() + (l * gld).sum()).backward()
      out, ld = hs.flow_eval(shape, direction, W, x, cond)
      gin, G = hs.flow_vjp(shape, direction, W, x, cond, gout, gld)
      Gor = pack(shape, {m: {k: v.grad for k, v in lv.items()} for m, lv in p.items()}, torch.float64)
      assert (out - o.detach()).abs().max() < 1e-10
      assert (ld - l.detach()).abs().max() < 1e-9
      assert (gin - xx.grad).abs().max() < 1e-7
      assert rel_err(G, Gor) < 1e-9


CASES = [
  ("ot", "free", {}), ("ot", "obstacle", {}), ("rwpo", "quadratic", {}), ("rwpo", "double_well", {}),
  ("fp", "gradient", {}), ("fp", "nongradient", {}),
  ("fp", "lorenz", dict(dim=3, L=3, sigma=0.1)),
  ("fp", "nongradient", dict(dim=10, sigma=0.05, B=128)),
  ("rwpo", "double_well", dict(dim=4, H=32, K=8, sigma=0.1, B=128)),
]


def _problem(cfg):
  from cnf_ot_b200.ops import problem_desc
  pd = problem_desc(cfg)
  return hs.ProblemDesc(pd.type, pd.subtype, pd.T, pd.beta, pd.a, pd.sigma, pd.dt, pd.dx)


@pytest.mark.parametrize("typ,sub,kw", CASES)
@pytest.mark.parametrize("dtype,tol_loss,tol_grad", [(torch.float64, 1e-6, 1e-6), (torch.float32, 2e-5, 5e-5)])
def test_step_value_and_grad(typ, sub, kw, dtype, tol_loss, tol_grad):
  """Whole train step (loss + parameter gradient) of the kernels' row functions vs
  jax.value_and_grad's restatement.  float32 tolerance: the finite-difference terms
  amplify rounding by 1/dt = 1/dx = 100 (SURVEY.md §7 hard part 1)."""
  kw = dict(kw)
  sigma = kw.pop("sigma", 0.3)
  cfg = make_cfg(typ, sub, Tn=2, lam=500.0, **({"B": 256} | kw))
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  if dtype == torch.float64:
    # exactness check of the hand-written adjoints: lift the reference's float32
    # `first` leaf so the oracle has no float32 arithmetic in it
    params["~"]["first"] = params["~"]["first"].double()
  inputs = make_inputs(cfg)
  loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
  Gor = pack(shape, grads, torch.float64)
  B = cfg["train"]["batch_size"]
  b = B // 32
  W = pack(shape, params, torch.float64)
  G, slots = hs.step(shape, _problem(cfg), W, inputs["latent"], inputs["latent"][:b], inputs["src"],
                     inputs["tgt"], inputs["t_batch"], B, b, 500.0, dtype=dtype)
  tot = float(slots.sum())
  assert abs(tot - float(loss)) <= tol_loss * abs(float(loss)), (tot, float(loss))
  gerr = float((G - Gor).abs().max() / Gor.abs().max())
  assert gerr <= tol_grad, gerr


def test_device_log1p_polynomial_accuracy():
  """softplus on the device uses log(1 + e) = e + e^2 Q(e) on [0, 1] (csrc/rqs_math.cuh:m_log1p, float32 FFMA chain).
  Re-evaluate the same chain here with the coefficients read from the header, rounding every step to float32:
  maximum relative error below 2e-7 (log1pf is ~1e-7), which is what DESIGN.md section 4.1 states."""
  import os
  import re
  import numpy as np
  src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cnf_ot_b200", "csrc", "rqs_math.cuh")).read()
  body = src[src.index("CNFOT_HD float m_log1p(float e) {"):]
  body = body[:body.index("}")]
  first = float(re.search(r"float q = (-?[0-9.eE+-]+)f;", body).group(1))
  rest = [float(m) for m in re.findall(r"q = fmaf\(q, e, (-?[0-9.eE+-]+)f\);", body)]
  assert len(rest) == 7 and "return fmaf(q * e, e, e);" in body
  e = np.linspace(0.0, 1.0, 200001)[1:].astype(np.float32)
  fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, dtype=np.float64)).astype(np.float32)
  q = np.full_like(e, np.float32(first))
  for c in rest:
    q = fma(q, e, np.float32(c))
  r = fma((q * e).astype(np.float32), e, e)
  ref = np.log1p(e.astype(np.float64))
  rel = np.abs(r.astype(np.float64) - ref) / ref
  assert float(rel.max()) < 2e-7 and float(rel.mean()) < 6e-8


def test_compile_time_spline_constants_equal_the_runtime_ones():
  """The fused kernels hold the reference's spline constants (flows.py:124-132) as immediates
  (FixedSplineConsts, rqs_math.cuh): the literals equal what make_spline_consts computes."""
  import ctypes
  import numpy as np
  for K in (3, 5, 8):
    f, r = np.zeros(6), np.zeros(6)
    hs.lib("rqs").hs_spline_consts(K, f.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p))
    assert np.allclose(f, r, rtol=0, atol=1e-15), (K, f, r)
