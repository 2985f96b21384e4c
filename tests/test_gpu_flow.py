"""GPU parity: cnfot_flow_* (C ABI) vs the oracle restatement of the reference's
conditional autoregressive flow (flows.py / autoregressive.py / conditional.py)."""
import pytest
import torch

from cnf_ot_b200 import ops
from cnf_ot_b200.layout import pack
from oracle import flow as oflow
from util import make_cfg, make_params, rel_err, shape_of

pytestmark = pytest.mark.gpu
TOL = 2e-5

SHAPES = [(2, 2, 2, 16, 5, 0.3), (3, 3, 1, 8, 3, 0.3), (10, 2, 2, 16, 5, 0.05),
          (4, 3, 2, 32, 8, 0.2), (2, 2, 3, 16, 5, 0.3), (2, 4, 1, 16, 5, 0.3)]


@pytest.mark.parametrize("D,L,M,H,K,sigma", SHAPES)
def test_forward_inverse_logprob(D, L, M, H, K, sigma):
  cfg = make_cfg(dim=D, L=L, M=M, H=H, K=K)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(5)
  n = 1000 + 3
  x = (torch.randn(n, D, generator=g, dtype=torch.float64) * 1.5).float()
  for per_row in (False, True):
    cond = torch.rand(n if per_row else 1, generator=g, dtype=torch.float64).float()
    c_or = cond.double().reshape(-1, 1) if per_row else cond.double()
    y_or, fld = oflow.flow_forward_and_log_det(spec, params, x.double(), c_or)
    x_or, ild = oflow.flow_inverse_and_log_det(spec, params, x.double(), c_or)
    y, ld = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=False)
    assert rel_err(y, y_or) < TOL and rel_err(ld, fld) < TOL
    xi, ldi = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=True)
    assert rel_err(xi, x_or) < TOL and rel_err(ldi, ild) < TOL
    # the densities ConditionalTransformed returns
    _, lp = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=True, add_base=True)
    assert rel_err(lp, oflow.base_log_prob(x_or) + ild) < TOL
    _, slp = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=False, add_base=True)
    assert rel_err(slp, oflow.base_log_prob(x.double()) - fld) < TOL
    # inverse(forward(x)) round trip on the device
    back, _ = ops.flow_eval(shape, W, y, cond.cuda(), inverse=True)
    assert rel_err(back, x) < 5e-5


def test_identity_at_reference_init():
  """flows.py:71-76: zero-initialised output layers and `first` => identity flow."""
  cfg = make_cfg(dim=3, L=3)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, 0.0)
  W = pack(shape, params).cuda()
  x = torch.randn(4096, 3, device="cuda")
  t = torch.rand(4096, device="cuda")
  y, ld = ops.flow_eval(shape, W, x, t, inverse=False)
  assert float((y - x).abs().max()) < 4e-6 and float(ld.abs().max()) < 4e-6
  _, lp = ops.flow_eval(shape, W, x, torch.tensor([0.5]), inverse=True, add_base=True)
  assert rel_err(lp, oflow.base_log_prob(x.double().cpu())) < 1e-5


@pytest.mark.parametrize("D,L,M,H,K,sigma", SHAPES[:4])
@pytest.mark.parametrize("inverse", [False, True])
def test_vjp(D, L, M, H, K, sigma, inverse):
  cfg = make_cfg(dim=D, L=L, M=M, H=H, K=K)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(11)
  n = 777
  x = (torch.randn(n, D, generator=g, dtype=torch.float64) * 1.2).float()
  gout = torch.randn(n, D, generator=g).float()
  gld = torch.randn(n, generator=g).float()
  cond = torch.rand(n, generator=g, dtype=torch.float64).float()
  for add_base in (False, True):
    xx = x.double().requires_grad_(True)
    p = oflow.clone_params(params, True)
    fn = oflow.flow_inverse_and_log_det if inverse else oflow.flow_forward_and_log_det
    o, l = fn(spec, p, xx, cond.double().reshape(-1, 1))
    if add_base:
      l = oflow.base_log_prob(o) + l if inverse else oflow.base_log_prob(xx) - l
    ((o * gout.double()).sum() + (l * gld.double()).sum()).backward()
    Gor = pack(shape, {m: {k: v.grad for k, v in lv.items()} for m, lv in p.items()}, torch.float64)
    gin, G = ops.flow_vjp(shape, W, x.cuda(), cond.cuda(), gout.cuda(), gld.cuda(), inverse=inverse,
                          add_base=add_base)
    assert rel_err(gin, xx.grad) < 20 * TOL
    assert float((G.cpu().double() - Gor).abs().max() / Gor.abs().max()) < 2e-5


def test_round_trip_at_full_batch():
  """inverse(forward(x)) over 2^18 rows of the benchmark shape (size-independent property)."""
  cfg = make_cfg()
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  W = pack(shape, params).cuda()
  x = torch.randn(1 << 18, 2, device="cuda")
  t = torch.rand(1 << 18, device="cuda")
  y, fld = ops.flow_eval(shape, W, x, t, inverse=False)
  back, ild = ops.flow_eval(shape, W, y, t, inverse=True)
  assert float(((back - x).abs() / (x.abs() + 1)).max()) < 5e-5
  assert float((fld + ild).abs().max()) < 5e-4


def test_empty_and_errors():
  from cnf_ot_b200._lib import CnfotError
  cfg = make_cfg()
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.3)
  W = pack(shape, params).cuda()
  y, ld = ops.flow_eval(shape, W, torch.empty(0, 2, device="cuda"), torch.tensor([0.1]), inverse=False)
  assert y.shape == (0, 2)
  with pytest.raises(CnfotError):
    ops.flow_eval(shape, W.cpu(), torch.zeros(4, 2), torch.tensor([0.1]), inverse=False)
  from cnf_ot_b200.layout import FlowShape
  bad = FlowShape(2, 2, 2, 24, 5)
  with pytest.raises(CnfotError):
    ops.flow_eval(bad, torch.zeros(bad.blob_size, device="cuda"), torch.zeros(4, 2, device="cuda"),
                  torch.tensor([0.1]), inverse=False)
