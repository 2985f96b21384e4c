"""GPU parity of the wide-conditioner engine (cnf_ot_b200/csrc/wide.cu: batched tcgen05 GEMMs over row
chunks; BASELINE config 5 = dim 32, 16 layers, hidden 512) against the oracle, through the same C-ABI
entry points as the fused kernels (cnfot_flow_*_ws, cnfot_flow_*_vjp, cnfot_mfc_step)."""
import pytest
import torch

from cnf_ot_b200 import _lib, ops
from cnf_ot_b200.layout import pack
from oracle import flow as oflow
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_wide_engine(monkeypatch):
  """Shapes both engines cover (hidden 16 / 32 with 5 bins) must run on the wide engine here."""
  monkeypatch.setenv("CNFOT_ENGINE", "wide")


# same float32 tolerances as the fused kernels (tests/test_gpu_flow.py, tests/test_gpu_step.py)
TOL, TOL_MAX = 2e-5, 2e-4
TOL_LOSS, TOL_GRAD = 2e-5, 5e-5

# (D, L, M, H, sigma)
# sigma keeps the flows well conditioned (|log-det| < 8): the raw spline parameters scale with sigma * sqrt(H)
SHAPES = [(4, 2, 2, 64, 0.05), (5, 3, 1, 128, 0.01), (3, 2, 3, 64, 0.1), (17, 2, 2, 64, 0.03), (3, 2, 2, 512, 0.01),
          (3, 2, 2, 32, 0.1), (3, 2, 4, 48, 0.1), (40, 2, 2, 64, 0.02)]   # narrow widths, 4 hidden layers, dim > 32


def close(a, b, tol=TOL, tol_max=TOL_MAX):
  a = a.detach().cpu().double().reshape(-1)
  b = b.detach().cpu().double().reshape(-1)
  e = (a - b).abs() / (b.abs() + 1.0)
  return float(e.quantile(0.999)) < tol and float(e.max()) < tol_max


@pytest.mark.parametrize("D,L,M,H,sigma", SHAPES)
def test_forward_inverse_logprob(D, L, M, H, sigma):
  cfg = make_cfg(dim=D, L=L, M=M, H=H)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(5)
  n = 700 + 3
  x = torch.randn(n, D, generator=g, dtype=torch.float64).float()
  for per_row in (False, True):
    cond = torch.rand(n if per_row else 1, generator=g, dtype=torch.float64).float()
    c_or = cond.double().reshape(-1, 1) if per_row else cond.double()
    y_or, fld = oflow.flow_forward_and_log_det(spec, params, x.double(), c_or)
    x_or, ild = oflow.flow_inverse_and_log_det(spec, params, x.double(), c_or)
    assert float(fld.abs().max()) < 8 and float(ild.abs().max()) < 8  # conditioning of the case
    y, ld = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=False)
    assert _lib.last_launch_info()["engine"] == "wide"
    assert close(y, y_or) and close(ld, fld)
    xi, ldi = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=True)
    assert close(xi, x_or) and close(ldi, ild)
    _, lp = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=True, add_base=True)
    assert close(lp, oflow.base_log_prob(x_or) + ild)
    _, slp = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=False, add_base=True)
    assert close(slp, oflow.base_log_prob(x.double()) - fld)
    back, _ = ops.flow_eval(shape, W, y, cond.cuda(), inverse=True)
    assert close(back, x, 5e-5, 1e-3)


def test_identity_at_reference_init():
  """flows.py:71-76: zero-initialised output layers and `first` => identity flow."""
  cfg = make_cfg(dim=6, L=3, H=64)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.0)
  W = pack(shape, params).cuda()
  x = torch.randn(1000, 6, device="cuda")
  y, ld = ops.flow_eval(shape, W, x, torch.rand(1000, device="cuda"), inverse=False)
  assert float((y - x).abs().max()) < 4e-6 and float(ld.abs().max()) < 4e-6


@pytest.mark.parametrize("D,L,M,H,sigma", SHAPES[:3])
@pytest.mark.parametrize("inverse", [False, True])
def test_vjp(D, L, M, H, sigma, inverse):
  cfg = make_cfg(dim=D, L=L, M=M, H=H)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(11)
  n = 900 + 5
  x = torch.randn(n, D, generator=g, dtype=torch.float64).float()
  gout0 = torch.randn(n, D, generator=g).float()
  gld0 = torch.randn(n, generator=g).float()
  cond = torch.rand(n, generator=g, dtype=torch.float64).float()

  def oracle(gout, gld, add_base):
    xx = x.double().requires_grad_(True)
    p = oflow.clone_params(params, True)
    fn = oflow.flow_inverse_and_log_det if inverse else oflow.flow_forward_and_log_det
    o, l = fn(spec, p, xx, cond.double().reshape(-1, 1))
    if add_base:
      l = oflow.base_log_prob(o) + l if inverse else oflow.base_log_prob(xx) - l
    ((o * gout.double()).sum() + (l * gld.double()).sum()).backward()
    G = pack(shape, {m: {k: v.grad for k, v in lv.items()} for m, lv in p.items()}, torch.float64)
    return xx.grad, G

  for add_base in (False, True):
    gin_or, _ = oracle(gout0, gld0, add_base)
    gin, _ = ops.flow_vjp(shape, W, x.cuda(), cond.cuda(), gout0.cuda(), gld0.cuda(), inverse=inverse,
                          add_base=add_base)
    err = ((gin.cpu().double() - gin_or).abs() / (gin_or.abs() + 1)).max(-1).values
    ties = err > 1e-3   # rows on a knot tie: the adjoint jumps there
    assert int(ties.sum()) <= 2, int(ties.sum())
    assert float(err[~ties].quantile(0.999)) < 1e-4
    keep = (~ties).float()
    gout, gld = gout0 * keep[:, None], gld0 * keep
    _, Gor = oracle(gout, gld, add_base)
    _, G = ops.flow_vjp(shape, W, x.cuda(), cond.cuda(), gout.cuda(), gld.cuda(), inverse=inverse,
                        add_base=add_base)
    assert float((G.cpu().double() - Gor).abs().max() / Gor.abs().max()) < 2e-5


def run_step(cfg, shape, params, inputs, lam, rows=None, sub_rows=None):
  B = cfg["train"]["batch_size"]
  b = B // 32
  W = pack(shape, params).cuda()
  f = lambda t: t.float().cuda()
  rs = slice(0, B) if rows is None else rows
  ss = slice(0, b) if sub_rows is None else sub_rows
  ot = cfg["general"]["type"] == "ot"
  out = ops.mfc_step(shape, ops.problem_desc(cfg), W, None if ot else f(inputs["latent"][rs]), f(inputs["latent"][:b][ss]),
                     f(inputs["src"][rs]) if ot else None, f(inputs["tgt"][rs]) if ot else None,
                     inputs["t_batch"].tolist(), lam, B, b)
  return out.cpu().double()


STEP_CASES = [("ot", "free", dict(dim=4, H=64, sigma=0.05)), ("ot", "obstacle", dict(dim=2, H=64, sigma=0.1)),
              ("ot", "free", dict(dim=3, H=128, M=1, sigma=0.02)), ("ot", "obstacle", dict(dim=5, H=64, M=3, L=3, sigma=0.05)),
              ("ot", "free", dict(dim=3, H=512, B=384, sigma=0.01)),
              ("rwpo", "quadratic", dict(dim=2, H=64, sigma=0.1)), ("rwpo", "double_well", dict(dim=3, H=64, M=1, sigma=0.03)),
              ("fp", "gradient", dict(dim=2, H=64, sigma=0.1)), ("fp", "nongradient", dict(dim=4, H=64, sigma=0.05)),
              ("fp", "lorenz", dict(dim=3, H=128, L=3, sigma=0.02)),
              ("ot", "obstacle", dict(dim=3, H=48, M=4, sigma=0.1)), ("rwpo", "quadratic", dict(dim=2, H=32, sigma=0.1)),
              # two layers of BASELINE configs[4] itself: dim 32, 2 x 512 (17.4 M parameters)
              ("ot", "free", dict(dim=32, H=512, B=256, sigma=0.005))]


@pytest.mark.parametrize("typ,sub,kw", STEP_CASES)
def test_loss_and_gradient(typ, sub, kw):
  kw = dict(kw)
  sigma = kw.pop("sigma")
  cfg = make_cfg(typ, sub, Tn=2, lam=500.0, **({"B": 640 + 64} | kw))
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  inputs = make_inputs(cfg)
  loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
  Gor = pack(shape, grads, torch.float64)
  out = run_step(cfg, shape, params, inputs, 500.0)
  assert _lib.last_launch_info()["engine"] == "wide"
  G, slots = out[:shape.blob_size], out[shape.blob_size:]
  assert abs(float(slots[0]) - float(loss)) <= TOL_LOSS * abs(float(loss)), (float(slots[0]), float(loss))
  assert abs(float(slots[1:5].sum()) - float(slots[0])) <= 1e-5 * abs(float(slots[0]))
  # Tie rows: a sample whose hidden pre-activation (ReLU kink) or spline input (knot) sits within float32 rounding
  # of the breakpoint takes the other branch than the float64 oracle, and that one row's contribution to the
  # affected leaves flips (north_star excludes knot ties).  With 64-512 wide layers a case holds ~1e7
  # pre-activations, so a few such rows are expected: 99.9 % of the gradient entries are held to TOL_GRAD, every
  # entry to 3e-4 of the largest one.
  err = (G - Gor).abs() / Gor.abs().max()
  q999 = float(err.kthvalue(max(1, int(0.999 * err.numel()))).values)   # (torch.quantile stops at 16 M entries)
  assert q999 <= TOL_GRAD, q999
  assert float(err.max()) <= 3e-4, float(err.max())


def test_chunks_and_shards_sum_to_whole_batch(monkeypatch):
  """Row chunks inside a call (CNFOT_WIDE_CHUNK) and row shards across calls (the data-parallel contract,
  SURVEY 8e) both add up to the whole-batch result."""
  cfg = make_cfg("ot", "obstacle", dim=4, H=64, B=2048, lam=100.0)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.05)
  inputs = make_inputs(cfg)
  whole = run_step(cfg, shape, params, inputs, 100.0)
  monkeypatch.setenv("CNFOT_WIDE_CHUNK", "384")   # 6 chunks of the fit rows, ragged tail
  chunked = run_step(cfg, shape, params, inputs, 100.0)
  assert float((chunked - whole).abs().max() / whole.abs().max()) < 2e-6
  monkeypatch.delenv("CNFOT_WIDE_CHUNK")
  B, b = 2048, 64
  parts = [run_step(cfg, shape, params, inputs, 100.0, rows=slice(i * B // 4, (i + 1) * B // 4),
                    sub_rows=slice(i * b // 4, (i + 1) * b // 4)) for i in range(4)]
  assert float((sum(parts) - whole).abs().max() / whole.abs().max()) < 2e-6


def test_unsupported_requests_fail_loudly():
  from cnf_ot_b200._lib import CnfotError
  from cnf_ot_b200.layout import FlowShape
  cfg = make_cfg("ot", "free", dim=2, H=64, B=256)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.1)
  W = pack(shape, params).cuda()
  lib = _lib.load()
  x = torch.zeros(4, 2, device="cuda")
  c = torch.zeros(1, device="cuda")
  with pytest.raises(CnfotError):   # the workspace-less entry cannot run a wide flow
    _lib.check(lib.cnfot_flow_forward(0, _lib.flow_desc(shape), W.data_ptr(), x.data_ptr(), c.data_ptr(), 0, 4,
                                      x.data_ptr(), 0, 0))
  bad = FlowShape(2, 2, 2, 64, 8)   # neither engine is instantiated for 8 bins at hidden 64
  with pytest.raises(CnfotError):
    ops.flow_eval(bad, torch.zeros(bad.blob_size, device="cuda"), x, c, inverse=False)


@pytest.mark.parametrize("with_score", [False, True])
def test_kinetic_energy_matches_oracle(with_score):
  """cnfot_kinetic_energy on the wide engine (utils.calc_kinetic_energy / calc_score_kinetic_energy,
  cnf_ot/utils.py:311-389): same tolerance as the fused kernels (tests/test_gpu_energies.py)."""
  from oracle import energies as oen
  D = 3
  cfg = make_cfg(dim=D, H=64)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, 0.05)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(21)
  n_t, batch = 4, 300
  latent = torch.randn(n_t, batch, D, generator=g, dtype=torch.float64).float()
  ts = torch.linspace(0.0, 1.5, n_t, dtype=torch.float64).float().tolist()
  if with_score:
    ref = oen.score_kinetic_energy(spec, params, latent.double(), ts, beta=2.0)
  else:
    ref = oen.kinetic_energy(spec, params, latent.double(), ts)
  got = ops.kinetic_energy(shape, W, latent.reshape(-1, D).cuda(), ts, with_score=with_score, kappa=0.5, latent_blocks=n_t)
  assert _lib.last_launch_info()["engine"] == "wide"
  assert abs(float(got) - float(ref)) <= 2e-4 * abs(float(ref)), (float(got), float(ref))


def test_reference_api_on_a_wide_flow():
  """The RQSFlow / applications mirror (flows.py:213-226, applications.py) runs unchanged on a flow the fused
  kernels do not cover: model.apply.* and value_and_grad(loss_fn) reach the wide engine through the same C ABI."""
  import functools
  from cnf_ot_b200 import applications, random
  from cnf_ot_b200.flows import RQSFlow
  model = RQSFlow((3, ), 2, [64, 64], 5)
  params = model.init(random.PRNGKey(0), torch.zeros(1, 3), torch.zeros(1))
  x = torch.randn(500, 3, device="cuda")
  lp = model.apply.log_prob(params, x, cond=torch.tensor([0.3]))
  assert _lib.last_launch_info()["engine"] == "wide"
  ref = oflow.base_log_prob(x.double().cpu())   # identity flow at the reference initialisation
  assert float((lp.double().cpu() - ref).abs().max()) < 1e-4
  y = model.apply.forward(params, x, torch.tensor([0.3]))
  assert float((y - x).abs().max()) < 4e-6
  loss_fn = functools.partial(applications.ot_loss_fn, model, 3, 1, 0.01, 1, "free")
  loss, grads = applications.value_and_grad(loss_fn)(params, random.PRNGKey(1), 50.0, 1024)
  assert torch.isfinite(torch.as_tensor(float(loss)))
  assert float(grads.blob.abs().max()) > 0


def test_host_entry_on_a_wide_flow():
  """cnfot_mfc_step_host (host buffers in and out) on the wide engine: pinned rows are read in place, pageable ones
  are staged; both agree with the device entry."""
  cfg = make_cfg("ot", "free", dim=3, H=64, B=1024, lam=100.0)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.05)
  inputs = make_inputs(cfg)
  dev = run_step(cfg, shape, params, inputs, 100.0)
  b = 1024 // 32
  W = pack(shape, params)
  for pinned in (True, False):
    mk = (lambda t: t.float().contiguous().pin_memory()) if pinned else (lambda t: t.float().contiguous().clone())
    out = torch.empty(shape.blob_size + 8, dtype=torch.float32)
    out = out.pin_memory() if pinned else out
    ops.mfc_step_host(shape, ops.problem_desc(cfg), mk(W), None, mk(inputs["latent"][:b]), mk(inputs["src"]),
                      mk(inputs["tgt"]), inputs["t_batch"].tolist(), 100.0, 1024, b, out)
    assert float((out.double() - dev).abs().max() / dev.abs().max()) < 2e-6


@pytest.mark.parametrize("typ,sub,kw", [("ot", "obstacle", dict(dim=3, H=64)), ("rwpo", "double_well", dict(dim=3, H=64)),
                                        ("fp", "nongradient", dict(dim=4, H=64))])
def test_stacked_passes_match_one_pass_per_chunk(typ, sub, kw, monkeypatch):
  """Small batches stack the passes that share rows into one chunk (fewer launches); CNFOT_WIDE_BATCH=0 runs one
  pass per chunk (what large batches do).  Both orders of the same arithmetic agree to float32 rounding."""
  cfg = make_cfg(typ, sub, Tn=2, lam=100.0, B=1024, **kw)
  shape = shape_of(cfg)
  _, params = make_params(cfg, 0.05)
  inputs = make_inputs(cfg)
  stacked = run_step(cfg, shape, params, inputs, 100.0)
  monkeypatch.setenv("CNFOT_WIDE_BATCH", "0")
  plain = run_step(cfg, shape, params, inputs, 100.0)
  assert float((stacked - plain).abs().max() / plain.abs().max()) < 5e-6
