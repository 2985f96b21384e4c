"""GPU: the host-side mirror of the reference API (RQSFlow / applications / solvers)."""
from functools import partial

import pytest
import torch

from cnf_ot_b200 import applications, random, solvers
from cnf_ot_b200.flows import ParamTree, RQSFlow
from cnf_ot_b200.layout import pack
from oracle import flow as oflow
from oracle import losses as olosses
from util import make_cfg, make_params, rel_err, shape_of

pytestmark = pytest.mark.gpu


def _model_and_params(cfg, sigma=0.3):
  c = cfg["cnf"]
  model = RQSFlow((cfg["general"]["dim"], ), c["flow_num_layers"], [c["hidden_size"]] * c["mlp_num_layers"],
                  c["num_bins"])
  spec, ref = make_params(cfg, sigma)
  params = ParamTree(model.shape, pack(model.shape, ref).cuda())
  return model, params, spec, ref


def test_model_api_matches_reference_semantics():
  cfg = make_cfg()
  model, params, spec, ref = _model_and_params(cfg)
  # the haiku-shaped tree aliases the blob
  assert set(params) == set(ref) and params["~"]["first"].shape == (1, 16)
  assert params["mlp_layer1_d1/~/linear_1"]["w"].shape == (16, 16)
  n = 512
  latent = random.normal(random.PRNGKey(1), (n, 2))
  cond = torch.full((n, 1), 0.4, device="cuda")
  y = model.apply.sample(params, cond=cond, seed=random.PRNGKey(1), sample_shape=(n, ))
  y2 = model.apply.sample(params, cond=cond, seed=random.PRNGKey(1), sample_shape=(n, ))
  assert torch.equal(y, y2)  # equal key => equal latent (the FD velocity relies on it)
  y_or = oflow.sample(spec, ref, latent.double().cpu(), cond.double().cpu())
  assert rel_err(y, y_or) < 2e-5
  ys, lp = model.apply.sample_and_log_prob(params, cond=cond, seed=random.PRNGKey(1), sample_shape=(n, ))
  _, lp_or = oflow.sample_and_log_prob(spec, ref, latent.double().cpu(), cond.double().cpu())
  assert torch.equal(ys, y) and rel_err(lp, lp_or) < 5e-5
  lq = model.apply.log_prob(params, y, cond=torch.tensor([0.4]))
  assert rel_err(lq, oflow.log_prob(spec, ref, y.double().cpu(), torch.tensor([0.4], dtype=torch.float64))) < 5e-5
  x = model.apply.inverse(params, model.apply.forward(params, latent, torch.tensor([0.4])), torch.tensor([0.4]))
  assert rel_err(x, latent) < 1e-4
  with pytest.raises(NotImplementedError):
    model.apply.forward_jac(params, latent, cond)
  # a plain dict of reference-shaped leaves is accepted too (packed on the fly)
  y3 = model.apply.sample(ref, cond=cond, latent=latent)
  assert torch.equal(y3, y)


def test_init_is_identity_flow():
  model = RQSFlow((3, ), 2, [16, 16], 5)
  params = model.init(random.PRNGKey(0), torch.zeros(1, 3), torch.zeros(1))
  assert float(params["linear_out_layer0_d1"]["w"].abs().max()) == 0.0
  w = params["mlp_layer0_d2/~/linear_0"]["w"]
  assert w.shape == (3, 16) and float(w.abs().max()) <= 2.0 / 3**0.5 + 1e-6 and float(w.std()) > 0.2
  x = random.normal(random.PRNGKey(5), (1000, 3))
  y = model.apply.sample(params, cond=torch.zeros(1000, 1, device="cuda"), latent=x)
  assert float((y - x).abs().max()) < 4e-6


@pytest.mark.parametrize("typ,sub", [("ot", "obstacle"), ("rwpo", "double_well"), ("fp", "nongradient")])
def test_value_and_grad_equals_forward_losses_and_oracle(typ, sub):
  """The fused step (value_and_grad) and the term-by-term forward functions of
  applications.py agree, and both agree with the oracle on the same draws."""
  cfg = make_cfg(typ, sub, B=2048, Tn=2, lam=50.0)
  model, params, spec, ref = _model_and_params(cfg)
  _, loss_fn, T = solvers.build(cfg)
  loss_fn = partial(loss_fn.func, model, *loss_fn.args[1:])
  rng = random.PRNGKey(7)
  vg = applications.value_and_grad(loss_fn)
  loss, grads = vg(params, rng, 50.0, 2048)
  fwd = loss_fn(params, rng, 50.0, 2048)
  assert abs(float(loss) - float(fwd)) <= 2e-5 * abs(float(fwd))
  # oracle on the very same draws
  sc = vg.step_config
  inp = applications.draw_step_inputs(model, sc["cfg"], sc["horizon"], rng, 2048)
  f64 = lambda t: None if t is None else t.double().cpu()
  o_in = {"latent": f64(inp.get("latent")), "src": f64(inp.get("src")), "tgt": f64(inp.get("tgt")),
          "t_batch": torch.tensor(inp["t_batch"], dtype=torch.float64)}
  o_in["latent_sub"] = f64(inp["latent_sub"])
  l_or, g_or = olosses.value_and_grad(cfg, spec, ref, o_in, 50.0)
  assert abs(float(loss) - float(l_or)) <= 2e-5 * abs(float(l_or))
  Gor = pack(model.shape, g_or, torch.float64)
  # the score terms of rwpo / fp are finite differences of float32 log-densities (1 / dx = 100): stated tolerance 2e-4
  tol = 5e-5 if typ == "ot" else 2e-4
  assert float((grads.blob.cpu().double() - Gor).abs().max() / Gor.abs().max()) < tol


def test_training_reduces_loss():
  """solvers.main on a short run: the loss goes down and stays finite (reference-style
  training loop with value_and_grad + Adam)."""
  cfg = make_cfg("ot", "free", B=4096, lam=50.0)
  cfg["train"].update(epochs=150, lr=5e-3)
  params, hist = solvers.main(cfg)
  h = torch.stack([l.detach() for l in hist]).cpu()
  assert bool(torch.isfinite(h).all())
  assert float(h[-10:].mean()) < 0.7 * float(h[:5].mean())


def test_params_out_and_params_in(tmp_path):
  """train.params_out writes the haiku-shaped pytree as .npz after training; train.params_in resumes from it
  (the reference has no parameter I/O)."""
  path = str(tmp_path / "params.npz")
  cfg = make_cfg("ot", "free", B=1024, lam=50.0)
  cfg["train"].update(epochs=3, lr=5e-3, params_out=path)
  params, _ = solvers.main(cfg)
  back = ParamTree.load(path)
  for mod in params:
    for leaf in params[mod]:
      assert torch.equal(back[mod][leaf], params[mod][leaf].cpu()), (mod, leaf)
  cfg2 = make_cfg("ot", "free", B=1024, lam=50.0)
  cfg2["train"].update(epochs=0, params_in=path)
  resumed, hist = solvers.main(cfg2)
  assert hist == [] and torch.equal(resumed.blob.cpu(), back.blob)
