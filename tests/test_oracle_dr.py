"""CPU checks of the oracle's restatement of the dimension-reduction loss (cnf_ot/dr/trainers.py:41-111) and of
the unconditional-flow layout (cond_shape=(0,)): closed form at the identity flow, autograd vs central
differences, parameter count, blob round trip."""
import torch

from cnf_ot_b200.layout import FlowShape, pack, unpack
from oracle import dr as odr
from oracle import flow as oflow


def _spec(dim=4, L=2, H=16):
  return oflow.FlowSpec(dim, L, [H, H], 5, conditional=False)


def test_identity_flow_closed_form():
  """Reference init = identity flows => x' = (x_1..x_sub, 0..0) and loss = mean sum_{c >= sub} x_c^2."""
  spec = _spec()
  enc, dec = oflow.init_params(spec, 1), oflow.init_params(spec, 2)
  x = torch.randn(500, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
  want = (x[:, 2:]**2).sum(-1).mean()
  got = odr.reconstruction_loss("enc_dec", spec, {"encoder": enc, "decoder": dec}, x, 2)
  assert abs(float(got) - float(want)) < 1e-12
  got = odr.reconstruction_loss("dec_only", spec, dec, x, 2)
  assert abs(float(got) - float(want)) < 1e-12


def test_unconditional_input_layers_have_no_time_row():
  """autoregressive.py:94-98: the condition is concatenated only if the flow is conditional."""
  spec = _spec(dim=3)
  p = oflow.init_params(spec, 0)
  assert tuple(p["mlp_layer0_d1/~/linear_0"]["w"].shape) == (1, 16)
  assert tuple(p["mlp_layer1_d2/~/linear_0"]["w"].shape) == (2, 16)
  shape = FlowShape(3, 2, 2, 16, 5, conditional=False)
  assert shape.param_count() == spec.param_count() == FlowShape(3, 2, 2, 16, 5).param_count() - 2 * 2 * 16
  # same blob layout as the conditional flow; the t rows are zero and are not leaves
  pp = oflow.perturb_params(p, 0.1)
  blob = pack(shape, pp, torch.float64)
  assert blob.numel() == FlowShape(3, 2, 2, 16, 5).blob_size
  off = shape.mlp_offset(0, 1)
  assert float(blob[off:off + 16].abs().max()) == 0.0
  back = unpack(shape, blob, like=pp)
  for mod in pp:
    for k in pp[mod]:
      assert torch.equal(back[mod][k].double(), pp[mod][k].double())


def test_autograd_matches_central_differences():
  spec = _spec(dim=3, H=8)
  g = torch.Generator().manual_seed(3)
  x = torch.randn(40, 3, dtype=torch.float64, generator=g) * 1.5
  for model in ("enc_dec", "dec_only"):
    dec = oflow.perturb_params(oflow.init_params(spec, 2), 0.2, seed=5)
    params = {"encoder": oflow.perturb_params(oflow.init_params(spec, 1), 0.2, seed=4), "decoder": dec} \
      if model == "enc_dec" else dec
    for mod in (params["encoder"], params["decoder"]) if model == "enc_dec" else (params, ):
      mod["~"]["first"] = mod["~"]["first"].double()   # float64 everywhere for the finite differences
    loss, grads = odr.value_and_grad(model, spec, params, x, 1)
    tree = params["decoder"] if model == "enc_dec" else params
    gtree = grads["decoder"] if model == "enc_dec" else grads
    for mod, leaf, idx in (("linear_out_layer0_d1", "b", (3, )), ("mlp_layer1_d2/~/linear_0", "w", (1, 2)), ("~", "first", (0, 7))):
      v = tree[mod][leaf]
      old = v[idx].item()
      h = 1e-6
      v[idx] = old + h
      lp = odr.reconstruction_loss(model, spec, params, x, 1)
      v[idx] = old - h
      lm = odr.reconstruction_loss(model, spec, params, x, 1)
      v[idx] = old
      fd = float(lp - lm) / (2 * h)
      assert abs(fd - float(gtree[mod][leaf][idx])) <= 1e-6 * max(1.0, abs(fd)), (model, mod, leaf, fd)
