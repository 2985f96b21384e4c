import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cnf_ot_b200 import dr
from cnf_ot_b200.flows import ParamTree
from cnf_ot_b200.layout import pack
from oracle import dr as odr, flow as oflow
def _params(spec, shape, seed, sigma):
  p = oflow.perturb_params(oflow.init_params(spec, seed), sigma, seed=seed + 10)
  for mod in p:
    for k in p[mod]: p[mod][k] = p[mod][k].to(torch.float32).to(p[mod][k].dtype)
  return p, ParamTree(shape, pack(shape, p).cuda())
for H, sigma in [(16, 0.08), (16, 0.2)]:
  cfg = {"cnf": {"flow_num_layers": 2, "mlp_num_layers": 2, "hidden_size": H, "num_bins": 5}}
  enc, dec = dr.build(4, cfg, "enc_dec"); shape = dec.shape
  spec = oflow.FlowSpec(4, 2, [H, H], 5, conditional=False)
  g = torch.Generator().manual_seed(7)
  x = (torch.randn(1007, 4, generator=g, dtype=torch.float64) * 1.5).float()
  pd, td = _params(spec, shape, 2, sigma); pe, te = _params(spec, shape, 1, sigma)
  loss_or, g_or = odr.value_and_grad("enc_dec", spec, {"encoder": pe, "decoder": pd}, x.double(), 2)
  loss, grads = dr.value_and_grad("enc_dec", enc, dec, 2)({"encoder": te, "decoder": td}, x.cuda())
  print(H, sigma, float(loss), float(loss_or))
  for name in ("encoder", "decoder"):
    G, Gor = grads[name].blob.cpu().double(), pack(shape, g_or[name], torch.float64)
    sc = float(Gor.abs().max()); worst = []
    for mod, leaf, shp, off, stride in shape.leaves():
      rows = 1
      for s_ in shp[:-1]: rows *= s_
      idx = torch.cat([torch.arange(off + r * stride, off + r * stride + shp[-1]) for r in range(rows)])
      worst.append((float((G[idx] - Gor[idx]).abs().max()) / sc, mod, leaf, float(Gor[idx].abs().max()) / sc))
    worst.sort(reverse=True)
    print(" ", name, "scale %.3g" % sc, [(("%.1e" % e), m, l, "%.1e" % mg) for e, m, l, mg in worst[:4]])
  # per-row check of the pieces
  y_or, _ = oflow.flow_forward_and_log_det(spec, pe, x.double())
  from cnf_ot_b200 import ops
  y, _ = ops.flow_eval(shape, te.blob, x.cuda(), (0.0,), inverse=False, want_logdet=False)
  e = (y.cpu().double() - y_or).abs().max(-1).values
  print("  encoder forward max err %.2e, rows with err > 1e-4: %d" % (float(e.max()), int((e > 1e-4).sum())))
