"""Wide engine vs fused kernels vs oracle on shapes both engines cover (CNFOT_ENGINE=wide forces the wide one)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cnf_ot_b200 import ops, _lib
from cnf_ot_b200.layout import pack
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of

def run(cfg, shape, params, inputs, lam):
  B = cfg["train"]["batch_size"]; b = B // 32
  W = pack(shape, params).cuda(); f = lambda t: t.float().cuda()
  out = ops.mfc_step(shape, ops.problem_desc(cfg), W, None, f(inputs["latent"][:b]), f(inputs["src"]), f(inputs["tgt"]),
                     inputs["t_batch"].tolist(), lam, B, b)
  return out.cpu().double(), _lib.last_launch_info()["engine"]

for sub, kw, sigma in [("obstacle", dict(dim=5, H=16, M=3, L=3), 0.1), ("free", dict(dim=3, H=16, B=384), 0.1),
                       ("obstacle", dict(dim=5, H=16, M=3, L=3), 0.2)]:
  for lam in (500.0, 0.0):
    cfg = make_cfg("ot", sub, Tn=2, lam=lam, **({"B": 704} | kw))
    shape = shape_of(cfg)
    spec, params = make_params(cfg, sigma)
    inputs = make_inputs(cfg)
    loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
    Gor = pack(shape, grads, torch.float64)
    os.environ.pop("CNFOT_ENGINE", None)
    fused, e1 = run(cfg, shape, params, inputs, lam)
    os.environ["CNFOT_ENGINE"] = "wide"
    wide, e2 = run(cfg, shape, params, inputs, lam)
    n = shape.blob_size
    sc = float(Gor.abs().max())
    print(sub, kw, "lam", lam, e1, e2, "loss or %.8g fused %.8g wide %.8g" % (float(loss), float(fused[n]), float(wide[n])))
    for name, g in (("fused", fused), ("wide", wide)):
      err = (g[:n] - Gor).abs() / sc
      i = int(err.argmax())
      print("  %-5s grad err max %.2e at %d (first-block err %.2e, rest %.2e)  slots %s" % (
        name, float(err.max()), i, float(err[:16].max()), float(err[16:].max()), [round(float(v), 6) for v in g[n:n + 5]]))
    print("  wide vs fused %.2e" % (float((wide[:n] - fused[:n]).abs().max()) / sc))
