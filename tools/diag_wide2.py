"""Where is the wide engine's gradient error?  Per-leaf error of the failing parity cases."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import pack
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of

CASES = [("rwpo", "double_well", dict(dim=3, H=64, M=1), 0.1), ("rwpo", "double_well", dict(dim=3, H=64, M=1), 0.03),
         ("rwpo", "double_well", dict(dim=3, H=64, M=2), 0.1), ("rwpo", "quadratic", dict(dim=3, H=64, M=1), 0.1)]
for typ, sub, kw, sigma in CASES:
  for lam in (500.0,):
    cfg = make_cfg(typ, sub, Tn=2, lam=lam, **({"B": 704} | kw))
    shape = shape_of(cfg)
    spec, params = make_params(cfg, sigma)
    inputs = make_inputs(cfg)
    loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
    Gor = pack(shape, grads, torch.float64)
    B = cfg["train"]["batch_size"]; b = B // 32
    f = lambda t: t.float().cuda()
    ot = typ == "ot"
    out = ops.mfc_step(shape, ops.problem_desc(cfg), pack(shape, params).cuda(), None if ot else f(inputs["latent"]),
                       f(inputs["latent"][:b]), f(inputs["src"]) if ot else None, f(inputs["tgt"]) if ot else None,
                       inputs["t_batch"].tolist(), lam, B, b).cpu().double()
    n = shape.blob_size
    sc = float(Gor.abs().max())
    print(typ, sub, kw, "sigma", sigma, "lam", lam, "loss or %.9g ours %.9g  slots %s  grad scale %.4g" % (float(loss), float(out[n]), out[n:n+5].tolist(), sc))
    worst = []
    for mod, leaf, shp, off, stride in shape.leaves():
      rows = 1
      for s_ in shp[:-1]: rows *= s_
      idx = torch.cat([torch.arange(off + r * stride, off + r * stride + shp[-1]) for r in range(rows)])
      e = float((out[idx] - Gor[idx]).abs().max()) / sc
      worst.append((e, mod, leaf, float(Gor[idx].abs().max()) / sc))
    worst.sort(reverse=True)
    for e, mod, leaf, mag in worst[:8]:
      print("   err %.2e  (leaf max %.2e of scale)  %s/%s" % (e, mag, mod, leaf))
