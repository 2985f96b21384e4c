"""Where is the wide engine's gradient error?  Per-leaf error of the failing parity cases."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import pack
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of

for sub, kw, sigma in [("obstacle", dict(dim=5, H=64, M=3, L=3), 0.05), ("free", dict(dim=3, H=512, B=384), 0.01)]:
  for lam in (500.0, 0.0):
    cfg = make_cfg("ot", sub, Tn=2, lam=lam, **({"B": 704} | kw))
    shape = shape_of(cfg)
    spec, params = make_params(cfg, sigma)
    inputs = make_inputs(cfg)
    loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
    Gor = pack(shape, grads, torch.float64)
    B = cfg["train"]["batch_size"]; b = B // 32
    f = lambda t: t.float().cuda()
    out = ops.mfc_step(shape, ops.problem_desc(cfg), pack(shape, params).cuda(), None, f(inputs["latent"][:b]),
                       f(inputs["src"]), f(inputs["tgt"]), inputs["t_batch"].tolist(), lam, B, b).cpu().double()
    n = shape.blob_size
    sc = float(Gor.abs().max())
    print(sub, kw, "lam", lam, "loss or %.9g ours %.9g  slots %s  grad scale %.4g" % (float(loss), float(out[n]), out[n:n+5].tolist(), sc))
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
