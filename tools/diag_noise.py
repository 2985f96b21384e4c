"""Error levels of the finite-difference (score) cases under two builds of the library (dev tool):
  CNFOT_LIB=<path to libcnfot.so> python tools/diag_noise.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from cnf_ot_b200 import _lib
if os.environ.get("CNFOT_LIB"):
  _lib.LIB_PATH = os.environ["CNFOT_LIB"]
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import FlowShape, pack
from oracle import energies as oen
from util import make_cfg, make_params, shape_of
import test_golden as tg, test_reference_golden as tr

def energies(D, sigma, seed):
  cfg = make_cfg(dim=D); shape = shape_of(cfg); spec, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(seed)
  n_t, batch = 5, 300
  latent = torch.randn(n_t, batch, D, generator=g, dtype=torch.float64).float()
  ts = torch.linspace(0.0, 1.5, n_t, dtype=torch.float64).float().tolist()
  ref = float(oen.score_kinetic_energy(spec, params, latent.double(), ts, beta=2.0))
  out = []
  for eng in ("mma", "cuda"):
    os.environ["CNFOT_ENGINE"] = eng
    got = float(ops.kinetic_energy(shape, W, latent.reshape(-1, D).cuda(), ts, with_score=True, kappa=0.5, latent_blocks=n_t))
    out.append("%s %.2e" % (eng, abs(got - ref) / abs(ref)))
  return out

def step(mod, name):
  g = mod.load(name)
  shape = (mod.spec_and_params(g) if mod is tg else mod.parts(g))[0]
  cfg = mod.step_cfg(name, g)
  B = int(g["latent"].shape[0]); b = B // 32; typ = cfg["general"]["type"]
  f = lambda t: t.float().cuda()
  out = []
  for eng in ("mma", "cuda"):
    os.environ["CNFOT_ENGINE"] = eng
    o = ops.mfc_step(shape, ops.problem_desc(cfg), f(g["blob"]), None if typ == "ot" else f(g["latent"]), f(g["latent"][:b]),
                     f(g["src"]) if typ == "ot" else None, f(g["tgt"]) if typ == "ot" else None, g["t_batch"].tolist(),
                     float(g["lam"]), B, b).cpu().double()
    G = o[:shape.blob_size]
    out.append("%s grad %.2e loss %.1e" % (eng, float((G - g["grad"]).abs().max() / g["grad"].abs().max()),
                                            abs(float(o[shape.blob_size]) - float(g["loss"])) / abs(float(g["loss"]))))
  return out

print("lib:", _lib.LIB_PATH)
for seed in (21, 22, 23):
  print("energy D=2 sigma=0.3 seed", seed, energies(2, 0.3, seed))
print("energy D=3 sigma=0.1", energies(3, 0.1, 21))
for name in ("step_fp_nongradient_d4", "step_rwpo_double_well", "step_ot_obstacle"):
  print(name, step(tg, name))
for name in ("ref_step_rwpo_double_well_d2", "ref_step_fp_gradient_d2", "ref_step_fp_lorenz_d3", "ref_step_ot_free_d2"):
  print(name, step(tr, name))
