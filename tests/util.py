"""Shared builders for the parity tests: seeded parameters and step inputs,
identical for the oracle (float64, CPU) and the kernels (float32, CUDA)."""
import copy

import torch

from cnf_ot_b200.layout import FlowShape, pack
from oracle import flow as oflow
from oracle import losses as olosses

BASE_CFG = {
  "general": {"type": "ot", "dim": 2, "dx": 0.01, "dt": 0.01, "t_batch_size": 1, "seed": 42},
  "ot": {"subtype": "free"},
  "rwpo": {"T": 2, "beta": 10, "a": 1, "pot_type": "double_well"},
  "fp": {"T": 1, "a": 1, "sigma": 0.5, "velocity_field_type": "gradient"},
  "cnf": {"flow_num_layers": 2, "mlp_num_layers": 2, "hidden_size": 16, "num_bins": 5},
  "train": {"epochs": 1, "lr": 1e-3, "_lambda": 5000.0, "batch_size": 2048, "eval_frequency": 100},
}


def make_cfg(typ="ot", sub=None, dim=2, L=2, M=2, H=16, K=5, B=256, Tn=1, lam=5000.0, **over):
  cfg = copy.deepcopy(BASE_CFG)
  cfg["general"].update(type=typ, dim=dim, t_batch_size=Tn)
  cfg["cnf"].update(flow_num_layers=L, mlp_num_layers=M, hidden_size=H, num_bins=K)
  cfg["train"].update(batch_size=B, _lambda=lam)
  if sub is not None:
    key = {"ot": "subtype", "rwpo": "pot_type", "fp": "velocity_field_type"}[typ]
    cfg[typ][key] = sub
  for k, v in over.items():
    sec, name = k.split("__")
    cfg[sec][name] = v
  return cfg


def shape_of(cfg) -> FlowShape:
  c = cfg["cnf"]
  return FlowShape(cfg["general"]["dim"], c["flow_num_layers"], c["mlp_num_layers"],
                   c["hidden_size"], c["num_bins"])


def make_params(cfg, sigma, seed=3):
  """Reference init + N(0, sigma^2) on biases / output layers / first.
  Rounded to float32 so the oracle and the kernels see the same numbers."""
  spec = olosses.spec_from_config(cfg)
  params = oflow.perturb_params(oflow.init_params(spec, seed=seed), sigma)
  for mod in params:
    for k in params[mod]:
      params[mod][k] = params[mod][k].to(torch.float32).to(params[mod][k].dtype)
  return spec, params


def make_inputs(cfg, seed=42):
  """Synthetic step inputs: latent N(0,I), mixture / Gaussian data,
  uniform times; float32-representable float64 tensors."""
  g = torch.Generator().manual_seed(seed)
  B, D = cfg["train"]["batch_size"], cfg["general"]["dim"]
  typ = cfg["general"]["type"]
  horizon = 1.0 if typ == "ot" else float(cfg[typ]["T"])
  if D == 2:
    src, tgt = olosses.source_mixture(g, B, D)
  else:
    src, tgt = olosses.source_gaussian(g, B, D)
  f32 = lambda t: t.to(torch.float32).to(torch.float64)
  return {
    "latent": f32(torch.randn(B, D, generator=g, dtype=torch.float64)),
    "src": f32(src), "tgt": f32(tgt),
    "t_batch": f32(torch.rand(cfg["general"]["t_batch_size"], generator=g, dtype=torch.float64) * horizon),
  }


def blob(shape, params, dtype=torch.float32):
  return pack(shape, params, dtype)


def rel_err(a, b):
  """max |a - b| / (|b| + 1): absolute near zero, relative for large values."""
  a = a.detach().to("cpu", torch.float64)
  b = b.detach().to("cpu", torch.float64)
  return float(((a - b).abs() / (b.abs() + 1.0)).max()) if a.numel() else 0.0
