"""Host-side mirror of the training part of /root/reference/cnf_ot/mfc/solvers.py
(:26-129): model construction, loss selection by `general.type`, the `update`
step (value_and_grad + Adam) and the training loop.  The evaluation / plotting
tail of the reference's `main` (:131-493) is out of scope.

    from cnf_ot_b200 import solvers
    params, loss_hist = solvers.main(yaml.safe_load(open("cnf_ot_b200/config/mfc.yaml")))
"""
from __future__ import annotations

from functools import partial
from typing import Dict, Tuple

import torch

from . import applications, ops, random
from .flows import ParamTree, RQSFlow, multi_transform, without_apply_rng


class AdamState:
  """optax.adam state (solvers.py:55-56): step count and the two moment blobs."""

  def __init__(self, params: ParamTree):
    self.count = 0
    self.mu = torch.zeros_like(params.blob)
    self.nu = torch.zeros_like(params.blob)


def build(config: Dict):
  """Model + loss selection, solvers.py:29-88.  Returns (model, loss_fn, T)."""
  g = config["general"]
  _type, dim, dt, dx = g["type"], g["dim"], g["dt"], g["dx"]
  t_batch_size = g["t_batch_size"]
  c = config["cnf"]
  model = RQSFlow(
    event_shape=(dim, ),
    num_layers=c["flow_num_layers"],
    hidden_sizes=[c["hidden_size"]] * c["mlp_num_layers"],
    num_bins=c["num_bins"],
    periodized=False,
  )
  model = without_apply_rng(multi_transform(model))
  if _type == "rwpo":
    r = config["rwpo"]
    T = r["T"]
    loss_fn = partial(applications.rwpo_loss_fn, model, dim, T, r["beta"], dt, dx, t_batch_size,
                      r["pot_type"], r["a"])
  elif _type == "fp":
    f = config["fp"]
    T = f["T"]
    loss_fn = partial(applications.fp_loss_fn, model, dim, T, f["a"], f["sigma"], dt, dx,
                      t_batch_size, f["velocity_field_type"])
  elif _type == "ot":
    T = 1
    loss_fn = partial(applications.ot_loss_fn, model, dim, T, dt, t_batch_size,
                      config["ot"]["subtype"])
  else:
    raise Exception(f"Unknown problem type: {_type}...")
  return model, loss_fn, T


def make_update(loss_fn, lr: float, batch_size: int):
  """solvers.py:90-97.  The parameter blob is updated IN PLACE (the returned params are
  the same object): B200-side the 'new pytree per step' of the functional original would
  only add a copy."""
  vg = applications.value_and_grad(loss_fn)

  def update(params: ParamTree, rng, _lambda, opt_state: AdamState) -> Tuple[torch.Tensor, ParamTree, AdamState]:
    loss, grads = vg(params, rng, _lambda, batch_size)
    opt_state.count += 1
    ops.adam_update(params.blob, grads.blob, opt_state.mu, opt_state.nu, lr, opt_state.count)
    return loss, params, opt_state

  return update


def main(config_dict: Dict, progress: bool = False):
  """Training loop of solvers.py:26-129; returns (params, loss_hist)."""
  config = config_dict
  rng = random.PRNGKey(config["general"]["seed"])
  tr = config["train"]
  epochs, batch_size, eval_frequency, _lambda = tr["epochs"], tr["batch_size"], tr["eval_frequency"], tr["_lambda"]
  model, loss_fn, T = build(config)
  dim = config["general"]["dim"]
  model_rng, rng = random.split(rng)
  params = model.init(model_rng, torch.zeros(1, dim), torch.zeros(1))
  opt_state = AdamState(params)
  update = make_update(loss_fn, tr["lr"], batch_size)
  loss_hist = []
  for step in range(epochs):
    update_rng, rng = random.split(rng)
    loss, params, opt_state = update(params, update_rng, _lambda, opt_state)
    loss_hist.append(loss)  # device scalar: no host sync in the hot loop (solvers.py:106)
    if progress and step % eval_frequency == 0:
      desc = f"step {step} loss={float(loss):.4e}"
      if config["general"]["type"] == "ot":
        eval_rng, rng = random.split(rng)
        KL = applications.density_fit_kl_loss_fn(model, dim, T, params, eval_rng, batch_size)
        desc += f" KL={float(KL):.4f}"
      print(desc, flush=True)
  return params, loss_hist


if __name__ == "__main__":
  import os
  import yaml
  here = os.path.dirname(os.path.abspath(__file__))
  with open(os.path.join(here, "config", "mfc.yaml"), "r") as file:
    config_dict = yaml.safe_load(file)
  main(config_dict, progress=True)
