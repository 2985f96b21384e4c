"""CPU restatement of the dimension-reduction loss of the reference (TEST INFRASTRUCTURE ONLY: imported by
tests/ and nothing else; the product path never touches it).

/root/reference/cnf_ot/dr/trainers.py:91-111 -- two UNCONDITIONAL flows (cond_shape=(0,), :41-68):

  enc_dec :  y = encoder.forward(x); y[:, sub_dim:] = 0; x' = decoder.forward(y)
  dec_only:  y = decoder.inverse(x); y[:, sub_dim:] = 0; x' = decoder.forward(y)
  loss = mean_rows( sum_dims (x - x')^2 )

Parity pinned to the reference's own `train()` run for one epoch in this container (its `loss_fn` and
`jax.value_and_grad(loss_fn)`, on torch-f64 stand-ins for jax / haiku / distrax / optax:
tests/golden/ref_dr_*.npz, tests/test_reference_golden.py, 1e-12), plus autograd-vs-finite-differences and the
identity-flow closed form (tests/test_oracle_dr.py).
"""
from __future__ import annotations

import torch

from . import flow as oflow


def reconstruction_loss(model: str, spec, params, x: torch.Tensor, sub_dim: int) -> torch.Tensor:
  """trainers.py:93-97 (enc_dec; params = {"encoder": ..., "decoder": ...}) / :106-110 (dec_only)."""
  if model == "enc_dec":
    y, _ = oflow.flow_forward_and_log_det(spec, params["encoder"], x)
    dec = params["decoder"]
  elif model == "dec_only":
    y, _ = oflow.flow_inverse_and_log_det(spec, params, x)
    dec = params
  else:
    raise ValueError(model)
  mask = torch.zeros(x.shape[-1], dtype=x.dtype)
  mask[:sub_dim] = 1.0
  xr, _ = oflow.flow_forward_and_log_det(spec, dec, y * mask)
  return ((x - xr)**2).sum(-1).mean()


def value_and_grad(model: str, spec, params, x: torch.Tensor, sub_dim: int):
  if model == "enc_dec":
    p = {k: oflow.clone_params(v, True) for k, v in params.items()}
    leaves = [(k, m, n) for k in p for m in p[k] for n in p[k][m]]
  else:
    p = oflow.clone_params(params, True)
  loss = reconstruction_loss(model, spec, p, x, sub_dim)
  loss.backward()
  if model == "enc_dec":
    grads = {k: {m: {n: v.grad for n, v in lv.items()} for m, lv in p[k].items()} for k in p}
  else:
    grads = {m: {n: v.grad for n, v in lv.items()} for m, lv in p.items()}
  return loss.detach(), grads
