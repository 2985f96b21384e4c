"""Host-side mirror of the reference's dimension-reduction trainer.

/root/reference/cnf_ot/dr/trainers.py:41-111 builds one or two UNCONDITIONAL RQS flows (cond_shape=(0,)) and
minimises the reconstruction error through a `sub_dim`-dimensional bottleneck:

  enc_dec :  y = encoder.forward(x); y[:, sub_dim:] = 0; x' = decoder.forward(y)      (:93-97)
  dec_only:  y = decoder.inverse(x); y[:, sub_dim:] = 0; x' = decoder.forward(y)      (:106-110)
  loss = mean_rows( sum_dims (x - x')^2 )

The flows are the same kernels as the MFC path (seam 1 of include/cnfot.h: cnfot_flow_*_ws and their VJPs) on a
blob whose t-rows are zero, called with t = 0; the bottleneck mask and the loss head are cnfot_mask_tail /
cnfot_recon_head.  Everything runs in libcnfot.so -- there is no CPU path.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch

from . import dist as _dist
from . import ops
from .flows import FlowModel, ParamTree, RQSFlow

_ZERO = (0.0, )


def build(dim: int, config: Dict, model: str = "enc_dec", device=None):
  """The flows trainers.py:41-68 creates from config["cnf"]: (encoder or None, decoder)."""
  c = config["cnf"]
  mk = lambda: RQSFlow(event_shape=(dim, ), num_layers=c["flow_num_layers"],
                       hidden_sizes=[c["hidden_size"]] * c["mlp_num_layers"], num_bins=c["num_bins"],
                       periodized=False, cond_shape=(0, ), device=device)
  if model == "enc_dec":
    return mk(), mk()
  if model == "dec_only":
    return None, mk()
  raise ValueError(f"unknown model: {model}")


def loss_fn(model: str, encoder: FlowModel, decoder: FlowModel, sub_dim: int) -> Callable:
  """The forward value of trainers.py's loss_fn(params, x)."""
  def fn(params, x):
    if model == "enc_dec":
      y = encoder.apply.forward(params["encoder"], x)
      dec = params["decoder"]
    else:
      y = decoder.apply.inverse(params, x)
      dec = params
    ops.mask_tail(y, sub_dim)
    xr = decoder.apply.forward(dec, y)
    return ops.recon_head(x, xr, x.shape[0])[0]
  return fn


def value_and_grad(model: str, encoder: FlowModel, decoder: FlowModel, sub_dim: int) -> Callable:
  """jax.value_and_grad(loss_fn)(params, data) of trainers.py:117.  Returns (loss, grads) with grads shaped like
  params (ParamTrees).  With torch.distributed initialised, x is this rank's row shard of a `global_rows` batch
  and loss / gradients are summed with one all-reduce each."""
  shape = decoder.shape

  def fn(params, x, global_rows: int = None) -> Tuple[torch.Tensor, object]:
    n = x.shape[0] if global_rows is None else global_rows
    if model == "enc_dec":
      enc_p, dec_p = params["encoder"], params["decoder"]
      y, _ = ops.flow_eval(shape, enc_p.blob, x, _ZERO, inverse=False, want_logdet=False)
    else:
      enc_p, dec_p = None, params
      y, _ = ops.flow_eval(shape, dec_p.blob, x, _ZERO, inverse=True, want_logdet=False)
    ops.mask_tail(y, sub_dim)
    xr, _ = ops.flow_eval(shape, dec_p.blob, y, _ZERO, inverse=False, want_logdet=False)
    loss, g_xr = ops.recon_head(x, xr, n)
    zeros = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
    g_y, g_dec = ops.flow_vjp(shape, dec_p.blob, y, _ZERO, g_xr, zeros, inverse=False)
    ops.mask_tail(g_y, sub_dim)
    if model == "enc_dec":
      _, g_enc = ops.flow_vjp(shape, enc_p.blob, x, _ZERO, g_y, zeros, inverse=False, want_g_in=False)
      loss = _dist.all_reduce_sum(loss.reshape(1))[0]
      return loss, {"encoder": ParamTree(shape, _dist.all_reduce_sum(g_enc)),
                    "decoder": ParamTree(shape, _dist.all_reduce_sum(g_dec))}
    _, g_inv = ops.flow_vjp(shape, dec_p.blob, x, _ZERO, g_y, zeros, inverse=True, want_g_in=False)
    g_dec.add_(g_inv)
    loss = _dist.all_reduce_sum(loss.reshape(1))[0]
    return loss, ParamTree(shape, _dist.all_reduce_sum(g_dec))

  return fn
