"""N>1 path on CPU (gloo, world_size 2): row sharding + ONE all-reduce of the
[gradient | loss] buffer reproduces the whole-batch result.  The per-rank compute
is the kernels' row math run by the host harness (tests/hostsim), the reduction
and the sharding are the product's `cnf_ot_b200.dist`."""
import os
import socket

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from cnf_ot_b200 import dist
from cnf_ot_b200.layout import pack


def test_shard_partitions_rows():
  for n in (0, 1, 7, 128, 8192, 262144):
    for world in (1, 2, 3, 4, 8):
      parts = [dist.shard(n, r, world) for r in range(world)]
      assert parts[0].start == 0 and parts[-1].stop == n
      for a, b in zip(parts[:-1], parts[1:]):
        assert a.stop == b.start
      sizes = [p.stop - p.start for p in parts]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    dist.shard(10, 2, 2)


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  import sys
  here = os.path.dirname(os.path.abspath(__file__))
  for p in (os.path.dirname(here), here):
    if p not in sys.path:
      sys.path.insert(0, p)
  import hostsim as hs
  from cnf_ot_b200.ops import problem_desc
  from oracle import losses as olosses
  from util import make_cfg, make_inputs, make_params, shape_of
  torch.set_num_threads(1)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  td.init_process_group("gloo", rank=rank, world_size=world)
  try:
    cfg = make_cfg("rwpo", "double_well", B=256, Tn=2, lam=100.0)
    shape = shape_of(cfg)
    spec, params = make_params(cfg, 0.3)
    inputs = make_inputs(cfg)
    B, b = 256, 8
    r, w = dist.rank_world()
    assert (r, w) == (rank, world)
    rs, ss = dist.shard(B, r, w), dist.shard(b, r, w)
    pd = problem_desc(cfg)
    pdh = hs.ProblemDesc(pd.type, pd.subtype, pd.T, pd.beta, pd.a, pd.sigma, pd.dt, pd.dx)
    G, slots = hs.step(shape, pdh, pack(shape, params, torch.float64), inputs["latent"][rs],
                       inputs["latent"][:b][ss], inputs["src"][rs], inputs["tgt"][rs],
                       inputs["t_batch"], B, b, 100.0, dtype=torch.float64)
    buf = torch.cat([G, slots]).float()  # the fp32 [gradient | loss] buffer of the C ABI
    dist.all_reduce_sum(buf)
    if rank == 0:
      loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
      Gor = pack(shape, grads, torch.float64)
      n = shape.blob_size
      gerr = float((buf[:n].double() - Gor).abs().max() / Gor.abs().max())
      lerr = abs(float(buf[n:].double().sum()) - float(loss)) / abs(float(loss))
      q.put((gerr, lerr))
  finally:
    td.destroy_process_group()


def test_two_rank_allreduce_matches_whole_batch():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(180)
    assert p.exitcode == 0
  gerr, lerr = q.get(timeout=5)
  assert gerr < 5e-6 and lerr < 5e-6, (gerr, lerr)
