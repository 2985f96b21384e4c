"""Multi-GPU check (run under torchrun, one rank per GPU): the fused step + all-reduce
(cnfot_mfc_step_dp over peer-mapped memory) against cnfot_mfc_step + NCCL all-reduce, and timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as td
import bench
from cnf_ot_b200 import ops, dist
from cnf_ot_b200.layout import FlowShape

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
shape = FlowShape(2, 2, 2, 16, 5)
B = 1 << 16; b = B // 32
cfg = bench.mfc_cfg("ot", "obstacle", 2, B * world)
pd = ops.problem_desc(cfg)
W = bench.make_blob(shape, dev, 0.3); td.broadcast(W, 0)
g = torch.Generator(device=dev).manual_seed(100 + rank)
src = torch.randn(B, 2, device=dev, generator=g) + 3; tgt = torch.randn(B, 2, device=dev, generator=g)
sub = torch.randn(b, 2, device=dev, generator=g)
px = dist.PeerExchange(shape, dev)
ok = True
for it in range(6):
  t = [0.1 + 0.1 * it]
  ref = ops.mfc_step(shape, pd, W, None, sub, src, tgt, t, 5000.0, B * world, b * world).clone()
  td.all_reduce(ref)
  if it == 3 and rank == world - 1:   # an empty shard still takes part
    got = ops.mfc_step(shape, pd, W, None, sub[:0], src[:0], tgt[:0], t, 5000.0, B * world, b * world, peers=px).clone()
  else:
    got = ops.mfc_step(shape, pd, W, None, sub, src, tgt, t, 5000.0, B * world, b * world, peers=px).clone()
  if it == 3:
    # reference for the empty-shard step: ranks 0..W-2 only
    part = ops.mfc_step(shape, pd, W, None, sub, src, tgt, t, 5000.0, B * world, b * world).clone()
    if rank == world - 1: part.zero_()
    td.all_reduce(part); ref = part
  err = float((got - ref).abs().max() / ref.abs().max())
  # every rank must hold the bit-identical result
  gathered = [torch.empty_like(got) for _ in range(world)]
  td.all_gather(gathered, got)
  same = all(torch.equal(gathered[0], x) for x in gathered)
  if rank == 0: print(f"step {it}: fused vs NCCL rel err {err:.2e}; identical on all ranks: {same}", flush=True)
  ok = ok and err < 2e-6 and same
# a small step (the reference's default batch): the kernel instantiation whose kinetic rows are spread over lane groups
Bs = 2048 // world; bs = Bs // 32
cfg_s = bench.mfc_cfg("rwpo", "double_well", 2, Bs * world)
pd_s = ops.problem_desc(cfg_s)
lat_s = torch.randn(Bs, 2, device=dev, generator=g); sub_s = torch.randn(max(bs, 1), 2, device=dev, generator=g)[:bs]
ref = ops.mfc_step(shape, pd_s, W, lat_s, sub_s, None, None, [0.4], 500.0, Bs * world, bs * world).clone()
td.all_reduce(ref)
got = ops.mfc_step(shape, pd_s, W, lat_s, sub_s, None, None, [0.4], 500.0, Bs * world, bs * world, peers=px).clone()
err = float((got - ref).abs().max() / ref.abs().max())
if rank == 0: print(f"small step (B = {Bs * world}, rwpo): fused vs NCCL rel err {err:.2e}", flush=True)
ok = ok and err < 2e-6
# on-chip draws: the sharded fused step (global row indices) == the whole batch on one GPU
gB, gb = B * world, b * world
rs, ss = dist.shard(gB, rank, world), dist.shard(gb, rank, world)
whole = ops.mfc_step_rng(shape, pd, W, 1234, 3, 1, 5000.0, gB, gb).clone()
got = ops.mfc_step_rng(shape, pd, W, 1234, 3, 1, 5000.0, gB, gb, rows_B=rs, rows_b=ss, peers=px).clone()
err = float((got - whole).abs().max() / whole.abs().max())
if rank == 0: print(f"on-chip draws: sharded fused step vs whole batch on one GPU rel err {err:.2e}", flush=True)
ok = ok and err < 2e-6
# device-resident update: K sharded updates with the all-reduce and Adam inside the kernel == K whole-batch updates
K = 8
Wd, Ws = W.clone(), W.clone()
st_d = ops.TrainState(shape, Wd, 99, peers=px)
st_s = ops.TrainState(shape, Ws, 99)
hd, hs_ = torch.zeros(K, device=dev), torch.zeros(K, device=dev)
for _ in range(K):
  ops.mfc_update(shape, pd, st_d, Wd, 1, 5000.0, gB, gb, 1e-3, rows_B=rs, rows_b=ss, loss_hist=hd)
  ops.mfc_update(shape, pd, st_s, Ws, 1, 5000.0, gB, gb, 1e-3, loss_hist=hs_)
torch.cuda.synchronize()
werr = float((Wd - Ws).abs().max())
lerr = float(((hd - hs_).abs() / hs_.abs()).max())
gw = [torch.empty_like(Wd) for _ in range(world)]
td.all_gather(gw, Wd)
same_w = all(torch.equal(gw[0], x) for x in gw)
if rank == 0:
  print(f"device-resident update x{K}: sharded vs whole-batch weights max abs diff {werr:.2e}, loss rel {lerr:.2e}; "
        f"weights identical on all ranks: {same_w}; status {st_d.status()}", flush=True)
ok = ok and werr < 5e-4 and lerr < 1e-5 and same_w and st_d.status() == 0 and st_d.step_count() == K
# a CUDA graph of sharded updates replays across ranks
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
  for _ in range(4):
    ops.mfc_update(shape, pd, st_d, Wd, 1, 5000.0, gB, gb, 1e-3, rows_B=rs, rows_b=ss)
for _ in range(3):
  g.replay()
  st_d.steps_issued += 4; px.epoch += 4
torch.cuda.synchronize()
st_d.steps_issued -= 4; px.epoch -= 4   # the capture itself ran nothing
for _ in range(12):
  ops.mfc_update(shape, pd, st_s, Ws, 1, 5000.0, gB, gb, 1e-3)
torch.cuda.synchronize()
werr = float((Wd - Ws).abs().max())
if rank == 0: print(f"graph of 4 sharded updates x3 replays vs 12 whole-batch updates: weights max abs diff {werr:.2e}; steps {st_d.step_count()}", flush=True)
ok = ok and werr < 2e-3 and st_d.step_count() == K + 12 and st_d.status() == 0
# solvers.main under torchrun: the reference's training loop, device-resident, row shards + the fused exchange, CUDA graphs
import copy, yaml
from cnf_ot_b200 import solvers
mcfg = yaml.safe_load(open(os.path.join(ROOT, "cnf_ot_b200", "config", "mfc.yaml")))
mcfg = copy.deepcopy(mcfg); mcfg["train"]["epochs"] = 60; mcfg["train"]["batch_size"] = 2048; mcfg["train"]["eval_frequency"] = 20
params, hist = solvers.main(mcfg)
torch.cuda.synchronize()
gp = [torch.empty_like(params.blob) for _ in range(world)]
td.all_gather(gp, params.blob)
same_p = all(torch.equal(gp[0], x) for x in gp)
hist = torch.as_tensor(hist).float()
if rank == 0:
  print(f"solvers.main x60 under torchrun (mfc.yaml, B = 2048): loss {float(hist[0]):.3e} -> {float(hist[-1]):.3e}; "
        f"parameters identical on all ranks: {same_p}", flush=True)
ok = ok and same_p and bool(torch.isfinite(hist).all()) and float(hist[-1]) < float(hist[0])
out = torch.empty(shape.blob_size + 8, device=dev)
def timed(fn, n=100):
  for _ in range(5): fn()
  td.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n * 1e3
def nccl_step():
  ops.mfc_step(shape, pd, W, None, sub, src, tgt, [0.3], 5000.0, B * world, b * world, out=out); td.all_reduce(out)
def fused_step():
  ops.mfc_step(shape, pd, W, None, sub, src, tgt, [0.3], 5000.0, B * world, b * world, out=out, peers=px)
def local_step():
  ops.mfc_step(shape, pd, W, None, sub, src, tgt, [0.3], 5000.0, B * world, b * world, out=out)
def update_step():
  ops.mfc_update(shape, pd, st_d, Wd, 1, 5000.0, gB, gb, 1e-3, rows_B=rs, rows_b=ss)
u = timed(update_step)   # before the stateless fused calls: those advance the exchange's epoch past the state's
a, c, l = timed(nccl_step), timed(fused_step), timed(local_step)
if rank == 0:
  print(f"B/GPU={B}: local step {l:.1f} us, + NCCL all-reduce {a:.1f} us, fused step+all-reduce {c:.1f} us, "
        f"device-resident update (draws + step + all-reduce + Adam) {u:.1f} us", flush=True)
  print("CHECK_DP", "PASS" if ok else "FAIL", flush=True)
td.barrier(); td.destroy_process_group()
