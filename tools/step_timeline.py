"""Where the time of one fused train step goes (builder tool): %globaltimer stamps written by the kernel
(cnfot_debug_step_timeline) for BASELINE cfg 2, on one GPU or under torchrun (fused all-reduce over peer memory).

  python tools/step_timeline.py [cfg2] [--rows=R] [--update]
  torchrun --nproc-per-node 2 tools/step_timeline.py cfg2
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import _lib, ops

name = next((a for a in sys.argv[1:] if not a.startswith("--")), "cfg2")
rows = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--rows=")), None)
world = int(os.environ.get("WORLD_SIZE", "1"))


class _D:
  world, rank, local = world, int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
  dev = torch.device("cuda", local)
  td = None


torch.cuda.set_device(_D.local)
if world > 1:
  import torch.distributed as td
  td.init_process_group("nccl", device_id=_D.dev)
  _D.td = td
w = bench.Workload(name, _D(), rows_override=rows)
w.attach_peer_exchange()
lib = _lib.load()
N = 40
init = torch.tensor([-1, 0, -1, 0, 0, 0, 0, 0], dtype=torch.int64)
words = torch.empty(N, 8, dtype=torch.int64, device=_D.dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
update = "--update" in sys.argv
state = ops.TrainState(w.shape, w.W, 1234, peers=w.px) if update else None
rs = slice(_D.rank * w.B, (_D.rank + 1) * w.B)
ss = slice(_D.rank * w.b, (_D.rank + 1) * w.b)


def step(i):
  if update:
    ops.mfc_update(w.shape, w.problem, state, w.W, 1, w.lam, w.gB, w.gb, 1e-4, rows_B=rs, rows_b=ss)
  else:
    w.step(i)


for i in range(5):
  step(i)
if world > 1:
  _D.td.barrier()
torch.cuda.synchronize()
for i in range(N):
  words[i].copy_(init, non_blocking=True)
torch.cuda.synchronize()
ev[0].record()
for i in range(N):
  lib.cnfot_debug_step_timeline(words[i].data_ptr())
  step(i)
  ev[i + 1].record()
lib.cnfot_debug_step_timeline(None)
torch.cuda.synchronize()
t = words.cpu().double()[5:]
span = t[:, 5] - t[:, 0]
per = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(5, N)]) * 1e3
gap = (t[1:, 0] - t[:-1, 5])
med = lambda x: float(x.median()) / 1e3
print(f"[rank {_D.rank}] {name} {'update' if update else 'step'} rows {w.B} world {world}: event time per step {float(per.median()):.1f} us; kernel span {med(span):.1f} us = "
      f"setup {med(t[:, 1] - t[:, 0]):.1f} (launch ramp {med(t[:, 7] - t[:, 0]):.1f}, longest CTA setup {med(t[:, 6]):.1f}) + tiles {med(t[:, 2] - t[:, 1]):.1f} (first CTA idle) .. {med(t[:, 3] - t[:, 1]):.1f} (last CTA idle) "
      f"+ flush {med(t[:, 4] - t[:, 3]):.1f} + tail {med(t[:, 5] - t[:, 4]):.1f}; gap between kernels {med(gap):.1f} us", flush=True)
if world > 1:
  _D.td.barrier(); _D.td.destroy_process_group()
