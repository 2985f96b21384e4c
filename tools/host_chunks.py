"""Time the host-buffer entry (cnfot_mfc_step_host) for different H2D chunk counts (dev tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import FlowShape
dev = torch.device("cuda", 0)
B = 1 << 18; b = B // 32
shape = FlowShape(2, 2, 2, 16, 5); cfg = bench.workload_cfg(B)
W = bench.make_blob(shape, dev)
pd = ops.problem_desc(cfg)
pin = lambda x: x.cpu().contiguous().pin_memory()
g = torch.Generator().manual_seed(1)
sets = [(pin(torch.randn(B, 2, generator=g) + 3), pin(torch.randn(B, 2, generator=g)), pin(torch.randn(b, 2, generator=g))) for _ in range(3)]
hW = pin(W); hout = torch.empty(shape.blob_size + 8).pin_memory()
for nc in ("1", "2", "3", "4", "1"):
  os.environ["CNFOT_HOST_CHUNKS"] = nc
  def step(i):
    src, tgt, sub = sets[i % 3]
    ops.mfc_step_host(shape, pd, hW, None, sub, src, tgt, [0.37], 5000.0, B, b, hout, device=dev)
  for i in range(5): step(i)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for i in range(50): step(i)
  torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 50
  print(f"chunks={nc}: {el*1e6:.1f} us/step  loss {float(hout[shape.blob_size]):.6e}")
