"""A few train steps of the benchmark workload (short, for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import FlowShape

typ = sys.argv[1] if len(sys.argv) > 1 else "ot"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
if typ == "ot":
  B = 1 << 18; shape = FlowShape(2, 2, 2, 16, 5); cfg = bench.workload_cfg(B)
elif typ == "rwpo":
  B = 1 << 20; shape = FlowShape(2, 2, 2, 16, 5); cfg = bench.workload_cfg(B); cfg["general"]["type"] = "rwpo"
else:
  B = 1 << 19; shape = FlowShape(10, 2, 2, 16, 5); cfg = bench.workload_cfg(B); cfg["general"].update(type="fp", dim=10)
b = B // 32
D = shape.dim
W = bench.make_blob(shape, dev) if D == 2 else torch.randn(shape.blob_size, device=dev) * 0.05
g = torch.Generator(device=dev).manual_seed(1)
lat = torch.randn(B, D, device=dev, generator=g); sub = torch.randn(b, D, device=dev, generator=g)
src = lat + 3.0; tgt = torch.randn(B, D, device=dev, generator=g)
pd = ops.problem_desc(cfg)
out = torch.empty(shape.blob_size + 8, device=dev)
for i in range(n):
  ops.mfc_step(shape, pd, W, None if typ == "ot" else lat, sub, src if typ == "ot" else None,
               tgt if typ == "ot" else None, [0.37], 5000.0, B, b, out=out)
torch.cuda.synchronize()
print("loss", float(out[shape.blob_size]))
