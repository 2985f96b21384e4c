"""Fused CUDA-core engine vs wide-conditioner engine on a hidden-32 flow both cover (ot/obstacle, D = 2, L = 2)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cnf_ot_b200 import ops, _lib
from cnf_ot_b200.layout import FlowShape
for H, D in ((64, 10), (32, 10), (64, 4)):
  shape = FlowShape(D, 2, 2, H, 5)
  g = torch.Generator(device="cuda").manual_seed(0)
  W = torch.randn(shape.blob_size, device="cuda", generator=g) * 0.02
  cfg = {"general": {"type": "ot", "dim": D, "dx": 0.01, "dt": 0.01}, "ot": {"subtype": "obstacle"}}
  prob = ops.problem_desc(cfg)
  for B in (4096, 1 << 16, 1 << 18):
    b = B // 32
    src = torch.randn(B, D, device="cuda", generator=g); tgt = torch.randn(B, D, device="cuda", generator=g)
    sub = torch.randn(b, D, device="cuda", generator=g)
    out = torch.empty(shape.blob_size + 8, device="cuda")
    for eng in (None, "wide"):
      if eng: os.environ["CNFOT_ENGINE"] = eng
      else: os.environ.pop("CNFOT_ENGINE", None)
      f = lambda: ops.mfc_step(shape, prob, W, None, sub, src, tgt, [0.37], 5000.0, B, b, out=out)
      for _ in range(3): f()
      torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(10): f()
      e1.record(); torch.cuda.synchronize()
      print(f"H={H} D={D} B={B}: engine {_lib.last_launch_info()['engine']:5s} {e0.elapsed_time(e1)/10:.3f} ms/step  loss {float(out[shape.blob_size]):.6g}", flush=True)
