"""Time the wide-conditioner engine on BASELINE config 5's flow (D = 32, 16 layers, hidden 512, ot/free) at a
reduced batch: python tools/time_wide.py [rows_B] [reps].  Prints ms/step and samples/s of one GPU."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from cnf_ot_b200 import ops, _lib
from cnf_ot_b200.layout import FlowShape

rows_B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
D, L, H = int(os.environ.get("WIDE_D", 32)), int(os.environ.get("WIDE_L", 16)), int(os.environ.get("WIDE_H", 512))
shape = FlowShape(D, L, 2, H, 5)
b = rows_B // 32
g = torch.Generator(device="cuda").manual_seed(0)
W = torch.randn(shape.blob_size, device="cuda", generator=g) * 0.02
src = torch.randn(rows_B, D, device="cuda", generator=g) - 3.0
tgt = torch.randn(rows_B, D, device="cuda", generator=g)
sub = torch.randn(b, D, device="cuda", generator=g)
cfg = {"general": {"type": "ot", "dim": D, "dx": 0.01, "dt": 0.01}, "ot": {"subtype": "free"}}
prob = ops.problem_desc(cfg)
out = torch.empty(shape.blob_size + 8, device="cuda")
def step():
  ops.mfc_step(shape, prob, W, None, sub, src, tgt, [0.37], 5000.0, rows_B, b, out=out)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# algorithmic conditioner flops: 2 L sum_d [(d+1) H + H^2 + 16 H] per row per flow pass, x3 (fwd, dgrad, wgrad)
per_pass = 2 * L * sum((d + 1) * H + H * H + 16 * H for d in range(1, D))
passes = 2 * rows_B + 2 * b
print("D %d L %d H %d rows_B %d b %d params %.1f M: %.1f ms/step, %.0f samples/s, loss %.6g, %.1f TFLOP/s algorithmic (3x per pass), engine %s, acc2=%s chunk=%s" % (
  D, L, H, rows_B, b, shape.blob_size / 1e6, ms, rows_B / ms * 1e3, float(out[shape.blob_size]),
  3 * per_pass * passes / ms / 1e9, _lib.last_launch_info()["engine"], os.environ.get("CNFOT_DENSE_ACC2", "0"),
  os.environ.get("CNFOT_WIDE_CHUNK", "default")))
