"""Launch the four stand-alone spline kernels (2^24 rows, K = 5) a few times each: the command ncu captures
(profiles/r01_rqs_ncu_full.txt).  Prints CUDA-event times when run plainly."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cnf_ot_b200 import _lib
lib = _lib.load()
n, K, P = 1 << 24, 5, 16
theta = torch.randn(n, P, device="cuda") * 0.3
v = torch.randn(n, device="cuda") * 3
y, ld = torch.empty_like(v), torch.empty_like(v)
go, gl = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
gi, gp = torch.empty_like(v), torch.empty_like(theta)
s = torch.cuda.current_stream().cuda_stream
calls = {
  "rqs_forward": (lambda: lib.cnfot_rqs_forward(s, v.data_ptr(), theta.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, y.data_ptr(), ld.data_ptr(), 0), 4 * P + 12),
  "rqs_inverse": (lambda: lib.cnfot_rqs_inverse(s, v.data_ptr(), theta.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, y.data_ptr(), ld.data_ptr(), 0), 4 * P + 12),
  "rqs_forward_vjp": (lambda: lib.cnfot_rqs_forward_vjp(s, v.data_ptr(), theta.data_ptr(), go.data_ptr(), gl.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, gi.data_ptr(), gp.data_ptr()), 8 * P + 16),
  "rqs_inverse_vjp": (lambda: lib.cnfot_rqs_inverse_vjp(s, v.data_ptr(), theta.data_ptr(), go.data_ptr(), gl.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, gi.data_ptr(), gp.data_ptr()), 8 * P + 16),
}
for name, (fn, bpr) in calls.items():
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(3):
    fn()
  e1.record(); torch.cuda.synchronize()
  us = e0.elapsed_time(e1) / 3 * 1e3
  print(f"{name}: {us:.1f} us  {n * bpr / us / 1e3:.0f} GB/s algorithmic ({bpr} B/row)")
