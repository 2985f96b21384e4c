"""Where the time of the host-buffer step goes (dev tool): H2D alone, kernel alone, Python overhead."""
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
W = bench.make_blob(shape, dev); pd = ops.problem_desc(cfg)
pin = lambda x: x.cpu().contiguous().pin_memory()
g = torch.Generator().manual_seed(1)
src, tgt, sub = pin(torch.randn(B, 2, generator=g) + 3), pin(torch.randn(B, 2, generator=g)), pin(torch.randn(b, 2, generator=g))
hW = pin(W); hout = torch.empty(shape.blob_size + 8).pin_memory()
dsrc, dtgt, dsub = src.cuda(), tgt.cuda(), sub.cuda()
def timeit(fn, n=50):
  for _ in range(5): fn()
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): fn()
  torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def h2d():
  dsrc.copy_(src, non_blocking=True); dtgt.copy_(tgt, non_blocking=True); dsub.copy_(sub, non_blocking=True)
  torch.cuda.synchronize()
print(f"H2D 3 tensors ({(src.numel()+tgt.numel()+sub.numel())*4/1e6:.2f} MB) + sync: {timeit(h2d):.1f} us")
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device=dev)
def h2d_big():
  dbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
us = timeit(h2d_big, 10); print(f"H2D 64 MiB: {us:.1f} us -> {64*1.048576/us*1e3:.1f} GB/s")
out = torch.empty(shape.blob_size + 8, device=dev)
def dev_step():
  ops.mfc_step(shape, pd, W, None, dsub, dsrc, dtgt, [0.37], 5000.0, B, b, out=out); torch.cuda.synchronize()
print(f"device step + sync: {timeit(dev_step):.1f} us")
def dev_step_d2h():
  ops.mfc_step(shape, pd, W, None, dsub, dsrc, dtgt, [0.37], 5000.0, B, b, out=out); hout.copy_(out, non_blocking=True); torch.cuda.synchronize()
print(f"device step + D2H + sync: {timeit(dev_step_d2h):.1f} us")
def host_step():
  ops.mfc_step_host(shape, pd, hW, None, sub, src, tgt, [0.37], 5000.0, B, b, hout, device=dev)
print(f"host step: {timeit(host_step):.1f} us")
e = src[:0]
def host_step_empty():
  ops.mfc_step_host(shape, pd, hW, None, sub[:0], e, e, [0.37], 5000.0, B, b, hout, device=dev)
print(f"host step, zero rows (API + Python overhead): {timeit(host_step_empty):.1f} us")
for zc in ("0", "1", "0", "1"):
  os.environ["CNFOT_HOST_ZEROCOPY"] = zc
  print(f"host step zero-copy={zc}: {timeit(host_step):.1f} us   loss {float(hout[shape.blob_size]):.6e}")
sets = [(pin(torch.randn(B, 2, generator=g) + 3), pin(torch.randn(B, 2, generator=g)), pin(torch.randn(b, 2, generator=g))) for _ in range(8)]
it = [0]
def host_step_rot():
  s_, t_, u_ = sets[it[0] % 8]; it[0] += 1
  ops.mfc_step_host(shape, pd, hW, None, u_, s_, t_, [0.37], 5000.0, B, b, hout, device=dev)
for zc in ("0", "1", "0", "1"):
  os.environ["CNFOT_HOST_ZEROCOPY"] = zc
  print(f"host step, 8 rotating host sets, zero-copy={zc}: {timeit(host_step_rot):.1f} us")
