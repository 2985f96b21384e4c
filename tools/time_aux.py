"""Timing of the kernels around the step (dev tool; CNFOT_LIB selects a build): densities, evaluation energies, the
model-API forward / inverse and their VJPs on the mfc.yaml flow."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import FlowShape

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
print("lib", os.environ.get("CNFOT_LIB", "libcnfot.so"), bench.density_block(dev))
shape = FlowShape(2, 2, 2, 16, 5)
W = bench.make_blob(shape, dev, 0.3)
n = 1 << 20
x = torch.randn(n, 2, device=dev); t = torch.rand(n, device=dev)
g = torch.randn(n, 2, device=dev); gl = torch.randn(n, device=dev)
for name, fn in (("flow_forward 2^20", lambda i: ops.flow_eval(shape, W, x, t, inverse=False)),
                 ("flow_inverse 2^20", lambda i: ops.flow_eval(shape, W, x, t, inverse=True)),
                 ("flow_forward_vjp 2^20", lambda i: ops.flow_vjp(shape, W, x, t, g, gl, inverse=False)),
                 ("flow_inverse_vjp 2^20", lambda i: ops.flow_vjp(shape, W, x, t, g, gl, inverse=True))):
  for i in range(3):
    fn(i)
  el = bench.time_region(fn, 10, torch.cuda.synchronize) / 10
  print(f"  {name}: {el * 1e6:.1f} us")
lat = torch.randn(65536, 2, device=dev)
ts = torch.linspace(0.01, 0.99, 64).tolist()
for ws in (False, True):
  fn = lambda i: ops.kinetic_energy(shape, W, lat, ts, 0.01, with_score=ws, kappa=0.1, dx=0.01)
  try:
    for i in range(2):
      fn(i)
    el = bench.time_region(fn, 5, torch.cuda.synchronize) / 5
    print(f"  kinetic_energy 65536 x 64 times with_score={ws}: {el * 1e6:.1f} us")
  except Exception as exc:
    print("  kinetic_energy:", repr(exc)[:200])
