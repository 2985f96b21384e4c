"""A few train steps of a BASELINE config (short: for ncu, and for A/B timing of tuning knobs).

  python tools/prof_step.py [cfg2|cfg1|cfg3|cfg4] [n] [explicit|rng|update] [--time] [--rows=R]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import _lib, ops

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "explicit"
timing = "--time" in sys.argv


class _D:
  world, rank, local = 1, 0, 0
  dev = torch.device("cuda", 0)
  td = None


torch.cuda.set_device(0)
rows = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--rows=")), None)
w = bench.Workload(name, _D(), rows_override=rows or ((1 << 19) if name == "cfg4" else None))
state = ops.TrainState(w.shape, w.W, 1234) if mode == "update" else None


def step(i):
  if mode == "explicit":
    w.step(i)
  elif mode == "rng":
    ops.mfc_step_rng(w.shape, w.problem, w.W, 1234, i, 1, w.lam, w.gB, w.gb, out=w.out)
  else:
    ops.mfc_update(w.shape, w.problem, state, w.W, 1, w.lam, w.gB, w.gb, 1e-4, out=w.out)


for i in range(3 if timing else 0):
  step(i)
torch.cuda.synchronize()
if timing:
  ts = []
  for r in range(5):
    ts.append(bench.time_region(step, n, torch.cuda.synchronize, first=r * n) / n * 1e3)
  print(f"{name} {mode} rows {w.B} env[{os.environ.get('CNFOT_STEP_UNIT', '-')},{os.environ.get('CNFOT_STEP_ROWS', '-')},{os.environ.get('CNFOT_ENGINE', '-')}]: "
        f"{sorted(ts)[2]:.4f} ms/step (min {min(ts):.4f}, max {max(ts):.4f}); {_lib.last_launch_info()}; loss {float(w.out[w.shape.blob_size]):.6e}")
else:
  for i in range(n):
    step(i)
  torch.cuda.synchronize()
  print("loss", float(w.out[w.shape.blob_size]))
