"""e2e of BASELINE cfg 2 through the C ABI with host buffers (cnfot_mfc_step_rng_host): wall time per call (dev tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import ops


class _D:
  world, rank, local = 1, 0, 0
  dev = torch.device("cuda", 0)
  td = None


torch.cuda.set_device(0)
w = bench.Workload("cfg2", _D())
n = w.shape.blob_size
hW = w.W.cpu().contiguous().pin_memory()
for pinned in (True, False):
  hout = torch.empty(n + 8, dtype=torch.float32)
  if pinned:
    hout = hout.pin_memory()
  for i in range(5):
    ops.mfc_step_rng_host(w.shape, w.problem, hW, 0x5EED, i, 1, w.lam, w.gB, w.gb, hout, device=_D.dev)
  ts = []
  for r in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20):
      ops.mfc_step_rng_host(w.shape, w.problem, hW, 0x5EED, 100 + i, 1, w.lam, w.gB, w.gb, hout, device=_D.dev)
    ts.append((time.perf_counter() - t0) / 20 * 1e6)
  ref = ops.mfc_step_rng(w.shape, w.problem, w.W, 0x5EED, 119, 1, w.lam, w.gB, w.gb).cpu()
  err = float((hout - ref).abs().max() / ref.abs().max())
  print(f"cnfot_mfc_step_rng_host, out {'pinned' if pinned else 'pageable'}: {sorted(ts)[2]:.1f} us per call (min {min(ts):.1f}); "
        f"vs device entry rel {err:.1e}", flush=True)
