"""Parity of the fused step against the CPU oracle at BASELINE hyper-parameters and sizes (builder tool):
  python tools/diag_parity.py [cfg2 cfg3 cfg4 ...]   (CNFOT_LIB / CNFOT_ENGINE select the build / engine)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from cnf_ot_b200 import _lib, ops


class _D:
  world, rank, local = 1, 0, 0
  dev = torch.device("cuda", 0)
  td = None


torch.cuda.set_device(0)
torch.set_num_threads(os.cpu_count() or 1)
for name in (sys.argv[1:] or ["cfg2", "cfg3"]):
  for rows in {"cfg2": (4096, 1 << 18), "cfg3": (4096, 1 << 16), "cfg4": (512, 4096), "cfg1": (4096, )}[name]:
    w = bench.Workload(name, _D(), oracle_rows=rows)
    s = w.sets[0]
    lat, sub, src, tgt, tb = w.args_of(s, 0)
    out = ops.mfc_step(w.shape, w.problem, w.W, lat, sub, src, tgt, tb, w.lam, w.gB, w.gb)
    torch.cuda.synchronize()
    r = bench.parity_against_oracle(w.cfg, w.shape, w.W, s, tb, w.lam, out)
    print(f"{name} rows {rows:7d} lib {os.environ.get('CNFOT_LIB', 'libcnfot.so')} engine {_lib.last_launch_info()['engine']}: "
          f"loss_rel {r['loss_rel']:.2e} grad_rel {r['grad_rel']:.2e} (oracle {r['oracle_seconds']:.1f} s)", flush=True)
