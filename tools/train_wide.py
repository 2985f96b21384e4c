"""solvers.main-style training on flows the fused kernels do not cover (wide-conditioner engine): the loss must go down."""
import os, sys, time, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from cnf_ot_b200 import solvers, _lib
cfg = yaml.safe_load(open(os.path.join(ROOT, "cnf_ot_b200", "config", "mfc.yaml")))
for typ, H, D, B, ep in (("ot", 64, 2, 4096, 200), ("rwpo", 64, 2, 2048, 100), ("ot", 512, 2, 4096, 60)):
  c = copy.deepcopy(cfg); c["general"]["type"] = typ; c["general"]["dim"] = D
  c["cnf"]["hidden_size"] = H; c["train"]["batch_size"] = B; c["train"]["epochs"] = ep
  torch.cuda.synchronize(); t0 = time.perf_counter()
  params, hist = solvers.main(c)
  torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / ep
  h = [float(v) for v in hist]
  print(f"{typ} hidden {H} dim {D} B={B}: engine {_lib.last_launch_info()['engine']}, {dt*1e3:.1f} ms/step wall, "
        f"loss {h[0]:.4e} -> min {min(h):.4e} / last {h[-1]:.4e}", flush=True)
  assert h[-1] < h[0] and all(v == v for v in h)
