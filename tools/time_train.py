"""Wall time per training step of solvers.main-style training (dev tool)."""
import os, sys, time, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from cnf_ot_b200 import solvers, random
cfg = yaml.safe_load(open(os.path.join(ROOT, "cnf_ot_b200", "config", "mfc.yaml")))
for typ, B in (("ot", 2048), ("ot", 4096), ("rwpo", 2048), ("fp", 2048), ("ot", 65536)):
  c = copy.deepcopy(cfg); c["general"]["type"] = typ; c["train"]["batch_size"] = B; c["train"]["epochs"] = 300
  solvers.main(c)  # warm-up
  torch.cuda.synchronize(); t0 = time.perf_counter()
  params, hist = solvers.main(c)
  torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 300
  print(f"{typ} B={B}: {dt*1e6:.0f} us/step wall (python loop), loss {float(hist[0]):.3e} -> {float(hist[-1]):.3e}", flush=True)
