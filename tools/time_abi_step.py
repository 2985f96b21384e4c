"""Time cnfot_mfc_step of ANY build of the library through raw ctypes (same signature since round 1), on the same
inputs: A/B of two builds inside one gpurun call (box-to-box variance is several percent).
  python tools/time_abi_step.py <lib.so> [cfg2|cfg3]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnf_ot_b200 import _lib

path, name = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "cfg2")
lib = ctypes.CDLL(os.path.abspath(path))
c = ctypes
F, P = c.POINTER(_lib.FlowDesc), c.POINTER(_lib.ProblemDesc)
lib.cnfot_mfc_step.restype = c.c_int32
lib.cnfot_mfc_step.argtypes = [c.c_void_p, F, P, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int32,
                               c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_float, c.c_void_p, c.c_void_p, c.c_int64]
lib.cnfot_mfc_step_workspace_bytes.restype = c.c_int64
lib.cnfot_mfc_step_workspace_bytes.argtypes = [F, c.c_int64, c.c_int64, c.c_int32]


class _D:
  world, rank, local = 1, 0, 0
  dev = torch.device("cuda", 0)
  td = None


torch.cuda.set_device(0)
w = bench.Workload(name, _D())
desc = _lib.flow_desc(w.shape)
ws = torch.empty(lib.cnfot_mfc_step_workspace_bytes(desc, w.B, w.b, 1), dtype=torch.uint8, device="cuda")
p = lambda t: 0 if t is None else t.data_ptr()
tb = torch.zeros(1)


def step(i):
  s = w.sets[i % w.n_sets]
  tb[0] = w.t_vals[i % 4096]
  rc = lib.cnfot_mfc_step(torch.cuda.current_stream().cuda_stream, desc, w.problem, p(w.W), p(s.get("latent")), p(s["latent_sub"]),
                          p(s.get("src")), p(s.get("tgt")), tb.data_ptr(), 1, w.B, w.b, w.gB, w.gb, w.lam, p(w.out), ws.data_ptr(),
                          ws.numel())
  assert rc == 0, rc


n = 20 if name == "cfg2" else 10
for i in range(3):
  step(i)
ts = sorted(bench.time_region(step, n, torch.cuda.synchronize, first=r * n) / n * 1e3 for r in range(7))
print(f"{os.path.basename(path)} {name}: {ts[3]:.4f} ms/step (min {ts[0]:.4f}, max {ts[-1]:.4f}); loss {float(w.out[w.shape.blob_size]):.6e}")
