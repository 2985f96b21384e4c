#!/usr/bin/env python
"""Benchmark of the cnf_ot flow train step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one evaluation of the configured MFC loss and its parameter gradient
(all flow passes forward + inverse + log-det + loss terms + backward; optimiser
excluded) over one synthetic batch.  Workload = BASELINE.json
configs[1]: mfc.yaml type=ot subtype=obstacle, 2-D, RQS flow (2 layers, 2x16
conditioner, 5 bins), batch 2^18 per GPU (weak scaling), lambda=5000, dt=0.01.

  value  samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e    the same through the C-ABI call with HOST (pinned) buffers: H2D of the batch
         and the weights, the step, D2H of [gradient | loss] inside the timed region
  --impl reference : the CPU restatement of the reference step (oracle/, torch f64,
         all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
  if _p not in sys.path:
    sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "flow train-step samples/s (fwd+inv+logdet+loss+grad)"
UNIT = "samples/s"
B_PER_GPU = 1 << 18
SIGMA = 0.3  # parameter perturbation


# stdout carries exactly ONE line (the JSON): everything a library writes to file descriptor 1 (NCCL prints its
# version banner there) is sent to stderr, the JSON goes to a private copy of the original stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
  _JSON_OUT.write(json.dumps(line) + "\n")
  _JSON_OUT.flush()


def workload_cfg(batch):
  return {
    "general": {"type": "ot", "dim": 2, "dx": 0.01, "dt": 0.01, "t_batch_size": 1, "seed": 42},
    "ot": {"subtype": "obstacle"},
    "rwpo": {"T": 1, "beta": 1, "a": 1, "pot_type": "double_well"},
    "fp": {"T": 1, "a": 1, "sigma": 0.5, "velocity_field_type": "nongradient"},
    "cnf": {"flow_num_layers": 2, "mlp_num_layers": 2, "hidden_size": 16, "num_bins": 5},
    "train": {"epochs": 1, "lr": 1e-3, "_lambda": 5000.0, "batch_size": batch, "eval_frequency": 100},
  }


def config_block(n_gpus, extra=None):
  c = {
    "workload": "mfc.yaml type=ot subtype=obstacle (BASELINE configs[1])",
    "dim": 2, "flow_num_layers": 2, "mlp": "2x16", "num_bins": 5, "params": 1200,
    "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * n_gpus, "sub_batch": "batch//32",
    "t_batch_size": 1, "lambda": 5000.0, "param_sigma": SIGMA,
    "rows_through_flow_per_step": "2B log-prob dir + 3(B/32) sample dir",
    "parallelism": f"dp{n_gpus} (rows sharded, one all-reduce of [grad|loss])",
  }
  if extra:
    c.update(extra)
  return c


# ---------------------------------------------------------------- clocks (NVML, in-process)
class ClockSampler:
  def __init__(self, index):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._t = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:  # no NVML: report nulls
      self.nv = None

  def _loop(self):
    nv = self.nv
    names = {
      getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
      getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
      getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
      getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
    }
    while not self._stop.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if r & bit:
            self.reasons.add(name)
      except Exception:
        pass
      time.sleep(0.002)

  def __enter__(self):
    if self.nv:
      self._t = threading.Thread(target=self._loop, daemon=True)
      self._t.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    if self._t:
      self._t.join()

  def summary(self):
    s = sorted(self.samples)
    return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------- CPU reference arm
def oracle_step_timer(rows):
  """One reference train step on the host: value_and_grad of the CPU restatement."""
  from oracle import losses as olosses
  from util import make_inputs, make_params
  cfg = workload_cfg(rows)
  spec, params = make_params(cfg, SIGMA)
  inputs = make_inputs(cfg)

  def step():
    t0 = time.perf_counter()
    loss, _ = olosses.value_and_grad(cfg, spec, params, inputs)
    return time.perf_counter() - t0, float(loss)

  return step


def cpu_baseline(budget_s=15.0):
  torch.set_num_threads(os.cpu_count() or 1)
  probe = oracle_step_timer(1 << 13)
  probe()
  dt, _ = probe()
  rate = (1 << 13) / dt
  rows = 1 << 13
  while rows < B_PER_GPU and (rows * 2) / rate * 3 < budget_s:
    rows *= 2
  step = oracle_step_timer(rows)
  step()
  ts = [step()[0] for _ in range(2)]
  best = min(ts)
  return {"value": rows / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
          "sample": f"oracle (torch f64 restatement of the reference step) on {rows} rows of the "
                    f"same workload, best of 2 after 1 warm-up"}


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  torch.set_num_threads(os.cpu_count() or 1)
  if args.workload == "cfg5":
    # BASELINE configs[4] on the CPU restatement: one bounded sample per step (128 rows of the same flow and loss)
    cb = cpu_baseline_cfg5(args.layers)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 1, "ms_per_step": 128 / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic scale-out, ot/free structure (BASELINE configs[4])", "dim": 32,
                       "flow_num_layers": args.layers, "mlp": "2x512", "rows_per_step_cpu": 128},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return
  # bounded sample per step so K + W steps end within minutes
  probe = oracle_step_timer(1 << 12)
  probe()
  dt, _ = probe()
  rate = (1 << 12) / dt
  total = args.steps + args.warmup
  rows = 1 << 12
  while rows < B_PER_GPU and (rows * 2) / rate * total < 120.0:
    rows *= 2
  step = oracle_step_timer(rows)
  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step()
  el = time.perf_counter() - t0
  value = rows * args.steps / el
  sample = (f"reference step restated on CPU (oracle/, torch f64, autograd), {rows} rows per step "
            f"(bounded sample of the 2^18-row workload), {torch.get_num_threads()} threads")
  line = {
    "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
    "data": "synthetic", "config": config_block(args.gpus, {"rows_per_step_cpu": rows}),
    "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                     "sample": sample},
    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  emit(line)


# ---------------------------------------------------------------- our arm
def make_blob(shape, device):
  """Reference init + N(0, sigma^2) on biases, output layers and `first` (seed 43)."""
  from cnf_ot_b200 import random as crandom
  from cnf_ot_b200.flows import FlowModel
  model = FlowModel(shape, device)
  params = model.init(crandom.PRNGKey(3))
  g = torch.Generator(device="cpu").manual_seed(43)
  for mod, leaves in params.items():
    for name, v in leaves.items():
      if name == "w" and mod.startswith("mlp_"):
        continue
      v.add_((torch.randn(v.shape, generator=g) * SIGMA).to(device))
  return params.blob


def peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f), "measured (MEASURED_PEAKS.json)"
  return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def time_region(fn, n, stream_sync):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  stream_sync()
  e0.record()
  for i in range(n):
    fn(i)
  e1.record()
  stream_sync()
  return e0.elapsed_time(e1) / 1e3  # seconds


def spline_rooflines(peak_gbs):
  """Stand-alone spline kernels (seam 2), HBM-bound: algorithmic bytes 4P+12 / 8P+16 per row."""
  from cnf_ot_b200 import _lib, ops
  lib = _lib.load()
  n, K, P = 1 << 24, 5, 16
  theta = torch.randn(n, P, device="cuda") * SIGMA
  v = torch.randn(n, device="cuda") * 3
  y, ld = torch.empty_like(v), torch.empty_like(v)
  go, gl = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
  gi, gp = torch.empty_like(v), torch.empty_like(theta)
  s = torch.cuda.current_stream().cuda_stream
  calls = {
    "rqs_forward": (lambda i: lib.cnfot_rqs_forward(s, v.data_ptr(), theta.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, y.data_ptr(), ld.data_ptr(), 0), 4 * P + 12),
    "rqs_inverse": (lambda i: lib.cnfot_rqs_inverse(s, v.data_ptr(), theta.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, y.data_ptr(), ld.data_ptr(), 0), 4 * P + 12),
    "rqs_forward_vjp": (lambda i: lib.cnfot_rqs_forward_vjp(s, v.data_ptr(), theta.data_ptr(), go.data_ptr(), gl.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, gi.data_ptr(), gp.data_ptr()), 8 * P + 16),
    "rqs_inverse_vjp": (lambda i: lib.cnfot_rqs_inverse_vjp(s, v.data_ptr(), theta.data_ptr(), go.data_ptr(), gl.data_ptr(), n, K, -10., 10., 1e-4, 1e-4, gi.data_ptr(), gp.data_ptr()), 8 * P + 16),
  }
  out = {}
  for name, (fn, bpr) in calls.items():
    for i in range(3):
      fn(i)
    el = time_region(fn, 10, torch.cuda.synchronize) / 10
    gbs = n * bpr / el / 1e9
    out[name] = {"bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s",
                 "frac": gbs / peak_gbs, "rows": n, "bytes_per_row": bpr, "us": el * 1e6}
  return out


def run_ours(args):
  import torch.distributed as td
  from cnf_ot_b200 import _lib, ops
  from cnf_ot_b200.layout import FlowShape

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: cnf_ot_b200 has no CPU path")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    # keep stdout to the ONE JSON line: NCCL's version banner goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    td.init_process_group("nccl", device_id=dev)
  _lib.load()

  n_gpus = world
  B, b = B_PER_GPU, B_PER_GPU // 32
  gB, gb = B * n_gpus, b * n_gpus
  cfg = workload_cfg(gB)
  shape = FlowShape(2, 2, 2, 16, 5)
  problem = ops.problem_desc(cfg)
  lam = 5000.0
  W = make_blob(shape, dev)
  if world > 1:
    td.broadcast(W, 0)

  # rotating input sets, > 2x L2 in aggregate, so every step streams its batch from HBM
  bytes_per_set = (2 * B + b) * 2 * 4
  n_sets = max(4, (2 * 126 * 2**20) // bytes_per_set + 1)
  g = torch.Generator(device=dev).manual_seed(42 + rank)
  centres = torch.tensor([[0., 5.], [5., 0.], [0., -5.], [-5., 0.], [3., 4.], [3., -4.], [-3., -4.], [-3., 4.]], device=dev)
  sets = []
  for _ in range(n_sets):
    z = torch.randn(B, 2, device=dev, generator=g)
    src = z + centres[torch.randint(0, 8, (B, ), device=dev, generator=g)]
    sets.append((src, z, torch.randn(b, 2, device=dev, generator=g)))
  tgen = torch.Generator().manual_seed(42)
  t_vals = torch.rand(4096, generator=tgen).tolist()
  out = torch.empty(shape.blob_size + 8, dtype=torch.float32, device=dev)

  # N > 1: the step's final reduction kernel also does the all-reduce of [gradient | loss] over
  # peer-mapped memory (NVLink / NVSwitch; cnfot_mfc_step_dp).  NCCL is the fallback transport.
  px, transport = None, "none"
  if world > 1:
    try:
      from cnf_ot_b200 import dist
      px = dist.PeerExchange(shape, dev)
      transport = "fused in the step's reduction kernel over peer-mapped memory (NVLink/NVSwitch)"
    except Exception as exc:  # symmetric memory unavailable
      print(f"[bench] PeerExchange unavailable ({exc!r}); using NCCL all-reduce", file=sys.stderr)
      transport = "NCCL all-reduce"

  def step(i):
    src, tgt, sub = sets[i % n_sets]
    ops.mfc_step(shape, problem, W, None, sub, src, tgt, [t_vals[i % 4096]], lam, gB, gb, out=out, peers=px)
    if world > 1 and px is None:
      td.all_reduce(out)

  def sync():
    if world > 1:
      td.barrier()
    torch.cuda.synchronize()

  for i in range(max(args.warmup, 3)):
    step(i)
  with ClockSampler(local) as clk:
    el = time_region(step, args.steps, sync)
  t = torch.tensor([el], dtype=torch.float64, device=dev)
  if world > 1:
    td.all_reduce(t, op=td.ReduceOp.MAX)
  el = float(t)
  value = gB * args.steps / el
  loss_dev = float(out[shape.blob_size])

  # ---- e2e: host buffers through the C ABI (N=1) / pinned copies + all-reduce (N>1)
  n_host = 8
  pin = lambda x: x.cpu().contiguous().pin_memory()
  hsets = [(pin(s[0]), pin(s[1]), pin(s[2])) for s in sets[:n_host]]
  hW = pin(W)
  hout = torch.empty(shape.blob_size + 8, dtype=torch.float32).pin_memory()
  dW2 = torch.empty_like(W)

  def step_e2e(i):
    src, tgt, sub = hsets[i % n_host]
    if world == 1:
      ops.mfc_step_host(shape, problem, hW, None, sub, src, tgt, [t_vals[i % 4096]], lam, gB, gb, hout, device=dev)
    else:
      # weights H2D; the pinned row buffers are read in place by the kernel (zero-copy over PCIe)
      dW2.copy_(hW, non_blocking=True)
      ops.mfc_step(shape, problem, dW2, None, sub, src, tgt, [t_vals[i % 4096]], lam, gB, gb, out=out, peers=px)
      if px is None:
        td.all_reduce(out)
      hout.copy_(out, non_blocking=True)
      torch.cuda.synchronize()

  for i in range(3):
    step_e2e(i)
  el_e = time_region(step_e2e, args.steps, sync)
  t = torch.tensor([el_e], dtype=torch.float64, device=dev)
  if world > 1:
    td.all_reduce(t, op=td.ReduceOp.MAX)
  el_e = float(t)
  h2d = (2 * B + b) * 2 * 4 + shape.blob_size * 4
  d2h = (shape.blob_size + 8) * 4

  if rank == 0:
    pk, pk_src = peaks()
    hbm = float(pk["hbm_gbs"])
    # dominant kernel: mfc_step_kernel (one launch per step; finalize + an 8-byte memset ride along)
    alg_bytes = (2 * B + b) * 2 * 4 + shape.blob_size * 4 + (shape.blob_size + 8) * 4
    t_launch = el / args.steps
    ach = alg_bytes / t_launch / 1e9
    # fp32 view of the same kernel: algorithmic FLOPs = conditioner fwd+dgrad+wgrad (3 x 2176/row/pass)
    rows_evals = 2 * B + 3 * b
    flops = rows_evals * 3 * 2176.0
    fp32_peak = 148 * 128 * 2 * float(pk.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    traffic, insts = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
      with open(tpath) as f:
        tj = json.load(f)
      traffic = tj.get("mfc_step_kernel_dram_bytes_per_launch")
      insts = tj.get("mfc_step_kernel_warp_insts_per_launch")
    issue_peak = 148 * 4 * float(pk.get("sm_max_mhz", 1965.0)) * 1e6  # warp instructions / s
    line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
      "warmup": max(args.warmup, 3), "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": config_block(n_gpus, {"all_reduce": transport, "l2_policy": f"{n_sets} rotating input sets ({n_sets * bytes_per_set >> 20} MiB > 2x L2)",
                                      "loss_last_step": loss_dev}),
      "e2e": {"value": gB * args.steps / el_e, "unit": UNIT, "h2d_bytes_per_step": h2d,
              "d2h_bytes_per_step": d2h, "ms_per_step": el_e / args.steps * 1e3,
              "api": "cnfot_mfc_step_host (C ABI; pinned host rows read in place by the kernel over PCIe, "
                     "weights H2D, [grad|loss] D2H)" if world == 1 else
                     "weights H2D + cnfot_mfc_step[_dp] on pinned host rows (zero-copy) + all-reduce + D2H"},
      "gpu_launches": 2 * args.steps,  # mfc_step_kernel + finalize[_allreduce]_kernel per step
      "clocks": clk.summary(),
      "roofline": {"bound": "hbm", "kernel": "mfc_step_kernel", "achieved": ach, "peak": hbm, "unit": "GB/s",
                   "frac": ach / hbm, "traffic": traffic, "peak_source": pk_src,
                   "note": "the fused step kernel is bound by instruction issue, not by HBM or the tensor pipe "
                           "(SURVEY.md §8d; ncu: profiles/): see roofline_issue / roofline_fp32; the HBM-bound kernels "
                           "of the path are the stand-alone spline kernels in roofline_spline"},
      "roofline_issue": {"bound": "issue", "kernel": "mfc_step_kernel",
                         "achieved": (insts / t_launch / 1e9) if insts else None, "peak": issue_peak / 1e9,
                         "unit": "G warp-inst/s", "frac": (insts / t_launch / issue_peak) if insts else None,
                         "warp_insts_per_launch": insts,
                         "note": "executed warp instructions per launch (ncu smsp__inst_executed.sum, profiles/) / "
                                 "live launch time, against 148 SM x 4 schedulers x max clock"},
      "roofline_fp32": {"bound": "fp32", "kernel": "mfc_step_kernel", "achieved": flops / t_launch / 1e12,
                        "peak": fp32_peak, "unit": "TFLOP/s", "frac": flops / t_launch / 1e12 / fp32_peak,
                        "flops_per_launch": flops,
                        "note": "algorithmic conditioner FLOPs (fwd+dgrad+wgrad) only, against the CUDA-core fp32 peak "
                                "148 SM x 128 FMA x 2 x max clock; the 16x16 layers run on the tensor pipe (mma.sync "
                                "tf32, 3 MMAs per product for fp32 fidelity), the input layers and splines on CUDA cores"},
    }
    if world == 1:
      line["roofline_spline"] = spline_rooflines(hbm)
      line["cpu_baseline"] = cpu_baseline()
    emit(line)
  if world > 1:
    td.barrier()
    td.destroy_process_group()


# ---------------------------------------------------------------- BASELINE configs[4] on the wide-conditioner engine
def run_cfg5(args):
  """Synthetic scale-out workload (BASELINE configs[4]): d = 32, 16 layers, conditioner 2 x 512, ot/free structure,
  Gaussian source N(-3, I) -> target N(0, I), rows sharded over the GPUs, ONE NCCL all-reduce of the 556 MB
  [gradient | loss] buffer per step.  Not the driver's default line (that is configs[1]); run with
  --workload cfg5 [--rows-per-gpu R] (BASELINE: 2^24 / 8 = 2^21 rows per GPU)."""
  import torch.distributed as td
  from cnf_ot_b200 import _lib, ops
  from cnf_ot_b200.layout import FlowShape
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    td.init_process_group("nccl", device_id=dev)
  D, L, H = 32, args.layers, 512
  shape = FlowShape(D, L, 2, H, 5)
  B = args.rows_per_gpu
  b = B // 32
  gB, gb = B * world, b * world
  cfg = {"general": {"type": "ot", "dim": D, "dx": 0.01, "dt": 0.01}, "ot": {"subtype": "free"}}
  problem = ops.problem_desc(cfg)
  g = torch.Generator(device=dev).manual_seed(43)
  # haiku-like scale for the hidden matrices (1/sqrt(fan_in) ~ 0.04 at 512), small output layers: a well-conditioned flow
  W = torch.randn(shape.blob_size, device=dev, generator=g) * 0.02
  if world > 1:
    td.broadcast(W, 0)
  g = torch.Generator(device=dev).manual_seed(42 + rank)
  n_sets = 2
  sets = [(torch.randn(B, D, device=dev, generator=g) - 3.0, torch.randn(B, D, device=dev, generator=g),
           torch.randn(b, D, device=dev, generator=g)) for _ in range(n_sets)]
  out = torch.empty(shape.blob_size + 8, dtype=torch.float32, device=dev)
  t_vals = torch.rand(64, generator=torch.Generator().manual_seed(42)).tolist()

  def step(i):
    src, tgt, sub = sets[i % n_sets]
    ops.mfc_step(shape, problem, W, None, sub, src, tgt, [t_vals[i % 64]], 5000.0, gB, gb, out=out)
    if world > 1:
      td.all_reduce(out)

  def sync():
    if world > 1:
      td.barrier()
    torch.cuda.synchronize()

  for i in range(max(args.warmup, 1)):
    step(i)
  with ClockSampler(local) as clk:
    el = time_region(step, args.steps, sync)
  t = torch.tensor([el], dtype=torch.float64, device=dev)
  if world > 1:
    td.all_reduce(t, op=td.ReduceOp.MAX)
  el = float(t)
  # e2e: pinned host rows read in place, weights H2D, [gradient | loss] D2H
  pin = lambda x: x.cpu().contiguous().pin_memory()
  hsrc, htgt, hsub = (pin(x) for x in sets[0])
  hW, hout = pin(W), torch.empty(shape.blob_size + 8, dtype=torch.float32).pin_memory()
  dW2 = torch.empty_like(W)

  def step_e2e(i):
    dW2.copy_(hW, non_blocking=True)
    ops.mfc_step(shape, problem, dW2, None, hsub, hsrc, htgt, [t_vals[i % 64]], 5000.0, gB, gb, out=out)
    if world > 1:
      td.all_reduce(out)
    hout.copy_(out, non_blocking=True)
    torch.cuda.synchronize()

  if args.no_e2e:   # BASELINE-size multi-GPU runs: one 42 s step per timed region is expensive; e2e measured at N = 1
    el_e = float("nan")
  else:
    step_e2e(0)
    el_e = time_region(step_e2e, max(1, args.steps // 2), sync) / max(1, args.steps // 2) * args.steps
  t = torch.tensor([el_e], dtype=torch.float64, device=dev)
  if world > 1:
    td.all_reduce(t, op=td.ReduceOp.MAX)
  el_e = float(t)
  if rank == 0:
    pk, pk_src = peaks()
    # dominant kernel: the hidden-layer GEMM (rows x 512 x 512, 3xTF32), timed alone on one chunk of rows
    rows = 4 * 148 * 128
    X = torch.randn(rows, H, device=dev)
    P = ops.PreparedDense(torch.randn(H, H, device=dev) / H**0.5)
    bias = torch.zeros(H, device=dev)
    Y = torch.empty(rows, H, device=dev)
    fn = lambda i: ops.dense_forward(X, P, bias=bias, epilogue="bias_relu", out=Y)
    for i in range(3):
      fn(i)
    tk = time_region(fn, 20, torch.cuda.synchronize) / 20
    tf32_peak = float(pk.get("bf16_tflops_sustained", 1392.5)) / 2.0
    pipe = 3 * 2.0 * rows * H * H / tk / 1e12   # tf32 MMA flops issued (3 per fp32-fidelity product)
    per_pass = 2 * L * sum((d + 1) * H + H * H + 16 * H for d in range(1, D))
    line = {
      "metric": METRIC, "value": gB * args.steps / el, "unit": UNIT, "n_gpus": world, "steps": args.steps,
      "warmup": max(args.warmup, 1), "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 on tcgen05, fp32 accumulate)", "data": "synthetic",
      "config": {"workload": "synthetic scale-out, ot/free structure (BASELINE configs[4])", "dim": D, "flow_num_layers": L,
                 "mlp": "2x512", "num_bins": 5, "params": shape.blob_size, "batch_per_gpu": B, "global_batch": gB,
                 "sub_batch": "batch//32", "engine": _lib.last_launch_info()["engine"],
                 "parallelism": f"dp{world} (rows sharded, one NCCL all-reduce of [grad|loss], {shape.blob_size * 4 >> 20} MiB)",
                 "l2_policy": "inputs and activations far larger than L2", "loss_last_step": float(out[shape.blob_size])},
      "e2e": {"value": None if args.no_e2e else gB * args.steps / el_e, "unit": UNIT,
              "h2d_bytes_per_step": (2 * B + b) * D * 4 + shape.blob_size * 4,
              "d2h_bytes_per_step": (shape.blob_size + 8) * 4,
              "api": "weights H2D + cnfot_mfc_step on pinned host rows (read in place) + all-reduce + D2H"},
      "gpu_launches": None, "clocks": clk.summary(),
      "roofline": {"bound": "tensor", "kernel": "dense_tc_kernel<256,1> (hidden layer, rows x 512 x 512, timed alone)",
                   "achieved": pipe, "peak": tf32_peak, "unit": "TFLOP/s", "frac": pipe / tf32_peak, "traffic": None,
                   "peak_source": pk_src + ": bf16_tflops_sustained / 2 (tf32)",
                   "fp32_fidelity_tflops": pipe / 3, "us_per_launch": tk * 1e6,
                   "step_algorithmic_tflops": 3 * per_pass * (2 * B + 2 * b) / (el / args.steps) / 1e12},
    }
    if world == 1 and not args.no_cpu_baseline:
      line["cpu_baseline"] = cpu_baseline_cfg5(L)
    emit(line)
  if world > 1:
    td.barrier()
    td.destroy_process_group()


def cpu_baseline_cfg5(L, rows=128):
  """The CPU restatement (oracle/, torch f64, autograd) on a bounded sample of the configs[4] workload."""
  from oracle import losses as olosses
  from util import make_cfg, make_inputs, make_params
  torch.set_num_threads(os.cpu_count() or 1)
  cfg = make_cfg("ot", "free", dim=32, L=L, M=2, H=512, K=5, B=rows, Tn=1, lam=5000.0)
  spec, params = make_params(cfg, 0.01)
  inputs = make_inputs(cfg)
  olosses.value_and_grad(cfg, spec, params, inputs)
  t0 = time.perf_counter()
  olosses.value_and_grad(cfg, spec, params, inputs)
  dt = time.perf_counter() - t0
  return {"value": rows / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
          "sample": f"oracle (torch f64 restatement of the reference step) on {rows} rows of the same workload "
                    f"(dim 32, {L} layers, 2x512), 1 timed step after 1 warm-up"}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"],
                  help="cfg2 = BASELINE configs[1] (the driver's line); cfg5 = configs[4] on the wide-conditioner engine")
  ap.add_argument("--rows-per-gpu", type=int, default=1 << 17, help="cfg5 only (BASELINE: 2^21)")
  ap.add_argument("--layers", type=int, default=16, help="cfg5 only: flow layers (BASELINE: 16)")
  ap.add_argument("--no-cpu-baseline", action="store_true", help="cfg5 only: skip the CPU oracle timing")
  ap.add_argument("--no-e2e", action="store_true", help="cfg5 only: skip the host-buffer end-to-end timing")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  elif args.workload == "cfg5":
    run_cfg5(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
