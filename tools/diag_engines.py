"""Loss / gradient error of every engine (CNFOT_ENGINE=cuda|mma|tc) vs the oracle, and step timing (dev tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cnf_ot_b200 import ops, _lib
from cnf_ot_b200.layout import pack, FlowShape
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of
from test_gpu_step import run_gpu
engines = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cuda", "mma"]
cases = [("ot","obstacle",{}),("rwpo","double_well",{}),("fp","nongradient",{}),("ot","free",dict(M=3)),("ot","obstacle",dict(M=1)),
         ("fp","lorenz",dict(dim=3,L=3,sigma=0.1)),("fp","nongradient",dict(dim=6,sigma=0.05,B=320)),("fp","nongradient",dict(dim=10,sigma=0.05,B=320))]
for typ, sub, kw in cases:
    kw = dict(kw); sigma = kw.pop("sigma", 0.3)
    cfg = make_cfg(typ, sub, Tn=2, lam=500.0, **({"B": 1088} | kw)); shape = shape_of(cfg)
    spec, params = make_params(cfg, sigma); inputs = make_inputs(cfg)
    loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
    Gor = pack(shape, grads, torch.float64)
    for eng in engines:
        os.environ["CNFOT_ENGINE"] = eng
        out = run_gpu(cfg, shape, params, inputs, 500.0)
        G, slots = out[:shape.blob_size], out[shape.blob_size:]
        print(f"{typ}/{sub} {kw} eng={eng} (ran {_lib.last_launch_info()}): loss rel {abs(float(slots[0])-float(loss))/abs(float(loss)):.2e} "
              f"grad relmax {float((G-Gor).abs().max()/Gor.abs().max()):.2e}", flush=True)
# timing on the bench workload
import bench
dev = torch.device("cuda", 0)
for name, typ, D, B in [("ot/obstacle", "ot", 2, 1 << 18), ("rwpo/double_well", "rwpo", 2, 1 << 20), ("fp/nongradient", "fp", 2, 1 << 19),
                        ("fp/nongradient D10 (cfg 4 per GPU)", "fp", 10, 1 << 19)]:
    shape = FlowShape(D, 2, 2, 16, 5); cfg = bench.workload_cfg(B); cfg["general"]["type"] = typ; cfg["general"]["dim"] = D
    b = B // 32
    W = bench.make_blob(shape, dev) if D == 2 else torch.randn(shape.blob_size, device=dev) * 0.05
    g = torch.Generator(device=dev).manual_seed(1)
    lat = torch.randn(B, D, device=dev, generator=g); sub = torch.randn(b, D, device=dev, generator=g)
    src = lat + 3.0; tgt = torch.randn(B, D, device=dev, generator=g)
    pd = ops.problem_desc(cfg); out = torch.empty(shape.blob_size + 8, device=dev)
    for eng in engines:
        os.environ["CNFOT_ENGINE"] = eng
        def step():
            ops.mfc_step(shape, pd, W, None if typ == "ot" else lat, sub, src if typ == "ot" else None,
                         tgt if typ == "ot" else None, [0.37], 5000.0, B, b, out=out)
        for _ in range(3): step()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): step()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"{name} B={B} eng={eng}: {us:.1f} us/step -> {B/us:.1f} M samples/s; loss {float(out[shape.blob_size]):.6e} {_lib.last_launch_info()}", flush=True)
