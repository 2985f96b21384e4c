import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cnf_ot_b200 import ops, _lib
from cnf_ot_b200.layout import pack
from oracle import losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of
from test_gpu_step import run_gpu
for typ, sub, M in [("rwpo","quadratic",1),("rwpo","quadratic",2),("rwpo","double_well",1),("ot","obstacle",1),("fp","nongradient",1),("fp","nongradient",2)]:
    cfg = make_cfg(typ, sub, Tn=2, lam=500.0, B=1088, M=M); shape = shape_of(cfg)
    spec, params = make_params(cfg, 0.3); inputs = make_inputs(cfg)
    loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
    Gor = pack(shape, grads, torch.float64)
    for tc in ("0","1"):
        os.environ["CNFOT_TC"] = tc
        out = run_gpu(cfg, shape, params, inputs, 500.0)
        G, slots = out[:shape.blob_size], out[shape.blob_size:]
        print(f"{typ}/{sub} M{M} tc={tc} ({_lib.last_launch_info()['tensor_cores']}): loss rel {abs(float(slots[0])-float(loss))/abs(float(loss)):.2e} grad relmax {float((G-Gor).abs().max()/Gor.abs().max()):.2e} slots {[round(float(x),3) for x in slots[1:5]]}")
