import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import pack, FlowShape
from oracle import flow as oflow
from util import make_cfg, make_params, shape_of
for (D,L,M,H,K,sigma) in [(2,2,2,16,5,0.3),(3,3,1,8,3,0.1)]:
    cfg = make_cfg(dim=D,L=L,M=M,H=H,K=K); shape = shape_of(cfg); spec, params = make_params(cfg, sigma)
    W = pack(shape, params).cuda()
    g = torch.Generator().manual_seed(5); n = 300
    x = torch.randn(n, D, generator=g, dtype=torch.float64).float()
    cond = torch.rand(n, generator=g, dtype=torch.float64).float()
    gout = torch.randn(n, D, generator=g).float(); gld = torch.randn(n, generator=g).float()
    for inverse in (False, True):
        xx = x.double().requires_grad_(True); p = oflow.clone_params(params, True)
        fn = oflow.flow_inverse_and_log_det if inverse else oflow.flow_forward_and_log_det
        o, l = fn(spec, p, xx, cond.double().reshape(-1,1))
        ((o*gout.double()).sum() + (l*gld.double()).sum()).backward()
        grads = {m: {k: v.grad for k, v in lv.items()} for m, lv in p.items()}
        Gor = pack(shape, grads, torch.float64)
        gin, G = ops.flow_vjp(shape, W, x.cuda(), cond.cuda(), gout.cuda(), gld.cuda(), inverse=inverse)
        G = G.cpu().double()
        print(f"D{D} H{H} inv{int(inverse)}: gin err {float((gin.cpu().double()-xx.grad).abs().max()):.2e}  G relmax {float((G-Gor).abs().max()/Gor.abs().max()):.2e}")
        for mod, leaf, shp, off, stride in shape.leaves():
            rows = 1
            for s_ in shp[:-1]: rows *= s_
            a = torch.stack([G[off+r*stride: off+r*stride+shp[-1]] for r in range(rows)])
            b = torch.stack([Gor[off+r*stride: off+r*stride+shp[-1]] for r in range(rows)])
            e = float((a-b).abs().max()/(b.abs().max()+1e-12))
            if e > 1e-4: print(f"    {mod}/{leaf} {shp}: rel err {e:.2e}  ours[0,:4]={a[0,:4].tolist()} ref={b[0,:4].tolist()}")
