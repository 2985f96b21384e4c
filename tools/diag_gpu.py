"""Prints kernel-vs-oracle error metrics and quick timings (diagnostics, not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cnf_ot_b200 import ops
from cnf_ot_b200.layout import pack
from oracle import rqs, flow as oflow, losses as olosses
from util import make_cfg, make_inputs, make_params, shape_of, rel_err

def q(a, b):
    a = a.detach().cpu().double().reshape(-1); b = b.detach().cpu().double().reshape(-1)
    e = (a - b).abs() / (b.abs() + 1)
    return f"max {float(e.max()):.2e} p99.9 {float(e.quantile(0.999)):.2e} med {float(e.median()):.2e}"

print("== rqs")
for K in (5, 8, 16):
  for inverse in (False, True):
    g = torch.Generator().manual_seed(K + 100 * inverse)
    n = 20000
    theta = (torch.randn(n, 3*K+1, generator=g, dtype=torch.float64) * 0.5).float()
    v = (torch.randn(n, generator=g, dtype=torch.float64) * 5).float()
    gout = torch.randn(n, generator=g).float(); gld = torch.randn(n, generator=g).float()
    f = rqs.rqs_inverse if inverse else rqs.rqs_forward
    vv, th = v.double().requires_grad_(True), theta.double().requires_grad_(True)
    o, l, idx = f(vv, th); (o*gout.double() + l*gld.double()).sum().backward()
    fn = ops.rqs_inverse if inverse else ops.rqs_forward
    out, ld, bins = fn(v.cuda(), theta.cuda(), K, want_bins=True)
    gin, gth = ops.rqs_vjp(inverse, v.cuda(), theta.cuda(), gout.cuda(), gld.cuda(), K)
    print(f"K{K} inv{int(inverse)}: out {q(out,o)} | ld {q(ld,l)} | bins!= {int((bins.cpu().long()!=idx).sum())} | gin {q(gin,vv.grad)} | gth {q(gth,th.grad)}")

print("== flow")
for (D,L,M,H,K,sigma) in [(2,2,2,16,5,0.3),(3,3,1,8,3,0.3),(10,2,2,16,5,0.05),(4,3,2,32,8,0.2),(2,4,1,16,5,0.3)]:
    cfg = make_cfg(dim=D,L=L,M=M,H=H,K=K); shape = shape_of(cfg); spec, params = make_params(cfg, sigma)
    W = pack(shape, params).cuda()
    g = torch.Generator().manual_seed(5); n = 4000
    x = (torch.randn(n, D, generator=g, dtype=torch.float64)*1.5).float()
    cond = torch.rand(n, generator=g, dtype=torch.float64).float()
    gout = torch.randn(n, D, generator=g).float(); gld = torch.randn(n, generator=g).float()
    for inverse in (False, True):
        xx = x.double().requires_grad_(True); p = oflow.clone_params(params, True)
        fn = oflow.flow_inverse_and_log_det if inverse else oflow.flow_forward_and_log_det
        o, l = fn(spec, p, xx, cond.double().reshape(-1,1))
        ((o*gout.double()).sum() + (l*gld.double()).sum()).backward()
        Gor = pack(shape, {m: {k: v.grad for k, v in lv.items()} for m, lv in p.items()}, torch.float64)
        y, ld = ops.flow_eval(shape, W, x.cuda(), cond.cuda(), inverse=inverse)
        gin, G = ops.flow_vjp(shape, W, x.cuda(), cond.cuda(), gout.cuda(), gld.cuda(), inverse=inverse)
        back, _ = ops.flow_eval(shape, W, y, cond.cuda(), inverse=not inverse)
        print(f"D{D} L{L} M{M} H{H} K{K} inv{int(inverse)}: out {q(y,o)} | ld {q(ld,l)} (|ld|max {float(l.abs().max()):.1f}) | gin {q(gin,xx.grad)} | G relmax {float((G.cpu().double()-Gor).abs().max()/Gor.abs().max()):.2e} | roundtrip {q(back, x)}")

print("== step timing")
for typ, sub, B, D, sigma in [("ot","obstacle",1<<18,2,0.3), ("rwpo","double_well",1<<20,2,0.3), ("fp","nongradient",1<<19,10,0.05), ("ot","free",4096,2,0.3)]:
    cfg = make_cfg(typ, sub, dim=D, B=B); shape = shape_of(cfg); spec, params = make_params(cfg, sigma)
    if typ == "rwpo": cfg["rwpo"].update(T=1, beta=1, a=1)
    W = pack(shape, params).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    lat = torch.randn(B, D, device="cuda", generator=g); sub_ = torch.randn(B//32, D, device="cuda", generator=g)
    src = lat + 3.0; tgt = torch.randn(B, D, device="cuda", generator=g)
    pd = ops.problem_desc(cfg)
    args = (shape, pd, W, None if typ=="ot" else lat, sub_, src if typ=="ot" else None, tgt if typ=="ot" else None, [0.37], 5000.0, B, B//32)
    out = ops.mfc_step(*args)
    for _ in range(3): ops.mfc_step(*args, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.mfc_step(*args, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    from cnf_ot_b200 import _lib
    print(f"{typ}/{sub} D{D} B={B}: {ms*1000:.1f} us/step -> {B/ms*1000/1e6:.1f} M samples/s; loss {float(out[shape.blob_size]):.4e} {_lib.last_launch_info()}")

print("== rqs kernel bandwidth (K=5)")
n = 1 << 24
theta = torch.randn(n, 16, device="cuda") * 0.3; v = torch.randn(n, device="cuda") * 3
for name, fn, nbytes in [("fwd", lambda: ops.rqs_forward(v, theta, 5), 76), ("inv", lambda: ops.rqs_inverse(v, theta, 5), 76)]:
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
    print(f"rqs {name}: {ms*1000:.1f} us for 2^24 rows -> {n*nbytes/ms/1e6:.0f} GB/s (incl. torch.empty)")
go = torch.randn(n, device="cuda"); gl = torch.randn(n, device="cuda")
for inv in (False, True):
    for _ in range(3): ops.rqs_vjp(inv, v, theta, go, gl, 5)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.rqs_vjp(inv, v, theta, go, gl, 5)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
    print(f"rqs vjp inv{int(inv)}: {ms*1000:.1f} us -> {n*144/ms/1e6:.0f} GB/s")
