"""Throughput of the tcgen05 dense layer (dev tool): rows x 512 x 512, 3xTF32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cnf_ot_b200 import ops
for rows, K, N in ((1 << 18, 512, 512), (1 << 20, 512, 512), (1 << 20, 512, 16)):
  X = torch.randn(rows, K, device="cuda"); W = torch.randn(K, N, device="cuda") / K**0.5; b = torch.randn(N, device="cuda")
  P = ops.PreparedDense(W); Y = torch.empty(rows, N, device="cuda")
  for _ in range(3): ops.dense_forward(X, P, bias=b, epilogue="bias_relu", out=Y)
  torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10): ops.dense_forward(X, P, bias=b, epilogue="bias_relu", out=Y)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  fl = 2.0 * rows * K * N
  print(f"rows={rows} K={K} N={N}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s algorithmic (x3 tf32 MMA passes = {3*fl/ms/1e9:.1f} TF/s on the pipe); "
        f"HBM {(rows*(K+N)*4)/ms/1e6:.0f} GB/s", flush=True)
  torch.backends.cuda.matmul.allow_tf32 = False
  for _ in range(2): torch.addmm(b, X, W)
  torch.cuda.synchronize(); e0.record()
  for _ in range(5): torch.relu(torch.addmm(b, X, W))
  e1.record(); torch.cuda.synchronize()
  print(f"   torch fp32 addmm+relu (cuBLAS): {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
for rows, Ka, Nb in ((1 << 20, 512, 512),):
  A = torch.randn(rows, Ka, device="cuda"); G = torch.randn(rows, Nb, device="cuda"); dW = torch.zeros(Ka, Nb, device="cuda"); db = torch.zeros(Nb, device="cuda")
  for _ in range(2): ops.dense_wgrad(A, G, dW, db)
  torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
  for _ in range(5): ops.dense_wgrad(A, G, dW, db)
  e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
  print(f"wgrad rows={rows} {Ka}x{Nb}: {ms:.3f} ms  {2.0*rows*Ka*Nb/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)
