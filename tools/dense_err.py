"""Error of the tcgen05 dense layer vs float64, next to torch's fp32 matmul (cuBLAS SIMT, TF32 off)."""
import torch
from cnf_ot_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
for rows, K, N in [(4096, 512, 512), (4096, 64, 64), (4096, 512, 16)]:
  g = torch.Generator().manual_seed(1)
  X = torch.randn(rows, K, generator=g).clamp_min(0).cuda()   # post-ReLU like activations
  W = (torch.randn(K, N, generator=g) / K**0.5).cuda()
  ref = X.double() @ W.double()
  y = ops.dense_forward(X, ops.PreparedDense(W), epilogue="none")
  t = X @ W
  sc = float(ref.abs().max())
  rms = float(ref.pow(2).mean().sqrt())
  print(rows, K, N, "ours max %.2e rms %.2e | torch fp32 max %.2e rms %.2e  (rel. to max|ref|; rms ref/max %.2f)" % (
    float((y.double() - ref).abs().max()) / sc, float((y.double() - ref).pow(2).mean().sqrt()) / sc,
    float((t.double() - ref).abs().max()) / sc, float((t.double() - ref).pow(2).mean().sqrt()) / sc, rms / sc))
  print("   mean signed err ours %.2e torch %.2e" % (float((y.double() - ref).mean()) / sc, float((t.double() - ref).mean()) / sc))
