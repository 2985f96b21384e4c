"""cuBLAS TF32 / BF16 GEMM throughput on this GPU (reference points for the tcgen05 kind::tf32 kernels)."""
import torch
def bench(dtype, tf32, n=8192, reps=20):
  torch.backends.cuda.matmul.allow_tf32 = tf32
  a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
  for _ in range(3): a @ b
  torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): a @ b
  e1.record(); torch.cuda.synchronize()
  return 2.0 * n**3 * reps / e0.elapsed_time(e1) / 1e9
print("bf16 8192^3: %.0f TFLOP/s" % bench(torch.bfloat16, False))
print("tf32 8192^3: %.0f TFLOP/s" % bench(torch.float32, True))
print("fp32 8192^3 (no tf32): %.0f TFLOP/s" % bench(torch.float32, False, reps=3))
# shape of the conditioner layer: (rows x 512) x (512 x 512)
torch.backends.cuda.matmul.allow_tf32 = True
a = torch.randn(1 << 20, 512, device="cuda"); b = torch.randn(512, 512, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): a @ b
e1.record(); torch.cuda.synchronize()
print("tf32 2^20 x 512 x 512: %.0f TFLOP/s (single pass tf32)" % (2.0 * (1 << 20) * 512 * 512 * 10 / e0.elapsed_time(e1) / 1e9))
