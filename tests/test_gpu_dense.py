"""GPU parity of the wide conditioner layers on tcgen05 (cnfot_dense_*; BASELINE config 5, hidden 512)
against a float64 matmul: fp32 fidelity through the 3xTF32 split, every epilogue, ragged row counts."""
import pytest
import torch

from cnf_ot_b200 import ops

pytestmark = pytest.mark.gpu


def _ref(X, W, b):
  return X.double() @ W.double() + (0 if b is None else b.double())


@pytest.mark.parametrize("rows,K,N", [(1000, 512, 512), (128, 512, 16), (333, 64, 64), (257, 32, 256), (5, 16, 48),
                                      (70, 512, 128)])
def test_dense_forward_matches_float64(rows, K, N):
  g = torch.Generator().manual_seed(rows + K + N)
  X = torch.randn(rows, K, generator=g).cuda()
  W = (torch.randn(K, N, generator=g) / K**0.5).cuda()
  b = torch.randn(N, generator=g).cuda()
  P = ops.PreparedDense(W)
  ref = _ref(X, W, b)
  scale = float(ref.abs().max())
  y = ops.dense_forward(X, P, bias=b, epilogue="bias")
  # single-pass tf32 would give ~1e-3 * scale; the stated bound is the path's fp32 tolerance 1e-5 (north_star)
  assert float((y.double() - ref).abs().max()) <= 1e-5 * scale
  yr = ops.dense_forward(X, P, bias=b, epilogue="bias_relu")
  assert float((yr.double() - ref.clamp_min(0)).abs().max()) <= 1e-5 * scale
  yn = ops.dense_forward(X, P, epilogue="none")
  assert float((yn.double() - _ref(X, W, None)).abs().max()) <= 1e-5 * scale
  m = torch.randn(rows, N, generator=g).cuda()
  ym = ops.dense_forward(X, P, mask_src=m, epilogue="relu_mask")
  assert float((ym.double() - torch.where(m.double() > 0, _ref(X, W, None), 0.0)).abs().max()) <= 1e-5 * scale


def test_dense_transposed_weights_give_the_data_gradient():
  """prepare(transpose=True) serves dX = G W^T (the conditioner's dgrad)."""
  g = torch.Generator().manual_seed(3)
  G = torch.randn(300, 16, generator=g).cuda()
  W = torch.randn(512, 16, generator=g).cuda()   # (in, out) like haiku's w
  P = ops.PreparedDense(W, transpose=True)        # contraction over `out`
  assert (P.K, P.N) == (16, 512)
  dx = ops.dense_forward(G, P, epilogue="none")
  ref = G.double() @ W.double().T
  assert float((dx.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_dense_rejects_bad_shapes():
  from cnf_ot_b200._lib import CnfotError
  with pytest.raises(CnfotError):
    ops.PreparedDense(torch.zeros(20, 16, device="cuda"))
  P = ops.PreparedDense(torch.zeros(16, 16, device="cuda"))
  with pytest.raises(CnfotError):
    ops.dense_forward(torch.zeros(4, 32, device="cuda"), P, epilogue="none")
  with pytest.raises(CnfotError):
    ops.dense_forward(torch.zeros(4, 16, device="cuda"), P, epilogue="bias")


@pytest.mark.parametrize("rows,Ka,Nb", [(5000, 512, 512), (777, 512, 16), (100, 33, 512), (4096, 64, 128), (17, 16, 64)])
def test_dense_wgrad_matches_float64(rows, Ka, Nb):
  g = torch.Generator().manual_seed(rows + Ka + Nb)
  A = torch.randn(rows, Ka, generator=g).cuda()
  G = torch.randn(rows, Nb, generator=g).cuda()
  dW0 = torch.randn(Ka, Nb, generator=g).cuda()
  db0 = torch.randn(Nb, generator=g).cuda()
  dW, db = dW0.clone(), db0.clone()
  ops.dense_wgrad(A, G, dW, db)
  ref = dW0.double() + A.double().T @ G.double()
  refb = db0.double() + G.double().sum(0)
  assert float((dW.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
  assert float((db.double() - refb).abs().max()) <= 1e-5 * float(refb.abs().max())
  # a strided destination (a matrix inside a larger blob) and inputs that are column slices
  blob = torch.zeros(Ka, Nb + 16, device="cuda")
  ops.dense_wgrad(A, G, blob[:, 16:])
  assert float((blob[:, 16:].double() - A.double().T @ G.double()).abs().max()) <= 1e-5 * float(ref.abs().max())
  assert float(blob[:, :16].abs().max()) == 0.0
