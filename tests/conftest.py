import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
  if _p not in sys.path:
    sys.path.insert(0, _p)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(params=["mma", "cuda"])
def engine(request, monkeypatch):
  """Conditioner engine of the fused kernels (CNFOT_ENGINE): the warp-level tensor-core engine is
  the default for 16-wide networks, the CUDA-core engine covers every other shape; shapes the
  requested engine does not support fall back to the CUDA-core one inside the library."""
  monkeypatch.setenv("CNFOT_ENGINE", request.param)
  return request.param
