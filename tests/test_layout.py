"""Host logic: blob layout vs the reference pytree, and vs the C ABI's offsets."""
import re

import pytest
import torch

from cnf_ot_b200 import _lib
from cnf_ot_b200.layout import FlowShape, pack, unpack
from oracle import flow as oflow

SHAPES = [(2, 2, 2, 16, 5), (10, 2, 2, 16, 5), (3, 3, 1, 8, 3), (4, 3, 2, 32, 8), (2, 2, 3, 16, 5)]


@pytest.mark.parametrize("D,L,M,H,K", SHAPES)
def test_pack_unpack_roundtrip(D, L, M, H, K):
  shape = FlowShape(D, L, M, H, K)
  spec = oflow.FlowSpec(D, L, [H] * M, K)
  params = oflow.perturb_params(oflow.init_params(spec, seed=1), 0.3)
  assert shape.param_count() == spec.param_count() == sum(v.numel() for v in oflow.leaves(params))
  b = pack(shape, params, torch.float64)
  assert b.numel() == shape.blob_size
  back = unpack(shape, b, like=params)
  for mod in params:
    for k in params[mod]:
      assert back[mod][k].dtype == params[mod][k].dtype
      assert torch.equal(back[mod][k].double(), params[mod][k].double()), (mod, k)
  # padding entries stay zero and every leaf lands on distinct floats
  marks = torch.zeros(shape.blob_size)
  for _, _, shp, off, stride in shape.leaves():
    rows = 1
    for s in shp[:-1]:
      rows *= s
    for r in range(rows):
      marks[off + r * stride: off + r * stride + shp[-1]] += 1
  assert marks.max() == 1
  assert int(marks.sum()) == shape.param_count()


def test_default_config_has_1200_params():
  assert FlowShape(2, 2, 2, 16, 5).param_count() == 1200  # config/mfc.yaml, SURVEY A.3
  assert FlowShape(10, 2, 2, 16, 5).param_count() == 11824


@pytest.mark.parametrize("D,L,M,H,K", SHAPES)
def test_offsets_match_c_abi(D, L, M, H, K):
  lib = _lib.load()
  shape = FlowShape(D, L, M, H, K)
  desc = _lib.flow_desc(shape)
  assert lib.cnfot_param_count(desc) == shape.blob_size
  assert lib.cnfot_spline_param_stride(desc) == shape.Pp
  assert lib.cnfot_offset_first(desc) == 0
  for l in range(L):
    for d in range(1, D):
      for m in range(M + 1):
        for bias in (0, 1):
          assert lib.cnfot_offset_linear(desc, l, d, m, bias) == shape.linear_offset(l, d, m, bool(bias))


def test_library_exports_every_declared_symbol():
  """include/cnfot.h <-> libcnfot.so <-> the ctypes table (no compute calls)."""
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  hdr = open(os.path.join(root, "include", "cnfot.h")).read()
  declared = set(re.findall(r"CNFOT_API\s+[\w\s\*]+?\b(cnfot_\w+)\s*\(", hdr))
  assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
  lib = _lib.load()
  for name in declared:
    assert hasattr(lib, name), name
  assert lib.cnfot_abi_version() == _lib.ABI_VERSION


def test_unsupported_shapes_fail_loudly():
  lib = _lib.load()
  bad = _lib.flow_desc(FlowShape(2, 2, 2, 24, 5))
  assert lib.cnfot_flow_supported(bad) != 0
  assert b"no fused kernel" in lib.cnfot_last_error()
  assert lib.cnfot_flow_supported(_lib.flow_desc(FlowShape(2, 2, 2, 16, 5))) == 0
  assert lib.cnfot_param_count(_lib.flow_desc(FlowShape(2, 2, 2, 18, 5))) == -1


def test_param_tree_npz_round_trip(tmp_path):
  """ParamTree.save / load: one array per haiku leaf ("<module>/<leaf>"), same values back, shape mismatches refused."""
  import numpy as np
  import pytest
  import torch
  from cnf_ot_b200.flows import ParamTree
  from cnf_ot_b200.layout import FlowShape
  for shape in (FlowShape(3, 2, 2, 16, 5), FlowShape(4, 2, 1, 8, 3, conditional=False)):
    g = torch.Generator().manual_seed(5)
    tree = ParamTree(shape, torch.randn(shape.blob_size, generator=g))
    # the padding slots of the blob are not parameters: compare leaves
    path = str(tmp_path / "params.npz")
    tree.save(path)
    z = np.load(path)
    assert "~/first" in z.files and z["~/first"].shape == (1, 3 * shape.num_bins + 1)
    assert sum(z[k].size for k in z.files if k != "__flow_shape__") == shape.param_count()
    back = ParamTree.load(path)
    assert back.shape == shape
    for mod in tree:
      for leaf in tree[mod]:
        assert torch.equal(back[mod][leaf], tree[mod][leaf]), (mod, leaf)
  bad = dict(np.load(path))
  bad.pop("~/first")
  np.savez(str(tmp_path / "bad.npz"), **bad)
  with pytest.raises(ValueError):
    ParamTree.load(str(tmp_path / "bad.npz"))
