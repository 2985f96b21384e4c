"""ctypes front-end of the host-side harness for the device math headers.

TEST ONLY.  Compiles tests/hostsim/*.cpp with g++ on first use.  The product
package never imports this.
"""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _build(name):
  src = os.path.join(HERE, f"hostsim_{name}.cpp")
  out = os.path.join(HERE, f"libhostsim_{name}.so")
  deps = [src] + [
    os.path.join(ROOT, "cnf_ot_b200", "csrc", f)
    for f in ("rqs_math.cuh", "flow_math.cuh", "step_math.cuh", "step_host.h", "philox.cuh")
  ] + [os.path.join(ROOT, "include", "cnfot.h")]
  if (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
    extra = os.environ.get("CNFOT_HOSTSIM_FLAGS", "").split()   # e.g. -DCNFOT_MERGED_SWEEP: the alternative build of the flow sweep
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC"] + extra + ["-x", "c++", src, "-o", out])
  return ctypes.CDLL(out)


_libs = {}


def lib(name):
  if name not in _libs:
    _libs[name] = _build(name)
  return _libs[name]


def _p(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class ProblemDesc(ctypes.Structure):
  _fields_ = [("type", ctypes.c_int32), ("subtype", ctypes.c_int32), ("T", ctypes.c_float),
              ("beta", ctypes.c_float), ("a", ctypes.c_float), ("sigma", ctypes.c_float),
              ("dt", ctypes.c_float), ("dx", ctypes.c_float)]


def rqs(K, direction, v, theta, gout=None, gld=None, dtype=torch.float64):
  n = v.numel()
  P = 3 * K + 1
  v = v.to(dtype).contiguous()
  theta = theta.to(dtype).contiguous()
  out = torch.empty(n, dtype=dtype)
  ld = torch.empty(n, dtype=dtype)
  idx = torch.empty(n, dtype=torch.int32)
  gin = gth = None
  if gout is not None:
    gout = gout.to(dtype).contiguous()
    gld = gld.to(dtype).contiguous()
    gin = torch.empty(n, dtype=dtype)
    gth = torch.empty(n, P, dtype=dtype)
  fn = lib("rqs").hs_rqs_f64 if dtype == torch.float64 else lib("rqs").hs_rqs_f32
  rc = fn(ctypes.c_int64(n), K, direction, _p(v), _p(theta), ctypes.c_int64(P), _p(gout), _p(gld),
          _p(out), _p(ld), _p(idx), _p(gin), _p(gth))
  assert rc == 0, rc
  return out, ld, idx, gin, gth


def flow_eval(shape, direction, W, x, cond, add_base=0, dtype=torch.float64):
  rows, D = x.shape
  W = W.to(dtype).contiguous()
  x = x.to(dtype).contiguous()
  cond = cond.to(dtype).contiguous().reshape(-1)
  cs = 0 if cond.numel() == 1 else 1
  out = torch.empty(rows, D, dtype=dtype)
  ld = torch.empty(rows, dtype=dtype)
  fn = lib("flow").hs_flow_eval_f64 if dtype == torch.float64 else lib("flow").hs_flow_eval_f32
  rc = fn(shape.dim, shape.num_layers, shape.mlp_layers, shape.hidden, shape.num_bins, direction,
          ctypes.c_int64(rows), _p(W), _p(x), _p(cond), ctypes.c_int64(cs), _p(out), _p(ld), add_base)
  assert rc == 0, rc
  return out, ld


def flow_vjp(shape, direction, W, x, cond, gout, gld, add_base=0, dtype=torch.float64):
  rows, D = x.shape
  W = W.to(dtype).contiguous()
  x = x.to(dtype).contiguous()
  cond = cond.to(dtype).contiguous().reshape(-1)
  cs = 0 if cond.numel() == 1 else 1
  gout = gout.to(dtype).contiguous()
  gld = gld.to(dtype).contiguous()
  gin = torch.empty(rows, D, dtype=dtype)
  G = torch.zeros(shape.blob_size, dtype=torch.float64)
  fn = lib("flow").hs_flow_vjp_f64 if dtype == torch.float64 else lib("flow").hs_flow_vjp_f32
  rc = fn(shape.dim, shape.num_layers, shape.mlp_layers, shape.hidden, shape.num_bins, direction,
          ctypes.c_int64(rows), _p(W), _p(x), _p(cond), ctypes.c_int64(cs), _p(gout), _p(gld),
          add_base, _p(gin), _p(G))
  assert rc == 0, rc
  return gin, G


def step(shape, pd, W, latent, latent_sub, src, tgt, t_batch, gB, gb, lam, dtype=torch.float64):
  c = lambda t: None if t is None else t.to(dtype).contiguous()
  W, latent, latent_sub, src, tgt = c(W), c(latent), c(latent_sub), c(src), c(tgt)
  tb = t_batch.to(torch.float64).contiguous()
  G = torch.zeros(shape.blob_size, dtype=torch.float64)
  slots = torch.zeros(8, dtype=torch.float64)
  fn = lib("flow").hs_step_f64 if dtype == torch.float64 else lib("flow").hs_step_f32
  rc = fn(shape.dim, shape.num_layers, shape.mlp_layers, shape.hidden, shape.num_bins,
          ctypes.byref(pd), _p(W), _p(latent), _p(latent_sub), _p(src), _p(tgt), _p(tb),
          tb.numel(), ctypes.c_int64(latent.shape[0]), ctypes.c_int64(latent_sub.shape[0]),
          ctypes.c_int64(gB), ctypes.c_int64(gb), ctypes.c_double(lam), _p(G), _p(slots))
  assert rc == 0, rc
  return G, slots


def philox_words(ctr, key):
  import numpy as np
  c = np.asarray(ctr, dtype=np.uint32)
  k = np.asarray(key, dtype=np.uint32)
  out = np.zeros(4, dtype=np.uint32)
  lib("philox").hs_philox4x32_10(c.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p),
                                 out.ctypes.data_as(ctypes.c_void_p))
  return out


def philox_rows(key, step, source, global_rows, dim, row0=0, rows=None):
  rows = global_rows - row0 if rows is None else rows
  out = torch.empty(rows, dim, dtype=torch.float32)
  lib("philox").hs_philox_rows(ctypes.c_uint64(key), ctypes.c_uint32(step), source, ctypes.c_int64(global_rows),
                               ctypes.c_int64(row0), ctypes.c_int64(rows), dim, _p(out))
  return out


def philox_times(key, step, n_t, horizon):
  out = torch.empty(n_t, dtype=torch.float32)
  lib("philox").hs_philox_times(ctypes.c_uint64(key), ctypes.c_uint32(step), n_t, ctypes.c_float(horizon), _p(out))
  return out
