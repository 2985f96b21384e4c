"""Generate tests/golden/ref_*.npz by running the REFERENCE'S OWN SOURCE FILES from /root/reference.

  python tests/golden/make_reference_golden.py        (needs /root/reference; run in the build container only)

PROVENANCE: `cnf_ot/models/{flows,autoregressive,conditional}.py` and `cnf_ot/mfc/applications.py` are imported
unmodified from /root/reference and executed on the torch-float64 stand-ins of `tests/golden/refshim.py` (JAX, haiku and
distrax cannot be installed here).  The model is built exactly as `cnf_ot/mfc/solvers.py:41-54,58-88` builds it
(`hk.without_apply_rng(hk.multi_transform(RQSFlow(...)))`, `partial(applications.<type>_loss_fn, model, ...)`), the
loss is the reference's loss function called with `(params, rng, _lambda, batch_size)` as `update` calls it
(`solvers.py:94`) and the gradient is autograd through the reference's own code.  The only restated piece on that
path is `distrax.RationalQuadraticSpline` (third-party, not vendored: delegates to oracle/rqs.py).
These fixtures pin oracle/flow.py and oracle/losses.py (tests/test_reference_golden.py, CPU, 1e-10) and the CUDA path
(same file, -m gpu, float32 tolerances) to the reference's code, not to a restatement of it.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
  if p not in sys.path:
    sys.path.insert(0, p)

import refshim  # noqa: E402

refshim.install()
sys.path.insert(0, "/root/reference")

import haiku as hk  # noqa: E402  (stand-in)
import jax  # noqa: E402  (stand-in)
import jax.numpy as jnp  # noqa: E402
from cnf_ot.mfc import applications  # noqa: E402  (the reference)
from cnf_ot.models.flows import RQSFlow  # noqa: E402  (the reference)

from cnf_ot_b200.layout import pack  # noqa: E402
from util import make_cfg, make_params, shape_of  # noqa: E402

f32 = lambda t: t.to(torch.float32).to(torch.float64)
CENTRES = torch.tensor([[0.0, 5.0], [5.0, 0.0], [0.0, -5.0], [-5.0, 0.0], [3.0, 4.0], [3.0, -4.0], [-3.0, -4.0],
                        [-3.0, 4.0]], dtype=torch.float64)


def build_model(cfg):
  """solvers.py:41-48."""
  c = cfg["cnf"]
  model = RQSFlow(event_shape=(cfg["general"]["dim"], ), num_layers=c["flow_num_layers"],
                  hidden_sizes=[c["hidden_size"]] * c["mlp_num_layers"], num_bins=c["num_bins"], periodized=False)
  return hk.without_apply_rng(hk.multi_transform(model))


def check_tree(model, cfg, params):
  """model.init (solvers.py:54) creates exactly the leaves (names, shapes) of our parameter tree."""
  dim = cfg["general"]["dim"]
  init = model.init(jax.random.PRNGKey(0), jnp.zeros((1, dim)), jnp.zeros((1, )))
  assert set(init) == set(params), set(init) ^ set(params)
  for mod in init:
    assert set(init[mod]) == set(params[mod]), mod
    for k in init[mod]:
      assert tuple(init[mod][k].shape) == tuple(params[mod][k].shape), (mod, k)
  assert init["~"]["first"].dtype == torch.float32 and float(init["~"]["first"].abs().max()) == 0.0
  return sum(v.numel() for lv in init.values() for v in lv.values())


def loss_fn_of(cfg, model):
  """solvers.py:58-88."""
  g = cfg["general"]
  dim, dt, dx, tbs = g["dim"], g["dt"], g["dx"], g["t_batch_size"]
  if g["type"] == "rwpo":
    r = cfg["rwpo"]
    return partial(applications.rwpo_loss_fn, model, dim, r["T"], r["beta"], dt, dx, tbs, r["pot_type"], r["a"]), r["T"]
  if g["type"] == "fp":
    f = cfg["fp"]
    return partial(applications.fp_loss_fn, model, dim, f["T"], f["a"], f["sigma"], dt, dx, tbs,
                   f["velocity_field_type"]), f["T"]
  return partial(applications.ot_loss_fn, model, dim, 1, dt, tbs, cfg["ot"]["subtype"]), 1


def grad_tree(params):
  return {m: {k: v.clone().requires_grad_(True) for k, v in lv.items()} for m, lv in params.items()}


def step_case(typ, sub, B, lam, Tn, seed, sigma=0.3, **kw):
  cfg = make_cfg(typ, sub, Tn=Tn, lam=lam, B=B, **kw)
  shape = shape_of(cfg)
  D = cfg["general"]["dim"]
  _, params = make_params(cfg, sigma)
  model = build_model(cfg)
  n_params = check_tree(model, cfg, params)
  g = torch.Generator().manual_seed(seed)
  z = f32(torch.randn(B, D, generator=g, dtype=torch.float64))
  u = f32(torch.rand(Tn, generator=g, dtype=torch.float64))
  idx = torch.randint(0, 8, (B, ), generator=g)
  refshim.Draws.set(normal=z, uniform=u, choice=idx)
  loss_fn, T = loss_fn_of(cfg, model)
  p = grad_tree(params)
  loss = loss_fn(p, jax.random.PRNGKey(cfg["general"]["seed"]), lam, B)
  loss.backward()
  grads = {m: {k: (v.grad if v.grad is not None else torch.zeros_like(v)).double() for k, v in lv.items()}
           for m, lv in p.items()}
  out = {"shape": np.array([D, shape.num_layers, shape.mlp_layers, shape.hidden, shape.num_bins]),
         "blob": pack(shape, params, torch.float64), "latent": z, "t_batch": u * T, "lam": np.array(lam),
         "loss": loss.detach(), "grad": pack(shape, grads, torch.float64), "n_params": np.array(n_params)}
  if typ == "ot":   # kl_loss_fn's draws (applications.py:34-82): the mixture noise and the target share the key
    out["src"], out["tgt"] = z + CENTRES[idx], z.clone()
  else:
    out["src"], out["tgt"] = torch.zeros(0, D, dtype=torch.float64), torch.zeros(0, D, dtype=torch.float64)
  print("  draws:", sorted(set(refshim.Draws.log)))
  return out


def flow_case(D, L, M, H, K, sigma, n, seed):
  """model.apply.{sample_and_log_prob, log_prob, forward, inverse} (flows.py:213-224)."""
  cfg = make_cfg(dim=D, L=L, M=M, H=H, K=K)
  shape = shape_of(cfg)
  _, params = make_params(cfg, sigma)
  model = build_model(cfg)
  check_tree(model, cfg, params)
  g = torch.Generator().manual_seed(seed)
  z = f32(torch.randn(n, D, generator=g, dtype=torch.float64))
  t = f32(torch.rand(n, generator=g, dtype=torch.float64))
  t0 = float(t[0])
  refshim.Draws.set(normal=z)
  key = jax.random.PRNGKey(0)
  with torch.no_grad():
    y, lp = model.apply.sample_and_log_prob(params, cond=t.reshape(-1, 1), seed=key, sample_shape=(n, ))
    ys = model.apply.sample(params, cond=t.reshape(-1, 1), seed=key, sample_shape=(n, ))
    x = f32(z * 1.5)
    c0 = jnp.ones((1, )) * t0
    out = {"shape": np.array([D, L, M, H, K]), "blob": pack(shape, params, torch.float64), "latent": z, "t": t,
           "sample_y": y, "sample_log_prob": lp, "x": x, "t0": np.array(t0),
           "log_prob_x": model.apply.log_prob(params, x, cond=c0), "forward_x": model.apply.forward(params, x, c0),
           "inverse_x": model.apply.inverse(params, x, c0)}
  assert torch.equal(y, ys)
  with torch.no_grad():   # applications.py:91-126 on the same latent rows
    out["ot_reverse_kl"] = applications.ot_reverse_kl_loss_fn(model, D, 1, params, key, n)
  return out


def energy_case(D, sigma, n, t_size, beta, T, seed):
  """utils.calc_kinetic_energy / calc_score_kinetic_energy (cnf_ot/utils.py:311-389) as solvers.py:138-160 calls them."""
  refshim.stub_plotting()
  from cnf_ot import utils as ref_utils  # the reference
  cfg = make_cfg(dim=D)
  shape = shape_of(cfg)
  _, params = make_params(cfg, sigma)
  model = build_model(cfg)
  g = torch.Generator().manual_seed(seed)
  z = f32(torch.randn(n, D, generator=g, dtype=torch.float64))
  refshim.Draws.set(normal=z)
  with torch.no_grad():
    e_kin = ref_utils.calc_kinetic_energy(model.apply.sample, params, jax.random.PRNGKey(1), batch_size=n, t_size=t_size,
                                          dim=D)
    e_score = ref_utils.calc_score_kinetic_energy(model.apply.sample, model.apply.log_prob, params, T=T, beta=beta, dim=D,
                                                  rng=jax.random.PRNGKey(1), batch_size=n, t_size=t_size)
  return {"shape": np.array([D, shape.num_layers, shape.mlp_layers, shape.hidden, shape.num_bins]),
          "blob": pack(shape, params, torch.float64), "latent": z, "t_size": np.array(t_size), "T": np.array(T),
          "beta": np.array(beta), "e_kin": e_kin, "e_score": e_score}


def density_case(sigma, n_mc, grid_size, seed):
  """SURVEY.md 8f row 3.  `utils.plot_density_snapshot` and `utils.plot_density_and_trajectory` (cnf_ot/utils.py:572-642)
  run UNMODIFIED against a recording matplotlib stand-in: the fixture holds the arrays they hand to `imshow` (100 x 100
  densities) and `scatter` (trajectories).  `rmse_mc_loss_fn` / `rmse_grid_loss_fn` are nested in `solvers.main`
  (solvers.py:254-301) and cannot be imported: they are restated here line by line on the reference's model API."""
  refshim.record_plots()
  from cnf_ot import utils as ref_utils  # the reference
  import math
  cfg = make_cfg(dim=2)
  shape = shape_of(cfg)
  _, params = make_params(cfg, sigma)
  model = build_model(cfg)
  g = torch.Generator().manual_seed(seed)
  r_ = f32(torch.randn(12, 2, generator=g, dtype=torch.float64) * 1.5)
  t_traj = np.linspace(0, 1, 5)
  dom = [-4.0, 5.0, -3.0, 6.0]
  with torch.no_grad():
    ref_utils.plot_density_snapshot(model.apply.log_prob, params)            # default t_array = linspace(0, 1, 10)
    snap = torch.stack(refshim.Plots.images)
    refshim.Plots.images, refshim.Plots.scatters = [], []
    ref_utils.plot_density_and_trajectory(model.apply.forward, model.apply.inverse, model.apply.log_prob, params, r_,
                                          t_traj, dom)
    dens2 = torch.stack(refshim.Plots.images)
    # every axes gets the same len(t) scatters: pixel coordinates of forward_fn(params, xi, t) (utils.py:626-633)
    sc = refshim.Plots.scatters[:len(t_traj)]
    traj = torch.stack([torch.stack([x * (dom[1] - dom[0]) / 100 + dom[0], y * (dom[3] - dom[2]) / 100 + dom[2]], dim=1)
                        for x, y in sc])
    # solvers.py:238-252 (dim = 2, a, T as in config/mfc.yaml fp block)
    a, T, cond = 1.0, 1.0, 1.0
    source_prob = lambda s: torch.exp(-0.5 * (s * s).sum(-1) / 4.0) / (2 * math.pi * 4.0)
    vt = math.exp(-2 * a * T) * (4 - 1 / 2 / a) + 1 / 2 / a
    target_prob = lambda s: torch.exp(-0.5 * (s * s).sum(-1) / vt) / (2 * math.pi * vt)
    # solvers.py:254-278
    z = f32(torch.randn(n_mc, 2, generator=g, dtype=torch.float64))
    refshim.Draws.set(normal=z)
    fake_cond_ = jnp.ones((n_mc, 1)) * cond
    samples, log_prob = model.apply.sample_and_log_prob(params, cond=fake_cond_, seed=jax.random.PRNGKey(1),
                                                        sample_shape=(n_mc, ))
    rmse_mc = torch.sqrt(((torch.exp(log_prob) - (source_prob(samples) * (1 - cond) + target_prob(samples) * cond))**2).mean())
    # solvers.py:282-301
    x = np.linspace(-5, 5, grid_size)
    X, Y = np.meshgrid(x, x)
    XY = jnp.hstack([X.reshape(-1, 1), Y.reshape(-1, 1)])
    rmse_grid = torch.sqrt(((torch.exp(model.apply.log_prob(params, XY, jnp.ones(1) * cond)) -
                             (source_prob(XY) * (1 - cond) + target_prob(XY) * cond))**2).mean())
  k = 3   # the fixture keeps every third grid point and the exact sums (the full arrays are 10 x 100 x 100 doubles)
  return {"shape": np.array([2, shape.num_layers, shape.mlp_layers, shape.hidden, shape.num_bins]),
          "blob": pack(shape, params, torch.float64), "snap_sub": snap[:, ::k, ::k], "snap_sum": snap.sum((1, 2)),
          "stride": np.array(k), "domain2": np.array(dom), "t_traj": t_traj, "dens2_sub": dens2[:, ::k, ::k],
          "dens2_sum": dens2.sum((1, 2)), "r": r_, "traj": traj, "latent_mc": z, "rmse_mc": rmse_mc,
          "rmse_grid": rmse_grid, "grid_size": np.array(grid_size), "fp_a": np.array(a), "fp_T": np.array(T)}


def dr_case(model, D, L, H, sub_dim, sigma, n, seed):
  """cnf_ot/dr/trainers.py:train run for ONE epoch with a no-op optimiser: its own `loss_fn` (:91-111) and
  `jax.value_and_grad(loss_fn)(params, data)` (:117) on the two unconditional flows it builds (:41-73)."""
  refshim.stub_trainer()
  from cnf_ot.dr import trainers as ref_trainers  # the reference
  from cnf_ot_b200.layout import FlowShape
  from oracle import flow as oflow
  spec = oflow.FlowSpec(D, L, [H, H], 5, conditional=False)
  shape = FlowShape(D, L, 2, H, 5, conditional=False)

  def mk(s):
    p = oflow.perturb_params(oflow.init_params(spec, s), sigma, seed=s + 10)
    return {m: {k: v.to(torch.float32).to(v.dtype) for k, v in lv.items()} for m, lv in p.items()}
  g = torch.Generator().manual_seed(seed)
  x = f32(torch.randn(n, D, generator=g, dtype=torch.float64) * 1.5)
  enc, dec = mk(1), mk(2)
  refshim.Trainer.init_queue = [enc, dec] if model == "enc_dec" else [dec]
  refshim.Trainer.captured = []
  cfg = {"cnf": {"flow_num_layers": L, "hidden_size": H, "mlp_num_layers": 2, "num_bins": 5}, "train": {"lr": 1e-3}}
  ret = ref_trainers.train(jax.random.PRNGKey(0), x, D, sub_dim, model, 1, cfg)
  (loss, grads), = refshim.Trainer.captured
  assert float(ret[-1][0]) == float(loss) and not refshim.Trainer.init_queue
  out = {"shape": np.array([D, L, 2, H, 5]), "sub_dim": np.array(sub_dim), "x": x, "loss": loss,
         "blob_decoder": pack(shape, dec, torch.float64),
         "grad_decoder": pack(shape, grads["decoder"] if model == "enc_dec" else grads, torch.float64)}
  if model == "enc_dec":
    out["blob_encoder"] = pack(shape, enc, torch.float64)
    out["grad_encoder"] = pack(shape, grads["encoder"], torch.float64)
  return out


def symbolic_spline_case(K, n, seed):
  """The reference's own symbolic statement of the rational-quadratic map (cnf_ot/models/nsf_symbol.py:3-10:
  f = yk + alpha / beta), lambdified with sympy and evaluated at the knots oracle/rqs.py derives from raw parameters
  drawn as the reference test draws them (tests/test_rqs_accuracy.py:60-69).  Pins the in-bin formula and its
  derivative (the log-det) of the restated spline to a file of the reference; the knot parametrisation (softmax
  widths / heights, softplus slopes) remains distrax's published one."""
  import contextlib
  import io
  import sympy
  with contextlib.redirect_stdout(io.StringIO()):   # the module prints a simplified derivative on import
    from cnf_ot.models import nsf_symbol as ns
  args = (ns.x, ns.xk, ns.xk1, ns.yk, ns.yk1, ns.deltak, ns.deltak1)
  f = sympy.lambdify(args, ns.f, "numpy")
  df = sympy.lambdify(args, sympy.diff(ns.f, ns.x), "numpy")
  from oracle import rqs as orqs
  g = torch.Generator().manual_seed(seed)
  theta = f32(torch.randn(n, 3 * K + 1, generator=g, dtype=torch.float64) * 0.5)
  x = f32((torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * 19.9)   # inside the knot range
  xp, yp, sl = (a.numpy() for a in orqs.normalize_knots(theta))
  xn = x.numpy()
  k = np.clip((xn[:, None] >= xp).sum(-1) - 1, 0, K - 1)
  r = np.arange(n)
  a = (xn, xp[r, k], xp[r, k + 1], yp[r, k], yp[r, k + 1], sl[r, k], sl[r, k + 1])
  return {"theta": theta, "x": x, "bin": k, "y": f(*a), "logdet": np.log(df(*a))}


def save(name, d):
  out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
  print(name, "loss" in out and float(out["loss"]), {k: v.shape for k, v in out.items()})


# name -> (type, subtype, B, lambda, t_batch_size, seed, overrides); the reference's mixture source and its
# nongradient / gradient drifts exist for dim 2 only, lorenz for dim 3
STEP_CASES = {
  "ref_step_ot_free_d2": ("ot", "free", 256, 500.0, 2, 101, {}),
  "ref_step_ot_obstacle_d2": ("ot", "obstacle", 256, 500.0, 2, 102, {}),
  "ref_step_rwpo_double_well_d2": ("rwpo", "double_well", 256, 500.0, 2, 103, {}),
  "ref_step_rwpo_quadratic_d3": ("rwpo", "quadratic", 128, 100.0, 1, 104, dict(dim=3, sigma=0.1)),
  "ref_step_fp_gradient_d2": ("fp", "gradient", 256, 500.0, 2, 105, {}),
  "ref_step_fp_nongradient_d2": ("fp", "nongradient", 256, 100.0, 1, 106, {}),
  "ref_step_fp_lorenz_d3": ("fp", "lorenz", 128, 100.0, 1, 107, dict(dim=3, sigma=0.1)),
  # other network shapes: hidden 64 (wide-conditioner engine on the GPU), 3 hidden layers, 3 flow layers, 8 bins
  "ref_step_ot_obstacle_d2_h64": ("ot", "obstacle", 256, 500.0, 1, 108, dict(H=64, sigma=0.05)),
  "ref_step_rwpo_double_well_d2_m3": ("rwpo", "double_well", 256, 500.0, 1, 109, dict(M=3, sigma=0.2)),
  "ref_step_fp_nongradient_d2_l3": ("fp", "nongradient", 256, 100.0, 1, 110, dict(L=3, sigma=0.2)),
  "ref_step_ot_free_d2_k8_h32": ("ot", "free", 256, 500.0, 2, 111, dict(K=8, H=32, sigma=0.1)),
}

if __name__ == "__main__":
  save("ref_flow_d2", flow_case(2, 2, 2, 16, 5, 0.3, 256, 21))
  save("ref_flow_d3_h8", flow_case(3, 3, 1, 8, 3, 0.1, 128, 22))
  save("ref_rqs_symbolic_k5", symbolic_spline_case(5, 512, 27))
  save("ref_rqs_symbolic_k8", symbolic_spline_case(8, 256, 28))
  save("ref_dr_enc_dec_d4", dr_case("enc_dec", 4, 2, 16, 2, 0.08, 192, 25))
  save("ref_dr_dec_only_d4", dr_case("dec_only", 4, 2, 16, 2, 0.08, 192, 26))
  save("ref_density_d2", density_case(0.3, 2048, 60, 29))
  save("ref_energy_d2", energy_case(2, 0.3, 64, 4, 2.0, 2.0, 23))
  save("ref_energy_d3", energy_case(3, 0.1, 32, 3, 4.0, 1.0, 24))
  for name, (typ, sub, B, lam, Tn, seed, kw) in STEP_CASES.items():
    save(name, step_case(typ, sub, B, lam, Tn, seed, **dict(kw)))
