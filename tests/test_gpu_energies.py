"""GPU parity of the evaluation energies: cnfot_kinetic_energy vs the oracle's
restatement of utils.calc_kinetic_energy / calc_score_kinetic_energy (cnf_ot/utils.py:311-389)."""
import pytest
import torch

from cnf_ot_b200 import ops, random, utils
from cnf_ot_b200.flows import ParamTree, RQSFlow
from cnf_ot_b200.layout import pack
from oracle import energies as oen
from util import make_cfg, make_params, shape_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,sigma", [(2, 0.3), (3, 0.1), (10, 0.05)])
@pytest.mark.parametrize("with_score", [False, True])
def test_kinetic_energy_matches_oracle(D, sigma, with_score, engine):
  cfg = make_cfg(dim=D)
  shape = shape_of(cfg)
  spec, params = make_params(cfg, sigma)
  W = pack(shape, params).cuda()
  g = torch.Generator().manual_seed(21)
  n_t, batch = 5, 300  # ragged tile
  latent = torch.randn(n_t, batch, D, generator=g, dtype=torch.float64).float()
  ts = torch.linspace(0.0, 1.5, n_t, dtype=torch.float64).float().tolist()
  if with_score:
    ref = oen.score_kinetic_energy(spec, params, latent.double(), ts, beta=2.0)
  else:
    ref = oen.kinetic_energy(spec, params, latent.double(), ts)
  got = ops.kinetic_energy(shape, W, latent.reshape(-1, D).cuda(), ts, with_score=with_score, kappa=0.5,
                           latent_blocks=n_t)
  # finite differences with 1/dt = 1/dx = 100 amplify float32 rounding: stated tolerance 2e-4 relative; with the score
  # (central differences of the log-prob across spline knots) 1e-3: the error is dominated by the few rows whose stencil
  # sits within rounding of a knot and varies 1e-5 ... 4e-4 with the seed on every engine (tools/diag_noise.py)
  tol = 1e-3 if with_score else 2e-4
  assert abs(float(got) - float(ref)) <= tol * abs(float(ref)), (float(got), float(ref))
  # one latent block reused for every time
  got1 = ops.kinetic_energy(shape, W, latent[0].cuda(), ts, with_score=with_score, kappa=0.5, latent_blocks=1)
  lat1 = latent[:1].expand(n_t, batch, D).double()
  if with_score:
    ref1 = oen.score_kinetic_energy(spec, params, lat1, ts, beta=2.0)
  else:
    ref1 = oen.kinetic_energy(spec, params, lat1, ts)
  assert abs(float(got1) - float(ref1)) <= tol * abs(float(ref1))


def test_reference_signatures_and_identity_flow():
  """utils.calc_kinetic_energy(sample_fn, params, rng, batch_size, t_size, dim): at the reference
  initialisation the flow is the identity for every t, so both energies have closed forms:
  kinetic = 0; score-corrected = mean(|x|^2) / (2 beta^2) -> dim / (2 beta^2)."""
  model = RQSFlow((2, ), 2, [16, 16], 5)
  params = model.init(random.PRNGKey(0), torch.zeros(1, 2), torch.zeros(1))
  e = utils.calc_kinetic_energy(model.apply.sample, params, random.PRNGKey(1), batch_size=4096, t_size=300, dim=2)
  assert abs(float(e)) < 1e-6
  beta = 2.0
  es = utils.calc_score_kinetic_energy(model.apply.sample, model.apply.log_prob, params, T=1.0, beta=beta, dim=2,
                                       rng=random.PRNGKey(2), batch_size=8192, t_size=130)
  assert abs(float(es) - 2 / (2 * beta**2)) < 0.01 * 2 / (2 * beta**2)
  e2 = utils.calc_score_kinetic_energy(model.apply.sample, model.apply.log_prob, params, T=1.0, beta=beta, dim=2,
                                       rng=random.PRNGKey(2), batch_size=8192, t_size=130)
  assert abs(float(e2) - float(es)) < 1e-12  # equal key => equal draws (partial sums may combine in another order)
  with pytest.raises(ValueError):
    utils.calc_kinetic_energy(model.apply.sample, params, random.PRNGKey(1), batch_size=64, t_size=4, dim=3)


def test_solver_evaluate_tail():
  """solvers.evaluate mirrors solvers.py:138-172; at the identity flow rwpo/quadratic has closed forms:
  e_kin = T dim / (2 beta^2), e_pot = dim / 2."""
  import copy
  from cnf_ot_b200 import solvers
  from util import BASE_CFG
  cfg = copy.deepcopy(BASE_CFG)
  cfg["general"]["type"] = "rwpo"
  cfg["rwpo"].update(T=2, beta=4, a=1, pot_type="quadratic")
  model, _, T = solvers.build(cfg)
  params = model.init(random.PRNGKey(0), torch.zeros(1, 2), torch.zeros(1))
  out = solvers.evaluate(cfg, model, params, random.PRNGKey(5), batch_size=16384, t_size=64, verbose=False)
  assert abs(out["e_kin"] - 2 * 2 / (2 * 16)) < 0.02 * 0.125 and abs(out["e_pot"] - 1.0) < 0.03
  import math
  assert abs(out["true_val"] - 2 * (1 + math.log(3.0)) / 4) < 1e-12
  cfg["general"]["type"] = "ot"
  out = solvers.evaluate(cfg, model, params, random.PRNGKey(5), batch_size=4096, t_size=50, verbose=False)
  assert abs(out["kinetic_more"]) < 1e-6 and abs(out["kinetic_less"]) < 1e-6
