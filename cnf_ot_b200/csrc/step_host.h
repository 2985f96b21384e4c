// Host-side helpers shared by the C-ABI implementation (api.cu) and the
// host test harness: the problem descriptor -> per-row weights mapping.
#pragma once

#include "../../include/cnfot.h"
#include "dispatch.h"
#include "step_math.cuh"

namespace cnfot {

// Scalings of applications.py folded into per-row weights:
//   ot    (:377-402)  lambda * [KL(0) + KL(T)] + sum_t kinetic / Tn (+ sum_t obstacle, not / Tn)
//   rwpo  (:405-421)  lambda * rKL(0) + potential(T) + sum_t kinetic_score * T / Tn
//   fp    (:424-441)  lambda * rKL(0) [beta = 4] + sum_t flow_matching * T / Tn
// mean over rows AND dims times D/2 == sum / (2 rows).
template <typename T>
inline int make_step_consts(const cnfot_problem_desc& p, int D, double lambda, int64_t B_global,
                            int64_t b_global, int n_t, StepConsts<T>* out, const char** err) {
  StepConsts<T> c;
  c.type = p.type;
  c.potential = kPotNone;
  c.drift = 0;
  double Tn = (double)n_t;
  double horizon = p.type == CNFOT_OT ? 1.0 : (double)p.T;
  double beta = p.type == CNFOT_FP ? 4.0 : (double)p.beta;
  double dt = p.dt, dx = p.dx;
  if (p.type == CNFOT_FP) dt = dx = 0.01;  // hard-coded at applications.py:286,301
  c.horizon = (T)horizon;
  c.a = (T)p.a;
  c.dt = (T)dt;
  c.dx = (T)dx;
  c.kappa = (T)0;
  c.var_src = (T)(2.0 / beta * (horizon + 1.0));
  c.var_tgt = (T)(2.0 / beta);
  c.w_fit = (T)(lambda / (double)B_global);
  c.w_pot = (T)0;
  if (p.type == CNFOT_OT) {
    if (p.subtype != CNFOT_OT_FREE && p.subtype != CNFOT_OT_OBSTACLE) { *err = "unknown ot subtype"; return 1; }
    c.w_kin = (T)(1.0 / (2.0 * (double)b_global * Tn));
    if (p.subtype == CNFOT_OT_OBSTACLE) {
      c.potential = kPotObstacle;
      c.a = (T)0;
      c.w_pot = (T)(1.0 / (double)b_global);
    }
  } else if (p.type == CNFOT_RWPO) {
    if (p.subtype != CNFOT_POT_QUADRATIC && p.subtype != CNFOT_POT_DOUBLE_WELL) { *err = "unknown rwpo pot_type"; return 1; }
    c.potential = p.subtype == CNFOT_POT_QUADRATIC ? kPotQuadratic : kPotDoubleWell;
    c.w_pot = (T)(1.0 / (double)B_global);
    c.w_kin = (T)(horizon / (2.0 * (double)b_global * Tn));
    c.kappa = (T)(1.0 / beta);
  } else if (p.type == CNFOT_FP) {
    c.w_kin = (T)(horizon / (2.0 * (double)b_global * Tn));
    c.kappa = (T)p.sigma;
    if (p.subtype == CNFOT_FP_GRADIENT) {
      if (D != 2) { *err = "gradient drift is defined for dim == 2 only"; return 1; }
      c.drift = kDriftGradient;
    } else if (p.subtype == CNFOT_FP_NONGRADIENT) {
      if (D % 2 != 0) { *err = "nongradient drift needs an even dim"; return 1; }
      c.drift = kDriftNonGradient;
    } else if (p.subtype == CNFOT_FP_LORENZ) {
      if (D != 3) { *err = "Lorenz dynamics is only defined for 3 dim!"; return 1; }
      c.drift = kDriftLorenz;
    } else { *err = "unknown fp velocity_field_type"; return 1; }
  } else {
    *err = "Unknown problem type";
    return 1;
  }
  *out = c;
  return 0;
}

}  // namespace cnfot
