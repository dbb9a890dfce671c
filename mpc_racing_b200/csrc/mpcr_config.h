// Defaults (control/ControllerParameters.py:4-32, control/MPC.py:152-161, IPOPT 3.14
// documented defaults) and the conversion of the public mpcr_config into the
// internal Config.  Host-only, shared by the C ABI and the test emulator.
#pragma once
#include "../../include/mpcr.h"
#include "mpcr_types.h"

namespace mpcr {

inline void config_default(mpcr_config* c) {
  c->N = 30;
  c->model = MPCR_MODEL_DYNAMIC;
  c->wrap_constraint = 1;
  c->lane_constraint = 0;
  c->device = 0;
  c->pointers_on_device = 0;
  c->max_iter = 500;
  c->acceptable_iter = 15;
  c->Ts = 0.05;
  c->lambda_s = 300.0;
  c->alpha_L = 500.0;
  c->min_steer = -0.9;
  c->max_steer = 0.9;
  c->min_throttle = -1.0;
  c->max_steer_delta = 0.2;
  c->min_steer_delta = -0.2;
  c->max_throttle_delta = 2.0;
  c->min_throttle_delta = -0.4;
  c->q_v_max = 2.0;
  c->v_max = 50.0;
  c->min_s_delta = 0.1;
  c->alpha_c = 1000.0;
  c->d_max = 0.85;
  c->q_v_y = 50.0;
  c->n_exp = 2.0;
  c->beta_delta = 5000.0;
  c->tol = 1e-4;
  c->acceptable_tol = 1e-2;
  c->dual_inf_tol = 1.0;
  c->constr_viol_tol = 1e-4;
  c->compl_inf_tol = 1e-4;
  c->mu_init = 0.1;
  c->solve_kernel = MPCR_KERNEL_AUTO;
  c->integrator = MPCR_INTEGRATOR_EULER;
  c->restoration = 0;
  c->schedule = MPCR_SCHEDULE_AUTO;
  c->polish_tol = 0.0;
}

inline void config_convert(const mpcr_config& c, Config* out) {
  Config& k = *out;
  k.N = c.N;
  k.Ts = c.Ts;
  k.model = c.model;
  k.wrap = c.wrap_constraint;
  k.lane_con = c.lane_constraint;
  k.integrator = c.integrator;
  Params& p = k.p;
  p.lambda_s = c.lambda_s;
  p.alpha_L = c.alpha_L;
  p.min_steer = c.min_steer;
  p.max_steer = c.max_steer;
  p.min_throttle = c.min_throttle;
  p.max_steer_delta = c.max_steer_delta;
  p.min_steer_delta = c.min_steer_delta;
  p.max_throttle_delta = c.max_throttle_delta;
  p.min_throttle_delta = c.min_throttle_delta;
  p.q_v_max = c.q_v_max;
  p.v_max = c.v_max;
  p.min_s_delta = c.min_s_delta;
  p.alpha_c = c.alpha_c;
  p.d_max = c.d_max;
  p.q_v_y = c.q_v_y;
  p.beta_delta = c.beta_delta;
  p.n_exp = (int)(c.n_exp + 0.5);
  IpmOptions& o = k.o;
  o.max_iter = c.max_iter;
  o.tol = c.tol;
  o.acceptable_tol = c.acceptable_tol;
  o.acceptable_iter = c.acceptable_iter;
  o.dual_inf_tol = c.dual_inf_tol;
  o.constr_viol_tol = c.constr_viol_tol;
  o.compl_inf_tol = c.compl_inf_tol;
  o.acceptable_dual_inf_tol = 1e10;
  o.acceptable_constr_viol_tol = 1e-2;
  o.acceptable_compl_inf_tol = 1e-2;
  o.nlp_scaling_max_gradient = 100.0;
  o.nlp_scaling_min_value = 1e-8;
  o.bound_relax_factor = 1e-8;
  o.bound_push = 0.01;
  o.bound_frac = 0.01;
  o.mu_init = c.mu_init;
  o.mu_linear_decrease_factor = 0.2;
  o.mu_superlinear_decrease_power = 1.5;
  o.barrier_tol_factor = 10.0;
  o.mu_min = 1e-11;
  o.tau_min = 0.99;
  o.kappa_d = 1e-5;
  o.kappa_sigma = 1e10;
  o.constr_mult_init_max = 1000.0;
  o.bound_mult_init_val = 1.0;
  o.s_max = 100.0;
  o.theta_max_fact = 1e4;
  o.theta_min_fact = 1e-4;
  o.gamma_theta = 1e-5;
  o.gamma_phi = 1e-8;
  o.delta = 1.0;
  o.s_theta = 1.1;
  o.s_phi = 2.3;
  o.eta_phi = 1e-8;
  o.alpha_min_frac = 0.05;
  o.alpha_red_factor = 0.5;
  o.max_soc = 4;
  o.kappa_soc = 0.99;
  o.delta_w_init = 1e-4;
  o.delta_w_min = 1e-20;
  o.delta_w_max = 1e40;
  o.kappa_w_minus = 1.0 / 3.0;
  o.kappa_w_plus = 8.0;
  o.kappa_w_plus_first = 100.0;
  o.restoration = c.restoration;
  o.resto_max_iter = 100;
  o.resto_weight = 1.0;
  o.kappa_resto = 0.9;
  o.polish_tol = c.polish_tol;
}

inline bool config_valid(const mpcr_config& c, const char** why) {
  if (c.N < 2 || c.N > kMaxN) { *why = "N must be in [2, 32]"; return false; }
  if (!(c.Ts > 0.0)) { *why = "Ts must be positive"; return false; }
  if (c.model < 0 || c.model > 2) { *why = "unknown model"; return false; }
  if (c.max_iter < 0) { *why = "max_iter must be >= 0"; return false; }
  if (c.n_exp < 1.0) { *why = "n_exp must be >= 1"; return false; }
  if (c.solve_kernel < 0 || c.solve_kernel > 1) { *why = "unknown solve_kernel"; return false; }
  if (c.integrator < 0 || c.integrator > 1) { *why = "unknown integrator"; return false; }
  if (c.restoration < 0 || c.restoration > 2) { *why = "restoration must be 0, 1 or 2"; return false; }
  if (c.schedule < 0 || c.schedule > 2) { *why = "unknown schedule"; return false; }
  if (!(c.polish_tol >= 0.0)) { *why = "polish_tol must be >= 0"; return false; }
  if (!(c.tol > 0.0) || !(c.acceptable_tol > 0.0)) { *why = "tol and acceptable_tol must be positive"; return false; }
  return true;
}

}  // namespace mpcr
