// RK4 discretisation of the three bicycle models with analytic Jacobian and
// multiplier-weighted Hessian, as fused fp64 device functions (north-star item 2; an
// extension: the reference's NLP uses the explicit-Euler maps, control/MPC.py:186-260).
//
// Vector field = the model's Euler increment per unit time, g(y, u) = (F(y, u; Ts) - y) / Ts
// (the ODE right-hand side for the dynamic single-track model).  With D_i = F(y_i, u) - y_i:
//     y_1 = x, y_2 = x + D_1/2, y_3 = x + D_2/2, y_4 = x + D_3,
//     x+  = x + (D_1 + 2 D_2 + 2 D_3 + D_4) / 6.
// Derivatives are carried over q = (yaw, v_x, v_y, r, thr, st) - X and Y enter only through the
// identity - by forward sensitivities of the stage points and adjoint stage weights mu_i:
//     H = sum_i Q_i' Hess(mu_i' D)(y_i, u) Q_i.
// Same restatement as oracle/vehicle.py:rk4_derivs (checked against torch.autograd).
#pragma once
#include "mpcr_model.cuh"

namespace mpcr {

// dense d(F - y)/dq (6 state rows x 6 q columns) from the sparse Euler linearisation
MPCR_HD void rk4_dense_increment_jac(const StageLin& L, double Ts, double D[6][6]) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) D[i][j] = 0.0;
  for (int k = 0; k < 3; ++k) { D[0][k] = L.aX[k]; D[1][k] = L.aY[k]; }
  D[2][3] = Ts;
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) D[3 + i][1 + k] = L.d[i][k] - (i == k ? 1.0 : 0.0);
    for (int k = 0; k < 2; ++k) D[3 + i][4 + k] = L.bu[i][k];
  }
}

// f[6] = x+; J[6][6] = d x+ / d q (may be null; d x+/d(X, Y) is the identity on rows X, Y);
// hm[21] = packed lam-weighted Hessian over q (may be null; needs lam and J)
MPCR_HD void rk4_model(int model, const double* x, const double* u, double Ts, const double* lam,
                       double* f, double (*J)[6], double* hm) {
  const double cw[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  const double cc[4] = {0.0, 0.5, 0.5, 1.0};
  double y[4][6], D[4][6][6], Q[4][4][6];   // Q_i: rows (yaw, v_x, v_y, r) of dy_i/dq
  double d[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
  double S[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) S[i][j] = 0.0;
  if (J)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) J[i][j] = (i >= 2 && j == i - 2) ? 1.0 : 0.0;
  for (int s = 0; s < 4; ++s) {
    for (int k = 0; k < 6; ++k) y[s][k] = x[k] + cc[s] * d[k];
    double F[6];
    if (J) {
      StageLin L;
      lin_model(model, y[s], u, Ts, nullptr, F, L, nullptr);
      rk4_dense_increment_jac(L, Ts, D[s]);
      for (int a = 0; a < 4; ++a)
        for (int j = 0; j < 6; ++j) Q[s][a][j] = (j == a ? 1.0 : 0.0) + cc[s] * S[2 + a][j];
      // S = D_s [Q_s; e_thr; e_st]
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double v = (j >= 4) ? D[s][i][j] : 0.0;
          for (int a = 0; a < 4; ++a) v += D[s][i][a] * Q[s][a][j];
          S[i][j] = v;
        }
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) J[i][j] += cw[s] * S[i][j];
    } else {
      f_model(model, y[s], u, Ts, F);
    }
    for (int k = 0; k < 6; ++k) {
      d[k] = F[k] - y[s][k];
      acc[k] += cw[s] * d[k];
    }
  }
  for (int k = 0; k < 6; ++k) f[k] = x[k] + acc[k];
  if (!hm || !J || !lam) return;
  for (int k = 0; k < kHmDoubles; ++k) hm[k] = 0.0;
  double mu_next[6] = {0, 0, 0, 0, 0, 0};
  for (int s = 3; s >= 0; --s) {
    double mu[6];
    for (int k = 0; k < 6; ++k) mu[k] = cw[s] * lam[k];
    if (s < 3) {
      // + c_{s+1} (dD_{s+1}/dy)' mu_{s+1}: only the (yaw, v_x, v_y, r) columns are non-zero
      for (int a = 0; a < 4; ++a) {
        double v = 0.0;
        for (int i = 0; i < 6; ++i) v += D[s + 1][i][a] * mu_next[i];
        mu[2 + a] += cc[s + 1] * v;
      }
    }
    StageLin L;
    double h[kHmDoubles];
    lin_model(model, y[s], u, Ts, mu, nullptr, L, h);
    // hm += W' h W with W = [Q_s; e_thr; e_st] (6 x 6)
    double W[6][6];
    for (int a = 0; a < 4; ++a)
      for (int j = 0; j < 6; ++j) W[a][j] = Q[s][a][j];
    for (int j = 0; j < 6; ++j) { W[4][j] = (j == 4) ? 1.0 : 0.0; W[5][j] = (j == 5) ? 1.0 : 0.0; }
    double T[6][6];   // h W
    for (int a = 0; a < 6; ++a)
      for (int j = 0; j < 6; ++j) {
        double v = 0.0;
        for (int b = 0; b < 6; ++b) v += h[hm_idx(a, b)] * W[b][j];
        T[a][j] = v;
      }
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) {
        double v = 0.0;
        for (int a = 0; a < 6; ++a) v += W[a][i] * T[a][j];
        hm[hm_idx(i, j)] += v;
      }
    for (int k = 0; k < 6; ++k) mu_next[k] = mu[k];
  }
}

// Dense outputs over z = (x, u): jac[6][8], hess[8][8] (either may be null), Euler or RK4.
MPCR_HD void model_eval_dense(int model, int integrator, const double* x, const double* u, double Ts,
                              const double* lam, double* f, double* jac, double* hess) {
  double Jq[6][6], hm[kHmDoubles];
  const bool want_h = hess != nullptr && lam != nullptr;
  if (integrator == 1) {
    rk4_model(model, x, u, Ts, lam, f, (jac || want_h) ? Jq : nullptr, want_h ? hm : nullptr);
  } else {
    StageLin L;
    double zero[6] = {0, 0, 0, 0, 0, 0};
    lin_model(model, x, u, Ts, want_h ? lam : zero, f, L, want_h ? hm : nullptr);
    double D[6][6];
    rk4_dense_increment_jac(L, Ts, D);
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) Jq[i][j] = D[i][j] + ((i >= 2 && j == i - 2) ? 1.0 : 0.0);
  }
  if (jac) {
    for (int i = 0; i < 6; ++i) {
      jac[i * 8 + 0] = (i == 0) ? 1.0 : 0.0;
      jac[i * 8 + 1] = (i == 1) ? 1.0 : 0.0;
      for (int j = 0; j < 6; ++j) jac[i * 8 + 2 + j] = Jq[i][j];
    }
  }
  if (hess) {
    for (int k = 0; k < 64; ++k) hess[k] = 0.0;
    if (want_h)
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) hess[(2 + i) * 8 + 2 + j] = hm[hm_idx(i, j)];
  }
}

}  // namespace mpcr
