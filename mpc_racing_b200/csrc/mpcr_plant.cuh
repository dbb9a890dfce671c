// The reference's numpy PLANT models (the only vehicle simulator available without CARLA;
// SURVEY 8a-21) as host/device functions:
//   models/Model.py:15-64      Fx: step-table motor efficiency, regen brake at throttle == 0,
//                              Gaussian "carla_penalty"
//   models/Model.py:66-81      steer_cmd_to_angle: step-table gain, np.deg2rad (true pi)
//   models/KinematicBicycleModel.py:11-48, models/DynamicBicycleModel.py:16-78,
//   models/BlendedBicycleModel.py:18-58 (explicit Euler, yaw wrapped by atan2(sin, cos))
// These are NOT the smooth models inside the NLP (mpcr_model.cuh).
#pragma once
#include <math.h>
#include "mpcr_model.cuh"

namespace mpcr {

MPCR_HD double plant_Fx(double throttle, double v_x) {
  const double regen = (throttle == 0.0) ? 0.2 * 9.81 * vp::m : 0.0;   // VehicleParameters.regen_brake_accel
  const double wheel_rpm = (v_x / vp::C_wheel) * 60;
  const double rpm = wheel_rpm * vp::R * 4.5;
  double eta;
  if (rpm < 9000) eta = 1.0;
  else if (rpm < 9500) eta = 0.88;
  else if (rpm < 10400) eta = 0.81;
  else if (rpm < 12500) eta = 0.71;
  else eta = 0.675;
  // scipy.stats.norm.pdf(throttle, loc=0.5, scale=0.0775) * m
  const double z = (throttle - 0.5) / 0.0775;
  const double penalty = exp(-0.5 * z * z) / (0.0775 * sqrt(2.0 * 3.141592653589793)) * vp::m;
  const double wheel_force = throttle * eta * vp::T_max * vp::R / vp::r_wheel;
  const double drag = 0.5 * vp::rho * vp::C_d * vp::A_f * (v_x * v_x);
  const double rolling = vp::C_roll * vp::m * vp::g;
  return wheel_force - drag - rolling - regen - penalty;
}

MPCR_HD double plant_steer_angle(double steer_cmd, double v_x, double v_y) {
  const double vel = sqrt(v_x * v_x + v_y * v_y) * 3.6;
  double gain;
  if (vel < 20.0) gain = 1.0;
  else if (vel < 60.0) gain = 0.9;
  else if (vel < 120.0) gain = 0.8;
  else gain = 0.7;
  return (steer_cmd * vp::max_steer * gain) * (3.141592653589793 / 180.0);   // np.deg2rad
}

MPCR_HD double wrap_angle(double a) { return atan2(sin(a), cos(a)); }

MPCR_HD void plant_kinematic(const double* x, const double* u, double dt, double* o) {
  const double yaw = x[2], v_x = x[3], v_y = x[4], r = x[5];
  const double delta = plant_steer_angle(u[1], v_x, v_y);
  const double Fx = plant_Fx(u[0], v_x);
  const double rr = (v_x / (vp::lr + vp::lf)) * tan(delta);
  o[0] = x[0] + (v_x * cos(yaw) - v_y * sin(yaw)) * dt;
  o[1] = x[1] + (v_x * sin(yaw) + v_y * cos(yaw)) * dt;
  o[2] = wrap_angle(yaw + rr * dt);
  o[3] = v_x + (Fx / vp::m) * dt;
  o[4] = r * vp::lr;
  o[5] = rr;
}

MPCR_HD void plant_dynamic(const double* x, const double* u, double dt, double* o) {
  const double yaw = x[2], v_x = x[3], v_y = x[4], r = x[5];
  const double Fx = plant_Fx(u[0], v_x);
  const double delta = plant_steer_angle(u[1], v_x, v_y);
  const double th_f = atan2(v_y + vp::lf * r, v_x);
  const double th_r = atan2(v_y - vp::lr * r, v_x);
  const double Fyf = vp::Cf * (delta - th_f);
  const double Fyr = vp::Cr * (-th_r);
  const double vxd = ((Fx - Fyf * sin(delta)) / vp::m) + (v_y * r);
  const double vyd = ((Fyf * cos(delta) + Fyr) / vp::m) - (v_x * r);
  const double rd = ((Fyf * cos(delta) * vp::lf) - (Fyr * vp::lr)) / vp::Iz;
  o[0] = x[0] + (v_x * cos(yaw) - v_y * sin(yaw)) * dt;
  o[1] = x[1] + (v_x * sin(yaw) + v_y * cos(yaw)) * dt;
  o[2] = wrap_angle(yaw + r * dt);
  o[3] = v_x + vxd * dt;
  o[4] = v_y + vyd * dt;
  o[5] = r + rd * dt;
}

MPCR_HD void plant_step(int model, const double* x, const double* u, double dt, double* o) {
  if (model == kKinematic) {
    plant_kinematic(x, u, dt, o);
  } else if (model == kDynamic) {
    plant_dynamic(x, u, dt, o);
  } else {
    double k[6], d[6];
    plant_kinematic(x, u, dt, k);
    plant_dynamic(x, u, dt, d);
    const double vel = hypot(x[3], x[4]);
    double lam = (vel - vp::Vblendmin) / (vp::Vblendmax - vp::Vblendmin);
    lam = fmin(fmax(lam, 0.0), 1.0);
    for (int i = 0; i < 6; ++i) o[i] = lam * d[i] + (1 - lam) * k[i];
  }
}

}  // namespace mpcr
