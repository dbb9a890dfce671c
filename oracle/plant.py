"""Oracle (test infrastructure): the numpy PLANT models of the reference, restated.

Follows models/Model.py:15-81 (Fx with step-table efficiency, regen brake, Gaussian
"carla_penalty"; steer_cmd_to_angle with step-table gain and np.deg2rad = true pi),
models/KinematicBicycleModel.py:11-48, models/DynamicBicycleModel.py:16-78 and
models/BlendedBicycleModel.py:18-58.  These are NOT the functions the NLP uses
(control/MPC.py:186-283 - see oracle/vehicle.py); they are the only simulator available
without CARLA and drive the closed-loop configurations.

PINNED against the reference's own classes: tests/golden/plant_golden.npz.
"""
import numpy as np

from . import vehicle as V

REGEN_BRAKE_ACCEL = 0.2      # models/VehicleParameters.py:25


def plant_Fx(throttle, v_x):
    """models/Model.py:15-64."""
    regen = REGEN_BRAKE_ACCEL * 9.81 * V.M if throttle == 0.0 else 0
    wheel_rpm = (v_x / V.C_WHEEL) * 60
    rpm = wheel_rpm * V.R_GEAR * 4.5
    if rpm < 9000:
        eta = 1.0
    elif rpm < 9500:
        eta = 0.88
    elif rpm < 10_400:
        eta = 0.81
    elif rpm < 12_500:
        eta = 0.71
    else:
        eta = 0.675
    # scipy.stats.norm.pdf(throttle, loc=0.5, scale=0.0775) * m
    z = (throttle - 0.5) / 0.0775
    carla_penalty = np.exp(-0.5 * z * z) / (0.0775 * np.sqrt(2 * np.pi)) * V.M
    wheel_force = throttle * eta * V.T_MAX * V.R_GEAR / V.R_WHEEL
    drag = 0.5 * V.RHO * V.C_D * V.A_F * (v_x ** 2)
    rolling = V.C_ROLL * V.M * V.G
    return wheel_force - drag - rolling - regen - carla_penalty


def plant_steer_angle(steer_cmd, v_x, v_y):
    """models/Model.py:66-81."""
    vel = np.sqrt(v_x ** 2 + v_y ** 2) * 3.6
    if vel < 20.0:
        gain = 1.0
    elif vel < 60.0:
        gain = 0.9
    elif vel < 120.0:
        gain = 0.8
    else:
        gain = 0.7
    return np.deg2rad(steer_cmd * V.MAX_STEER_DEG * gain)


def _wrap(yaw):
    return np.arctan2(np.sin(yaw), np.cos(yaw))


def kinematic_step(x, u, dt):
    """models/KinematicBicycleModel.py:11-48."""
    X, Y, yaw, v_x, v_y, r = x
    delta = plant_steer_angle(u[1], v_x, v_y)
    Fx = plant_Fx(u[0], v_x)
    x_new = X + (v_x * np.cos(yaw) - v_y * np.sin(yaw)) * dt
    y_new = Y + (v_x * np.sin(yaw) + v_y * np.cos(yaw)) * dt
    yaw_new = yaw + ((v_x / (V.LR + V.LF)) * np.tan(delta)) * dt
    v_x_new = v_x + (Fx / V.M) * dt
    r_new = (v_x / (V.LR + V.LF)) * np.tan(delta)
    v_y_new = r * V.LR
    return np.array([x_new, y_new, _wrap(yaw_new), v_x_new, v_y_new, r_new])


def dynamic_step(x, u, dt):
    """models/DynamicBicycleModel.py:16-78."""
    X, Y, yaw, v_x, v_y, r = x
    Fx = plant_Fx(u[0], v_x)
    delta = plant_steer_angle(u[1], v_x, v_y)
    th_f = np.arctan2(v_y + V.LF * r, v_x)
    th_r = np.arctan2(v_y - V.LR * r, v_x)
    Fyf = V.CF * (delta - th_f)
    Fyr = V.CR * (-th_r)
    vxd = ((Fx - Fyf * np.sin(delta)) / V.M) + (v_y * r)
    vyd = ((Fyf * np.cos(delta) + Fyr) / V.M) - (v_x * r)
    rd = ((Fyf * np.cos(delta) * V.LF) - (Fyr * V.LR)) / V.IZ
    x_new = X + (v_x * np.cos(yaw) - v_y * np.sin(yaw)) * dt
    y_new = Y + (v_x * np.sin(yaw) + v_y * np.cos(yaw)) * dt
    yaw_new = yaw + r * dt
    return np.array([x_new, y_new, _wrap(yaw_new), v_x + vxd * dt, v_y + vyd * dt, r + rd * dt])


def blended_step(x, u, dt):
    """models/BlendedBicycleModel.py:18-58."""
    vel = np.hypot(x[3], x[4])
    lam = np.clip((vel - V.VBLENDMIN) / (V.VBLENDMAX - V.VBLENDMIN), 0, 1)
    return lam * dynamic_step(x, u, dt) + (1 - lam) * kinematic_step(x, u, dt)


PLANT_STEP = {"kinematic": kinematic_step, "dynamic": dynamic_step, "blended": blended_step}
