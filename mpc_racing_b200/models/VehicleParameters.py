"""Vehicle constants (reference: models/VehicleParameters.py:8-41).  The CUDA
kernels carry the same numbers in csrc/mpcr_model.cuh (namespace vp); this class
is the Python-side view with the reference's attribute names."""
from dataclasses import dataclass


@dataclass
class VehicleParameters:
    m: float = 1845.0
    max_steer = 70.0
    min_steer = -70.0
    eta_motor = 0.9
    T_max = 743.0
    r_wheel = 0.37
    C_wheel = 2 * 3.14 * r_wheel       # pi ~ 3.14, exactly as the reference computes it
    R = 9.0
    rho = 1.225
    C_d = 0.23
    A_f = 2.2
    C_roll = 0.012
    max_rpm = 15000
    regen_brake_accel = 0.2
    Iz: float = 3960.0
    lf: float = 0.8
    lr: float = 2
    Cf: float = 65_000
    Cr: float = 65_000
    g: float = 9.81
    Vblendmin: float = 2
    Vblendmax: float = 15
    Ts: int = 0.05
    car_width = 1.85
