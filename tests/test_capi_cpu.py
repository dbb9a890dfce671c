"""No-GPU checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol
include/mpcr.h declares, and the host-side mirrors behave like the reference's classes."""
import ctypes
import os
import re

import numpy as np
import pytest

from mpc_racing_b200 import _capi, build as mbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(mbuild.build())


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "mpcr.h")).read()
    declared = set(re.findall(r"\b(mpcr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_version_and_default_config(lib):
    assert lib.mpcr_version() == 105
    cfg = _capi.MpcrConfig()
    lib.mpcr_config_default.argtypes = [ctypes.POINTER(_capi.MpcrConfig)]
    lib.mpcr_config_default(ctypes.byref(cfg))
    # control/ControllerParameters.py:4-32 and control/MPC.py:152-161
    assert (cfg.N, cfg.Ts, cfg.max_iter) == (30, 0.05, 500)
    assert (cfg.lambda_s, cfg.alpha_L, cfg.q_v_max, cfg.v_max) == (300, 500, 2, 50)
    assert (cfg.alpha_c, cfg.d_max, cfg.q_v_y, cfg.n_exp, cfg.beta_delta) == (1000, 0.85, 50, 2, 5000)
    assert (cfg.tol, cfg.acceptable_tol) == (1e-4, 1e-2)
    assert (cfg.min_throttle_delta, cfg.max_throttle_delta) == (-0.4, 2.0)
    assert cfg.wrap_constraint == 1 and cfg.lane_constraint == 0


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = _capi.MpcrConfig()
    lib.mpcr_config_default.argtypes = [ctypes.POINTER(_capi.MpcrConfig)]
    lib.mpcr_config_default(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    lib.mpcr_create.argtypes = [ctypes.POINTER(_capi.MpcrConfig), ctypes.POINTER(ctypes.c_void_p)]
    rc = lib.mpcr_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -3          # MPCR_ERR_NO_DEVICE: there is no CPU fallback
    from mpc_racing_b200.solver import BatchedMPC
    with pytest.raises(_capi.MpcrError):
        BatchedMPC()


def test_parameter_mirrors_match_reference_surface():
    from mpc_racing_b200.control.ControllerParameters import (FixedControllerParameters,
                                                               RuntimeControllerParameters)
    from mpc_racing_b200.models import State, VehicleParameters
    assert FixedControllerParameters.lookahead_distance == 75.0
    assert RuntimeControllerParameters().as_row() == [1000.0, 0.85, 50.0, 2.0, 5000.0]
    s = State(1, 2, 3, 4, 5, 6).set_controls(0.1, 0.2)
    assert (s.throttle, s.steer) == (0.1, 0.2)
    assert s == State(1, 2, 3, 4, 5, 6, steer=0.2, throttle=0.1)
    assert s != State(1, 2, 3, 4, 5, 6)
    assert VehicleParameters.C_wheel == 2 * 3.14 * 0.37


def test_no_product_import_of_oracle():
    """The product package never imports oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "mpc_racing_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_unknown_layout_is_rejected(lib):
    """solve_kernel values other than AUTO / WARP are rejected at create time (the config check
    runs before any device is touched)."""
    import ctypes as C
    from mpc_racing_b200 import _capi
    cfg = _capi.default_config(lib)
    assert cfg.integrator == 0 and cfg.solve_kernel == 0 and cfg.polish_tol == 0.0
    cfg.solve_kernel = 2
    h = C.c_void_p()
    assert lib.mpcr_create(C.byref(cfg), C.byref(h)) == -1      # MPCR_ERR_INVALID_ARG
