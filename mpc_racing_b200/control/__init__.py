"""Mirror of the reference's ``control`` package for the solve path.

Import the controller the way the reference's callers do (agent.py:12):
``from mpc_racing_b200.control.MPC import MPC``.
"""
from .ControllerParameters import FixedControllerParameters, RuntimeControllerParameters  # noqa: F401
