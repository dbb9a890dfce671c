"""GPU: closed-loop throughput of BASELINE config 4 (16k parallel rollouts, blended plant, N = 15,
warm-started): simulated control steps per second, and the solve kernel's share of a step."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
from mpc_racing_b200.splines.ParameterizedCenterline import ParameterizedCenterline

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
track = sys.argv[3] if len(sys.argv) > 3 else "t4"
cl = ParameterizedCenterline(track)
ro = ClosedLoopRollouts(cl, N=15, Ts=0.05, plant="blended")
x0, s0 = start_states(cl, R, seed=1234, v0=5.0)
ro.run(x0, s0, max_steps=5)                       # warm-up (allocations, first-launch costs)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = ro.run(x0, s0, max_steps=K)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{track}: {R} rollouts x {K} steps in {dt:.2f} s = {R * K / dt:.0f} control steps/s "
      f"({dt / K * 1e3:.1f} ms per batched step; last solve kernel {ro.solver.last_kernel_ms():.1f} ms); "
      f"failed solves {int(res['fail_steps'].sum())} of {R * K}")
