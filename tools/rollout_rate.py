"""GPU: closed-loop throughput (BASELINE config 4 shape: parallel rollouts, blended plant, N = 15,
warm-started) and where a control step's time goes: per-kernel CUDA-event times of ungraphed ticks,
then the graphed engine's steps per second.

    python tools/rollout_rate.py [R] [K] [track] [warp|auto] [keyed dispatch 0|1]
"""
import ctypes as C
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
from mpc_racing_b200.splines.ParameterizedCenterline import ParameterizedCenterline

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
track = sys.argv[3] if len(sys.argv) > 3 else "t4"
layout = sys.argv[4] if len(sys.argv) > 4 else "auto"
keyed = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
cl = ParameterizedCenterline(track, lanes=False, error=False)
ro = ClosedLoopRollouts(cl, N=15, Ts=0.05, plant="blended", restoration=True, solve_kernel=layout)
ro.keyed_dispatch = keyed
x0, s0 = start_states(cl, R, seed=1234, v0=5.0)
S = ro.begin(x0, s0)
ro.advance(S, 6)                                   # warm-up (allocations, first-launch costs, graph capture)
# ---- per-kernel times of 10 ungraphed ticks -------------------------------------------------------
ev = lambda: torch.cuda.Event(enable_timing=True)
tp, ts, ta = [], [], []
with torch.cuda.stream(ro.stream):
    for _ in range(10):
        k = S.k
        e = [ev() for _ in range(4)]
        st = ro.stream.cuda_stream
        e[0].record(ro.stream)
        ro._chk(ro.lib.mpcr_rollout_prepare(ro.track, C.byref(S.bufs), st))
        e[1].record(ro.stream)
        cur, prev = S.plan[k % 2], S.plan[(k + 1) % 2]
        ro.solver.solve(S.x_mpc, S.u_prev, S.progress, S.cx, S.cy, params=S.params, u_warm=prev["U"],
                        want_lam_g=False, out=cur, stream=st)
        e[2].record(ro.stream)
        ro._chk(ro.lib.mpcr_rollout_apply(ro.solver._h, C.byref(S.bufs), cur["U"].data_ptr(), cur["status"].data_ptr(), st))
        e[3].record(ro.stream)
        S.k += 1
        ro.stream.synchronize()
        tp.append(e[0].elapsed_time(e[1])); ts.append(e[1].elapsed_time(e[2])); ta.append(e[2].elapsed_time(e[3]))
print(f"{track} R={R} layout={layout} keyed_dispatch={keyed}: prepare {np.median(tp):.3f} ms, solve {np.median(ts):.3f} ms "
      f"(kernel {ro.solver.last_kernel_ms():.3f}), apply {np.median(ta):.3f} ms; launch {ro.solver.launch_info()}")
if S.k % 2:
    ro.advance(S, 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
r = ro.advance(S, K)
dt = time.perf_counter() - t0
print(f"graphed: {R} rollouts x {K} steps in {dt:.3f} s = {R * K / dt:.0f} control steps/s "
      f"({dt / K * 1e3:.2f} ms per batched step); converged solves {r['converged_solves']} of {R * K}")
