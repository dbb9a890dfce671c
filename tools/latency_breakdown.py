"""B = 1 through the host-pointer C ABI: wall time of the call against the solve kernel's own time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch
h = make_batch(64)
s = BatchedMPC(N=30, Ts=0.05)
wall, kern = [], []
for i in range(64):
    a = [h[k][i:i + 1] for k in ("x0", "u_prev", "s0", "cx", "cy")]
    s.solve(*a)
    t = time.perf_counter(); r = s.solve(*a); wall.append((time.perf_counter() - t) * 1e3); kern.append(s.last_kernel_ms())
wall, kern = np.array(wall), np.array(kern)
print("wall p50 %.3f ms, kernel p50 %.3f ms, overhead p50 %.3f ms (min %.3f)" % (np.median(wall), np.median(kern), np.median(wall - kern), (wall - kern).min()))
