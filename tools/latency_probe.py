"""Per-iteration time of the warp kernel against the number of resident warps, for (a) one problem
replicated (all warps walk the same instruction stream, nearly in lockstep) and (b) distinct problems.
Separates instruction-fetch sharing from latency hiding.  python tools/latency_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch

dev = torch.device("cuda", 0)
host = make_batch(16384)
s = BatchedMPC(N=30, Ts=0.05, model="dynamic", device=0, on_device=True, solve_kernel="warp")
keys = ("x0", "u_prev", "s0", "cx", "cy")


def run(idx):
    din = {k: torch.from_numpy(np.ascontiguousarray(host[k][idx])).to(dev) for k in keys}
    best = 1e9
    for _ in range(4):
        r = s.solve(*[din[k] for k in keys])
        torch.cuda.synchronize()
        best = min(best, s.last_kernel_ms())
    it = r.iters.cpu().numpy()
    return best, it


for B in (1, 148, 296, 592, 1036, 2072, 4144):
    rep = np.full(B, 17)
    ms_r, it_r = run(rep)
    mix = np.arange(B)
    ms_m, it_m = run(mix)
    waves = max(1.0, B / 1036.0)
    print(f"B={B:5d}  replicated: {ms_r:8.3f} ms, iters {it_r[0]:3d}, {1e3 * ms_r / it_r[0] / waves:7.1f} us/iter/wave   "
          f"distinct: {ms_m:8.3f} ms, iters mean {it_m.mean():5.1f} max {it_m.max():3d}, "
          f"{1e3 * ms_m / (it_m.sum() / B) / waves:7.1f} us/mean-iter/wave", flush=True)
