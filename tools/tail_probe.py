"""How much of the kernel time is tail (stragglers dispatched late)?  Solves the bench batch in index
order, then longest-first (iteration counts from the first run: an oracle schedule, not a usable one)
and shortest-first."""
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
    return best, r.iters.cpu().numpy()


s2 = BatchedMPC(N=30, Ts=0.05, model="dynamic", device=0, on_device=True, solve_kernel="warp", schedule="curvature")
s_idx = s
s = BatchedMPC(N=30, Ts=0.05, model="dynamic", device=0, on_device=True, solve_kernel="warp", schedule="index")
base = np.arange(16384)
ms0, it = run(base)
print("index order        ", ms0)
order = np.argsort(-it, kind="stable")
print("longest first      ", run(order)[0])
print("shortest first     ", run(order[::-1])[0])
rng = np.random.default_rng(0)
print("random permutation ", run(rng.permutation(16384))[0])
# long ones (top 2 %) first, rest in index order
top = order[:328]
rest = np.setdiff1d(base, top)
print("top 2 % first      ", run(np.concatenate([top, rest]))[0])
s = s2
print("curvature schedule (library, index-order input) ", run(base)[0])
print("curvature schedule on a random permutation      ", run(rng.permutation(16384))[0])
