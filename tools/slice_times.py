"""GPU: kernel time of each rank's slice of the weak-scaling batch, all on ONE GPU - tells a slow
GPU from a slice that is harder to pack (development aid for the multi-GPU record).
    python tools/slice_times.py [world] [batch per rank]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch
from mpc_racing_b200.sharding import shard_range
world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
per = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
full = make_batch(world * per, fixture="cfg2_shanghai_4096.npz")
dev = torch.device("cuda", 0)
s = BatchedMPC(N=30, Ts=0.05, model="dynamic", device=0, on_device=True, restoration=True)
for r in range(world):
    lo, hi = shard_range(world * per, r, world)
    din = {k: torch.from_numpy(np.ascontiguousarray(full[k][lo:hi])).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
    ms = []
    for _ in range(4):
        res = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
        ms.append(s.last_kernel_ms())
    it = res.iters.cpu().numpy()
    print(f"slice {r} of {world}: kernel ms {['%.2f' % m for m in ms]} iterations {int(it.sum())} longest {int(it.max())} "
          f"converged {int((res.status >= 0).sum())}")
