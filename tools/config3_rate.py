"""GPU: BASELINE config 3 at full size - 16384 dynamic-bicycle problems on t1_triple with the lane
row (512 reference-sampled start states x 32 seeded perturbations), with and without restoration."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch
B = 16384
host = make_batch(B, fixture="cfg3_t1_triple_512.npz")
dev = torch.device("cuda", 0)
din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy", "max_err")}
for kw in (dict(), dict(restoration=True)):
    s = BatchedMPC(N=30, Ts=0.05, lane_constraint=True, device=0, on_device=True, **kw)
    for _ in range(3):
        r = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"], max_err=din["max_err"])
        torch.cuda.synchronize()
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    ms = s.last_kernel_ms()
    print(kw, "kernel ms %.1f" % ms, "converged %d of %d (%.2f %%)" % ((st >= 0).sum(), B, 100 * (st >= 0).mean()),
          "-> %.0f converged solves/s" % ((st >= 0).sum() / ms * 1e3), "mean iters %.1f" % it.mean(),
          {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}, flush=True)
