"""GPU: restoration in the warp kernel - default-path timing unchanged? config-3 rescue and cost."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch, load_fixture
dev = torch.device("cuda", 0)
host = make_batch(16384)
din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
for kw in (dict(), dict(restoration=True)):
    s = BatchedMPC(N=30, Ts=0.05, device=0, on_device=True, **kw)
    for _ in range(3):
        r = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
        torch.cuda.synchronize()
    st = r.status.cpu().numpy()
    print("bench batch", kw, "kernel ms %.2f" % s.last_kernel_ms(), "converged", int((st >= 0).sum()),
          {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}, flush=True)
d = load_fixture("cfg3_t1_triple_512.npz")
for kw in (dict(), dict(restoration=True)):
    s = BatchedMPC(N=30, Ts=0.05, lane_constraint=True, **kw)
    for _ in range(2):
        r = s.solve(d["x0"], d["u_prev"], d["s0"], d["cx"], d["cy"], max_err=d["max_err"])
    print("config 3", kw, "kernel ms %.1f" % s.last_kernel_ms(), "converged", int((r.status >= 0).sum()), "of 512")
