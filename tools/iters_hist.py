"""Iteration-count / status distribution of the bench workload and a tail estimate (GPU)."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
host = make_batch(B)
s = BatchedMPC(N=30, Ts=0.05, model="dynamic", device=0, on_device=True)
din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
for _ in range(2):
    res = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
torch.cuda.synchronize()
ms = s.last_kernel_ms()
it = res.iters.cpu().numpy(); st = res.status.cpu().numpy()
print("kernel ms", ms, "launch", s.launch_info())
print("status counts", {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))})
q = [0, 1, 5, 25, 50, 75, 90, 95, 99, 99.9, 100]
print("iters percentiles", dict(zip(q, np.percentile(it, q).tolist())), "mean", it.mean(), "sum", it.sum())
top = np.argsort(it)[-20:]
print("top iters", [(int(b), int(it[b]), int(st[b])) for b in top])
nw = s.launch_info()["blocks"]
print("ideal ms if perfectly balanced over %d warps at %.3f us/iter:" % (nw, 1e3 * ms * nw / it.sum()), ms)
print("max-iter problem alone would take ms:", it.max() * ms * nw / it.sum())
