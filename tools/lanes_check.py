"""GPU: lane-per-problem kernel vs warp-per-problem kernel on the bench workload."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
host = make_batch(B)
din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
res = {}
for name, sel in (("warp", 1), ("lanes", 2)):
    s = BatchedMPC(N=30, Ts=0.05, model="dynamic", device=0, on_device=True, solve_kernel=sel)
    for _ in range(3):
        r = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
        torch.cuda.synchronize()
        ms = s.last_kernel_ms()
        print(name, "kernel ms %.2f" % ms, "solves/s %.0f" % (B / ms * 1e3), s.launch_info(), flush=True)
    res[name] = {k: getattr(r, k).cpu().numpy() for k in ("U", "X", "S", "obj", "status", "iters", "lam_g")}
a, b = res["warp"], res["lanes"]
print("status equal", (a["status"] == b["status"]).mean(), "iters equal", (a["iters"] == b["iters"]).mean())
print("status counts lanes", {int(k): int(v) for k, v in zip(*np.unique(b["status"], return_counts=True))})
same = a["iters"] == b["iters"]
rel = np.abs(a["obj"] - b["obj"]) / np.abs(a["obj"])
print("obj rel diff: same-iters max %.3e, all max %.3e" % (rel[same].max(), rel.max()))
du = np.abs(a["U"] - b["U"]); du[:, 0, -1] = 0
print("U diff: same-iters max %.3e, all p99 %.3e max %.3e" % (du[same].max(), np.percentile(du.max(axis=(1, 2)), 99), du.max()))
print("mean iters warp %.3f lanes %.3f" % (a["iters"].mean(), b["iters"].mean()))
