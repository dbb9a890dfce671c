"""GPU: kernel time of one configuration (development aid).
    python tools/kernel_time.py <B> <N> <warp|auto> [restoration 0|1] [reps] [euler|rk4] [dynamic|kinematic|blended]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch
B = int(sys.argv[1]); N = int(sys.argv[2]); sel = sys.argv[3]
resto = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
integrator = sys.argv[6] if len(sys.argv) > 6 else "euler"
model = sys.argv[7] if len(sys.argv) > 7 else "dynamic"
dev = torch.device("cuda", 0)
host = make_batch(B, fixture="prod_t4_256.npz" if N <= 16 else "cfg2_shanghai_4096.npz")
din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
s = BatchedMPC(N=N, Ts=0.05, model=model, device=0, on_device=True, solve_kernel=sel, restoration=resto,
               integrator=integrator)
ms = []
for _ in range(reps):
    r = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
    ms.append(s.last_kernel_ms())
print(os.environ.get("MPCR_LIB", "libmpcr.so"), "B", B, "N", N, sel, model, integrator, "resto", int(resto), "kernel ms", ["%.2f" % m for m in ms],
      "conv", int((r.status >= 0).sum()), "mean iters %.2f" % r.iters.double().mean().item(),
      "us per problem-iteration (all warps busy) %.3f" % (1e3 * min(ms) / float(r.iters.double().sum().item())), s.launch_info())
