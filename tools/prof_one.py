"""One solve launch for ncu:  python tools/prof_one.py <B> <warp|auto> [N] [launches] [restoration]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch
B = int(sys.argv[1]); sel = sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
resto = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
dev = torch.device("cuda", 0)
host = make_batch(B, fixture="prod_t4_256.npz" if N <= 16 else "cfg2_shanghai_4096.npz")
din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
s = BatchedMPC(N=N, Ts=0.05, model="dynamic", device=0, on_device=True, solve_kernel=sel, restoration=resto)
for _ in range(n):
    r = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
torch.cuda.synchronize()
print(sel, B, "N", N, "resto", resto, "kernel ms", s.last_kernel_ms(), "conv", int((r.status >= 0).sum()),
      "mean iters %.2f" % r.iters.double().mean().item(), s.launch_info())
