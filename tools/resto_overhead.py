"""GPU: per-iteration overhead of the restoration BUILD on a batch where no problem needs it."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import make_batch
dev = torch.device("cuda", 0)
for fixture, N in (("prod_t4_256.npz", 15), ("cfg2_shanghai_4096.npz", 30)):
    host = make_batch(16384, fixture=fixture)
    din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
    for kw in (dict(), dict(restoration=True)):
        s = BatchedMPC(N=N, Ts=0.05, device=0, on_device=True, **kw)
        for _ in range(3):
            r = s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
            torch.cuda.synchronize()
        st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
        print(fixture, N, kw, "kernel ms %.2f" % s.last_kernel_ms(), "failed", int((st < 0).sum()), "iters sum", int(it.sum()),
              "max", int(it.max()), "us per iteration-slot %.3f" % (1e3 * s.last_kernel_ms() * 1036 / it.sum()), flush=True)
