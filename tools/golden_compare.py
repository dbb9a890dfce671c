"""GPU: whole fixture batches against the committed oracle solutions (tests/golden/*_oracle.npz)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import load_fixture
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, kw, N in (("cfg2_shanghai_4096", {}, 30), ("cfg3_t1_triple_512", dict(lane_constraint=True), 30),
                    ("prod_t4_256", {}, 15)):
    d = load_fixture(name + ".npz"); g = np.load(os.path.join(G, name + "_oracle.npz"))
    for kern in ("warp", "lanes"):
        s = BatchedMPC(N=N, Ts=0.05, solve_kernel=kern, **kw)
        r = s.solve(d["x0"], d["u_prev"], d["s0"], d["cx"], d["cy"],
                    max_err=d["max_err"] if kw else None)
        same_st = r.status == g["status"]
        both = (r.status >= 0) & (g["status"] >= 0)
        rel = np.abs(r.obj - g["obj"]) / np.abs(g["obj"])
        dU = np.abs(r.U - g["U"]) / np.maximum(1.0, np.abs(g["U"])); last = dU[:, 0, -1].copy(); dU[:, 0, -1] = 0
        du = dU.reshape(len(dU), -1).max(axis=1)
        print(f"{name:22s} {kern:6s} status equal {same_st.mean():.5f} ({(~same_st).sum()} differ: "
              f"{[(int(i), int(r.status[i]), int(g['status'][i])) for i in np.where(~same_st)[0][:8]]}) "
              f"conv-class equal {((r.status>=0)==(g['status']>=0)).mean():.5f} iters equal {(r.iters==g['iters']).mean():.4f}")
        print(f"{'':29s} both-converged {both.sum()}: obj rel max {rel[both].max():.2e} (>1e-6: {(rel[both]>1e-6).sum()}), "
              f"dU max {du[both].max():.2e} (>1e-4: {(du[both]>1e-4).sum()}), last-throttle max {last[both].max():.2e}")
