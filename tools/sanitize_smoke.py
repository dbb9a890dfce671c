"""GPU: small invocations of every kernel of libmpcr.so for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpc_racing_b200.solver import BatchedMPC
from mpc_racing_b200.workload import load_fixture
from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
from mpc_racing_b200.splines import ParameterizedCenterline

d = load_fixture(); d3 = load_fixture("cfg3_t1_triple_512.npz"); p = load_fixture("prod_t4_256.npz")
a = lambda f, n: tuple(f[k][:n] for k in ("x0", "u_prev", "s0", "cx", "cy"))
for kw, args, extra in ((dict(N=30), a(d, 40), {}), (dict(N=30, restoration=True), a(d, 40), {}),
                        (dict(N=30, lane_constraint=True, restoration=True), a(d3, 40), dict(max_err=d3["max_err"][:40])),
                        (dict(N=15, model="blended"), a(p, 33), {}), (dict(N=12, integrator="rk4"), a(p, 5), {}),
                        (dict(N=2, model="kinematic", wrap_constraint=False), a(p, 3), {}),
                        (dict(N=32, polish_tol=1e-9), a(d, 3), {})):
    r = BatchedMPC(Ts=0.05, schedule="curvature", **kw).solve(*args, **extra)
    print(kw, "status", np.unique(r.status, return_counts=True))
cl = ParameterizedCenterline("t4")
s = np.linspace(0, cl.length, 50)
cl.eval_all(s); cl.coeffs(s, 45.0); cl.project(cl.eval_all(s)[:, 0] + 0.3, cl.eval_all(s)[:, 1] - 0.2); cl.lookup_error(s, 45.0)
ro = ClosedLoopRollouts(cl, N=15, Ts=0.05, plant="blended", restoration=True, estimate_state=True, use_graph=False)
x0, s0 = start_states(cl, 48, seed=3, v0=8.0)
res = ro.run(x0, s0, max_steps=6, seg_bounds=[0.0, 5.0, 10.0])
print("rollout steps", res["steps"][:5], "fails", int(res["fail_steps"].sum()))
print("SANITIZE SMOKE OK")
