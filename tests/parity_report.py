"""GPU: whole fixture batches against (a) the committed numpy-oracle solutions and (b) the compiled
oracle (oracle/cxx) run on the box's host cores with the same options - plain and polished
(polish_tol = 1e-9 on both sides).  Prints the parity table of DESIGN.md and the outlier lists.

    python tests/parity_report.py [polish_tol]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_racing_b200.solver import BatchedMPC          # noqa: E402
from mpc_racing_b200.workload import load_fixture      # noqa: E402
from oracle import cxx                                  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
PT = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-9


def cmp(a, b, tag):
    st_a, st_b = np.asarray(a["status"]), np.asarray(b["status"])
    both = (st_a >= 0) & (st_b >= 0)
    rel = np.abs(a["obj"] - b["obj"]) / np.abs(b["obj"])
    dU = np.abs(a["U"] - b["U"]) / np.maximum(1.0, np.abs(b["U"]))
    last = dU[:, 0, -1].copy()
    dU[:, 0, -1] = 0
    du = dU.reshape(len(dU), -1).max(axis=1)
    same_min = both & (rel <= 1e-6)
    out = dict(tag=tag, n=int(len(st_a)), status_equal=float((st_a == st_b).mean()),
               class_equal=float(((st_a >= 0) == (st_b >= 0)).mean()),
               iters_equal=float((np.asarray(a["iters"]) == np.asarray(b["iters"])).mean()),
               both_converged=int(both.sum()), obj_within_1e6=float((rel[both] <= 1e-6).mean()),
               du_within_1e4=float((du[both] <= 1e-4).mean()),
               same_min_du_within_1e4=float((du[same_min] <= 1e-4).mean()) if same_min.any() else None,
               same_min_du_max=float(du[same_min].max()) if same_min.any() else None,
               different_minimum=[int(i) for i in np.where(both & (rel > 1e-6))[0]],
               du_outliers_same_min=[int(i) for i in np.where(same_min & (du > 1e-4))[0]],
               status_differs=[(int(i), int(st_a[i]), int(st_b[i])) for i in np.where(st_a != st_b)[0][:16]],
               last_throttle_max=float(last[both].max()))
    if "lam_g" in a and "lam_g" in b and a["lam_g"] is not None:
        m = min(a["lam_g"].shape[1], b["lam_g"].shape[1])
        den = np.maximum(1.0, np.abs(b["lam_g"][:, :m]).max(axis=1))
        dl = np.abs(a["lam_g"][:, :m] - b["lam_g"][:, :m]).max(axis=1) / den
        out["lam_g_within_1e4"] = float((dl[same_min] <= 1e-4).mean()) if same_min.any() else None
        out["lam_g_max"] = float(dl[same_min].max()) if same_min.any() else None
    print(json.dumps(out), flush=True)
    return out


def as_dict(r):
    return dict(status=r.status, iters=r.iters, obj=r.obj, U=r.U, lam_g=r.lam_g)


for name, kw, okw, N in (("cfg2_shanghai_4096", {}, {}, 30),
                         ("cfg2_shanghai_4096", dict(restoration=True), dict(restoration=True), 30),
                         ("cfg3_t1_triple_512", dict(lane_constraint=True), dict(lane=True), 30),
                         ("cfg3_t1_triple_512", dict(lane_constraint=True, restoration=True), dict(lane=True, restoration=True), 30),
                         ("prod_t4_256", {}, {}, 15),
                         ("prod_t4_256", dict(model="blended"), dict(model="blended"), 15)):
    d = load_fixture(name + ".npz")
    args = (d["x0"], d["u_prev"], d["s0"], d["cx"], d["cy"])
    me = d["max_err"] if kw.get("lane_constraint") else None
    tag = name + ("+" + "+".join(sorted(kw)) if kw else "")
    r = as_dict(BatchedMPC(N=N, Ts=0.05, **kw).solve(*args, max_err=me))
    gname = name + ("_resto" if kw.get("restoration") else "") + "_oracle.npz"
    if os.path.exists(os.path.join(G, gname)) and not kw.get("model"):
        g = np.load(os.path.join(G, gname))
        cmp(r, g, tag + " | kernel vs numpy-oracle golden")
    t = time.time()
    c = cxx.solve_batch(*args, N=N, max_err=me, **okw)
    print(f"# oracle/cxx {len(d['x0'])} problems in {time.time() - t:.1f} s", flush=True)
    cmp(r, c, tag + " | kernel vs oracle/cxx")
    rp = as_dict(BatchedMPC(N=N, Ts=0.05, polish_tol=PT, **kw).solve(*args, max_err=me))
    cp = cxx.solve_batch(*args, N=N, max_err=me, polish_tol=PT, **okw)
    cmp(rp, cp, tag + f" | POLISHED {PT:g}: kernel vs oracle/cxx")
