"""CPU study behind DESIGN.md section 5 ("The 1 -> N curve"): why one rank's slice of the weak-scaling
batch takes 10 % longer than another's for the same number of IPM iterations.

Iteration counts come from the compiled oracle (oracle/cxx - this script is test infrastructure, like
everything that touches oracle/); the persistent launch is modelled as a list schedule of one
problem per worker over 1036 workers (B200: 148 SMs x 7 resident warps) with cost = iterations x
0.128 ms.  It prints, per slice: the ideal makespan (work / workers), index order, the kernel's
curvature-bin order (mpcr.cu:schedule_bin_kernel restated), the perfect longest-first order, and -
if scikit-learn is importable - the order by a gradient-boosted prediction of the iteration count
from eleven input features, with the rank quantiles of the long problems under that order.

    python tests/tail_study.py [slices, default "0,3"] [world, default 4]      (about 80 s per slice on 8 cores)
"""
import heapq
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_racing_b200.workload import make_batch          # noqa: E402
from oracle import cxx                                    # noqa: E402

WORKERS, MS_PER_ITER, PER_RANK = 1036, 0.128, 16384


def makespan(order, cost, workers=WORKERS):
    h = [0.0] * workers
    heapq.heapify(h)
    for b in order:
        heapq.heappush(h, heapq.heappop(h) + cost[b])
    return max(h)


def path_curvature(d, reduce):
    B = len(d["s0"])
    L = np.maximum(5.0, 1.2 * d["x0"][:, 3] * 30 * 0.05)
    k = np.zeros(B)
    for b in range(B):
        s = d["s0"][b] + L[b] * np.arange(16) / 15.0
        x1, x2 = np.polyval(np.polyder(d["cx"][b]), s), np.polyval(np.polyder(d["cx"][b], 2), s)
        y1, y2 = np.polyval(np.polyder(d["cy"][b]), s), np.polyval(np.polyder(d["cy"][b], 2), s)
        n2 = x1 * x1 + y1 * y1
        k[b] = reduce(np.abs(x1 * y2 - y1 * x2) / (n2 * np.sqrt(n2)))
    return k


def curvature_bins(k):
    t = np.clip((np.log2(np.maximum(k, 1e-300)) + 13.287712379549449) / 13.287712379549449, 0, 1)
    return np.argsort(255 - np.clip((t * 255).astype(int), 0, 255), kind="stable")


def features(d, kbar):
    B, x0, s0 = len(d["s0"]), d["x0"], d["s0"]
    ev = lambda c, der: np.array([np.polyval(np.polyder(c[b], der) if der else c[b], s0[b]) for b in range(B)])
    px, py, dx, dy = ev(d["cx"], 0), ev(d["cy"], 0), ev(d["cx"], 1), ev(d["cy"], 1)
    nrm = np.hypot(dx, dy)
    ex, ey = x0[:, 0] - px, x0[:, 1] - py
    kmax = path_curvature(d, np.max)
    head = np.abs(np.angle(np.exp(1j * (x0[:, 2] - np.arctan2(dy, dx)))))
    return np.stack([np.abs((-dy * ex + dx * ey) / nrm), np.abs((dx * ex + dy * ey) / nrm), head, x0[:, 3],
                     np.abs(x0[:, 4]), np.abs(x0[:, 5]), np.log(kbar), np.log(kmax), np.abs(d["u_prev"][:, 0]),
                     np.abs(d["u_prev"][:, 1]), x0[:, 3] ** 2 * kmax], axis=1)


def main():
    slices = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,3").split(",")]
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    full = make_batch(world * PER_RANK, fixture="cfg2_shanghai_4096.npz")
    data = {}
    for r in slices:
        d = {k: np.ascontiguousarray(full[k][r * PER_RANK:(r + 1) * PER_RANK]) for k in ("x0", "u_prev", "s0", "cx", "cy")}
        t = time.time()
        out = cxx.solve_batch(d["x0"], d["u_prev"], d["s0"], d["cx"], d["cy"], N=30, Ts=0.05, restoration=True)
        it = out["iters"].astype(float)
        cost = it * MS_PER_ITER + 0.05
        kbar = path_curvature(d, np.mean)
        print(f"slice {r} of {world} ({time.time() - t:.0f} s on the host): iterations {int(it.sum())}, longest {int(it.max())}, "
              f"{int((it > 60).sum())} problems above 60; makespan ideal {cost.sum() / WORKERS:.2f} ms, index order "
              f"{makespan(np.arange(len(it)), cost):.2f}, curvature bins {makespan(curvature_bins(kbar), cost):.2f}, "
              f"perfect longest-first {makespan(np.argsort(-cost), cost):.2f}; corr(log curvature, iterations) "
              f"{np.corrcoef(np.log(kbar), it)[0, 1]:.2f}")
        data[r] = (features(d, kbar), it, cost)
    if len(slices) < 2:
        return
    try:
        from sklearn.ensemble import GradientBoostingRegressor
    except ImportError:
        print("scikit-learn not importable: predictability part skipped")
        return
    (X0, it0, _), (X1, it1, cost1) = data[slices[0]], data[slices[1]]
    g = GradientBoostingRegressor(n_estimators=200, max_depth=4).fit(X0, it0)
    p = g.predict(X1)
    rank = np.argsort(np.argsort(-p)) / len(p)
    q = np.percentile(rank[it1 > 80], [50, 90, 99, 100])
    print(f"trained on slice {slices[0]}, applied to slice {slices[1]}: corr(prediction, iterations) {np.corrcoef(p, it1)[0, 1]:.2f}, "
          f"makespan in predicted-longest-first order {makespan(np.argsort(-p), cost1):.2f} ms; dispatch rank of the "
          f"problems above 80 iterations: median {q[0]:.3f}, 90 % {q[1]:.3f}, 99 % {q[2]:.3f}, last {q[3]:.3f} "
          f"(0 = dispatched first): a few long problems are not predictable from the inputs and start late")


if __name__ == "__main__":
    main()
