"""Oracle solutions of whole fixture batches, committed as golden vectors so that the GPU parity
tests can compare EVERY problem of BASELINE config 2 (4096), config 3 (512, lane rows) and the
production-horizon batch (256, N = 15) without running the slow numpy oracle on the GPU box.

    python tests/golden/make_oracle_solutions.py [cores]       (about 10 minutes on 8 cores)

Stores per problem: status, iters, obj, U (2 x N).  Generator of the inputs: make_golden.py."""
import os
import sys
from multiprocessing import get_context

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

CASES = {
    "cfg2_shanghai_4096": dict(N=30, Ts=0.05, lane=False),
    "cfg3_t1_triple_512": dict(N=30, Ts=0.05, lane=True),
    "prod_t4_256": dict(N=15, Ts=0.05, lane=False),
    # the lane-row batch once more with the restoration phase (mpcr_config.restoration)
    "cfg3_t1_triple_512+resto": dict(N=30, Ts=0.05, lane=True, restoration=True),
}


def _init():
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass


def _one(args):
    from parity_util import oracle_solve
    x0, up, s0, cx, cy, me, kw = args
    extra = dict(lane=True, max_error=float(me)) if kw["lane"] else {}
    opts = dict(restoration=True) if kw.get("restoration") else None
    r = oracle_solve(x0, up, s0, cx, cy, N=kw["N"], Ts=kw["Ts"], options=opts, **extra)
    return r["status"], r["iters"], r["obj"], r["U"]


def main():
    from mpc_racing_b200.workload import load_fixture
    cores = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    ctx = get_context("fork")
    with ctx.Pool(cores, initializer=_init) as pool:
        only = sys.argv[2] if len(sys.argv) > 2 else None
        for name, kw in CASES.items():
            if only and name != only:
                continue
            d = load_fixture(name.split("+")[0] + ".npz")
            B = d["x0"].shape[0]
            jobs = [(d["x0"][b], d["u_prev"][b], d["s0"][b], d["cx"][b], d["cy"][b], d["max_err"][b], kw)
                    for b in range(B)]
            res = pool.map(_one, jobs, chunksize=4)
            np.savez_compressed(os.path.join(HERE, name.replace("+", "_") + "_oracle.npz"),
                                status=np.array([r[0] for r in res], dtype=np.int32),
                                iters=np.array([r[1] for r in res], dtype=np.int32),
                                obj=np.array([r[2] for r in res]),
                                U=np.array([r[3] for r in res]))
            st = np.array([r[0] for r in res])
            print(name, "done:", {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}, flush=True)


if __name__ == "__main__":
    main()
