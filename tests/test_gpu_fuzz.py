"""GPU fuzz: the CUDA solver against the independent compiled CPU oracle (oracle/cxx: AD derivatives,
block-tridiagonal LDL^T) over random configurations - horizon, step, model, lane row, wrap rows, restoration, NaN previous inputs,
per-problem controller parameters (well outside the reference's ranges: badly scaled objectives),
warm starts.  Success / failure must agree problem by problem; where the iteration counts agree the
objectives agree to round-off."""
import numpy as np
import pytest

from mpc_racing_b200.workload import load_fixture

pytestmark = pytest.mark.gpu


def fuzz_trials(n_trials, seed=2024):
    rng = np.random.default_rng(seed)
    fx = {k: load_fixture(f) for k, f in (("cfg2", "cfg2_shanghai_4096.npz"), ("cfg3", "cfg3_t1_triple_512.npz"),
                                          ("prod", "prod_t4_256.npz"))}
    for trial in range(n_trials):
        name = str(rng.choice(["cfg2", "cfg3", "prod"]))
        d = fx[name]
        B = int(rng.choice([1, 7, 33, 200]))
        idx = rng.choice(d["x0"].shape[0], B, replace=False)
        N = int(rng.choice([3, 10, 15, 30, 32]))
        kw = dict(N=N, Ts=float(rng.choice([0.05, 0.04])), model=str(rng.choice(["dynamic", "kinematic", "blended"])),
                  wrap_constraint=bool(rng.integers(2)), lane_constraint=bool(name == "cfg3" and rng.integers(2)),
                  restoration=bool(rng.integers(2)))
        up = d["u_prev"][idx].copy()
        if rng.integers(3) == 0:
            up[:, rng.integers(2)] = np.nan
        prm = None
        if rng.integers(2):
            prm = np.stack([rng.uniform(300, 1500, B), rng.uniform(0.5, 0.85, B), rng.uniform(10, 80, B),
                            rng.choice([2.0, 4.0], B), rng.uniform(1000, 8000, B)], axis=1)
        uw = None
        if rng.integers(2):
            uw = np.stack([rng.uniform(-0.5, 0.8, (B, N)), rng.uniform(-0.3, 0.3, (B, N))], axis=1)
        args = (d["x0"][idx], up, d["s0"][idx], d["cx"][idx], d["cy"][idx])
        opt = dict(max_err=d["max_err"][idx] if kw["lane_constraint"] else None, params=prm, u_warm=uw)
        yield trial, name, B, kw, args, opt


def compare(a, b):
    both = (a.status >= 0) & (b.status >= 0)
    same = (a.iters == b.iters) & both
    rel = np.abs(a.obj - b.obj) / np.maximum(1.0, np.abs(a.obj))
    return dict(status_eq=float((a.status == b.status).mean()),
                class_eq=float(((a.status >= 0) == (b.status >= 0)).mean()),
                iters_eq=float(same.mean()),
                worst_same=float(rel[same].max()) if same.any() else 0.0,
                obj_close=float((rel[both] <= 1e-6).mean()) if both.any() else 1.0,
                conv=float(both.mean()),
                finite=bool(np.isfinite(a.U).all() and np.isfinite(b.U).all()))


class _R:
    def __init__(self, d):
        self.__dict__.update(d)


def oracle_cxx(kw, args, opt):
    from oracle import cxx
    o = cxx.solve_batch(*args, N=kw["N"], Ts=kw["Ts"], model=kw["model"], wrap=kw["wrap_constraint"],
                        lane=kw["lane_constraint"], restoration=kw["restoration"], **opt)
    return _R(o)


def test_kernel_agrees_with_the_compiled_oracle_on_random_configurations():
    from mpc_racing_b200.solver import BatchedMPC
    for trial, name, B, kw, args, opt in fuzz_trials(30):
        a = BatchedMPC(solve_kernel="warp", **kw).solve(*args, **opt)
        b = oracle_cxx(kw, args, opt)
        m = compare(a, b)
        ctx = f"trial {trial} {name} B={B} {kw}: {m}"
        assert m["finite"], ctx
        assert m["class_eq"] >= 0.97, ctx            # success / failure, problem by problem
        assert m["worst_same"] <= 1e-7, ctx          # same iteration count => same result
        assert m["obj_close"] >= 0.95, ctx
