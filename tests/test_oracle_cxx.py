"""The compiled CPU oracle (oracle/cxx, C++17 + OpenMP) against the numpy oracle.

The two share no derivative code (AD of the reference's expressions vs hand-derived formulas)
and no linear algebra (block-tridiagonal Bunch-Kaufman LDL^T vs scipy's dense LDL^T); they share
the algorithm text (oracle/ipm.py = IPOPT's published algorithm).  Agreement of iteration counts
and optima therefore pins the NLP restatement and the solver transcription against each other.
"""
import os

import numpy as np
import pytest

from oracle import cxx, vehicle as V
from parity_util import oracle_solve

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mpc_racing_b200", "data", "fixtures")


def test_bunch_kaufman_solve_and_inertia():
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 16, 40):
        for trial in range(6):
            Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            ev = rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)
            A = (Q * ev) @ Q.T
            A = 0.5 * (A + A.T)
            if trial % 2 and n > 2:          # KKT shape: zero block on the diagonal
                k = n // 3
                A[-k:, -k:] = 0.0
            b = rng.standard_normal(n)
            x, inertia = cxx.bk_solve(A, b)
            w = np.linalg.eigvalsh(A)
            assert inertia == (int((w > 0).sum()), int((w < 0).sum()), 0)
            assert np.abs(A @ x - b).max() <= 1e-8 * max(1.0, np.abs(x).max() * np.abs(A).max())


@pytest.mark.parametrize("model", [V.MODEL_DYNAMIC, V.MODEL_KINEMATIC, V.MODEL_BLENDED])
@pytest.mark.parametrize("integrator", ["euler", "rk4"])
def test_ad_derivatives_match_hand_derived(model, integrator):
    """AD of control/MPC.py:186-283 (C++) vs the hand-derived numpy derivatives (oracle/vehicle.py)."""
    rng = np.random.default_rng(7 + model)
    for _ in range(20):
        x = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-3, 3), rng.uniform(3, 30),
                      rng.normal(0, 0.5), rng.normal(0, 0.2)])
        if model == V.MODEL_BLENDED:
            x[3] = rng.uniform(3, 14)       # inside the blending band: the lambda derivatives are live
        u = np.array([rng.uniform(-1, 0.85), rng.uniform(-0.5, 0.5)])
        lam = rng.standard_normal(6)
        f, J, H = cxx.model_eval(model, x, u, 0.05, lam, integrator)
        if integrator == "euler":
            fr = V.f_model(model, x, u, 0.05)
            Jr, Hr = V.model_derivs(model, x, u, 0.05, lam)
        else:
            fr = V.rk4_step(model, x, u, 0.05)
            Jr, Hr = V.rk4_derivs(model, x, u, 0.05, lam)
        assert np.abs(f - fr).max() <= 1e-12 * max(1.0, np.abs(fr).max())
        assert np.abs(J - Jr).max() <= 1e-9 * max(1.0, np.abs(Jr).max())
        assert np.abs(H - Hr).max() <= 1e-8 * max(1.0, np.abs(Hr).max())


@pytest.mark.parametrize("case", ["plain", "kinematic", "blended", "nowrap", "lane", "lane+resto", "N15", "rk4"])
def test_live_against_numpy_oracle(case):
    """Same problems through oracle/ipm.py (live) and the C++ oracle: same status, iteration count
    within one, objective to 1e-8, inputs to 1e-6."""
    name = "cfg3_t1_triple_512" if case.startswith("lane") else "cfg2_shanghai_4096"
    d = np.load(os.path.join(FIX, name + ".npz"))
    idx = [0, 3] if case != "lane+resto" else [5, 21]
    kw_np, kw_cx, opts, N = {}, {}, None, 30
    if case == "kinematic":
        kw_np, kw_cx = dict(model=V.MODEL_KINEMATIC), dict(model="kinematic")
    elif case == "blended":
        kw_np, kw_cx = dict(model=V.MODEL_BLENDED), dict(model="blended")
    elif case == "nowrap":
        kw_np, kw_cx = dict(wrap=False), dict(wrap=False)
    elif case == "N15":
        N = 15
    elif case == "rk4":
        kw_np, kw_cx = dict(integrator="rk4"), dict(integrator="rk4")
    if case.startswith("lane"):
        kw_cx = dict(lane=True, max_err=d["max_err"][idx])
        if case == "lane+resto":
            opts = dict(restoration=True)
            kw_cx["restoration"] = True
    r = cxx.solve_batch(d["x0"][idx], d["u_prev"][idx], d["s0"][idx], d["cx"][idx], d["cy"][idx], N=N, **kw_cx)
    for j, b in enumerate(idx):
        if case.startswith("lane"):
            kw_np = dict(lane=True, max_error=float(d["max_err"][b]))
        ref = oracle_solve(d["x0"][b], d["u_prev"][b], d["s0"][b], d["cx"][b], d["cy"][b], N=N, options=opts, **kw_np)
        assert r["status"][j] == ref["status"]
        assert abs(int(r["iters"][j]) - ref["iters"]) <= 1
        if ref["status"] >= 0:
            assert abs(r["obj"][j] - ref["obj"]) <= 1e-8 * abs(ref["obj"])
            dU = np.abs(r["U"][j] - ref["U"])
            dU[0, -1] = 0.0
            assert dU.max() <= 1e-6
            lam = r["lam_g"][j][:ref["m"]]
            assert np.abs(lam - ref["lam_g"]).max() <= 1e-6 * max(1.0, np.abs(ref["lam_g"]).max())


def test_dense_and_block_tridiagonal_linear_algebra_agree():
    d = np.load(os.path.join(FIX, "cfg2_shanghai_4096.npz"))
    sl = slice(0, 16)
    a = cxx.solve_batch(d["x0"][sl], d["u_prev"][sl], d["s0"][sl], d["cx"][sl], d["cy"][sl], linsol="blocktri")
    b = cxx.solve_batch(d["x0"][sl], d["u_prev"][sl], d["s0"][sl], d["cx"][sl], d["cy"][sl], linsol="dense")
    assert (a["status"] == b["status"]).all()
    assert (np.abs(a["iters"] - b["iters"]) <= 1).all()
    assert np.abs(a["obj"] - b["obj"]).max() <= 1e-8 * np.abs(b["obj"]).max()


@pytest.mark.parametrize("name,gold,kw,N", [
    ("cfg2_shanghai_4096", "cfg2_shanghai_4096_oracle", {}, 30),
    ("cfg3_t1_triple_512", "cfg3_t1_triple_512_oracle", dict(lane=True), 30),
    ("cfg3_t1_triple_512", "cfg3_t1_triple_512_resto_oracle", dict(lane=True, restoration=True), 30),
    ("prod_t4_256", "prod_t4_256_oracle", {}, 15),
])
def test_whole_batches_against_numpy_oracle_golden(name, gold, kw, N):
    """EVERY problem of the fixture batches: the committed numpy-oracle solutions
    (tests/golden/*_oracle.npz, generator make_oracle_solutions.py) against the C++ oracle."""
    d = np.load(os.path.join(FIX, name + ".npz"))
    g = np.load(os.path.join(G, gold + ".npz"))
    r = cxx.solve_batch(d["x0"], d["u_prev"], d["s0"], d["cx"], d["cy"], N=N, max_err=d["max_err"], **kw)
    assert (r["status"] == g["status"]).all()
    assert (r["iters"] == g["iters"]).mean() >= 0.97
    both = g["status"] >= 0
    rel = np.abs(r["obj"] - g["obj"]) / np.abs(g["obj"])
    assert rel[both].max() <= 1e-6
    dU = np.abs(r["U"] - g["U"])
    dU[:, 0, -1] = 0.0
    assert dU[both].max() <= 1e-4


def test_polish_keeps_status_and_tightens():
    d = np.load(os.path.join(FIX, "cfg2_shanghai_4096.npz"))
    sl = slice(0, 32)
    a = cxx.solve_batch(d["x0"][sl], d["u_prev"][sl], d["s0"][sl], d["cx"][sl], d["cy"][sl])
    p = cxx.solve_batch(d["x0"][sl], d["u_prev"][sl], d["s0"][sl], d["cx"][sl], d["cy"][sl], polish_tol=1e-9)
    assert (a["status"] == p["status"]).all() and (a["iters"] == p["iters"]).all()
    assert (p["polish_iters"] >= 0).all() and p["polish_iters"].max() > 0
    assert (p["obj"] <= a["obj"] + 1e-9 * np.abs(a["obj"])).all()
