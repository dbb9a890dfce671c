"""Golden vectors of the UNMODIFIED reference solve (control/MPC.py:10-184: CasADi Opti + IPOPT).

Needs the reference tree and the CasADi wheel (casadi==3.6.5, requirements.txt:8).  Neither is
available in the build container of this repository (no network, CasADi not in the offline
wheelhouse), so this script is the hook a maintainer - or a later round on a box that has CasADi -
runs ONCE to pin the IPOPT iteration path:

    python baseline/casadi_ref.py --reference /path/to/mpc-racing [--cases cfg1,cfg2,cfg3,prod] [--limit 256]

For every problem of the committed input fixtures (mpc_racing_b200/data/fixtures/*.npz) it
constructs ``control.MPC.MPC`` exactly as agent.py:171-184 / script/test_mpc.py:42-53 do, with
``PYTHONBREAKPOINT=0`` (the reference calls breakpoint() on failure, control/MPC.py:173) and
IPOPT's console output sent to /dev/null, and stores

    ret (States [6,N+1], U [2,N], S_hat [N+1], e_hat_c [N], e_hat_l [N]), dual = lam_g, the
    return status, iteration count and final objective (``sol.stats()``)

into tests/golden/<case>_ipopt.npz.  tests/test_casadi_golden.py gates the CUDA path (and both
oracles) on those files at the north-star tolerances and skips while they do not exist.

The reference's lane row is commented out (control/MPC.py:135), so the cfg3 goldens are the
reference's NLP WITHOUT the lane row; the lane-row extension has no IPOPT counterpart to pin.
"""
import argparse
import contextlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIX = os.path.join(ROOT, "mpc_racing_b200", "data", "fixtures")
GOLD = os.path.join(ROOT, "tests", "golden")

# case -> (fixture, N, Ts): the horizons / steps the fixtures were generated for (tests/golden/make_golden.py)
CASES = {
    "cfg1": ("cfg1_test_mpc.npz", 25, 0.1),             # script/test_mpc.py's own input
    "cfg2": ("cfg2_shanghai_4096.npz", 30, 0.05),
    "cfg3": ("cfg3_t1_triple_512.npz", 30, 0.05),
    "prod": ("prod_t4_256.npz", 15, 0.05),
}


def have_casadi():
    try:
        import casadi  # noqa: F401
        return True
    except Exception:
        return False


@contextlib.contextmanager
def quiet_fd1():
    """IPOPT prints through the C stdout: point file descriptor 1 at /dev/null for the solve."""
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(null)
        os.close(saved)


def solve_reference(x0, u_prev, s0, cx, cy, max_err, N, Ts):
    """One unmodified ``control.MPC.MPC`` construction; returns the golden record of the solve."""
    from control.MPC import MPC
    from control.ControllerParameters import RuntimeControllerParameters
    from models.State import State
    thr = None if u_prev[0] != u_prev[0] else float(u_prev[0])
    st = None if u_prev[1] != u_prev[1] else float(u_prev[1])
    state0 = State(x=float(x0[0]), y=float(x0[1]), yaw=float(x0[2]), v_x=float(x0[3]), v_y=float(x0[4]),
                   yaw_dot=float(x0[5]), throttle=thr, steer=st)
    with quiet_fd1():
        mpc = MPC(state0=state0, s0=float(s0), centerline_x_poly_coeffs=list(map(float, cx)),
                  centerline_y_poly_coeffs=list(map(float, cy)), max_error=float(max_err),
                  runtime_params=RuntimeControllerParameters(), Ts=Ts, N=N)
    sol, ret, dual = mpc.solution()
    States, U, S_hat, eC, eL = ret
    rec = dict(X=np.asarray(States, float), U=np.asarray(U, float), S=np.asarray(S_hat, float),
               eC=np.asarray(eC, float), eL=np.asarray(eL, float))
    if sol is not None:
        stats = sol.stats()
        rec.update(success=1, return_status=stats.get("return_status", ""), iters=int(stats.get("iter_count", -1)),
                   obj=float(stats["iterations"]["obj"][-1]) if "iterations" in stats else float("nan"),
                   lam_g=np.asarray(dual, float).ravel())
    else:
        rec.update(success=0, return_status="failed", iters=-1, obj=float("nan"), lam_g=np.zeros(0))
    return rec


STATUS_CODES = {"Solve_Succeeded": 0, "Solved_To_Acceptable_Level": 1, "Maximum_Iterations_Exceeded": -1,
                "Restoration_Failed": -2, "Error_In_Step_Computation": -3, "Invalid_Number_Detected": -4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("MPC_RACING_REFERENCE", "/root/reference"))
    ap.add_argument("--cases", default="cfg1,cfg2,cfg3,prod")
    ap.add_argument("--limit", type=int, default=0, help="first K problems of every fixture (0 = all)")
    a = ap.parse_args()
    if not have_casadi():
        print("casadi is not importable here: nothing written (tests/test_casadi_golden.py keeps skipping)")
        return 2
    os.environ["PYTHONBREAKPOINT"] = "0"
    sys.path.insert(0, a.reference)
    for case in a.cases.split(","):
        fixture, N, Ts = CASES[case]
        d = np.load(os.path.join(FIX, fixture))
        B = d["x0"].shape[0] if not a.limit else min(a.limit, d["x0"].shape[0])
        recs = [solve_reference(d["x0"][b], d["u_prev"][b], d["s0"][b], d["cx"][b], d["cy"][b],
                                d["max_err"][b] if "max_err" in d.files else 0.0, N, Ts) for b in range(B)]
        ls = max(r["lam_g"].size for r in recs)
        lam = np.zeros((B, ls))
        for b, r in enumerate(recs):
            lam[b, :r["lam_g"].size] = r["lam_g"]
        out = dict(N=N, Ts=Ts, index=np.arange(B),
                   status=np.array([STATUS_CODES.get(r["return_status"], -9) for r in recs], np.int32),
                   success=np.array([r["success"] for r in recs], np.int32),
                   iters=np.array([r["iters"] for r in recs], np.int32), obj=np.array([r["obj"] for r in recs]),
                   U=np.array([r["U"] for r in recs]), X=np.array([r["X"] for r in recs]),
                   S=np.array([r["S"] for r in recs]), eC=np.array([r["eC"] for r in recs]),
                   eL=np.array([r["eL"] for r in recs]), lam_g=lam)
        np.savez_compressed(os.path.join(GOLD, f"{case}_ipopt.npz"), **out)
        print(case, "written:", B, "problems,", int(out["success"].sum()), "solved by IPOPT")
    return 0


if __name__ == "__main__":
    sys.exit(main())
