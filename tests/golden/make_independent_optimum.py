"""Solver-independent check of the OPTIMUM of the restated NLP: the same
RacingNLP functions the oracle IPM uses (objective, constraints, exact derivatives) are handed to
scipy.optimize.minimize(method="trust-constr") - a different algorithm (trust-region barrier /
SQP, no filter line search, no IPOPT heuristics).  Its optima are committed as
tests/golden/independent_optimum.npz; oracle/ipm.py and the CUDA kernels have to land on the same
points.  This is NOT an IPOPT result and is not presented as one.

    python tests/golden/make_independent_optimum.py        (about a minute per problem)
"""
import os
import sys
import time

import numpy as np
from scipy.optimize import Bounds, NonlinearConstraint, minimize
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.nlp import RacingNLP                      # noqa: E402
from mpc_racing_b200.workload import load_fixture     # noqa: E402

PROBLEMS = [0, 7, 11]          # indices into cfg2_shanghai_4096.npz (benign: IPM converges in < 30 its)


def solve_one(nlp):
    eq = nlp.gl == nlp.gu
    con = NonlinearConstraint(nlp.g, nlp.gl, nlp.gu, jac=lambda w: csr_matrix(nlp.jac(w)),
                              hess=lambda w, v: nlp.hess(w, 0.0, v))
    res = minimize(nlp.f, nlp.w0, jac=nlp.grad_f, hess=lambda w: nlp.hess(w, 1.0, np.zeros(nlp.m)),
                   constraints=[con], method="trust-constr",
                   options=dict(gtol=1e-9, xtol=1e-12, barrier_tol=1e-10, maxiter=3000,
                                initial_barrier_parameter=0.1, verbose=0))
    return res, int(eq.sum())


def main():
    d = load_fixture("cfg2_shanghai_4096.npz")
    out = dict(index=np.array(PROBLEMS))
    objs, Us, info = [], [], []
    for b in PROBLEMS:
        nlp = RacingNLP(d["x0"][b], d["u_prev"][b], d["s0"][b], d["cx"][b], d["cy"][b], N=30, Ts=0.05)
        t = time.time()
        res, neq = solve_one(nlp)
        U, S, X = nlp.split(res.x)
        print(f"problem {b}: f* = {res.fun:.9f}  iters {res.nit}  optimality {res.optimality:.2e}  "
              f"constr_violation {res.constr_violation:.2e}  {time.time() - t:.1f} s", flush=True)
        objs.append(res.fun)
        Us.append(U.T.copy())
        info.append([res.nit, res.optimality, res.constr_violation])
    out.update(obj=np.array(objs), U=np.array(Us), info=np.array(info))
    np.savez(os.path.join(HERE, "independent_optimum.npz"), **out)


if __name__ == "__main__":
    main()
