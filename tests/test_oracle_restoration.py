"""Restoration in the oracle: (i) options['restoration'] = True is the kernels' phase (quadratic
penalty over the same (x, s), restated in oracle/ipm.py); (ii) 'l1' / 'quad' are the studies of
oracle/restoration.py (IPOPT's l1 formulation, with and without relaxed dynamics rows, and the
penalty variant as a nested NLP).  All of them rescue the lane-row problems on which the plain filter
line search stops with status -2 and land on the optimum an independent solver finds."""
import numpy as np

from mpc_racing_b200.workload import load_fixture
from oracle import ipm
from oracle.nlp import RacingNLP

INDEPENDENT_OPTIMUM_478 = -112828.3055199883      # scipy trust-constr, constraint violation 3e-14


def test_restoration_rescues_a_lane_row_failure():
    d = load_fixture("cfg3_t1_triple_512.npz")
    b = 478
    nlp = RacingNLP(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b], N=30, Ts=0.05, lane=True,
                    max_error=float(d['max_err'][b]))
    plain = ipm.solve(nlp)
    assert plain.status == ipm.RESTORATION_NEEDED
    # the kernels' phase and IPOPT's l1 formulation; the nested quadratic variant ('quad') and the l1 one
    # with relaxed dynamics rows (resto_relax_eq=True) give the same optimum (slower; run by hand)
    for opts in (dict(restoration=True), dict(restoration='l1', resto_relax_eq=False)):
        r = ipm.solve(nlp, opts)
        assert r.status == ipm.SOLVE_SUCCEEDED and 1 <= r.n_resto <= 20
        assert abs(r.obj - INDEPENDENT_OPTIMUM_478) <= 1e-6 * abs(INDEPENDENT_OPTIMUM_478)
        g = nlp.g(r.x)
        assert np.maximum(nlp.gl - g, 0).max() <= 1e-4 and np.maximum(g - nlp.gu, 0).max() <= 1e-4
