"""Golden vectors for the spline layer, produced by the reference's own classes."""
import os

import numpy as np

TRACKS = ["shanghai_intl_circuit", "t1_triple", "t2_triple", "t3", "t4"]


def make_splines(outdir):
    from splines.ParameterizedCenterline import ParameterizedCenterline
    from splines.TrackSegments import TrackSegments
    rng = np.random.default_rng(1234)
    out = {}
    for t in TRACKS:
        cl = ParameterizedCenterline(t)
        L = cl.length
        s = np.concatenate([rng.uniform(-20, 2 * L, 120), [0.0, L - 1e-9, L, 1e-12], cl.spline_x.t[3:9]])
        ev = np.array([[cl.Gx(v), cl.Gy(v), cl.dGx(v), cl.dGy(v), cl.ddGx(v), cl.ddGy(v)] for v in s])
        geo = np.array([[cl.unit_tangent_yaw(v), cl.curvature(v), *cl.unit_principal_normal(v)] for v in s])
        sq = rng.uniform(0, L, 24)
        cxy = np.array([cl.x_as_coeffs(v - 5, la, 4) + cl.y_as_coeffs(v - 5, la, 4)
                        for v in sq for la in (45.0, 75.0)])
        lerr = np.array([[cl.lookup_error(v, 45), cl.lookup_error(v, 75)] for v in sq])
        # projection queries: points +-N(0, 2 m) off the centerline (script/test_projection.py recipe)
        pq, pr, sgn = [], [], []
        for v in sq:
            n = np.array(cl.unit_principal_normal(v))
            off = rng.normal(0, 2.0)
            X, Y = cl.Gx(v) + off * n[0], cl.Gy(v) + off * n[1]
            lo, hi = v - 2, v + 2
            sp, dist = cl.projection_local(X, Y, bounds=(lo, hi))
            pq.append([X, Y, lo, hi])
            pr.append([sp, dist])
            sgn.append(cl.error_sign(X, Y, sp))
        # lanes: evaluate both lane splines (note the reference's left/right swap)
        sl = rng.uniform(0, cl.right_lane.length, 16)
        lane_ev = np.array([[cl.right_lane.Gx(v), cl.right_lane.Gy(v), cl.left_lane.Gx(v % cl.left_lane.length),
                             cl.left_lane.Gy(v % cl.left_lane.length)] for v in sl])
        out.update({f"{t}/length": np.array([L, cl.right_lane.length, cl.left_lane.length]),
                    f"{t}/s": s, f"{t}/eval": ev, f"{t}/geo": geo, f"{t}/sq": sq, f"{t}/coeffs": cxy,
                    f"{t}/lookup_error": lerr, f"{t}/proj_q": np.array(pq), f"{t}/proj_r": np.array(pr),
                    f"{t}/error_sign": np.array(sgn), f"{t}/lane_s": sl, f"{t}/lane_eval": lane_ev,
                    f"{t}/knots": cl.spline_x.t})
    # TrackSegments on the two shortest tracks (9.5 s on Shanghai) + GA call signature values
    for t in ("t1_triple", "t4"):
        cl = ParameterizedCenterline(t)
        seg = TrackSegments(cl, 5, 30, 1500, 1500, 250)
        out[f"{t}/segments"] = np.array([seg.lap_time] + list(seg.bounds))
    np.savez_compressed(os.path.join(outdir, "splines_golden.npz"), **out)
