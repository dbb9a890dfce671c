"""GPU-backed mirror of splines/ParameterizedLine.py.

Same method names and return conventions; every method also accepts arrays (that is the
point: one kernel launch per batch of queries).  Scalars in -> Python floats out, like the
reference.  Table construction is a host precompute; all evaluation goes through libmpcr.so
(no CPU evaluation path).
"""
import ctypes as C

import numpy as np

from .. import _capi
from . import tables

_DEFAULT_HANDLE = {}


def _handle(device=0):
    """A solver handle is the device context of the track kernels."""
    h = _DEFAULT_HANDLE.get(device)
    if h is None:
        from ..solver import BatchedMPC
        h = BatchedMPC(device=device, on_device=False)
        _DEFAULT_HANDLE[device] = h
    return h


class ParameterizedLine:
    def __init__(self, device=0):
        self.length = None
        self.waypoints = None
        self._track = None
        self._device = device
        self._tables = None
        self._err = None

    # ---- construction -----------------------------------------------------------------
    def from_waypoints(self, waypoints):
        """splines/ParameterizedLine.py:162-178."""
        t = tables.line_tables(waypoints)
        self._tables = t
        self.length = t["length"]
        self.waypoints = [(float(p[0]), float(p[1])) for p in waypoints]
        self._upload()

    def from_file(self, fp):
        raise NotImplementedError()

    def _upload(self):
        self._free()
        owner = _handle(self._device)
        self._owner = owner
        t = self._tables
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        er = el = None
        n_err = 0
        if self._err is not None:
            er = np.ascontiguousarray(self._err[:, 1], dtype=np.float64)
            el = np.ascontiguousarray(self._err[:, 2], dtype=np.float64)
            n_err = len(er)
        out = C.c_void_p()
        rc = owner.lib.mpcr_track_create(owner._h, len(t["brk"]) - 1, p(t["brk"]), p(t["cx"]), p(t["cy"]),
                                         n_err, p(er), p(el), C.byref(out))
        if rc != 0:
            raise _capi.MpcrError(owner.lib.mpcr_last_error(owner._h).decode())
        self._track = out

    def device_track(self, owner):
        """A second mpcr_track bound to ``owner`` (a BatchedMPC created with on_device=True):
        query arrays are then device pointers.  Returns the raw handle; the caller destroys it
        with ``owner.lib.mpcr_track_destroy``."""
        t = self._tables
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        er = el = None
        n_err = 0
        if self._err is not None:
            er = np.ascontiguousarray(self._err[:, 1], dtype=np.float64)
            el = np.ascontiguousarray(self._err[:, 2], dtype=np.float64)
            n_err = len(er)
        out = C.c_void_p()
        rc = owner.lib.mpcr_track_create(owner._h, len(t["brk"]) - 1, p(t["brk"]), p(t["cx"]), p(t["cy"]),
                                         n_err, p(er), p(el), C.byref(out))
        if rc != 0:
            raise _capi.MpcrError(owner.lib.mpcr_last_error(owner._h).decode())
        return out

    def _free(self):
        if getattr(self, "_track", None):
            self._owner.lib.mpcr_track_destroy(self._track)
            self._track = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # ---- batched primitives -----------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise _capi.MpcrError(self._owner.lib.mpcr_last_error(self._owner._h).decode())

    def eval_all(self, s):
        """[B,6] = (Gx, Gy, dGx, dGy, ddGx, ddGy) at s[B]."""
        s = np.ascontiguousarray(np.atleast_1d(s), dtype=np.float64)
        out = np.empty((len(s), 6))
        self._chk(self._owner.lib.mpcr_track_eval(self._track, len(s), s.ctypes.data, out.ctypes.data, None))
        return out

    def coeffs(self, s, lookahead):
        """(cx[B,5], cy[B,5]) = x/y_as_coeffs(s, lookahead, deg=4) for s[B]."""
        s = np.ascontiguousarray(np.atleast_1d(s), dtype=np.float64)
        cx = np.empty((len(s), 5))
        cy = np.empty((len(s), 5))
        self._chk(self._owner.lib.mpcr_track_coeffs(self._track, len(s), s.ctypes.data, float(lookahead),
                                                    cx.ctypes.data, cy.ctypes.data, None))
        return cx, cy

    def project(self, X, Y, lo=None, hi=None):
        """(s*, dist, sign) for points (X[B], Y[B]); lo/hi None or NaN = global search."""
        X = np.ascontiguousarray(np.atleast_1d(X), dtype=np.float64)
        Y = np.ascontiguousarray(np.atleast_1d(Y), dtype=np.float64)
        B = len(X)
        lo = np.full(B, np.nan) if lo is None else np.ascontiguousarray(np.broadcast_to(lo, B), dtype=np.float64)
        hi = np.full(B, np.nan) if hi is None else np.ascontiguousarray(np.broadcast_to(hi, B), dtype=np.float64)
        s = np.empty(B)
        d = np.empty(B)
        sg = np.empty(B, dtype=np.int32)
        self._chk(self._owner.lib.mpcr_track_project(self._track, B, X.ctypes.data, Y.ctypes.data,
                                                     lo.ctypes.data, hi.ctypes.data, s.ctypes.data,
                                                     d.ctypes.data, sg.ctypes.data, None))
        return s, d, sg

    # ---- the reference's method surface -----------------------------------------------
    @staticmethod
    def _ret(v, like):
        return float(v[0]) if np.ndim(like) == 0 else v

    def Gx(self, s): return self._ret(self.eval_all(s)[:, 0], s)
    def Gy(self, s): return self._ret(self.eval_all(s)[:, 1], s)
    def dGx(self, s): return self._ret(self.eval_all(s)[:, 2], s)
    def dGy(self, s): return self._ret(self.eval_all(s)[:, 3], s)
    def ddGx(self, s): return self._ret(self.eval_all(s)[:, 4], s)
    def ddGy(self, s): return self._ret(self.eval_all(s)[:, 5], s)

    def x_as_coeffs(self, s, lookahead, deg=4):
        assert deg == 4, "the device kernel fits degree-4 polynomials (the only degree the MPC uses)"
        return list(self.coeffs(s, lookahead)[0][0])

    def y_as_coeffs(self, s, lookahead, deg=4):
        assert deg == 4
        return list(self.coeffs(s, lookahead)[1][0])

    def projection(self, X, Y, bounds=None):
        if bounds is None or 5 < abs(bounds[1] - bounds[0]):
            return self.projection_global(X, Y)
        return self.projection_local(X, Y, bounds=bounds)

    def projection_local(self, X, Y, bounds=None, warn=True):
        if bounds is None:
            bounds = (0, self.length)
        if warn is True and (bounds[1] - bounds[0]) > 10:
            print("Warning: local projection over with large bounds", bounds)
        # wide windows go to the kernel's global path; the reference would run Brent on them
        s, d, _ = self.project(X, Y, bounds[0], bounds[1])
        return float(s[0]), float(d[0])

    def projection_global(self, X, Y):
        s, d, _ = self.project(X, Y)
        return float(s[0]), float(d[0])

    def unit_tangent(self, s):
        e = self.eval_all(s)
        dr = e[:, 2:4]
        ut = dr / np.linalg.norm(dr, axis=1, keepdims=True)
        return ut[0] if np.ndim(s) == 0 else ut

    def unit_tangent_yaw(self, s):
        ut = np.atleast_2d(self.unit_tangent(s))
        return self._ret(np.arctan2(ut[:, 1], ut[:, 0]), s)

    def curvature(self, s):
        e = self.eval_all(s)
        return self._ret(np.abs(e[:, 2] * e[:, 5] - e[:, 3] * e[:, 4]), s)

    def mean_curvature(self, s, lookahead, N=10):
        ss = np.linspace(s, s + lookahead, N)
        return (1 / N) * float(np.sum(self.curvature(ss)))

    def unit_principal_normal(self, s):
        ut = np.atleast_2d(self.unit_tangent(s))
        if np.ndim(s) == 0:
            return float(ut[0, 1]), float(-ut[0, 0])
        return np.stack([ut[:, 1], -ut[:, 0]], axis=1)
