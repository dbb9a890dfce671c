"""Mirror of splines/ParameterizedCenterline.py (minus ``plot``).

Track data comes from mpc_racing_b200/data/tracks/<track>.npz (converted from the
reference's pickles/CSVs by tools/convert_tracks.py) or, via ``from_file``, from a reference
waypoint pickle.
"""
import os
import pickle

import numpy as np

from .ParameterizedLane import ParameterizedLane
from .ParameterizedLine import ParameterizedLine
from .util import euclidean, midpoint

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "tracks")
TRACKS = ["shanghai_intl_circuit", "t1_triple", "t2_triple", "t3", "t4"]


class ParameterizedCenterline(ParameterizedLine):
    def __init__(self, track: str = "shanghai_intl_circuit", lanes=True, error=True, device=0):
        super().__init__(device)
        d = np.load(os.path.join(DATA, track + ".npz"))
        self.track = track
        if lanes:
            # the reference builds right_lane from *_left.csv and left_lane from *_right.csv
            # (splines/ParameterizedCenterline.py:17-21); kept for parity
            self.right_lane = ParameterizedLane(device)
            self.right_lane.from_points(d["left"])
            self.left_lane = ParameterizedLane(device)
            self.left_lane.from_points(d["right"])
        if error:
            self._err = np.ascontiguousarray(d["max_error"], dtype=np.float64)
        self._from_points([tuple(p) for p in d["waypoints"]])

    def _from_points(self, waypoints):
        # close the gap with a point 90 % of the way from the last to the first waypoint (:99-103)
        if euclidean(waypoints[-1], waypoints[0]) > 0.1:
            waypoints = list(waypoints) + [midpoint(waypoints[-1], waypoints[0], alpha=0.9)]
        self.from_waypoints(waypoints)

    def from_file(self, fp):
        with open(fp, "rb") as f:
            waypoints = pickle.load(f)
        waypoints.pop(0)                      # first element is the track id
        self._from_points([(p[0], p[1]) for p in waypoints])

    def lookup_error(self, s, lookahead):
        sa = np.ascontiguousarray(np.atleast_1d(s), dtype=np.float64)
        out = np.empty(len(sa))
        self._chk(self._owner.lib.mpcr_track_lookup_error(self._track, len(sa), sa.ctypes.data,
                                                          float(lookahead), out.ctypes.data, None))
        return self._ret(out, s)

    def error_sign(self, X, Y, s):
        """+1 to the right of the centerline.  Evaluated with the projection kernel's sign output
        on a zero-width window at s."""
        sa = np.atleast_1d(np.asarray(s, dtype=np.float64))
        _, _, sg = self.project(X, Y, sa, sa)
        return int(sg[0]) if np.ndim(s) == 0 else sg
