"""Mirror of splines/ParameterizedLane.py: a lane-boundary spline that remembers the last
projection for warm-started bounds."""
import numpy as np

from .ParameterizedLine import ParameterizedLine


class ParameterizedLane(ParameterizedLine):
    def __init__(self, device=0):
        super().__init__(device)
        self.last_progress = None

    def progress_bounds(self, step=0.5):
        if self.last_progress is None:
            return None
        return (self.last_progress - (2 * step), self.last_progress + (step * 6))

    def projection(self, X, Y, bounds=None):
        ret = super().projection(X, Y, bounds)
        self.last_progress = ret[0]
        return ret

    def from_points(self, xy):
        self.from_waypoints([(float(a), float(b)) for a, b in np.asarray(xy)])

    def from_file(self, fp):
        """CSV with columns x,y (lanes/<track>_{left,right}.csv)."""
        data = np.genfromtxt(fp, delimiter=",", names=True)
        self.from_waypoints(list(zip(data["x"], data["y"])))
