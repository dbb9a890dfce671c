"""Vehicle state container with the reference's field names and semantics
(reference: models/State.py:5-36)."""
from dataclasses import dataclass
from typing import Optional


@dataclass(eq=False)
class State:
    x: float
    y: float
    yaw: float
    v_x: float
    v_y: float
    yaw_dot: float
    steer: Optional[float] = None       # last applied commands; None = unknown
    throttle: Optional[float] = None

    def set_controls(self, throttle, steer):
        self.throttle, self.steer = throttle, steer
        return self

    def as_vector(self):
        return [self.x, self.y, self.yaw, self.v_x, self.v_y, self.yaw_dot]

    def __eq__(self, other):
        if not isinstance(other, State):
            return False
        mine = self.as_vector() + [self.steer, self.throttle]
        theirs = other.as_vector() + [other.steer, other.throttle]
        return all(a == b for a, b in zip(mine, theirs))
