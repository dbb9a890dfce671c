"""Track geometry with the reference's class names (splines/*.py); evaluation runs on the GPU
through the mpcr_track_* entry points of libmpcr.so."""
from .ParameterizedLine import ParameterizedLine  # noqa: F401
from .ParameterizedLane import ParameterizedLane  # noqa: F401
from .ParameterizedCenterline import ParameterizedCenterline  # noqa: F401
