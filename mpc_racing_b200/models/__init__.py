from .State import State  # noqa: F401
from .VehicleParameters import VehicleParameters  # noqa: F401
