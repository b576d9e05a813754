"""Drop-in module name of the reference (`from VelocityTracker import VelocityTracker`)."""
from .velocity import VelocityTracker, Phase, RunningAverage  # noqa: F401
