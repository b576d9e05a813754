"""Drop-in module name of the reference (`from RunningAverage import RunningAverage`)."""
from .velocity import RunningAverage  # noqa: F401
