"""Drop-in module name of the reference (`from Phase import Phase`, Phase.py:6)."""
from .velocity import Phase  # noqa: F401
