from . import _lebedev  # noqa: F401
