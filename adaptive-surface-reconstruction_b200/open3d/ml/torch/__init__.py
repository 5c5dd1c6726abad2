from . import layers, ops  # noqa: F401
