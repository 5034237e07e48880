from . import ray, render  # noqa: F401
