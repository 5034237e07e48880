from . import ray, render, turntable  # noqa: F401
