from .api import add_depth  # noqa: F401
