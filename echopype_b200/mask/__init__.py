from .api import apply_mask, frequency_differencing  # noqa: F401
