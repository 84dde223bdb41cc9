"""echopype_b200 - B200-native (sm_100a) implementation of echopype's calibrate -> clean -> commongrid
array-compute path behind the reference's Python API."""

from . import calibrate, clean, commongrid, consolidate, mask, pipeline, utils  # noqa: F401
from .dataset import DataArray, Dataset, EchoData  # noqa: F401

__version__ = "0.1.0"
