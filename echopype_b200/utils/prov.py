"""Provenance / processing-level attributes (echopype/utils/prov.py:24-43,181-331), metadata only."""

import datetime
import functools
import re

import numpy as np

from ..dataset import Dataset
from .log import _init_logger

logger = _init_logger(__name__)
VERSION = "0.1.0+b200"

PROCESSING_LEVELS = dict(
    L0="Level 0", L1A="Level 1A", L1B="Level 1B", L2A="Level 2A", L2B="Level 2B", L3A="Level 3A", L3B="Level 3B", L4="Level 4"
)
_URL = "https://echopype.readthedocs.io/en/stable/processing-levels.html"


def echopype_prov_attrs(process_type):
    now = datetime.datetime.now(datetime.timezone.utc).isoformat(timespec="seconds")
    return {
        f"{process_type}_software_name": "echopype_b200",
        f"{process_type}_software_version": VERSION,
        f"{process_type}_time": now,
    }


def source_files_vars(source_paths):
    """prov.py:85-152 reduced to the variables attached by compute_Sv (calibrate/api.py:236-241)."""
    if source_paths is None:
        paths = []
    elif isinstance(source_paths, (str, bytes)) or not hasattr(source_paths, "__iter__"):
        paths = [str(source_paths)]
    else:
        paths = [str(p) for p in source_paths]
    return {
        "source_files_var": {
            "source_filenames": (("filenames",), np.array(paths, dtype=object), {"long_name": "Source filenames"}),
        },
        "source_files_coord": {
            "filenames": (("filenames",), np.arange(len(paths)), {"long_name": "Index for data and metadata source filenames"}),
        },
    }


def _check_valid_latlon(ds):
    for name in ("longitude", "latitude"):
        if name not in ds or bool(np.all(np.isnan(np.asarray(ds[name].values, dtype=float)))):
            return False
    return True


def add_processing_level(processing_level_code, is_echodata=False):
    """prov.py:181-308 for stand-alone functions returning a Dataset."""

    def wrapper(func):
        if not (processing_level_code in PROCESSING_LEVELS or re.fullmatch(r"L\*[A|B]|L[1-4]\*", processing_level_code)):
            raise ValueError(f"Decorator processing_level_code {processing_level_code} used in {func.__qualname__} is invalid.")

        @functools.wraps(func)
        def inner(*args, **kwargs):
            ds = func(*args, **kwargs)
            if not isinstance(ds, Dataset):
                raise RuntimeError(
                    f"{func.__qualname__}: Processing level decorator function cannot be used "
                    "with a function that does not return an xarray Dataset or EchoData object"
                )
            if _check_valid_latlon(ds):
                if processing_level_code in PROCESSING_LEVELS:
                    level = PROCESSING_LEVELS[processing_level_code]
                elif "*" in processing_level_code and "input_processing_level" in ds.attrs:
                    if processing_level_code[-1] == "*":
                        sub, lev = ds.attrs["input_processing_level"][-1], processing_level_code[1]
                    else:
                        sub, lev = processing_level_code[-1], ds.attrs["input_processing_level"][-2]
                    level = PROCESSING_LEVELS[f"L{lev}{sub}"]
                    del ds.attrs["input_processing_level"]
                else:
                    raise RuntimeError(
                        "Processing level attributes (processing_level_code {processing_level_code}) "
                        f"cannot be added. Please ensure that {func.__qualname__} "
                        "uses the function insert_input_processing_level."
                    )
                ds = ds.assign_attrs({"processing_level": level, "processing_level_url": _URL})
            else:
                logger.info("xarray Dataset does not contain valid location data. Processing level attributes will not be added.")
                ds.attrs.pop("input_processing_level", None)
            return ds

        return inner

    return wrapper


def insert_input_processing_level(ds, input_ds):
    if "processing_level" in input_ds.attrs:
        return ds.assign_attrs({"input_processing_level": input_ds.attrs["processing_level"]})
    return ds
