"""Logging helpers with the reference's behaviour (echopype/utils/log.py:1-114): per-module loggers,
stdout below WARNING / stderr from WARNING, ``verbose(override=)`` toggles ``logging.disable``."""

import logging
import sys
from typing import Optional

_FORMATTER = logging.Formatter("{asctime}:{name}:{levelname}: {message}", style="{")


class _BelowWarning(logging.Filter):
    def filter(self, record):
        return record.levelno < logging.WARNING


def _init_logger(name) -> logging.Logger:
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    if not any(getattr(h, "name", "") == "stdout_stream_handler" for h in logger.handlers):
        out = logging.StreamHandler(sys.stdout)
        out.setLevel(logging.INFO)
        out.set_name("stdout_stream_handler")
        out.setFormatter(_FORMATTER)
        out.addFilter(_BelowWarning())
        err = logging.StreamHandler(sys.stderr)
        err.setLevel(logging.WARNING)
        err.set_name("stderr_stream_handler")
        err.setFormatter(_FORMATTER)
        logger.addHandler(out)
        logger.addHandler(err)
    return logger


def verbose(logfile: Optional[str] = None, override: bool = False) -> None:
    """echopype/utils/log.py:19-60: ``override=False`` turns log output on, ``True`` silences it."""
    if not isinstance(override, bool):
        raise ValueError("override argument must be a boolean!")
    logging.disable(logging.NOTSET if override is False else logging.WARNING)
    if logfile is not None:
        pkg = __name__.split(".")[0]
        for name in list(logging.root.manager.loggerDict):
            if pkg in name:
                lg = logging.getLogger(name)
                if not any(getattr(h, "name", "") == "logfile_file_handler" for h in lg.handlers):
                    fh = logging.FileHandler(logfile)
                    fh.set_name("logfile_file_handler")
                    fh.setFormatter(_FORMATTER)
                    lg.addHandler(fh)
