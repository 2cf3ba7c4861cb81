"""`parallax.log` — logger named PARALLAX, level from PARALLAX_LOG_LEVEL.

Parity: reference `common/lib.py:58-67`.
"""
import logging
import os
import sys

from .consts import PARALLAX_LOG_LEVEL

_LEVELS = {
    "DEBUG": logging.DEBUG, "INFO": logging.INFO, "WARNING": logging.WARNING,
    "WARN": logging.WARNING, "ERROR": logging.ERROR, "CRITICAL": logging.CRITICAL,
}


def _build():
    logger = logging.getLogger("PARALLAX")
    level = _LEVELS.get(os.environ.get(PARALLAX_LOG_LEVEL, "INFO").upper(), logging.INFO)
    logger.setLevel(level)
    if not logger.handlers:
        handler = logging.StreamHandler(sys.stderr)
        handler.setFormatter(logging.Formatter(
            "%(asctime)s %(name)s %(levelname)s %(message)s", "%H:%M:%S"))
        logger.addHandler(handler)
        logger.propagate = False
    return logger


parallax_log = _build()
