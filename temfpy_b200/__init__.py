"""temfpy_b200 -- B200-native implementation of TeMFpy's mean-field -> MPS conversion hot path.

Same public entry points as ``temfpy`` (reference src/temfpy/__init__.py:21-50): the sub-modules
``slater``, ``schmidt_utils``, ``utils``, ``testing``, ``iMPS``, ``gutzwiller``, ``pfaffian`` are
loaded lazily; ``setup_logging`` mirrors __init__.py:12-15.
"""
import importlib
import logging

__version__ = "0.1.0"
_SUBMODULES = ("slater", "pfaffian", "gutzwiller", "iMPS", "schmidt_utils", "utils", "testing", "engine",
               "mps", "dist")
__all__ = list(_SUBMODULES) + ["setup_logging"]


def setup_logging(level="INFO"):
    """Basic logging configuration of the package loggers (reference __init__.py:12-15)."""
    logging.basicConfig(level=level, format="%(levelname)-8s: %(message)s")


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
