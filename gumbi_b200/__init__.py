"""gumbi_b200 -- B200-native exact-GP inference core behind Gumbi's Regressor backend API."""
from ._lib import BackendUnavailable, lib_path  # noqa: F401
from .engine import GPEngine  # noqa: F401
from .backend import ArrayGP, ArrayRegressor, B200Backend, make_backend  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "B200GP":
        from . import backend

        return backend.B200GP
    raise AttributeError(name)
