"""Import alias for the ``grid-fed-rl-gym_b200/`` source directory.

The product directory carries the reference's repository name, which is not a
valid Python identifier; this shim points the importable name
``grid_fed_rl_b200`` at it and re-exports its public API.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "grid-fed-rl-gym_b200")
__path__.insert(0, _SRC)  # submodules resolve inside the product directory

from . import _api  # noqa: E402
from ._api import __all__  # noqa: E402,F401

for _n in __all__:
    if _n in _api.__dict__:
        globals()[_n] = _api.__dict__[_n]


def __getattr__(name):
    return _api.__getattr__(name)
