"""Import alias for the ``grid-fed-rl-gym_b200/`` source directory.

The product directory carries the reference's repository name, which is not a
valid Python identifier; this shim points the importable name
``grid_fed_rl_b200`` at it and re-exports its public API.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "grid-fed-rl-gym_b200")
__path__.insert(0, _SRC)  # submodules resolve inside the product directory

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
