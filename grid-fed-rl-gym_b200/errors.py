"""Exception names at the Python boundary.

The first four keep the names (and hierarchy) of the reference's
``grid_fed_rl/utils/exceptions.py:48-75`` so that callers' ``except`` clauses keep working.
Per-instance numerical trouble (non-convergence, NaN) is never an exception here: it is
data (``converged`` / ``error`` flags), as in the reference's ``step`` (grid_env.py:440-467).
"""


class GridEnvironmentError(Exception):
    """Base class (reference utils/exceptions.py:48)."""


class PowerFlowError(GridEnvironmentError):
    """The solver could not be set up or run (reference utils/exceptions.py:53)."""


class NetworkTopologyError(GridEnvironmentError):
    """Feeder cannot be compiled for the radial solvers (reference utils/exceptions.py:63)."""


class InvalidActionError(GridEnvironmentError):
    """Action batch of the wrong shape / dtype / device (reference utils/exceptions.py:68)."""


class InvalidConfigurationError(GridEnvironmentError, ValueError):
    """GFR_E_ARG from the native library."""


class GridLimitError(GridEnvironmentError):
    """GFR_E_LIMIT: feeder too large for the compiled kernels' shared-memory plan."""


class NativeRuntimeError(PowerFlowError):
    """GFR_E_CUDA: CUDA runtime failure, or no device (there is no CPU fallback)."""
