"""Public names of the package (what ``import grid_fed_rl_b200`` exposes)."""
from .components import Box, Bus, FeederParameters, Line, Load, PowerFlowSolution
from .errors import (GridEnvironmentError, GridLimitError, InvalidActionError,
                     InvalidConfigurationError, NativeRuntimeError, NetworkTopologyError,
                     PowerFlowError)
from .feeders import (BaseFeeder, CustomFeeder, IEEE13Bus, IEEE34Bus, IEEE123Bus, NetworkConfig,
                      ScalableFeeder, SimpleRadialFeeder, SyntheticFeeder)
from .topology import (FeederSoA, RepairedFeeder, TopologyError, auto_lanes, compile_feeder,
                       compile_for_solver, repair_topology)

__all__ = [
    "FeederSoA", "RepairedFeeder", "TopologyError", "auto_lanes", "compile_feeder",
    "compile_for_solver", "repair_topology",
    "Box", "Bus", "FeederParameters", "Line", "Load", "PowerFlowSolution",
    "BaseFeeder", "CustomFeeder", "IEEE13Bus", "IEEE34Bus", "IEEE123Bus", "NetworkConfig",
    "ScalableFeeder", "SimpleRadialFeeder", "SyntheticFeeder",
    "GridEnvironmentError", "GridLimitError", "InvalidActionError", "InvalidConfigurationError",
    "NativeRuntimeError", "NetworkTopologyError", "PowerFlowError",
    "BatchedGridEnvironment", "B200PowerFlowSolver", "shard_range",
    "GridEnvironment", "VectorizedEnvironment", "RolloutBuffer", "collect_random_data", "GraphedCollector", "HostStepper",
    "MultiAgentEnvironmentWrapper", "AgentConfig",
]


def __getattr__(name):
    # the torch-backed classes load lazily so that feeder / topology tooling imports stay light
    if name == "BatchedGridEnvironment":
        from . import env
        return getattr(env, name)
    if name == "shard_range":
        from .distributed import shard_range
        return shard_range
    if name == "B200PowerFlowSolver":
        from .solver import B200PowerFlowSolver
        return B200PowerFlowSolver
    if name in ("GridEnvironment", "VectorizedEnvironment", "RolloutBuffer", "collect_random_data", "GraphedCollector"):
        from . import compat
        return getattr(compat, name)
    if name in ("MultiAgentEnvironmentWrapper", "AgentConfig"):
        from . import multi_agent
        return getattr(multi_agent, name)
    if name == "HostStepper":
        from .pipeline import HostStepper
        return HostStepper
    raise AttributeError(name)
