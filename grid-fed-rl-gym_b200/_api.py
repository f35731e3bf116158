"""Public names of the package (what ``import grid_fed_rl_b200`` exposes)."""
from .components import Box, Bus, FeederParameters, Line, Load, PowerFlowSolution
from .feeders import (BaseFeeder, CustomFeeder, IEEE13Bus, IEEE34Bus, IEEE123Bus, NetworkConfig,
                      ScalableFeeder, SimpleRadialFeeder, SyntheticFeeder)
from .topology import (FeederSoA, RepairedFeeder, TopologyError, compile_feeder,
                       repair_topology)

__all__ = [
    "FeederSoA", "RepairedFeeder", "TopologyError", "compile_feeder", "repair_topology",
    "Box", "Bus", "FeederParameters", "Line", "Load", "PowerFlowSolution",
    "BaseFeeder", "CustomFeeder", "IEEE13Bus", "IEEE34Bus", "IEEE123Bus", "NetworkConfig",
    "ScalableFeeder", "SimpleRadialFeeder", "SyntheticFeeder",
]
