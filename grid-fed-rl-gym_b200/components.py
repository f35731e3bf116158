"""Plain records for the feeder inputs of the batched power-flow step.

These carry exactly the attributes the hot path reads from the reference's
component classes, so a reference feeder and one of ours are interchangeable
wherever this package takes a ``feeder`` (duck-typed on attribute names):

* ``Bus``   - reference ``grid_fed_rl/environments/base.py:197-227``
* ``Line``  - reference ``grid_fed_rl/environments/base.py:230-264``
* ``Load``  - reference ``grid_fed_rl/environments/base.py:267-295``
* ``PowerFlowSolution`` - reference ``grid_fed_rl/environments/power_flow.py:12-22``
* ``Box``   - reference ``grid_fed_rl/environments/base.py:37-63`` (shape/low/high/sample only)
"""

from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Any, Union

import numpy as np

BusId = Union[int, str]


class Bus:
    __slots__ = ("id", "voltage_level", "bus_type", "base_voltage",
                 "voltage_magnitude", "voltage_angle")

    def __init__(self, id: BusId, voltage_level: float, bus_type: str = "pq",
                 base_voltage: float = 1.0) -> None:
        self.id = id
        self.voltage_level = voltage_level
        self.bus_type = bus_type          # "slack" | "pv" | "pq"
        self.base_voltage = base_voltage
        self.voltage_magnitude = 1.0
        self.voltage_angle = 0.0

    def __repr__(self) -> str:
        return f"Bus({self.id!r}, {self.bus_type})"


class Line:
    __slots__ = ("id", "from_bus", "to_bus", "resistance", "reactance", "rating",
                 "power_flow", "loading")

    def __init__(self, id: BusId, from_bus: BusId, to_bus: BusId, resistance: float,
                 reactance: float, rating: float) -> None:
        self.id = id
        self.from_bus = from_bus
        self.to_bus = to_bus
        self.resistance = resistance      # pu
        self.reactance = reactance        # pu
        self.rating = rating              # VA
        self.power_flow = 0.0
        self.loading = 0.0

    def update_state(self, power_flow: float = None) -> None:
        # reference base.py:261-264: loading is |P| / rating, not |S| / rating
        if power_flow is not None:
            self.power_flow = power_flow
            self.loading = abs(power_flow) / self.rating if self.rating > 0 else 0.0

    def __repr__(self) -> str:
        return f"Line({self.id!r}, {self.from_bus!r}->{self.to_bus!r})"


class Load:
    __slots__ = ("id", "bus", "base_power", "power_factor", "active_power", "reactive_power")

    def __init__(self, id: BusId, bus: BusId, base_power: float, power_factor: float = 0.95) -> None:
        self.id = id
        self.bus = bus
        self.base_power = base_power      # W
        self.power_factor = power_factor
        # The reference never refreshes these two after construction
        # (base.py:282-283); the observation and the frequency model read them
        # as constants.
        self.active_power = base_power
        self.reactive_power = base_power * np.tan(np.arccos(power_factor))

    def __repr__(self) -> str:
        return f"Load({self.id!r}@{self.bus!r}, {self.base_power:.0f} W)"


@dataclass
class FeederParameters:
    base_voltage: float   # kV
    base_power: float     # MVA
    frequency: float      # Hz


@dataclass
class PowerFlowSolution:
    """Same field names as the reference dataclass; B=1 results hold numpy
    arrays / scalars, batched results hold tensors with a leading B axis."""
    converged: Any
    iterations: Any
    bus_voltages: Any
    bus_angles: Any
    line_flows: Any
    line_loadings: Any
    losses: Any
    max_mismatch: Any


class Box:
    """Shape/bounds holder with the reference's ``sample()`` contract."""

    def __init__(self, low, high, shape=None, dtype=None) -> None:
        self.low = np.asarray(low, dtype=np.float64).reshape(-1)
        self.high = np.asarray(high, dtype=np.float64).reshape(-1)
        self.shape = tuple(shape) if shape is not None else self.low.shape
        self.dtype = dtype or np.float32

    def _bound(self, arr: np.ndarray, i: int, default: float) -> float:
        if arr.size == 1:
            return float(arr[0])
        return float(arr[i]) if i < arr.size else default

    def sample(self):
        n = self.shape[0] if self.shape else 1
        vals = [random.uniform(self._bound(self.low, i, -1.0), self._bound(self.high, i, 1.0))
                for i in range(n)]
        return vals if n > 1 else vals[0]
